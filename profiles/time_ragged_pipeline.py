"""Real datasets: every image has its own ground-truth size and 1-4 classes.  Post-processing time (both reference passes, lattice
builds included) of 35 VOC-like images through pipeline.batch_confusion's bucket loop, for
  (a) round 1's form: one bucket per (class count, shape), one stream;
  (b) buckets keyed by (padded channel count, background, shape) with per-image class counts, spread over BUCKET_STREAMS streams.
Shapes: VOC's common 500x375 / 375x500 / 500x333 plus random others (`--all-distinct`: 35 different shapes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import synth
from pnp_ovss_b200 import pipeline

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
B, S, n = 35, 336, 21
distinct = "--all-distinct" in sys.argv
common = [(375, 500)] * 18 + [(500, 375)] * 7 + [(333, 500)] * 4
shapes = [(int(rng.integers(280, 500)), int(rng.integers(300, 500))) for _ in range(B)] if distinct else \
         (common + [(int(rng.integers(280, 500)), int(rng.integers(300, 500))) for _ in range(B - len(common))])
counts = rng.choice([1, 2, 3, 4], size=B, p=[0.55, 0.30, 0.10, 0.05])
names = ["aeroplane", "bicycle", "bird", "boat", "motorbike", "television"]
tok = synth.SyntheticWordPieceTokenizer()
class_lists = [names[:int(c)] for c in counts]
caps = ["A picture of " + " ".join(c) for c in class_lists]
tokens = tok(caps, padding="max_length", max_length=500)
T = max(len(tok.encode(c)) for c in caps)
P = S // 16
fn = synth.SynthGradcamFn(3, B, T, P)
rows = tokens.attention_mask[:, 1:T].float()
g_fixed = fn(torch.zeros(B, 3, S, S), rows).to(dev)
gts = [synth.gt_labels(10 + b, *shapes[b], n) for b in range(B)]
guides = [synth.guide_image(20 + b, *shapes[b]) for b in range(B)]
ids = [[1 + names.index(c) for c in cl] for cl in class_lists]
imgs = torch.zeros(B, 3, S, S, device=dev)
bad = torch.zeros(1, dtype=torch.int32, device=dev)


def run():
    return pipeline.batch_confusion(lambda x: g_fixed, imgs, tokens.input_ids.tolist(), tok.decode, class_lists, ids, gts, guides, drop_iter=1,
                                    patch_num=P, threshold=0.15, data_type="voc", mode="blur+crf", n_class=n, bad_count=bad)


def timed(label):
    run(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        h = run()[0]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print("%-70s %7.1f ms per batch, %5.2f ms per image   (hist sum %d)" % (label, dt * 1e3, dt * 1e3 / B, int(h.sum())))
    return h


print("35 images, %d distinct shapes, class counts %s" % (len(set(shapes)), np.bincount(counts)[1:].tolist()))
pipeline.PAD_CLASSES_IN_BUCKETS, pipeline.BUCKET_STREAMS = False, 1
ha = timed("(a) exact-count buckets, one stream")
pipeline.PAD_CLASSES_IN_BUCKETS, pipeline.BUCKET_STREAMS = True, 1
hb = timed("(b1) buckets by padded channel count, one stream")
for ns in (2, 4, 8):
    pipeline.BUCKET_STREAMS = ns
    hc = timed("(b%d) buckets by padded channel count, %d streams" % (ns, ns))
    assert torch.equal(ha, hc)
assert torch.equal(ha, hb)
# the uniform reference: 35 images of one shape and one class count
shapes_u = [(375, 500)] * B
gts = [synth.gt_labels(10 + b, *shapes_u[b], n) for b in range(B)]
guides = [synth.guide_image(20 + b, *shapes_u[b]) for b in range(B)]
class_lists = [names[:2]] * B
caps = ["A picture of " + " ".join(c) for c in class_lists]
tokens = tok(caps, padding="max_length", max_length=500)
ids = [[1 + names.index(c) for c in cl] for cl in class_lists]
T2 = max(len(tok.encode(c)) for c in caps)
g_fixed = synth.SynthGradcamFn(3, B, T2, P)(torch.zeros(B, 3, S, S), tokens.attention_mask[:, 1:T2].float()).to(dev)
timed("(u) uniform batch: 35 x 375x500, 2 classes")
