"""Real datasets: every image has its own ground-truth size and 1-3 classes, so a batch falls apart into per-shape
buckets.  Time of the whole post-processing (both passes) for 35 such images, bucket by bucket (what pipeline does today)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnp_ovss_b200 import ops, pipeline, synthetic as synth
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
B, P, n = 35, 21, 21
shapes = [(int(rng.integers(280, 500)), int(rng.integers(300, 500))) for _ in range(B)]   # VOC-like sizes
counts = rng.choice([1, 2, 3, 4], size=B, p=[0.68, 0.25, 0.06, 0.01])
items = []
for b in range(B):
    H, W = shapes[b]
    C = int(counts[b])
    items.append((synth.saliency_maps(100 + b, C, P).unsqueeze(0).to(dev),
                  torch.from_numpy(synth.guide_image(200 + b, H, W)[None]).to(dev),
                  torch.from_numpy(synth.gt_labels(300 + b, H, W, n)[None]).to(dev),
                  torch.arange(C + 1, dtype=torch.int32, device=dev)[None]))

def run():
    hist = torch.zeros((n, n), dtype=torch.int64, device=dev)
    for maps, guide, gt, lut in items:
        H, W = gt.shape[1:]
        lat_b = ops.build_lattice(H, W, 50.0, rgb=guide, srgb=5.0)
        for rescale in (True, False):
            pipeline.postprocess_batch(maps, guide, gt, lut, hist, threshold=0.15, rescale=rescale, with_background=True,
                                       mode="blur+crf", n_class=n, bilateral=lat_b)
    return hist

run(); torch.cuda.synchronize()
t = time.perf_counter(); h = run(); torch.cuda.synchronize()
dt = time.perf_counter() - t
print("35 images, 35 distinct shapes, 1-4 classes: %.1f ms total, %.2f ms per image (both passes, lattice build included)" % (dt * 1e3, dt * 1e3 / B))
