"""How long does the model pass take on one B200 the way the REFERENCE drives it (BITM:386-457: weights require grad,
loss.backward() through ViT-L + BERT, capture in the block) versus the trimmed pass the product uses?  fp32, B=35 @336."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import bench
from pnp_ovss_b200.blip_itm import BlipITM

dev = torch.device("cuda:0")
w = bench.make_workload(0)
torch.manual_seed(4321)
model = BlipITM(img_size=336, tokenizer=w["tok"]).to(dev).eval()
tokens = w["tokens"].to(dev)
imgs = w["imgs"].to(dev)
for full in (False, True):
    model.requires_grad_(full)
    for _ in range(2):
        model.gradcam(imgs, w["captions"], tokens, layer=7, head=9, full_backward=full)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        model.gradcam(imgs, w["captions"], tokens, layer=7, head=9, full_backward=full)
    torch.cuda.synchronize()
    print("full_backward=%s: %.1f ms per 35-image pass, peak mem %.1f GB" % (full, (time.perf_counter() - t0) / 3 * 1e3,
                                                                         torch.cuda.max_memory_allocated() / 2**30))
