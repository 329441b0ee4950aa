"""Profiling driver: the post-processing half of one bench step (configs[1] shapes: B=35, C'=21, 336x336, blur+crf)
without the torch model, so that ncu captures only the custom kernels.  Usage: python profiles/run_postprocess.py [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np
import torch

from pnp_ovss_b200 import synthetic as synth
from pnp_ovss_b200 import ops, pipeline

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kind = sys.argv[2] if len(sys.argv) > 2 else "natural"
dev = torch.device("cuda:0")
B, C, P, H, W, n = 35, 20, 21, 336, 336, 21
maps = torch.stack([synth.saliency_maps(100 + b, C, P) for b in range(B)]).to(dev)
guides = torch.from_numpy(np.stack([synth.guide_image(5000 + b, H, W, kind) for b in range(B)])).to(dev)
gts = torch.from_numpy(np.stack([synth.gt_labels(7000 + b, H, W, n) for b in range(B)])).to(dev)
luts = torch.arange(C + 1, dtype=torch.int32, device=dev).repeat(B, 1)
lat_b = ops.build_lattice(H, W, 50.0, rgb=guides, srgb=5.0)
for it in range(iters):
    hist = torch.zeros((n, n), dtype=torch.int64, device=dev)
    stats = {}
    pipeline.postprocess_batch(maps, guides, gts, luts, hist, threshold=0.15, rescale=False, with_background=True, mode="blur+crf",
                               n_class=n, stats=stats, bilateral=lat_b)
torch.cuda.synchronize()
print("ok", stats, int(hist.sum()))
