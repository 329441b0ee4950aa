"""Tuning run: time the CRF kernel classes for one PNP_GRID_MULT (CTAs per SM of the grid-stride kernels)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pnp_ovss_b200 import _lib, ops, pipeline, synthetic as synth
dev = torch.device("cuda:0")
B, C, P, H, W, n = 35, 20, 21, 336, 336, 21
maps = torch.stack([synth.saliency_maps(100 + b, C, P) for b in range(B)]).to(dev)
guides = torch.from_numpy(np.stack([synth.guide_image(5000 + b, H, W) for b in range(B)])).to(dev)
gts = torch.from_numpy(np.stack([synth.gt_labels(7000 + b, H, W, n) for b in range(B)])).to(dev)
luts = torch.arange(C + 1, dtype=torch.int32, device=dev).repeat(B, 1)
lat_b = ops.build_lattice(H, W, 50.0, rgb=guides, srgb=5.0)
lib = _lib.load()
tot = (ctypes.c_float * 19)(); cnt = (ctypes.c_int * 19)()
for it in range(3):
    if it == 1:
        lib.pnp_profile_start(ctypes.c_uint(0xFFFFFFFE))
    hist = torch.zeros((n, n), dtype=torch.int64, device=dev)
    pipeline.postprocess_batch(maps, guides, gts, luts, hist, threshold=0.15, rescale=False, with_background=True, mode="blur+crf",
                               n_class=n, bilateral=lat_b)
torch.cuda.synchronize()
lib.pnp_profile_stop(tot, cnt, 19)
print({k: v for k, v in os.environ.items() if k.startswith("PNP_GRID")}, {lib.pnp_profile_kernel_name(i).decode(): round(tot[i] / 2, 2) for i in (12, 13, 14, 17, 18)})
