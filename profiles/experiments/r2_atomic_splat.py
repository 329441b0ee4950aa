"""Deterministic CSR-gather splat vs the pixel-order vector-atomic splat (PNP_SPLAT_ATOMIC=1, red.global.add.v4.f32) on the cfg1
post-processing shapes: time of a CRF pass, per-kernel time of the bilateral splat, run-to-run label stability and the
disagreement with the deterministic labels.  The switch is read once per process: run this script once per setting; it saves
its labels to gpurun_out/ so that the second run can compare.
    python profiles/experiments/r2_atomic_splat.py ; PNP_SPLAT_ATOMIC=1 python profiles/experiments/r2_atomic_splat.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np
import torch

from pnp_ovss_b200 import _lib, ops, pipeline
from pnp_ovss_b200 import synthetic as synth

mode = "atomic" if os.environ.get("PNP_SPLAT_ATOMIC") == "1" else "deterministic"
dev = torch.device("cuda:0")
lib = _lib.load()
B, C, P, H, W, n = 35, 20, 21, 336, 336, 21
maps = torch.stack([synth.saliency_maps(100 + b, C, P) for b in range(B)])
maps[:, :, :4, :4] = 0
maps = maps.to(dev)
guides = torch.from_numpy(np.stack([synth.guide_image(5000 + b, H, W) for b in range(B)])).to(dev)
gts = torch.from_numpy(np.stack([synth.gt_labels(7000 + b, H, W, n) for b in range(B)])).to(dev)
luts = torch.arange(C + 1, dtype=torch.int32, device=dev).repeat(B, 1)
lat_b = ops.build_lattice(H, W, 50.0, rgb=guides, srgb=5.0)
n_ids = lib.pnp_profile_num_kernels()
tot, cnt = (ctypes.c_float * n_ids)(), (ctypes.c_int * n_ids)()
labels = []
for it in range(4):
    hist = torch.zeros((n, n), dtype=torch.int64, device=dev)
    if it == 3:
        lib.pnp_profile_start(ctypes.c_uint(0xFFFFFFFE))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    pred = pipeline.postprocess_batch(maps, guides, gts, luts, hist, threshold=0.15, rescale=False, with_background=True, mode="blur+crf",
                                      n_class=n, bilateral=lat_b, return_labels=True)
    t1.record()
    torch.cuda.synchronize()
    labels.append(pred.cpu().numpy())
lib.pnp_profile_stop(tot, cnt, n_ids)
per = {lib.pnp_profile_kernel_name(i).decode(): (tot[i], cnt[i]) for i in range(1, n_ids) if cnt[i]}
print("%s splat: post-processing pass %.2f ms; bilateral splat %.3f ms per launch (%d launches), update %.3f, lattice blur %.3f" % (
    mode, t0.elapsed_time(t1), per["crf_splat_bilateral"][0] / per["crf_splat_bilateral"][1], per["crf_splat_bilateral"][1],
    per["crf_meanfield_update"][0] / per["crf_meanfield_update"][1], per["crf_blur_axis_bilateral"][0] / per["crf_blur_axis_bilateral"][1]))
print("   run-to-run: %d of %d pixels differ between run 2 and run 3" % (int((labels[1] != labels[2]).sum()), labels[1].size))
out = os.path.join(ROOT, "gpurun_out", "splat_labels_%s.npy" % mode)
np.save(out, labels[2])
other = os.path.join(ROOT, "gpurun_out", "splat_labels_%s.npy" % ("deterministic" if mode == "atomic" else "atomic"))
if os.path.exists(other):
    o = np.load(other)
    print("   vs the other setting: %d of %d pixels differ (%.2e)" % (int((o != labels[2]).sum()), o.size, float((o != labels[2]).mean())))
