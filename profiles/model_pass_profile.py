"""Where one trimmed model pass (35 images @336, block 8 / head 9) spends its time, per GEMM mode: CUDA-event time per pass,
GradCAM distance from the native-fp32 pass and from an fp64 autograd pass (4 images), and the torch.profiler kernel table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pnp_ovss_b200.blip_itm import BlipITM

dev = torch.device("cuda:0")
cfg = int(os.environ.get("PNP_CFG", "1"))
w = bench.make_workload(0) if not hasattr(bench, "CONFIGS") else bench.make_workload(0, bench.CONFIGS[cfg])
torch.manual_seed(4321)
model = BlipITM(img_size=w["S"], tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
imgs, caps = w["imgs"].to(dev), w["captions"]
tok = w["tokens"].to(dev)
cams = {}
for mode in sys.argv[1:] or ["fp32", "3xtf32"]:
    model.gemm_precision = mode
    for _ in range(2):
        model.gradcam(imgs, caps, tok, layer=7, head=9)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        cam, _ = model.gradcam(imgs, caps, tok, layer=7, head=9)
    t1.record()
    torch.cuda.synchronize()
    cams[mode] = cam
    print("== %s: model pass %.1f ms" % (mode, t0.elapsed_time(t1) / 5), flush=True)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model.gradcam(imgs, caps, tok, layer=7, head=9)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=90))
if "fp32" in cams:
    for mode, cam in cams.items():
        print("%s vs fp32: max |diff| / max = %.3e" % (mode, ((cam - cams["fp32"]).abs().max() / cams["fp32"].abs().max()).item()))
n = 4
tok4 = w["tok"](caps[:n], padding="max_length", max_length=500).to(dev)
truth = bench.gradcam_fp64(model, imgs[:n].contiguous(), caps[:n], tok4, 7, 9, w["P"])
for mode in cams:
    model.gemm_precision = mode
    got, _ = model.gradcam(imgs[:n].contiguous(), caps[:n], tok4, layer=7, head=9)
    print("%s GradCAM vs fp64: max |diff| / max = %.3e" % (mode, ((got.double() - truth).abs().max() / truth.abs().max()).item()))
