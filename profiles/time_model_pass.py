"""One trimmed model pass (35 images @336, block 8 / head 9) in the current process environment: time per pass and the
GradCAM's distance from an fp64 autograd pass (4 images).  Run plain, and under
    LD_PRELOAD=/usr/local/cuda/lib64/libcublasLt.so.12:/usr/local/cuda/lib64/libcublas.so.12 CUBLAS_EMULATE_SINGLE_PRECISION=1
to see what cuBLAS 12.9's BF16x9 FP32 emulation (absent from the cuBLAS 12.8 that torch bundles) does to torch's own sgemm calls."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pnp_ovss_b200.blip_itm import BlipITM

dev = torch.device("cuda:0")
w = bench.make_workload(0)
torch.manual_seed(4321)
model = BlipITM(img_size=336, tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
imgs, caps = w["imgs"].to(dev), w["captions"]
tok = w["tokens"].to(dev)
for _ in range(2):
    model.gradcam(imgs, caps, tok, layer=7, head=9)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(3):
    cam, _ = model.gradcam(imgs, caps, tok, layer=7, head=9)
t1.record()
torch.cuda.synchronize()
print("cublas version (runtime): %s   env: emulate=%s strategy=%s" % (
    torch.backends.cuda.cublas_version() if hasattr(torch.backends.cuda, "cublas_version") else "?",
    os.environ.get("CUBLAS_EMULATE_SINGLE_PRECISION"), os.environ.get("CUBLAS_EMULATION_STRATEGY")))
print("model pass: %.1f ms" % (t0.elapsed_time(t1) / 3))
n = 4
tok4 = w["tok"](caps[:n], padding="max_length", max_length=500).to(dev)
truth = bench.gradcam_fp64(model, imgs[:n].contiguous(), caps[:n], tok4, 7, 9, 21)
got, _ = model.gradcam(imgs[:n].contiguous(), caps[:n], tok4, layer=7, head=9)
print("GradCAM vs fp64: max |diff| / max = %.3e" % ((got.double() - truth).abs().max() / truth.abs().max()).item())
