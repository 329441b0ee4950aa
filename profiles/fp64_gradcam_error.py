"""Ground truth for the GEMM-precision question: the block-8/head-9 GradCAM of the BLIP ITM-large-shaped model computed in
fp64 (plain torch autograd, no custom kernel), against the product's fp32 pass and its 3xTF32 / TF32 / bf16 variants.
If fp32 and 3xTF32 sit at the same distance from the fp64 result, 3xTF32 is fp32-grade for this path."""
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
n = 4
imgs, caps = w["imgs"][:n].to(dev), w["captions"][:n]
tok = w["tok"](caps, padding="max_length", max_length=500).to(dev)
LAYER, HEAD = 7, 9


truth = bench.gradcam_fp64(model, imgs, caps, tok, LAYER, HEAD, 21)
scale = truth.abs().max().item()
big = truth > 1e-3 * scale
print("fp64 GradCAM: max %.3e, %d of %d cells above 1e-3*max" % (scale, int(big.sum()), truth.numel()))
for mode in ("fp32", "3xtf32", "tf32", "bf16"):
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    model.gemm_precision = mode
    got, _ = model.gradcam(imgs, caps, tok, layer=LAYER, head=HEAD)
    torch.backends.cuda.matmul.allow_tf32 = False
    model.gemm_precision = "fp32"
    d = (got.double() - truth).abs()
    print("%-7s vs fp64: max |diff| / max = %.3e   mean rel (cells > 1e-3*max) = %.3e" % (
        mode, d.max().item() / scale, (d / truth.abs().clamp_min(1e-300))[big].mean().item()))
