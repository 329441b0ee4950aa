"""How far does the block-8/head-9 GradCAM move when the model's GEMMs run in TF32 or bf16 instead of fp32?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pnp_ovss_b200.blip_itm import BlipITM
dev = torch.device("cuda:0")
w = bench.make_workload(0)
torch.manual_seed(4321)
model = BlipITM(img_size=336, tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
tokens = w["tokens"].to(dev)
imgs = w["imgs"][:8].to(dev)
caps = w["captions"][:8]
tok8 = w["tok"](caps, padding="max_length", max_length=500).to(dev)
ref, _ = model.gradcam(imgs, caps, tok8, layer=7, head=9)
for mode in ("3xtf32", "tf32", "bf16"):
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    model.gemm_precision = mode
    got, _ = model.gradcam(imgs, caps, tok8, layer=7, head=9)
    torch.backends.cuda.matmul.allow_tf32 = False
    model.gemm_precision = "fp32"
    scale = ref.abs().max().item()
    print("%s: max |diff| / max|ref| = %.3e   mean rel (where ref>1e-3*max) = %.3e" % (
        mode, (got - ref).abs().max().item() / scale,
        ((got - ref).abs() / ref.abs().clamp_min(1e-30))[ref > 1e-3 * scale].mean().item()))
