"""Profiling driver: trimmed model passes of cfg1 (35 images @336) in the shipped GEMM mode, eagerly (no CUDA graphs), so that ncu
sees every launch.  Usage: python profiles/run_model_pass.py [passes] [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pnp_ovss_b200.blip_itm import BlipITM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = sys.argv[2] if len(sys.argv) > 2 else "3xfp16"
dev = torch.device("cuda:0")
w = bench.make_workload(0)
torch.manual_seed(4321)
model = BlipITM(img_size=w["S"], tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
model.gemm_precision = mode
model.USE_VIT_GRAPH = model.USE_TEXT_GRAPH = False
imgs, caps, tok = w["imgs"].to(dev), w["captions"], w["tokens"].to(dev)
for _ in range(n):
    cam, _ = model.gradcam(imgs, caps, tok, layer=7, head=9)
torch.cuda.synchronize()
print("ok", float(cam.abs().max()))
