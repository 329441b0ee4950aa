import os, sys
sys.path.insert(0, "/root/repo")
import torch
from pnp_ovss_b200 import ops
dev = torch.device("cuda:0")
qkv = torch.randn(35, 442, 3, 16, 64, device=dev) * 2048.0
for _ in range(3):
    ops.attention_fp16x3(qkv, 1.0 / 2048.0, 0.125, None, split_hi_scale=1.0)
torch.cuda.synchronize()
