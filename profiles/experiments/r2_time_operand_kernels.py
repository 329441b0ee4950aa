"""Time the model-pass operand kernels at the ViT-L block shapes of configs[1] (35 x 442 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pnp_ovss_b200 import ops
dev = torch.device("cuda:0")
M = 35 * 442
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
f = torch.randn(M, 4096, device=dev) * 2048; bias = torch.randn(4096, device=dev)
x = torch.randn(M, 1024, device=dev); r = torch.randn(M, 1024, device=dev) * 2048; g = torch.ones(1024, device=dev); bt = torch.zeros(1024, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
us = t(lambda: ops.gelu_fp16_split3(f, bias, 1 / 2048, 1.0, flag))
print("gelu_fp16_split3 [%d, 4096]: %.1f us  (%.0f GB/s of 4 + 6 B per element)" % (M, us, M * 4096 * 10 / us / 1e3))
us = t(lambda: ops.layernorm_fp16_split3(x, g, bt, 1e-6, residual=r, residual_scale=1 / 2048, residual_bias=bt, hi_scale=1.0, flag=flag, bias_one=64.0))
print("layernorm_fp16_split3 [%d, 1024] with residual: %.1f us  (%.0f GB/s of 4 + 4 + 4 + 6 B per element)" % (M, us, M * 1024 * 18 / us / 1e3))
us = t(lambda: ops.fp16_split3(x, 1.0, 1.0, flag))
print("fp16_split3 [%d, 1024]: %.1f us" % (M, us))
xs = torch.randn(875, 768, device=dev)
us = t(lambda: ops.tf32_split3(xs))
print("tf32_split3 [875, 768]: %.1f us" % us)
