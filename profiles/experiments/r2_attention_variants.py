"""Time pnp_attention_fp16x3 at the cfg1 encoder shape (35 images, 442 tokens, 16 heads) for the fragment-load variant selected
by PNP_ATT_VARIANT (read once per process: bit 0 = K fragments via ldmatrix.x4, bit 1 = V fragments via ldmatrix.x4.trans)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pnp_ovss_b200 import ops
dev = torch.device("cuda:0")
for (B, L, H) in [(35, 442, 16), (35, 785, 16)]:
    qkv = torch.randn(B, L, 3, H, 64, device=dev) * 2048.0
    for split in (None, 1.0):
        for _ in range(3):
            ops.attention_fp16x3(qkv, 1.0 / 2048.0, 0.125, None, split_hi_scale=split)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            ops.attention_fp16x3(qkv, 1.0 / 2048.0, 0.125, None, split_hi_scale=split)
        b.record()
        torch.cuda.synchronize()
        print("variant %s  B=%d L=%d H=%d  %s output: %.3f ms per call (split pass + main kernel)" % (
            os.environ.get("PNP_ATT_VARIANT", "0"), B, L, H, "operand-split" if split else "fp32", a.elapsed_time(b) / 20))
