"""Does the bias ride for free in the fp16 -> fp32 GEMM?  torch.mm(out_dtype=fp32) vs torch.addmm(bias, ..., out_dtype=fp32) vs
mm followed by torch.add(bias, y, alpha) at the qkv / KV shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

dev = torch.device("cuda:0")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for (M, K, N) in [(15470, 3072, 3072), (15470, 3072, 18432)]:
    A = torch.randn(M, K, device=dev).half()
    W = torch.randn(N, K, device=dev).half()
    bias = torch.randn(N, device=dev)
    t_mm = timeit(lambda: torch.mm(A, W.t(), out_dtype=torch.float32))
    t_addmm = timeit(lambda: torch.addmm(bias, A, W.t(), out_dtype=torch.float32))
    t_two = timeit(lambda: torch.add(bias, torch.mm(A, W.t(), out_dtype=torch.float32), alpha=1.0 / 2048))
    y1 = torch.addmm(bias, A, W.t(), out_dtype=torch.float32)
    y2 = torch.mm(A, W.t(), out_dtype=torch.float32) + bias
    print("M=%d K=%d N=%d: mm %.3f ms, addmm(bias) %.3f ms, mm + add %.3f ms; addmm == mm + bias: max diff %.2e" % (
        M, K, N, t_mm, t_addmm, t_two, (y1 - y2).abs().max().item()))
q = torch.randn(35, 16, 442, 64, device=dev)
for sc in (1.0, 2048.0):
    t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q * sc, q * sc, q * sc, scale=0.125 / sc / sc))
    print("sdpa with operands scaled by %g: %.3f ms (incl. 3 scalings)" % (sc, t))
o1 = torch.nn.functional.scaled_dot_product_attention(q, q, q)
o2 = torch.nn.functional.scaled_dot_product_attention(q * 2048, q * 2048, q * 2048, scale=0.125 / 2048 / 2048) / 2048
print("scaled sdpa max diff %.2e" % (o1 - o2).abs().max().item())
