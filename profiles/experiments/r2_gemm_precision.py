"""Accuracy and speed of the candidate fp32-grade GEMM formulations on a ViT-L-shaped product (M=15470), against fp64.
 native   torch fp32 (SIMT sgemm)
 tf32     plain TF32
 x3       one TF32 GEMM over [x_hi|x_lo|x_hi] x [W_hi|W_hi|W_lo]^T (depth 3K)
 x3_2     two TF32 GEMMs: small = [x_lo|x_hi] x [W_hi|W_lo]^T (depth 2K), y = small + x_hi W_hi^T (depth K, beta=1)
 h16      fp16 hi/lo split (lo scaled by 2^11), fp32 accumulate/output, two GEMMs
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pnp_ovss_b200 import ops
from pnp_ovss_b200.blip_itm import _w3, _tf32_split

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def timeit(fn, n=10):
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


for (M, K, N) in [(15470, 1024, 3072), (15470, 4096, 1024), (15470, 1024, 4096)]:
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(dev)
    truth = (x[:2048].double() @ w.double().t())
    scale = truth.abs().max().item()
    res = {}

    def err(y):
        return ((y[:2048].double() - truth).abs().max().item() / scale, ((y[:2048].double() - truth).pow(2).mean().sqrt().item()) / scale)

    res["native"] = (err(x @ w.t()), timeit(lambda: x @ w.t()))
    torch.backends.cuda.matmul.allow_tf32 = True
    res["tf32"] = (err(x @ w.t()), timeit(lambda: x @ w.t()))
    x3, w3 = ops.tf32_split3(x), _w3(w)
    res["x3"] = (err(torch.nn.functional.linear(x3, w3)), timeit(lambda: torch.nn.functional.linear(x3, w3)))

    def two():
        small = torch.mm(x3[:, K:], w3[:, K:].t())
        return torch.addmm(small, x3[:, :K], w3[:, :K].t())
    res["x3_2"] = (err(two()), timeit(two))

    def two_b():  # small terms accumulated in place into the main product
        y = torch.mm(x3[:, :K], w3[:, :K].t())
        return y.addmm_(x3[:, K:], w3[:, K:].t())
    res["x3_2b"] = (err(two_b()), timeit(two_b))
    torch.backends.cuda.matmul.allow_tf32 = False
    # hi-only accumulation error: x_hi W_hi in TF32 against fp64 of the same operands
    xh, wh = _tf32_split(x)[0], _tf32_split(w)[0]
    torch.backends.cuda.matmul.allow_tf32 = True
    yh = xh @ wh.t()
    torch.backends.cuda.matmul.allow_tf32 = False
    th = xh[:2048].double() @ wh.double().t()
    res["hi_only_accum"] = ((((yh[:2048].double() - th).abs().max() / scale).item(), ((yh[:2048].double() - th).pow(2).mean().sqrt() / scale).item()), 0.0)
    yn = xh @ wh.t()
    res["hi_only_native"] = ((((yn[:2048].double() - th).abs().max() / scale).item(), ((yn[:2048].double() - th).pow(2).mean().sqrt() / scale).item()), 0.0)
    try:
        xh16 = x.half()
        xl16 = ((x - xh16.float()) * 2048.0).half()
        wh16 = w.half()
        wl16 = ((w - wh16.float()) * 2048.0).half()
        a2 = torch.cat([xl16, xh16], 1).contiguous()
        b2 = torch.cat([wh16, wl16], 1).contiguous()

        def h16():
            main = torch.mm(xh16, wh16.t(), out_dtype=torch.float32)
            small = torch.mm(a2, b2.t(), out_dtype=torch.float32)
            return main.add_(small, alpha=1.0 / 2048.0)
        res["h16"] = (err(h16()), timeit(h16))
    except Exception as e:  # noqa
        print("h16 unavailable:", str(e)[:200])
    print("M=%d K=%d N=%d  (errors relative to max |y| = %.3g: max, rms)" % (M, K, N, scale))
    for k, ((emax, erms), ms) in res.items():
        print("   %-16s max %.3e  rms %.3e   %.3f ms  (%.0f TF/s fp32-equivalent)" % (k, emax, erms, ms, (2.0 * M * K * N / (ms * 1e-3) / 1e12) if ms else 0))
