"""fp16 hi/lo split GEMM candidates (fp32-grade result on the fp16 tensor cores, 2x the TF32 rate) against fp64:
 h16_2     main = h Wh^T ; y = main + 2^-11 ([l|h] [Wh|Wl]^T)  through torch.addmm(out_dtype=fp32, alpha=2^-11, beta=1)
 h16_1     one depth-3K GEMM on [h*8 | l | h] x [Wh*256 | Wh | Wl]^T, result scaled by 2^11 (consumer must undo)
 x3_2b     the TF32 formulation in use
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pnp_ovss_b200 import ops
from pnp_ovss_b200.blip_itm import _w3

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


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


def split16(t):
    h = t.half()
    l = ((t - h.float()) * 2048.0).half()
    return h, l


for (M, K, N) in [(15470, 1024, 3072), (15470, 1024, 1024), (15470, 1024, 4096), (15470, 4096, 1024), (15470, 1024, 18432)]:
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    truth = x[:2048].double() @ w.double().t() + bias.double()
    scale = truth.abs().max().item()

    def err(y):
        d = y[:2048].double() - truth
        return d.abs().max().item() / scale, d.pow(2).mean().sqrt().item() / scale
    res = {}
    xh, xl = split16(x)
    wh, wl = split16(w)
    A = torch.cat([xh, xl, xh], 1).contiguous()
    Bm = torch.cat([wh, wh, wl], 1).contiguous()

    def h16_2():
        main = torch.addmm(bias, A[:, :K], Bm[:, :K].t(), out_dtype=torch.float32)
        return torch.addmm(main, A[:, K:], Bm[:, K:].t(), out_dtype=torch.float32, alpha=2.0 ** -11)
    try:
        res["h16_2 (addmm alpha)"] = (err(h16_2()), timeit(h16_2))
    except Exception as e:  # noqa
        print("h16_2 addmm path unavailable:", str(e)[:300])

        def h16_2b():
            main = torch.mm(A[:, :K], Bm[:, :K].t(), out_dtype=torch.float32)
            small = torch.mm(A[:, K:], Bm[:, K:].t(), out_dtype=torch.float32)
            return main.add_(small, alpha=2.0 ** -11).add_(bias)
        res["h16_2 (mm + add_)"] = (err(h16_2b()), timeit(h16_2b))
    A1 = torch.cat([(x * 8).half() if False else (xh.float() * 8).half(), xl, xh], 1).contiguous()
    B1 = torch.cat([(wh.float() * 256).half(), wh, wl], 1).contiguous()

    def h16_1():
        return torch.mm(A1, B1.t(), out_dtype=torch.float32)
    y1 = h16_1() * 2.0 ** -11 + bias
    res["h16_1 (one chain, x2^11)"] = (err(y1), timeit(h16_1))
    x3, w3 = ops.tf32_split3(x), _w3(w)
    torch.backends.cuda.matmul.allow_tf32 = True

    def x3_2b():
        y = torch.nn.functional.linear(x3[:, :K], w3[:, :K], bias)
        return y.addmm_(x3[:, K:], w3[:, K:].t())
    res["x3_2b (TF32, in use)"] = (err(x3_2b()), timeit(x3_2b))
    torch.backends.cuda.matmul.allow_tf32 = False
    res["native fp32"] = (err(torch.nn.functional.linear(x, w, bias)), timeit(lambda: torch.nn.functional.linear(x, w, bias), 5))
    print("M=%d K=%d N=%d" % (M, K, N))
    for k, ((emax, erms), ms) in res.items():
        print("   %-28s max %.3e  rms %.3e   %.3f ms  (%.0f TF/s fp32-equivalent)" % (k, emax, erms, ms, 2.0 * M * K * N / (ms * 1e-3) / 1e12))
