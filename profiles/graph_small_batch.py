"""Launch-bound regime: one 336x336 image, 3 channels (the live class counts of the reference are 1-3 per image).
Eager launches vs one CUDA-graph replay of the same pnp_crf_inference call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnp_ovss_b200 import ops, synthetic as synth
dev = torch.device("cuda:0")
H = W = 336
for C in (3, 21):
    guides = torch.from_numpy(synth.guide_image(5, H, W)[None]).to(dev)
    lat_s = ops.build_lattice(H, W, 3.0, device=dev)
    lat_b = ops.build_lattice(H, W, 50.0, rgb=guides, srgb=5.0)
    Cp = ops.crf_pad_channels(C)
    U = torch.rand(1, H * W, Cp, device=dev)
    U[:, :, C:] = 0
    run = lambda: ops.crf_inference([lat_s, lat_b], [7.0, 10.0], U, C, 10)
    for _ in range(3): run()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): run()
    torch.cuda.synchronize(); eager = (time.perf_counter() - t) / 20
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): out = run()
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): g.replay()
    torch.cuda.synchronize(); graph = (time.perf_counter() - t) / 20
    print("C'=%d  B=1  eager %.3f ms  graph replay %.3f ms" % (C, eager * 1e3, graph * 1e3))
