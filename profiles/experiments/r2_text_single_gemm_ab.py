"""A/B of the text side's 3xTF32 linears as one depth-3K GEMM per linear (one launch) against main + correction GEMMs (two launches):
time per model pass (CUDA graphs on) and GradCAM error against an fp64 pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, bench
from pnp_ovss_b200 import blip_itm
from pnp_ovss_b200.blip_itm import BlipITM
dev = torch.device("cuda:0")
w = bench.make_workload(0)
torch.manual_seed(4321)
model = BlipITM(img_size=336, tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
model.gemm_precision = "3xfp16"
imgs, caps = w["imgs"].to(dev), w["captions"]
tok = w["tokens"].to(dev)
n = 4
tok4 = w["tok"](caps[:n], padding="max_length", max_length=500).to(dev)
truth = bench.gradcam_fp64(model, imgs[:n].contiguous(), caps[:n], tok4, 7, 9, 21)
for use in (False, True, False, True):
    blip_itm._Linear3.SINGLE_GEMM = use
    model.__dict__.pop("_text_graphs", None)      # the text pass is replayed from a CUDA graph: capture it again in this mode
    for _ in range(3):
        model.gradcam(imgs, caps, tok, layer=7, head=9)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(8):
        cam, _ = model.gradcam(imgs, caps, tok, layer=7, head=9)
    t1.record(); torch.cuda.synchronize()
    got, _ = model.gradcam(imgs[:n].contiguous(), caps[:n], tok4, layer=7, head=9)
    print("SINGLE_GEMM=%s: model pass %.2f ms   GradCAM vs fp64 %.3e" % (use, t0.elapsed_time(t1) / 8, ((got.double() - truth).abs().max() / truth.abs().max()).item()))
