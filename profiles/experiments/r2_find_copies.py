"""Which Python lines launch aten::copy_ kernels inside one eager encoder pass (3xFP16 mode)?"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import bench
from torch.profiler import profile, ProfilerActivity
from pnp_ovss_b200.blip_itm import BlipITM

dev = torch.device("cuda:0")
w = bench.make_workload(0)
torch.manual_seed(4321)
model = BlipITM(img_size=336, tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
model.gemm_precision = "3xfp16"
model.USE_VIT_GRAPH = False
tokens = w["tokens"].to(dev)
imgs = w["imgs"].to(dev)
with torch.no_grad():
    for _ in range(2):
        model._vit3(imgs, mode="3xfp16")
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        model._vit3(imgs, mode="3xfp16")
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
agg = collections.Counter(); cnt = collections.Counter()
for ev in prof.events():
    if ev.name in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::cat"):
        frames = [f for f in (ev.stack or []) if "pnp_ovss_b200" in f]
        key = (ev.name, frames[0] if frames else "?")
        agg[key] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        cnt[key] += 1
for k, v in agg.most_common(12):
    print("%8.1f us  x%d  %s  %s" % (v, cnt[k], k[0], k[1]))
