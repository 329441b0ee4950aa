"""Where does one model pass (ViT-L + BERT + trimmed backward, fp32) spend its time?  torch.profiler top kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
tokens = w["tokens"].to(dev)
imgs = w["imgs"].to(dev)
for _ in range(2):
    model.gradcam(imgs, w["captions"], tokens, layer=7, head=9)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.gradcam(imgs, w["captions"], tokens, layer=7, head=9)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
