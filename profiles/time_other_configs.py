"""Post-processing time (upsample -> blur -> unary -> CRF -> confusion, one pass) at the other BASELINE.json shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnp_ovss_b200 import _lib, ops, pipeline, synthetic as synth
dev = torch.device("cuda:0")
lib = _lib.load()
for name, B, C, P, S, with_bg, n in (("voc21@336", 35, 20, 21, 336, True, 21), ("context59@336", 35, 59, 21, 336, False, 60),
                                     ("ade150@336", 35, 150, 21, 336, False, 151),
                                     ("coco_obj81@448", 35, 80, 28, 448, True, 91), ("coco_stuff171@512", 8, 171, 21, 512, False, 183)):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    maps = torch.stack([synth.saliency_maps(100 + b, C, P) for b in range(B)])
    maps[:, :, :4, :4] = 0
    maps = maps.to(dev)
    guides = torch.from_numpy(np.stack([synth.guide_image(5000 + b, S, S) for b in range(B)])).to(dev)
    gts = torch.from_numpy(np.stack([synth.gt_labels(7000 + b, S, S, n) for b in range(B)])).to(dev)
    Cc = C + (1 if with_bg else 0)
    luts = (torch.arange(Cc, dtype=torch.int32, device=dev) + (0 if with_bg else 1)).repeat(B, 1)
    lat_b = ops.build_lattice(S, S, 50.0, rgb=guides, srgb=5.0)
    tot = (ctypes.c_float * 19)(); cnt = (ctypes.c_int * 19)()
    for it in range(2):
        if it == 1:
            lib.pnp_profile_start(ctypes.c_uint(0xFFFFFFFE))
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t0.record()
        hist = torch.zeros((n, n), dtype=torch.int64, device=dev)
        pipeline.postprocess_batch(maps, guides, gts, luts, hist, threshold=0.15, rescale=False, with_background=with_bg,
                                   mode="blur+crf", n_class=n, bilateral=lat_b)
    t1.record(); torch.cuda.synchronize()
    lib.pnp_profile_stop(tot, cnt, 19)
    k = {lib.pnp_profile_kernel_name(i).decode(): round(tot[i], 2) for i in range(1, 19) if cnt[i]}
    print("%-18s B=%d  pass %.1f ms (%.2f ms/image)  M_b=%d  " % (name, B, t0.elapsed_time(t1), t0.elapsed_time(t1) / B, lat_b.M),
          {a: b for a, b in sorted(k.items(), key=lambda kv: -kv[1])[:5]})
    del maps, guides, gts, lat_b
    torch.cuda.empty_cache()
