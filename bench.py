#!/usr/bin/env python
"""bench.py -- images/sec of end-to-end mask extraction (BASELINE.json metric) on N B200s of one node.

A step = one batch of BASELINE.json configs[1] ("VOC-shaped batch 35 @336, 21 classes, drop_iter 4, prune_att_head 9,
blur+crf") through the whole path: 4 Salience-DropOut rounds (random-init BLIP ITM-large pass in torch fp32 GEMMs,
fused softmax/GradCAM kernel, DropOut kernel), token merge, and -- for both maps the VOC driver scores (round-0 and
accumulated, DRV:348-403 / 424-481) -- threshold/upsample, Gaussian blur, dense CRF (10 mean-field iterations),
argmax + relabel + confusion matrix.  Data-parallel over images: every rank runs its own batch (weak scaling) and
the int64 confusion matrices are all-reduced once over NCCL at the end of the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path on the host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "images/sec end-to-end mask extraction @336"
VOC = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "table", "dog", "horse",
       "motorbike", "person", "plant", "sheep", "sofa", "train", "television"]
WORKLOAD = dict(name="voc21_b35_336_drop4_head9_blur+crf", B=35, S=336, P=21, C=20, n_class=21, drop_iter=4, layer=7, head=9,
                threshold=0.15, mode="blur+crf", data_type="voc")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "3xtf32", "tf32", "bf16"],
                    help="precision of the model's torch GEMMs (the reference runs fp32; anything else is reported in dtype)")
    ap.add_argument("--guide", default="natural", choices=["natural", "noise"], help="CRF guide image flavour (SURVEY 8d)")
    ap.add_argument("--ref-images", type=int, default=0, help="images per step of the CPU arm (0 = sized to the time budget)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra 3xtf32 measurement reported under `alt_gemm`")
    ap.add_argument("--cublas-emulation", action="store_true",
                    help="run torch's own fp32 GEMMs through cuBLAS 12.9's BF16x9 FP32 emulation: re-executes this script with the "
                         "system libcublas/libcublasLt 12.9 preloaded over the 12.8 that torch bundles (which has no emulation) and "
                         "CUBLAS_EMULATE_SINGLE_PRECISION=1; reported in `dtype`")
    ap.add_argument("--no-overlap", action="store_true", help="run the round-0 pass on the main stream instead of a side stream")
    ap.add_argument("--classes", default="all20", choices=["all20", "live"],
                    help="all20: every image is captioned with the 20 VOC classes (BASELINE configs[1]); live: classes per image drawn "
                         "from the reference's own GPT-4o answers for VOC (1:983, 2:366, 3:88, 4:11, 6:1 of 1449 images, mean 1.40)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------- workload
def make_workload(rank):
    from pnp_ovss_b200 import synthetic as synth
    w = dict(WORKLOAD)
    B, S, C, n = w["B"], w["S"], w["C"], w["n_class"]
    g = torch.Generator().manual_seed(1234 + rank)
    tok = synth.SyntheticWordPieceTokenizer()
    if w.get("classes", "all20") == "live":  # the caption the reference really feeds: GPT-4o classes with p > 70 (DRV:764-783)
        rng = np.random.default_rng(99 + rank)
        counts = rng.choice([1, 2, 3, 4, 6], size=B, p=np.array([983, 366, 88, 11, 1]) / 1449.0)
        idx = [sorted(rng.choice(20, size=int(c), replace=False).tolist()) for c in counts]
        w["name"] = w["name"].replace("voc21", "voc_live_classes")
    else:
        idx = [list(range(C)) for _ in range(B)]
    class_lists = [[VOC[i] for i in ids] for ids in idx]
    captions = ["A picture of " + " ".join(cl) for cl in class_lists]
    w.update(tok=tok, captions=captions, tokens=tok(captions, padding="max_length", max_length=500),
             class_lists=class_lists, dataset_ids=[[i + 1 for i in ids] for ids in idx],
             imgs=torch.randn(B, 3, S, S, generator=g),
             guides=np.stack([synth.guide_image(5000 + 97 * rank + b, S, S, w.get("guide", "natural")) for b in range(B)]),
             gts=np.stack([synth.gt_labels(7000 + 97 * rank + b, S, S, n) for b in range(B)]))
    return w


# --------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thread = [], None, None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------- roofline bookkeeping
LATTICE_LAUNCHES = {2: 22, 5: 25}  # kernels inside one pnp_lattice_build (see lattice.cu build_impl)


def algorithmic_bytes(kernel, w, stats, T):
    """Algorithmic bytes per LAUNCH of each custom kernel class (DESIGN.md 'Kernels and rooflines')."""
    B, C, P, S = w["B"], w["C"], w["P"], w["S"]
    Cc, N, K = C + 1, S * S, P * P + 1
    Ms, Mb = stats.get("M_s", 0), stats.get("M_b", 0)
    return {
        "softmax_fwd": 8 * B * 12 * T * K,
        "softmax_bwd_gradcam": 8 * B * (T - 1) * K + 4 * B * (T - 1) * (K - 1),
        "token_merge": 4 * B * (T - 1) * P * P + 4 * B * C * P * P,
        "salience_dropout_round": 4 * B * (T - 1) * P * P * 3 + B * 10 * 3 * 256 * 4,
        "upsample_write": 4 * B * C * P * P + 4 * B * Cc * N,
        "blur_vertical": 8 * B * Cc * N,
        "blur_horizontal": 8 * B * Cc * N,
        "crf_unary": 8 * B * Cc * N,
        "crf_splat_bilateral": 4 * Cc * (B * N + Mb) + 8 * 6 * B * N,
        # one launch fuses two axis passes (blur_axis2_kernel): two units of SURVEY 8(d)'s "read + write per blur pass"
        "crf_blur_axis_bilateral": 2 * (8 * Cc + 8) * Mb,
        "crf_splat_spatial": 4 * Cc * B * (N + Ms) + 8 * 3 * N,
        "crf_blur_axis_spatial": 1.5 * (8 * Cc + 8) * Ms * B,  # 3 axes in 2 launches (one fused pair + one single)
        "crf_meanfield_update": 4 * Cc * B * (2 * N) + 4 * Cc * (B * Ms + Mb) + 8 * 9 * B * N,
        "confusion": 12 * B * N,
        "argmax_channels": 4 * Cc * B * N + 4 * B * N,
    }.get(kernel)


# --------------------------------------------------------------------------------------------------- our arm
SYSTEM_CUBLAS = ("/usr/local/cuda/lib64/libcublasLt.so.12", "/usr/local/cuda/lib64/libcublas.so.12")


def emulation_env():
    """Environment in which torch's sgemm calls run as cuBLAS BF16x9 FP32 emulation, or None if the system cuBLAS is absent."""
    if not all(os.path.exists(p) for p in SYSTEM_CUBLAS):
        return None
    env = dict(os.environ)
    env["LD_PRELOAD"] = ":".join(list(SYSTEM_CUBLAS) + ([env["LD_PRELOAD"]] if env.get("LD_PRELOAD") else []))
    env["CUBLAS_EMULATE_SINGLE_PRECISION"] = "1"
    env["PNP_BENCH_CUBLAS_EMULATION"] = "1"
    return env


def run_emulated_child(args):
    """The same steps in a child process under emulation_env(); returns the `alt_gemm_emulated` object."""
    env = emulation_env()
    if env is None:
        return {"unavailable": "no system cuBLAS >= 12.9 under /usr/local/cuda/lib64"}
    cmd = [sys.executable, os.path.abspath(__file__), "--cublas-emulation", "--steps", str(args.steps), "--warmup", str(args.warmup),
           "--no-cpu-baseline", "--no-alt", "--guide", args.guide, "--classes", args.classes]
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    except (subprocess.SubprocessError, IndexError, ValueError) as e:
        return {"unavailable": "child run failed: %s" % str(e)[:160]}
    return {"gemm": "torch's own fp32 GEMM calls, executed by cuBLAS 12.9 as BF16x9 FP32 emulation (system libcublas preloaded over "
                    "torch's bundled 12.8, CUBLAS_EMULATE_SINGLE_PRECISION=1); `python bench.py --cublas-emulation` runs it as the main mode",
            "value": line["value"], "unit": "images/s", "ms_per_step": line["ms_per_step"],
            "e2e": (line.get("e2e") or {}).get("value"), "gradcam_max_err_vs_fp64_rel_to_max": line.get("gradcam_err_vs_fp64")}


def gradcam_fp64(model, imgs, captions, tokens, layer, head, P):
    """GradCAM of (layer, head) from an fp64 copy of the model with plain torch autograd (MED:228-300, BITM:399-433
    restated in torch; no custom kernel) -- the ground truth the GEMM-precision modes are judged against."""
    import copy
    import math
    m = copy.deepcopy(model).double().requires_grad_(True)
    m.gemm_precision = "fp32"
    xa = m.layer[layer].crossattention.self
    kept = {}

    def forward(hidden, enc, enc_mask=None):
        q, k, v = xa._split(xa.query(hidden)), xa._split(xa.key(enc)), xa._split(xa.value(enc))
        probs = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1]), -1)
        probs.retain_grad()
        kept["probs"] = probs
        ctx = torch.matmul(probs, v)
        B, h, T, d = ctx.shape
        return ctx.permute(0, 2, 1, 3).reshape(B, T, h * d)

    xa.forward = forward
    with torch.enable_grad():
        out = m(imgs.double(), captions)
        out[:, 1].sum().backward()
    p = kept["probs"]
    T = p.shape[2]
    mask = tokens.attention_mask[:, :T].double()
    cam = p[:, head, :, 1:] * p.grad[:, head, :, 1:].clamp(min=0) * mask[:, :, None]
    return cam[:, 1:].reshape(imgs.shape[0], T - 1, P, P).detach()


def run_ours(args):
    import torch.distributed as dist
    from pnp_ovss_b200 import _lib, pipeline
    from pnp_ovss_b200.blip_itm import BlipITM

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    WORKLOAD["guide"] = args.guide
    WORKLOAD["classes"] = args.classes
    w = make_workload(rank)
    B = w["B"]
    torch.manual_seed(4321)  # same random-init weights on every rank
    model = BlipITM(img_size=w["S"], tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
    model.gemm_precision = args.gemm
    if args.gemm == "tf32":
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    tokens_dev = w["tokens"].to(dev)
    token_ids = w["tokens"].input_ids.tolist()
    T = max(len(w["tok"].encode(c)) for c in w["captions"])

    # pinned host buffers (e2e) and resident device copies (value)
    imgs_h, guides_h, gts_h = w["imgs"].pin_memory(), torch.from_numpy(w["guides"]).pin_memory(), torch.from_numpy(w["gts"]).pin_memory()
    imgs_src = imgs_h.to(dev)
    imgs_d, guides_d, gts_d = torch.empty_like(imgs_src), guides_h.to(dev), gts_h.to(dev)
    total_hist = torch.zeros((w["n_class"], w["n_class"]), dtype=torch.int64, device=dev)
    hist_host = torch.empty((w["n_class"], w["n_class"]), dtype=torch.int64).pin_memory()

    def gradcam_fn(x):
        return model.gradcam(x, w["captions"], tokens_dev, layer=w["layer"], head=w["head"])[0]

    def step(e2e, stats=None):
        if e2e:  # host buffers in, host result out
            imgs_d.copy_(imgs_h, non_blocking=True)
            guides_d.copy_(guides_h, non_blocking=True)
            gts_d.copy_(gts_h, non_blocking=True)
        else:    # inputs already resident; DropOut zeroes pixel blocks in place, so restore the working copy
            imgs_d.copy_(imgs_src)
        h0, hagg, _ = pipeline.batch_confusion(gradcam_fn, imgs_d, token_ids, w["tok"].decode, w["class_lists"], w["dataset_ids"],
                                               gts_d, guides_d, drop_iter=w["drop_iter"], patch_num=w["P"],
                                               threshold=w["threshold"], data_type=w["data_type"], mode=w["mode"],
                                               n_class=w["n_class"], stats=stats, overlap=not args.no_overlap)
        total_hist.add_(hagg)
        if e2e:
            hist_host.copy_(hagg, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return h0, hagg

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, n_steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n_steps):
            step(e2e)
        if world > 1:
            dist.all_reduce(total_hist, op=dist.ReduceOp.SUM)  # the path's one exchange step
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    n_ids = 19
    tot = (ctypes.c_float * n_ids)()
    cnt = (ctypes.c_int * n_ids)()

    # ---- warm-up; the last warm-up step is profiled per kernel class to find the dominant custom kernel
    stats = {}
    for i in range(max(args.warmup, 3)):
        if i == max(args.warmup, 3) - 1:
            stats = {"events": []}
            lib.pnp_profile_start(ctypes.c_uint(0xFFFFFFFE))
        step(False, stats)
    torch.cuda.synchronize()
    lib.pnp_profile_stop(tot, cnt, n_ids)
    per_kernel = {lib.pnp_profile_kernel_name(i).decode(): (float(tot[i]), int(cnt[i])) for i in range(1, n_ids) if cnt[i]}
    launches_per_step = sum(c for k, (_, c) in per_kernel.items() if k != "lattice_build")
    launches_per_step += per_kernel.get("lattice_build", (0, 0))[1] * LATTICE_LAUNCHES[5]
    ev = stats.pop("events")
    stages = {}
    for (n0, e0), (n1, e1) in zip(ev, ev[1:]):
        stages[n1] = stages.get(n1, 0.0) + e0.elapsed_time(e1)
    dominant = max((k for k in per_kernel if k != "lattice_build"), key=lambda k: per_kernel[k][0])
    dom_id = next(i for i in range(1, n_ids) if lib.pnp_profile_kernel_name(i).decode() == dominant)

    # ---- timed region (inputs resident), dominant kernel bracketed by events in situ
    sampler = ClockSampler(local_rank)
    sampler.start()
    # kernels of the round-0 pass run on a side stream under the model's GEMMs; the roofline is taken from the launches on
    # the main stream (the all-drop pass), which own the GPU while they run
    lib.pnp_profile_filter_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), 1)
    lib.pnp_profile_start(ctypes.c_uint(1 << dom_id))
    ms_total = timed(False, args.steps)
    lib.pnp_profile_stop(tot, cnt, n_ids)
    lib.pnp_profile_filter_stream(ctypes.c_void_p(0), 0)
    clocks = sampler.stop()
    dom_ms = float(tot[dom_id]) / max(int(cnt[dom_id]), 1)
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end through host buffers
    e2e = None
    if not args.no_e2e:
        step(True)
        e2e_ms = timed(True, args.steps)
        e2e = {"value": world * B * args.steps / (e2e_ms / 1e3), "unit": "images/s",
               "h2d_bytes_per_step": int(imgs_h.numel() * 4 + guides_h.numel() + gts_h.numel() * 4),
               "d2h_bytes_per_step": int(hist_host.numel() * 8), "ms_per_step": e2e_ms / args.steps}

    # ---- the same steps with the ViT linears as error-compensated TF32 products (reported beside the fp32 headline)
    alt = None
    if args.gemm == "fp32" and not args.no_alt:
        model.gemm_precision = "3xtf32"
        step(False)
        alt_ms = timed(False, args.steps)
        probe = imgs_src[:4].contiguous()
        tok4 = w["tok"](w["captions"][:4], padding="max_length", max_length=500).to(dev)
        cam_3x = model.gradcam(probe, w["captions"][:4], tok4, layer=w["layer"], head=w["head"])[0]
        model.gemm_precision = "fp32"
        cam_32 = model.gradcam(probe, w["captions"][:4], tok4, layer=w["layer"], head=w["head"])[0]
        dev_rel = float(((cam_3x - cam_32).abs().max() / cam_32.abs().max()).item())
        vs_fp64 = None
        if rank == 0:
            try:  # both modes against an fp64 torch-autograd pass of the same model (ground truth), outside any timed region
                truth = gradcam_fp64(model, probe, w["captions"][:4], tok4, w["layer"], w["head"], w["P"])
                sc = truth.abs().max()
                vs_fp64 = {"fp32": float(((cam_32.double() - truth).abs().max() / sc).item()),
                           "3xtf32": float(((cam_3x.double() - truth).abs().max() / sc).item())}
                del truth
            except RuntimeError as e:  # e.g. out of memory next to the resident workload: the timing above stands
                vs_fp64 = {"error": str(e).splitlines()[0][:120]}
            torch.cuda.empty_cache()
        alt = {"gemm": "3xtf32 (x_hi W_hi + x_hi W_lo + x_lo W_hi on TF32 tensor cores, fp32 accumulate, ViT linears only)",
               "value": world * B * args.steps / (alt_ms / 1e3), "unit": "images/s", "ms_per_step": alt_ms / args.steps,
               "gradcam_max_dev_vs_fp32_rel_to_max": dev_rel, "gradcam_max_err_vs_fp64_rel_to_max": vs_fp64}

    emulated = bool(os.environ.get("PNP_BENCH_CUBLAS_EMULATION"))
    err_fp64 = None
    if emulated and rank == 0:  # this process IS the emulated run: how far is its GradCAM from the fp64 pass?
        probe = imgs_src[:4].contiguous()
        tok4 = w["tok"](w["captions"][:4], padding="max_length", max_length=500).to(dev)
        cam = model.gradcam(probe, w["captions"][:4], tok4, layer=w["layer"], head=w["head"])[0]
        truth = gradcam_fp64(model, probe, w["captions"][:4], tok4, w["layer"], w["head"], w["P"])
        err_fp64 = float(((cam.double() - truth).abs().max() / truth.abs().max()).item())
        del truth
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    alt_emulated = None
    if args.gemm == "fp32" and not args.no_alt and not emulated and world == 1:
        alt_emulated = run_emulated_child(args)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    abytes = algorithmic_bytes(dominant, w, stats, T) if args.classes == "all20" else None  # ragged buckets: no single figure
    achieved = abytes / (dom_ms * 1e-3) / 1e9 if abytes else None
    traffic = None
    try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture, if any
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(dominant)
    except (OSError, ValueError):
        pass
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": abytes, "avg_launch_ms": dom_ms, "launches_timed": int(cnt[dom_id]),
                "share_of_step": per_kernel[dominant][0] / ms_per_step}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline = run_reference_steps(w, n_images=2, steps=1, warmup=0)

    kernels_ms = {k: {"ms_per_step": round(v[0], 3), "launches": v[1],
                      "GBps": (round(algorithmic_bytes(k, w, stats, T) * v[1] / (v[0] * 1e-3) / 1e9, 1)
                               if args.classes == "all20" and algorithmic_bytes(k, w, stats, T) and v[0] > 0 else None)}
                  for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])}
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("f32 (cuBLAS 12.9 BF16x9 FP32 emulation of torch's sgemm calls)" if emulated else
                                     {"fp32": "f32", "3xtf32": "f32 (ViT GEMMs as 3 error-compensated TF32 products)"}.get(args.gemm, args.gemm)), "data": "synthetic",
            "config": {"workload": w["name"], "images_per_step_per_gpu": B, "img_size": w["S"], "patch_grid": w["P"],
                       "classes": w["C"] if args.classes == "all20" else "live (mean %.2f per image)" % (sum(len(c) for c in w["class_lists"]) / B),
                       "channels": w["C"] + 1 if args.classes == "all20" else "classes + background", "drop_iter": w["drop_iter"], "block": w["layer"] + 1,
                       "head": w["head"], "postprocess": w["mode"], "crf_iters": 10, "tokens_T": T, "guide": args.guide,
                       "model": "BLIP ITM-large shape, random init, torch %s GEMMs, trimmed backward" % args.gemm,
                       "passes": "round0 + all_drop (DRV:348-403, 424-481)",
                       "overlap": "off" if args.no_overlap else "lattice build + round-0 pass on a side stream under DropOut rounds 1-3", "parallelism": "dp%d over images" % world,
                       "M_s": stats.get("M_s"), "M_b_per_batch": stats.get("M_b"),
                       "l2": "per-step working set (>3 GB) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "alt_gemm": alt, "alt_gemm_emulated": alt_emulated,
            "gradcam_err_vs_fp64": err_fp64,
            "custom_kernels": {"ms_per_step": round(sum(v[0] for v in per_kernel.values()), 3),
                               "images_per_s": round(B / (sum(v[0] for v in per_kernel.values()) * 1e-3), 1),
                               "note": "sum of the in-situ event times of every pnp:: kernel in one (warm-up) step; the rest of "
                                       "the step is the model's torch fp32 GEMMs"},
            "stages_ms_per_step": {k: round(v, 3) for k, v in stages.items()}, "kernels": kernels_ms,
            "hist_total": int(total_hist.sum().item())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------- CPU arm
def run_reference_steps(w, n_images, steps, warmup):
    """The reference's CPU path (oracle/reference_arm.py) on the first n_images of the batch; returns the
    cpu_baseline object.  All host threads: torch intra-op threads for the model, one process per image afterwards."""
    from oracle import reference_arm as RA
    from pnp_ovss_b200.blip_itm import BlipITM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(4321)
    model = RA.install_reference_capture(BlipITM(img_size=w["S"], tokenizer=w["tok"]).eval())
    n = n_images
    tok = w["tok"]
    tokens = tok(w["captions"][:n], padding="max_length", max_length=500)
    pool = RA.make_pool(min(cores, 2 * n))
    pool.map(abs, range(min(cores, 2 * n)))  # spin the workers up outside the timed region
    times = []
    timings = {}
    for i in range(warmup + steps):
        if i == warmup:
            timings.clear()
        t0 = time.perf_counter()
        RA.reference_batch_confusion(model, w["imgs"][:n].clone(), w["captions"][:n], tokens, tok.decode, w["class_lists"][:n],
                                     w["dataset_ids"][:n], list(w["gts"][:n]), list(w["guides"][:n]),
                                     drop_iter=w["drop_iter"], layer=w["layer"], head=w["head"], threshold=w["threshold"],
                                     data_type=w["data_type"], mode=w["mode"], n_class=w["n_class"], pool=pool, timings=timings)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    pool.close()
    pool.join()
    sec = sum(times) / len(times)
    return {"value": n / sec, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d of the %d images of one batch per step (same seeds), model pass as BITM:386-457 in torch-CPU fp32 "
                      "(12-block capture, full backward), post-processing as DRV:424-481 via oracle/ (scipy gaussian_filter, "
                      "C restatement of pydensecrf), both reference passes; %.1f s per step" % (n, w["B"], sec),
            "sec_per_step": sec, "model_sec_per_step": timings.get("model_s", 0.0) / len(times),
            "post_sec_per_step": timings.get("post_s", 0.0) / len(times), "post_workers": min(cores, 2 * n)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    WORKLOAD["guide"] = args.guide
    WORKLOAD["classes"] = args.classes
    w = make_workload(0)
    n_steps = args.steps + args.warmup
    n = args.ref_images or max(1, min(4, int(160.0 / (max(n_steps, 1) * 12.0))))
    cb = run_reference_steps(w, n, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "images_per_step": n, "img_size": w["S"], "classes": w["C"],
                       "drop_iter": w["drop_iter"], "postprocess": w["mode"], "device": "host CPU"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse_args()
    if a.cublas_emulation and a.impl == "ours" and not os.environ.get("PNP_BENCH_CUBLAS_EMULATION"):
        env = emulation_env()
        if env is None:
            raise SystemExit("bench.py --cublas-emulation: no system cuBLAS >= 12.9 under /usr/local/cuda/lib64")
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
