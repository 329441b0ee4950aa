"""CPU: bench.py's bookkeeping -- workload shapes of BASELINE.json configs[1], the algorithmic-bytes table covering
the kernel classes the library can time, and the JSON keys the driver contract names."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_is_baseline_config_1():
    w = bench.make_workload(0)
    assert (w["B"], w["S"], w["P"], w["C"], w["n_class"], w["drop_iter"], w["head"], w["layer"] + 1, w["mode"]) == \
        (35, 336, 21, 20, 21, 4, 9, 8, "blur+crf")
    assert tuple(w["imgs"].shape) == (35, 3, 336, 336) and w["guides"].shape == (35, 336, 336, 3) and w["guides"].dtype == np.uint8
    assert w["gts"].shape == (35, 336, 336) and w["gts"].dtype == np.float32
    assert all(len(c) == 20 for c in w["class_lists"]) and w["dataset_ids"][0] == list(range(1, 21))
    w1 = bench.make_workload(1)
    assert not np.array_equal(w["guides"], w1["guides"])          # ranks get different images (weak scaling)


def test_live_class_distribution():
    bench.WORKLOAD["classes"] = "live"
    try:
        w = bench.make_workload(0)
    finally:
        bench.WORKLOAD.pop("classes")
    counts = [len(c) for c in w["class_lists"]]
    assert min(counts) >= 1 and max(counts) <= 6 and 1.0 <= np.mean(counts) <= 2.2
    assert all(ids == sorted(ids) and all(1 <= i <= 20 for i in ids) for ids in w["dataset_ids"])


def test_algorithmic_bytes_cover_the_timed_kernel_classes():
    from pnp_ovss_b200 import _lib
    lib = _lib.load()
    w = bench.make_workload(0)
    stats = {"M_s": 14755, "M_b": 3245257}
    names = [lib.pnp_profile_kernel_name(i).decode() for i in range(1, 19)]
    assert len(set(names)) == 18 and all(names)
    latency_bound = {"threshold_prep", "blur_normalize", "lattice_build"}   # reported in ms, not GB/s
    for n in names:
        b = bench.algorithmic_bytes(n, w, stats, 31)
        assert (b is None) == (n in latency_bound), n
        if b is not None:
            assert b > 0
    # the dominant kernels' figures of DESIGN.md section 3
    assert bench.algorithmic_bytes("crf_blur_axis_bilateral", w, stats, 31) == 2 * (8 * 21 + 8) * 3245257
    assert bench.algorithmic_bytes("softmax_fwd", w, stats, 31) == 8 * 35 * 12 * 31 * 442


def test_reference_arm_sample_is_bounded():
    # the CPU arm sizes its per-step sample so that warm-up + steps stay within a few minutes
    for steps, warmup in ((3, 3), (10, 3), (1, 0), (20, 5)):
        n = max(1, min(4, int(160.0 / (max(steps + warmup, 1) * 12.0))))
        assert 1 <= n <= 4


def test_cublas_emulation_environment(monkeypatch):
    """bench.py --cublas-emulation: the system cuBLAS 12.9 pair is preloaded ahead of anything already there and the
    emulation switch is set; without the libraries the mode reports itself unavailable instead of failing."""
    import bench
    monkeypatch.setenv("LD_PRELOAD", "/opt/other.so")
    env = bench.emulation_env()
    if all(os.path.exists(p) for p in bench.SYSTEM_CUBLAS):
        assert env["LD_PRELOAD"].split(":") == list(bench.SYSTEM_CUBLAS) + ["/opt/other.so"]
        assert env["CUBLAS_EMULATE_SINGLE_PRECISION"] == "1" and env["PNP_BENCH_CUBLAS_EMULATION"] == "1"
    else:
        assert env is None
    monkeypatch.setattr(bench, "SYSTEM_CUBLAS", ("/nonexistent/libcublasLt.so.12", "/nonexistent/libcublas.so.12"))
    assert bench.emulation_env() is None
    args = bench.parse_args.__globals__["argparse"].Namespace(steps=1, warmup=3, guide="natural", classes="all20")
    assert "unavailable" in bench.run_emulated_child(args)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` end to end on the host CPU (one image, one step): the JSON line the driver parses."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-images", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == bench.WORKLOAD["name"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: on a box without CUDA the product arm of bench.py stops with a clear message (and exit code != 0)."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
