"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> profiles/dram_traffic.json:
average DRAM bytes per launch of every kernel CLASS bench.py reports (the names of pnp_profile_kernel_name)."""
import csv
import datetime
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

# kernel function (substring of the demangled name) -> class; splat / lattice blur serve both lattices: the spatial one is
# launched with grid.y = images, the batched bilateral one with grid.y = 1
FUNCTIONS = [("softmax_fwd_kernel", "softmax_fwd"), ("softmax_bwd_gradcam_kernel", "softmax_bwd_gradcam"), ("token_merge_kernel", "token_merge"),
             ("salience_dropout_round_kernel", "salience_dropout_round"), ("threshold_prep_kernel", "threshold_prep"),
             ("upsample_write_kernel", "upsample_write"), ("blur_vertical", "blur_vertical"), ("blur_horizontal_kernel", "blur_horizontal"),
             ("blur_normalize_kernel", "blur_normalize"), ("unary_from_maps_kernel", "crf_unary"), ("unary_write_kernel", "crf_unary"),
             ("lowrank_minmax_kernel", "lowrank_blur"), ("lowrank_unary_kernel", "lowrank_unary"), ("meanfield_update", "crf_meanfield_update"),
             ("argmax_channels", "argmax_channels"), ("confusion_kernel", "confusion"), ("split3_kernel", "tf32_split3")]


def classify(name, grid):
    if "splat_kernel" in name and "norm_splat" not in name:
        return "crf_splat_spatial" if grid[1] > 1 else "crf_splat_bilateral"
    if "blur_axis" in name:
        return "crf_blur_axis_spatial" if grid[1] > 1 else "crf_blur_axis_bilateral"
    for sub, cls in FUNCTIONS:
        if sub in name:
            return cls
    return None


def main(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[1:]:
        if len(r) != len(hdr) or not r[col["ID"]].isdigit():
            continue
        key = int(r[col["ID"]])
        d = launches.setdefault(key, {"name": r[col["Kernel Name"]], "grid": tuple(int(x) for x in r[col["Grid Size"]].strip("()").split(","))})
        val, unit = float(r[col["Metric Value"]].replace(",", "")), r[col["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1)
        d[r[col["Metric Name"]]] = val * scale
    per = {}
    for d in launches.values():
        cls = classify(d["name"], d["grid"])
        if cls is None:
            continue
        e = per.setdefault(cls, {"bytes": 0.0, "ms": 0.0, "n": 0})
        e["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        e["ms"] += d.get("gpu__time_duration.sum", 0.0)
        e["n"] += 1
    out = {"source_sha16": bench.source_fingerprint(), "workload": bench.CONFIGS[1]["name"],
           "captured": datetime.datetime.utcnow().strftime("%Y-%m-%d") + ", ncu metric pass over profiles/run_postprocess.py (cfg1 post-processing shapes, cold cache)",
           "kernels": {k: int(round(v["bytes"] / v["n"])) for k, v in sorted(per.items())},
           "launches": {k: v["n"] for k, v in sorted(per.items())},
           "ncu_ms_per_launch": {k: round(v["ms"] / v["n"], 4) for k, v in sorted(per.items())}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
