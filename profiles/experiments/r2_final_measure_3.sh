# configs[2..4] again after the width-dependent splat / update grids
set -x
timeout 500 python -m pytest tests/test_gpu_crf.py -x -q -m gpu 2>&1 | tail -3
for k in 2 3 4; do
  timeout 600 python bench.py --config $k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_cfg$k.json 2> gpurun_out/r2g_bench_cfg$k.err
  tail -c 300 gpurun_out/r2g_bench_cfg$k.err
done
python - <<'PY'
import json
for k in (2, 3, 4):
    d = json.load(open("gpurun_out/r2g_bench_cfg%d.json" % k))
    print(k, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["ms_per_step"], 1), d["stages_ms_per_step"], d["clocks"]["sm_mhz"], d.get("parity"))
    print("   ", {kk: (vv["ms_per_step"], vv["frac"]) for kk, vv in d["kernels"].items() if vv["ms_per_step"] > 5})
PY
