# the other BASELINE configurations on one GPU with the final kernels (CPU legs skipped: unchanged since profiles/r2_bench_line_cfg*.json)
set -x
for k in 0 2 3 4; do
  timeout 600 python bench.py --config $k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_cfg$k.json 2> gpurun_out/r2f_bench_cfg$k.err
  tail -c 300 gpurun_out/r2f_bench_cfg$k.err
done
python - <<'PY'
import json
for k in (0, 2, 3, 4):
    try:
        d = json.load(open("gpurun_out/r2f_bench_cfg%d.json" % k))
        print(k, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["ms_per_step"], 1), d["stages_ms_per_step"], d["clocks"]["sm_mhz"], d.get("parity"))
    except Exception as e:
        print(k, "failed", e)
PY
