# last step of the round: DRAM traffic re-measured on the final sources, then the default bench line
set -x
bash profiles/refresh_traffic.sh; cp profiles/dram_traffic.json gpurun_out/dram_traffic.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_cfg1.json 2> gpurun_out/r2h_bench_cfg1.err
tail -c 300 gpurun_out/r2h_bench_cfg1.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_cfg1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['ref_gpu']['value'])"
