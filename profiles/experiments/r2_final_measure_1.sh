set -x
bash profiles/refresh_traffic.sh; cp profiles/dram_traffic.json gpurun_out/dram_traffic.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_cfg1.json 2> gpurun_out/r2f_bench_cfg1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-parity > gpurun_out/r2f_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention_tc5_kernel|attention_split|layernorm_split3|gelu_split3" -c 4 -o gpurun_out/prof_r2f_model python profiles/run_model_pass.py 1 > gpurun_out/prof_r2f_model.log 2>&1
timeout 200 python profiles/model_pass_profile.py 3xfp16 > gpurun_out/r2f_model_pass_profile.txt 2>&1
tail -3 gpurun_out/r2f_bench_cfg1.err; ls -la gpurun_out | tail -8
