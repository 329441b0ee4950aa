#!/bin/bash
# Re-measures the per-launch DRAM traffic of every custom kernel class on the cfg1 post-processing shapes and rewrites
# profiles/dram_traffic.json, stamped with the sha16 of the CUDA sources (bench.py reports `roofline.traffic` from it and
# marks it stale when the sources have changed since).  Run on a GPU box from the repo root:
#     gpurun -- bash profiles/refresh_traffic.sh
# ncu's metric pass is cold-cache and serialised: the BYTES are what is kept, not the times.
set -e
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:pnp:: --csv --log-file gpurun_out/traffic_r2.csv python profiles/run_postprocess.py 1 > gpurun_out/traffic_r2.log 2>&1
python profiles/parse_traffic.py gpurun_out/traffic_r2.csv profiles/dram_traffic.json
