#!/bin/bash
# ncu launch list of the bench command at the end of round 2 (cold-cache, serialised kernel times: shares, not a step time)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2u_train_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2u_bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2u_train_launches.csv
python tools/launch_summary.py gpurun_out/r2u_train_launches.csv --step | head -40
