mkdir -p gpurun_out
T0=$(date +%s)
for W in cfg2 cfg3 cfg4; do
timeout -k 5 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t32_train_$W.log 2>&1; tail -1 gpurun_out/t32_train_$W.log | cut -c1-220
done
timeout -k 5 300 python bench.py --workload cfg4 --mode forward --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t32_fwd_cfg4.log 2>&1; tail -1 gpurun_out/t32_fwd_cfg4.log | cut -c1-220
timeout -k 5 300 python bench.py --workload cfg2 --mode forward --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t32_fwd_cfg2.log 2>&1; tail -1 gpurun_out/t32_fwd_cfg2.log | cut -c1-220
echo "bench done $(( $(date +%s) - T0 )) s"
timeout -k 5 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 130 -c 4 -o gpurun_out/t32_gemm_full python bench.py --mode forward --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t32_ncu1.log 2>&1; tail -1 gpurun_out/t32_ncu1.log | cut -c1-200
echo "ncu1 done $(( $(date +%s) - T0 )) s"
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1650 -c 700 --csv --log-file gpurun_out/t32_train_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/t32_ncu2.log 2>&1; tail -1 gpurun_out/t32_ncu2.log | cut -c1-200
echo "all done $(( $(date +%s) - T0 )) s"
