mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests_all.log 2>&1; tail -5 gpurun_out/r2g_tests_all.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_train.log 2>&1; tail -c 300 gpurun_out/r2g_train.log
W2V2_DGRAD_ACCUM=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_train_noaccum.log 2>&1; tail -c 300 gpurun_out/r2g_train_noaccum.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2g_train_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ncu_bench.log 2>&1
for w in cfg2 cfg3 cfg4; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_train_$w.log 2>&1; tail -c 200 gpurun_out/r2g_train_$w.log; done
python bench.py --workload cfg4 --mode forward --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_fwd_cfg4.log 2>&1
