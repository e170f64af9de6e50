mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/t15_tests.log 2>&1; tail -5 gpurun_out/t15_tests.log
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t15_train1.log 2>&1; tail -1 gpurun_out/t15_train1.log | cut -c1-300
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reg > gpurun_out/t15_train1_noreg.log 2>&1; tail -1 gpurun_out/t15_train1_noreg.log | cut -c1-300
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1100 --csv --log-file gpurun_out/t15_train_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/t15_ncu.log 2>&1; tail -2 gpurun_out/t15_ncu.log | cut -c1-200
