mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/t14_tests.log 2>&1; tail -5 gpurun_out/t14_tests.log
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t14_train1.log 2>&1; tail -1 gpurun_out/t14_train1.log | cut -c1-300
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1100 --csv --log-file gpurun_out/t14_train_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/t14_ncu.log 2>&1; tail -2 gpurun_out/t14_ncu.log | cut -c1-200
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd -s 30 -c 2 -o gpurun_out/t14_lnbwd -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t14_ncu2.log 2>&1; tail -2 gpurun_out/t14_ncu2.log | cut -c1-200
