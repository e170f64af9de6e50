mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_training.py -m gpu -q -x -s 2>&1 | tail -25 | cut -c1-400
timeout -k 5 900 python bench.py --steps 10 --warmup 3 --mode train > gpurun_out/t9_bench_train.log 2>&1; tail -2 gpurun_out/t9_bench_train.log | cut -c1-3000
