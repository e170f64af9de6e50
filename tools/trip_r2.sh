mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ragged.py tests/test_gpu_round2.py -m gpu -q > gpurun_out/r2i_tests_new.log 2>&1; tail -12 gpurun_out/r2i_tests_new.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_ragged.py --deselect tests/test_gpu_round2.py > gpurun_out/r2i_tests_all.log 2>&1; tail -4 gpurun_out/r2i_tests_all.log
timeout 300 python tools/eval_throughput.py 256 > gpurun_out/r2i_eval_throughput.log 2>&1; tail -5 gpurun_out/r2i_eval_throughput.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_train.log 2>&1; tail -c 300 gpurun_out/r2i_train.log
