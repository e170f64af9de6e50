mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 300 python -m pytest tests/test_gpu_modules.py tests/test_gpu_backward.py tests/test_gpu_ops.py -m gpu -x -q -k "cosine or posconv or weight_norm" > gpurun_out/t47_tests.log 2>&1; tail -6 gpurun_out/t47_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
