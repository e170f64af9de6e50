mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 150 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm or conv1d" > gpurun_out/t20_gemm.log 2>&1; tail -15 gpurun_out/t20_gemm.log
echo "gemm tests done $(( $(date +%s) - T0 )) s"
if grep -q "passed" gpurun_out/t20_gemm.log && ! grep -q "failed" gpurun_out/t20_gemm.log; then
timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/t20_tests.log 2>&1; tail -5 gpurun_out/t20_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t20_train1.log 2>&1; tail -1 gpurun_out/t20_train1.log | cut -c1-300
timeout -k 5 300 python bench.py --mode forward --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t20_fwd1.log 2>&1; tail -1 gpurun_out/t20_fwd1.log | cut -c1-300
echo "bench done $(( $(date +%s) - T0 )) s"
timeout -k 5 200 python tools/time_ops.py > gpurun_out/t20_time_ops.log 2>&1; tail -30 gpurun_out/t20_time_ops.log
fi
echo "all done $(( $(date +%s) - T0 )) s"
