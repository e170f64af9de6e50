mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/t49_tests.log 2>&1; tail -6 gpurun_out/t49_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/t49_train1.log 2>&1; tail -1 gpurun_out/t49_train1.log | cut -c1-200
echo "all done $(( $(date +%s) - T0 )) s"
