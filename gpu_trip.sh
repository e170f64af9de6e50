mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/t28_tests.log 2>&1; tail -6 gpurun_out/t28_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
timeout -k 5 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t28_train1.log 2>&1; tail -1 gpurun_out/t28_train1.log | cut -c1-250
echo "all done $(( $(date +%s) - T0 )) s"
