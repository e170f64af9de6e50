# the standard validation trip (tools/trip.sh ships this file to a B200 box): GPU tests, smoke, default bench, forward bench
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 900 python -m pytest tests -m gpu -q > gpurun_out/trip_tests.log 2>&1; tail -5 gpurun_out/trip_tests.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/trip_smoke.log 2>&1; tail -1 gpurun_out/trip_smoke.log
echo "tests+smoke done $(( $(date +%s) - T0 )) s"
timeout -k 5 600 python bench.py > gpurun_out/trip_bench_default.log 2>&1; tail -1 gpurun_out/trip_bench_default.log | cut -c1-260
timeout -k 5 300 python bench.py --mode forward --no-cpu-baseline > gpurun_out/trip_bench_fwd.log 2>&1; tail -1 gpurun_out/trip_bench_fwd.log | cut -c1-260
echo "all done $(( $(date +%s) - T0 )) s"
