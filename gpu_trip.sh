mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/t31_tests.log 2>&1; tail -6 gpurun_out/t31_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
for P in 1 0; do
W2V2_PDL=$P timeout -k 5 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t31_train_pdl$P.log 2>&1; tail -1 gpurun_out/t31_train_pdl$P.log | cut -c1-200
W2V2_PDL=$P timeout -k 5 300 python bench.py --mode forward --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t31_fwd_pdl$P.log 2>&1; tail -1 gpurun_out/t31_fwd_pdl$P.log | cut -c1-200
done
echo "bench done $(( $(date +%s) - T0 )) s"
timeout -k 5 200 python tools/timeline.py --mode train > gpurun_out/t31_timeline_train.log 2>&1; head -3 gpurun_out/t31_timeline_train.log | cut -c1-200
echo "all done $(( $(date +%s) - T0 )) s"
