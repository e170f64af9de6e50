# First GPU trip of round 2 (ship with: cp tools/gpu_trip_round2_first.sh tools/gpu_trip.sh && tools/trip.sh 1500 /tmp/r2a.log).
# Validates what was written after round 1's GPU budget ran out, A/B-measures the two experimental switches, runs the
# stock-kernel yardstick and captures the two kernels DESIGN.md section 7 puts first.  ~8 GPU-minutes.
mkdir -p gpurun_out
T0=$(date +%s)
# 1. the regular suite (includes the fixture-based tests added blind), then the experimental kernels' own test
timeout -k 5 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_tests.log 2>&1; tail -5 gpurun_out/r2a_tests.log
W2V2_EXPERIMENTAL=1 timeout -k 5 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k experimental > gpurun_out/r2a_experimental.log 2>&1; tail -3 gpurun_out/r2a_experimental.log
# 2. the experimental switches through the training parity tests
W2V2_SAVE_GELU_GRAD=1 W2V2_PREP_V2=1 timeout -k 5 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_regularise.py -m gpu -q > gpurun_out/r2a_tests_switches.log 2>&1; tail -3 gpurun_out/r2a_tests_switches.log
echo "tests done $(( $(date +%s) - T0 )) s"
# 3. A/B: default, +gelu' kept in the forward, +prepare v2, both
for cfg in "" "W2V2_SAVE_GELU_GRAD=1" "W2V2_PREP_V2=1" "W2V2_SAVE_GELU_GRAD=1 W2V2_PREP_V2=1"; do
  tag=$(echo "$cfg" | tr -c 'A-Z0-9\n' '_'); tag=${tag:-default}
  env $cfg timeout -k 5 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_$tag.log 2>&1
  echo "$tag: $(tail -1 gpurun_out/r2a_bench_$tag.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"])' 2>/dev/null)"
done
echo "bench done $(( $(date +%s) - T0 )) s"
# 4. stock kernels on the same box
timeout -k 5 600 python tools/hf_eager_bench.py > gpurun_out/r2a_hf_eager.log 2>&1; grep '"impl"' gpurun_out/r2a_hf_eager.log | cut -c1-200
# 5. ncu --set full: GELU-backward GEMM, attention backward (one launch each; names as the launch list prints them)
timeout -k 5 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 70 -c 1 -f -o gpurun_out/r2a_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_gemm.log 2>&1
timeout -k 5 300 ncu --set full --import-source on --clock-control none -k regex:attention_bwd_fused_kernel -s 5 -c 1 -f -o gpurun_out/r2a_attn_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_attn.log 2>&1
echo "all done $(( $(date +%s) - T0 )) s"
