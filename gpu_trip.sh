mkdir -p gpurun_out
for k in "test_gemm_wgrad" "dgrad or layernorm_bwd or gelu_bwd or adam" "test_attention_bwd"; do
  timeout -k 5 200 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "$k" > gpurun_out/t7_$RANDOM.log 2>&1
  echo "group ($k) exit $?"; tail -n 14 gpurun_out/t7_*.log | tail -n 14 | cut -c1-300
  rm -f gpurun_out/t7_*.log
done
timeout -k 5 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_modules.py -m gpu -q 2>&1 | tail -5
timeout -k 5 100 python tools/op_bench.py conv0
timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t7_bench.log 2>&1; tail -2 gpurun_out/t7_bench.log | cut -c1-600
