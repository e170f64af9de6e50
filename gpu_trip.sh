mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
i=0
for k in "test_gemm_f16 and 128-256-64-0-False-False" "test_gemm_f16" "test_conv1d or rejects" "test_conv0" "test_layernorm or stat_pool or asp" "softmax or split3"; do
  i=$((i+1))
  timeout -k 5 170 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$k" > gpurun_out/t1_$i.log 2>&1
  echo "group $i ($k) exit $?" | tee -a gpurun_out/t1_summary.txt
  tail -n 12 gpurun_out/t1_$i.log
done
