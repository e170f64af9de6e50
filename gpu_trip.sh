mkdir -p gpurun_out
i=0
for k in "test_attention" "test_posconv" "split3"; do
  i=$((i+1))
  timeout -k 5 170 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$k" > gpurun_out/t2_$i.log 2>&1
  echo "group $i ($k) exit $?" | tee -a gpurun_out/t2_summary.txt
  tail -n 25 gpurun_out/t2_$i.log
done
timeout -k 5 300 python -m pytest tests/test_gpu_engine.py -m gpu -q > gpurun_out/t2_engine.log 2>&1
echo "engine exit $?" | tee -a gpurun_out/t2_summary.txt
tail -n 30 gpurun_out/t2_engine.log
timeout -k 5 300 python tools/time_ops.py 64 > gpurun_out/t2_time.log 2>&1
echo "time exit $?"
cat gpurun_out/t2_time.log | tail -40
