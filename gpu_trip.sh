mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q -s -k "unfrozen" > gpurun_out/t33_cnn.log 2>&1; grep -E "unfrozen-CNN|passed|failed|Error|error" gpurun_out/t33_cnn.log | cut -c1-1500 | head -20
echo "cnn tests done $(( $(date +%s) - T0 )) s"
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/t33_tests.log 2>&1; tail -5 gpurun_out/t33_tests.log
echo "all done $(( $(date +%s) - T0 )) s"
