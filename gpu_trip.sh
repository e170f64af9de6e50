mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -x -q -k "attention or full_utterance or posconv" > gpurun_out/t39_tests.log 2>&1; tail -12 gpurun_out/t39_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
