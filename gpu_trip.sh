mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_regularise.py -m gpu -q 2>&1 | tail -15 | cut -c1-300
timeout -k 5 600 python -m pytest tests/test_gpu_training.py -m gpu -q 2>&1 | tail -15 | cut -c1-300
timeout -k 5 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_modules.py -m gpu -q 2>&1 | tail -4
