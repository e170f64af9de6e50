mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_gpu_split_path.py -m gpu -q -k "paired_input_model_matches_oracle" -s > gpurun_out/t52_paired.log 2>&1; tail -6 gpurun_out/t52_paired.log | cut -c1-300
