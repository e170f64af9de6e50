#!/bin/bash
# compute-sanitizer over the kernels added at the end of round 2: key-tiled attention with dropout, attention backward per
# key block, LayerNorm backward ring, weight preparation on 64 x 64 tiles
mkdir -p gpurun_out
SEL='layernorm_bwd_from_output or refresh or (beyond_256 and 301) or two_cta'
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_attention_persist.py -q -m gpu -k "$SEL" > gpurun_out/r2t_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -a "passed\|failed\|ERROR SUMMARY" gpurun_out/r2t_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -q -m gpu -k "layernorm_bwd_from_output or refresh" > gpurun_out/r2t_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -a "passed\|failed\|ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/r2t_racecheck.log | tail -3
