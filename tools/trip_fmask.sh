#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_regularise.py -x -q -m gpu > gpurun_out/fmask.log 2>&1
echo "rc=$?"; grep -a "^E  \|passed\|failed" gpurun_out/fmask.log | cut -c1-300 | tail -8
