#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ragged.py -x -q -m gpu > gpurun_out/lv60r.log 2>&1
echo "rc=$?"; grep -a "^E  \|passed\|failed" gpurun_out/lv60r.log | cut -c1-300 | tail -6
