#!/bin/bash
# the -lv60 / XLSR variant: evaluation forward and training step against the oracle
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_round2.py -x -q -m gpu -k "stable or large" -s > gpurun_out/lv60.log 2>&1
echo "lv60 rc=$?"; grep -a "^E  \|passed\|failed\|worst" gpurun_out/lv60.log | cut -c1-300 | tail -12
