#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "attentive or asp or ragged or pool or fixture or ensemble" > gpurun_out/asp.log 2>&1
echo "rc=$?"; grep -a "^E  \|passed\|failed\|FAILED" gpurun_out/asp.log | cut -c1-300 | tail -10
