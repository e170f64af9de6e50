#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "freeze_protocol" > gpurun_out/freeze.log 2>&1
echo "freeze rc=$?"; grep -a "^E  \|passed\|failed" gpurun_out/freeze.log | cut -c1-300 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log | cut -c1-300
