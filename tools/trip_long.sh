#!/bin/bash
# the whole GPU suite (no -x) after the long-sequence training changes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/long_all.log 2>&1
echo "all rc=$?"; tail -6 gpurun_out/long_all.log
