#!/bin/bash
# usage: tools/trip.sh <timeout_s> <logfile> [gpus]  -- builds (aborts on failure), then ships tools/gpu_trip.sh to a B200 box
set -e
cd /root/repo
python -m w2v2_speaker_b200.build > /tmp/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 /tmp/build.log; exit 1; }
python -c "from w2v2_speaker_b200 import _lib; _lib.load()" || { echo "LOAD FAILED"; exit 1; }
if [ -n "$3" ]; then
  exec /usr/local/graft/bin/gpurun --gpus "$3" --timeout "$1" -- 'bash tools/gpu_trip.sh' > "$2" 2>&1
else
  exec /usr/local/graft/bin/gpurun --timeout "$1" -- 'bash tools/gpu_trip.sh' > "$2" 2>&1
fi
