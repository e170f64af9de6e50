#!/bin/bash
# LayerNorm backward with rows streamed through a shared-memory ring: parity, A/B of the train step, kernel time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -x -q -m gpu > gpurun_out/lnb_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/lnb_tests.log
for v in 1 0; do
  W2V2_LNB_RING=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/lnb_bench_$v.log 2>&1
  python - "$v" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/lnb_bench_%s.log" % sys.argv[1]) if x.startswith("{")]
d = json.loads(l[-1]); print("LNB_RING=%s train" % sys.argv[1], round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4))
PY
done
for v in 1 0; do
W2V2_LNB_RING=$v timeout 300 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:layernorm_bwd -c 6 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -a "layernorm_bwd\|gpu__time\|warps_active\|dram_throughput" | cut -c1-150 | tail -8
done
