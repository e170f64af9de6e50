#!/bin/bash
# weight preparation on 64 x 64 tiles: parity (bit-identical copies), then A/B of the train step and the kernel's time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -x -q -m gpu -k "refresh or trainer or adam" > gpurun_out/prep_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/prep_tests.log
for v in 1 0; do
  W2V2_PREP_V3=$v timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/prep_bench_$v.log 2>&1
  python - "$v" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/prep_bench_%s.log" % sys.argv[1]) if x.startswith("{")]
d = json.loads(l[-1]); print("PREP_V3=%s train" % sys.argv[1], round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4))
PY
done
W2V2_PREP_V3=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:prepare_weights -c 4 python bench.py --steps 2 --warmup 2 2>&1 | grep -a "prepare_weights\|gpu__time\|dram__bytes" | tail -9
