#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final_cfg4.log 2>&1
python - <<'PY'
import json
l = [x for x in open("gpurun_out/final_cfg4.log") if x.startswith("{")]
d = json.loads(l[-1]); print("cfg4 train", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4))
PY
timeout 100 python bench.py --mode forward --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_fwd.log 2>&1
python - <<'PY'
import json
l = [x for x in open("gpurun_out/final_fwd.log") if x.startswith("{")]
d = json.loads(l[-1]); print("cfg1 forward", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), round(d["value"]))
PY
