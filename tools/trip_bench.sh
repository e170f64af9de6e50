#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/final_bench.log 2>&1
python - <<'PY'
import json
l = [x for x in open("gpurun_out/final_bench.log") if x.startswith("{")]
d = json.loads(l[-1]); print("train", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "gemm", round(d["roofline"]["ms_per_step"], 3), round(d["roofline"]["frac"], 4), "conv0", round(d["roofline_hbm"]["ms_per_step"], 3), round(d["roofline_hbm"]["frac"], 3), d["clocks"])
PY
tail -3 gpurun_out/final_bench.log | cut -c1-300
