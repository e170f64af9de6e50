#!/bin/bash
# end-of-round validation: whole GPU suite, smoke(), one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log | cut -c1-250
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/final_bench.log 2>&1
python - <<'PY'
import json
l = [x for x in open("gpurun_out/final_bench.log") if x.startswith("{")]
d = json.loads(l[-1]); print("train", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "gemm", round(d["roofline"]["ms_per_step"], 3), round(d["roofline"]["frac"], 4), "conv0", round(d["roofline_hbm"]["frac"], 3), d["clocks"]["sm_mhz"])
PY
