#!/bin/bash
# attention forward with two resident CTAs per SM: which kernel runs, and timings with the switch on and off
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from w2v2_speaker_b200 import ops
B, T, H, heads = 64, 149, 768, 12
q = torch.randn(B * T, 3 * H).cuda().half()
for _ in range(3):
    o = ops.attention(q, B, T, H, heads, want_lse=True, drop_p=0.1, drop_seed=5)
torch.cuda.synchronize()
PY
for v in 1 0; do
  echo "== W2V2_ATTN_2CTA=$v"
  W2V2_ATTN_2CTA=$v timeout 200 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none python /tmp/one.py 2>&1 | grep -i "attention\|gpu__time\|occupancy\|registers\|warps_active" | tail -12
done
