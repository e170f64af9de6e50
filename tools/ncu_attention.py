"""One forward and one backward attention launch at the cfg1 shape with dropout (for `ncu -k regex:attention`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from w2v2_speaker_b200 import ops
B, T, H, heads = 64, 149, 768, 12
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B * T, 3 * H, generator=g).cuda().half()
qkv[:, :H] *= 0.35
d_o = torch.randn(B * T, H, generator=g).cuda().half()
dbias = torch.zeros(3 * H, device="cuda")
for _ in range(2):
    out, lse = ops.attention(qkv, B, T, H, heads, want_lse=True, drop_p=0.1, drop_seed=5)
    ops.attention_bwd(qkv, out, d_o, lse, B, T, H, heads, drop_p=0.1, drop_seed=5, qscale=0.125, dbias=dbias)
torch.cuda.synchronize()
