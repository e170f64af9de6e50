"""Single-op microbenchmarks (attention / posconv / conv0 / layernorm) for ncu captures.
usage: python tools/op_bench.py <op> [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from w2v2_speaker_b200 import ops
op = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T, H = 149, 768
dev = "cuda"
if op == "attention":
    qkv = torch.randn(B * T, 3 * H, device=dev).half()
    fn = lambda: ops.attention(qkv, B, T, H, 12)
elif op == "posconv":
    x = torch.randn(B, T, H, device=dev).half()
    v = torch.randn(H, 48, 128, device=dev) * 0.01
    w = ops.posconv_fold_weight(v, v.pow(2).sum(dim=(0, 1)).sqrt(), 16, ops.posconv_taps_per_mma(T, H, 16))
    b = torch.zeros(H, device=dev)
    fn = lambda: ops.posconv(x, w, b, 16, 128)
elif op == "conv0":
    wav = torch.randn(B, 48000, device=dev)
    w0 = torch.randn(512, 10, device=dev) * 0.4
    g = torch.ones(512, device=dev); bt = torch.zeros(512, device=dev)
    fn = lambda: ops.conv0_gn_gelu(wav, w0, g, bt)
elif op == "layernorm":
    x = torch.randn(B * T, H, device=dev); r = torch.randn(B * T, H, device=dev)
    g = torch.ones(H, device=dev); bt = torch.zeros(H, device=dev)
    fn = lambda: ops.layernorm(x, g, bt, bias=bt, residual=r)
for _ in range(3):
    fn()
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record(); fn(); e.record(); torch.cuda.synchronize()
print(op, f"{s.elapsed_time(e)*1e3:.1f} us")
