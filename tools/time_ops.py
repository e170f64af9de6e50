"""Per-kernel timings at the cfg1 shapes (B=64, 3 s, base) with CUDA events -- development aid.
usage: python tools/time_ops.py [B]"""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from w2v2_speaker_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(name, fn, flops=0.0, bytes_=0.0, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    msg = f"{name:34s} {t*1e3:9.1f} us"
    if flops:
        msg += f"  {flops / t / 1e9:8.1f} TFLOP/s"
    if bytes_:
        msg += f"  {bytes_ / t / 1e6:8.1f} GB/s"
    print(msg, flush=True)
    return t


N = 48000
Ls = [9599, 4799, 2399, 1199, 599, 299, 149]
T, H, FF, C = 149, 768, 3072, 512
M = B * T
total = 0.0
wav = torch.randn(B, N, device=dev)
w0 = torch.randn(C, 10, device=dev) * 0.4
g = torch.ones(C, device=dev); bta = torch.zeros(C, device=dev)
total += timeit("conv0+GN+GELU", lambda: ops.conv0_gn_gelu(wav, w0, g, bta), bytes_=B * (N * 4 * 2 + Ls[0] * C * 2))
x = torch.randn(B, Ls[0], C, device=dev).half()
ks = [3, 3, 3, 3, 2, 2]
for i, k in enumerate(ks):
    w = (torch.randn(C, k * C, device=dev) * math.sqrt(2.0 / (k * C))).half()
    Lo = Ls[i + 1]
    fl = 2.0 * B * Lo * C * C * k
    by = B * (Ls[i] + Lo) * C * 2
    total += timeit(f"conv{i+1} k={k} L={Lo}", lambda: ops.conv1d_cl_f16(x, w, k, 2, 1), fl, by)
    x = ops.conv1d_cl_f16(x, w, k, 2, 1)
a512 = torch.randn(M, C, device=dev)
gg = torch.ones(C, device=dev); bb = torch.zeros(C, device=dev)
total += timeit("LN512", lambda: ops.layernorm(a512, gg, bb, want32=False), bytes_=M * C * 6)
h16 = torch.randn(M, H, device=dev).half()
h32 = torch.randn(M, H, device=dev)
bH = torch.zeros(H, device=dev); gH = torch.ones(H, device=dev)
def mk(n, k): return (torch.randn(n, k, device=dev) * 0.02).half()
wproj = mk(H, C); a16 = a512.half()
total += timeit("proj GEMM 512->768", lambda: ops.gemm_f16(a16, wproj, bH, 0, torch.float32), 2.0 * M * H * C)
total += timeit("cast f16 [M,H]", lambda: ops.cast_f16(h32), bytes_=M * H * 6)
v = torch.randn(H, H // 16, 128, device=dev) * 0.01
gn = v.pow(2).sum(dim=(0, 1)).sqrt()
pw = ops.posconv_fold_weight(v, gn, 16, ops.posconv_taps_per_mma(T, H, 16))
x16 = h16.view(B, T, H)
total += timeit("posconv", lambda: ops.posconv(x16, pw, bH, 16, 128), 2.0 * M * H * (H // 16) * 128)
total += timeit("LN768 (+bias+res)", lambda: ops.layernorm(h32, gH, bH, bias=bH, residual=h32), bytes_=M * H * (4 + 4 + 4 + 2))
wqkv = mk(3 * H, H); b3 = torch.zeros(3 * H, device=dev)
lay = 0.0
lay += timeit("QKV GEMM", lambda: ops.gemm_f16(h16, wqkv, b3, 0, torch.float16), 2.0 * M * 3 * H * H)
qkv = ops.gemm_f16(h16, wqkv, b3, 0, torch.float16)
lay += timeit("attention", lambda: ops.attention(qkv, B, T, H, 12), 4.0 * B * 12 * T * T * 64)
wo = mk(H, H)
lay += timeit("out_proj GEMM", lambda: ops.gemm_f16(h16, wo, None, 0, torch.float32), 2.0 * M * H * H)
lay += timeit("LN768", lambda: ops.layernorm(h32, gH, bH, bias=bH, residual=h32), bytes_=M * H * 14)
w1 = mk(FF, H); b1 = torch.zeros(FF, device=dev)
lay += timeit("FFN1 GEMM+GELU", lambda: ops.gemm_f16(h16, w1, b1, 1, torch.float16), 2.0 * M * FF * H)
f1 = ops.gemm_f16(h16, w1, b1, 1, torch.float16)
w2 = mk(H, FF)
lay += timeit("FFN2 GEMM", lambda: ops.gemm_f16(f1, w2, None, 0, torch.float32), 2.0 * M * FF * H)
lay += timeit("LN768", lambda: ops.layernorm(h32, gH, bH, bias=bH, residual=h32), bytes_=M * H * 14)
print(f"one transformer layer: {lay*1e3:.1f} us  -> x12 = {12*lay:.3f} ms")
total += 12 * lay
hb = h32.view(B, T, H)
total += timeit("mean pool", lambda: ops.stat_pool(hb, 0), bytes_=M * H * 4)
print(f"estimated forward total: {total:.3f} ms -> {B / total * 1e3:.0f} utt/s")
