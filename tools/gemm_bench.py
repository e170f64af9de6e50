"""Microbenchmark of the tcgen05 tap-GEMM on the cfg1 shapes.  usage: python tools/gemm_bench.py [shape ...]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from w2v2_speaker_b200 import ops

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
B, T = 64, 149
M = B * T
SH = {  # name: (M, N, K, act, f32out, bias)
    "qkv": (M, 2304, 768, 0, False, True),
    "oproj": (M, 768, 768, 0, True, False),
    "ffn1": (M, 3072, 768, 1, False, True),
    "ffn2": (M, 768, 3072, 0, True, False),
    "big": (8192, 8192, 8192, 0, False, False),
}
names = sys.argv[1:] or ["qkv", "oproj", "ffn1", "ffn2", "conv1", "big"]
iters = int(os.environ.get("ITERS", 5))
for nm in names:
    if nm == "conv1":
        x = torch.randn(B, 9599, 512, device=dev).half()
        w = (torch.randn(512, 1536, device=dev) * 0.03).half()
        fn = lambda: ops.conv1d_cl_f16(x, w, 3, 2, 1)
        fl = 2.0 * B * 4799 * 512 * 1536
    else:
        m, n, k, act, f32, ub = SH[nm]
        a = torch.randn(m, k, device=dev).half()
        w = (torch.randn(n, k, device=dev) / math.sqrt(k)).half()
        bias = torch.zeros(n, device=dev) if ub else None
        out = torch.empty(m, n, dtype=torch.float32 if f32 else torch.float16, device=dev)
        fn = lambda: ops.gemm_f16(a, w, bias, act, out.dtype, out=out)
        fl = 2.0 * m * n * k
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    print(f"{os.environ.get('W2V2_GEMM_EPI','direct'):6s} {nm:6s} {t*1e3:9.1f} us  {fl/t/1e9:8.1f} TFLOP/s", flush=True)
