"""Attention forward / backward device times at the cfg1 shape (B=64, T=149, 12 heads), the cfg4 shape (B=32, T=249,
16 heads) and a single-tile shape, measured on a replayed CUDA graph (no host launch cost inside the timed region)
over four rotating input sets (4 x 59 MB > L2) -- development aid.   usage: python tools/time_attention.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from w2v2_speaker_b200 import ops

NSETS, REPS = (1, 12) if "--hot" in sys.argv else (4, 3)      # --hot: one input set, L2-resident (as after the QKV GEMM)


def graph_time(name, fns):
    """fns: one closure per input set.  Captures REPS x len(fns) calls, replays 5 times, reports per-call time."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            for f in fns:
                f()
    ts = []
    for _ in range(5):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / (REPS * len(fns)))
    ts.sort()
    print(f"{name:44s} median {ts[len(ts)//2]*1e3:8.1f} us   min {ts[0]*1e3:8.1f} us", flush=True)


for B, T, H, heads in ((64, 149, 768, 12), (32, 249, 1024, 16)):
    g = torch.Generator().manual_seed(0)
    sets = []
    for i in range(NSETS):
        qkv = torch.randn(B * T, 3 * H, generator=g).cuda().half()
        qkv[:, :H] *= 0.35
        d_o = torch.randn(B * T, H, generator=g).cuda().half()
        sets.append((qkv, d_o))
    dbias = torch.zeros(3 * H, device="cuda")
    for p in (0.0, 0.1):
        outs = [ops.attention(q, B, T, H, heads, want_lse=True, drop_p=p, drop_seed=5) for q, _ in sets]
        tag = f"B={B} T={T} heads={heads} drop={p}"
        graph_time("attention fwd  " + tag,
                   [lambda q=q: ops.attention(q, B, T, H, heads, want_lse=True, drop_p=p, drop_seed=5) for q, _ in sets])
        graph_time("attention bwd  " + tag,
                   [lambda q=q, d=d, o=o: ops.attention_bwd(q, o[0], d, o[1], B, T, H, heads, drop_p=p, drop_seed=5,
                                                            qscale=0.125, dbias=dbias) for (q, d), o in zip(sets, outs)])
