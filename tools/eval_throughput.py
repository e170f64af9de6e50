"""Evaluation throughput on a VoxCeleb-like mix of utterance lengths: embeddings per second when the test
utterances are embedded one at a time (the reference's test loop) vs in length buckets (ragged.plan_buckets + the
length-masked kernels).  Prints one JSON line per mode.   usage: python tools/eval_throughput.py [n_utterances]"""
import json
import math
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
m = bench.build_module(dev, train=False)
g = torch.Generator().manual_seed(0)
# VoxCeleb1-O: 4 s minimum, mean ~8.2 s, long tail -> 4 s + lognormal
secs = [min(4.0 + math.exp(1.0 + 0.8 * float(torch.randn(1, generator=g))), 40.0) for _ in range(n)]
utts = []
for s in secs:
    x = torch.randn(int(s * 16000), generator=g)
    utts.append(((x - x.mean()) / (x.std() + 1e-5)).to(dev))
total_s = sum(secs)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


with torch.no_grad():
    t1, e1 = timed(lambda: torch.cat([m.compute_speaker_embedding(u[None]).reshape(1, -1) for u in utts]))
    res = {}
    for mb, pad in ((16, 0.1), (32, 0.15), (64, 0.25)):
        t2, e2 = timed(lambda: m.compute_speaker_embeddings_ragged(utts, max_batch=mb, max_pad_fraction=pad))
        err = ((e2 - e1).norm(dim=1) / e1.norm(dim=1)).max().item()
        res[(mb, pad)] = (t2, err)
print(json.dumps({"mode": "one utterance per forward (reference test loop)", "utterances": n, "audio_s": round(total_s, 1),
                  "mean_s": round(total_s / n, 2), "utt_per_s": round(n / t1, 1), "audio_s_per_s": round(total_s / t1, 1)}))
for (mb, pad), (t2, err) in res.items():
    print(json.dumps({"mode": f"length buckets (max_batch {mb}, max padding {pad})", "utterances": n, "utt_per_s": round(n / t2, 1),
                      "audio_s_per_s": round(total_s / t2, 1), "speedup": round(t1 / t2, 2), "max_rel_diff_vs_single": err}))
