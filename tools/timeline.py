"""Kernel timeline of bench steps from CUPTI (torch.profiler): per-kernel busy time, idle gaps between
consecutive kernels on the compute stream, and the per-kernel-name totals.  usage:
    python tools/timeline.py [--mode train|forward] [--steps 3]"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="train"); ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
from oracle.params import make_inputs
from w2v2_speaker_b200 import trainer as T
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
train = a.mode == "train"
m = bench.build_module(dev, train, True)
tr = T.FlatAdamTrainer(m, lr=1e-4) if train else None
wav, lab = make_inputs(64, 48000, 5994, seed=1)
wav = wav[:, None, :].contiguous().to(dev); lab = lab.to(dev)
def step():
    if train: return tr.step(wav, lab)
    with torch.no_grad():
        e, p = m(wav); return m.loss_fn(p, lab)
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda x: x[0])
busy = sum(e - s for s, e, _ in ks)
span = ks[-1][1] - ks[0][0]
gaps = []
last_end = ks[0][1]
for s, e, n in ks[1:]:
    if s > last_end: gaps.append((s - last_end, n))
    last_end = max(last_end, e)
idle = sum(g for g, _ in gaps)
print(f"# {a.mode}: {len(ks) // a.steps} kernels/step, span {span / a.steps / 1e3:.3f} ms/step, busy {busy / a.steps / 1e3:.3f} ms/step, "
      f"idle between kernels {idle / a.steps / 1e3:.3f} ms/step ({len(gaps) // a.steps} gaps, mean {idle / max(1, len(gaps)):.2f} us)")
agg = collections.OrderedDict()
for s, e, n in ks:
    n = n.split("(")[0].replace("void ", "")[:64]
    x = agg.setdefault(n, [0, 0.0]); x[0] += 1; x[1] += e - s
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{n:66s} n={c // a.steps:4d} {t / a.steps:9.1f} us {100 * t / busy:5.1f}%  ({t / c:7.1f} us each)")
gb = collections.Counter()
for g, n in gaps: gb[n.split("(")[0].replace("void ", "")[:64]] += g
print("# idle time in front of:")
for n, t in gb.most_common(12): print(f"  {n:64s} {t / a.steps:8.1f} us/step")
