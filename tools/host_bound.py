"""Is the step host-bound?  Host enqueue time per step (no synchronisation inside the loop) vs device time.
usage: python tools/host_bound.py [--mode train|forward]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="train"); a = ap.parse_args()
from w2v2_speaker_b200.synthetic import synthetic_batch as make_inputs
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
K = 10
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); s.record()
for _ in range(K): step()
e.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"{a.mode}: host enqueue {1e3*(t1-t0)/K:.2f} ms/step, device {s.elapsed_time(e)/K:.2f} ms/step, wall {1e3*(t2-t0)/K:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
