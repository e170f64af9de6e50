"""Per-entry-point device time of one bench step, measured with CUDA events around every C-ABI call
(warm caches, real clocks -- unlike the serialised cold-cache ncu launch list).  usage:
    python tools/step_profile.py [--mode train|forward] [--no-reg] [--reps 3]
Event pairs bracket each call on the launching stream, so a kernel's figure includes whatever the device
was still finishing when it was enqueued; with the device saturated the per-name sums add up to the step."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="train", choices=["train", "forward"])
    ap.add_argument("--no-reg", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from oracle.params import make_inputs
    from w2v2_speaker_b200 import ops, training, engine, trainer as trainer_mod
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    train = args.mode == "train"
    module = bench.build_module(dev, train, not args.no_reg)
    tr = trainer_mod.FlatAdamTrainer(module, lr=1e-4) if train else None
    wav, labels = make_inputs(bench.BATCH, bench.SAMPLES, bench.NUM_SPEAKERS, seed=1234)
    wav = wav[:, None, :].contiguous().to(dev)
    labels = labels.to(dev)

    def step():
        if train:
            return tr.step(wav, labels)
        with torch.no_grad():
            emb, pred = module(wav)
            return module.loss_fn(pred, labels)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    rec = []
    orig = ops.call

    def traced(name, *a):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(name, *a)
        e.record()
        tag = name
        if name == "w2v2_gemm_f16":      # (A, rows, row_stride, batch_stride, batch, ntaps, tap_stride, cin, W, ldw, N, bias, act, out, f32, ...)
            tag = f"w2v2_gemm_f16[N={a[10]},K={a[5] * a[7]},{'f32' if a[14] else 'f16'}{',gelu' if a[12] else ''}]"
        elif name == "w2v2_gemm_wgrad_f16":
            tag = f"w2v2_gemm_wgrad_f16[N={a[5]},K={a[6]}]"
        rec.append((tag, s, e))
        return r

    for mod in (ops, training, engine, trainer_mod):
        if hasattr(mod, "call"):
            mod.call = traced
    ops.call = traced
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.reps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ops.call = orig
    agg = collections.OrderedDict()
    for tag, s, e in rec:
        a = agg.setdefault(tag, [0, 0.0])
        a[0] += 1
        a[1] += s.elapsed_time(e) * 1e3
    tot = sum(a[1] for a in agg.values()) / args.reps
    wall = t0.elapsed_time(t1) * 1e3 / args.reps
    print(f"# {args.mode} step, {len(rec) // args.reps} C-ABI calls per step; bracketed device time {tot / 1e3:.3f} ms of "
          f"{wall / 1e3:.3f} ms per (instrumented) step")
    for tag, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{tag:60s} n={n // args.reps:4d} {t / args.reps:9.1f} us {100 * t / args.reps / tot:5.1f}%  ({t / n:7.1f} us each)")


main()
