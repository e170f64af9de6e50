#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: utterances/sec, 3 s @ 16 kHz, wav2vec2-base).

    python bench.py --gpus N --steps K --warmup W [--mode train|forward]   # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...                # the reference path on the host CPU cores

Workload = configs[1] of BASELINE.json: wav2vec2-base + mean pool + Linear(768->5994) + CE on synthetic
3 s utterances, 64 per GPU, driven through the public module API (w2v2_speaker_b200/speaker_module.py).
--workload cfg2|cfg3|cfg4 selects the other BASELINE.json configurations (attentive pooling + AAM, mean+std +
AAM, wav2vec2-large 5 s) for both arms; cfg1 stays the default the metric is quoted on.

  --mode train   (default) one step = forward + backward + gradient all-reduce (N > 1) + Adam update
                 (w2v2_speaker_b200/trainer.py); CNN feature extractor frozen and dropout 0.1 (feature
                 projection / hidden / attention), LayerDrop 0.05, SpecAugment 0.05 switched ON -- the reference's
                 default training configuration (R:config/network/wav2vec2_fc.yaml, R:src/models/wav2vec2.py:83-94);
                 --no-reg sets every regularisation probability to 0.
  --mode forward one step = eval-mode embedding + logits + softmax/CE.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM), `e2e` = the
same call with pinned-host inputs (H2D inside the timed region) and loss / arg-max (and the embedding in
forward mode) read back to the host every step.  `roofline` aggregates every tensor-core GEMM launch
(forward tap-GEMM + dgrad + wgrad) of one step: FLOPs from the launch arguments / CUDA-event time;
`roofline_hbm` is the HBM-bound conv0+GroupNorm+GELU stage; `cpu_baseline` times the CPU oracle (a
torch-fp32 port of the reference path, same mode) on a bounded sample.
Multi-GPU: one process per GPU (torchrun), pure data parallel over utterances (weak scaling); train mode
all-reduces the flat fp32 gradient over NCCL, forward mode has no collective.  Time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "utterances/sec (3 s@16 kHz, w2v2-base)"
NUM_SPEAKERS = 5994
SAMPLES = 48000
BATCH = 64
DEFAULT_MODE = "train"

# BASELINE.json configs[1..4]; cfg1 is the one the metric is quoted on (the default).  The others are selectable
# with --workload so that every configuration has a measured line (profiles/).
WORKLOADS = {
    "cfg1": dict(hf_id="facebook/wav2vec2-base", arch="base", pooling="mean", loss="ce", samples=48000, batch=64,
                 desc="cfg1: wav2vec2-base + mean-pool + CE(5994), 3 s@16 kHz, batch 64 per GPU"),
    "cfg2": dict(hf_id="facebook/wav2vec2-base", arch="base", pooling="attentive", loss="aam", samples=48000, batch=64,
                 desc="cfg2: wav2vec2-base + attentive-stat-pool + AAM-softmax(m=0.2, s=30), 3 s@16 kHz, batch 64 per GPU"),
    "cfg3": dict(hf_id="facebook/wav2vec2-base", arch="base", pooling="mean+std", loss="aam", samples=48000, batch=64,
                 desc="cfg3: wav2vec2-base + mean+std pool (reference default) + AAM-softmax(m=0.2, s=30), 3 s@16 kHz, "
                      "batch 64 per GPU (global batch 512 at 8 GPUs)"),
    "cfg4": dict(hf_id="facebook/wav2vec2-large", arch="large", pooling="mean+std", loss="ce", samples=80000, batch=32,
                 desc="cfg4: wav2vec2-large (24 layers) + mean+std pool + CE(5994), 5 s@16 kHz, batch 32 per GPU"),
}
WL = WORKLOADS["cfg1"]
TRAIN_CNN = False


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def workload_name(mode: str, reg: bool = True) -> str:
    if mode == "train":
        return (WL["desc"] + ", TRAIN step = forward + backward + grad all-reduce + Adam; " +
                ("CNN UNFROZEN, " if TRAIN_CNN else "CNN frozen, ") +
                ("dropout 0.1 / LayerDrop 0.05 / SpecAugment 0.05 on (reference defaults)" if reg else
                 "regularisation probabilities 0"))
    return WL["desc"] + ", eval forward (embedding + logits + softmax/loss)"


# ------------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(", ") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows)
            out["samples"] = len(rows)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][2])
                out["power_w_max"] = max(float(r[3]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            seen = set()
            for r in rows:
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower() == "active":
                        seen.add(n)
            out["reasons"] = sorted(seen)
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (torch-fp32 port of the reference path) on the host cores


def cpu_reference_step_fn(batch: int, mode: str):
    """One step of the reference path on the host CPU.  The encoder is the reference's own third-party code -- the
    `transformers.Wav2Vec2Model` it wraps (R:src/models/wav2vec2.py:37-53; present in this image, unlike the reference
    checkout, which cannot travel to the GPU box), in train mode with the reference's regularisation defaults (dropout,
    LayerDrop and SpecAugment are HF's own), eager attention, fp32; pooling and loss heads are the oracle's restatement
    of R:src/layers/pooling.py / R:src/optim/loss (a few lines of torch each); optimizer = torch.optim.Adam
    (R:config/optim/algo/adam.yaml).  Falls back to the all-oracle port if transformers cannot build the model."""
    from oracle import w2v2_oracle as O
    from oracle import params as OP
    arch = OP.LARGE if WL["arch"] == "large" else OP.BASE
    p = OP.make_params(arch, seed=0)
    E = arch.hidden * (1 if WL["pooling"] == "mean" else 2)
    hp = OP.make_head_params(E, NUM_SPEAKERS, seed=1)
    asp = OP.make_asp_params(arch.hidden, seed=2) if WL["pooling"] == "attentive" else None
    wav, labels = OP.make_inputs(batch, WL["samples"], NUM_SPEAKERS, seed=1234)
    train = mode == "train"

    hf = None
    try:
        from transformers import Wav2Vec2Config, Wav2Vec2Model
        size = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096) \
            if WL["arch"] == "large" else {}
        cfg = Wav2Vec2Config(activation_dropout=0.0, attention_dropout=0.1, feat_proj_dropout=0.1, hidden_dropout=0.1,
                             layerdrop=0.05, mask_time_prob=0.05, mask_time_length=10, mask_feature_prob=0.0, **size)
        cfg._attn_implementation = "eager"
        hf = Wav2Vec2Model(cfg)
        res = hf.load_state_dict(p, strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        hf.train(train)
        hf.feature_extractor.requires_grad_(False)          # completely_freeze_feature_extractor: true
    except Exception as e:                                   # pragma: no cover - depends on the installed transformers
        print(f"[bench] transformers model unavailable ({e}); timing the oracle port", file=sys.stderr)
        hf = None
    cpu_reference_step_fn.kind = "HF Wav2Vec2Model (eager, the reference's encoder) + oracle heads" if hf is not None \
        else "oracle/w2v2_oracle.py"

    def embed():
        h = hf(wav).last_hidden_state if hf is not None else O.wav2vec2_forward(wav, p, arch)
        if WL["pooling"] == "attentive":
            return O.attentive_stat_pool(h, asp, training=train)
        return O.mean_pool(h) if WL["pooling"] == "mean" else O.mean_std_pool(h)

    def head(emb):
        if WL["loss"] == "aam":
            return O.aam_softmax(emb, hp["aam.fc_weights"], labels, 0.2, 30.0)
        return O.cross_entropy_head(emb, hp["fc.weight"], hp["fc.bias"], labels)

    if train:
        if hf is not None:
            tr = [q for q in hf.parameters() if q.requires_grad]
        else:
            O.TRAIN_REG = {"feat": 0.1, "hidden": 0.1, "attn": 0.1, "layerdrop": 0.05}     # reference defaults
            tr = [p[k] for k in p if not k.startswith("feature_extractor") and k != "masked_spec_embed"]
        tr += [hp["aam.fc_weights"]] if WL["loss"] == "aam" else [hp["fc.weight"], hp["fc.bias"]]
        if asp is not None:
            tr += [v for k, v in asp.items() if "running" not in k]
        for t in tr:
            t.requires_grad_(True)
        opt = torch.optim.Adam(tr, lr=1e-4)

        def step():
            opt.zero_grad()
            _, loss, _ = head(embed())
            loss.backward()
            opt.step()
            return float(loss.detach())
        return step

    O.TRAIN_REG = None

    def step():
        with torch.no_grad():
            _, loss, _ = head(embed())
        return float(loss)
    return step


def pick_cpu_threads() -> int:
    """torch CPU ops do not always scale to every hardware thread of a large host: probe a few thread
    counts on a small batch and keep the fastest (the baseline gets its best configuration)."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (cores, cores // 2, 64, 32, 16) if 1 <= c <= cores}, reverse=True)
    if len(cands) == 1:
        return cands[0]
    probe = cpu_reference_step_fn(2, "forward")
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        probe()
        t0 = time.perf_counter()
        probe()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def time_cpu(batch: int, steps: int, warmup: int, mode: str):
    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    step = cpu_reference_step_fn(batch, mode)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores


def run_reference_arm(args):
    """`--impl reference`: the reference path on the host cores, same workload / mode / step counts as the product arm.
    The per-step batch is the configuration's own (64 at cfg1) unless W + K steps of it would take more than ~4 minutes
    on this host, in which case the largest batch that fits is used and reported."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    probe_b = 4
    probe = cpu_reference_step_fn(probe_b, args.mode)
    probe()
    t0 = time.perf_counter()
    probe()
    per_utt = (time.perf_counter() - t0) / probe_b
    budget = float(os.environ.get("W2V2_REF_ARM_SECONDS", "240"))
    batch = int(max(1, min(args.batch, budget / ((steps + warmup) * per_utt))))
    step = cpu_reference_step_fn(batch, args.mode)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    uts, ms = batch * steps / dt, dt / steps * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": uts, "unit": "utt/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.mode), "batch_per_step": batch, "device": "host CPU", "mode": args.mode},
        "cpu_baseline": {"value": uts, "unit": "utt/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {batch} utterances of {WL['samples'] / 16000:g} s ({cpu_reference_step_fn.kind}, "
                                   f"torch fp32, {cores} threads, mode {args.mode})"},
        "e2e": {"value": uts, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm


def build_module(device, train: bool, reg: bool = True, train_cnn: bool = False):
    from w2v2_speaker_b200.optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    torch.manual_seed(0)
    kw = dict(activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
              mask_time_prob=0.0, mask_feature_prob=0.0) if (train and not reg) else {}
    cfg = Wav2vec2FCModuleConfig(wav2vec_hunggingface_id=WL["hf_id"], stat_pooling_type=WL["pooling"],
                                 test_stat_pooling_type=WL["pooling"], **kw)
    loss = CrossEntropyLoss if WL["loss"] == "ce" else (lambda: AngularAdditiveMarginSoftMaxLoss(1, 1, margin=0.2, scale=30))
    m = Wav2vec2FCModule(cfg, NUM_SPEAKERS, loss).to(device)
    if train:
        m.train()
        m.wav2vec.model.feature_extractor.requires_grad_(train_cnn)   # R:.../wav2vec2_fc.py:346-347 (default: frozen)
    else:
        m.eval()
    return m


def instrumented(step_fn, lib):
    """One extra (untimed) step with CUDA events around every tensor-core GEMM launch (recorded inside the
    library, w2v2_gemm_profile_*: most launches are issued by the native per-layer schedules, not from python)
    and around the conv0 stage; FLOPs come from the launch arguments."""
    import ctypes
    from w2v2_speaker_b200 import ops
    rec = []
    orig_call = ops.call

    def traced(name, *a):
        if name in ("w2v2_conv0_gn_gelu", "w2v2_conv0_raw", "w2v2_conv0_gn_lens", "w2v2_conv0_gn_ex"):
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig_call(name, *a)
            e.record()
            rec.append((s, e))
            return r
        return orig_call(name, *a)
    ops.call = traced
    ms, fl, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
    try:
        lib.w2v2_gemm_profile_start()
        step_fn()
    finally:
        lib.w2v2_gemm_profile_stop(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
        ops.call = orig_call
    torch.cuda.synchronize()
    return {"gemm_ms": ms.value, "gemm_flops": fl.value, "gemm_launches": n.value,
            "conv0": [s.elapsed_time(e) for s, e in rec]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=DEFAULT_MODE, choices=["train", "forward"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration (cfg1 = the one the metric is quoted on)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reg", action="store_true", help="train mode without dropout / LayerDrop / SpecAugment")
    ap.add_argument("--input", default="f32", choices=["f32", "int16"],
                    help="f32: normalised float waveforms (what the reference's data pipeline hands the model); int16: raw 16-bit "
                         "PCM, the input normaliser folded into conv layer 0 on the device (half the H2D bytes, SURVEY 8f-2)")
    ap.add_argument("--train-cnn", action="store_true",
                    help="train mode with the CNN feature extractor unfrozen (completely_freeze_feature_extractor: false)")
    args = ap.parse_args()
    global WL, TRAIN_CNN
    WL = WORKLOADS[args.workload]
    TRAIN_CNN = bool(args.train_cnn)
    if args.batch is None:
        args.batch = WL["batch"]
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the sm_100a path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from w2v2_speaker_b200 import _lib
    lib = _lib.load()
    from w2v2_speaker_b200.synthetic import synthetic_batch as make_inputs
    B = args.batch
    K, W = args.steps, max(3, args.warmup)
    train = args.mode == "train"
    module = build_module(dev, train, not args.no_reg, args.train_cnn)
    trainer = None
    if train:
        from w2v2_speaker_b200.trainer import FlatAdamTrainer
        trainer = FlatAdamTrainer(module, lr=1e-4)
    wav_cpu, labels_cpu = make_inputs(B, WL["samples"], NUM_SPEAKERS, seed=1234 + rank)
    if args.input == "int16":        # raw PCM: the synthetic waveform at a speech-like level, quantised to 16 bits
        wav_cpu = (wav_cpu * 3000.0).round().clamp(-32768, 32767).to(torch.int16)
    wav_pin = wav_cpu[:, None, :].contiguous().pin_memory()          # [B,1,N] as the reference batches
    labels_pin = labels_cpu.pin_memory()
    wav_dev = wav_pin.to(dev)
    labels_dev = labels_pin.to(dev)

    def run(w, l):
        if train:
            loss, prob = trainer.step(w, l)
            return None, loss, prob
        with torch.no_grad():
            emb, pred = module(w)
            loss, prob = module.loss_fn(pred, l)
        return emb, loss, prob

    def step_device():
        return run(wav_dev, labels_dev)

    emb_dim = (1024 if WL["arch"] == "large" else 768) * (1 if WL["pooling"] == "mean" else 2)
    emb_host = torch.empty(B, emb_dim, dtype=torch.float32).pin_memory()
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    arg_host = torch.empty(B, dtype=torch.int64).pin_memory()

    # End-to-end step: the batch comes from pinned host memory and the result (loss / arg-max, and the embedding in
    # forward mode) is read back every step.  Like a data loader with pinned memory would, the upload of the NEXT
    # batch is issued on a copy stream before this step's kernels are enqueued, so it runs under them; every step
    # still uploads exactly one batch and ends with the host holding that step's result.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [dict(w=torch.empty_like(wav_dev), l=torch.empty_like(labels_dev),
                  ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    state = {"i": 0, "primed": False}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot["free"])          # the step that last used this slot has consumed it
            slot["w"].copy_(wav_pin, non_blocking=True)
            slot["l"].copy_(labels_pin, non_blocking=True)
            slot["ready"].record(copy_stream)

    def step_e2e():
        cur = torch.cuda.current_stream()
        i = state["i"]
        if not state["primed"]:
            for sl in slots:
                sl["free"].record(cur)
            upload(slots[i % 2])                          # nothing was prefetched for the first step
            state["primed"] = True
        sl = slots[i % 2]
        upload(slots[(i + 1) % 2])                        # next step's batch, under this step's compute
        cur.wait_event(sl["ready"])
        emb, loss, prob = run(sl["w"], sl["l"])
        sl["free"].record(cur)
        if emb is not None:
            emb_host.copy_(emb, non_blocking=True)
        loss_host.copy_(loss.view(1), non_blocking=True)
        arg_host.copy_(prob.argmax(1), non_blocking=True)
        state["i"] = i + 1
        cur.synchronize()                                 # the user reads the step's result on the host

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if use_dist:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(W):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.w2v2_launch_count()
    total_ms = timed(step_device, K)
    launches = (lib.w2v2_launch_count() - l0) // K
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, K)
    clocks = sampler.stop() if rank == 0 else None

    value = world * B * K / (total_ms / 1e3)
    e2e_value = world * B * K / (e2e_ms / 1e3)
    # the instrumented step contains the gradient all-reduce in train mode: every rank must take part
    # (three instrumented steps, the median by achieved rate: one step is a single sample of the clock / power state)
    ts = [instrumented(step_device, lib) for _ in range(3)]
    ts.sort(key=lambda r: (r["gemm_flops"] / r["gemm_ms"]) if r["gemm_ms"] else 0.0)
    t = ts[1]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json: cuBLAS bf16 sustained; fp16 runs at the same tensor rate)" \
            if peaks else "fallback (B200_PROFILING.md)"
        roof = None
        try:        # DRAM traffic per launch comes from committed ncu captures (it cannot be counted inside this run)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        except Exception:
            traffic = {}
        if t["gemm_launches"]:
            g_ms = t["gemm_ms"]
            fl = t["gemm_flops"]
            ach = fl / (g_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_tc_kernel + gemm_wgrad_kernel (tcgen05), all launches of one step",
                    "achieved": ach, "peak": tens_peak, "unit": "TFLOP/s", "frac": ach / tens_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of the largest single f32-output instance
                    "traffic": traffic.get("gemm", {}).get("bytes") if WL is WORKLOADS["cfg1"] else None,
                    "traffic_note": "FROM PROFILE, not this run: " + traffic.get("gemm", {}).get("kernel", "-") + " (" +
                                    traffic.get("gemm", {}).get("source", "-") + ")",
                    "peak_source": peak_src, "launches": t["gemm_launches"], "ms_per_step": g_ms, "tflop_per_step": fl / 1e12,
                    # train mode: the FFN2 data-gradient launches carry the GELU backward and the bias-gradient sums in
                    # their epilogue (counted in the time, not in the FLOPs); per-role rates: profiles/r01_g_kernel_roofline_table.txt
                    "note": "GEMM FLOPs only; epilogue work (GELU / GELU-backward / bias sums) is inside the timed launches"}
        roof_hbm = None
        if t["conv0"]:
            c0_bytes = B * (WL["samples"] * 4 * 2 + ((WL["samples"] - 10) // 5 + 1) * 512 * 2)
            c_ms = sum(t["conv0"])
            ach = c0_bytes / (c_ms * 1e-3) / 1e9
            roof_hbm = {"bound": "hbm", "kernel": "conv0+GroupNorm+GELU stage (moments, stats, im2col, tensor-core GEMM "
                        "with GN+GELU epilogue)", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": traffic.get("conv0_stage", {}).get("bytes") if WL is WORKLOADS["cfg1"] else None,
                        "traffic_note": "FROM PROFILE, not this run (" + traffic.get("conv0_stage", {}).get("source", "-") + ")",
                        "algorithmic_bytes": c0_bytes, "ms_per_step": c_ms}
        cpu = None
        if not args.no_cpu_baseline:
            cb = 4 if train else 8
            uts, ms, cores = time_cpu(cb, 2 if train else 4, 1, args.mode)
            cpu = {"value": uts, "unit": "utt/s", "cores": cores, "kind": "port",
                   "sample": f"{2 if train else 4} steps x {cb} utterances of {WL['samples'] / 16000:g} s "
                             f"({cpu_reference_step_fn.kind}, torch fp32{' + autograd + torch Adam' if train else ''}, "
                             f"{cores} threads)"}
        line = {
            "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate+statistics+master weights", "data": "synthetic",
            "config": {"workload": workload_name(args.mode, not args.no_reg) +
                       (", raw int16 PCM input (normaliser folded into conv 0)" if args.input == "int16" else ""),
                       "mode": args.mode, "global_batch": world * B,
                       "parallelism": f"dp{world}" + (f" ({trainer.collective} of the flat fp32 gradient)" if train else
                                                       " (independent utterances, no collective)"),
                       "l2": "per-step working set (> 1.5 GB activations + 0.19 GB fp16 weights) >> 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "utt/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": B * WL["samples"] * (2 if args.input == "int16" else 4) + B * 8,
                    "d2h_bytes_per_step": (0 if train else B * emb_dim * 4) + 4 + B * 8},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
