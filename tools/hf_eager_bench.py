"""'Same box, stock kernels' line: the HuggingFace ``Wav2Vec2Model`` the reference wraps, run eagerly by
PyTorch (cuDNN / cuBLAS / SDPA) on the workload bench.py measures -- wav2vec2-base, 64 utterances of 3 s, mean pooling,
Linear(768 -> 5994) + cross-entropy; forward (eval) and training step (forward + backward + torch.optim.Adam, CNN
frozen, the reference's default regularisation) -- in fp32, tf32 and fp16 autocast (the paper's `precision: 16`).
Not part of the product or of bench.py's contract: a yardstick to quote next to it.

usage (GPU box): python tools/hf_eager_bench.py [--batch 64] [--seconds 3] [--steps 10] [--warmup 3] [--device cuda]
prints one JSON line per (mode, precision)."""
import argparse
import json
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


def build(device, train: bool):
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    reg = dict(activation_dropout=0.0, attention_dropout=0.1, feat_proj_dropout=0.1, hidden_dropout=0.1, layerdrop=0.05,
               mask_time_prob=0.05, mask_time_length=10, mask_feature_prob=0.0)        # R:config/network/wav2vec2_fc.yaml
    cfg = Wav2Vec2Config(**reg)
    model = Wav2Vec2Model(cfg).to(device)
    head = nn.Linear(768, 5994).to(device)
    model.train(train); head.train(train)
    if train:
        model.feature_extractor.requires_grad_(False)       # completely_freeze_feature_extractor: true
    return model, head


def run(mode: str, precision: str, args):
    dev = torch.device(args.device)
    train = mode == "train"
    torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
    torch.backends.cudnn.allow_tf32 = precision == "tf32"
    model, head = build(dev, train)
    params = [p for p in list(model.parameters()) + list(head.parameters()) if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-5) if train else None
    scaler = torch.amp.GradScaler(enabled=train and precision == "fp16" and dev.type == "cuda")
    g = torch.Generator().manual_seed(0)
    wav = torch.randn(args.batch, int(16000 * args.seconds), generator=g).to(dev)
    labels = torch.randint(0, 5994, (args.batch,), generator=g).to(dev)
    amp = torch.autocast(dev.type, dtype=torch.float16 if dev.type == "cuda" else torch.bfloat16, enabled=precision == "fp16")

    def step():
        if train:
            opt.zero_grad(set_to_none=True)
            with amp:
                loss = F.cross_entropy(head(model(wav).last_hidden_state.mean(1)).float(), labels)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
        else:
            with torch.no_grad(), amp:
                loss = F.cross_entropy(head(model(wav).last_hidden_state.mean(1)).float(), labels)
        return loss

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    sync()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    print(json.dumps({"impl": "hf-eager", "mode": mode, "precision": precision, "ms_per_step": ms,
                      "value": args.batch / ms * 1e3, "unit": "utt/s", "batch": args.batch, "seconds": args.seconds,
                      "steps": args.steps, "loss": float(loss.detach()), "torch": torch.__version__,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9 if dev.type == "cuda" else None}))
    del model, head, opt
    if dev.type == "cuda":
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--modes", default="forward,train")
    ap.add_argument("--precisions", default="fp32,tf32,fp16")
    args = ap.parse_args()
    for mode in args.modes.split(","):
        for precision in args.precisions.split(","):
            try:
                run(mode, precision, args)
            except torch.OutOfMemoryError as e:            # fp32 activations of 64 x 3 s are ~10 GB: fits, but be explicit
                print(json.dumps({"impl": "hf-eager", "mode": mode, "precision": precision, "error": str(e)[:200]}))


if __name__ == "__main__":
    main()
