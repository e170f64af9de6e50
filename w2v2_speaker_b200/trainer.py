"""Data-parallel training step around the reference-facing module:

    forward (sm_100a kernels) -> loss.backward() (hand-written backward via training.*Fn)
    -> NCCL all-reduce of the flat fp32 gradient buffer over NVLink (world > 1)
    -> fused Adam over the flat fp32 parameter buffer (w2v2_adam_step).

Parameters and gradients of the trainable tensors are views into two flat fp32 buffers (so the
all-reduce is one message -- bucketed below -- and Adam is one launch), which is also what the
reference's Lightning DDP + ``torch.optim.Adam`` (R:config/optim/algo/adam.yaml, R:config/trainer/trainer.yaml:6-9)
amount to.  LayerDrop-skipped or unused parameters simply keep a zero gradient.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops


class FlatAdamTrainer:
    def __init__(self, module: torch.nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 bucket_bytes: int = 64 << 20):
        self.module = module
        self.lr, self.betas, self.eps = lr, betas, eps
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            k = p.numel()
            self.flat_p[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + k].view_as(p)
            p.grad = self.flat_g[o:o + k].view_as(p)
            o += k
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None

    def _refresh_module_weights(self):
        for m in self.module.modules():
            if hasattr(m, "refresh"):
                m.refresh()
            for attr in ("_w_split", "_w_sig"):
                if hasattr(m, attr):
                    setattr(m, attr, None)

    def allreduce_grads(self):
        """Sum-all-reduce of the flat gradient in buckets on a side stream (the mean's 1/world is folded
        into the Adam gradient scale)."""
        if self.world == 1:
            return
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            n = self.flat_g.numel()
            for s in range(0, n, self.bucket_elems):
                dist.all_reduce(self.flat_g[s:min(n, s + self.bucket_elems)], op=dist.ReduceOp.SUM)
        cur.wait_stream(self.comm_stream)

    def step(self, wav: torch.Tensor, labels: torch.Tensor):
        """One optimisation step; returns (loss, softmax) like the reference's training_step uses them."""
        self.flat_g.zero_()
        emb, pred = self.module(wav)
        loss, prob = self.module.loss_fn(pred, labels)
        loss.backward()
        self.allreduce_grads()
        self.step_count += 1
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.step_count, grad_scale=1.0 / self.world)
        self._refresh_module_weights()
        return loss.detach(), prob
