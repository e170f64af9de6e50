"""Data-parallel training step around the reference-facing module:

    forward (sm_100a kernels) -> loss.backward() (hand-written backward via training.*Fn)
    -> NCCL all-reduce of the flat fp32 gradient buffer over NVLink (world > 1)
    -> fused Adam over the flat fp32 parameter buffer (w2v2_adam_step).

Parameters and gradients of the trainable tensors are views into two flat fp32 buffers (so the
all-reduce is one message -- bucketed below -- and Adam is one launch per segment), which is also what
the reference's Lightning DDP + ``torch.optim.Adam`` (R:config/optim/algo/adam.yaml,
R:config/trainer/trainer.yaml:6-9) amount to.  The encoder's backward writes its weight gradients
straight into the flat buffer (``model._grad_sink``), so no per-parameter accumulation pass exists;
they stay loss-scaled and Adam undoes the scale.  Unused / LayerDrop-skipped parameters keep a zero
gradient.

The optimizer runs on its own stream: Adam, the re-derivation of the fp16 operand copies and the zeroing of
the gradient buffer are HBM-bound and depend on nothing of the next step's CNN forward (frozen feature
extractor: its weights do not change), which is tensor-bound -- so step n's update overlaps step n+1's
feature extractor.  The encoder waits for the update right after the CNN (``model._pre_encoder_hook``).
"""
from __future__ import annotations

from typing import List

import os

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import nvtx_range
from .dist_utils import remaining_spans
from .models.wav2vec2 import Wav2Vec2ModelB200
from .training import LOSS_SCALE, GradBook, encoder_grad_order


class FlatAdamTrainer:
    def __init__(self, module: torch.nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 bucket_bytes: int = 64 << 20, lr_schedule=None):
        """`lr_schedule`: optional callable step (1-based) -> learning rate, evaluated once per iteration like the
        reference steps its scheduler (`interval: step`, R:config/optim/schedule/one_cycle.yaml; `one_cycle_lr` below)."""
        self.module = module
        self.lr, self.betas, self.eps = lr, betas, eps
        self.lr_schedule = lr_schedule
        self.model = next((m for m in module.modules() if isinstance(m, Wav2Vec2ModelB200)), None)
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bucket_elems = max(1, bucket_bytes // 4)
        self._overlapped = []
        self._refreshable = None
        self._pending = None
        # transformer layers per overlapped all-reduce, deepest group first.  Large groups while the rest of the backward
        # still hides them (fewer NCCL launches competing with the GEMMs), small ones at the end: what is sent after layer 0
        # is exposed, so the last group is a single layer (W2V2_AR_SCHEDULE="4,4,2,1,1"; W2V2_AR_LAYERS=n: uniform groups)
        if "W2V2_AR_LAYERS" in os.environ:
            self.ar_schedule = [max(1, int(os.environ["W2V2_AR_LAYERS"]))]
        else:
            self.ar_schedule = [max(1, int(x)) for x in os.environ.get("W2V2_AR_SCHEDULE", "4,4,2,1,1").split(",") if x.strip()]
        self._ar_group = 0
        self._ar_from_env = "W2V2_AR_LAYERS" in os.environ or "W2V2_AR_SCHEDULE" in os.environ
        self.params: List[torch.nn.Parameter] = []
        self._state = {}                       # id(parameter) -> (exp_avg, exp_avg_sq) views of the previous layout
        self._layout()
        dev = self.params[0].device
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        # optimizer stream (see module docstring); with a trainable CNN the update must finish before the next forward
        self.opt_stream = torch.cuda.Stream(device=dev)
        if self.world > 1:
            self._broadcast_initial_state()

    # ---- flat buffers -------------------------------------------------------------------------------------------
    def _trainable_signature(self):
        ps = self.__dict__.get("_all_params")
        if ps is None:                     # the Parameter objects never change after construction: walk the tree once
            ps = self._all_params = list(self.module.parameters())
        return tuple(p.requires_grad for p in ps)

    def _layout(self):
        """(Re)build the flat parameter / gradient / Adam-state buffers for the CURRENT set of trainable parameters.
        Called at construction and again whenever `requires_grad` changed between two steps -- the reference's freeze
        protocol (R:src/lightning_modules/speaker/wav2vec2_fc.py:339-361) trains the heads alone for `num_frozen_steps`
        and then releases the encoder; Adam state of parameters that stay trainable is carried over, newly released
        parameters start from zero moments like a freshly added `torch.optim` param group."""
        module = self.module
        named = dict(module.named_parameters())
        enc_named = dict(self.model.named_parameters()) if self.model is not None else {}
        # keep the state of the previous layout (views into the old flat buffers stay alive through these references)
        old_state = {}
        o = 0
        for p_ in self.params:
            k = p_.numel()
            old_state[id(p_)] = (self.m[o:o + k], self.v[o:o + k])
            o += k
        if self.model is not None:
            self.model._grad_sink = None
            self.model._grad_ready_hook = None
            self.model._backward_start_hook = None
            self.model._pre_encoder_hook = None
        # segment 0: encoder parameters in the order the backward wants (q|k|v adjacent), loss-scaled grads
        enc_order = [k for k in (encoder_grad_order(self.model.arch) if self.model is not None else [])
                     if enc_named[k].requires_grad]
        if self.model is not None and len(enc_order) != len(encoder_grad_order(self.model.arch)):
            enc_order = []          # (partially) frozen encoder: its trainable parameters go through autograd accumulation
        enc_ids = {id(enc_named[k]) for k in enc_order}
        seg0 = [enc_named[k] for k in enc_order]
        seg1 = [p for p in named.values() if p.requires_grad and id(p) not in enc_ids]
        self.params = seg0 + seg1
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.n0 = sum(p.numel() for p in seg0)
        n = sum(p.numel() for p in self.params)
        flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = self._alloc_gradient(n, dev)
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        v = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for i, p in enumerate(self.params):
            k = p.numel()
            flat_p[o:o + k].copy_(p.data.reshape(-1))
            p.data = flat_p[o:o + k].view_as(p)
            if id(p) in old_state:
                m[o:o + k].copy_(old_state[id(p)][0])
                v[o:o + k].copy_(old_state[id(p)][1])
            if i >= len(seg0):
                p.grad = self.flat_g[o:o + k].view_as(p)      # autograd accumulates the (unscaled) head grads here
            else:
                p.grad = None                                  # the backward writes straight into the flat buffer
            o += k
        self.flat_p, self.m, self.v = flat_p, m, v
        if seg0:
            self.model._grad_sink = GradBook({k: enc_named[k].shape for k in enc_order}, enc_order, dev,
                                             flat=self.flat_g[:self.n0])
        cnn_trainable = self.model is not None and any(q.requires_grad for q in self.model._items()[3])
        self._overlap_update = (self.model is not None and not cnn_trainable and
                                os.environ.get("W2V2_OPT_STREAM", "1") != "0")        # =0: update in stream order (A/B)
        # the encoder joins the optimizer stream right after its CNN; a frozen encoder never runs that hook (step() joins)
        self._encoder_joins = self._overlap_update and bool(seg0)
        if self._encoder_joins:
            self.model._pre_encoder_hook = self._join_update
        if seg0 and self.world > 1:
            self.model._grad_ready_hook = self._layer_ready      # spans of flat_g[:n0] == GradBook offsets
            self.model._backward_start_hook = self._heads_ready
        for h in getattr(self, "_seg1_hooks", []):
            h.remove()
        self._seg1_count, self._seg1_seen = len(seg1), 0
        self._seg1_hooks = [p.register_post_accumulate_grad_hook(self._count_seg1) for p in seg1] if self.world > 1 else []
        self._refreshable = None
        self._sig = self._trainable_signature()
        if self.model is not None:
            self.model.refresh()          # parameter storage moved: re-derive the operand copies on next use

    def _alloc_gradient(self, n: int, dev) -> torch.Tensor:
        """The flat gradient.  With several ranks on one NVSwitch domain it is allocated as SYMMETRIC memory with a
        multicast mapping (torch.distributed._symmetric_memory: allocation + rendezvous + barriers are plumbing), so that
        the exchange step can be this package's NVLS kernel (csrc/allreduce.cu) instead of ncclAllReduce; anything
        missing (no multicast support, a torch without symmetric memory, W2V2_NVLS=0) falls back to NCCL, loudly."""
        self._symm, self._mc_ptr = None, 0
        # light CTAs (256 threads, ~24 registers, no shared memory: they fit next to a GEMM CTA), one per SM at most: what
        # limits the kernel is bytes in flight towards the switch (128 x 256 x 4 x 16 B = 2 MB)
        self._nvls_ctas = int(os.environ.get("W2V2_NVLS_CTAS", "128"))
        if self.world > 1 and os.environ.get("W2V2_NVLS", "1") != "0":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                group = dist.group.WORLD
                try:
                    symm_mem.enable_symm_mem_for_group(group.group_name)
                except Exception:
                    pass                                   # newer torch: enabled implicitly
                n_pad = (n + 3) // 4 * 4                   # the kernel moves 16-byte vectors
                buf = symm_mem.empty(n_pad, dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(buf, group)
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                if mc == 0:
                    raise RuntimeError("the symmetric allocation has no multicast mapping (NVLS not available)")
                buf.zero_()
                self._symm, self._mc_ptr, self._symm_buf = hdl, mc, buf
                if not self._ar_from_env:
                    # the NVLS kernel runs next to the GEMMs instead of displacing them, so small spans cost nothing and
                    # leave the shortest tail: one transformer layer per all-reduce (8 GPUs: 10.76 -> 10.64 ms per step)
                    self.ar_schedule = [1]
                self.collective = f"NVLS two-shot all-reduce (own kernel, {self._nvls_ctas} CTAs, symmetric memory)"
                return buf[:n]
            except Exception as e:                         # pragma: no cover - depends on the machine
                if dist.get_rank() == 0:
                    print(f"[w2v2_speaker_b200] NVLS all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", flush=True)
        self.collective = "NCCL all-reduce" if self.world > 1 else "none (single rank)"
        return torch.zeros(n, dtype=torch.float32, device=dev)

    def _broadcast_initial_state(self):
        """What DistributedDataParallel does at construction (R:config/trainer/trainer.yaml:6-9 `accelerator: ddp`): every
        rank starts from rank 0's parameters and buffers (BatchNorm running statistics of the attentive pooling)."""
        for p in self.module.parameters():                 # frozen parameters included
            dist.broadcast(p.data, src=0)
        for b in self.module.buffers():
            if b.is_floating_point() or b.dtype in (torch.int64, torch.int32):
                dist.broadcast(b, src=0)
        self._refresh_module_weights()

    def _refresh_module_weights(self):
        if self._refreshable is None:          # the module tree is fixed: walk it once
            mods = list(self.module.modules())
            self._refreshable = ([m for m in mods if hasattr(m, "refresh")],
                                 [m for m in mods if hasattr(m, "_w_split") or hasattr(m, "_w_sig")])
        for m in self._refreshable[0]:
            m.refresh()
        for m in self._refreshable[1]:
            for attr in ("_w_split", "_w_sig"):
                if hasattr(m, attr):
                    setattr(m, attr, None)

    def _layer_ready(self, lo: int, hi: int):
        """Backward hook: layer spans arrive deepest first and are contiguous; merge `ar_layers` of them into one
        all-reduce (fewer NCCL launches competing with the backward GEMMs for SMs)."""
        if self._pending is None:
            self._pending = [lo, hi, 1]
        else:
            self._pending[0] = min(self._pending[0], lo)
            self._pending[1] = max(self._pending[1], hi)
            self._pending[2] += 1
        if self._pending[2] >= self.ar_schedule[min(self._ar_group, len(self.ar_schedule) - 1)]:
            self._ar_group += 1
            self._flush_pending()

    def _count_seg1(self, _param):
        self._seg1_seen += 1

    def _heads_ready(self):
        """Backward hook, called when the encoder's backward starts: autograd has already accumulated the gradients of
        everything in front of it (pooling, FC head / AAM weights -- segment 1 of the flat buffer, 18-37 MB), so their
        all-reduce goes out first and runs under the whole encoder backward instead of after it."""
        # ... provided every one of them has been accumulated in this backward: a trainable tensor in FRONT of the encoder
        # (none exists in the reference's modules) would get its gradient later, and the span then waits for the end
        if self.flat_g.numel() > self.n0 and self._seg1_seen >= self._seg1_count:
            self._reduce_span(self.n0, self.flat_g.numel())

    def _flush_pending(self):
        if self._pending is not None:
            lo, hi, _ = self._pending
            self._pending = None
            self._reduce_span(lo, hi)

    def _reduce_span(self, lo: int, hi: int):
        """Enqueue the sum-all-reduce of flat_g[lo:hi] on the communication stream, ordered after everything
        the compute stream has enqueued so far (called while the backward is still being scheduled: NCCL
        then runs concurrently with the remaining layers)."""
        if hi <= lo:
            return
        self._overlapped.append((lo, hi))
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            if self._symm is not None:
                # every rank's replica of the span is complete -> pull / broadcast through the switch -> every slice landed
                lo4, hi4 = lo // 4 * 4, (hi + 3) // 4 * 4          # (span edges are parameter boundaries: multiples of 4
                self._symm.barrier(channel=0)                      #  except the padded end of the buffer)
                _lib.call("w2v2_nvls_allreduce_f32", _lib.c_void_p(self._mc_ptr), lo4, hi4, dist.get_rank(), self.world,
                          self._nvls_ctas, _lib.stream_ptr())
                self._symm.barrier(channel=0)
                return
            for s in range(lo, hi, self.bucket_elems):
                dist.all_reduce(self.flat_g[s:min(hi, s + self.bucket_elems)], op=dist.ReduceOp.SUM)

    def allreduce_grads(self):
        """Sum-all-reduce of whatever part of the flat gradient the backward has not already sent (the
        per-layer spans go out from inside the backward, deepest layer first), then join the streams.
        The mean's 1/world is folded into the Adam gradient scale.  Every rank issues the same spans in
        the same order."""
        if self.world == 1:
            return
        self._flush_pending()
        for lo, hi in remaining_spans(self._overlapped, self.flat_g.numel()):
            self._reduce_span(lo, hi)
        self._overlapped = []
        self._ar_group = 0
        self._seg1_seen = 0
        torch.cuda.current_stream().wait_stream(self.comm_stream)

    def _join_update(self):
        """Called by the encoder right after the CNN forward: everything from here on reads updated parameters
        (and, in the backward, writes the gradient buffer the optimizer stream has just zeroed)."""
        torch.cuda.current_stream().wait_stream(self.opt_stream)

    def synchronize(self):
        """Make the current stream wait for the pending parameter update (call before reading parameters
        outside step(): evaluation, checkpointing)."""
        torch.cuda.current_stream().wait_stream(self.opt_stream)

    def step(self, wav: torch.Tensor, labels: torch.Tensor):
        """One optimisation step; returns (loss, softmax) like the reference's training_step uses them."""
        cur = torch.cuda.current_stream()
        if self._trainable_signature() != self._sig:
            # the set of trainable parameters changed (freeze protocol released the encoder, or froze something):
            # finish the pending update, then rebuild the flat buffers and the gradient sink for the new set
            cur.wait_stream(self.opt_stream)
            self._layout()
        if not self._encoder_joins:
            # in-order update, or an encoder without trainable parameters: its (evaluation-path) forward never calls the
            # join hook, and the heads read weights / operand copies the optimizer stream may still be writing
            cur.wait_stream(self.opt_stream)
        with nvtx_range("forward"):
            emb, pred = self.module(wav)
            cur.wait_stream(self.opt_stream)      # no-op if the encoder already joined (it always does when it runs)
            loss, prob = self.module.loss_fn(pred, labels)
        with nvtx_range("backward"):
            loss.backward()
        with nvtx_range("allreduce"):
            self.allreduce_grads()
        self.step_count += 1
        if self.lr_schedule is not None:
            self.lr = float(self.lr_schedule(self.step_count))
        b1, b2 = self.betas
        n0, n = self.n0, self.flat_p.numel()
        opt_stream = self.opt_stream if self._overlap_update else cur
        opt_stream.wait_stream(cur)
        with nvtx_range("optimizer"), torch.cuda.stream(opt_stream):
            if n0:
                ops.adam_step(self.flat_p[:n0], self.flat_g[:n0], self.m[:n0], self.v[:n0], self.lr, b1, b2, self.eps,
                              self.step_count, grad_scale=1.0 / (LOSS_SCALE * self.world), zero_grad=True)
            if n > n0:
                ops.adam_step(self.flat_p[n0:], self.flat_g[n0:], self.m[n0:], self.v[n0:], self.lr, b1, b2, self.eps,
                              self.step_count, grad_scale=1.0 / self.world, zero_grad=True)
            self._refresh_module_weights()         # (the gradient buffer was cleared by the Adam pass itself)
        return loss.detach(), prob


def one_cycle_lr(max_lr: float, total_steps: int, pct_start: float = 0.3, div_factor: float = 25.0,
                 final_div_factor: float = 1e4):
    """`torch.optim.lr_scheduler.OneCycleLR` (cosine annealing, two phases) as a step -> lr callable for
    FlatAdamTrainer(lr_schedule=...): the reference's default schedule (R:config/optim/schedule/one_cycle.yaml,
    stepped every iteration).  step is 1-based: the value returned for step s is the rate torch's scheduler holds
    after s - 1 calls of `scheduler.step()`."""
    import math
    initial, min_lr = max_lr / div_factor, max_lr / div_factor / final_div_factor
    up_end = float(pct_start * total_steps) - 1.0
    last = float(total_steps) - 1.0

    def cos(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1.0)

    def lr(step: int) -> float:
        n = min(max(step - 1, 0), total_steps - 1)
        if n <= up_end:
            return cos(initial, max_lr, n / up_end if up_end > 0 else 1.0)
        return cos(max_lr, min_lr, (n - up_end) / (last - up_end))
    return lr
