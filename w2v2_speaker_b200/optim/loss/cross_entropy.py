"""Mirror of R:src/optim/loss/cross_entropy.py:14-33 on the fused softmax/CE/argmax kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops


class CrossEntropyLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, logits: torch.Tensor, label_indexes: torch.Tensor):
        return self._ce_loss(logits, label_indexes)

    def _ce_loss(self, logits: torch.Tensor, label_indexes: torch.Tensor):
        # logits [BATCH_SIZE, NUM_SPEAKERS] (unnormalised), label indexes [BATCH_SIZE] int64
        if logits.dim() != 2 or label_indexes.shape[0] != logits.shape[0]:
            raise ValueError("expected logits [BATCH_SIZE, NUM_SPEAKERS] and labels [BATCH_SIZE]")
        if torch.is_grad_enabled() and logits.requires_grad:
            from ...training import CrossEntropyFn
            return CrossEntropyFn.apply(logits, label_indexes.to(torch.int64))
        logits = logits.float()
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        prob, loss_rows, _ = ops.softmax_ce(logits, label_indexes.to(torch.int64))
        return ops.mean_rows(loss_rows), prob
