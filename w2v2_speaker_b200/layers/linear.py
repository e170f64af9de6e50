"""Classifier Linear(E -> num_speakers) of the CE head (R:src/lightning_modules/speaker/wav2vec2_fc.py:199-210)
on the tensor cores with error-compensated fp16 operands (fp32-accurate logits: argmax parity)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops


class SpeakerLinear(nn.Linear):
    """``nn.Linear`` parameters / state_dict, forward through w2v2_split3_rows + w2v2_gemm_f16."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__(in_features, out_features, bias)
        self._w_split = None
        self._w_sig = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        sig = (self.weight.data_ptr(), self.weight._version)
        if self._w_split is None or sig != self._w_sig:
            self._w_split = ops.split3_rows(self.weight.detach().float(), 1)
            self._w_sig = sig
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            from ..training import SpeakerLinearFn
            return SpeakerLinearFn.apply(x, self.weight, self.bias, self._w_split)
        xa = ops.split3_rows(x.detach().float().contiguous(), 0)
        b = self.bias.detach().float() if self.bias is not None else None
        return ops.gemm_f16(xa, self._w_split, b, 0, torch.float32)
