"""Mirror of R:src/optim/loss/aam_softmax.py:22-74 (ArcFace / AAM-softmax) on the sm_100a kernels.

cosine = normalize(x) . normalize(W)^T runs on the tensor cores with error-compensated fp16
operands (hi/lo split, ~fp32-accurate; see w2v2_l2norm_rows_split3), the margin / scale /
softmax / CE / argmax are one fused warp-primitive kernel."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops


class AngularAdditiveMarginSoftMaxLoss(nn.Module):
    def __init__(self, input_features, output_features, margin=0.3, scale=15, easy_margin=False):
        super().__init__()
        self.margin = margin
        self.scale = scale
        self.input_features = input_features
        self.fc_weights = nn.Parameter(torch.empty(output_features, input_features), requires_grad=True)
        nn.init.xavier_normal_(self.fc_weights, gain=1)
        self.easy_margin = easy_margin
        self.cos_m = math.cos(self.margin)
        self.sin_m = math.sin(self.margin)
        # make the function cos(theta+m) monotonic decreasing while theta in [0, 180] degrees
        self.th = math.cos(math.pi - self.margin)
        self.mm = math.sin(math.pi - self.margin) * self.margin
        self._w_split = None
        self._w_sig = None

    def _weights(self) -> torch.Tensor:
        sig = (self.fc_weights.data_ptr(), self.fc_weights._version)
        if self._w_split is None or sig != self._w_sig:
            self._w_split = ops.l2norm_rows_split3(self.fc_weights.detach().float(), 1)
            self._w_sig = sig
        return self._w_split

    def forward(self, x, label=None):
        assert x.size()[0] == label.size()[0]
        assert x.size()[1] == self.input_features
        if torch.is_grad_enabled() and (x.requires_grad or self.fc_weights.requires_grad):
            from ...training import AamSoftmaxFn
            return AamSoftmaxFn.apply(x, self.fc_weights, label.to(torch.int64), self.margin, self.scale,
                                      self.easy_margin, self._weights())
        xa = ops.l2norm_rows_split3(x.detach().float(), 0)
        cosine = ops.gemm_f16(xa, self._weights(), None, 0, torch.float32)       # [B, S] (padded pitch)
        prob, loss_rows, _ = ops.aam_softmax_ce(cosine, label.to(torch.int64), self.margin, self.scale,
                                                self.easy_margin)
        return ops.mean_rows(loss_rows), prob
