"""Drop-in mirror of the reference's ``src/layers/pooling.py`` (same classes / signatures).

mean, mean+std, max and attentive-statistics pooling run in the sm_100a kernels
(w2v2_stat_pool / w2v2_asp_*); quantile / index / none are not on any measured configuration
 and stay the reference's plain tensor ops.
Inputs follow the reference: ``[B, T, C]`` with ``dim_to_reduce=1`` (as built at
R:src/lightning_modules/speaker/wav2vec2_fc.py:238-272) or ``[B, C, T]`` with ``dim_to_reduce=2``.
"""
from __future__ import annotations

import math
import random

import torch
import torch.nn as nn

from .. import ops


def _as_btc(tensor: torch.Tensor, dim_to_reduce: int) -> torch.Tensor:
    if tensor.dim() != 3:
        raise ValueError("pooling expects a [BATCH, TIME, FEATURE] or [BATCH, FEATURE, TIME] tensor")
    if dim_to_reduce == 1:
        return tensor.float().contiguous()
    if dim_to_reduce == 2:
        return tensor.float().transpose(1, 2).contiguous()
    raise ValueError("can only pool dimension 1 or 2")


class MeanStatPool1D(nn.Module):
    """R:src/layers/pooling.py:24-30."""

    def __init__(self, dim_to_reduce: int = 2):
        super().__init__()
        self.dim_to_reduce = dim_to_reduce

    def forward(self, tensor: torch.Tensor, lengths=None):
        """lengths (this package's extension, evaluation only): int32 CUDA vector of valid frames per utterance of a
        zero-padded ragged batch -- the statistic then covers each utterance's own frames."""
        x = _as_btc(tensor, self.dim_to_reduce)
        if lengths is not None:
            return ops.stat_pool(x.detach(), 0, lengths)
        if torch.is_grad_enabled() and x.requires_grad:
            from ..training import MeanPoolFn
            return MeanPoolFn.apply(x)
        return ops.stat_pool(x, 0)


class MeanStdStatPool1D(nn.Module):
    """R:src/layers/pooling.py:38-44: ``cat(std_mean(x))`` => [std (unbiased) || mean]."""

    def __init__(self, dim_to_reduce: int = 2):
        super().__init__()
        self.dim_to_reduce = dim_to_reduce

    def forward(self, tensor: torch.Tensor, lengths=None):
        x = _as_btc(tensor, self.dim_to_reduce)
        if lengths is not None:                     # ragged evaluation batch (see MeanStatPool1D.forward)
            return ops.stat_pool(x.detach(), 1, lengths)
        if torch.is_grad_enabled() and x.requires_grad:
            from ..training import MeanStdPoolFn
            return MeanStdPoolFn.apply(x)
        return ops.stat_pool(x, 1)


class MaxPool1D(nn.Module):
    """R:src/layers/pooling.py:74-80."""

    def __init__(self, dim_to_reduce: int = 2):
        super().__init__()
        self.dim_to_reduce = dim_to_reduce

    def forward(self, tensor: torch.Tensor, lengths=None):
        x = _as_btc(tensor, self.dim_to_reduce)
        if lengths is not None:
            return ops.stat_pool(x.detach(), 2, lengths)
        if torch.is_grad_enabled() and x.requires_grad:
            # training: torch's differentiable reduction (same values; max pooling is on no measured configuration and
            # its backward is a scatter of B*C numbers -- the kernel below has no autograd node)
            return x.max(dim=1).values
        return ops.stat_pool(x, 2)


class QuantilePool1D(nn.Module):
    """R:src/layers/pooling.py:51-67 (plain tensor op; not on a measured configuration)."""

    def __init__(self, dim_to_reduce: int = 2):
        super().__init__()
        self.dim_to_reduce = dim_to_reduce
        self.quantiles = torch.Tensor([0, 0.25, 0.5, 0.75, 1]).detach()

    def forward(self, tensor: torch.Tensor):
        q = torch.quantile(tensor, self.quantiles.to(tensor.device), dim=self.dim_to_reduce)
        return torch.flatten(torch.transpose(q, 0, 1), start_dim=1, end_dim=2)


class _AttentiveStatisticsPooling(nn.Module):
    """Parameters of speechbrain's ``AttentiveStatisticsPooling(channels, attention_channels=128,
    global_context=True)`` under speechbrain's state_dict names (tdnn.conv.conv / tdnn.norm.norm /
    conv.conv); forward in the sm_100a kernels.  x: [N, C, L] -> [N, 2C, 1]."""

    def __init__(self, channels: int, attention_channels: int = 128):
        super().__init__()
        self.channels, self.attention_channels = channels, attention_channels
        self.eps = 1e-12

        def holder(name, mod):
            h = nn.Module()
            h.add_module(name, mod)
            return h
        self.tdnn = nn.Module()
        self.tdnn.add_module("conv", holder("conv", nn.Conv1d(channels * 3, attention_channels, 1)))
        self.tdnn.add_module("norm", holder("norm", nn.BatchNorm1d(attention_channels)))
        self.conv = holder("conv", nn.Conv1d(attention_channels, channels, 1))

    def forward(self, x_ncl: torch.Tensor) -> torch.Tensor:
        return self.forward_btc(x_ncl.transpose(1, 2).contiguous()).unsqueeze(2)

    def forward_btc(self, x: torch.Tensor, lengths=None) -> torch.Tensor:
        c1, bn, c2 = self.tdnn.conv.conv, self.tdnn.norm.norm, self.conv.conv
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(q.requires_grad for q in self.parameters()))
        if lengths is not None and (self.training or needs_grad):
            raise NotImplementedError("ragged batches (lengths=...) are an evaluation feature: call .eval() under no_grad")
        if self.training or needs_grad:
            # batch-statistics BatchNorm and / or a backward pass: the autograd Function over the same kernels
            from ..training import AspPoolFn
            return AspPoolFn.apply(x.float().contiguous(), c1.weight, c1.bias, bn.weight, bn.bias, c2.weight, c2.bias, self)
        B, T, C = x.shape
        x = x.float().contiguous()
        # TDNN 1x1 conv over [x | mean | std] (3C channels) WITHOUT materialising the concatenation: the mean / std columns are
        # constant over an utterance's frames, so their share of the product is a per-utterance bias
        #     z[b, t] = W1x x[b, t] + (W1m mean_b + W1s std_b + b1)
        # -- the frame GEMM runs over K = C (x3 for the error-compensated fp16 split) instead of 3C, and the operand is a
        # third of the size (44 MB instead of 132 MB at 64 x 149 x 768)
        A = self.attention_channels
        w1 = c1.weight.detach().float().reshape(A, 3 * C)
        xs = ops.split3_rows(x.view(B * T, C), 0)                            # [B*T, 3C] fp16: [hi | lo | hi]
        z = ops.gemm_f16(xs, ops.split3_rows(w1[:, :C].contiguous(), 1), None, 0, torch.float32)
        stats = ops.stat_pool(x, 3, lengths)                                 # [B, 2C] = [mean | std] (uniform weights)
        u = ops.gemm_f16(ops.split3_rows(stats, 0), ops.split3_rows(w1[:, C:].contiguous(), 1), c1.bias.detach().float(), 0,
                         torch.float32)                                      # [B, A]
        scale = (bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
        shift = (bn.bias.detach() - bn.running_mean * scale).float().contiguous()
        y16 = ops.asp_relu_bn_tanh(z.contiguous(), scale, shift, u.contiguous(), T)      # tanh(BN(ReLU(.)))
        logits = ops.gemm_f16(y16, ops.cast_f16(c2.weight.detach().view(C, self.attention_channels)),
                              c2.bias.detach().float(), 0, torch.float32)
        return ops.asp_pool(x, logits.contiguous().view(B, T, C), lengths)   # [B, 2C] = [mean || std]


class AttentiveStatPool1D(nn.Module):
    """R:src/layers/pooling.py:87-106."""

    def __init__(self, embedding_size: int, dim_to_reduce: int = 2):
        super().__init__()
        self.pooling_layer = _AttentiveStatisticsPooling(embedding_size)
        self.dim_to_reduce = dim_to_reduce

    def forward(self, tensor: torch.Tensor, lengths=None):
        if self.dim_to_reduce == 2:
            pooled_embedding = self.pooling_layer.forward_btc(tensor.transpose(1, 2).contiguous(), lengths)
        elif self.dim_to_reduce == 1:
            pooled_embedding = self.pooling_layer.forward_btc(tensor, lengths)
        else:
            raise ValueError("can only pool dimension 1 or 2")
        pooled_embedding = pooled_embedding.squeeze()
        if len(pooled_embedding.shape) == 1:
            pooled_embedding = pooled_embedding[None, :]
        return pooled_embedding


class IndexPool1D(nn.Module):
    """R:src/layers/pooling.py:112-154, including the upstream quirk that "middle" selects the
    last frame."""

    def __init__(self, selection_method: str, dim_to_reduce: int):
        super().__init__()
        self.selection_method = selection_method
        self.dim_to_reduce = dim_to_reduce

    def forward(self, tensor: torch.Tensor):
        n = tensor.shape[self.dim_to_reduce]
        if self.selection_method in ("first", "first+cls"):
            idx = 0
        elif self.selection_method in ("middle", "last"):
            idx = n - 1
        elif self.selection_method == "random":
            idx = random.randint(0, int(n) - 1)
        else:
            raise ValueError(f"unknown index {self.selection_method}")
        view = tensor[:, idx, :] if self.dim_to_reduce == 1 else tensor[:, :, idx]
        return torch.clone(view)


class NoPooling(nn.Module):
    """R:src/layers/pooling.py:161-166."""

    def __init__(self):
        super().__init__()

    def forward(self, tensor: torch.Tensor):
        return tensor
