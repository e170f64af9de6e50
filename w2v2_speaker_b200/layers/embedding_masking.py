"""Mirror of the reference's ``src/layers/embedding_masking.py`` (R:...:18-123).

On the hot path this layer is the identity: ``Wav2vec2FCModule`` builds it with
``timestep_mask_prob=0`` (R:src/lightning_modules/speaker/wav2vec2_fc.py:162-169) and upstream only
draws the channel mask when the *timestep* probability is positive (R:...:79).  The
masking branch is kept for interface completeness with the same (quirky) semantics.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class EmbeddingMasker(nn.Module):
    def __init__(self, timestep_mask_prob: float, timestep_mask_width: int, channel_mask_prob: float,
                 channel_mask_width: int, time_dim: int = 1, embedding_dim: int = 2):
        if not (0 <= channel_mask_prob <= 1):
            raise ValueError(f"probability channel_mask_prob {channel_mask_prob} expected to be in range [0,1]")
        if not (0 <= timestep_mask_prob <= 1):
            raise ValueError(f"probability timestep_mask_prob {timestep_mask_prob} expected to be in range [0,1]")
        if time_dim == 0 or embedding_dim == 0:
            raise ValueError("dimensions to mask cannot be dim 0 (batch dimension)")
        super().__init__()
        self.timestep_mask_prob = timestep_mask_prob
        self.timestep_mask_width = timestep_mask_width
        self.channel_mask_prob = channel_mask_prob
        self.channel_mask_width = channel_mask_width
        self.time_dim = time_dim
        self.embedding_dim = embedding_dim

    @staticmethod
    def _span_mask(n: int, prob: float, width: int) -> torch.Tensor:
        """1 = keep, 0 = masked; every drawn start masks `width` consecutive positions."""
        keep = torch.ones(n)
        for s in torch.nonzero(torch.rand(n) <= prob).flatten().tolist():
            keep[s:s + width] = 0
        return keep

    def forward(self, embedding_tensor: torch.Tensor):
        if not self.training or (self.timestep_mask_prob + self.channel_mask_prob == 0):
            return embedding_tensor
        assert len(embedding_tensor.shape) == 3
        if self.timestep_mask_prob <= 0:
            return embedding_tensor * 1.0        # upstream draws neither mask in this case
        m = torch.ones(embedding_tensor.shape, device=embedding_tensor.device)
        for dim, prob, width in ((self.time_dim, self.timestep_mask_prob, self.timestep_mask_width),
                                 (self.embedding_dim, self.channel_mask_prob, self.channel_mask_width)):
            keep = self._span_mask(embedding_tensor.shape[dim], prob, width).to(m.device)
            shape = [1, 1, 1]
            shape[dim] = -1
            m = m * keep.view(shape)
        return m * embedding_tensor
