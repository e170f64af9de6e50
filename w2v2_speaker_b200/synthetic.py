"""Synthetic utterance batches for benchmarks and smoke runs (there is no dataset offline).

Product-side generator: bench.py's product arm and the examples use this, nothing under ``oracle/`` (test
infrastructure).  The waveform is white noise standardised per utterance, which is what the reference's input
normaliser hands the network (R:src/data/preprocess/input_normalisation.py:53-67, ``channel_wise=False``)."""
from __future__ import annotations

from typing import Tuple

import torch


def synthetic_batch(batch: int, num_samples: int, num_speakers: int = 5994, seed: int = 1234) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (waveform f32 [batch, num_samples] with zero mean / unit variance per utterance, labels int64 [batch])."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, num_samples, generator=g, dtype=torch.float32)
    x = (x - x.mean(dim=1, keepdim=True)) / (x.std(dim=1, keepdim=True) + 1e-5)
    labels = torch.randint(0, num_speakers, (batch,), generator=g, dtype=torch.int64)
    return x, labels
