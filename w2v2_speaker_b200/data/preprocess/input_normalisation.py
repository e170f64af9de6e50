"""Device-side mirror of the reference's waveform normaliser (R:src/data/preprocess/input_normalisation.py:38-94), the
step directly in front of the hot path.

``InputNormalizer2D.normalize(x, channel_wise=False)`` keeps the reference's signature and return triple for a single
``[1, N]`` / ``[F, N]`` tensor; ``normalize_batch`` is what a GPU input pipeline wants: a whole ``[B, N]`` batch of
utterances -- float32, or raw 16-bit PCM as stored in the wav files (half the host-to-device bytes) -- standardised per
utterance by one kernel launch.  ``channel_wise=True`` (per-feature statistics of spectrogram inputs) is not on the
waveform path and stays the reference's tensor expression."""
from __future__ import annotations

import torch

from ... import ops


class InputNormalizer2D:
    def __init__(self, normalize_over_channels: bool = True):
        self.channel_wise = normalize_over_channels

    @staticmethod
    def normalize(spectogram: torch.Tensor, channel_wise: bool):
        if len(spectogram.shape) != 2:
            raise ValueError("expect to normalize over 2D input")
        if channel_wise or not spectogram.is_cuda:
            std, mean = torch.std_mean(spectogram, dim=0) if channel_wise else torch.std_mean(spectogram)
            return (spectogram - mean) / (std + 1e-5), mean, std
        flat = spectogram.reshape(1, -1)
        out, mean, std = ops.normalize_wav(flat)
        return out.view_as(spectogram), mean[0], std[0]

    @staticmethod
    def normalize_batch(wav: torch.Tensor) -> torch.Tensor:
        """[B, N] float32 or int16 PCM on the GPU -> float32 [B, N], each utterance (x - mean) / (std + 1e-5)."""
        return ops.normalize_wav(wav)[0]
