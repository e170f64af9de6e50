"""Deterministic parameter sets for the oracle, the golden fixtures and the CUDA path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parameters are keyed by the
HuggingFace ``Wav2Vec2Model.state_dict()`` names (transformers 5.x naming, i.e. the
pos-conv weight-norm pair is ``parametrizations.weight.original0/1``) so the very
same dict can be ``load_state_dict``-ed into the reference's wrapped HF model
(R:src/models/wav2vec2.py:108) and fed to the oracle / the CUDA engine.

Distributions follow HF ``_init_weights`` (HF:968-1003) for the big matrices so
activations stay O(1) through the stack; norm affine parameters and biases are
perturbed away from (1, 0) so that parity tests are sensitive to them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch


@dataclass(frozen=True)
class ArchConfig:
    """Architecture numbers of the encoder (HF Wav2Vec2Config fields we depend on)."""
    name: str = "base"
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    ffn: int = 3072
    conv_dim: int = 512
    conv_kernel: tuple = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: tuple = (5, 2, 2, 2, 2, 2, 2)
    pos_kernel: int = 128
    pos_groups: int = 16
    eps: float = 1e-5
    # the "-lv60" / XLSR family (HF configs with feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True):
    feat_extract_norm: str = "group"      # "layer": every conv layer is Conv1d(+bias) -> LayerNorm(C) -> GELU (HF:275-299)
    conv_bias: bool = False
    stable_layer_norm: bool = False       # pre-LN encoder layers + final encoder LayerNorm (HF:632-655, HF:731-799)

    def conv_lengths(self, n: int):
        out = []
        for k, s in zip(self.conv_kernel, self.conv_stride):
            n = (n - k) // s + 1          # HF:1012-1018
            out.append(n)
        return out


BASE = ArchConfig()
LARGE = ArchConfig(name="large", hidden=1024, layers=24, heads=16, ffn=4096)
LARGE_LV60 = ArchConfig(name="large-lv60", hidden=1024, layers=24, heads=16, ffn=4096, feat_extract_norm="layer",
                        conv_bias=True, stable_layer_norm=True)
# a tiny architecture for fast CPU tests of host logic (same structure, fewer/lighter layers)
TINY = ArchConfig(name="tiny", hidden=128, layers=2, heads=2, ffn=256, conv_dim=64,
                  pos_kernel=16, pos_groups=4)


def arch_from_id(huggingface_id: str) -> ArchConfig:
    """Size detection by substring, as the reference does (R:src/models/wav2vec2.py:112-117)."""
    if "base" in huggingface_id:
        return BASE
    if "large" in huggingface_id:
        return LARGE_LV60 if ("lv60" in huggingface_id or "xlsr" in huggingface_id) else LARGE
    if "tiny" in huggingface_id:
        return TINY
    raise ValueError("cannot determine num features")


def make_params(arch: ArchConfig = BASE, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def uni(*shape, lo=-1.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    C, H = arch.conv_dim, arch.hidden
    p: Dict[str, torch.Tensor] = {}
    p["masked_spec_embed"] = uni(H, lo=0.0, hi=1.0)
    cin = 1
    for i, k in enumerate(arch.conv_kernel):
        # kaiming_normal_, fan_in, gain sqrt(2)  (HF:998-1003)
        p[f"feature_extractor.conv_layers.{i}.conv.weight"] = randn(C, cin, k, std=math.sqrt(2.0 / (cin * k)))
        if arch.conv_bias:
            p[f"feature_extractor.conv_layers.{i}.conv.bias"] = randn(C, std=0.05)
        if i == 0 or arch.feat_extract_norm == "layer":
            p[f"feature_extractor.conv_layers.{i}.layer_norm.weight"] = 1.0 + randn(C, std=0.1)
            p[f"feature_extractor.conv_layers.{i}.layer_norm.bias"] = randn(C, std=0.1)
        cin = C
    p["feature_projection.layer_norm.weight"] = 1.0 + randn(C, std=0.1)
    p["feature_projection.layer_norm.bias"] = randn(C, std=0.1)
    kk = math.sqrt(1.0 / C)
    p["feature_projection.projection.weight"] = uni(H, C, lo=-kk, hi=kk)
    p["feature_projection.projection.bias"] = uni(H, lo=-kk, hi=kk)
    gsz = H // arch.pos_groups
    p["encoder.pos_conv_embed.conv.bias"] = randn(H, std=0.02)
    v = randn(H, gsz, arch.pos_kernel, std=2.0 * math.sqrt(1.0 / (arch.pos_kernel * H)))
    p["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
    # weight-norm gain initialised to the norm of v over dims (0,1) (as torch weight_norm does),
    # then perturbed so that g != ||v|| is exercised
    p["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = (
        v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt() * (1.0 + randn(1, 1, arch.pos_kernel, std=0.05)))
    p["encoder.layer_norm.weight"] = 1.0 + randn(H, std=0.1)
    p["encoder.layer_norm.bias"] = randn(H, std=0.1)
    for l in range(arch.layers):
        pre = f"encoder.layers.{l}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            p[pre + f"attention.{nm}.weight"] = randn(H, H, std=0.02)
            p[pre + f"attention.{nm}.bias"] = randn(H, std=0.02)
        p[pre + "layer_norm.weight"] = 1.0 + randn(H, std=0.1)
        p[pre + "layer_norm.bias"] = randn(H, std=0.1)
        p[pre + "feed_forward.intermediate_dense.weight"] = randn(arch.ffn, H, std=0.02)
        p[pre + "feed_forward.intermediate_dense.bias"] = randn(arch.ffn, std=0.02)
        p[pre + "feed_forward.output_dense.weight"] = randn(H, arch.ffn, std=0.02)
        p[pre + "feed_forward.output_dense.bias"] = randn(H, std=0.02)
        p[pre + "final_layer_norm.weight"] = 1.0 + randn(H, std=0.1)
        p[pre + "final_layer_norm.bias"] = randn(H, std=0.1)
    return p


def make_head_params(embed: int, num_speakers: int, seed: int = 1) -> Dict[str, torch.Tensor]:
    """CE head Linear(E->S) (R:.../wav2vec2_fc.py:199-210), AAM ``fc_weights[S,E]`` (xavier normal,
    R:src/optim/loss/aam_softmax.py:33-37) and the speechbrain ASP parameters."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    bound = 1.0 / math.sqrt(embed)
    p["fc.weight"] = (torch.rand(num_speakers, embed, generator=g) * 2 - 1) * bound
    p["fc.bias"] = (torch.rand(num_speakers, generator=g) * 2 - 1) * bound
    p["aam.fc_weights"] = torch.randn(num_speakers, embed, generator=g) * math.sqrt(2.0 / (embed + num_speakers))
    return p


def make_asp_params(channels: int, attention_channels: int = 128, seed: int = 2) -> Dict[str, torch.Tensor]:
    """Parameters of speechbrain ``AttentiveStatisticsPooling(channels)`` with speechbrain's
    state_dict names (tdnn = TDNNBlock(3C -> 128, k=1), conv = Conv1d(128 -> C, k=1))."""
    g = torch.Generator().manual_seed(seed)
    A, C = attention_channels, channels

    def uni(*shape, bound):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    p = {}
    b1 = 1.0 / math.sqrt(3 * C)
    p["tdnn.conv.conv.weight"] = uni(A, 3 * C, 1, bound=b1)
    p["tdnn.conv.conv.bias"] = uni(A, bound=b1)
    p["tdnn.norm.norm.weight"] = 1.0 + torch.randn(A, generator=g) * 0.1
    p["tdnn.norm.norm.bias"] = torch.randn(A, generator=g) * 0.1
    p["tdnn.norm.norm.running_mean"] = torch.randn(A, generator=g) * 0.1
    p["tdnn.norm.norm.running_var"] = 1.0 + torch.rand(A, generator=g) * 0.5
    b2 = 1.0 / math.sqrt(A)
    p["conv.conv.weight"] = uni(C, A, 1, bound=b2)
    p["conv.conv.bias"] = uni(C, bound=b2)
    return p


def make_inputs(batch: int, num_samples: int, num_speakers: int = 5994, seed: int = 1234):
    """Synthetic batch as SURVEY 8(d): randn waveform standardised per utterance exactly like
    InputNormalizer2D.normalize(channel_wise=False) (R:src/data/preprocess/input_normalisation.py:53-67)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, num_samples, generator=g, dtype=torch.float32)
    mean = x.mean(dim=1, keepdim=True)
    std = x.std(dim=1, keepdim=True)
    x = (x - mean) / (std + 1e-5)
    labels = torch.randint(0, num_speakers, (batch,), generator=g, dtype=torch.int64)
    return x, labels
