"""CPU restatement (plain fp32/fp64 PyTorch ops) of the reference hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites the
reference (``R:`` = /root/reference) or the HuggingFace file that holds the arithmetic
(``HF:`` = transformers/models/wav2vec2/modeling_wav2vec2.py, transformers 5.5.0; the
reference pins ``transformers ^4.8.2``, R:pyproject.toml:40 -- eval-mode math is unchanged).

Layout convention here is the reference's: activations are ``[B, C, T]`` in the conv
stack and ``[B, T, H]`` in the transformer, exactly as HF produces them.

Eval-mode semantics only (dropout / LayerDrop / SpecAugment are identity in eval and are
switched off for training-parity runs).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .params import ArchConfig, BASE

P = Dict[str, torch.Tensor]

# Train-mode regularisation of the reference defaults (R:src/models/wav2vec2.py:83-94), used by bench.py's CPU
# train-step baseline only (parity runs keep it None = eval semantics): {"feat":, "hidden":, "attn":, "layerdrop":}
TRAIN_REG: Optional[dict] = None


def _drop(x: torch.Tensor, key: str) -> torch.Tensor:
    if TRAIN_REG is None or TRAIN_REG.get(key, 0.0) <= 0.0:
        return x
    return F.dropout(x, TRAIN_REG[key], training=True)


# --------------------------------------------------------------------------------------
# feature extractor: HF:409-419 (Wav2Vec2FeatureEncoder.forward)


def conv_layer0(wav: torch.Tensor, p: P, arch: ArchConfig = BASE) -> torch.Tensor:
    """Wav2Vec2GroupNormConvLayer (HF:302-323): Conv1d(1->C,k=10,s=5,no bias) ->
    GroupNorm(C groups, C channels, eps 1e-5, affine) -> exact GELU.  wav [B,N] -> [B,C,L0]."""
    h = F.conv1d(wav[:, None, :], p["feature_extractor.conv_layers.0.conv.weight"],
                 stride=arch.conv_stride[0])
    h = F.group_norm(h, arch.conv_dim, p["feature_extractor.conv_layers.0.layer_norm.weight"],
                     p["feature_extractor.conv_layers.0.layer_norm.bias"], eps=arch.eps)
    return F.gelu(h)


def conv_layer(h: torch.Tensor, i: int, p: P, arch: ArchConfig = BASE) -> torch.Tensor:
    """Wav2Vec2NoLayerNormConvLayer (HF:254-272): Conv1d(C->C,k,s=2,no bias) -> exact GELU."""
    h = F.conv1d(h, p[f"feature_extractor.conv_layers.{i}.conv.weight"], stride=arch.conv_stride[i])
    return F.gelu(h)


def conv_layer_ln(h: torch.Tensor, i: int, p: P, arch: ArchConfig) -> torch.Tensor:
    """Wav2Vec2LayerNormConvLayer (HF:275-299; the "-lv60" / XLSR feature extractor, every layer): Conv1d(+bias) ->
    LayerNorm over the channels (transpose, LayerNorm(C), transpose back) -> exact GELU."""
    pre = f"feature_extractor.conv_layers.{i}."
    h = F.conv1d(h, p[pre + "conv.weight"], p.get(pre + "conv.bias"), stride=arch.conv_stride[i])
    h = F.layer_norm(h.transpose(1, 2), (arch.conv_dim,), p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"],
                     arch.eps).transpose(1, 2)
    return F.gelu(h)


def feature_extractor(wav: torch.Tensor, p: P, arch: ArchConfig = BASE,
                      stages: Optional[list] = None) -> torch.Tensor:
    layer_mode = arch.feat_extract_norm == "layer"
    h = conv_layer_ln(wav[:, None, :], 0, p, arch) if layer_mode else conv_layer0(wav, p, arch)
    if stages is not None:
        stages.append(h)
    for i in range(1, len(arch.conv_kernel)):
        h = conv_layer_ln(h, i, p, arch) if layer_mode else conv_layer(h, i, p, arch)
        if stages is not None:
            stages.append(h)
    return h  # [B, C, T]


# --------------------------------------------------------------------------------------
# feature projection: HF:429-434


def feature_projection(feat_btc: torch.Tensor, p: P, arch: ArchConfig = BASE) -> torch.Tensor:
    """LayerNorm(C) -> Linear(C->H) (-> dropout, identity in eval).  [B,T,C] -> [B,T,H]."""
    n = F.layer_norm(feat_btc, (arch.conv_dim,), p["feature_projection.layer_norm.weight"],
                     p["feature_projection.layer_norm.bias"], arch.eps)
    return _drop(F.linear(n, p["feature_projection.projection.weight"], p["feature_projection.projection.bias"]), "feat")


# --------------------------------------------------------------------------------------
# positional conv embedding: HF:326-379


def pos_conv_weight(p: P) -> torch.Tensor:
    """weight_norm over dim=2: w[o,i,k] = g[0,0,k] * v[o,i,k] / ||v[:,:,k]||_F (HF:340-358)."""
    g = p["encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    v = p["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    return g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()


def pos_conv_embed(h_bth: torch.Tensor, p: P, arch: ArchConfig = BASE) -> torch.Tensor:
    """Conv1d(H->H,k=128,pad=64,groups=16,bias) -> drop last frame (even kernel, HF:371-379) -> GELU."""
    x = h_bth.transpose(1, 2)
    y = F.conv1d(x, pos_conv_weight(p), p["encoder.pos_conv_embed.conv.bias"],
                 padding=arch.pos_kernel // 2, groups=arch.pos_groups)
    if arch.pos_kernel % 2 == 0:
        y = y[:, :, :-1]
    return F.gelu(y).transpose(1, 2)


# --------------------------------------------------------------------------------------
# transformer: HF:668-727 (encoder), HF:592-609 (layer), HF:500-549 (attention), HF:566-573 (FFN)


def attention(h: torch.Tensor, l: int, p: P, arch: ArchConfig = BASE) -> torch.Tensor:
    """softmax(q k^T * d^-0.5) v with no mask (the reference never passes attention_mask,
    R:src/models/wav2vec2.py:71); eager form HF:438-463."""
    B, T, H = h.shape
    nh, d = arch.heads, H // arch.heads
    pre = f"encoder.layers.{l}.attention."
    q = F.linear(h, p[pre + "q_proj.weight"], p[pre + "q_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    k = F.linear(h, p[pre + "k_proj.weight"], p[pre + "k_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    v = F.linear(h, p[pre + "v_proj.weight"], p[pre + "v_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * (d ** -0.5)
    a = _drop(torch.softmax(s, dim=-1), "attn")
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, T, H)
    return F.linear(o, p[pre + "out_proj.weight"], p[pre + "out_proj.bias"])


def encoder_layer(h: torch.Tensor, l: int, p: P, arch: ArchConfig = BASE) -> torch.Tensor:
    """Post-LN block (HF:592-609): h = LN1(h + attn(h)); h = LN2(h + FFN(h))."""
    pre = f"encoder.layers.{l}."
    h = h + _drop(attention(h, l, p, arch), "hidden")
    h = F.layer_norm(h, (arch.hidden,), p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"], arch.eps)
    f = F.gelu(F.linear(h, p[pre + "feed_forward.intermediate_dense.weight"],
                        p[pre + "feed_forward.intermediate_dense.bias"]))
    f = F.linear(f, p[pre + "feed_forward.output_dense.weight"], p[pre + "feed_forward.output_dense.bias"])
    h = F.layer_norm(h + _drop(f, "hidden"), (arch.hidden,), p[pre + "final_layer_norm.weight"],
                     p[pre + "final_layer_norm.bias"], arch.eps)
    return h


def encoder_layer_stable(h: torch.Tensor, l: int, p: P, arch: ArchConfig) -> torch.Tensor:
    """Pre-LN block of the stable-layer-norm models (Wav2Vec2EncoderLayerStableLayerNorm, HF:632-655):
    h = h + attn(LN1(h)); h = h + FFN(LN2(h))."""
    pre = f"encoder.layers.{l}."
    a = F.layer_norm(h, (arch.hidden,), p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"], arch.eps)
    h = h + _drop(attention(a, l, p, arch), "hidden")
    c = F.layer_norm(h, (arch.hidden,), p[pre + "final_layer_norm.weight"], p[pre + "final_layer_norm.bias"], arch.eps)
    f = F.gelu(F.linear(c, p[pre + "feed_forward.intermediate_dense.weight"],
                        p[pre + "feed_forward.intermediate_dense.bias"]))
    f = F.linear(f, p[pre + "feed_forward.output_dense.weight"], p[pre + "feed_forward.output_dense.bias"])
    return h + _drop(f, "hidden")


def encoder_stable(h: torch.Tensor, p: P, arch: ArchConfig, hidden_states: Optional[list] = None) -> torch.Tensor:
    """Wav2Vec2EncoderStableLayerNorm.forward (HF:731-799), eval mode: h = h + pos_conv(h); L x pre-LN layer; final
    LayerNorm.  hidden_states collects the input of every layer and the normalised output, as HF does."""
    h = _drop(h + pos_conv_embed(h, p, arch), "hidden")
    for l in range(arch.layers):
        if hidden_states is not None:
            hidden_states.append(h)
        if TRAIN_REG is not None and torch.rand([]).item() < TRAIN_REG.get("layerdrop", 0.0):
            continue
        h = encoder_layer_stable(h, l, p, arch)
    h = F.layer_norm(h, (arch.hidden,), p["encoder.layer_norm.weight"], p["encoder.layer_norm.bias"], arch.eps)
    if hidden_states is not None:
        hidden_states.append(h)
    return h


def encoder(h: torch.Tensor, p: P, arch: ArchConfig = BASE, hidden_states: Optional[list] = None) -> torch.Tensor:
    """Wav2Vec2Encoder.forward (HF:668-727), eval mode: h = LN(h + pos_conv(h)); 12 x layer."""
    if arch.stable_layer_norm:
        return encoder_stable(h, p, arch, hidden_states)
    h = h + pos_conv_embed(h, p, arch)
    h = _drop(F.layer_norm(h, (arch.hidden,), p["encoder.layer_norm.weight"], p["encoder.layer_norm.bias"], arch.eps),
              "hidden")
    if hidden_states is not None:
        hidden_states.append(h)
    for l in range(arch.layers):
        if TRAIN_REG is not None and torch.rand([]).item() < TRAIN_REG.get("layerdrop", 0.0):
            continue                                        # LayerDrop (HF:701-713)
        h = encoder_layer(h, l, p, arch)
        if hidden_states is not None:
            hidden_states.append(h)
    return h


def wav2vec2_forward(wav: torch.Tensor, p: P, arch: ArchConfig = BASE, trace: Optional[dict] = None) -> torch.Tensor:
    """Wav2Vec2Model.forward (HF:1327-1383), eval: returns last_hidden_state [B,T,H].
    The reference wrapper then returns its transpose [B,H,T] (R:src/models/wav2vec2.py:62-76)."""
    stages = [] if trace is not None else None
    feat = feature_extractor(wav, p, arch, stages)
    proj = feature_projection(feat.transpose(1, 2), p, arch)
    hs = [] if trace is not None else None
    out = encoder(proj, p, arch, hs)
    if trace is not None:
        trace["conv"] = stages
        trace["proj"] = proj
        trace["hidden_states"] = hs
    return out


# --------------------------------------------------------------------------------------
# pooling: R:src/layers/pooling.py  (input [B,T,C], dim_to_reduce=1 as built at
# R:src/lightning_modules/speaker/wav2vec2_fc.py:238-272)


def mean_pool(x: torch.Tensor) -> torch.Tensor:
    """MeanStatPool1D (R:src/layers/pooling.py:24-30)."""
    return torch.mean(x, 1)


def mean_std_pool(x: torch.Tensor) -> torch.Tensor:
    """MeanStdStatPool1D (R:src/layers/pooling.py:38-44): cat(std_mean) => [std(unbiased) || mean]."""
    return torch.cat(torch.std_mean(x, 1), 1)


def max_pool(x: torch.Tensor) -> torch.Tensor:
    """MaxPool1D (R:src/layers/pooling.py:74-80)."""
    return torch.max(x, dim=1).values


def quantile_pool(x: torch.Tensor) -> torch.Tensor:
    """QuantilePool1D (R:src/layers/pooling.py:51-67)."""
    q = torch.quantile(x, torch.tensor([0, 0.25, 0.5, 0.75, 1.0]), dim=1)
    return torch.flatten(q.transpose(0, 1), 1, 2)


def attentive_stat_pool(x_btc: torch.Tensor, asp: P, training: bool = False,
                        bn_eps: float = 1e-5) -> torch.Tensor:
    """speechbrain 0.5.x ``AttentiveStatisticsPooling(channels, attention_channels=128,
    global_context=True)`` as called from AttentiveStatPool1D (R:src/layers/pooling.py:87-106)
    with dim_to_reduce=1 (input transposed to [N,C,L]).  Published algorithm (ECAPA_TDNN.py):

        mean, std = stats(x, 1/L);   stats(x,m): mean = sum(m*x), std = sqrt(clamp(sum(m*(x-mean)^2), eps=1e-12))
        attn = cat[x, mean.repeat(L), std.repeat(L)]                       (3C channels)
        attn = conv(tanh(tdnn(attn)));  tdnn = Conv1d(3C,128,1) -> ReLU -> BatchNorm1d(128)
        attn = softmax(attn, dim=L);  mean, std = stats(x, attn);  out = cat[mean, std]  -> [N, 2C]

    Source unavailable offline; pinned against transformers' port of the same layer except for the BatchNorm position,
    which stays "parity unpinned" (oracle/__init__.py, tests/test_oracle_golden.py)."""
    x = x_btc.transpose(1, 2)                                   # [N, C, L]
    L = x.shape[-1]

    def stats(x, m, eps=1e-12):
        mean = (m * x).sum(2)
        std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
        return mean, std

    m0 = torch.full((x.shape[0], 1, L), 1.0 / L, dtype=x.dtype)
    mean, std = stats(x, m0)
    a = torch.cat([x, mean.unsqueeze(2).repeat(1, 1, L), std.unsqueeze(2).repeat(1, 1, L)], dim=1)
    a = F.conv1d(a, asp["tdnn.conv.conv.weight"], asp["tdnn.conv.conv.bias"])
    a = F.relu(a)
    if training:
        a = F.batch_norm(a, None, None, asp["tdnn.norm.norm.weight"], asp["tdnn.norm.norm.bias"],
                         training=True, eps=bn_eps)
    else:
        a = F.batch_norm(a, asp["tdnn.norm.norm.running_mean"], asp["tdnn.norm.norm.running_var"],
                         asp["tdnn.norm.norm.weight"], asp["tdnn.norm.norm.bias"], training=False, eps=bn_eps)
    a = torch.tanh(a)
    a = F.conv1d(a, asp["conv.conv.weight"], asp["conv.conv.bias"])
    a = F.softmax(a, dim=2)
    mean, std = stats(x, a)
    return torch.cat([mean, std], dim=1)                        # [N, 2C]


def split_path_forward(wavs, p: P, tokens, arch: ArchConfig = BASE) -> torch.Tensor:
    """The encoder on a caller-built sequence, as the reference's CLS-token wrapper (R:src/models/wav2vec2.py:128-140,
    one utterance, tokens = [cls]) and its paired-input model (R:src/lightning_modules/speaker/wav2vec2_paired_input.py:
    162-207, two utterances, tokens = [cls, sep, sep]) run it: every utterance goes through the CNN and the feature
    projection on its own, a constant-valued token row precedes each of them, the remaining tokens close the
    sequence, and the transformer stack runs on the concatenation.  -> last_hidden_state [B, T', H]."""
    B = wavs[0].shape[0]
    parts = []
    for i, wav in enumerate(wavs):
        parts.append(torch.ones(B, 1, arch.hidden) * tokens[i])
        parts.append(feature_projection(feature_extractor(wav, p, arch).transpose(1, 2), p, arch))
    for t in tokens[len(wavs):]:
        parts.append(torch.ones(B, 1, arch.hidden) * t)
    return encoder(torch.cat(parts, dim=1), p, arch)


# --------------------------------------------------------------------------------------
# heads / losses


def cross_entropy_head(emb: torch.Tensor, fc_w: torch.Tensor, fc_b: torch.Tensor,
                       labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Last fc_list Linear (R:.../wav2vec2_fc.py:199-210) + CrossEntropyLoss
    (R:src/optim/loss/cross_entropy.py:19-33): returns (logits, loss, softmax)."""
    logits = F.linear(emb, fc_w, fc_b)
    return logits, F.cross_entropy(logits, labels), F.softmax(logits, dim=1)


def aam_softmax(x: torch.Tensor, fc_weights: torch.Tensor, labels: torch.Tensor,
                margin: float = 0.2, scale: float = 30.0, easy_margin: bool = False):
    """AngularAdditiveMarginSoftMaxLoss.forward (R:src/optim/loss/aam_softmax.py:50-74).
    Returns (scaled margin logits, loss, softmax)."""
    cos_m, sin_m = math.cos(margin), math.sin(margin)
    th, mm = math.cos(math.pi - margin), math.sin(math.pi - margin) * margin
    cosine = F.linear(F.normalize(x), F.normalize(fc_weights))
    sine = torch.sqrt((1.0 - cosine * cosine).clamp(0, 1))
    phi = cosine * cos_m - sine * sin_m
    if easy_margin:
        phi = torch.where(cosine > 0, phi, cosine)
    else:
        phi = torch.where((cosine - th) > 0, phi, cosine - mm)
    one_hot = torch.zeros_like(cosine)
    one_hot.scatter_(1, labels.view(-1, 1), 1)
    out = ((one_hot * phi) + ((1.0 - one_hot) * cosine)) * scale
    return out, F.cross_entropy(out, labels), F.softmax(out, dim=1)


# --------------------------------------------------------------------------------------
# the composed path: R:src/lightning_modules/speaker/wav2vec2_fc.py:414-438


def speaker_embedding(wav: torch.Tensor, p: P, pooling: str = "mean", asp: Optional[P] = None,
                      arch: ArchConfig = BASE, trace: Optional[dict] = None) -> torch.Tensor:
    """compute_speaker_embedding (R:.../wav2vec2_fc.py:414-431): encoder -> [B,T,H] -> pool ->
    EmbeddingMasker (identity on this path)."""
    h = wav2vec2_forward(wav, p, arch, trace)
    if pooling == "mean":
        return mean_pool(h)
    if pooling == "mean+std":
        return mean_std_pool(h)
    if pooling == "attentive":
        return attentive_stat_pool(h, asp)
    if pooling == "max":
        return max_pool(h)
    if pooling == "quantile":
        return quantile_pool(h)
    raise ValueError(pooling)
