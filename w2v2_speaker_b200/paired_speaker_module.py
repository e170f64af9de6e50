"""Sibling head on the same encoder: the reference's paired-input speaker model, a framework-free
mirror of ``Wav2vec2PairedSpeakerModule`` (R:src/lightning_modules/speaker/wav2vec2_paired_input.py:26-207) and of
the forward protocol of its base class (R:src/lightning_modules/speaker/paired_speaker_recognition_module.py:62-87).

Two utterances go through the CNN feature extractor and the feature projection separately, are joined into ONE
sequence ``[CLS, frames_1, SEP, frames_2, SEP]`` (constant-valued tokens), the transformer stack runs on that
sequence and a Linear(H -> 1) on the CLS position scores "same speaker".  The split call path
(``model.feature_extractor`` / ``model.feature_projection`` / ``model.encoder``) runs on the same kernels as the
fused one and is differentiable (training.FeatureExtractorFn / FeatureProjectionFn / EncoderStackFn); SpecAugment is
not applied on it, as upstream (it lives in ``Wav2Vec2Model.forward``, which this path bypasses).

    module = Wav2vec2PairedSpeakerModule(cfg, BinaryCrossEntropyLoss).cuda()
    scores = module(wav_a, wav_b)                       # [B, 1]
    loss, prediction = module.loss_fn(scores, same_speaker_labels)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.nn as nn

from .models.wav2vec2 import Wav2Vec2RegularisationConfig, Wav2Vec2WrapperModule


@dataclass
class Wav2vec2PairedSpeakerModuleConfig:
    """R:src/lightning_modules/speaker/wav2vec2_paired_input.py:26-62 (same field names; defaults follow
    R:config/network/wav2vec2_paired.yaml where it sets them)."""
    wav2vec_hunggingface_id: str = "facebook/wav2vec2-base"
    reset_weights: bool = False
    wav2vec_initially_frozen: bool = False
    num_frozen_steps: Optional[int] = 10000
    completely_freeze_feature_extractor: bool = True
    completely_freeze_feature_projector: bool = False
    activation_dropout: float = 0.0
    attention_dropout: float = 0.1
    feat_proj_dropout: float = 0.1
    hidden_dropout: float = 0.1
    layerdrop: float = 0.05
    mask_feature_length: int = 10
    mask_feature_prob: float = 0.0
    mask_time_length: int = 10
    mask_time_prob: float = 0.05
    final_channel_mask_prob: float = 0.0
    final_channel_mask_width: int = 5
    # class attributes upstream (no annotation => not dataclass fields): the token fill values
    cls_token_constant = 1
    sep_token_constant = -1


class Wav2vec2PairedSpeakerModule(nn.Module):
    def __init__(self, cfg: Wav2vec2PairedSpeakerModuleConfig, loss_fn_constructor: Callable[[], nn.Module]):
        super().__init__()
        self.cfg = cfg
        self.loss_fn = loss_fn_constructor()
        self.wav2vec = Wav2Vec2WrapperModule(
            wav2vec2_huggingface_id=cfg.wav2vec_hunggingface_id, reset_weights=cfg.reset_weights,
            reg_cfg=Wav2Vec2RegularisationConfig(
                gradient_checkpointing=False, activation_dropout=cfg.activation_dropout,
                attention_dropout=cfg.attention_dropout, feat_proj_dropout=cfg.feat_proj_dropout,
                hidden_dropout=cfg.hidden_dropout, layerdrop=cfg.layerdrop,
                mask_feature_length=cfg.mask_feature_length, mask_feature_prob=cfg.mask_feature_prob,
                mask_time_length=cfg.mask_time_length, mask_time_prob=cfg.mask_time_prob))
        self._is_wav2vec_frozen = False
        self.linear = nn.Linear(in_features=self._get_wav2vec2_embedding_size(), out_features=1)
        self.steps = 0

    def _get_wav2vec2_embedding_size(self) -> int:
        if "base" in self.cfg.wav2vec_hunggingface_id:
            return 768
        if "large" in self.cfg.wav2vec_hunggingface_id:
            return 1024
        raise ValueError("unknown wav2ec2 embedding size")

    def generate_example_input(self, include_batch_dimension: bool, batch_size: Optional[int] = None):
        shape = [batch_size, 16000] if include_batch_dimension else [16000]
        return torch.rand(size=shape), torch.rand(size=shape)

    # ---- freeze protocol (R:...wav2vec2_paired_input.py:129-160) -------------------------------------------
    def _apply_permanent_freezes(self) -> None:
        if self.cfg.completely_freeze_feature_extractor:
            self.wav2vec.model.feature_extractor.requires_grad_(False)
        if self.cfg.completely_freeze_feature_projector:
            self.wav2vec.model.feature_projection.requires_grad_(False)

    def on_train_start(self) -> None:
        self.steps = 0
        if self.cfg.wav2vec_initially_frozen:
            self.wav2vec.freeze()
            self._is_wav2vec_frozen = True
        self._apply_permanent_freezes()

    def on_after_backward(self) -> None:
        self.steps += 1
        if (self._is_wav2vec_frozen and self.cfg.num_frozen_steps is not None
                and self.steps >= self.cfg.num_frozen_steps):
            self.wav2vec.unfreeze()
            self._is_wav2vec_frozen = False
            self._apply_permanent_freezes()

    # ---- forward (R:...wav2vec2_paired_input.py:162-207) ----------------------------------------------------
    def compute_speaker_equality(self, wav_tensor: torch.Tensor, other_wav_tensor: torch.Tensor) -> torch.Tensor:
        assert wav_tensor.shape[0] == other_wav_tensor.shape[0]
        assert wav_tensor.device == other_wav_tensor.device
        model = self.wav2vec.model
        features_1 = model.feature_extractor(wav_tensor).transpose(1, 2)
        features_2 = model.feature_extractor(other_wav_tensor).transpose(1, 2)
        features_1, _ = model.feature_projection(features_1)
        features_2, _ = model.feature_projection(features_2)
        # the reference hard-codes the token width to 768 (it only ever runs wav2vec2-base here)
        ones = torch.ones((wav_tensor.shape[0], 1, 768), device=wav_tensor.device)
        cls_token = ones * self.cfg.cls_token_constant
        sep_token = ones * self.cfg.sep_token_constant
        sequence = torch.cat([cls_token, features_1, sep_token, features_2, sep_token], dim=1)
        tokens = model.encoder(sequence).last_hidden_state
        return self.linear(tokens[:, 0, :])

    def forward(self, input_tensor: torch.Tensor, other_input_tensor: torch.Tensor) -> torch.Tensor:
        return self.compute_speaker_equality(input_tensor, other_input_tensor)
