"""The caller of the hot path: a framework-free mirror of the reference's ``Wav2vec2FCModule``
forward protocol (R:src/lightning_modules/speaker/wav2vec2_fc.py:48-98, 101-236, 363-438 and
R:src/lightning_modules/speaker/speaker_recognition_module.py:109-130, 207-220).

The reference's Lightning module itself is outside the hot path ("boundary, kept
as-is"); this class exists so that the path can be driven, tested and benchmarked end to end through
one public call without pytorch_lightning / hydra (absent from this image):

    module = Wav2vec2FCModule(cfg, num_speakers, loss_fn_constructor).cuda().eval()
    embedding, prediction = module(wav)                 # forward()  (speaker_recognition_module.py:121-130)
    loss, softmax = module.loss_fn(prediction, labels)  # _train_step_ce_loss (:207-220)

Same config field names, pooling dispatch, AAM head surgery and squeeze semantics as upstream.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch
import torch.nn as nn

from .layers.embedding_masking import EmbeddingMasker
from .layers.linear import SpeakerLinear
from .layers.pooling import (AttentiveStatPool1D, IndexPool1D, MaxPool1D, MeanStatPool1D, MeanStdStatPool1D,
                             NoPooling, QuantilePool1D)
from .models.wav2vec2 import Wav2Vec2RegularisationConfig, Wav2Vec2WrapperModule, Wav2vecLiteWrapperModule
from .optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss


@dataclass
class Wav2vec2FCModuleConfig:
    """R:src/lightning_modules/speaker/wav2vec2_fc.py:48-98 (same field names; defaults follow
    R:config/network/wav2vec2_fc.yaml)."""
    wav2vec_hunggingface_id: str = "facebook/wav2vec2-base"
    reset_weights: bool = False
    wav2vec_feature_encoder_only: bool = False
    wav2vec_initially_frozen: bool = False
    num_frozen_steps: Optional[int] = None
    completely_freeze_feature_extractor: bool = True
    hidden_fc_layers_out: List[int] = field(default_factory=list)
    embedding_layer_idx: int = -1
    stat_pooling_type: str = "mean"
    test_stat_pooling_type: str = "mean"
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
    explicit_stat_pool_embedding_size: Optional[int] = None
    explicit_num_speakers: Optional[int] = None
    use_transformers_as_ensembles: bool = False
    num_ensembles: int = 12


class Wav2vec2FCModule(nn.Module):
    def __init__(self, cfg: Wav2vec2FCModuleConfig, num_speakers: int,
                 loss_fn_constructor: Callable[[], nn.Module]):
        super().__init__()
        self.cfg = cfg
        self.num_speakers = num_speakers
        if cfg.completely_freeze_feature_extractor and cfg.wav2vec_feature_encoder_only:
            raise ValueError("can not freeze the whole network! Either `completely_freeze_feature_extractor` or "
                             "`wav2vec_feature_encoder_only` need to be set to False")
        self.loss_fn = loss_fn_constructor()
        if cfg.wav2vec_feature_encoder_only:
            self.wav2vec = Wav2vecLiteWrapperModule(cfg.wav2vec_hunggingface_id, cfg.reset_weights)
        else:
            self.wav2vec = Wav2Vec2WrapperModule(
                wav2vec2_huggingface_id=cfg.wav2vec_hunggingface_id, reset_weights=cfg.reset_weights,
                reg_cfg=Wav2Vec2RegularisationConfig(
                    gradient_checkpointing=False, activation_dropout=cfg.activation_dropout,
                    attention_dropout=cfg.attention_dropout, feat_proj_dropout=cfg.feat_proj_dropout,
                    hidden_dropout=cfg.hidden_dropout, layerdrop=cfg.layerdrop,
                    mask_feature_length=cfg.mask_feature_length, mask_feature_prob=cfg.mask_feature_prob,
                    mask_time_length=cfg.mask_time_length, mask_time_prob=cfg.mask_time_prob),
                insert_clc_token=cfg.stat_pooling_type == "first+cls")
        self.embedding_masker = EmbeddingMasker(timestep_mask_prob=0, timestep_mask_width=1,
                                                channel_mask_prob=cfg.final_channel_mask_prob,
                                                channel_mask_width=cfg.final_channel_mask_width,
                                                time_dim=2, embedding_dim=1)
        self.stat_pooling = self._determine_pooling_layer(cfg.stat_pooling_type, only_at_test_time=False)
        self.stat_pool_dimension = self._determine_stat_pool_embedding_size()
        if cfg.test_stat_pooling_type != cfg.stat_pooling_type:
            self.test_stat_pooling = self._determine_pooling_layer(cfg.test_stat_pooling_type, only_at_test_time=True)
        else:
            self.test_stat_pooling = self.stat_pooling
        outs = list(cfg.hidden_fc_layers_out)
        self.fc_list = nn.ModuleList([
            nn.Sequential(SpeakerLinear(self.stat_pool_dimension if i == 0 else outs[i - 1], n), nn.ReLU())
            for i, n in enumerate(outs)])
        self.fc_list.append(nn.Sequential(SpeakerLinear(
            self.stat_pool_dimension if not outs else outs[-1],
            cfg.explicit_num_speakers if cfg.explicit_num_speakers else num_speakers)))
        if isinstance(self.loss_fn, AngularAdditiveMarginSoftMaxLoss):
            # R:.../wav2vec2_fc.py:212-224: drop the last FC, rebuild AAM with the right feature sizes
            del self.fc_list[-1]
            self.loss_fn = AngularAdditiveMarginSoftMaxLoss(
                input_features=self.stat_pool_dimension,
                output_features=cfg.explicit_num_speakers if cfg.explicit_num_speakers is not None else num_speakers,
                margin=self.loss_fn.margin, scale=self.loss_fn.scale)

        self._is_wav2vec_frozen = False
        self.steps = 0
        self.test_with_ensemble = cfg.use_transformers_as_ensembles

    # ---- freeze protocol (R:.../wav2vec2_fc.py:339-361): which parameters get gradients, and with
    # them which backward kernels run (frozen encoder: the heads' only; frozen CNN: everything behind it) -----------
    def on_train_start(self) -> None:
        self.steps = 0
        if self.cfg.wav2vec_initially_frozen:
            self.wav2vec.freeze()
            self._is_wav2vec_frozen = True
        if self.cfg.completely_freeze_feature_extractor:
            self.wav2vec.model.feature_extractor.requires_grad_(False)

    def on_after_backward(self) -> None:
        self.steps += 1
        if (self._is_wav2vec_frozen and self.cfg.num_frozen_steps is not None
                and self.steps >= self.cfg.num_frozen_steps):
            self.wav2vec.unfreeze()
            self._is_wav2vec_frozen = False
            if self.cfg.completely_freeze_feature_extractor:
                self.wav2vec.model.feature_extractor.requires_grad_(False)

    def generate_example_input(self, include_batch_dimension: bool, batch_size: Optional[int] = None):
        # R:.../wav2vec2_fc.py:321-337: one second of uniform noise
        return torch.rand(size=[batch_size, 16000] if include_batch_dimension else [16000])

    # R:.../wav2vec2_fc.py:238-272
    def _determine_pooling_layer(self, stat_pooling_type: str, only_at_test_time: bool):
        if stat_pooling_type == "mean":
            return MeanStatPool1D(dim_to_reduce=1)
        if stat_pooling_type == "mean+std":
            return MeanStdStatPool1D(dim_to_reduce=1)
        if stat_pooling_type == "attentive":
            if only_at_test_time:
                raise ValueError("attention can not be learned at test time")
            return AttentiveStatPool1D(dim_to_reduce=1, embedding_size=self.wav2vec.num_features)
        if stat_pooling_type == "quantile":
            return QuantilePool1D(dim_to_reduce=1)
        if stat_pooling_type in ["first", "first+cls", "last", "middle", "random"]:
            return IndexPool1D(selection_method=self.cfg.stat_pooling_type, dim_to_reduce=1)
        if stat_pooling_type == "max":
            return MaxPool1D(dim_to_reduce=1)
        if stat_pooling_type.lower() == "none":
            return NoPooling()
        raise ValueError(f"unknown value {stat_pooling_type=}, should be one of ['mean', 'mean+std', 'attentive', "
                         f"'quantile', 'max', 'first', 'last', 'middle', 'random', 'none']")

    # R:.../wav2vec2_fc.py:284-319
    def _determine_stat_pool_embedding_size(self):
        if self.cfg.explicit_stat_pool_embedding_size is not None:
            return self.cfg.explicit_stat_pool_embedding_size
        t = self.cfg.stat_pooling_type
        base = self.wav2vec.num_features
        if t.lower() in ["mean", "first", "first+cls", "last", "middle", "random", "max", "none"]:
            return base
        if t in ("mean+std", "attentive"):
            return base * 2
        if t == "quantile":
            return base * 5
        raise ValueError(f"unknown value for {t=}")

    # R:.../wav2vec2_fc.py:363-397
    def _fc_head_ops_pre_spk_embedding(self, wav2vec_embedding: torch.Tensor):
        pool = self.stat_pooling if self.training else self.test_stat_pooling
        pooled = pool(wav2vec_embedding)
        if not isinstance(self.stat_pooling, NoPooling):
            assert pooled.shape[1] == self.stat_pool_dimension
            assert len(pooled.shape) == 2
            assert pooled.shape[0] == wav2vec_embedding.shape[0]
        x = torch.squeeze(self.embedding_masker(pooled[:, :, None]))
        if self.cfg.embedding_layer_idx < 0:
            return x
        for idx, fc_layer in enumerate(self.fc_list):
            x = fc_layer(x)
            if self.cfg.embedding_layer_idx == idx:
                break
        return x

    # R:.../wav2vec2_fc.py:399-412
    def _fc_head_ops_post_spk_embedding(self, embedding_tensor: torch.Tensor):
        x = embedding_tensor
        if x.dim() == 1:
            x = x[None, :]
        for idx, fc_layer in enumerate(self.fc_list):
            if idx <= self.cfg.embedding_layer_idx:
                continue
            x = fc_layer(x)
        return x

    # R:.../wav2vec2_fc.py:414-431
    def compute_speaker_embedding(self, input_tensor: torch.Tensor) -> torch.Tensor:
        if len(input_tensor.shape) == 3 and input_tensor.shape[1] == 1:
            input_tensor = torch.squeeze(input_tensor)
        if len(input_tensor.shape) == 1:
            input_tensor = torch.stack([input_tensor])
        wav2vec_embeddings = self.wav2vec(input_tensor)                       # [BS, C, T]
        wav2vec_embeddings = torch.transpose(wav2vec_embeddings, 2, 1)        # [BS, T, C]
        return self._fc_head_ops_pre_spk_embedding(wav2vec_embeddings)

    def compute_speaker_embeddings_ragged(self, utterances, max_batch: int = 32, max_pad_fraction: float = 0.15):
        """Speaker embeddings of full-length utterances of DIFFERENT lengths (evaluation; this package's extension of
        the reference's one-utterance-per-step test loop, R:src/lightning_modules/speaker/speaker_recognition_module.py:462-500).
        utterances: sequence of 1-D waveforms (already normalised, any device).  They are sorted into length buckets
        (ragged.plan_buckets), zero-padded per bucket and run with per-utterance length masks, so each row equals
        `compute_speaker_embedding(u[None])` of that utterance alone.  -> [len(utterances), E] in the input order."""
        from .ragged import plan_buckets
        if self.training or torch.is_grad_enabled():
            raise RuntimeError("compute_speaker_embeddings_ragged is an evaluation call: use .eval() under torch.no_grad()")
        pool = self.test_stat_pooling
        if isinstance(pool, (IndexPool1D, NoPooling, QuantilePool1D)) or isinstance(self.wav2vec, Wav2vecLiteWrapperModule):
            raise NotImplementedError("ragged batches are built for mean / mean+std / max / attentive pooling")
        dev = next(self.parameters()).device
        lengths = [int(u.numel()) for u in utterances]
        eng = self.wav2vec.model._engine()
        out = None
        for bucket in plan_buckets(lengths, max_batch, max_pad_fraction):
            n_max = max(lengths[i] for i in bucket)
            wav = torch.zeros(len(bucket), n_max, dtype=torch.float32, device=dev)
            for r, i in enumerate(bucket):
                wav[r, :lengths[i]] = utterances[i].reshape(-1).to(dev, torch.float32)
            lens = [lengths[i] for i in bucket]
            hidden = self.wav2vec(wav, lengths=lens)                              # [B, C, T]
            frames = torch.tensor(eng.frame_lengths(lens), dtype=torch.int32, device=dev)
            pooled = pool(torch.transpose(hidden, 2, 1), lengths=frames)
            x = torch.squeeze(self.embedding_masker(pooled[:, :, None]), 2)
            if self.cfg.embedding_layer_idx >= 0:
                for idx, fc_layer in enumerate(self.fc_list):
                    x = fc_layer(x)
                    if self.cfg.embedding_layer_idx == idx:
                        break
            if out is None:
                out = torch.empty(len(utterances), x.shape[1], dtype=x.dtype, device=dev)
            out[torch.tensor(bucket, device=dev)] = x
        return out

    # R:.../wav2vec2_fc.py:433-438
    def compute_speaker_prediction(self, embedding_tensor: torch.Tensor) -> torch.Tensor:
        return self._fc_head_ops_post_spk_embedding(embedding_tensor).squeeze()

    # R:.../wav2vec2_fc.py:440-463 (test-time option `use_transformers_as_ensembles`): one pooled embedding per encoder
    # output, the last `num_ensembles` of the 13 hidden states (the engine's evaluation forward returns them all)
    def compute_ensemble_embedding(self, input_tensor: torch.Tensor):
        if len(input_tensor.shape) == 3 and input_tensor.shape[1] == 1:
            input_tensor = torch.squeeze(input_tensor)
        if len(input_tensor.shape) == 1:
            input_tensor = torch.stack([input_tensor])
        out = self.wav2vec.model(input_tensor, output_hidden_states=True)
        n = self.wav2vec.model.arch.layers + 1
        embeddings = []
        for hidden in out.hidden_states[n - self.cfg.num_ensembles:n]:
            embeddings.append(torch.squeeze(self.stat_pooling(hidden)))
        return embeddings

    # R:src/lightning_modules/speaker/speaker_recognition_module.py:121-130
    def forward(self, input_tensor: torch.Tensor):
        embedding = self.compute_speaker_embedding(input_tensor)
        prediction = self.compute_speaker_prediction(embedding)
        return embedding, prediction
