"""Drop-in mirror of the reference's ``src/models/wav2vec2.py`` on the sm_100a kernels.

Same class names, constructor arguments, attributes and tensor contracts as
R:src/models/wav2vec2.py:83-169; the arithmetic of the wrapped HuggingFace ``Wav2Vec2Model``
(R:src/models/wav2vec2.py:37-53, HF:1327-1383) is done by ``engine.EncoderEngine`` through the
C ABI.  Parameters are ``nn.Parameter``s under HF's state_dict names, so ``parameters()``,
``state_dict()``, ``load_state_dict(strict=False)`` and ``requires_grad_`` keep working
(SURVEY 8b).

With gradients enabled the forward runs through ``training.EncoderFn`` (hand-written backward, frozen or
unfrozen CNN, train-mode regularisation); there is no PyTorch fallback to fall into.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from ..engine import ArchConfig, EncoderEngine, PreparedWeights, arch_from_id

try:  # the reference derives from LightningModule; it is absent from this image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover - depends on the environment
    class _Base(nn.Module):
        """nn.Module with LightningModule's freeze / unfreeze semantics."""

        def freeze(self) -> None:
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        def unfreeze(self) -> None:
            for p in self.parameters():
                p.requires_grad = True
            self.train()


@dataclass
class Wav2Vec2RegularisationConfig:
    """R:src/models/wav2vec2.py:83-94 (same fields and defaults)."""
    gradient_checkpointing: bool = False
    activation_dropout: float = 0.0
    attention_dropout: float = 0.1
    feat_proj_dropout: float = 0.1
    hidden_dropout: float = 0.1
    layerdrop: float = 0.05
    mask_feature_length: int = 10
    mask_feature_prob: float = 0.0
    mask_time_length: int = 10
    mask_time_prob: float = 0.05


class _Holder(nn.Module):
    """Anonymous container used to reproduce HF's module tree (and therefore its state_dict keys)."""


def _set_param(root: nn.Module, dotted: str, value: torch.Tensor) -> None:
    parts = dotted.split(".")
    m = root
    for name in parts[:-1]:
        if not hasattr(m, name):
            m.add_module(name, _Holder())
        m = getattr(m, name)
    m.register_parameter(parts[-1], nn.Parameter(value))


@dataclass
class Wav2Vec2BaseModelOutput:
    last_hidden_state: torch.Tensor
    extract_features: Optional[torch.Tensor] = None
    hidden_states: Optional[tuple] = None


class _EncoderOutput:
    def __init__(self, last_hidden_state, hidden_states=None):
        self.last_hidden_state = last_hidden_state
        self.hidden_states = hidden_states


class _FeatureExtractor(_Holder):
    """``model.feature_extractor``: [B,N] -> [B,512,T] (HF:409-419)."""

    def forward(self, input_values: torch.Tensor) -> torch.Tensor:
        model = self._owner()
        model._check_mode()
        cnn = model._items()[3]
        if torch.is_grad_enabled() and any(q.requires_grad for q in cnn):
            from ..training import FeatureExtractorFn
            names = [n for n in model._items()[0] if n.startswith("feature_extractor.")]
            return FeatureExtractorFn.apply(input_values.float(), model, names, *cnn).transpose(1, 2)
        return model._engine().feature_extractor(input_values).transpose(1, 2)


class _FeatureProjection(_Holder):
    """``model.feature_projection``: [B,T,512] -> (hidden [B,T,H], normed input) (HF:429-434)."""

    def forward(self, hidden_states: torch.Tensor):
        model = self._owner()
        model._check_mode()
        names, params = model._split_params("feature_projection.")
        if torch.is_grad_enabled() and (hidden_states.requires_grad or any(q.requires_grad for q in params)):
            from ..training import FeatureProjectionFn
            return FeatureProjectionFn.apply(hidden_states, model, names, *params), None
        if model.training and model._stochastic():          # train mode, nothing to differentiate: dropout still applies
            from ..training import projection_forward_train
            B, T, C = hidden_states.shape
            with torch.no_grad():
                h0, _ = projection_forward_train(model._engine(), hidden_states.detach().float().contiguous().view(B * T, C),
                                                 model._draw_split_plan(B, T, hidden_states.device))
            return h0.view(B, T, -1), None
        return model._engine().feature_projection(hidden_states.contiguous()), None


class _Encoder(_Holder):
    """``model.encoder``: [B,T',H] -> object with ``last_hidden_state`` (HF:668-727)."""

    def forward(self, hidden_states: torch.Tensor, output_hidden_states: bool = False, **_):
        model = self._owner()
        model._check_mode()
        names, params = model._split_params("encoder.")
        if torch.is_grad_enabled() and (hidden_states.requires_grad or any(q.requires_grad for q in params)):
            if output_hidden_states:
                raise NotImplementedError("output_hidden_states is only available without gradients")
            from ..training import EncoderStackFn
            return _EncoderOutput(EncoderStackFn.apply(hidden_states, model, names, *params), None)
        if model.training and model._stochastic() and not output_hidden_states:
            from ..training import stack_forward_train
            B, T, H = hidden_states.shape
            with torch.no_grad():
                plan = model._draw_split_plan(B, T, hidden_states.device)
                out = stack_forward_train(model._engine(), hidden_states.detach().float().contiguous().view(B * T, H), B, T,
                                          plan, {"plan": plan})
            return _EncoderOutput(out, None)
        hs = [] if output_hidden_states else None
        out = model._engine().encoder(hidden_states.float(), hs)
        return _EncoderOutput(out, tuple(hs) if hs is not None else None)


class Wav2Vec2ModelB200(nn.Module):
    """Stand-in for ``transformers.Wav2Vec2Model`` (the object the reference keeps in ``.model``)."""

    def __init__(self, arch: ArchConfig, reg_cfg: Optional[Wav2Vec2RegularisationConfig] = None):
        super().__init__()
        self.arch = arch
        self.reg_cfg = reg_cfg or Wav2Vec2RegularisationConfig()
        self.feature_extractor = _FeatureExtractor()
        self.feature_projection = _FeatureProjection()
        self.encoder = _Encoder()
        for sub in (self.feature_extractor, self.feature_projection, self.encoder):
            object.__setattr__(sub, "_owner", lambda self=self: self)
        for name, value in init_hf_parameters(arch).items():
            _set_param(self, name, value)
        self._prepared = None
        self._prepared_sig = None
        from ..checkpoint import install_key_translation
        install_key_translation(self)          # transformers 4.x checkpoints (weight_g / weight_v) load as they are

    # ---- weights in kernel form, rebuilt when any parameter changed -------------------------
    def _items(self):
        """(names, parameters, name -> parameter) cached: walking the module tree costs ~1 ms per call and
        the training step needs it several times.  The Parameter objects never change after construction
        (``.to()`` / optimizers mutate ``.data``)."""
        c = self.__dict__.get("_items_cache")
        if c is None:
            named = list(self.named_parameters())
            c = ([n for n, _ in named], [q for _, q in named], dict(named),
                 [q for n, q in named if n.startswith("feature_extractor.")])
            self.__dict__["_items_cache"] = c
        return c

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self._items()[1])

    def refresh(self) -> None:
        """Call after the parameters were updated in place behind autograd's back (fused optimizer):
        re-derives the fp16 / transposed operand copies with one batched launch."""
        eng = self._prepared
        if eng is not None and self._prepared_sig == self._signature() and eng.w.update():
            tw = getattr(eng, "_train_weights", None)
            if tw is not None:              # same for the data-gradient form of the folded positional-conv weight
                from .. import ops
                for u in list(tw._pos_dgrad.keys()):
                    tw._pos_dgrad[u] = ops.posconv_fold_weight(eng.w._pos_v, eng.w._pos_g, eng.w._groups, u, mode=1)
            return
        self._prepared = None

    def _engine(self) -> EncoderEngine:
        if any(not p.is_cuda for p in self._items()[1]):
            raise RuntimeError("Wav2Vec2ModelB200 runs on a CUDA device only: move the module with .cuda() "
                               "(there is no CPU path)")
        sig = self._signature()
        if self._prepared is None or sig != self._prepared_sig:
            self._prepared = EncoderEngine(PreparedWeights(dict(self.named_parameters()), self.arch))
            self._prepared_sig = sig
        return self._prepared

    def _train_weights(self, eng: EncoderEngine):
        """Transposed fp16 weight copies for the data-gradient GEMMs, cached per prepared weight set."""
        from ..training import TrainWeights
        if getattr(eng, "_train_weights", None) is None:
            eng._train_weights = TrainWeights(eng.w, dict(self.named_parameters()))
        return eng._train_weights

    def _needs_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self._items()[1])

    def _stochastic(self) -> bool:
        r = self.reg_cfg
        return (r.activation_dropout + r.attention_dropout + r.feat_proj_dropout + r.hidden_dropout +
                r.layerdrop + r.mask_time_prob + r.mask_feature_prob) > 0

    def _draw_reg_plan(self, wav: torch.Tensor, eng: EncoderEngine):
        """Host-side draw of this step's regularisation (None in eval mode or when every probability is 0)."""
        if not self.training or not self._stochastic():
            return None
        import numpy as np
        from ..training import RegPlan
        if getattr(self, "_rng", None) is None:
            self._rng = np.random.default_rng(torch.initial_seed() % (1 << 63))
        T = self.arch.conv_lengths(wav.shape[1])[-1]
        if self.reg_cfg.mask_time_prob <= 0:
            return RegPlan(self.reg_cfg, self.arch.layers, wav.shape[0], T, self._rng, wav.device, hidden=self.arch.hidden)
        # ring of pinned staging buffers for the SpecAugment mask (the host may run a few steps ahead of the device)
        ring = self.__dict__.setdefault("_mask_ring", {})
        key = (wav.shape[0], T)
        if key not in ring:
            pin = torch.cuda.is_available()          # (the kernel-stubbed dry run of the CPU test tier has no driver)
            ring[key] = [[torch.empty(wav.shape[0] * T, dtype=torch.uint8, pin_memory=pin) for _ in range(8)], 0]
        bufs, i = ring[key]
        ring[key][1] = (i + 1) % len(bufs)
        return RegPlan(self.reg_cfg, self.arch.layers, wav.shape[0], T, self._rng, wav.device, bufs[i],
                       hidden=self.arch.hidden)

    def _split_params(self, prefix: str):
        """(names, parameters) of one part of the model, in registration order (cached)."""
        c = self.__dict__.setdefault("_split_cache", {})
        if prefix not in c:
            names, params = self._items()[:2]
            keep = [i for i, n in enumerate(names) if n.startswith(prefix)]
            c[prefix] = ([names[i] for i in keep], [params[i] for i in keep])
        return c[prefix]

    def _draw_split_plan(self, B: int, T: int, device):
        """Regularisation draw for one call of the split path (model.feature_projection / model.encoder called on
        their own): dropouts and LayerDrop as usual, no SpecAugment -- HF applies the time mask in
        ``Wav2Vec2Model.forward`` only, which that path bypasses."""
        if not self.training or not self._stochastic():
            return None
        import dataclasses
        import numpy as np
        from ..training import RegPlan
        if getattr(self, "_rng", None) is None:
            self._rng = np.random.default_rng(torch.initial_seed() % (1 << 63))
        return RegPlan(dataclasses.replace(self.reg_cfg, mask_time_prob=0.0), self.arch.layers, B, T, self._rng, device)

    def _check_mode(self):
        if self.arch.feat_extract_norm == "layer" and torch.is_grad_enabled() and any(
                q.requires_grad for q in self._items()[3]):
            raise NotImplementedError(
                "the layer-norm feature extractor (-lv60, XLSR checkpoints) trains with its CNN frozen only "
                "(completely_freeze_feature_extractor: true, the reference default): its backward is not built")

    def forward(self, input_values: torch.Tensor, output_hidden_states: bool = False, lengths=None,
                normalize_input: Optional[bool] = None, **_):
        """HF:1327-1383.  input_values f32 [B,N].  With gradients enabled the forward keeps what the
        hand-written backward needs (training.EncoderFn) so that loss.backward() works.
        lengths (extension, evaluation only): host sequence of sample counts of a zero-padded ragged batch
        (w2v2_speaker_b200/ragged.py); rows behind an utterance's last frame are padding."""
        self._check_mode()
        raw = input_values.dtype == torch.int16 or bool(normalize_input)
        if raw:
            # this package's extension of the boundary: raw 16-bit PCM (or an un-normalised float waveform)
            # comes in, the reference's InputNormalizer2D is folded into conv layer 0 on the device.  Evaluation, or
            # training with the feature extractor frozen (its backward would need the normalised waveform).
            if any(q.requires_grad for q in self._items()[3]) and torch.is_grad_enabled():
                raise NotImplementedError("raw / int16 input needs the feature extractor frozen "
                                          "(completely_freeze_feature_extractor: true, the reference default)")
        else:
            input_values = input_values.float()
        if lengths is not None:
            if self._needs_grad() or (self.training and self._stochastic()):
                raise NotImplementedError("ragged batches (lengths=...) are an evaluation feature: call .eval() under no_grad")
            out = self._engine().forward(input_values, None, lengths, normalize=raw)
            return Wav2Vec2BaseModelOutput(last_hidden_state=out)
        if self._needs_grad():
            from ..training import EncoderFn
            names, params = self._items()[:2]
            self._raw_next = raw                      # (read by EncoderFn.forward: fold the input normaliser into conv 0)
            if output_hidden_states and self.arch.stable_layer_norm:
                raise NotImplementedError("output_hidden_states under gradients is not built for the stable-layer-norm "
                                          "variant: call it in eval mode under no_grad")
            self._keep_saved = bool(output_hidden_states)
            out = EncoderFn.apply(input_values, self, names, *params)
            hs = None
            if output_hidden_states:
                # the per-layer outputs the training forward keeps anyway, as DETACHED views: gradients flow through
                # last_hidden_state only (what the reference differentiates; its hidden-state consumer, the test-time
                # ensemble of R:src/lightning_modules/speaker/wav2vec2_fc.py:440-463, runs without gradients)
                S = self.__dict__.pop("_last_saved")
                B, T, H = out.shape
                hs, cur = [S["top"][0].view(B, T, H)], S["top"][0].view(B, T, H)
                for L in S["layers"]:
                    if L is not None:                 # (a LayerDrop-skipped layer passes its input on, HF:701-713)
                        cur = L["arena"].tensor("h2_32", torch.float32, (B, T, H))
                    hs.append(cur)
                hs = tuple(h.detach() for h in hs)
            return Wav2Vec2BaseModelOutput(last_hidden_state=out, hidden_states=hs)
        eng = self._engine()
        if self.training and self._stochastic():
            # train mode without gradients: the reference reaches this routinely (`wav2vec_initially_frozen`: freeze() calls
            # eval(), Lightning calls .train() again after every validation loop while the encoder is still frozen) and HF
            # then runs a dropout-active, gradient-less forward.  Same here: the training forward with this step's
            # regularisation draw, its saved state discarded.
            if output_hidden_states:
                raise NotImplementedError("output_hidden_states is only available in eval mode")
            from ..training import encoder_forward_train
            with torch.no_grad():
                out, _ = encoder_forward_train(eng, input_values, self._draw_reg_plan(input_values, eng),
                                               self.masked_spec_embed.detach(), False,
                                               getattr(self, "_pre_encoder_hook", None), normalize=raw)
            return Wav2Vec2BaseModelOutput(last_hidden_state=out)
        trace = {} if output_hidden_states else None
        out = eng.forward(input_values, trace, None, normalize=raw)
        hs = tuple(trace["hidden_states"]) if trace is not None else None
        return Wav2Vec2BaseModelOutput(last_hidden_state=out, hidden_states=hs)


def init_hf_parameters(arch: ArchConfig) -> dict:
    """Random initialisation following HF ``_init_weights`` (HF:968-1003) and the module defaults."""
    C, H = arch.conv_dim, arch.hidden
    p = {}
    p["masked_spec_embed"] = torch.rand(H)
    cin = 1
    for i, k in enumerate(arch.conv_kernel):
        w = torch.empty(C, cin, k)
        nn.init.kaiming_normal_(w)
        p[f"feature_extractor.conv_layers.{i}.conv.weight"] = w
        if arch.conv_bias:                                   # (-lv60 / XLSR: HF:275-299)
            bound = math.sqrt(1.0 / (cin * k))
            p[f"feature_extractor.conv_layers.{i}.conv.bias"] = torch.empty(C).uniform_(-bound, bound)
        if i == 0 or arch.feat_extract_norm == "layer":
            p[f"feature_extractor.conv_layers.{i}.layer_norm.weight"] = torch.ones(C)
            p[f"feature_extractor.conv_layers.{i}.layer_norm.bias"] = torch.zeros(C)
        cin = C
    p["feature_projection.layer_norm.weight"] = torch.ones(C)
    p["feature_projection.layer_norm.bias"] = torch.zeros(C)
    kk = math.sqrt(1.0 / C)
    p["feature_projection.projection.weight"] = torch.empty(H, C).uniform_(-kk, kk)
    p["feature_projection.projection.bias"] = torch.empty(H).uniform_(-kk, kk)
    gsz = H // arch.pos_groups
    v = torch.randn(H, gsz, arch.pos_kernel) * (2.0 * math.sqrt(1.0 / (arch.pos_kernel * H)))
    p["encoder.pos_conv_embed.conv.bias"] = torch.zeros(H)
    p["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    p["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
    p["encoder.layer_norm.weight"] = torch.ones(H)
    p["encoder.layer_norm.bias"] = torch.zeros(H)
    for l in range(arch.layers):
        pre = f"encoder.layers.{l}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            p[pre + f"attention.{nm}.weight"] = torch.randn(H, H) * 0.02
            p[pre + f"attention.{nm}.bias"] = torch.zeros(H)
        p[pre + "layer_norm.weight"] = torch.ones(H)
        p[pre + "layer_norm.bias"] = torch.zeros(H)
        p[pre + "feed_forward.intermediate_dense.weight"] = torch.randn(arch.ffn, H) * 0.02
        p[pre + "feed_forward.intermediate_dense.bias"] = torch.zeros(arch.ffn)
        p[pre + "feed_forward.output_dense.weight"] = torch.randn(H, arch.ffn) * 0.02
        p[pre + "feed_forward.output_dense.bias"] = torch.zeros(H)
        p[pre + "final_layer_norm.weight"] = torch.ones(H)
        p[pre + "final_layer_norm.bias"] = torch.zeros(H)
    return p


def load_base_wav2vec2_model(huggingface_id: str, reg_cfg: Optional[Wav2Vec2RegularisationConfig] = None,
                             device=torch.device("cpu")) -> Wav2Vec2ModelB200:
    """R:src/models/wav2vec2.py:25-55.  No network / HF cache exists here, so the architecture named
    by ``huggingface_id`` is random-initialised; load weights with ``load_state_dict`` (HF key names)."""
    return Wav2Vec2ModelB200(arch_from_id(huggingface_id), reg_cfg).to(device)


def wav2vec2_embed_raw_audio(input_tensor: torch.Tensor, model: Wav2Vec2ModelB200, lengths=None) -> torch.Tensor:
    """R:src/models/wav2vec2.py:62-76: [B, num_samples] -> [B, H, num_frames]."""
    output = model(input_tensor, lengths=lengths) if lengths is not None else model(input_tensor)
    return output.last_hidden_state.transpose(1, 2)


def reset_model(model: nn.Module) -> None:
    """R:src/util.py:214-226 re-initialises every leaf with its PyTorch default ``reset_parameters``.
    The parameter holders here have no torch leaf modules, so the equivalent is a fresh HF-style init."""
    fresh = init_hf_parameters(model.arch)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            prm.copy_(fresh[name].to(prm.device))


class Wav2Vec2WrapperModule(_Base):
    """R:src/models/wav2vec2.py:97-146."""

    def __init__(self, wav2vec2_huggingface_id: str, reset_weights: bool,
                 reg_cfg: Optional[Wav2Vec2RegularisationConfig] = None, insert_clc_token: bool = False,
                 cls_token_constant: float = 1):
        super().__init__()
        self.model = load_base_wav2vec2_model(wav2vec2_huggingface_id, reg_cfg)
        self.insert_cls_token = insert_clc_token
        self.cls_token_constant = cls_token_constant
        if "base" in wav2vec2_huggingface_id:
            self.num_features = 768
        elif "large" in wav2vec2_huggingface_id:
            self.num_features = 1024
        else:
            raise ValueError("cannot determine num features")
        if reset_weights:
            reset_model(self.model)

    @property
    def num_embedding_features(self):
        return self.num_features

    def forward(self, wav_input: torch.Tensor, lengths=None):
        # wav_input has shape [BATCH_SIZE, NUM_SAMPLES]
        if lengths is not None:
            if self.insert_cls_token:
                raise NotImplementedError("ragged batches are not built for the CLS-token path")
            return wav2vec2_embed_raw_audio(wav_input, self.model, lengths)
        if self.insert_cls_token:
            # R:src/models/wav2vec2.py:128-140 (the reference hard-codes the CLS width to 768)
            features = self.model.feature_extractor(wav_input).transpose(1, 2)
            features, _ = self.model.feature_projection(features)
            cls_token = torch.ones((wav_input.shape[0], 1, 768), device=wav_input.device) * self.cls_token_constant
            sequence = torch.cat([cls_token, features], dim=1)
            embedding = self.model.encoder(sequence).last_hidden_state.transpose(1, 2)
        else:
            embedding = wav2vec2_embed_raw_audio(wav_input, self.model)
        # [BATCH_SIZE, NUM_FEATURES, NUM_FRAMES]
        return embedding


class Wav2vecLiteWrapperModule(_Base):
    """R:src/models/wav2vec2.py:149-169: CNN feature extractor only."""
    num_features = 512

    def __init__(self, wav2vec2_huggingface_id: str, reset_weights: bool):
        super().__init__()
        self.model = load_base_wav2vec2_model(wav2vec2_huggingface_id)
        if reset_weights:
            reset_model(self.model)

    @property
    def num_embedding_features(self):
        return self.num_features

    def forward(self, wav_input: torch.Tensor):
        return self.model.feature_extractor(wav_input)
