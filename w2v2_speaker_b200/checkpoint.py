"""Checkpoint interoperability: weights trained with the reference load into the sm_100a modules.

Two sources exist in the reference's world:

* HuggingFace ``Wav2Vec2Model`` state dicts (``facebook/wav2vec2-base``, R:src/models/wav2vec2.py:25-55).  Under
  transformers 4.x (the reference pins ``^4.8.2``) the weight-normed positional conv stores ``weight_g`` /
  ``weight_v``; transformers 5.x -- and therefore this package, which mirrors the 5.x names -- stores
  ``parametrizations.weight.original0`` / ``original1``.  Task-head checkpoints (``Wav2Vec2ForCTC``,
  ``Wav2Vec2ForPreTraining``) prefix every encoder key with ``wav2vec2.`` and carry extra heads.
* Lightning checkpoints of the reference's own modules (R:src/main.py:279-281 ``load_from_checkpoint(..., strict=False)``):
  ``{"state_dict": {"wav2vec.model.<hf key>": ..., "fc_list.0.0.weight": ..., "loss_fn.fc_weights": ...}}``.
  ``Wav2vec2FCModule`` reproduces that module tree, so only the positional-conv names need translating.

Pure key bookkeeping on the host; no arithmetic."""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Tuple

import torch

_POS = "pos_conv_embed.conv."
_RENAMES = ((_POS + "weight_g", _POS + "parametrizations.weight.original0"),
            (_POS + "weight_v", _POS + "parametrizations.weight.original1"))
# heads of HF task models that have no counterpart in the bare encoder
_HF_HEAD_PREFIXES = ("quantizer.", "project_q.", "project_hid.", "lm_head.", "classifier.", "projector.", "dropout_features.")


def translate_key(key: str) -> str:
    """transformers 4.x name -> the 5.x name this package uses (identity for everything else)."""
    for old, new in _RENAMES:
        if key.endswith(old):
            return key[:-len(old)] + new
    return key


def convert_hf_state_dict(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Any HuggingFace wav2vec2 checkpoint -> keys of ``Wav2Vec2ModelB200.state_dict()``: strips the ``wav2vec2.``
    prefix of task models, drops their heads, translates the 4.x weight-norm names."""
    out = {}
    for k, v in state_dict.items():
        if k.startswith("wav2vec2."):
            k = k[len("wav2vec2."):]
        elif k.startswith(_HF_HEAD_PREFIXES):
            continue
        out[translate_key(k)] = v
    return out


def convert_lightning_state_dict(checkpoint: Mapping) -> Dict[str, torch.Tensor]:
    """A Lightning checkpoint (or its bare ``state_dict``) of the reference's ``Wav2vec2FCModule`` -> keys of this
    package's ``Wav2vec2FCModule``."""
    sd = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    return {translate_key(k): v for k, v in sd.items()}


def load_reference_checkpoint(module: torch.nn.Module, checkpoint: Mapping, strict: bool = False) -> Tuple[Iterable[str], Iterable[str]]:
    """``network_class.load_from_checkpoint(path, strict=False)`` of R:src/main.py:279-281 for an already constructed
    module: returns (missing_keys, unexpected_keys) like ``load_state_dict``."""
    res = module.load_state_dict(convert_lightning_state_dict(checkpoint), strict=strict)
    return res.missing_keys, res.unexpected_keys


def install_key_translation(model: torch.nn.Module) -> None:
    """Make ``model.load_state_dict`` accept the 4.x names directly (a load_state_dict pre-hook that renames in place)."""

    def hook(module, state_dict, prefix, *args):
        for k in list(state_dict.keys()):
            if k.startswith(prefix):
                nk = translate_key(k)
                if nk != k and nk not in state_dict:
                    state_dict[nk] = state_dict.pop(k)

    model.register_load_state_dict_pre_hook(hook)
