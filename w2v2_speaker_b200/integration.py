"""Switch a checkout of the reference (nikvaessen/w2v2-speaker) onto the sm_100a path WITHOUT editing its files.

    import w2v2_speaker_b200.integration as b200
    b200.install()                      # before anything imports src.lightning_modules / src.main
    from src.lightning_modules.speaker.wav2vec2_fc import Wav2vec2FCModule      # the reference's own class

``install()`` registers this package's mirrors under the reference's module names, so that the reference's own
callers -- ``Wav2vec2FCModule`` (R:src/lightning_modules/speaker/wav2vec2_fc.py:24-39 imports),
``Wav2vec2PairedSpeakerModule``, ``src.main`` (hydra ``_target_``s) -- pick them up:

    src.models.wav2vec2            -> w2v2_speaker_b200.models.wav2vec2
    src.layers.pooling             -> w2v2_speaker_b200.layers.pooling
    src.layers.embedding_masking   -> w2v2_speaker_b200.layers.embedding_masking
    src.optim.loss.aam_softmax     -> w2v2_speaker_b200.optim.loss.aam_softmax
    src.optim.loss.cross_entropy   -> w2v2_speaker_b200.optim.loss.cross_entropy
    src.optim.loss.binary_cross_entropy -> w2v2_speaker_b200.optim.loss.binary_cross_entropy
    src.evaluation.speaker.cosine_distance (CosineDistanceEvaluator) and src.eval_metrics stay the reference's unless
    ``evaluation=True`` (they are CPU code outside the hot path; the mirrors are drop-in for them too).

The one thing a module swap cannot reach is the classifier the reference builds itself,
``nn.Linear(stat_pool_dimension, num_speakers)`` (R:.../wav2vec2_fc.py:176-210): left alone it runs on cuBLAS.
``accelerate_heads(module)`` re-types those layers in place to ``SpeakerLinear`` (same Parameter objects, same
state_dict keys); ``install(patch_heads=True)`` (the default) wraps ``Wav2vec2FCModule.__init__`` to do it on
construction.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict

import torch.nn as nn

_MAP: Dict[str, str] = {
    "src.models.wav2vec2": "w2v2_speaker_b200.models.wav2vec2",
    "src.layers.pooling": "w2v2_speaker_b200.layers.pooling",
    "src.layers.embedding_masking": "w2v2_speaker_b200.layers.embedding_masking",
    "src.optim.loss.aam_softmax": "w2v2_speaker_b200.optim.loss.aam_softmax",
    "src.optim.loss.cross_entropy": "w2v2_speaker_b200.optim.loss.cross_entropy",
    "src.optim.loss.binary_cross_entropy": "w2v2_speaker_b200.optim.loss.binary_cross_entropy",
}
_EVAL_MAP: Dict[str, str] = {
    "src.evaluation.speaker.cosine_distance": "w2v2_speaker_b200.evaluation.speaker.cosine_distance",
    "src.eval_metrics": "w2v2_speaker_b200.eval_metrics",
}


def accelerate_heads(module: nn.Module) -> int:
    """Re-type every plain ``nn.Linear`` of ``module.fc_list`` to ``SpeakerLinear`` in place (parameters, hooks and
    state_dict keys untouched; only ``forward`` changes: error-compensated fp16 operands on the tensor cores, fp32
    accurate logits).  Returns the number of layers converted.  Idempotent."""
    from .layers.linear import SpeakerLinear
    n = 0
    for m in getattr(module, "fc_list", nn.ModuleList()).modules():
        if type(m) is nn.Linear:
            m.__class__ = SpeakerLinear
            m._w_split = None
            m._w_sig = None
            n += 1
    return n


def install(patch_heads: bool = True, evaluation: bool = False) -> None:
    """Register the mirrors under the reference's module names (see the module docstring).  Call it before the
    reference's lightning modules are imported; modules that were imported earlier keep the classes they bound."""
    already = [k for k in _MAP if k in sys.modules and sys.modules[k].__name__ == k]
    if already:
        raise RuntimeError("install() must run before the reference imports %s" % ", ".join(already))
    table = dict(_MAP)
    if evaluation:
        table.update(_EVAL_MAP)
    for ref_name, own_name in table.items():
        sys.modules[ref_name] = importlib.import_module(own_name)
    if patch_heads:
        _PatchHeads.arm()


class _PatchHeads:
    """Import hook: the first time ``src.lightning_modules.speaker.wav2vec2_fc`` is imported, wrap
    ``Wav2vec2FCModule.__init__`` so that the classifier layers it builds are converted on construction."""
    TARGET = "src.lightning_modules.speaker.wav2vec2_fc"
    armed = False

    @classmethod
    def arm(cls):
        if cls.TARGET in sys.modules:
            cls.patch(sys.modules[cls.TARGET])
        elif not cls.armed:
            sys.meta_path.insert(0, cls())
            cls.armed = True

    @staticmethod
    def patch(mod):
        klass = getattr(mod, "Wav2vec2FCModule", None)
        if klass is None or getattr(klass, "_b200_heads", False):
            return
        orig = klass.__init__

        def __init__(self, *a, **k):
            orig(self, *a, **k)
            accelerate_heads(self)
        __init__.__wrapped__ = orig
        klass.__init__ = __init__
        klass._b200_heads = True

    # importlib.abc.MetaPathFinder protocol: let the normal machinery import the target, then patch it
    def find_spec(self, fullname, path, target=None):
        if fullname != self.TARGET:
            return None
        sys.meta_path.remove(self)
        type(self).armed = False
        spec = importlib.util.find_spec(fullname)
        if spec is None or spec.loader is None:
            return spec
        loader = spec.loader
        exec_module = loader.exec_module

        def exec_and_patch(module):
            exec_module(module)
            _PatchHeads.patch(module)
        loader.exec_module = exec_and_patch
        return spec


import importlib.util  # noqa: E402  (used by _PatchHeads.find_spec)
