"""B200-native (sm_100a) implementation of the w2v2-speaker hot path.

Host side mirrors the reference's Python interface for the path (``models.wav2vec2``,
``layers.pooling``, ``optim.loss``); the arithmetic runs in hand-written CUDA kernels behind the
C ABI declared in ``include/w2v2_b200.h`` (``libw2v2_b200.so``).  There is no CPU fallback.
"""
__all__ = ["_lib", "engine"]
