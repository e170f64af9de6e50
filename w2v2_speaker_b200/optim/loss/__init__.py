"""Mirror of the reference's ``src/optim/loss`` exports for the hot path and its sibling heads
(R:src/optim/loss/__init__.py:1-4, R:src/optim/loss/binary_cross_entropy.py)."""
from .aam_softmax import AngularAdditiveMarginSoftMaxLoss
from .binary_cross_entropy import BinaryCrossEntropyLoss
from .cross_entropy import CrossEntropyLoss

__all__ = ["AngularAdditiveMarginSoftMaxLoss", "BinaryCrossEntropyLoss", "CrossEntropyLoss"]
