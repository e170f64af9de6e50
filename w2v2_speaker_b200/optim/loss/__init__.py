"""Mirror of the reference's ``src/optim/loss`` exports for the hot path (R:src/optim/loss/__init__.py:1-4)."""
from .aam_softmax import AngularAdditiveMarginSoftMaxLoss
from .cross_entropy import CrossEntropyLoss

__all__ = ["AngularAdditiveMarginSoftMaxLoss", "CrossEntropyLoss"]
