"""Binary cross-entropy head loss of the paired-input model (R:src/optim/loss/binary_cross_entropy.py:16-38).

One logit per utterance pair: nothing here is worth a kernel (B numbers); the arithmetic stays in torch and its
gradient enters the hand-written encoder backward through the ordinary autograd graph (the Functions in
``training.py`` take and return UNSCALED fp32 gradients, so any torch head composes with them)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class BinaryCrossEntropyLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.softmax = nn.LogSoftmax(dim=1)          # attribute kept for state-dict / repr parity; unused upstream too

    def forward(self, logits: torch.Tensor, label_indexes: torch.Tensor):
        return self._bce_loss(logits, label_indexes)

    def _bce_loss(self, logits: torch.Tensor, label_indexes: torch.Tensor):
        # logits [BATCH_SIZE, 1] (pre-sigmoid), labels [BATCH_SIZE] in {0, 1}
        scores = logits.squeeze().to(torch.float32)
        target = label_indexes.squeeze().to(torch.float32)
        loss = F.binary_cross_entropy_with_logits(scores, target)
        with torch.no_grad():
            prediction = torch.sigmoid(scores).detach()      # in [0, 1], what the accuracy metric consumes
        return loss, prediction
