"""TEST INFRASTRUCTURE (see oracle/__init__.py): the reference's own formulation of the evaluation metrics, restated
with the very libraries it uses -- sklearn ``roc_curve`` + scipy ``interp1d`` / ``brentq`` for the EER
(R:src/eval_metrics.py:54-79) and the list-based Kaldi-derived loops for the minimum detection cost
(R:src/eval_metrics.py:91-206) -- plus torch's CosineSimilarity with the evaluator's centring
(R:src/evaluation/speaker/cosine_distance.py:107-132, speaker_recognition_evaluator.py:154-172).
Only tests/ may import this module."""
from operator import itemgetter

import numpy as np
import torch


def eer(groundtruth_scores, predicted_scores, pos_label=1):
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from sklearn.metrics import roc_curve
    fpr, tpr, thresholds = roc_curve(groundtruth_scores, predicted_scores, pos_label=pos_label)
    e = brentq(lambda x: 1.0 - x - interp1d(fpr, tpr)(x), 0.0, 1.0)
    return e, interp1d(fpr, thresholds)(e).item()


def mdc(groundtruth_scores, predicted_scores, c_miss=1, c_fa=1, p_target=0.05):
    sorted_indexes, thresholds = zip(*sorted([(i, t) for i, t in enumerate(predicted_scores)], key=itemgetter(1)))
    gt = [groundtruth_scores[i] for i in sorted_indexes]
    fnrs, fprs = [], []
    for i in range(len(gt)):
        fnrs.append((fnrs[i - 1] if i else 0) + gt[i])
        fprs.append((fprs[i - 1] if i else 0) + 1 - gt[i])
    fn_norm = sum(gt)
    fp_norm = len(gt) - fn_norm
    fnrs = [x / float(fn_norm) for x in fnrs]
    fprs = [1 - x / float(fp_norm) for x in fprs]
    best, best_t = float("inf"), thresholds[0]
    for i in range(len(fnrs)):
        c = c_miss * fnrs[i] * p_target + c_fa * fprs[i] * (1 - p_target)
        if c < best:
            best, best_t = c, thresholds[i]
    return best / min(c_miss * p_target, c_fa * (1 - p_target)), best_t


def cosine_scores(left: torch.Tensor, right: torch.Tensor, mean=None, std=None) -> torch.Tensor:
    if mean is not None:
        left = (left - mean) / (std + 1e-12)
        right = (right - mean) / (std + 1e-12)
    return torch.nn.CosineSimilarity()(left, right)
