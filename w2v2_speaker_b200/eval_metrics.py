"""Equal error rate and minimum detection cost (mirror of R:src/eval_metrics.py:54-79, 91-206).

The reference computes the EER with sklearn's ``roc_curve`` + scipy ``interp1d`` / ``brentq`` and the min-DCF
with python loops over sorted score lists; here both are numpy restatements of the same definitions (a sort, two
cumulative sums, one linear interpolation), checked against the reference formulation in
tests/test_host_logic.py (oracle/eval_oracle.py)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def _verify_correct_scores(groundtruth_scores, predicted_scores):
    """R:src/eval_metrics.py:21-47."""
    if len(groundtruth_scores) != len(predicted_scores):
        raise ValueError(f"length of input lists should match, while groundtruth_scores={len(groundtruth_scores)} and "
                         f"predicted_scores={len(predicted_scores)}")
    if not all(np.isin(groundtruth_scores, [0, 1])):
        raise ValueError(f"groundtruth values should be either 0 and 1, while they are actually one of "
                         f"{np.unique(groundtruth_scores)}")


def _roc(gt: np.ndarray, sc: np.ndarray, pos_label: int):
    """(fpr, tpr, thresholds) like sklearn.metrics.roc_curve: one point per distinct score, interior points of
    straight runs dropped (that keeps the curve but decides between which thresholds the reference interpolates),
    preceded by the (0, 0) point with threshold +inf."""
    y = (gt == pos_label).astype(np.float64)
    order = np.argsort(-sc, kind="mergesort")
    sc, y = sc[order], y[order]
    last = np.r_[np.where(np.diff(sc))[0], sc.size - 1]          # last index of every run of equal scores
    tps = np.cumsum(y)[last]
    fps = (1 + last) - tps
    thr = sc[last]
    P, N = y.sum(), y.size - y.sum()
    if P == 0 or N == 0:
        raise ValueError("EER needs both positive and negative trials")
    if fps.size > 2:
        keep = np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True]
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    return np.r_[0.0, fps / N], np.r_[0.0, tps / P], np.r_[np.inf, thr]


def calculate_eer(groundtruth_scores: List[int], predicted_scores: List[float], pos_label: int = 1) -> Tuple[float, float]:
    """Point where the ROC curve (linear between operating points) crosses fpr = 1 - tpr, and the (linearly
    interpolated) threshold there -- R:src/eval_metrics.py:54-79."""
    _verify_correct_scores(groundtruth_scores, predicted_scores)
    if pos_label not in (0, 1):
        raise ValueError(f"The positive label should be either 0 or 1, not {pos_label}")
    gt = np.asarray(groundtruth_scores)
    sc = np.asarray(predicted_scores, dtype=np.float64)
    if not np.all(np.isfinite(sc)):
        raise ValueError("scores contain NaN or Inf")
    fpr, tpr, thr = _roc(gt, sc, pos_label)
    f = 1.0 - fpr - tpr                               # non-increasing along the curve, 1 at the start, -1 at the end
    i = int(np.searchsorted(-f, 0.0, side="left"))    # first point with f <= 0
    if f[i] == 0.0:
        return float(fpr[i]), float(thr[i])
    x0, y0, x1, y1 = fpr[i - 1], tpr[i - 1], fpr[i], tpr[i]
    t = (1.0 - x0 - y0) / ((x1 - x0) + (y1 - y0))
    eer = x0 + t * (x1 - x0)
    if x1 == x0:
        # vertical piece: the curve is not a function of fpr there; the reference's root finder (brentq on
        # 1 - x - interp1d(fpr, tpr)(x)) closes in on the jump from below, where the threshold is that of its
        # lower end
        return float(eer), float(thr[i - 1] if np.isfinite(thr[i - 1]) else thr[i])
    w = (eer - x0) / (x1 - x0)
    th = thr[i] if not np.isfinite(thr[i - 1]) else thr[i - 1] + w * (thr[i] - thr[i - 1])
    return float(eer), float(th)


def calculate_mdc(groundtruth_scores: List[int], predicted_scores: List[float], c_miss: float = 1, c_fa: float = 1,
                  p_target: float = 0.05) -> Tuple[float, float]:
    """Minimum of the normalised detection cost over all score thresholds -- R:src/eval_metrics.py:91-206
    (NIST SRE 2016 evaluation plan, section 3)."""
    _verify_correct_scores(groundtruth_scores, predicted_scores)
    if c_miss < 1:
        raise ValueError(f"c_miss={c_miss} should be >= 1")
    if c_fa < 1:
        raise ValueError(f"c_fa={c_fa} should be >= 1")
    if p_target < 0 or p_target > 1:
        raise ValueError(f"p_target={p_target} should be between 0 and 1")
    gt = np.asarray(groundtruth_scores, dtype=np.float64)
    sc = np.asarray(predicted_scores, dtype=np.float64)
    order = np.argsort(sc, kind="mergesort")           # ascending, stable like python's sorted()
    gt, thr = gt[order], sc[order]
    if gt.sum() == 0 or gt.sum() == gt.size:
        # the reference divides python ints here (R:src/eval_metrics.py:130-139) and lets the ZeroDivisionError reach its
        # evaluator, which maps it to mdc = 1, threshold = 1337 (R:src/evaluation/speaker/speaker_recognition_evaluator.py)
        raise ZeroDivisionError("calculate_mdc needs trials of both classes (all ground-truth labels are equal)")
    fnrs = np.cumsum(gt) / gt.sum()
    fprs = 1.0 - np.cumsum(1.0 - gt) / (gt.size - gt.sum())
    c_det = c_miss * fnrs * p_target + c_fa * fprs * (1 - p_target)
    i = int(np.argmin(c_det))                          # first minimum, like the reference's strict `<`
    return float(c_det[i] / min(c_miss * p_target, c_fa * (1 - p_target))), float(thr[i])
