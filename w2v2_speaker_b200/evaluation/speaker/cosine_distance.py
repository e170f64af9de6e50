"""Drop-in mirror of the reference's trial evaluator (R:src/evaluation/speaker/cosine_distance.py:64-132 and its
base class R:src/evaluation/speaker/speaker_recognition_evaluator.py:25-152): same data classes, constructor
arguments, ``fit_parameters`` / ``reset_parameters`` / ``evaluate`` contract and result dictionary.

The scoring itself runs on the GPU: the pooled embeddings of all samples form one table in HBM and every trial is a
(row, row) pair scored by ``w2v2_cosine_pairs`` -- the reference stacks two [num_pairs, E] copies of the embeddings
on the host and calls torch's CosineSimilarity on them.  Ensembles of embeddings and non-pooled embeddings (the
reference's two "dirty hack" branches) are not on the measured path and raise."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union
from warnings import warn

import numpy as np
import torch

from ... import ops
from ...eval_metrics import calculate_eer, calculate_mdc


@dataclass
class EvaluationPair:
    same_speaker: bool
    sample1_id: str
    sample2_id: str


@dataclass
class EmbeddingSample:
    sample_id: str
    embedding: Union[torch.Tensor, List[torch.Tensor]]


class CosineDistanceEvaluator:
    def __init__(self, center_before_scoring: bool, length_norm_before_scoring: bool, max_num_training_samples: int,
                 device: Optional[torch.device] = None):
        self.max_num_training_samples = max_num_training_samples
        self.center_before_scoring = center_before_scoring
        self.length_norm_before_scoring = length_norm_before_scoring     # a cosine is invariant to it
        self.device = device
        self.mean = None
        self.std = None

    # R:src/evaluation/speaker/cosine_distance.py:83-104
    def fit_parameters(self, embedding_tensors: List[torch.Tensor], _label_tensors: List[torch.Tensor]):
        if not self.center_before_scoring:
            return
        if len(embedding_tensors) <= 2:
            raise ValueError("mean/std calculation requires more than 2 samples")
        self.std, self.mean = torch.std_mean(torch.stack(embedding_tensors, dim=0).float(), dim=0)

    def reset_parameters(self):
        if self.center_before_scoring:
            self.mean = None
            self.std = None

    def _device(self, samples) -> torch.device:
        if self.device is not None:
            return torch.device(self.device)
        for s in samples:
            if torch.is_tensor(s.embedding) and s.embedding.is_cuda:
                return s.embedding.device
        return torch.device("cuda", torch.cuda.current_device())

    def score_pairs(self, pairs: List[EvaluationPair], samples: List[EmbeddingSample]) -> Optional[np.ndarray]:
        """Raw cosine scores in [-1, 1], one per trial, or None if a trial names an unknown sample."""
        index: Dict[str, int] = {}
        for i, s in enumerate(samples):
            if s.sample_id in index:
                raise ValueError(f"duplicate key {s.sample_id}")
            if isinstance(s.embedding, list) or s.embedding.dim() != 1:
                raise NotImplementedError("ensembles / non-pooled embeddings are not scored by the sm_100a path")
            index[s.sample_id] = i
        ia, ib = [], []
        for p in pairs:
            if p.sample1_id not in index or p.sample2_id not in index:
                warn(f"{p.sample1_id} or {p.sample2_id} not in sample_map")
                return None
            ia.append(index[p.sample1_id])
            ib.append(index[p.sample2_id])
        dev = self._device(samples)
        table = torch.stack([s.embedding.detach().float() for s in samples]).to(dev)
        mean = std = None
        if self.center_before_scoring:
            if self.mean is None:
                raise ValueError("fit_parameters has to be called before scoring with centering")
            mean, std = self.mean.to(dev).float().contiguous(), self.std.to(dev).float().contiguous()
        sc = ops.cosine_pairs(table, torch.tensor(ia, dtype=torch.int32, device=dev),
                              torch.tensor(ib, dtype=torch.int32, device=dev), mean, std)
        return sc.cpu().numpy().astype(np.float64)

    # R:src/evaluation/speaker/speaker_recognition_evaluator.py:46-115
    def evaluate(self, pairs: List[EvaluationPair], samples: List[EmbeddingSample]):
        scores = self.score_pairs(pairs, samples)
        if scores is None:
            return {"eer": -1, "eer_threshold": -1, "mdc": -1, "mdc_threshold": -1}
        ground_truth_scores = [1 if p.same_speaker else 0 for p in pairs]
        prediction_scores = np.clip((scores + 1) / 2, 0, 1).tolist()        # to [0, 1] like the reference
        try:
            eer, eer_threshold = calculate_eer(ground_truth_scores, prediction_scores, pos_label=1)
        except (ValueError, ZeroDivisionError) as e:
            print(f"EER calculation had {e}")
            eer, eer_threshold = 1, 1337
        try:
            mdc, mdc_threshold = calculate_mdc(ground_truth_scores, prediction_scores)
        except (ValueError, ZeroDivisionError) as e:
            print(f"mdc calculation had {e}")
            mdc, mdc_threshold = 1, 1337
        return {"eer": eer, "eer_threshold": eer_threshold, "mdc": mdc, "mdc_threshold": mdc_threshold}
