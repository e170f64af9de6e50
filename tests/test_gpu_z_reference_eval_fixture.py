"""GPU evaluator against outputs of the reference's OWN CosineDistanceEvaluator (tests/golden/ref_eval.npz, produced by
oracle/make_golden.py eval).  Sorted last on purpose: it was added after the round's GPU budget was spent (its CPU
counterpart, tests/test_host_logic.py::test_eval_metrics_match_the_reference_functions, runs everywhere)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("center", [False, True])
def test_cosine_evaluator_matches_the_reference_evaluator(center):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from w2v2_speaker_b200.evaluation.speaker import CosineDistanceEvaluator, EmbeddingSample, EvaluationPair
    g = golden("ref_eval.npz")
    emb = torch.from_numpy(g["embeddings"])
    ids = [f"utt{i}" for i in range(emb.shape[0])]
    samples = [EmbeddingSample(i, e) for i, e in zip(ids, emb)]
    pairs = [EvaluationPair(bool(s), ids[a], ids[b]) for s, a, b in zip(g["same"], g["left"], g["right"])]
    ev = CosineDistanceEvaluator(center_before_scoring=center, length_norm_before_scoring=True, max_num_training_samples=0)
    ev.fit_parameters(list(emb[::int(g["fit_stride"])]), [])
    res = ev.evaluate(pairs, samples)
    key = "center" if center else "plain"
    # fp32 scores from a different summation order can swap near-tied trials: one trial of 2000 is 5e-4
    assert abs(res["eer"] - float(g[key + ".eer"])) < 1e-3
    assert abs(res["eer_threshold"] - float(g[key + ".eer_threshold"])) < 1e-3
    assert abs(res["mdc"] - float(g[key + ".mdc"])) < 2e-3
    assert abs(res["mdc_threshold"] - float(g[key + ".mdc_threshold"])) < 2e-3
