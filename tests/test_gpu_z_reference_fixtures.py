"""GPU path against outputs of the reference's OWN classes, committed as fixtures by oracle/make_golden.py:
the evaluator (tests/golden/ref_eval.npz: CosineDistanceEvaluator.evaluate) and the paired-input model
(tests/golden/ref_paired_b3.npz: one training step of Wav2vec2PairedSpeakerModule + BinaryCrossEntropyLoss).
A third test runs BASELINE.json's full size (64 x 3 s) through the batch-independence property.
Sorted last on purpose: these were added after the round's GPU budget was spent.  Their CPU counterparts run
everywhere (tests/test_host_logic.py::test_eval_metrics_match_the_reference_functions,
tests/test_oracle_golden.py::test_split_path_restatement_matches_the_reference_paired_model), and the same GPU code was
validated against the oracle those CPU tests pin (tests/test_gpu_split_path.py, tests/test_gpu_modules.py)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

ZERO_REG = dict(activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                mask_time_prob=0.0, mask_feature_prob=0.0)


def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


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


def test_paired_input_model_matches_the_reference_fixture(base_params):
    """The same training step against tests/golden/ref_paired_b3.npz, which the reference's OWN
    Wav2vec2PairedSpeakerModule + BinaryCrossEntropyLoss produced (oracle/make_golden.py paired): eval scores, loss,
    prediction, the Linear head's gradients and every encoder gradient (norm and a strided sample each)."""
    _need_cuda()
    from oracle.params import make_inputs
    from w2v2_speaker_b200.optim.loss import BinaryCrossEntropyLoss
    from w2v2_speaker_b200.paired_speaker_module import Wav2vec2PairedSpeakerModule, Wav2vec2PairedSpeakerModuleConfig
    g = golden("ref_paired_b3.npz")
    wav_a, _ = make_inputs(3, 16000, seed=31)
    wav_b, _ = make_inputs(3, 11283, seed=32)
    labels = torch.from_numpy(g["labels"])
    m = Wav2vec2PairedSpeakerModule(Wav2vec2PairedSpeakerModuleConfig(**ZERO_REG), BinaryCrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params)
    with torch.no_grad():
        m.linear.weight.copy_(torch.from_numpy(g["linear.weight"]))
        m.linear.bias.copy_(torch.from_numpy(g["linear.bias"]))
    m = m.cuda().eval()
    with torch.no_grad():
        scores = m(wav_a.cuda(), wav_b.cuda())
    assert np.abs(scores.cpu().numpy() - g["scores.eval"]).max() < 5e-3
    m.train()
    m.on_train_start()
    scores = m(wav_a.cuda(), wav_b.cuda())
    loss, prediction = m.loss_fn(scores, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < 5e-3
    assert np.abs(prediction.cpu().numpy() - g["prediction"]).max() < 2e-3

    def rel(a, b):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)

    assert rel(m.linear.weight.grad.cpu().numpy(), g["grad.linear.weight"]) < 1e-2
    assert rel(m.linear.bias.grad.cpu().numpy(), g["grad.linear.bias"]) < 1e-2
    worst = (-1.0, "")
    for k, q in m.wav2vec.model.named_parameters():
        if k.startswith("feature_extractor.") or k == "masked_spec_embed":
            assert q.grad is None or q.grad.abs().max().item() == 0.0, k
            continue
        if k.endswith("k_proj.bias"):                 # exactly 0 in exact arithmetic: rounding noise on both sides
            continue
        grad = q.grad.detach().cpu()
        flat = grad.reshape(-1)
        step = max(1, flat.numel() // 256)
        norm_err = abs(grad.double().norm().item() - float(g[f"grad.{k}.norm"])) / float(g[f"grad.{k}.norm"])
        # 256 strided elements of a tensor: a noisier statistic than the norm-wise error of the whole tensor
        err = rel(flat[::step][:256].numpy(), g[f"grad.{k}.sample"])
        worst = max(worst, (err, k))
        assert norm_err < 1e-2 and err < 3e-2, (k, norm_err, err)
    print("worst sampled gradient error against the reference fixture", worst)


def test_full_size_batch_is_consistent_with_small_batches(base_params):
    """BASELINE.json's headline size (64 utterances of 3 s) through the property the domain offers: utterances are
    independent, so every row of the full-size result must equal what the same utterance gives in a batch of two (which
    is the size the oracle comparison runs at -- anchored here for the first two utterances), and permuting the batch
    must permute the outputs.  GEMM tiles span utterances and the stream-K split depends on the tile count, so the
    agreement is to fp32 summation order, not bit-exact."""
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import make_head_params, make_inputs
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    S = 5994
    m = Wav2vec2FCModule(Wav2vec2FCModuleConfig(stat_pooling_type="mean", test_stat_pooling_type="mean"), S, CrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    head = make_head_params(768, S, seed=1)
    with torch.no_grad():
        m.fc_list[-1][0].weight.copy_(head["fc.weight"]); m.fc_list[-1][0].bias.copy_(head["fc.bias"])
    m = m.cuda().eval()
    wav, labels = make_inputs(64, 48000, S, seed=5)
    x = wav[:, None, :].cuda()

    def rows(a, b):
        a, b = a.double().cpu(), b.double().cpu()
        return ((a - b).norm(dim=-1) / b.norm(dim=-1)).max().item()

    with torch.no_grad():
        emb, pred = m(x)
        loss, prob = m.loss_fn(pred, labels.cuda())
        assert emb.shape == (64, 768) and prob.shape == (64, S) and torch.isfinite(loss)
        for a in (0, 31, 62):
            emb2, pred2 = m(x[a:a + 2])
            assert rows(emb[a:a + 2], emb2) < 1e-4, a
            assert torch.equal(pred[a:a + 2].argmax(1), pred2.argmax(1)), a
        perm = torch.arange(63, -1, -1)
        emb_p, pred_p = m(x[perm.cuda()])
        assert rows(emb_p, emb[perm.cuda()]) < 1e-4
        assert torch.equal(pred_p.argmax(1), pred[perm.cuda()].argmax(1))
        torch.set_num_threads(8)
        ref = O.speaker_embedding(wav[:2], base_params, "mean")                  # 2 x 3 s on the CPU: a few seconds
    # north_star: 1e-3 on fp32 embeddings
    assert rows(emb[:2], ref) < 1e-3


def test_full_size_training_step_gradient_is_the_mean_of_half_batches(base_params):
    """The backward at BASELINE.json's full size (64 utterances of 3 s, regularisation off): the loss is a mean over
    utterances, so every parameter gradient of the full batch must be the average of the gradients of its two halves
    (same kernels, other tile counts / split-K factors / atomics order; the halves' activation gradients are exactly
    twice as large, which fp16 rounding commutes with)."""
    _need_cuda()
    from oracle.params import make_head_params, make_inputs
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    S = 5994
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type="mean", test_stat_pooling_type="mean", **ZERO_REG)
    m = Wav2vec2FCModule(cfg, S, CrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    head = make_head_params(768, S, seed=1)
    with torch.no_grad():
        m.fc_list[-1][0].weight.copy_(head["fc.weight"]); m.fc_list[-1][0].bias.copy_(head["fc.bias"])
    m = m.cuda().train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    wav, labels = make_inputs(64, 48000, S, seed=6)
    x, y = wav[:, None, :].cuda(), labels.cuda()

    def grads(lo, hi):
        m.zero_grad(set_to_none=True)
        emb, pred = m(x[lo:hi])
        loss, _ = m.loss_fn(pred, y[lo:hi])
        loss.backward()
        return loss.item(), {k: q.grad.detach().double() for k, q in m.named_parameters() if q.grad is not None}

    la, ga = grads(0, 32)
    lb, gb = grads(32, 64)
    lf, gf = grads(0, 64)
    assert abs(lf - 0.5 * (la + lb)) < 1e-5 * abs(lf)
    assert set(gf) == set(ga) == set(gb) and len(gf) > 190
    worst = (-1.0, "")
    for k, g in gf.items():
        if k.endswith("k_proj.bias"):                 # exactly 0 in exact arithmetic: rounding noise in all three runs
            continue
        want = 0.5 * (ga[k] + gb[k])
        err = ((g - want).norm() / want.norm().clamp_min(1e-30)).item()
        worst = max(worst, (err, k))
        assert err < 2e-3, (k, err)
    print("worst full-batch vs half-batch gradient difference", worst)
