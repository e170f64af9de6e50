"""GPU parity of the reference-facing Python classes (the drop-in boundary) against the fixtures
generated from the reference's own Wav2vec2FCModule and against the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
S = 5994


def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def rel_rows(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm(dim=-1) / b.norm(dim=-1)).max().item()


def build(pooling, loss, base_params):
    from oracle.params import make_asp_params, make_head_params
    from w2v2_speaker_b200.optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type=pooling, test_stat_pooling_type=pooling)
    ctor = CrossEntropyLoss if loss == "ce" else (
        lambda: AngularAdditiveMarginSoftMaxLoss(input_features=1, output_features=1, margin=0.2, scale=30))
    m = Wav2vec2FCModule(cfg, S, ctor)
    res = m.wav2vec.model.load_state_dict(base_params, strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    E = 768 if pooling == "mean" else 1536
    head = make_head_params(E, S, seed=1)
    with torch.no_grad():
        if loss == "ce":
            m.fc_list[-1][0].weight.copy_(head["fc.weight"]); m.fc_list[-1][0].bias.copy_(head["fc.bias"])
        else:
            m.loss_fn.fc_weights.copy_(head["aam.fc_weights"])
        if pooling == "attentive":
            r = m.stat_pooling.pooling_layer.load_state_dict(make_asp_params(768, seed=2), strict=False)
            assert not r.missing_keys or all("num_batches_tracked" in k for k in r.missing_keys)
    return m.cuda().eval()


@pytest.mark.parametrize("fname,B,N", [("ref_cfg0_b2_1s.npz", 2, 16000), ("ref_b3_ragged_0p7s.npz", 3, 11283)])
@pytest.mark.parametrize("pooling,loss", [("mean", "ce"), ("mean+std", "aam"), ("attentive", "aam")])
def test_module_matches_reference_fixture(base_params, fname, B, N, pooling, loss):
    _need_cuda()
    from oracle.params import make_inputs
    g = golden(fname)
    wav, labels = make_inputs(B, N, S, seed=1234)
    m = build(pooling, loss, base_params)
    with torch.no_grad():
        emb, pred = m(wav[:, None, :].cuda())                  # [B,1,N] batches as the reference feeds them
        loss_v, prob = m.loss_fn(pred, labels.cuda())
    key = f"{pooling}.{loss}"
    ref_emb = g[key + ".embedding"]
    assert tuple(emb.shape) == ref_emb.shape
    assert rel_rows(emb, ref_emb) < 1e-3                        # north_star: 1e-3 rel on fp32 embeddings
    assert abs(loss_v.item() - float(g[key + ".loss"])) / float(g[key + ".loss"]) < 1e-3
    assert np.array_equal(prob.argmax(1).cpu().numpy(), g[key + ".argmax"])      # bit-exact speaker id
    flat = prob.reshape(-1).cpu()
    step = max(1, flat.numel() // 4096)
    assert rel_rows(flat[::step][:4096][None], g[key + ".softmax.sample"][None]) < 5e-3
    if loss == "ce":
        assert rel_rows(pred, g[key + ".logits"]) < 1e-3


def test_pooling_layers_both_orientations():
    _need_cuda()
    from w2v2_speaker_b200.layers.pooling import MaxPool1D, MeanStatPool1D, MeanStdStatPool1D
    x = torch.randn(3, 37, 768, generator=torch.Generator().manual_seed(5)).cuda()
    assert rel_rows(MeanStatPool1D(1)(x), x.mean(1)) < 1e-6
    assert rel_rows(MeanStatPool1D(2)(x.transpose(1, 2)), x.mean(1)) < 1e-6
    assert rel_rows(MeanStdStatPool1D(1)(x), torch.cat(torch.std_mean(x, 1), 1)) < 1e-5
    assert torch.equal(MaxPool1D(1)(x), x.max(1).values)
    with pytest.raises(ValueError):
        MeanStatPool1D(0)(x)


def test_wrapper_cls_token_and_lite_paths(base_params):
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import BASE, make_inputs
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2WrapperModule, Wav2vecLiteWrapperModule
    wav, _ = make_inputs(2, 8000, seed=3)
    w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False, insert_clc_token=True)
    w.model.load_state_dict(base_params)
    w = w.cuda().eval()
    with torch.no_grad():
        got = w(wav.cuda())                                      # [B, H, T+1]
        feat = O.feature_extractor(wav, base_params, BASE).transpose(1, 2)
        proj = O.feature_projection(feat, base_params, BASE)
        seq = torch.cat([torch.ones(2, 1, 768), proj], 1)       # R:src/models/wav2vec2.py:128-140
        ref = O.encoder(seq, base_params, BASE).transpose(1, 2)
    assert got.shape == ref.shape
    assert rel_rows(got.transpose(1, 2).mean(1), ref.transpose(1, 2).mean(1)) < 1e-3
    lite = Wav2vecLiteWrapperModule("facebook/wav2vec2-base", False)
    lite.model.load_state_dict(base_params)
    lite = lite.cuda().eval()
    with torch.no_grad():
        f = lite(wav.cuda())
    assert f.shape == (2, 512, 24)
    assert ((f.cpu().double() - feat.transpose(1, 2).double()).norm() / feat.double().norm()).item() < 2e-3


def test_long_utterance_evaluates_and_trains_and_very_long_training_fails_loudly(base_params):
    """Utterances longer than 256 frames: the evaluation forward handles any length (full-utterance test_step).
    TRAINING runs the key-tiled attention kernels up to what the single-slab positional conv holds (about 1150 frames
    at wav2vec2-base; parity at 301 frames: tests/test_gpu_round2.py::test_paired_model_trains_at_the_reference_crop_length);
    beyond that it must raise, not fall back."""
    _need_cuda()
    from w2v2_speaker_b200._lib import W2V2Error
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2RegularisationConfig, Wav2Vec2WrapperModule
    # dropout / SpecAugment at the reference's defaults, LayerDrop off so that every layer must see a gradient
    w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False,
                              reg_cfg=Wav2Vec2RegularisationConfig(layerdrop=0.0)).cuda().eval()
    with torch.no_grad():
        out = w(torch.randn(1, 16000 * 6, device="cuda"))           # 299 frames > 256
    assert out.shape == (1, 768, 299) and torch.isfinite(out).all()
    w.train()
    w.model.feature_extractor.requires_grad_(False)
    out = w(torch.randn(2, 16000 * 6, device="cuda"))
    # (a random readout: sum(out) and sum(out^2) are constants of the final LayerNorm -- their gradients vanish)
    (out * torch.randn(out.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(1))).mean().backward()
    amax = [dict(w.model.named_parameters())[f"encoder.layers.{i}.attention.q_proj.weight"].grad.abs().max().item()
            for i in range(12)]
    assert all(0 < a < float("inf") for a in amax), amax
    with pytest.raises(W2V2Error):
        w(torch.randn(1, 16000 * 40, device="cuda")).sum().backward()      # 1999 frames


@pytest.mark.parametrize("center", [False, True])
def test_cosine_evaluator_matches_reference_formulation(center):
    """CosineDistanceEvaluator: GPU trial scoring against torch's CosineSimilarity with the evaluator's centring, EER /
    min-DCF against the reference's sklearn / scipy formulation (oracle/eval_oracle.py)."""
    _need_cuda()
    from oracle import eval_oracle as EO
    from w2v2_speaker_b200.evaluation.speaker import CosineDistanceEvaluator, EmbeddingSample, EvaluationPair
    g = torch.Generator().manual_seed(9)
    n_spk, per, E = 40, 6, 1536
    centers = torch.randn(n_spk, E, generator=g)
    emb = (centers[:, None, :] + 5.0 * torch.randn(n_spk, per, E, generator=g) + 0.5).reshape(-1, E)
    ids = [f"spk{s}/utt{u}" for s in range(n_spk) for u in range(per)]
    samples = [EmbeddingSample(i, e) for i, e in zip(ids, emb)]
    rng = np.random.default_rng(4)
    pairs = []
    for _ in range(3000):
        a, b = rng.integers(0, len(ids), 2)
        if a != b:
            pairs.append(EvaluationPair(ids[a].split("/")[0] == ids[b].split("/")[0], ids[a], ids[b]))
    ev = CosineDistanceEvaluator(center_before_scoring=center, length_norm_before_scoring=True, max_num_training_samples=0)
    ev.fit_parameters(list(emb[::3]), [])
    scores = ev.score_pairs(pairs, samples)
    idx = {k: i for i, k in enumerate(ids)}
    left = torch.stack([emb[idx[p.sample1_id]] for p in pairs]); right = torch.stack([emb[idx[p.sample2_id]] for p in pairs])
    mean = std = None
    if center:
        std, mean = torch.std_mean(torch.stack(list(emb[::3])), dim=0)
    ref = EO.cosine_scores(left, right, mean, std).numpy()
    assert np.abs(scores - ref).max() < 2e-6
    res = ev.evaluate(pairs, samples)
    gt = [1 if p.same_speaker else 0 for p in pairs]
    pred = np.clip((ref.astype(np.float64) + 1) / 2, 0, 1).tolist()
    re, _ = EO.eer(gt, pred)
    rm, _ = EO.mdc(gt, pred)
    assert abs(res["eer"] - re) < 1e-4 and abs(res["mdc"] - rm) < 1e-3
    assert 0.0 < res["eer"] < 0.5
    missing = ev.evaluate([EvaluationPair(True, "nope", ids[0])], samples)
    assert missing == {"eer": -1, "eer_threshold": -1, "mdc": -1, "mdc_threshold": -1}



def test_input_normaliser_on_device():
    """InputNormalizer2D.normalize(channel_wise=False) per utterance on the GPU, from float32 and from
    16-bit PCM, against the reference's tensor expression."""
    _need_cuda()
    from w2v2_speaker_b200.data.preprocess import InputNormalizer2D
    g = torch.Generator().manual_seed(12)
    x = (torch.randn(5, 48000, generator=g) * 0.05 + 0.01).clamp(-1, 1)
    std, mean = torch.std_mean(x, dim=1, keepdim=True)
    ref = (x - mean) / (std + 1e-5)
    got = InputNormalizer2D.normalize_batch(x.cuda())
    assert rel_rows(got, ref) < 1e-5
    pcm = (x * 32768).round().clamp(-32768, 32767).to(torch.int16)
    xq = pcm.float() / 32768
    std, mean = torch.std_mean(xq, dim=1, keepdim=True)
    got16 = InputNormalizer2D.normalize_batch(pcm.cuda())
    assert rel_rows(got16, (xq - mean) / (std + 1e-5)) < 1e-5
    one, m1, s1 = InputNormalizer2D.normalize(x[:1].cuda(), channel_wise=False)
    r1, rm, rs = InputNormalizer2D.normalize(x[:1], channel_wise=False)          # CPU tensors: the reference expression
    assert rel_rows(one, r1) < 1e-5 and abs(m1.item() - rm.item()) < 1e-7 and abs(s1.item() - rs.item()) < 1e-6
