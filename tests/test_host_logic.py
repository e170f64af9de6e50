"""CPU tests of the host-side mirror of the reference interface (no kernels are launched)."""
import pytest
import torch

from w2v2_speaker_b200.engine import BASE, LARGE, LARGE_LV60, arch_from_id
from w2v2_speaker_b200.models.wav2vec2 import (Wav2Vec2RegularisationConfig, Wav2Vec2WrapperModule,
                                               Wav2vecLiteWrapperModule)
from w2v2_speaker_b200.optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss
from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig


def test_arch_detection_follows_reference_substring_rule():
    assert arch_from_id("facebook/wav2vec2-base") is BASE
    assert arch_from_id("facebook/wav2vec2-large") is LARGE
    # the stable-layer-norm checkpoints (HF configs: feat_extract_norm="layer", conv_bias, do_stable_layer_norm)
    assert arch_from_id("facebook/wav2vec2-large-lv60") is LARGE_LV60
    assert arch_from_id("facebook/wav2vec2-large-xlsr-53") is LARGE_LV60
    assert LARGE_LV60.stable_layer_norm and LARGE_LV60.conv_bias and LARGE_LV60.feat_extract_norm == "layer"
    with pytest.raises(ValueError):
        arch_from_id("facebook/hubert")
    assert BASE.conv_lengths(48000) == [9599, 4799, 2399, 1199, 599, 299, 149]
    assert BASE.conv_lengths(16000)[-1] == 49 and LARGE.conv_lengths(80000)[-1] == 249


def test_regularisation_config_defaults_match_reference():
    r = Wav2Vec2RegularisationConfig()
    assert (r.activation_dropout, r.attention_dropout, r.feat_proj_dropout, r.hidden_dropout, r.layerdrop) == \
        (0.0, 0.1, 0.1, 0.1, 0.05)
    assert (r.mask_time_prob, r.mask_time_length, r.mask_feature_prob, r.mask_feature_length) == (0.05, 10, 0.0, 10)


@pytest.fixture(scope="module")
def wrapper():
    return Wav2Vec2WrapperModule("facebook/wav2vec2-base", reset_weights=False)


def test_wrapper_surface(wrapper):
    assert wrapper.num_features == 768 and wrapper.num_embedding_features == 768
    assert sum(p.numel() for p in wrapper.model.parameters()) == 94_371_712
    for attr in ("feature_extractor", "feature_projection", "encoder"):
        assert hasattr(wrapper.model, attr)
    with pytest.raises(ValueError):
        Wav2Vec2WrapperModule("facebook/other", reset_weights=False)
    assert Wav2vecLiteWrapperModule.num_features == 512


def test_state_dict_keys_match_huggingface(wrapper):
    tr = pytest.importorskip("transformers")
    hf = tr.Wav2Vec2Model(tr.Wav2Vec2Config())
    ours = {k: tuple(v.shape) for k, v in wrapper.model.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in hf.state_dict().items()}
    assert ours == theirs


def test_freeze_unfreeze_semantics(wrapper):
    wrapper.freeze()
    assert not wrapper.training and all(not p.requires_grad for p in wrapper.parameters())
    wrapper.unfreeze()
    assert wrapper.training and all(p.requires_grad for p in wrapper.parameters())
    wrapper.model.feature_extractor.requires_grad_(False)        # R:.../wav2vec2_fc.py:346-347
    assert all(not p.requires_grad for p in wrapper.model.feature_extractor.parameters())
    assert any(p.requires_grad for p in wrapper.model.encoder.parameters())
    wrapper.eval()


def test_no_cpu_fallback(wrapper):
    wrapper.eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            wrapper(torch.zeros(1, 4000))


def test_training_mode_without_cuda_fails_loudly(wrapper):
    wrapper.train()
    with pytest.raises((NotImplementedError, RuntimeError)):
        wrapper(torch.zeros(1, 4000))
    wrapper.eval()


def test_time_mask_follows_hf_span_rules():
    import numpy as np
    from w2v2_speaker_b200.training import compute_time_mask
    rng = np.random.default_rng(0)
    m = compute_time_mask(64, 149, 0.05, 10, 2, rng).reshape(64, 149)
    per_utt = m.sum(1)
    assert per_utt.max() <= 20 and per_utt.min() >= 10       # two spans of 10 frames, possibly overlapping
    assert compute_time_mask(2, 5, 0.05, 10, 2, rng).sum() == 0   # utterance shorter than one span


@pytest.mark.parametrize("pooling,dim", [("mean", 768), ("mean+std", 1536), ("attentive", 1536), ("max", 768),
                                         ("quantile", 3840), ("first", 768), ("none", 768)])
def test_fc_module_pooling_dispatch(pooling, dim):
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type=pooling, test_stat_pooling_type=pooling)
    m = Wav2vec2FCModule(cfg, 100, CrossEntropyLoss)
    assert m.stat_pool_dimension == dim
    assert len(m.fc_list) == 1 and m.fc_list[0][0].out_features == 100
    with pytest.raises(ValueError):
        Wav2vec2FCModule(Wav2vec2FCModuleConfig(stat_pooling_type="bogus", test_stat_pooling_type="bogus"), 10,
                         CrossEntropyLoss)


def test_aam_head_surgery_and_constants():
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type="mean+std", test_stat_pooling_type="mean+std")
    ctor = lambda: AngularAdditiveMarginSoftMaxLoss(input_features=1, output_features=1, margin=0.2, scale=30)
    m = Wav2vec2FCModule(cfg, 5994, ctor)
    assert len(m.fc_list) == 0                                     # R:.../wav2vec2_fc.py:212-214
    assert tuple(m.loss_fn.fc_weights.shape) == (5994, 1536)
    assert m.loss_fn.margin == 0.2 and m.loss_fn.scale == 30
    assert abs(m.loss_fn.th - (-0.980067)) < 1e-5 and abs(m.loss_fn.mm - 0.0397339) < 1e-6
    with pytest.raises(ValueError):
        Wav2vec2FCModule(Wav2vec2FCModuleConfig(wav2vec_feature_encoder_only=True), 10, CrossEntropyLoss)


def test_embedding_masker_is_identity_on_the_path():
    from w2v2_speaker_b200.layers.embedding_masking import EmbeddingMasker
    mk = EmbeddingMasker(0, 1, 0.5, 5, time_dim=2, embedding_dim=1).train()
    x = torch.randn(3, 8, 1)
    assert torch.equal(mk(x), x)
    with pytest.raises(ValueError):
        EmbeddingMasker(1.5, 1, 0, 1)


@pytest.mark.parametrize("seed,n,ties", [(0, 2000, False), (1, 500, True), (2, 37, False), (3, 10000, True)])
def test_eer_and_min_dcf_match_the_reference_formulation(seed, n, ties):
    """eval_metrics (numpy) against the reference's sklearn / scipy / list-loop formulation (oracle/eval_oracle.py)."""
    import numpy as np
    from oracle import eval_oracle as EO
    from w2v2_speaker_b200.eval_metrics import calculate_eer, calculate_mdc
    rng = np.random.default_rng(seed)
    gt = rng.integers(0, 2, n)
    sc = np.clip(0.5 + 0.18 * rng.standard_normal(n) + 0.15 * (gt - 0.5) * 2, 0, 1)
    if ties:
        sc = np.round(sc, 2)
    e, t = calculate_eer(gt.tolist(), sc.tolist())
    re, rt = EO.eer(gt.tolist(), sc.tolist())
    assert abs(e - re) < 1e-9
    if np.isfinite(rt):
        assert abs(t - rt) < 1e-6
    m, mt = calculate_mdc(gt.tolist(), sc.tolist())
    rm, rmt = EO.mdc(gt.tolist(), sc.tolist())
    assert abs(m - rm) < 1e-9 and abs(mt - rmt) < 1e-12
    with pytest.raises(ValueError):
        calculate_eer([0, 2, 1], [0.1, 0.2, 0.3])
    with pytest.raises(ValueError):
        calculate_mdc([0, 1], [0.1])


def test_reference_era_checkpoints_load(wrapper):
    """transformers 4.x key names (weight_g / weight_v), HF task-model checkpoints (wav2vec2. prefix + heads)
    and Lightning checkpoints of the reference's Wav2vec2FCModule load into the modules of this package."""
    import torch
    from w2v2_speaker_b200 import checkpoint as C
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    sd = {k: torch.randn_like(v) for k, v in wrapper.model.state_dict().items()}
    g5, v5 = ("encoder.pos_conv_embed.conv.parametrizations.weight.original0",
              "encoder.pos_conv_embed.conv.parametrizations.weight.original1")
    old = dict(sd)
    old["encoder.pos_conv_embed.conv.weight_g"] = old.pop(g5)
    old["encoder.pos_conv_embed.conv.weight_v"] = old.pop(v5)
    # 1. plain load_state_dict with 4.x names (pre-hook)
    res = wrapper.model.load_state_dict(dict(old))
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(wrapper.model.state_dict()[g5], sd[g5]) and torch.equal(wrapper.model.state_dict()[v5], sd[v5])
    # 2. a Wav2Vec2ForCTC-style checkpoint
    ctc = {"wav2vec2." + k: v for k, v in old.items()}
    ctc["lm_head.weight"] = torch.zeros(32, 768)
    conv = C.convert_hf_state_dict(ctc)
    assert set(conv) == set(sd)
    # 3. a Lightning checkpoint of the reference's speaker module
    m = Wav2vec2FCModule(Wav2vec2FCModuleConfig(stat_pooling_type="mean", test_stat_pooling_type="mean"), 11, CrossEntropyLoss)
    ours = m.state_dict()
    ckpt = {"state_dict": {}}
    for k, v in ours.items():
        k4 = k.replace("parametrizations.weight.original0", "weight_g").replace("parametrizations.weight.original1", "weight_v")
        ckpt["state_dict"][k4] = torch.randn_like(v) if v.is_floating_point() else v.clone()
    missing, unexpected = C.load_reference_checkpoint(m, ckpt)
    assert not list(missing) and not list(unexpected)
    k = "wav2vec.model.encoder.pos_conv_embed.conv.parametrizations.weight.original1"
    assert torch.equal(m.state_dict()[k], ckpt["state_dict"][k.replace("parametrizations.weight.original1", "weight_v")])
    assert torch.equal(m.state_dict()["fc_list.0.0.weight"], ckpt["state_dict"]["fc_list.0.0.weight"])


def test_profile_tools_reproduce_the_committed_tables():
    """tools/launch_summary.py and tools/roofline_table.py on the committed end-of-round launch list: the step they cut
    out, the launch count and the headline rows are the ones profiles/ holds."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r01_g_train_launches.csv")
    # (round 1's LayerNorm backward read two gradient streams: the "two" argument selects its byte model)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "roofline_table.py"), csv_path, "1373.2", "6549.0", "two"],
                         check=True, capture_output=True, text=True).stdout
    assert out == open(os.path.join(root, "profiles", "r01_g_kernel_roofline_table.txt")).read()
    # round 2's list (persistent attention, accumulating data-gradient GEMMs, one gradient stream into the LayerNorm backward)
    out2 = subprocess.run([sys.executable, os.path.join(root, "tools", "roofline_table.py"),
                           os.path.join(root, "profiles", "r02_c_train_launches.csv")], check=True, capture_output=True, text=True).stdout
    assert out2 == open(os.path.join(root, "profiles", "r02_c_kernel_roofline_table.txt")).read()
    assert "attention backward, persistent" in out2 and "one gradient stream in" in out2
    # the launch list of the final bench command of round 2 (ring LayerNorm backward, 64 x 64-tile weight preparation)
    out3 = subprocess.run([sys.executable, os.path.join(root, "tools", "roofline_table.py"),
                           os.path.join(root, "profiles", "r02_i_train_launches.csv")], check=True, capture_output=True, text=True).stdout
    assert out3 == open(os.path.join(root, "profiles", "r02_i_kernel_roofline_table.txt")).read()
    assert "290 launches" in out3.splitlines()[1] and "re-derivation of the fp16 operand copies" in out3
    lines = out.splitlines()
    assert "270 launches" in lines[1]
    shares = [float(l.split("%")[0].split()[-1]) for l in lines[3:]]
    assert abs(sum(shares) - 100.0) < 1.0
    summ = subprocess.run([sys.executable, os.path.join(root, "tools", "launch_summary.py"), csv_path, "--step"], check=True,
                          capture_output=True, text=True).stdout
    assert summ.splitlines()[0].startswith("# 270 launches")


def test_eval_metrics_match_the_reference_functions():
    """tests/golden/ref_eval.npz holds outputs of the reference's OWN calculate_eer / calculate_mdc and
    CosineDistanceEvaluator.evaluate (oracle/make_golden.py eval).  The product's host-side metric code
    (w2v2_speaker_b200/eval_metrics.py) reproduces them; so does the restatement the GPU evaluator test uses
    (oracle/eval_oracle.py), fed with CPU cosine scores."""
    import numpy as np
    from conftest import golden
    from oracle import eval_oracle as EO
    from w2v2_speaker_b200.eval_metrics import calculate_eer, calculate_mdc
    g = golden("ref_eval.npz")
    gt, pred = g["metric.gt"].tolist(), g["metric.pred"].tolist()
    e, et = calculate_eer(gt, pred)
    m, mt = calculate_mdc(gt, pred)
    assert abs(e - float(g["metric.eer"])) < 1e-9 and abs(et - float(g["metric.eer_threshold"])) < 1e-9
    assert abs(m - float(g["metric.mdc"])) < 1e-9 and abs(mt - float(g["metric.mdc_threshold"])) < 1e-9
    oe, oet = EO.eer(gt, pred)
    om, omt = EO.mdc(gt, pred)
    assert abs(oe - e) < 1e-12 and abs(oet - et) < 1e-12 and abs(om - m) < 1e-12 and abs(omt - mt) < 1e-12
    emb = torch.from_numpy(g["embeddings"])
    left, right = emb[g["left"]], emb[g["right"]]
    same = g["same"].astype(int).tolist()
    for key, center in (("plain", False), ("center", True)):
        mean = std = None
        if center:
            std, mean = torch.std_mean(emb[::int(g["fit_stride"])], dim=0)
        scores = EO.cosine_scores(left, right, mean, std).numpy().astype(np.float64)
        pred = np.clip((scores + 1) / 2, 0, 1).tolist()
        e, et = calculate_eer(same, pred)
        m, mt = calculate_mdc(same, pred)
        assert abs(e - float(g[key + ".eer"])) < 1e-6, key
        assert abs(et - float(g[key + ".eer_threshold"])) < 1e-5, key
        assert abs(m - float(g[key + ".mdc"])) < 1e-6 and abs(mt - float(g[key + ".mdc_threshold"])) < 1e-5, key


def test_bucket_planner_covers_every_utterance_once():
    """ragged.plan_buckets: every index exactly once, batch size / padding / memory limits respected."""
    import random
    from w2v2_speaker_b200.ragged import plan_buckets
    random.seed(1)
    lengths = [random.randint(16000, 400000) for _ in range(300)] + [48000] * 40
    for max_batch, pad in ((32, 0.15), (8, 0.0), (64, 0.5)):
        buckets = plan_buckets(lengths, max_batch, pad, max_batch_samples=32 * 160000)
        assert sorted(i for b in buckets for i in b) == list(range(len(lengths)))
        for b in buckets:
            longest = max(lengths[i] for i in b)
            assert len(b) <= max_batch
            assert min(lengths[i] for i in b) >= (1.0 - pad) * longest - 1e-9
            assert len(b) == 1 or len(b) * longest <= 32 * 160000
    assert plan_buckets([], 4) == []
    with pytest.raises(ValueError):
        plan_buckets([100, 0], 4)


def test_accelerate_heads_retypes_plain_linears_in_place():
    """integration.accelerate_heads: the classifier layers the reference builds itself (nn.Linear, R:src/lightning_modules/
    speaker/wav2vec2_fc.py:176-210) become SpeakerLinear without touching parameters or state_dict keys; idempotent."""
    import torch.nn as nn
    from w2v2_speaker_b200.integration import accelerate_heads
    from w2v2_speaker_b200.layers.linear import SpeakerLinear

    class Holder(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc_list = nn.ModuleList([nn.Sequential(nn.Linear(8, 4), nn.ReLU()), nn.Sequential(nn.Linear(4, 3))])
            self.other = nn.Linear(2, 2)
    m = Holder()
    keys, w = list(m.state_dict().keys()), m.fc_list[0][0].weight
    assert accelerate_heads(m) == 2 and accelerate_heads(m) == 0
    assert all(isinstance(seq[0], SpeakerLinear) for seq in m.fc_list) and type(m.other) is nn.Linear
    assert list(m.state_dict().keys()) == keys and m.fc_list[0][0].weight is w


def test_synthetic_batch_is_standardised_and_reproducible():
    from w2v2_speaker_b200.synthetic import synthetic_batch
    x, y = synthetic_batch(4, 16000, 100, seed=3)
    x2, y2 = synthetic_batch(4, 16000, 100, seed=3)
    assert torch.equal(x, x2) and torch.equal(y, y2) and x.shape == (4, 16000) and y.dtype == torch.int64
    assert x.mean(1).abs().max() < 1e-5 and (x.std(1) - 1).abs().max() < 1e-3 and int(y.max()) < 100


def test_stable_layer_norm_variant_has_hf_parameter_names_and_is_evaluation_only():
    """-lv60 / XLSR mirror: the parameter names are those of the HF model with feat_extract_norm="layer", conv_bias,
    do_stable_layer_norm (checked against a 2-layer HF instance of that configuration); training with the CNN unfrozen
    raises (the layer-norm feature extractor has no backward here)."""
    import dataclasses
    import transformers as tr
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2ModelB200, init_hf_parameters
    arch = dataclasses.replace(LARGE_LV60, layers=2)
    cfg = tr.Wav2Vec2Config(hidden_size=1024, num_hidden_layers=2, num_attention_heads=16, intermediate_size=4096,
                            feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True)
    hf = tr.Wav2Vec2Model(cfg).state_dict()
    mine = init_hf_parameters(arch)
    assert set(mine) == set(hf)
    assert all(tuple(mine[k].shape) == tuple(hf[k].shape) for k in hf)
    m = Wav2Vec2ModelB200(arch).train()
    with pytest.raises(NotImplementedError, match="CNN frozen"):
        m(torch.zeros(1, 16000))
