"""Pins the CPU oracle (oracle/w2v2_oracle.py) to outputs of the reference itself
(tests/golden/ref_*.npz, produced by oracle/make_golden.py from /root/reference's own
Wav2vec2FCModule) and, where transformers is importable, to the live HF Wav2Vec2Model."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import w2v2_oracle as O
from oracle.params import BASE, make_asp_params, make_head_params, make_inputs

CASES = [("ref_cfg0_b2_1s.npz", 2, 16000), ("ref_b3_ragged_0p7s.npz", 3, 11283)]


def rel(a, b):
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def check_summary(name, t, g, tol=2e-5):
    t = t.detach().float()
    assert list(t.shape) == list(g[name + ".shape"])
    flat = t.reshape(-1)
    step = max(1, flat.numel() // 4096)
    assert rel(flat[::step][:4096], g[name + ".sample"]) < tol, name
    assert abs(t.double().norm().item() - float(g[name + ".norm"])) / float(g[name + ".norm"]) < tol, name


@pytest.mark.parametrize("fname,B,N", CASES)
def test_oracle_matches_reference_fixtures(fname, B, N, base_params):
    torch.set_num_threads(8)
    g = golden(fname)
    wav, labels = make_inputs(B, N, 5994, seed=1234)
    assert np.array_equal(labels.numpy(), g["labels"])
    assert np.allclose(wav.reshape(-1)[::97][:2048].numpy(), g["wav.sample"])
    trace = {}
    with torch.no_grad():
        h = O.wav2vec2_forward(wav, base_params, BASE, trace)
        assert rel(h, g["last_hidden_state"]) < 2e-5
        for i, c in enumerate(trace["conv"]):
            check_summary(f"conv.{i}", c, g)
        for i, hs in enumerate(trace["hidden_states"]):
            check_summary(f"hidden_states.{i}", hs, g)
        # mean + CE
        emb = O.mean_pool(h)
        assert rel(emb, g["mean.ce.embedding"]) < 2e-5
        hp = make_head_params(768, 5994, seed=1)
        logits, loss, sm = O.cross_entropy_head(emb, hp["fc.weight"], hp["fc.bias"], labels)
        assert rel(logits, g["mean.ce.logits"]) < 2e-5
        assert abs(loss.item() - float(g["mean.ce.loss"])) < 1e-4
        assert np.array_equal(sm.argmax(1).numpy(), g["mean.ce.argmax"])
        check_summary("mean.ce.softmax", sm, g, tol=1e-4)
        # mean+std + AAM
        emb = O.mean_std_pool(h)
        assert rel(emb, g["mean+std.aam.embedding"]) < 2e-5
        hp = make_head_params(1536, 5994, seed=1)
        _, loss, sm = O.aam_softmax(emb, hp["aam.fc_weights"], labels, 0.2, 30.0)
        assert abs(loss.item() - float(g["mean+std.aam.loss"])) < 1e-4
        assert np.array_equal(sm.argmax(1).numpy(), g["mean+std.aam.argmax"])
        # attentive + AAM
        emb = O.attentive_stat_pool(h, make_asp_params(768, seed=2))
        assert rel(emb, g["attentive.aam.embedding"]) < 2e-5
        _, loss, sm = O.aam_softmax(emb, hp["aam.fc_weights"], labels, 0.2, 30.0)
        assert abs(loss.item() - float(g["attentive.aam.loss"])) < 1e-4
        assert np.array_equal(sm.argmax(1).numpy(), g["attentive.aam.argmax"])


def test_oracle_matches_live_hf_model(base_params):
    """The encoder arithmetic lives in transformers (not in /root/reference): check the
    restatement against the installed HF Wav2Vec2Model directly (eager attention)."""
    tr = pytest.importorskip("transformers")
    torch.set_num_threads(8)
    cfg = tr.Wav2Vec2Config(mask_time_prob=0.0)
    cfg._attn_implementation = "eager"
    m = tr.Wav2Vec2Model(cfg).eval()
    res = m.load_state_dict(base_params, strict=False)
    assert not res.missing_keys
    wav, _ = make_inputs(2, 8000, seed=7)
    with torch.no_grad():
        ref = m(wav).last_hidden_state
        got = O.wav2vec2_forward(wav, base_params, BASE)
    assert rel(got, ref) < 1e-5


def test_stable_layer_norm_restatement_matches_live_hf_model():
    """The -lv60 / XLSR variant of the oracle (LayerNorm conv layers with bias, HF:275-299; pre-LN encoder layers and
    the final encoder LayerNorm, HF:632-655 / HF:731-799) against the live HF model of that configuration: last hidden
    state and every entry of ``hidden_states``."""
    tr = pytest.importorskip("transformers")
    from oracle import w2v2_oracle as O
    from oracle.params import ArchConfig, make_params
    arch = ArchConfig(name="small-lv60", hidden=256, layers=3, heads=4, ffn=512, conv_dim=64, feat_extract_norm="layer",
                      conv_bias=True, stable_layer_norm=True)
    p = make_params(arch, seed=3)
    cfg = tr.Wav2Vec2Config(hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=512,
                            conv_dim=(64,) * 7, feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
                            mask_time_prob=0.0)
    m = tr.Wav2Vec2Model(cfg).eval()
    missing, unexpected = m.load_state_dict(p, strict=False)
    assert not missing and unexpected == ["masked_spec_embed"]
    wav = torch.randn(2, 8000, generator=torch.Generator().manual_seed(1))
    trace = {}
    with torch.no_grad():
        out = m(wav, output_hidden_states=True)
        ref = O.wav2vec2_forward(wav, p, arch, trace)
    assert (out.last_hidden_state - ref).abs().max().item() < 1e-5
    assert len(out.hidden_states) == len(trace["hidden_states"]) == 4
    for a, b in zip(out.hidden_states, trace["hidden_states"]):
        assert (a - b).abs().max().item() < 1e-5


def test_pos_conv_weight_norm_formula(base_params):
    w = O.pos_conv_weight(base_params)
    v = base_params["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    g = base_params["encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    ref = torch._weight_norm(v, g, 2)
    assert rel(w, ref) < 1e-6


def test_attentive_pooling_restatement_matches_an_independent_port():
    """speechbrain's source is absent offline; transformers ships a port of the same ECAPA-TDNN layer (Qwen2.5-Omni's
    speaker encoder).  With the BatchNorm of speechbrain's TDNN block neutralised (the port has none) the oracle's
    restatement and the port agree exactly: statistics + clamp, [x | mean | std] context, tanh, 1x1 convs, softmax over
    time, [mean | std] output."""
    mod = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    from oracle import w2v2_oracle as O
    from oracle.params import make_asp_params
    C = 96
    asp = make_asp_params(C, seed=2)
    asp["tdnn.norm.norm.running_mean"].zero_()
    asp["tdnn.norm.norm.running_var"].fill_(1.0 - 1e-5)          # + eps = 1: the eval-mode BatchNorm is the identity
    asp["tdnn.norm.norm.weight"].fill_(1.0)
    asp["tdnn.norm.norm.bias"].zero_()
    port = mod.AttentiveStatisticsPooling(C, attention_channels=asp["tdnn.conv.conv.weight"].shape[0]).eval()
    with torch.no_grad():
        port.tdnn.conv.weight.copy_(asp["tdnn.conv.conv.weight"]); port.tdnn.conv.bias.copy_(asp["tdnn.conv.conv.bias"])
        port.conv.weight.copy_(asp["conv.conv.weight"]); port.conv.bias.copy_(asp["conv.conv.bias"])
        for T in (37, 149):
            x = torch.randn(3, T, C, generator=torch.Generator().manual_seed(T))
            ours = O.attentive_stat_pool(x, asp, training=False)
            theirs = port(x.transpose(1, 2)).squeeze(2)
            assert ours.shape == theirs.shape == (3, 2 * C)
            assert (ours - theirs).abs().max().item() <= 1e-6 * theirs.abs().max().item()


def test_split_path_restatement_matches_the_reference_paired_model(base_params):
    """tests/golden/ref_paired_b3.npz was produced by the reference's own Wav2vec2PairedSpeakerModule +
    BinaryCrossEntropyLoss (one training step, regularisation off) and by the CLS-token path of its
    Wav2Vec2WrapperModule (oracle/make_golden.py paired).  The oracle's composition -- split_path_forward, a Linear on the
    CLS position, BCE, torch autograd -- reproduces scores, loss and every gradient."""
    import torch.nn.functional as F
    g = golden("ref_paired_b3.npz")
    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    wav_a, _ = make_inputs(3, 16000, seed=31)
    wav_b, _ = make_inputs(3, 11283, seed=32)
    labels = torch.from_numpy(g["labels"])
    lw = torch.from_numpy(g["linear.weight"]).clone().requires_grad_(True)
    lb = torch.from_numpy(g["linear.bias"]).clone().requires_grad_(True)
    tokens = O.split_path_forward([wav_a, wav_b], p, [1.0, -1.0, -1.0])
    scores = F.linear(tokens[:, 0, :], lw, lb)
    loss = F.binary_cross_entropy_with_logits(scores.squeeze(), labels.float())
    loss.backward()
    assert np.abs(scores.detach().numpy() - g["scores.train"]).max() < 2e-5
    assert np.abs(scores.detach().numpy() - g["scores.eval"]).max() < 2e-5          # regularisation off: train == eval
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert np.abs(torch.sigmoid(scores.detach().squeeze()).numpy() - g["prediction"]).max() < 1e-5

    def close(a, b, tol=2e-4):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        return np.linalg.norm(a - b) <= tol * max(np.linalg.norm(b), 1e-12)

    assert close(lw.grad.numpy(), g["grad.linear.weight"]) and close(lb.grad.numpy(), g["grad.linear.bias"])
    checked = 0
    for k, v in p.items():
        if v.grad is None:
            assert k.startswith("feature_extractor.") or k == "masked_spec_embed", k
            continue
        if k.endswith("k_proj.bias"):                 # exactly 0 in exact arithmetic: both sides are rounding noise
            continue
        flat = v.grad.reshape(-1)
        step = max(1, flat.numel() // 256)
        assert tuple(g[f"grad.{k}.shape"]) == tuple(v.grad.shape), k
        assert abs(v.grad.double().norm().item() - float(g[f"grad.{k}.norm"])) <= 2e-4 * float(g[f"grad.{k}.norm"]), k
        assert close(flat[::step][:256].numpy(), g[f"grad.{k}.sample"], 1e-3), k
        checked += 1
    assert checked > 150
    with torch.no_grad():
        cls = O.split_path_forward([wav_b], {k: v.detach() for k, v in p.items()}, [1.0])
    assert cls.shape[1] == 36
    assert close(cls[:, 0, :].numpy(), g["cls.first_token"], 2e-5)


@pytest.mark.parametrize("pooling,loss", [("mean", "ce"), ("mean+std", "aam"), ("attentive", "aam")])
def test_oracle_autograd_matches_the_reference_training_step(base_params, pooling, loss):
    """tests/golden/ref_train_b2_1s.npz: loss and every parameter gradient of ONE training step of the reference's own
    Wav2vec2FCModule (oracle/make_golden.py train; regularisation off, CNN frozen, BatchNorm of the attentive pooling on
    batch statistics).  The GPU training tests compare the CUDA backward with autograd of the oracle; this pins that
    autograd to the reference's."""
    g = golden("ref_train_b2_1s.npz")
    torch.set_num_threads(8)
    S = 5994
    wav, labels = make_inputs(2, 16000, S, seed=1234)
    assert np.array_equal(labels.numpy(), g["labels"])
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    E = 768 if pooling == "mean" else 1536
    head = {k: v.clone().requires_grad_(True) for k, v in make_head_params(E, S, seed=1).items()}
    asp = None
    h = O.wav2vec2_forward(wav, p, BASE)
    if pooling == "attentive":
        asp = {k: (v.clone().requires_grad_(True) if "running" not in k else v.clone())
               for k, v in make_asp_params(768, seed=2).items()}
        emb = O.attentive_stat_pool(h, asp, training=True)
    else:
        emb = O.mean_pool(h) if pooling == "mean" else O.mean_std_pool(h)
    if loss == "ce":
        _, loss_v, _ = O.cross_entropy_head(emb, head["fc.weight"], head["fc.bias"], labels)
    else:
        _, loss_v, _ = O.aam_softmax(emb, head["aam.fc_weights"], labels, 0.2, 30.0)
    loss_v.backward()
    key = f"{pooling}.{loss}"
    assert abs(loss_v.item() - float(g[key + ".loss"])) < 2e-5 * float(g[key + ".loss"])
    ours = {"wav2vec.model." + k: v for k, v in p.items()}
    ours.update({"fc_list.0.0.weight": head.get("fc.weight"), "fc_list.0.0.bias": head.get("fc.bias"),
                 "loss_fn.fc_weights": head.get("aam.fc_weights")})
    if asp is not None:
        ours.update({"stat_pooling.pooling_layer." + k: v for k, v in asp.items()})
    names, norms, samples = g[key + ".grad.names"], g[key + ".grad.norms"], g[key + ".grad.samples"]
    assert len(names) > 200
    for name, norm, smp in zip(names, norms, samples):
        t = ours[str(name)]
        assert t is not None and t.grad is not None, name
        if str(name).endswith("k_proj.bias") or str(name).endswith("pooling_layer.conv.conv.bias"):
            continue                                   # exactly 0 in exact arithmetic (softmax shift invariance): noise
        flat = t.grad.reshape(-1)
        step = max(1, flat.numel() // 128)
        got = flat[::step][:128].double().numpy()
        ref = smp[:got.size].astype(np.float64)
        assert abs(t.grad.double().norm().item() - norm) <= 5e-4 * norm, name
        assert np.linalg.norm(got - ref) <= 2e-3 * max(np.linalg.norm(ref), 1e-12), name
    frozen = [k for k, v in p.items() if v.grad is None]
    assert all(k.startswith("feature_extractor.") or k == "masked_spec_embed" for k in frozen)
