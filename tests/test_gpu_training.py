"""GPU parity of the training step (forward + hand-written backward through torch.autograd) against
autograd of the CPU oracle: loss within 1e-3, every parameter gradient within 1e-2 norm-wise
(gradients under reduced-precision operands), CNN feature extractor frozen as in the
reference default, regularisation probabilities 0."""
import pytest
import torch

pytestmark = pytest.mark.gpu
S = 5994


def _module(base_params, pooling="mean"):
    from oracle.params import make_head_params
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type=pooling, test_stat_pooling_type=pooling, activation_dropout=0.0,
                                 attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                                 mask_time_prob=0.0, mask_feature_prob=0.0)
    m = Wav2vec2FCModule(cfg, S, CrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    head = make_head_params(768, S, seed=1)
    with torch.no_grad():
        m.fc_list[-1][0].weight.copy_(head["fc.weight"]); m.fc_list[-1][0].bias.copy_(head["fc.bias"])
    m = m.cuda().train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)          # R:.../wav2vec2_fc.py:346-347
    return m, head


@pytest.mark.parametrize("B,N", [(2, 16000), (3, 11283), (2, 80000)])      # 80000 samples = 5 s -> 249 frames
def test_training_step_gradients_match_oracle_autograd(base_params, B, N):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import w2v2_oracle as O
    from oracle.params import BASE, make_inputs
    wav, labels = make_inputs(B, N, S, seed=1234)
    m, head = _module(base_params)
    emb, pred = m(wav[:, None, :].cuda())
    loss, prob = m.loss_fn(pred, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    fw = head["fc.weight"].clone().requires_grad_(True)
    fb = head["fc.bias"].clone().requires_grad_(True)
    ref_emb = O.speaker_embedding(wav, p, "mean")
    _, ref_loss, _ = O.cross_entropy_head(ref_emb, fw, fb, labels)
    ref_loss.backward()

    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-3
    got = dict(m.wav2vec.model.named_parameters())
    worst = (0.0, None)
    for k, v in p.items():
        if k.startswith("feature_extractor"):
            assert got[k].grad is None
            continue
        if k == "masked_spec_embed":                  # unused without SpecAugment: no or zero gradient
            assert got[k].grad is None or got[k].grad.abs().max().item() == 0.0
            continue
        assert got[k].grad is not None, k
        g, r = got[k].grad.detach().cpu().double(), v.grad.double()
        if k.endswith("k_proj.bias"):
            # softmax is invariant to a per-row constant, so the key-bias gradient is exactly 0 in exact
            # arithmetic: both sides are rounding noise -> compare against the scale of the q-bias gradient
            scale = p[k.replace("k_proj", "q_proj")].grad.double().norm()
            assert g.norm() < 1e-2 * scale and r.norm() < 1e-2 * scale, k
            continue
        rel = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
        worst = max(worst, (rel, k))
        assert rel < 1e-2, (k, rel)
    lin = m.fc_list[-1][0]
    for g, r, k in ((lin.weight.grad, fw.grad, "fc.weight"), (lin.bias.grad, fb.grad, "fc.bias")):
        rel = ((g.cpu().double() - r.double()).norm() / r.double().norm()).item()
        assert rel < 1e-2, (k, rel)
    print("worst parameter-gradient error", worst)


def test_torch_adam_step_reduces_loss(base_params):
    """The drop-in contract: loss.backward() + a stock torch optimizer train the module."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle.params import make_inputs
    wav, labels = make_inputs(4, 16000, S, seed=7)
    m, _ = _module(base_params)
    opt = torch.optim.Adam([q for q in m.parameters() if q.requires_grad], lr=1e-4)
    x, y = wav[:, None, :].cuda(), labels.cuda()
    losses = []
    for _ in range(3):
        opt.zero_grad()
        emb, pred = m(x)
        loss, _ = m.loss_fn(pred, y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("B,N", [(2, 16000), (3, 11283)])
def test_unfrozen_feature_extractor_gradients_match_oracle_autograd(base_params, B, N):
    """completely_freeze_feature_extractor: false -- the CNN backward (GELU / GroupNorm / strided-conv data and
    weight gradients on the tap-GEMM and batched wgrad kernels) against autograd of the CPU oracle."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import w2v2_oracle as O
    from oracle.params import make_inputs
    wav, labels = make_inputs(B, N, S, seed=4321)
    m, head = _module(base_params)
    m.wav2vec.model.feature_extractor.requires_grad_(True)
    emb, pred = m(wav[:, None, :].cuda())
    loss, prob = m.loss_fn(pred, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(True) for k, v in base_params.items()}
    fw = head["fc.weight"].clone().requires_grad_(True)
    fb = head["fc.bias"].clone().requires_grad_(True)
    ref_emb = O.speaker_embedding(wav, p, "mean")
    _, ref_loss, _ = O.cross_entropy_head(ref_emb, fw, fb, labels)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-3
    got = dict(m.wav2vec.model.named_parameters())
    errs = {}
    for k, v in p.items():
        if k.startswith("feature_extractor") or k in ("feature_projection.projection.weight",
                                                       "encoder.layers.0.attention.q_proj.weight",
                                                       "encoder.layers.11.feed_forward.output_dense.weight"):
            assert got[k].grad is not None, k
            g, r = got[k].grad.detach().cpu().double(), v.grad.double()
            errs[k] = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
    print("unfrozen-CNN gradient errors", {k.replace("feature_extractor.conv_layers.", "conv"): f"{e:.2e}" for k, e in errs.items()})
    # the conv gradients travel through 12 transformer layers and up to 7 conv layers of fp16-operand arithmetic
    assert all(e < 2e-2 for e in errs.values()), errs


def test_training_step_meanstd_aam_matches_oracle_autograd(base_params):
    """configs[3] shape of the path: mean+std pooling + AAM-softmax (margin 0.2, scale 30)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import w2v2_oracle as O
    from oracle.params import make_head_params, make_inputs
    from w2v2_speaker_b200.optim.loss import AngularAdditiveMarginSoftMaxLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    B, N = 3, 16000
    wav, labels = make_inputs(B, N, S, seed=99)
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type="mean+std", test_stat_pooling_type="mean+std", activation_dropout=0.0,
                                 attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                                 mask_time_prob=0.0)
    m = Wav2vec2FCModule(cfg, S, lambda: AngularAdditiveMarginSoftMaxLoss(1, 1, margin=0.2, scale=30))
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    head = make_head_params(1536, S, seed=1)
    with torch.no_grad():
        m.loss_fn.fc_weights.copy_(head["aam.fc_weights"])
    m = m.cuda().train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    emb, pred = m(wav[:, None, :].cuda())
    loss, prob = m.loss_fn(pred, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    fw = head["aam.fc_weights"].clone().requires_grad_(True)
    ref_emb = O.speaker_embedding(wav, p, "mean+std")
    _, ref_loss, _ = O.aam_softmax(ref_emb, fw, labels, 0.2, 30.0)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-3
    got = dict(m.wav2vec.model.named_parameters())
    for k in ("encoder.layers.11.feed_forward.output_dense.weight", "encoder.layers.0.attention.q_proj.weight",
              "encoder.layers.5.layer_norm.weight", "feature_projection.projection.weight",
              "encoder.pos_conv_embed.conv.parametrizations.weight.original1"):
        g, r = got[k].grad.cpu().double(), p[k].grad.double()
        assert ((g - r).norm() / r.norm()).item() < 1e-2, k
    g, r = m.loss_fn.fc_weights.grad.cpu().double(), fw.grad.double()
    assert ((g - r).norm() / r.norm()).item() < 1e-2


@pytest.mark.parametrize("train_bn", [True, False])
def test_attentive_pooling_backward_matches_oracle_autograd(train_bn):
    """AttentiveStatPool1D in training (batch-statistics BatchNorm) and in eval mode with gradients: output,
    d x, every parameter gradient and the running statistics against autograd of the CPU oracle."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import w2v2_oracle as O
    from oracle.params import make_asp_params
    from w2v2_speaker_b200.layers.pooling import AttentiveStatPool1D
    B, T, C = 5, 149, 768
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, T, C, generator=g)
    dout = torch.randn(B, 2 * C, generator=g) * 1e-3
    asp = make_asp_params(C, seed=2)
    layer = AttentiveStatPool1D(C, dim_to_reduce=1)
    r = layer.pooling_layer.load_state_dict(asp, strict=False)
    assert not r.missing_keys or all("num_batches" in k for k in r.missing_keys)
    layer = layer.cuda().train(train_bn)
    xg = x.cuda().requires_grad_(True)
    out = layer(xg)
    out.backward(dout.cuda())                 # Function boundaries carry plain unscaled gradients
    torch.cuda.synchronize()

    ref_p = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in asp.items()}
    xr = x.clone().requires_grad_(True)
    ref = O.attentive_stat_pool(xr, ref_p, training=train_bn)
    ref.backward(dout)

    def rel(a, b):
        a = a.detach().cpu().double(); b = b.detach().double()
        return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

    got = dict(layer.pooling_layer.named_parameters())
    errs = {"out": rel(out, ref), "dx": rel(xg.grad, xr.grad)}
    for k, v in ref_p.items():
        if v.grad is not None:
            errs[k] = rel(got[k].grad.reshape(v.shape), v.grad)
    # softmax over time is invariant to a per-channel constant: the gradient of the second conv's bias is exactly
    # 0 in exact arithmetic, both sides are rounding noise -> compare against the scale of the weight gradient
    k2 = "conv.conv.bias"
    errs.pop(k2)
    assert got[k2].grad.norm().item() < 1e-3 * got["conv.conv.weight"].grad.norm().item()
    print("ASP errors", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["out"] < 1e-3
    assert all(v < 1e-2 for v in errs.values()), errs
    bn = layer.pooling_layer.tdnn.norm.norm
    if train_bn:
        # torch semantics of the running statistics (momentum 0.1, unbiased variance)
        a = torch.relu(torch.nn.functional.conv1d(torch.cat([
            x.transpose(1, 2), x.mean(1)[:, :, None].expand(B, C, T),
            x.var(1, unbiased=False).clamp(1e-12).sqrt()[:, :, None].expand(B, C, T)], 1),
            asp["tdnn.conv.conv.weight"], asp["tdnn.conv.conv.bias"]))
        rm = 0.9 * asp["tdnn.norm.norm.running_mean"] + 0.1 * a.mean((0, 2))
        rv = 0.9 * asp["tdnn.norm.norm.running_var"] + 0.1 * a.transpose(1, 2).reshape(-1, a.shape[1]).var(0, unbiased=True)
        assert rel(bn.running_mean, rm) < 2e-3 and rel(bn.running_var, rv) < 2e-3
    else:
        assert torch.equal(bn.running_mean.cpu(), asp["tdnn.norm.norm.running_mean"])


def test_training_step_attentive_aam_matches_oracle_autograd(base_params):
    """configs[2]: attentive-statistics pooling + AAM-softmax (margin 0.2, scale 30), training mode."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import w2v2_oracle as O
    from oracle.params import make_asp_params, make_head_params, make_inputs
    from w2v2_speaker_b200.optim.loss import AngularAdditiveMarginSoftMaxLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    B, N = 3, 16000
    wav, labels = make_inputs(B, N, S, seed=77)
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type="attentive", test_stat_pooling_type="attentive", activation_dropout=0.0,
                                 attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                                 mask_time_prob=0.0)
    m = Wav2vec2FCModule(cfg, S, lambda: AngularAdditiveMarginSoftMaxLoss(1, 1, margin=0.2, scale=30))
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    asp = make_asp_params(768, seed=2)
    m.stat_pooling.pooling_layer.load_state_dict(asp, strict=False)
    head = make_head_params(1536, S, seed=1)
    with torch.no_grad():
        m.loss_fn.fc_weights.copy_(head["aam.fc_weights"])
    m = m.cuda().train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    emb, pred = m(wav[:, None, :].cuda())
    loss, prob = m.loss_fn(pred, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    ap = {k: v.clone().requires_grad_("running" not in k) for k, v in asp.items()}
    fw = head["aam.fc_weights"].clone().requires_grad_(True)
    h = O.wav2vec2_forward(wav, p)
    ref_emb = O.attentive_stat_pool(h, ap, training=True)
    _, ref_loss, _ = O.aam_softmax(ref_emb, fw, labels, 0.2, 30.0)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-3
    got = dict(m.wav2vec.model.named_parameters())
    for k in ("encoder.layers.11.feed_forward.output_dense.weight", "encoder.layers.0.attention.q_proj.weight",
              "encoder.layers.5.layer_norm.weight", "feature_projection.projection.weight"):
        g, r = got[k].grad.cpu().double(), p[k].grad.double()
        assert ((g - r).norm() / r.norm()).item() < 1e-2, k
    gp = dict(m.stat_pooling.pooling_layer.named_parameters())
    for k, v in ap.items():
        if v.grad is not None and k != "conv.conv.bias":          # exactly-zero gradient (shift invariance of softmax)
            g, r = gp[k].grad.cpu().double().reshape(v.shape), v.grad.double()
            # The TDNN conv sits behind a ReLU whose input here carries the encoder's ~1e-3 reduced-precision error:
            # units within that distance of zero flip their derivative, and over 147 rows x 128 units a handful of
            # flips is several percent of the gradient norm.  With identical inputs the same kernels meet 1e-2
            # (test_attentive_pooling_backward_matches_oracle_autograd).
            tol = 0.15 if k.startswith("tdnn.conv.conv") else 1e-2
            assert ((g - r).norm() / r.norm()).item() < tol, k
    g, r = m.loss_fn.fc_weights.grad.cpu().double(), fw.grad.double()
    assert ((g - r).norm() / r.norm()).item() < 1e-2


@pytest.mark.parametrize("pooling", ["mean", "first+cls"])
def test_flat_adam_trainer_matches_torch_adam(base_params, pooling):
    """trainer.FlatAdamTrainer (gradient sink, fused Adam, optimizer stream overlapped with the next step's CNN
    forward, in-place refresh of the fp16 operand copies) against loss.backward() + torch.optim.Adam on an
    identical module: same losses step by step, same parameter updates.  "first+cls" drives the split call path
    (CLS token in front of the sequence), whose Functions write into the same sink."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle.params import make_inputs
    from w2v2_speaker_b200.trainer import FlatAdamTrainer
    wav, labels = make_inputs(4, 16000, S, seed=5)
    x, y = wav[:, None, :].cuda(), labels.cuda()
    ma, _ = _module(base_params, pooling)
    mb, _ = _module(base_params, pooling)
    start = {k: v.detach().clone() for k, v in mb.named_parameters()}
    tr = FlatAdamTrainer(ma, lr=1e-4)
    opt = torch.optim.Adam([q for q in mb.parameters() if q.requires_grad], lr=1e-4)
    la, lb = [], []
    for _ in range(3):
        loss, _ = tr.step(x, y)
        la.append(loss.item())
        opt.zero_grad()
        emb, pred = mb(x)
        loss_b, _ = mb.loss_fn(pred, y)
        loss_b.backward()
        opt.step()
        lb.append(loss_b.item())
    tr.synchronize()
    torch.cuda.synchronize()
    assert la[-1] < la[0]
    for a, b in zip(la, lb):
        assert abs(a - b) / abs(b) < 2e-3, (la, lb)
    pa = dict(ma.named_parameters())
    errs = {}
    for k, vb in mb.named_parameters():
        if not vb.requires_grad or k.endswith("k_proj.bias"):      # key bias: exactly-zero gradient, pure rounding noise
            continue
        da = (pa[k].detach() - start[k]).double()
        db = (vb.detach() - start[k]).double()
        if db.norm().item() == 0.0:
            assert da.norm().item() == 0.0, k
            continue
        errs[k] = ((da - db).norm() / db.norm()).item()
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("largest update differences", [(k, f"{e:.2e}") for k, e in worst])
    # Adam normalises the gradient (the update is O(lr) per element whatever the gradient's size), so elements whose
    # gradient is at rounding level move differently in the two runs (the wgrad reduce-adds are not ordered); the
    # bulk of every tensor must agree, and the large matrices closely
    vals = sorted(errs.values())
    if pooling == "mean":
        assert vals[len(vals) // 2] < 0.05, worst
        assert errs["wav2vec.model.encoder.layers.5.feed_forward.intermediate_dense.weight"] < 0.1, worst
        assert max(vals) < 0.8, worst
    else:      # one token carries the whole gradient: more elements sit at rounding level, the bulk must still agree
        assert vals[len(vals) // 2] < 0.1, worst


def test_inplace_weight_refresh_equals_rebuild(base_params):
    """After a fused optimizer step the fp16 / transposed / folded operand copies are re-derived by ONE batched
    launch (w2v2_prepare_weights); they must equal what a from-scratch preparation of the new parameters gives."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from w2v2_speaker_b200.engine import PreparedWeights
    from w2v2_speaker_b200.training import TrainWeights
    m, _ = _module(base_params)
    model = m.wav2vec.model
    eng = model._engine()
    tw = model._train_weights(eng)
    ptrs = (eng.w.layers[3]["wqkv"].data_ptr(), tw.layers[3]["w1T"].data_ptr())
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in model.parameters():                       # raw in-place update, invisible to autograd versions
            torch.add(p.data, torch.randn(p.shape, generator=g).to(p.device) * 0.01, out=p.data)
    model.refresh()
    assert model._engine() is eng                          # updated in place, not rebuilt
    assert ptrs == (eng.w.layers[3]["wqkv"].data_ptr(), tw.layers[3]["w1T"].data_ptr())
    named = dict(model.named_parameters())
    fresh = PreparedWeights(named, model.arch)
    fresh_t = TrainWeights(fresh, named)
    H = model.arch.hidden
    scale = float(H // model.arch.heads) ** -0.5
    for l in (0, 5, 11):
        for k in ("wqkv", "bqkv", "wo", "w1", "w2"):
            assert torch.equal(eng.w.layers[l][k], fresh.layers[l][k]), (l, k)
        for k in ("wqkvT", "woT", "w1T", "w2T"):
            assert torch.equal(tw.layers[l][k], fresh_t.layers[l][k]), (l, k)
        pre = f"encoder.layers.{l}."
        ref = torch.cat([named[pre + "attention.q_proj.weight"] * scale, named[pre + "attention.k_proj.weight"],
                         named[pre + "attention.v_proj.weight"]], 0).half()
        assert torch.equal(eng.w.layers[l]["wqkv"], ref)
        # the transposed (data-gradient) copy keeps the q block UNSCALED: the attention backward scales dq instead
        ref_t = torch.cat([named[pre + "attention.q_proj.weight"], named[pre + "attention.k_proj.weight"],
                           named[pre + "attention.v_proj.weight"]], 0).half()
        assert torch.equal(tw.layers[l]["wqkvT"], ref_t.t())
        assert torch.equal(tw.layers[l]["w2T"], named[pre + "feed_forward.output_dense.weight"].half().t())
    assert torch.equal(eng.w.fp_w, named["feature_projection.projection.weight"].half())
    assert torch.equal(tw.fp_wT, eng.w.fp_w.t())
