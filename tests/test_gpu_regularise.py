"""Train-mode regularisation kernels (dropout masks as a pure function of (seed, index), attention
dropout inside the tcgen05 attention forward/backward, SpecAugment time mask) against torch, using a
numpy replica of the counter-based mask; plus an end-to-end step with the reference's default
regularisation (R:src/models/wav2vec2.py:83-94)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
S = 5994


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from w2v2_speaker_b200 import ops as _ops
    return _ops


def keep_mask(seed: int, n: int, p: float) -> np.ndarray:
    """numpy replica of csrc/common.cuh::dropout_hash + the 16-bit threshold test (n even)."""
    thr = int(p * 65536.0 + 0.5)
    M32 = 0xFFFFFFFF

    def fmix32(x):          # python ints
        x ^= x >> 16
        x = (x * 0x85EBCA6B) & M32
        x ^= x >> 13
        x = (x * 0xC2B2AE35) & M32
        return x ^ (x >> 16)

    k0 = fmix32(((seed & M32) + 0x9E3779B9) & M32)
    k1 = fmix32(((seed >> 32) & M32) ^ k0 ^ 0x7F4A7C15)
    m32 = np.uint64(M32)
    idx = np.arange(n // 2, dtype=np.uint64)
    x = (((idx & m32) ^ np.uint64(k0)) * np.uint64(0x9E3779B1)) & m32
    x ^= x >> np.uint64(15)
    x = (x + np.uint64(k1) + (idx >> np.uint64(32)) * np.uint64(0x7FEB352D)) & m32
    x = (x * np.uint64(0x85EBCA6B)) & m32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & m32
    x ^= x >> np.uint64(16)
    bits = x
    lo, hi = bits & np.uint64(0xFFFF), bits >> np.uint64(16)
    keep = np.empty(n, dtype=bool)
    keep[0::2] = lo >= thr
    keep[1::2] = hi >= thr
    return keep, 1.0 / (1.0 - thr / 65536.0)


def rel(a, b):
    a = a.double(); b = b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_dropout_kernel_matches_replica(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 768, generator=g).cuda()
    bias = torch.randn(768, generator=g).cuda()
    seed, p = 123456789012345, 0.1
    keep, inv = keep_mask(seed, x.numel(), p)
    keep_t = torch.from_numpy(keep).view(300, 768).cuda()
    ref = torch.where(keep_t, (x + bias) * inv, torch.zeros_like(x))
    y, y16 = ops.dropout_(x.clone(), p, seed, bias=bias, want16=True)
    assert rel(y, ref) < 1e-6
    assert rel(y16.float(), ref) < 5e-4
    assert abs(keep.mean() - 0.9) < 5e-3
    xh = x.half()
    yh, _ = ops.dropout_(xh.clone(), 0.25, 7)
    k2, inv2 = keep_mask(7, x.numel(), 0.25)
    ref2 = torch.where(torch.from_numpy(k2).view(300, 768).cuda(), xh.float() * inv2, torch.zeros_like(x))
    assert rel(yh.float(), ref2) < 5e-4
    # p = 0 is the identity
    y0, _ = ops.dropout_(x.clone(), 0.0, 99)
    assert torch.equal(y0, x)


@pytest.mark.parametrize("B,T,H,heads", [(2, 49, 768, 12), (2, 149, 768, 12), (1, 249, 1024, 16)])
def test_attention_dropout_forward_backward(ops, B, T, H, heads):
    d = H // heads
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * T, 3 * H, generator=g).cuda().half()
    qkv[:, :H] *= 0.35
    d_o = (torch.randn(B * T, H, generator=g) * 0.7).cuda().half()
    p, seed = 0.1, 424242
    out, lse = ops.attention(qkv, B, T, H, heads, want_lse=True, drop_p=p, drop_seed=seed)
    dqkv = ops.attention_bwd(qkv, out, d_o, lse, B, T, H, heads, drop_p=p, drop_seed=seed)
    torch.cuda.synchronize()
    TK = (T + 15) // 16 * 16
    keep, inv = keep_mask(seed, B * heads * T * TK, p)
    m = torch.from_numpy(keep).view(B, heads, T, TK)[..., :T].cuda()
    x = qkv.float().requires_grad_(True)
    q, k, v = (x[:, i * H:(i + 1) * H].view(B, T, heads, d).transpose(1, 2) for i in range(3))
    a = torch.softmax(q @ k.transpose(2, 3), -1)
    a = torch.where(m, a * inv, torch.zeros_like(a))
    o = (a @ v).transpose(1, 2).reshape(B * T, H)
    ref = torch.autograd.grad(o, x, d_o.float())[0]
    assert rel(out.float(), o.detach()) < 2e-3
    for i in range(3):
        assert rel(dqkv[:, i * H:(i + 1) * H].float(), ref[:, i * H:(i + 1) * H]) < 5e-3, "qkv"[i]


def test_time_mask_apply_and_backward(ops):
    from w2v2_speaker_b200.training import compute_time_mask
    B, T, H = 4, 149, 768
    rng = np.random.default_rng(3)
    mask = torch.from_numpy(compute_time_mask(B, T, 0.05, 10, 2, rng)).cuda()
    g = torch.Generator().manual_seed(2)
    h = torch.randn(B * T, H, generator=g).cuda()
    emb = torch.rand(H, generator=g).cuda()
    out = ops.time_mask_apply_(h.clone(), mask, emb)
    ref = torch.where(mask.bool()[:, None], emb[None, :].expand(B * T, H), h)
    assert torch.equal(out, ref)
    dh = torch.randn(B * T, H, generator=g).cuda()
    dembed = torch.zeros(H, device="cuda")
    d2 = ops.time_mask_bwd_(dh.clone(), mask, dembed, 0.5)
    assert torch.equal(d2, torch.where(mask.bool()[:, None], torch.zeros_like(dh), dh))
    assert rel(dembed, 0.5 * dh[mask.bool()].sum(0)) < 1e-5


def test_feature_axis_specaugment_kernel_and_training_step(ops, base_params):
    """SpecAugment along the feature axis (HF:1312-1322, `mask_feature_prob`; off in every reference configuration, built
    for completeness): the masking kernel against torch, then one training step with ONLY that regularisation on against
    autograd of the oracle with the same mask applied between the feature projection and the encoder."""
    import torch.nn.functional as F
    from oracle import w2v2_oracle as O
    from oracle.params import make_head_params, make_inputs
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    from w2v2_speaker_b200.training import compute_time_mask
    B, T, H = 3, 49, 768
    rng = np.random.default_rng(5)
    fm = torch.from_numpy(compute_time_mask(B, H, 0.3, 10, 0, rng)).cuda()
    assert 0.1 < fm.float().mean().item() < 0.4
    h = torch.randn(B, T, H, generator=torch.Generator().manual_seed(1)).cuda()
    out = ops.feature_mask_(h.clone(), fm, B, T)
    assert torch.equal(out, torch.where(fm.view(B, 1, H).bool(), torch.zeros_like(h), h))

    zero = dict(activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                mask_time_prob=0.0)
    m = Wav2vec2FCModule(Wav2vec2FCModuleConfig(mask_feature_prob=0.3, mask_feature_length=10, **zero), S, CrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    head = make_head_params(768, S, seed=1)
    with torch.no_grad():
        m.fc_list[-1][0].weight.copy_(head["fc.weight"]); m.fc_list[-1][0].bias.copy_(head["fc.bias"])
    m = m.cuda().train()
    m.on_train_start()
    model = m.wav2vec.model
    plans = []
    draw = model._draw_reg_plan
    model._draw_reg_plan = lambda *a, **k: (plans.append(draw(*a, **k)), plans[-1])[1]
    wav, labels = make_inputs(B, 16000, S, seed=9)
    emb, pred = m(wav[:, None, :].cuda())
    loss, _ = m.loss_fn(pred, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert len(plans) == 1 and plans[0].fmask is not None and plans[0].mask is None
    keep = 1.0 - plans[0].fmask.view(B, 1, H).float().cpu()

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    fw, fb = head["fc.weight"].clone().requires_grad_(True), head["fc.bias"].clone().requires_grad_(True)
    proj = O.feature_projection(O.feature_extractor(wav, p).transpose(1, 2), p) * keep
    ref_emb = O.encoder(proj, p).mean(1)
    ref_loss = F.cross_entropy(F.linear(ref_emb, fw, fb), labels)
    ref_loss.backward()
    assert rel(emb.detach().cpu(), ref_emb.detach()) < 1.5e-3
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-3
    got = dict(model.named_parameters())
    for k in ("feature_projection.projection.weight", "feature_projection.projection.bias", "feature_projection.layer_norm.weight",
              "encoder.layers.0.attention.q_proj.weight", "encoder.layers.11.feed_forward.output_dense.weight",
              "encoder.pos_conv_embed.conv.bias"):
        assert rel(got[k].grad.cpu(), p[k].grad) < 1e-2, k
    # the masked hidden units see no gradient through the projection: their bias gradient is exactly zero for an
    # utterance-independent check only where EVERY utterance masks the unit
    allm = plans[0].fmask.view(B, H).bool().all(0).cpu()
    if allm.any():
        assert got["feature_projection.projection.bias"].grad.cpu()[allm].abs().max().item() == 0.0


def test_training_step_with_reference_default_regularisation(base_params):
    """dropout 0.1 x3, LayerDrop 0.05, SpecAugment 0.05 (the reference defaults): the step runs, every
    gradient is finite, LayerDrop-skipped layers get exactly zero gradient, masked_spec_embed trains."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle.params import make_head_params, make_inputs
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    torch.manual_seed(0)
    cfg = Wav2vec2FCModuleConfig(layerdrop=0.3)                   # other probabilities: reference defaults
    m = Wav2vec2FCModule(cfg, S, CrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    m = m.cuda().train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    wav, labels = make_inputs(4, 48000, S, seed=11)
    emb, pred = m(wav[:, None, :].cuda())
    loss, prob = m.loss_fn(pred, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    model = m.wav2vec.model
    skipped = 0
    for l in range(12):
        g = dict(model.named_parameters())[f"encoder.layers.{l}.feed_forward.output_dense.weight"].grad
        assert torch.isfinite(g).all()
        skipped += int(g.abs().max().item() == 0.0)
    assert 1 <= skipped <= 9                                       # p = 0.3 over 12 layers
    assert model.masked_spec_embed.grad.abs().max().item() > 0     # SpecAugment rows feed the embedding
    for n, q in model.named_parameters():
        if q.grad is not None:
            assert torch.isfinite(q.grad).all(), n
    # eval mode is deterministic and unaffected
    m.eval()
    with torch.no_grad():
        e1, _ = m(wav[:, None, :].cuda())
        e2, _ = m(wav[:, None, :].cuda())
    assert torch.equal(e1, e2)


@pytest.mark.parametrize("rows,H", [(1000, 768), (37, 1024), (300, 512)])
def test_layernorm_with_fused_dropout_equals_two_pass(rows, H):
    """LayerNorm(dropout(x + bias) + residual) with the mask generated inside the kernel == the standalone
    dropout pass followed by the plain LayerNorm (bit-identical: same hash, same arithmetic), forward and backward."""
    from w2v2_speaker_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(rows, H, generator=g).cuda()
    res = torch.randn(rows, H, generator=g).cuda()
    bias = torch.randn(H, generator=g).cuda()
    gamma = (torch.rand(H, generator=g) + 0.5).cuda()
    beta = torch.randn(H, generator=g).cuda()
    dy = torch.randn(rows, H, generator=g).cuda()
    p, seed = 0.1, 0x1234567890AB
    y32, y16 = ops.layernorm(x, gamma, beta, 1e-5, bias=bias, residual=res, drop_p=p, drop_seed=seed)
    xd = x.clone()
    ops.dropout_(xd, p, seed, bias=bias)
    r32, r16 = ops.layernorm(xd, gamma, beta, 1e-5, residual=res)
    assert torch.equal(y32, r32) and torch.equal(y16, r16)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    dx32, dx16 = ops.layernorm_bwd(dy, x, gamma, 1e-5, bias=bias, residual=res, dgamma=dg, dbeta=db, drop_p=p, drop_seed=seed)
    dg2, db2 = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    e32, e16 = ops.layernorm_bwd(dy, xd, gamma, 1e-5, residual=res, dgamma=dg2, dbeta=db2)
    assert torch.equal(dx32, e32)                       # residual gradient: unmasked
    ops.dropout_(e16, p, seed)                          # branch gradient: masked and rescaled
    assert (dx16.float() - e16.float()).abs().max().item() <= 2e-3 * e16.float().abs().max().item()
    kept = (e16 != 0)
    assert torch.equal(dx16 != 0, kept) or ((dx16 != 0) ^ kept).float().mean().item() < 1e-3
    assert torch.allclose(dg, dg2, rtol=1e-4, atol=1e-4) and torch.allclose(db, db2, rtol=1e-4, atol=1e-4)
