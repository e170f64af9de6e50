"""GPU parity of the backward-pass entry points against torch autograd (fp32/fp64) on the same inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from w2v2_speaker_b200 import ops as _ops
    return _ops


def rel(a, b):
    a = a.double(); b = b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("M,N,K", [(9536, 768, 768), (9536, 3072, 768), (9536, 768, 3072), (9536, 2304, 768),
                                   (64, 5994, 768), (300, 512, 768), (149 * 3, 768, 512), (70, 128, 2304)])
def test_gemm_wgrad(ops, M, N, K):
    ldy = (N + 63) // 64 * 64                      # row pitch must be a multiple of 16 bytes (TMA)
    dy = torch.zeros(M, ldy, dtype=torch.float16, device="cuda")[:, :N]
    dy.copy_(_rand((M, N), 1, 0.5))
    x = _rand((M, K), 2).half()
    dw = torch.zeros(N, K, device="cuda")
    ops.gemm_wgrad_f16(dy, x, dw)
    torch.cuda.synchronize()
    ref = dy.double().t() @ x.double()
    assert torch.isfinite(dw).all()
    assert rel(dw, ref) < 2e-5
    ops.gemm_wgrad_f16(dy, x, dw)                 # accumulates like .grad
    assert rel(dw, 2 * ref) < 2e-5


def test_dgrad_via_transposed_weight(ops):
    M, N, K = 1000, 768, 3072                      # y = x W^T, W [N, K];  dx = dy W
    dy = _rand((M, N), 3).half()
    W = _rand((N, K), 4, 0.02)
    wt = ops.cast_f16_transpose(W)                  # [K, N]
    assert wt.shape == (K, 768)
    dx = ops.gemm_f16(dy, wt, None, 0, torch.float32)
    ref = dy.double() @ W.half().double()
    assert rel(dx, ref) < 2e-5
    # padded / scaled variant (classifier: R=5994 -> 6016, rows scaled)
    W2 = _rand((5994, 768), 5, 0.02)
    sc = torch.rand(5994, generator=torch.Generator().manual_seed(6)).cuda() + 0.5
    wt2 = ops.cast_f16_transpose(W2, 6016, sc)
    assert wt2.shape == (768, 6016)
    assert rel(wt2[:, :5994].float(), (W2 * sc[:, None]).t()) < 4e-4
    assert (wt2[:, 5994:] == 0).all()


@pytest.mark.parametrize("rows,H,use_b", [(9536, 768, True), (300, 512, False), (77, 1024, True)])
def test_layernorm_bwd(ops, rows, H, use_b):
    xa = _rand((rows, H), 7, 2.0)
    bias = _rand((H,), 8)
    res = _rand((rows, H), 9)
    gamma = (1 + 0.1 * _rand((H,), 10)).requires_grad_(True)
    beta = (0.1 * _rand((H,), 11)).requires_grad_(True)
    dy_a = _rand((rows, H), 12)
    dy_b = _rand((rows, H), 13) if use_b else None
    x = (xa + bias + res).double().requires_grad_(True)
    y = F.layer_norm(x, (H,), gamma.double(), beta.double(), 1e-5)
    dy = (dy_a + (dy_b if use_b else 0)).double()
    gx, gg, gb = torch.autograd.grad(y, [x, gamma, beta], dy)
    dgamma = torch.zeros(H, device="cuda"); dbeta = torch.zeros(H, device="cuda")
    dx32, dx16 = ops.layernorm_bwd(dy_a, xa, gamma.detach(), 1e-5, dy_b=dy_b, bias=bias, residual=res, dgamma=dgamma,
                                   dbeta=dbeta)
    torch.cuda.synchronize()
    assert rel(dx32, gx) < 2e-5
    assert rel(dx16.float(), gx) < 5e-4
    assert rel(dgamma, gg) < 2e-5
    assert rel(dbeta, gb) < 2e-5


def test_gelu_bwd_colsum_ce_pool(ops):
    z = _rand((9536, 3072), 14, 1.5).half()
    dg = _rand((9536, 3072), 15).half()
    dz = ops.gelu_bwd(dg, z)
    zz = z.float().requires_grad_(True)
    ref = torch.autograd.grad(F.gelu(zz), zz, dg.float())[0]
    assert rel(dz.float(), ref) < 5e-4
    out = torch.ones(3072, device="cuda")
    ops.colsum(dz, out, 0.5)
    assert rel(out, 1 + 0.5 * dz.double().sum(0)) < 1e-5
    x32 = _rand((777, 768), 16)
    out2 = torch.zeros(768, device="cuda")
    ops.colsum(x32, out2)
    assert rel(out2, x32.double().sum(0)) < 1e-5
    # CE backward
    B, S = 64, 5994
    logits = _rand((B, S), 17, 2.0).requires_grad_(True)
    labels = torch.randint(0, S, (B,), generator=torch.Generator().manual_seed(18)).cuda()
    loss = F.cross_entropy(logits, labels)
    g = torch.autograd.grad(loss, logits)[0]
    prob = torch.softmax(logits.detach(), 1)
    dl = ops.softmax_ce_bwd(prob, labels, 1024.0 / B, 6016)
    assert dl.shape == (B, 6016) and (dl[:, S:] == 0).all()
    assert rel(dl[:, :S].float() / 1024.0, g) < 6e-4
    # mean pool backward
    demb = _rand((5, 768), 19)
    dh = ops.mean_pool_bwd(demb, 149)
    assert rel(dh, (demb / 149)[:, None, :].expand(5, 149, 768)) < 1e-6


@pytest.mark.parametrize("B,T,H,heads", [(2, 49, 768, 12), (3, 149, 768, 12), (2, 128, 768, 12), (1, 16, 768, 12),
                                         (2, 192, 1024, 16), (1, 1, 768, 12), (2, 249, 1024, 16), (1, 256, 768, 12),
                                         (1, 200, 768, 12), (2, 64, 768, 12), (1, 130, 768, 12)])
def test_attention_bwd(ops, B, T, H, heads):
    d = H // heads
    qkv = _rand((B * T, 3 * H), 20).half()
    qkv[:, :H] *= 0.35
    d_o = _rand((B * T, H), 21, 0.7).half()
    out, lse = ops.attention(qkv, B, T, H, heads, want_lse=True)
    dqkv = ops.attention_bwd(qkv, out, d_o, lse, B, T, H, heads)
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    q, k, v = (x[:, i * H:(i + 1) * H].view(B, T, heads, d).transpose(1, 2) for i in range(3))
    a = torch.softmax(q @ k.transpose(2, 3), -1)
    o = (a @ v).transpose(1, 2).reshape(B * T, H)
    ref = torch.autograd.grad(o, x, d_o.float())[0]
    assert torch.isfinite(dqkv.float()).all()
    ref_lse = torch.logsumexp(q.detach() @ k.detach().transpose(2, 3), -1)
    assert rel(lse, ref_lse) < 1e-3
    for i, name in enumerate("qkv"):
        got, want = dqkv[:, i * H:(i + 1) * H].float(), ref[:, i * H:(i + 1) * H]
        if want.abs().max().item() < 1e-6:          # T == 1: softmax over one key has zero q / k gradient
            assert got.abs().max().item() < 1e-3, name
        else:
            assert rel(got, want) < 4e-3, name


@pytest.mark.parametrize("B,T,H,heads", [(3, 149, 768, 12), (2, 249, 1024, 16)])
def test_attention_bwd_scaled_dq_and_bias_gradients(ops, B, T, H, heads):
    """w2v2_attention_bwd_ex2: dq multiplied by qscale on the way out, q/k/v bias gradients = column sums of dqkv."""
    qkv = _rand((B * T, 3 * H), 40).half()
    qkv[:, :H] *= 0.35
    d_o = _rand((B * T, H), 41, 0.7).half()
    out, lse = ops.attention(qkv, B, T, H, heads, want_lse=True)
    base = ops.attention_bwd(qkv, out, d_o, lse, B, T, H, heads).float()
    dbias = torch.zeros(3 * H, device="cuda")
    got = ops.attention_bwd(qkv, out, d_o, lse, B, T, H, heads, qscale=0.125, dbias=dbias).float()
    torch.cuda.synchronize()
    want = base.clone()
    want[:, :H] *= 0.125
    assert rel(got, want) < 1e-3
    ref_b = want.double().sum(0)
    ref_b[H:2 * H] = 0           # the key-bias gradient is exactly 0 in exact arithmetic: compare it against the others' scale
    assert (dbias.double()[:H] - ref_b[:H]).norm() / ref_b[:H].norm() < 2e-3
    assert (dbias.double()[2 * H:] - ref_b[2 * H:]).norm() / ref_b[2 * H:].norm() < 2e-3
    assert dbias[H:2 * H].double().norm() < 1e-2 * ref_b[2 * H:].norm()


@pytest.mark.parametrize("rows,H,p", [(9536, 768, 0.1), (1000, 1024, 0.0), (333, 512, 0.25)])
def test_layernorm_bwd_from_output_equals_bwd_from_inputs(ops, rows, H, p):
    """The training layers run the LayerNorm backward from the LayerNorm OUTPUT + saved rstd (one fp32 stream less):
    it must agree with the backward that recomputes the statistics from the inputs."""
    x = _rand((rows, H), 60); res = _rand((rows, H), 61); bias = _rand((H,), 62)
    gamma = 1 + 0.2 * _rand((H,), 63); beta = 0.3 * _rand((H,), 64)
    dy_a = _rand((rows, H), 65); dy_b = _rand((rows, H), 66)
    seed = 987654321
    y32, y16, rstd = ops.layernorm_with_rstd(x, gamma, beta, 1e-5, bias=bias, residual=res, drop_p=p, drop_seed=seed)
    r32, r16 = ops.layernorm(x, gamma, beta, 1e-5, bias=bias, residual=res, drop_p=p, drop_seed=seed)
    assert torch.equal(y32, r32) and torch.equal(y16, r16)
    g1, b1, d1 = (torch.zeros(H, device="cuda") for _ in range(3))
    g2, b2, d2 = (torch.zeros(H, device="cuda") for _ in range(3))
    a32, a16 = ops.layernorm_bwd(dy_a, x, gamma, 1e-5, dy_b=dy_b, bias=bias, residual=res, dgamma=g1, dbeta=b1, dbias=d1,
                                 drop_p=p, drop_seed=seed)
    o32, o16 = ops.layernorm_bwd_from_output(dy_a, y32, rstd, gamma, beta, dy_b=dy_b, dgamma=g2, dbeta=b2, dbias=d2,
                                             drop_p=p, drop_seed=seed)
    torch.cuda.synchronize()
    assert rel(o32, a32) < 2e-5
    assert rel(o16.float(), a16.float()) < 1e-3
    assert torch.equal(o16 != 0, a16 != 0) or ((o16 != 0) ^ (a16 != 0)).float().mean().item() < 1e-3
    assert rel(g2, g1) < 2e-5 and rel(b2, b1) < 1e-6 and rel(d2, d1) < 2e-5


def test_adam_matches_torch(ops):
    n = 100003
    p0 = _rand((n,), 22)
    g = _rand((n,), 23, 0.1)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for step in range(1, 4):
        ref.grad = g * step
        opt.step()
        ops.adam_step(p, g * step * 128.0, m, v, 1e-3, 0.9, 0.999, 1e-8, step, grad_scale=1.0 / 128.0)
    torch.cuda.synchronize()
    assert rel(p, ref.detach()) < 1e-6


@pytest.mark.parametrize("B,T,H,G", [(3, 149, 768, 16), (2, 49, 768, 16), (2, 249, 1024, 16), (1, 1, 768, 16),
                                     (150, 149, 768, 16)])
def test_posconv_wgrad_matches_conv1d_autograd(ops, B, T, H, G):
    """dW of the grouped k=128 conv (HF:360-368): shifted-slab tensor-core kernel vs torch's conv1d backward."""
    K, I = 128, H // G
    x16 = _rand((B, T, H), 11).half()
    dz16 = _rand((B, T, H), 12, 0.25).half()
    dw = torch.zeros(H, K * I, device="cuda")
    ops.posconv_wgrad(dz16, x16, G, K, dw)
    torch.cuda.synchronize()
    w = torch.zeros(H, I, K, dtype=torch.float64, requires_grad=True)
    y = F.conv1d(x16.double().cpu().transpose(1, 2), w, None, padding=K // 2, groups=G)[:, :, :T]
    y.backward(dz16.double().cpu().transpose(1, 2))
    ref = w.grad.permute(0, 2, 1).reshape(H, K * I).cuda()          # [o][k][i]
    assert torch.isfinite(dw).all()
    assert rel(dw, ref) < 2e-5
    ops.posconv_wgrad(dz16, x16, G, K, dw)                          # accumulates like .grad
    assert rel(dw, 2 * ref) < 2e-5


def test_fused_bias_gradients(ops):
    """Column sums fused into their producers: gelu_bwd(dbias=...) and layernorm_bwd(dbias=...) must equal the
    standalone passes."""
    M, FF, H = 1003, 3072, 768
    dg = _rand((M, FF), 21, 0.3).half()
    z = _rand((M, FF), 22, 1.5).half()
    db = torch.zeros(FF, device="cuda")
    dz = ops.gelu_bwd(dg, z, dbias=db)
    dz_ref = ops.gelu_bwd(dg, z)
    assert torch.equal(dz, dz_ref)
    assert rel(db, dz_ref.double().sum(0)) < 1e-5
    dz = ops.gelu_bwd(dg, z, dbias=db)                      # accumulates
    assert rel(db, 2 * dz_ref.double().sum(0)) < 1e-5
    x = _rand((M, H), 23)
    res = _rand((M, H), 24)
    dy = _rand((M, H), 25)
    gamma = _rand((H,), 26).abs() + 0.5
    bias = _rand((H,), 27)
    dbias = torch.zeros(H, device="cuda")
    dx32, dx16 = ops.layernorm_bwd(dy, x, gamma, 1e-5, bias=bias, residual=res, dbias=dbias)
    assert rel(dbias, dx32.double().sum(0)) < 1e-5
    dbias.zero_()
    dx32, dx16 = ops.layernorm_bwd(dy, x, gamma, 1e-5, bias=bias, residual=res, dbias=dbias, drop_p=0.1, drop_seed=77)
    assert rel(dbias, dx16.double().sum(0)) < 2e-3         # dx16 is the fp16-rounded branch gradient
