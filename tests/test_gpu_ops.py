"""GPU parity tests of the individual C-ABI entry points against plain fp32 PyTorch ops.
Run on a B200 with:  python -m pytest tests -m gpu"""
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


@pytest.mark.parametrize("M,N,K,act,f32out,use_bias", [
    (128, 256, 64, 0, False, False),
    (128, 256, 512, 0, True, False),
    (300, 768, 768, 0, True, True),
    (1000, 128, 2304, 0, True, True),
    (9536, 3072, 768, 1, False, True),
    (9536, 768, 3072, 0, True, True),
    (64, 5994, 768, 0, True, True),
    (149, 512, 1024, 1, False, False),
])
def test_gemm_f16(ops, M, N, K, act, f32out, use_bias):
    a = _rand((M, K), 1).half()
    w = _rand((N, K), 2, 1.0 / math.sqrt(K)).half()
    bias = _rand((N,), 3) if use_bias else None
    out = ops.gemm_f16(a, w, bias, act, torch.float32 if f32out else torch.float16)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if use_bias:
        ref = ref + bias
    if act:
        ref = F.gelu(ref)
    assert out.shape == (M, N)
    assert torch.isfinite(out.float()).all()
    tol = 2e-5 if f32out else 6e-4
    assert rel(out.float(), ref) < tol
    assert (out.float() - ref).abs().max().item() < (1e-3 if f32out else 2e-2)


@pytest.mark.parametrize("B,L,C,k", [(3, 399, 512, 3), (2, 1199, 512, 3), (4, 299, 512, 2), (1, 3, 512, 3), (5, 98, 512, 2)])
def test_conv1d_channels_last(ops, B, L, C, k):
    x = _rand((B, L, C), 4).half()
    w = _rand((C, C, k), 5, math.sqrt(2.0 / (C * k)))
    wt = ops.conv_weight_tapmajor(w)
    out = ops.conv1d_cl_f16(x, wt, k, 2, act=1)
    torch.cuda.synchronize()
    ref = F.gelu(F.conv1d(x.float().transpose(1, 2), w.half().float(), stride=2)).transpose(1, 2)
    assert out.shape == ref.shape
    assert rel(out.float(), ref) < 6e-4


@pytest.mark.parametrize("M,N,K,use_bias", [
    (9536, 768, 768, False),      # out_proj: 225 tiles on 148 SMs -> stream-K
    (9536, 768, 3072, True),      # FFN2 shape, bias added by the leading segment only
    (9536, 768, 2304, False),     # QKV data gradient
    (5000, 1000, 192, True),      # 160 tiles, 3 k-blocks each: ranges barely longer than a tile
    (7968, 1024, 4096, False),    # wav2vec2-large, 5 s
])
def test_gemm_f16_stream_k_repeatable(ops, M, N, K, use_bias):
    """fp32-output GEMMs whose tile count leaves the last wave mostly idle run the stream-K schedule
    (two CTAs share a tile through a flag hand-shake + TMA reduce-add): results must match the dense
    reference and stay correct over repeated launches (the flags re-arm themselves)."""
    a = _rand((M, K), 11).half()
    w = _rand((N, K), 12, 1.0 / math.sqrt(K)).half()
    bias = _rand((N,), 13) if use_bias else None
    ref = a.float() @ w.float().t()
    if use_bias:
        ref = ref + bias
    for _ in range(12):              # more launches than flag-ring slots
        out = ops.gemm_f16(a, w, bias, 0, torch.float32)
        torch.cuda.synchronize()
        assert rel(out, ref) < 2e-5
        assert (out - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(9536, 3072, 768), (1000, 4096, 1024), (300, 256, 64)])
def test_gemm_dual_output_gelu(ops, M, N, K):
    """FFN1 of the training forward: pre-activation and GELU output from one GEMM, identical to the two-pass form."""
    a = _rand((M, K), 51).half()
    w = _rand((N, K), 52, 1.0 / math.sqrt(K)).half()
    bias = _rand((N,), 53)
    act, pre = ops.gemm_f16_dual_gelu(a, w, bias)
    torch.cuda.synchronize()
    ref_pre = a.float() @ w.float().t() + bias
    assert rel(pre.float(), ref_pre) < 6e-4
    assert rel(act.float(), F.gelu(ref_pre)) < 6e-4
    z = ops.gemm_f16(a, w, bias, 0, torch.float16)
    assert torch.equal(pre, z)                            # same accumulation, same rounding
    g, _ = ops.gelu_fwd(z.contiguous(), torch.float16)
    # the fused epilogue applies the GELU to the fp32 accumulator, the two-pass form to the fp16-rounded z
    assert rel(act.float(), g.float()) < 6e-4


@pytest.mark.parametrize("M,N,K", [(9536, 3072, 768), (1000, 4096, 1024), (300, 384, 64), (77, 160, 128)])
def test_gemm_gelu_backward_epilogue(ops, M, N, K):
    """FFN2 data gradient of the training backward: dz = (a W^T) * gelu'(z) and the column sums of dz from one GEMM,
    against autograd through F.gelu in fp32 and against the two-pass form (GEMM, then w2v2_gelu_bwd_colsum)."""
    a = _rand((M, K), 61).half()
    w = _rand((N, K), 62, 1.0 / math.sqrt(K)).half()
    z = _rand((M, N), 63, 1.5).half()
    dbias = torch.full((N,), 0.25, device="cuda")          # accumulated into, not overwritten
    for _ in range(2):                                    # twice: the shared-memory accumulators re-arm themselves
        dz = ops.gemm_f16_gelu_bwd(a, w, z, dbias)
    torch.cuda.synchronize()
    zf = z.float().requires_grad_(True)
    dg = a.float() @ w.float().t()
    F.gelu(zf).backward(dg)
    assert rel(dz.float(), zf.grad) < 6e-4
    ref_sum = 0.25 + 2 * dz.double().sum(0)                # the sums of the fp16 values the wgrad GEMM reads
    scale = dz.double().abs().sum(0) + 1.0
    assert ((dbias.double() - ref_sum).abs() / scale).max().item() < 2e-6
    dg16 = ops.gemm_f16(a, w, None, 0, torch.float16)
    db2 = torch.zeros(N, device="cuda")
    dz2 = ops.gelu_bwd(dg16, z, db2)
    assert rel(dz.float(), dz2.float()) < 8e-4           # two-pass rounds dg to fp16 before the multiply
    assert ((2 * db2.double() + 0.25 - dbias.double()).abs() / scale).max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(9536, 3072, 768), (300, 384, 64)])
def test_gemm_pair_keeping_the_gelu_derivative(ops, M, N, K):
    """FFN1 forward keeping gelu'(z) (w2v2_gemm_f16_dual_gelu_grad) and the multiply-only backward epilogue
    (w2v2_gemm_f16_mul_colsum) against torch fp32 and against the validated z-keeping pair."""
    a = _rand((M, K), 71).half()
    w = _rand((N, K), 72, 1.0 / math.sqrt(K)).half()
    bias = _rand((N,), 73)
    act, grad = ops.gemm_f16_dual_gelu_grad(a, w, bias)
    act_ref, pre = ops.gemm_f16_dual_gelu(a, w, bias)
    torch.cuda.synchronize()
    assert rel(act.float(), act_ref.float()) < 1e-6                   # same accumulator, same GELU polynomial
    zf = (a.float() @ w.float().t() + bias).requires_grad_(True)
    F.gelu(zf).sum().backward()
    assert rel(grad.float(), zf.grad) < 6e-4
    dy = _rand((M, K), 74).half()
    w2t = _rand((N, K), 75, 1.0 / math.sqrt(K)).half()
    db_a = torch.zeros(N, device="cuda"); db_b = torch.zeros(N, device="cuda")
    dz_a = ops.gemm_f16_mul_colsum(dy, w2t, grad, db_a)
    dz_b = ops.gemm_f16_gelu_bwd(dy, w2t, pre, db_b)
    torch.cuda.synchronize()
    assert rel(dz_a.float(), dz_b.float()) < 1.5e-3                   # gelu' rounded to fp16 vs recomputed from fp16 z
    scale = dz_b.double().abs().sum(0) + 1.0
    assert ((db_a.double() - dz_a.double().sum(0)).abs() / scale).max().item() < 2e-6
    assert ((db_a.double() - db_b.double()).abs() / scale).max().item() < 2e-3


def test_gemm_rejects_bad_k(ops):
    from w2v2_speaker_b200._lib import W2V2Error
    a = torch.zeros(8, 40, dtype=torch.float16, device="cuda")
    w = torch.zeros(8, 40, dtype=torch.float16, device="cuda")
    with pytest.raises(W2V2Error):
        ops.gemm_f16(a, w)


@pytest.mark.parametrize("B,N", [(2, 16000), (3, 11283), (1, 400), (2, 48000)])
def test_conv0_groupnorm_gelu(ops, B, N):
    C = 512
    wav = _rand((B, N), 6)
    wav = wav + 0.3            # non-zero mean exercises the moment-based statistics
    w = _rand((C, 1, 10), 7, math.sqrt(2.0 / 10))
    gamma = 1 + 0.1 * _rand((C,), 8)
    beta = 0.1 * _rand((C,), 9)
    out = ops.conv0_gn_gelu(wav, w.view(C, 10), gamma, beta, 1e-5)
    torch.cuda.synchronize()
    h = F.conv1d(wav[:, None, :].double(), w.double(), stride=5)
    ref = F.gelu(F.group_norm(h, C, gamma.double(), beta.double(), 1e-5)).transpose(1, 2).float()
    assert out.shape == ref.shape
    assert rel(out.float(), ref) < 4e-4          # fp16 output rounding ~ 2^-11/sqrt(3)
    assert (out.float() - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("rows,H,xf32,use_res", [(149 * 2, 768, True, True), (1000, 512, True, False), (77, 1024, False, True), (5, 768, False, False)])
def test_layernorm(ops, rows, H, xf32, use_res):
    x = _rand((rows, H), 10, 2.0) + 0.5
    xin = x if xf32 else x.half()
    bias = _rand((H,), 11)
    res = _rand((rows, H), 12) if use_res else None
    gamma = 1 + 0.1 * _rand((H,), 13)
    beta = 0.1 * _rand((H,), 14)
    y32, y16 = ops.layernorm(xin, gamma, beta, 1e-5, bias=bias, residual=res)
    torch.cuda.synchronize()
    t = xin.float() + bias + (res if use_res else 0)
    ref = F.layer_norm(t.double(), (H,), gamma.double(), beta.double(), 1e-5).float()
    assert rel(y32, ref) < 1e-6
    assert rel(y16.float(), ref) < 4e-4


@pytest.mark.parametrize("B,T,H", [(2, 49, 768), (3, 149, 768), (1, 2, 1024)])
def test_stat_pool(ops, B, T, H):
    x = _rand((B, T, H), 15) + 0.7
    assert rel(ops.stat_pool(x, 0), x.mean(1)) < 1e-6
    assert rel(ops.stat_pool(x, 1), torch.cat(torch.std_mean(x, 1), 1)) < 2e-6
    assert torch.equal(ops.stat_pool(x, 2), x.max(1).values)


def test_asp_pieces(ops):
    from oracle.params import make_asp_params
    from oracle import w2v2_oracle as O
    B, T, H = 3, 49, 768
    x = _rand((B, T, H), 16)
    asp = {k: v.cuda() for k, v in make_asp_params(H, seed=2).items()}
    # front
    cat = ops.asp_concat(x)
    mean = x.mean(1)
    std = ((x - mean[:, None]) ** 2).mean(1).clamp(1e-12).sqrt()
    refcat = torch.cat([x, mean[:, None].expand(B, T, H), std[:, None].expand(B, T, H)], 2).reshape(B * T, 3 * H)
    assert rel(cat.float(), refcat) < 4e-4
    # tail
    lg = _rand((B, T, H), 17, 2.0)
    out = ops.asp_pool(x, lg)
    a = torch.softmax(lg, 1)
    m = (a * x).sum(1)
    s = (a * (x - m[:, None]) ** 2).sum(1).clamp(1e-12).sqrt()
    assert rel(out, torch.cat([m, s], 1)) < 2e-6
    # middle
    z = _rand((B * T, 128), 18)
    scale = asp["tdnn.norm.norm.weight"] / torch.sqrt(asp["tdnn.norm.norm.running_var"] + 1e-5)
    shift = asp["tdnn.norm.norm.bias"] - asp["tdnn.norm.norm.running_mean"] * scale
    y = ops.asp_relu_bn_tanh(z, scale, shift)
    assert rel(y.float(), torch.tanh(torch.relu(z) * scale + shift)) < 4e-4


def test_softmax_ce_and_aam(ops):
    from oracle import w2v2_oracle as O
    B, S, E = 7, 5994, 1536
    logits = _rand((B, S), 19, 3.0)
    labels = torch.randint(0, S, (B,), generator=torch.Generator().manual_seed(20)).cuda()
    prob, loss, am = ops.softmax_ce(logits, labels)
    assert rel(prob, torch.softmax(logits, 1)) < 1e-6
    assert rel(loss, F.cross_entropy(logits, labels, reduction="none")) < 1e-6
    assert torch.equal(am.long(), logits.argmax(1))
    assert abs(ops.mean_rows(loss).item() - F.cross_entropy(logits, labels).item()) < 1e-5
    # AAM on exact cosines (the GEMM part is covered by the split3 test below)
    x = _rand((B, E), 21); W = _rand((S, E), 22)
    cos = F.linear(F.normalize(x), F.normalize(W))
    # force both branches of the margin at the label column
    cos[0, labels[0]] = -0.995
    cos[1, labels[1]] = 0.9
    ref_logits, ref_loss, ref_sm = None, None, None
    one_hot = torch.zeros_like(cos).scatter_(1, labels.view(-1, 1), 1)
    m, s = 0.2, 30.0
    sine = torch.sqrt((1 - cos * cos).clamp(0, 1))
    phi = cos * math.cos(m) - sine * math.sin(m)
    phi = torch.where((cos - math.cos(math.pi - m)) > 0, phi, cos - math.sin(math.pi - m) * m)
    ref_logits = (one_hot * phi + (1 - one_hot) * cos) * s
    c2 = cos.clone()
    prob, loss, am = ops.aam_softmax_ce(c2, labels, m, s)
    assert rel(c2, ref_logits) < 1e-6
    assert rel(prob, torch.softmax(ref_logits, 1)) < 2e-6
    assert rel(loss, F.cross_entropy(ref_logits, labels, reduction="none")) < 2e-6
    assert torch.equal(am.long(), ref_logits.argmax(1))


def test_split3_classifier_gemm_is_fp32_accurate(ops):
    B, S, E = 64, 5994, 1536
    x = _rand((B, E), 23); W = _rand((S, E), 24, math.sqrt(2.0 / (S + E)))
    xa = ops.l2norm_rows_split3(x, 0)
    wb = ops.l2norm_rows_split3(W, 1)
    cos = ops.gemm_f16(xa, wb, None, 0, torch.float32)
    ref = F.linear(F.normalize(x.double()), F.normalize(W.double()))
    assert (cos.double() - ref).abs().max().item() < 2e-6
    # plain (non-normalised) split
    xa = ops.split3_rows(x, 0); wb = ops.split3_rows(W, 1)
    lg = ops.gemm_f16(xa, wb, None, 0, torch.float32)
    ref = x.double() @ W.double().t()
    assert rel(lg.double(), ref) < 2e-5      # lo parts of ~0.016-scale weights are fp16 subnormals


@pytest.mark.parametrize("B,T,H,heads", [(2, 49, 768, 12), (3, 149, 768, 12), (1, 249, 1024, 16), (2, 128, 768, 12), (1, 1, 768, 12), (2, 16, 768, 12)])
def test_attention(ops, B, T, H, heads):
    qkv = _rand((B * T, 3 * H), 30, 1.0).half()
    qkv[:, :H] *= 0.35          # q already carries the d^-0.5 scale on this path
    out = ops.attention(qkv, B, T, H, heads)
    torch.cuda.synchronize()
    d = H // heads
    q, k, v = (qkv[:, i * H:(i + 1) * H].float().view(B, T, heads, d).transpose(1, 2) for i in range(3))
    a = torch.softmax(q @ k.transpose(2, 3), -1)
    ref = (a @ v).transpose(1, 2).reshape(B * T, H)
    assert torch.isfinite(out.float()).all()
    assert rel(out.float(), ref) < 1.5e-3


@pytest.mark.parametrize("B,T,H,heads", [(2, 300, 768, 12), (1, 513, 768, 12), (1, 1000, 1024, 16), (3, 257, 768, 12),
                                         (1, 384, 768, 12)])
def test_attention_long_sequences(ops, B, T, H, heads):
    """T > 256 (full-utterance evaluation): the key-tiled two-pass kernel, output and log-sum-exp."""
    qkv = _rand((B * T, 3 * H), 31, 1.0).half()
    qkv[:, :H] *= 0.35
    out, lse = ops.attention(qkv, B, T, H, heads, want_lse=True)
    torch.cuda.synchronize()
    d = H // heads
    q, k, v = (qkv[:, i * H:(i + 1) * H].float().view(B, T, heads, d).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(2, 3)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, H)
    assert torch.isfinite(out.float()).all()
    assert rel(out.float(), ref) < 1.5e-3
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-3


def test_attention_dropout_on_long_sequences_keeps_the_expected_share(ops):
    """Attention dropout beyond 256 frames (key-tiled kernel; mask parity: tests/test_gpu_attention_persist.py): with
    V = 1 every output element is the kept probability mass of its row / (1 - p): mean 1, never above 1 / (1 - p)."""
    B, T, H, heads, p = 2, 300, 768, 12, 0.1
    qkv = torch.zeros(B * T, 3 * H, dtype=torch.float16, device="cuda")
    qkv[:, 2 * H:] = 1.0
    out = ops.attention(qkv, B, T, H, heads, drop_p=p, drop_seed=1)
    out = (out[0] if isinstance(out, tuple) else out).float()
    assert abs(out.mean().item() - 1.0) < 5e-3 and out.max().item() <= 1.0 / (1.0 - p) + 1e-2
    assert out.std().item() > 5e-3                                   # rows differ: a mask was applied


@pytest.mark.parametrize("B,T,H,G", [(2, 49, 768, 16), (5, 149, 768, 16), (3, 249, 1024, 16), (1, 7, 768, 16), (9, 35, 768, 16)])
def test_posconv(ops, B, T, H, G):
    K = 128
    x = _rand((B, T, H), 31).half()
    v = _rand((H, H // G, K), 32, 2.0 * math.sqrt(1.0 / (K * H)))
    g = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt() * (1 + 0.05 * _rand((1, 1, K), 33))
    bias = _rand((H,), 34, 0.02)
    w16 = ops.posconv_fold_weight(v, g.view(-1), G, ops.posconv_taps_per_mma(T, H, G))
    out = ops.posconv(x, w16, bias, G, K)
    torch.cuda.synchronize()
    w = (g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()).half().float()
    y = F.conv1d(x.float().transpose(1, 2), w, bias, padding=K // 2, groups=G)[:, :, :-1]
    ref = F.gelu(y).transpose(1, 2)
    assert out.shape == ref.shape
    assert torch.isfinite(out).all()
    # the tap norms are summed in a different (but fixed) order than torch's: a handful of folded weights land on
    # the other side of an fp16 rounding boundary (one ulp = 5e-4 of ONE weight), hence a little more than 2e-5
    assert rel(out, ref) < 4e-5
