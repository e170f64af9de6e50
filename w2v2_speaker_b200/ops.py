"""Thin torch-tensor wrappers over the C ABI (one function per entry point of include/w2v2_b200.h).

torch is used for device memory and streams only; every function launches hand-written CUDA on
torch's current stream via libw2v2_b200.so and raises if the library is missing.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

F16 = torch.float16
F32 = torch.float32


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.W2V2Error(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise _lib.W2V2Error(f"{name}: expected dtype {dtype}, got {t.dtype}")


def gemm_f16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = 0,
             out_dtype=F16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias).  a, w fp16 row-major (row pitch = stride(0))."""
    _chk(a, F16, "a"); _chk(w, F16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        ld = (N + 7) // 8 * 8                      # keep the row pitch a multiple of 16 bytes
        buf = torch.empty(M, ld, dtype=out_dtype, device=a.device)
        out = buf[:, :N]
    call("w2v2_gemm_f16", ptr(a), M, a.stride(0), 0, 1, 1, 0, K, ptr(w), w.stride(0), N, ptr(bias), act,
         ptr(out), 1 if out.dtype == F32 else 0, out.stride(0), 0, stream_ptr())
    return out


def gemm_f16_dual_gelu(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor):
    """-> (gelu(a @ w^T + bias), a @ w^T + bias), both f16 [M, N], from one GEMM."""
    _chk(a, F16, "a"); _chk(w, F16, "w")
    M, K = a.shape
    N = w.shape[0]
    act = torch.empty(M, N, dtype=F16, device=a.device)
    pre = torch.empty(M, N, dtype=F16, device=a.device)
    call("w2v2_gemm_f16_dual_gelu", ptr(a), M, a.stride(0), K, ptr(w), w.stride(0), N, ptr(bias), ptr(act), ptr(pre), N,
         stream_ptr())
    return act, pre


def gemm_f16_dual_gelu_grad(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor):
    """EXPERIMENTAL -> (gelu(a @ w^T + bias), gelu'(a @ w^T + bias)), both f16 [M, N], from one GEMM."""
    _chk(a, F16, "a"); _chk(w, F16, "w")
    M, K = a.shape
    N = w.shape[0]
    act = torch.empty(M, N, dtype=F16, device=a.device)
    grad = torch.empty(M, N, dtype=F16, device=a.device)
    call("w2v2_gemm_f16_dual_gelu_grad", ptr(a), M, a.stride(0), K, ptr(w), w.stride(0), N, ptr(bias), ptr(act), ptr(grad), N,
         stream_ptr())
    return act, grad


def gemm_f16_mul_colsum(a: torch.Tensor, w: torch.Tensor, mul: torch.Tensor, colsum: torch.Tensor) -> torch.Tensor:
    """EXPERIMENTAL -> out = (a @ w^T) * mul, f16 [M, N]; colsum[n] += sum_r out[r, n] (fp32, in place)."""
    _chk(a, F16, "a"); _chk(w, F16, "w"); _chk(mul, F16, "mul"); _chk(colsum, F32, "colsum")
    M, K = a.shape
    N = w.shape[0]
    assert mul.shape == (M, N) and colsum.numel() == N
    out = torch.empty(M, N, dtype=F16, device=a.device)
    call("w2v2_gemm_f16_mul_colsum", ptr(a), M, a.stride(0), K, ptr(w), w.stride(0), N, ptr(mul), mul.stride(0), ptr(out), N,
         ptr(colsum), stream_ptr())
    return out


def gemm_f16_gelu_bwd(a: torch.Tensor, w: torch.Tensor, z: torch.Tensor, dbias: torch.Tensor) -> torch.Tensor:
    """-> dz = (a @ w^T) * gelu'(z), f16 [M, N]; dbias[n] += sum_r dz[r, n] (fp32, accumulated in place)."""
    _chk(a, F16, "a"); _chk(w, F16, "w"); _chk(z, F16, "z"); _chk(dbias, F32, "dbias")
    M, K = a.shape
    N = w.shape[0]
    assert z.shape == (M, N) and dbias.numel() == N
    dz = torch.empty(M, N, dtype=F16, device=a.device)
    call("w2v2_gemm_f16_gelu_bwd", ptr(a), M, a.stride(0), K, ptr(w), w.stride(0), N, ptr(z), z.stride(0), ptr(dz), N,
         ptr(dbias), stream_ptr())
    return dz


def conv1d_cl_f16(x: torch.Tensor, w_tap: torch.Tensor, ksize: int, stride: int, act: int = 1,
                  out_dtype=F16) -> torch.Tensor:
    """Strided Conv1d (no bias) + activation over a channels-last fp16 activation x[B,L,C] with the
    tap-major weight w_tap[Cout, ksize*C] (HF:254-272)."""
    _chk(x, F16, "x"); _chk(w_tap, F16, "w_tap")
    B, L, C = x.shape
    assert x.is_contiguous()
    Lout = (L - ksize) // stride + 1
    Cout = w_tap.shape[0]
    out = torch.empty(B, Lout, Cout, dtype=out_dtype, device=x.device)
    call("w2v2_gemm_f16", ptr(x), Lout, stride * C, L * C, B, ksize, C, C, ptr(w_tap), w_tap.stride(0), Cout,
         None, act, ptr(out), 1 if out_dtype == F32 else 0, Cout, Lout * Cout, stream_ptr())
    return out


def conv_weight_tapmajor(w: torch.Tensor) -> torch.Tensor:
    _chk(w, F32, "w")
    cout, cin, k = w.shape
    o = torch.empty(cout, k * cin, dtype=F16, device=w.device)
    call("w2v2_conv_weight_tapmajor", ptr(w.contiguous()), ptr(o), cout, cin, k, stream_ptr())
    return o


def cast_f16(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    _chk(x, F32, "x")
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=F16, device=x.device)
    call("w2v2_cast_f16", ptr(x), ptr(y), x.numel(), float(scale), stream_ptr())
    return y


def _chk_lens(lens: torch.Tensor, B: int) -> torch.Tensor:
    if not lens.is_cuda or lens.dtype != torch.int32 or lens.numel() != B or not lens.is_contiguous():
        raise _lib.W2V2Error(f"lens must be a contiguous CUDA int32 vector of {B} entries")
    return lens


def conv0_gn_gelu(wav: torch.Tensor, w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float = 1e-5, lens: Optional[torch.Tensor] = None, normalize: bool = False) -> torch.Tensor:
    """HF:302-323.  wav [B,N] -> f16 channels-last [B, L0, C].  wav: float32, or int16 PCM (x = pcm / 32768).
    lens (int32 [B], samples): zero-padded ragged batch, the GroupNorm statistics of utterance b cover its own frames only.
    normalize: fold the reference's per-utterance input normaliser (x - mean) / (std + 1e-5) into the GroupNorm affine."""
    if wav.dtype not in (F32, torch.int16) or not wav.is_cuda:
        raise _lib.W2V2Error(f"wav: expected a CUDA float32 or int16 tensor, got {wav.dtype} on {wav.device}")
    B, N = wav.shape
    C = w.shape[0]
    L0 = (N - 10) // 5 + 1
    lib = _lib.load()
    ws = torch.empty(lib.w2v2_conv0_workspace_bytes(B, N, C), dtype=torch.uint8, device=wav.device)
    out = torch.empty(B, L0, C, dtype=F16, device=wav.device)
    call("w2v2_conv0_raw", ptr(wav.contiguous()), 1 if wav.dtype == F32 else 0, int(bool(normalize)), B, N,
         ptr(_chk_lens(lens, B)) if lens is not None else None, ptr(w.contiguous()), ptr(gamma), ptr(beta), eps, ptr(ws),
         ptr(out), C, 1, stream_ptr())
    return out


def cast_f16_rowmask(x: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """f32 [B, T, H] -> f16 with the rows t >= lens[b] zeroed."""
    _chk(x, F32, "x")
    B, T, H = x.shape
    y = torch.empty(x.shape, dtype=F16, device=x.device)
    call("w2v2_cast_f16_rowmask", ptr(x.contiguous()), ptr(y), B, T, H, ptr(_chk_lens(lens, B)), stream_ptr())
    return y


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
              want32: bool = True, want16: bool = True, drop_p: float = 0.0, drop_seed: int = 0):
    """LayerNorm(dropout(x + bias) + residual) over the last dim -> (y32 | None, y16 | None)."""
    assert x.is_contiguous()
    H = x.shape[-1]
    rows = x.numel() // H
    y32 = torch.empty(x.shape, dtype=F32, device=x.device) if want32 else None
    y16 = torch.empty(x.shape, dtype=F16, device=x.device) if want16 else None
    call("w2v2_layernorm_ex", ptr(x), 1 if x.dtype == F32 else 0, ptr(bias), ptr(residual), ptr(gamma), ptr(beta),
         eps, ptr(y32), ptr(y16), rows, H, float(drop_p), int(drop_seed), stream_ptr())
    return y32, y16


def layernorm_with_rstd(x, gamma, beta, eps=1e-5, bias=None, residual=None, drop_p: float = 0.0, drop_seed: int = 0):
    """LayerNorm(dropout(x + bias) + residual) -> (y32, y16, rstd [rows]): the forward of the training layers."""
    assert x.is_contiguous()
    H = x.shape[-1]
    rows = x.numel() // H
    y32 = torch.empty(x.shape, dtype=F32, device=x.device)
    y16 = torch.empty(x.shape, dtype=F16, device=x.device)
    rstd = torch.empty(rows, dtype=F32, device=x.device)
    call("w2v2_layernorm_ex2", ptr(x), 1 if x.dtype == F32 else 0, ptr(bias), ptr(residual), ptr(gamma), ptr(beta), eps,
         ptr(y32), ptr(y16), ptr(rstd), rows, H, float(drop_p), int(drop_seed), stream_ptr())
    return y32, y16, rstd


def layernorm_bwd_from_output(dy_a, y32, rstd, gamma, beta, dy_b=None, dgamma=None, dbeta=None, dbias=None,
                              drop_p: float = 0.0, drop_seed: int = 0):
    """LayerNorm backward from the LayerNorm output y32 and the saved rstd -> (dx32, dx16)."""
    H = y32.shape[-1]
    rows = y32.numel() // H
    dx32 = torch.empty(rows, H, dtype=F32, device=y32.device)
    dx16 = torch.empty(rows, H, dtype=F16, device=y32.device)
    call("w2v2_layernorm_bwd_from_output", ptr(dy_a), ptr(dy_b), ptr(y32), ptr(rstd), ptr(gamma), ptr(beta), ptr(dx32),
         ptr(dx16), ptr(dgamma), ptr(dbeta), ptr(dbias), rows, H, float(drop_p), int(drop_seed), stream_ptr())
    return dx32, dx16


def posconv_taps_per_mma(T: int, H: int, groups: int) -> int:
    u = _lib.load().w2v2_posconv_taps_per_mma(T, H, groups)
    if u < 1:
        raise _lib.W2V2Error(f"w2v2_posconv: T={T} frames does not fit the single-slab kernel (needs time tiling)")
    return u


def posconv_fold_weight(v: torch.Tensor, g: torch.Tensor, groups: int, taps_per_mma: int, mode: int = 0) -> torch.Tensor:
    """mode 0: forward weight; mode 1: data-gradient weight (in/out channels swapped, taps reversed)."""
    H, I, K = v.shape
    buf = torch.empty(H * I * K + 2 * K, dtype=F16, device=v.device)
    call("w2v2_posconv_fold_weight", ptr(v.contiguous()), ptr(g.contiguous()), ptr(buf), H, groups, K, taps_per_mma,
         mode, stream_ptr())
    return buf


def posconv(x16: torch.Tensor, w16: torch.Tensor, bias: torch.Tensor, groups: int, K: int) -> torch.Tensor:
    _chk(x16, F16, "x16")
    B, T, H = x16.shape
    out = torch.empty(B, T, H, dtype=F32, device=x16.device)
    call("w2v2_posconv", ptr(x16), ptr(w16), ptr(bias), ptr(out), B, T, H, groups, K, stream_ptr())
    return out


def posconv_ex(x16, w16, bias, groups: int, K: int, act: int, in_shift: int) -> torch.Tensor:
    B, T, H = x16.shape
    out = torch.empty(B, T, H, dtype=F32, device=x16.device)
    call("w2v2_posconv_ex", ptr(x16), ptr(w16), ptr(bias), ptr(out), B, T, H, groups, K, act, in_shift, stream_ptr())
    return out


def gelu_fwd(x: torch.Tensor, out_dtype, want_x16: bool = False):
    """-> (gelu(x) in out_dtype, f16 copy of x | None)."""
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    x16 = torch.empty(x.shape, dtype=F16, device=x.device) if want_x16 else None
    call("w2v2_gelu_fwd", ptr(x), 1 if x.dtype == F32 else 0, ptr(y), 1 if out_dtype == F32 else 0, ptr(x16), x.numel(),
         stream_ptr())
    return y, x16


def posconv_im2col(x16: torch.Tensor, groups: int, K: int, g: int, out: torch.Tensor) -> torch.Tensor:
    B, T, H = x16.shape
    call("w2v2_posconv_im2col", ptr(x16), ptr(out), B, T, H, groups, K, g, stream_ptr())
    return out


def posconv_wgrad(dz16: torch.Tensor, x16: torch.Tensor, groups: int, K: int, dw_hki: torch.Tensor) -> torch.Tensor:
    """dw_hki[o, k*I + i] += sum_{b,t} dz[b,t,o] x[b,t+k-K/2, g(o)*I+i]   (f32 [H, K*I], accumulated)."""
    _chk(dz16, F16, "dz16")
    _chk(x16, F16, "x16")
    B, T, H = x16.shape
    assert dz16.shape == x16.shape and dw_hki.dtype == F32 and dw_hki.is_contiguous()
    assert dw_hki.shape == (H, K * (H // groups))
    call("w2v2_posconv_wgrad", ptr(dz16), ptr(x16), ptr(dw_hki), B, T, H, groups, K, stream_ptr())
    return dw_hki


def weight_norm_bwd(dw_hki, v, g, scale, dv, dg):
    H, I, K = v.shape
    scratch = torch.empty(2 * K, dtype=F32, device=v.device)
    call("w2v2_weight_norm_bwd", ptr(dw_hki), ptr(v), ptr(g), ptr(scratch), float(scale), ptr(dv), ptr(dg), H, I, K,
         stream_ptr())


def attention(qkv16: torch.Tensor, B: int, T: int, H: int, heads: int, want_lse: bool = False, drop_p: float = 0.0,
              drop_seed: int = 0):
    """-> out f16 [B*T, H]  (and lse f32 [B, heads, T] when want_lse, for the backward pass)."""
    _chk(qkv16, F16, "qkv16")
    out = torch.empty(B * T, H, dtype=F16, device=qkv16.device)
    lse = torch.empty(B, heads, T, dtype=F32, device=qkv16.device) if want_lse else None
    call("w2v2_attention_ex", ptr(qkv16), ptr(out), ptr(lse), B, T, H, heads, float(drop_p), int(drop_seed),
         stream_ptr())
    return (out, lse) if want_lse else out


def attention_lens(qkv16: torch.Tensor, B: int, T: int, H: int, heads: int, lens: torch.Tensor) -> torch.Tensor:
    """Attention over a zero-padded ragged batch: keys t >= lens[b] are excluded (evaluation) -> out f16 [B*T, H]."""
    _chk(qkv16, F16, "qkv16")
    out = torch.empty(B * T, H, dtype=F16, device=qkv16.device)
    call("w2v2_attention_lens", ptr(qkv16), ptr(out), B, T, H, heads, ptr(_chk_lens(lens, B)), stream_ptr())
    return out


# ---- backward ------------------------------------------------------------------------------------


def attention_bwd(qkv16, o16, do16, lse, B: int, T: int, H: int, heads: int, drop_p: float = 0.0,
                  drop_seed: int = 0, qscale: float = 1.0, dbias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> dqkv f16 [B*T, 3H] with the q block multiplied by `qscale`; `dbias` (f32 [3H]) accumulates its column sums."""
    dqkv = torch.empty(B * T, 3 * H, dtype=F16, device=qkv16.device)
    call("w2v2_attention_bwd_ex2", ptr(qkv16), ptr(o16), ptr(do16), ptr(lse), ptr(dqkv), B, T, H, heads, float(drop_p),
         int(drop_seed), float(qscale), ptr(dbias), stream_ptr())
    return dqkv


def gemm_wgrad_f16(dy16: torch.Tensor, x16: torch.Tensor, dw: torch.Tensor) -> torch.Tensor:
    """dw[N,K] (f32, accumulated) += dy16[M,N]^T @ x16[M,K]."""
    _chk(dy16, F16, "dy16"); _chk(x16, F16, "x16"); _chk(dw, F32, "dw")
    M, N = dy16.shape
    K = x16.shape[1]
    assert x16.shape[0] == M and tuple(dw.shape) == (N, K)
    call("w2v2_gemm_wgrad_f16", ptr(dy16), dy16.stride(0), ptr(x16), x16.stride(0), M, N, K, ptr(dw), dw.stride(0),
         stream_ptr())
    return dw


def prepare_tile_edge() -> int:
    """Edge of the square tiles the job table's `tile_begin` counts (host-only query)."""
    return int(_lib.load().w2v2_prepare_tile_edge())


def prepare_weights(table_u8: torch.Tensor, njobs: int, tiles: int):
    """table_u8: device copy of an array of w2v2_prep_job records (see engine.WeightPrep)."""
    assert table_u8.dtype == torch.uint8 and table_u8.numel() == 64 * njobs
    call("w2v2_prepare_weights", ptr(table_u8), njobs, tiles, stream_ptr())


def cast_f16_transpose(w: torch.Tensor, ldt: Optional[int] = None, row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """w f32 [R,C] -> f16 [C, ldt] (transposed; ldt >= R, zero padded; multiple of 8)."""
    R, C = w.shape
    ldt = ldt or (R + 63) // 64 * 64
    out = torch.empty(C, ldt, dtype=F16, device=w.device)
    call("w2v2_cast_f16_transpose", ptr(w.contiguous()), ptr(out), R, C, ldt, ptr(row_scale), stream_ptr())
    return out


def layernorm_bwd(dy_a, xa, gamma, eps=1e-5, dy_b=None, bias=None, residual=None, dgamma=None, dbeta=None,
                  want32=True, want16=True, drop_p: float = 0.0, drop_seed: int = 0, dbias=None):
    """-> (dx32: gradient of the residual input, dx16: gradient of the (dropped) branch input xa);
    dbias (f32 [H]) accumulates the column sums of the branch gradient (= gradient of `bias`)."""
    H = xa.shape[-1]
    rows = xa.numel() // H
    dx32 = torch.empty(rows, H, dtype=F32, device=xa.device) if want32 else None
    dx16 = torch.empty(rows, H, dtype=F16, device=xa.device) if want16 else None
    call("w2v2_layernorm_bwd_ex", ptr(dy_a), ptr(dy_b), ptr(xa), 1 if xa.dtype == F32 else 0, ptr(bias), ptr(residual),
         ptr(gamma), eps, ptr(dx32), ptr(dx16), ptr(dgamma), ptr(dbeta), ptr(dbias), rows, H, float(drop_p),
         int(drop_seed), stream_ptr())
    return dx32, dx16


def gelu_bwd(dg16, z16, dbias=None):
    """dz = dg * gelu'(z); with dbias (f32 [cols]): also dbias += column sums of dz, in the same pass."""
    dz = torch.empty_like(dg16)
    if dbias is None:
        call("w2v2_gelu_bwd", ptr(dg16), ptr(z16), ptr(dz), dg16.numel(), stream_ptr())
    else:
        assert dg16.is_contiguous() and z16.is_contiguous()
        cols = dg16.shape[-1]
        call("w2v2_gelu_bwd_colsum", ptr(dg16), ptr(z16), ptr(dz), dg16.numel() // cols, cols, ptr(dbias), stream_ptr())
    return dz


def colsum(x: torch.Tensor, out: torch.Tensor, scale: float = 1.0):
    rows, cols = x.shape
    call("w2v2_colsum", ptr(x), 1 if x.dtype == F32 else 0, rows, cols, x.stride(0), float(scale), ptr(out), stream_ptr())
    return out


def softmax_ce_bwd(prob, labels, coef: float, ldd: int):
    B, S = prob.shape
    dl = torch.empty(B, ldd, dtype=F16, device=prob.device)
    call("w2v2_softmax_ce_bwd", ptr(prob), ptr(labels), float(coef), ptr(dl), B, S, ldd, stream_ptr())
    return dl


def mean_pool_bwd(demb, T: int):
    B, H = demb.shape
    dh = torch.empty(B, T, H, dtype=F32, device=demb.device)
    call("w2v2_mean_pool_bwd", ptr(demb.contiguous()), ptr(dh), B, T, H, stream_ptr())
    return dh


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step: int, grad_scale: float = 1.0, zero_grad: bool = False):
    call("w2v2_adam_step_ex", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
         int(step), float(grad_scale), int(zero_grad), stream_ptr())


def stat_pool(x: torch.Tensor, mode: int, lens: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(x, F32, "x")
    B, T, H = x.shape
    out = torch.empty(B, H * (2 if mode in (1, 3) else 1), dtype=F32, device=x.device)
    call("w2v2_stat_pool_lens", ptr(x.contiguous()), ptr(out), B, T, H, mode,
         ptr(_chk_lens(lens, B)) if lens is not None else None, stream_ptr())
    return out


def asp_concat(x: torch.Tensor) -> torch.Tensor:
    B, T, H = x.shape
    cat = torch.empty(B * T, 3 * H, dtype=F16, device=x.device)
    call("w2v2_asp_concat", ptr(x), ptr(cat), B, T, H, stream_ptr())
    return cat


def asp_concat_split3(x: torch.Tensor, lens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[x | mean | std] as error-compensated operand [hi | lo | hi], f16 [B*T, 9H] (pairs with split3_rows(W, 1))."""
    B, T, H = x.shape
    cat = torch.empty(B * T, 9 * H, dtype=F16, device=x.device)
    call("w2v2_asp_concat_split3_lens", ptr(x), ptr(cat), B, T, H, ptr(_chk_lens(lens, B)) if lens is not None else None,
         stream_ptr())
    return cat


def asp_relu_bn_tanh(z: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, ubias: Optional[torch.Tensor] = None,
                     rows_per_utt: int = 1) -> torch.Tensor:
    """tanh(BN(ReLU(z (+ ubias[utterance])))) -> f16 [rows, A]; ubias f32 [rows / rows_per_utt, A]."""
    rows, A = z.shape
    y = torch.empty(rows, A, dtype=F16, device=z.device)
    if ubias is None:
        call("w2v2_asp_relu_bn_tanh", ptr(z), ptr(scale), ptr(shift), ptr(y), rows, A, stream_ptr())
    else:
        _chk(ubias, F32, "ubias")
        assert ubias.is_contiguous() and ubias.shape == (rows // rows_per_utt, A)
        call("w2v2_asp_relu_bn_tanh_ubias", ptr(z), ptr(scale), ptr(shift), ptr(ubias), rows_per_utt, ptr(y), rows, A,
             stream_ptr())
    return y


def asp_pool(x: torch.Tensor, logits: torch.Tensor, lens: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, T, H = x.shape
    out = torch.empty(B, 2 * H, dtype=F32, device=x.device)
    call("w2v2_asp_pool_lens", ptr(x), ptr(logits), ptr(out), B, T, H, ptr(_chk_lens(lens, B)) if lens is not None else None,
         stream_ptr())
    return out


def _chk_labels(labels: torch.Tensor, B: int) -> torch.Tensor:
    """Class indices as the kernels read them: a CUDA, contiguous int64 vector of B entries (the range check
    0 <= label < S is done on the device: the kernel traps, like torch's device-side assert)."""
    if not isinstance(labels, torch.Tensor) or not labels.is_cuda:
        raise ValueError("labels must be a CUDA tensor (there is no CPU path)")
    if labels.dtype != torch.int64:
        raise TypeError(f"labels must be int64 class indices, got {labels.dtype}")
    if labels.numel() != B or labels.dim() != 1:
        raise ValueError(f"labels must have shape [{B}], got {tuple(labels.shape)}")
    return labels if labels.is_contiguous() else labels.contiguous()


def softmax_ce(logits: torch.Tensor, labels: torch.Tensor, want_prob: bool = True):
    """-> (prob [B,S] | None, loss_rows [B], argmax [B] int32).  logits f32 [B,S] (any row pitch)."""
    B, S = logits.shape
    labels = _chk_labels(labels, B)
    prob = torch.empty(B, S, dtype=F32, device=logits.device) if want_prob else None
    loss = torch.empty(B, dtype=F32, device=logits.device)
    am = torch.empty(B, dtype=torch.int32, device=logits.device)
    call("w2v2_softmax_ce", ptr(logits), logits.stride(0), ptr(labels), ptr(prob), ptr(loss), ptr(am), B, S,
         stream_ptr())
    return prob, loss, am


def aam_softmax_ce(cosine: torch.Tensor, labels: torch.Tensor, margin: float, scale: float,
                   easy_margin: bool = False, want_prob: bool = True):
    """In place on `cosine` (becomes the scaled margin logits).  -> (prob, loss_rows, argmax)."""
    B, S = cosine.shape
    labels = _chk_labels(labels, B)
    prob = torch.empty(B, S, dtype=F32, device=cosine.device) if want_prob else None
    loss = torch.empty(B, dtype=F32, device=cosine.device)
    am = torch.empty(B, dtype=torch.int32, device=cosine.device)
    call("w2v2_aam_softmax_ce", ptr(cosine), cosine.stride(0), ptr(labels), float(margin), float(scale),
         int(easy_margin), ptr(prob), ptr(loss), ptr(am), B, S, stream_ptr())
    return prob, loss, am


def l2norm_rows_split3(x: torch.Tensor, which: int) -> torch.Tensor:
    rows, E = x.shape
    y = torch.empty(rows, 3 * E, dtype=F16, device=x.device)
    call("w2v2_l2norm_rows_split3", ptr(x.contiguous()), ptr(y), rows, E, which, stream_ptr())
    return y


def l2norm_rows_f16(x: torch.Tensor) -> torch.Tensor:
    rows, E = x.shape
    y = torch.empty(rows, E, dtype=F16, device=x.device)
    call("w2v2_l2norm_rows_f16", ptr(x.contiguous()), ptr(y), rows, E, stream_ptr())
    return y


def split3_rows(x: torch.Tensor, which: int) -> torch.Tensor:
    rows, E = x.shape
    y = torch.empty(rows, 3 * E, dtype=F16, device=x.device)
    call("w2v2_split3_rows", ptr(x.contiguous()), ptr(y), rows, E, which, stream_ptr())
    return y


def mean_rows(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty(1, dtype=F32, device=x.device)
    call("w2v2_mean_rows", ptr(x), ptr(out), x.numel(), stream_ptr())
    return out[0]


def add2_cast(a, b=None, want32=True, want16=True):
    o32 = torch.empty_like(a) if want32 else None
    o16 = torch.empty(a.shape, dtype=F16, device=a.device) if want16 else None
    call("w2v2_add2_cast", ptr(a), ptr(b), ptr(o32), ptr(o16), a.numel(), stream_ptr())
    return o32, o16


def cast_f16_rows(x: torch.Tensor, ldy: int, scale: float = 1.0) -> torch.Tensor:
    rows, cols = x.shape
    y = torch.empty(rows, ldy, dtype=F16, device=x.device)
    call("w2v2_cast_f16_rows", ptr(x), x.stride(0), ptr(y), ldy, rows, cols, float(scale), stream_ptr())
    return y


def scale_f32_(x: torch.Tensor, s: float):
    call("w2v2_scale_f32", ptr(x), x.numel(), float(s), stream_ptr())
    return x


def scaled_copy_f32(x: torch.Tensor, s: float) -> torch.Tensor:
    """x * s into a fresh tensor (x: f32, any shape, contiguous)."""
    _chk(x, F32, "x")
    assert x.is_contiguous()
    y = torch.empty_like(x)
    call("w2v2_scale_copy_f32", ptr(x), ptr(y), x.numel(), float(s), stream_ptr())
    return y


def grad_entry_scale(x: torch.Tensor, s: float, lo: float = 2.0 ** -16, hi: float = 2.0 ** 8, mid: float = 1.0):
    """-> (x * s * k as a fresh f32 tensor, inv = device scalar holding 1 / k).  k is 1 unless amax|x| * s left [lo, hi]
    (w2v2_grad_entry_scale): the hook that keeps an outer GradScaler's 2^16 out of the fp16 gradient operands."""
    _chk(x, F32, "x")
    assert x.is_contiguous()
    y = torch.empty_like(x)
    state = torch.empty(2, dtype=F32, device=x.device)
    call("w2v2_grad_entry_scale", ptr(x), ptr(y), x.numel(), float(s), float(lo), float(hi), float(mid), ptr(state),
         stream_ptr())
    return y, state[1:]


def scale_f32_dev_(x: torch.Tensor, s: float, dev_scale: torch.Tensor):
    """x *= s * dev_scale[0] in place."""
    call("w2v2_scale_f32_dev", ptr(x), x.numel(), float(s), ptr(dev_scale), stream_ptr())
    return x


def softmax_ce_bwd_f32(prob, labels, dloss, coef: float):
    B, S = prob.shape
    dl = torch.empty(B, S, dtype=F32, device=prob.device)
    call("w2v2_softmax_ce_bwd_f32", ptr(prob), ptr(labels), ptr(dloss), float(coef), ptr(dl), B, S, stream_ptr())
    return dl


def dropout_(x: torch.Tensor, p: float, seed: int, bias: Optional[torch.Tensor] = None, want16: bool = False):
    """In place: x = keep ? (x + bias) / (1-p) : 0.  Returns (x, f16 copy | None)."""
    H = x.shape[-1]
    y16 = torch.empty(x.shape, dtype=F16, device=x.device) if want16 else None
    call("w2v2_dropout", ptr(x), 1 if x.dtype == F32 else 0, ptr(bias), H, ptr(x), ptr(y16), x.numel(), float(p), int(seed),
         stream_ptr())
    return x, y16


def time_mask_apply_(h: torch.Tensor, mask_u8: torch.Tensor, embed: torch.Tensor):
    rows, H = h.shape
    call("w2v2_time_mask_apply", ptr(h), ptr(mask_u8), ptr(embed), rows, H, stream_ptr())
    return h


def feature_mask_(h: torch.Tensor, mask_u8: torch.Tensor, B: int, T: int):
    """In place: h[b, t, c] = 0 where mask_u8[b * H + c] (SpecAugment along the feature axis; also its backward)."""
    _chk(h, F32, "h")
    H = h.shape[-1]
    assert h.is_contiguous() and h.numel() == B * T * H and mask_u8.dtype == torch.uint8 and mask_u8.numel() == B * H
    call("w2v2_feature_mask", ptr(h), ptr(mask_u8), B, T, H, stream_ptr())
    return h


def time_mask_bwd_(dh: torch.Tensor, mask_u8: torch.Tensor, dembed: torch.Tensor, scale: float = 1.0):
    rows, H = dh.shape
    call("w2v2_time_mask_bwd", ptr(dh), ptr(mask_u8), ptr(dembed), rows, H, float(scale), stream_ptr())
    return dh


def meanstd_pool_bwd(x: torch.Tensor, dout: torch.Tensor) -> torch.Tensor:
    B, T, H = x.shape
    dx = torch.empty_like(x)
    call("w2v2_meanstd_pool_bwd", ptr(x), ptr(dout.contiguous()), ptr(dx), B, T, H, stream_ptr())
    return dx


def aam_softmax_ce_train(cosine, labels, margin, scale, easy_margin=False):
    """Like aam_softmax_ce, additionally returning the pre-margin label cosines (saved for backward)."""
    B, S = cosine.shape
    labels = _chk_labels(labels, B)
    prob = torch.empty(B, S, dtype=F32, device=cosine.device)
    loss = torch.empty(B, dtype=F32, device=cosine.device)
    am = torch.empty(B, dtype=torch.int32, device=cosine.device)
    cl = torch.empty(B, dtype=F32, device=cosine.device)
    call("w2v2_aam_softmax_ce_ex", ptr(cosine), cosine.stride(0), ptr(labels), float(margin), float(scale),
         int(easy_margin), ptr(prob), ptr(loss), ptr(am), ptr(cl), B, S, stream_ptr())
    return prob, loss, am, cl


def aam_bwd_dcos(prob, cos_label, labels, dloss, coef, margin, scale, easy_margin, ldd):
    B, S = prob.shape
    dc = torch.empty(B, ldd, dtype=F16, device=prob.device)
    call("w2v2_aam_bwd_dcos", ptr(prob), ptr(cos_label), ptr(labels), ptr(dloss), float(coef), float(margin), float(scale),
         int(easy_margin), ptr(dc), B, S, ldd, stream_ptr())
    return dc


def row_inv_norm(x: torch.Tensor) -> torch.Tensor:
    rows, E = x.shape
    inv = torch.empty(rows, dtype=F32, device=x.device)
    call("w2v2_row_inv_norm", ptr(x.contiguous()), ptr(inv), rows, E, stream_ptr())
    return inv


def l2norm_rows_bwd(x: torch.Tensor, dxh: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    rows, E = x.shape
    dx = torch.empty(rows, E, dtype=F32, device=x.device)
    call("w2v2_l2norm_rows_bwd", ptr(x.contiguous()), ptr(dxh), dxh.stride(0), ptr(dx), rows, E, float(scale), 0,
         stream_ptr())
    return dx


# ---- attentive-statistics pooling, training ------------------------------------------------------


def asp_bn_batch_stats(z: torch.Tensor, gamma, beta, eps: float, momentum: float, running_mean, running_var):
    """Training-mode BatchNorm1d over the rows of relu(z) -> (scale, shift, mean, rstd); updates the running stats."""
    rows, A = z.shape
    dev = z.device
    ws = torch.empty(2 * A, dtype=torch.float64, device=dev)
    scale, shift, mean, rstd = (torch.empty(A, dtype=F32, device=dev) for _ in range(4))
    call("w2v2_asp_bn_batch_stats", ptr(z), rows, A, ptr(gamma), ptr(beta), float(eps), float(momentum), ptr(running_mean),
         ptr(running_var), ptr(ws), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), stream_ptr())
    return scale, shift, mean, rstd


def asp_pool_bwd(x: torch.Tensor, logits: torch.Tensor, out: torch.Tensor, dout: torch.Tensor):
    """-> (dlogits f16 [B*T, C], dx_direct f32 [B,T,C])."""
    B, T, C = x.shape
    dlg = torch.empty(B * T, C, dtype=F16, device=x.device)
    dx = torch.empty_like(x)
    call("w2v2_asp_pool_bwd", ptr(x), ptr(logits), ptr(out), ptr(dout.contiguous()), ptr(dlg), ptr(dx), B, T, C, stream_ptr())
    return dlg, dx


def asp_act_bwd(dh: torch.Tensor, z: torch.Tensor, scale, shift, mean, rstd, batch_stats: bool, dgamma, dbeta,
                grad_scale: float) -> torch.Tensor:
    rows, A = z.shape
    ws = torch.empty(2 * A, dtype=torch.float64, device=z.device)
    dz = torch.empty(rows, A, dtype=F16, device=z.device)
    call("w2v2_asp_act_bwd", ptr(dh), ptr(z), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), int(batch_stats), ptr(ws), ptr(dz),
         ptr(dgamma), ptr(dbeta), float(grad_scale), rows, A, stream_ptr())
    return dz


def asp_front_bwd_(x: torch.Tensor, dcat: torch.Tensor, dx: torch.Tensor) -> torch.Tensor:
    B, T, C = x.shape
    assert dcat.stride(1) == 1 and dcat.shape == (B * T, 3 * C)
    call("w2v2_asp_front_bwd", ptr(x), ptr(dcat), dcat.stride(0), ptr(dx), B, T, C, stream_ptr())
    return dx


# ---- CNN feature extractor, training ---------------------------------------------------------------


def conv0_gn(wav: torch.Tensor, w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, act: int):
    """conv layer 0 + GroupNorm (+ GELU if act) -> (f16 channels-last [B, L0, C], workspace tensor, (scale, im2col) views).
    The workspace keeps the per-(b, c) affine and the im2col operand the backward needs."""
    import ctypes
    _chk(wav, F32, "wav")
    B, N = wav.shape
    C = w.shape[0]
    L0 = (N - 10) // 5 + 1
    lib = _lib.load()
    ws = torch.empty(lib.w2v2_conv0_workspace_bytes(B, N, C), dtype=torch.uint8, device=wav.device)
    out = torch.empty(B, L0, C, dtype=F16, device=wav.device)
    call("w2v2_conv0_gn_ex", ptr(wav.contiguous()), B, N, ptr(w.contiguous()), ptr(gamma), ptr(beta), eps, ptr(ws), ptr(out), C,
         int(act), stream_ptr())
    so, sho, ao = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    call("w2v2_conv0_workspace_offsets", B, N, C, ctypes.byref(so), ctypes.byref(sho), ctypes.byref(ao))
    scale = ws[so.value:so.value + B * C * 4].view(F32).view(B, C)
    im2col = ws[ao.value:ao.value + B * L0 * 64 * 2].view(F16).view(B, L0, 64)
    return out, ws, scale, im2col


def gemm_f16_taps(a16: torch.Tensor, tap_rows, w16: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[b, r, :] = sum_t a16[b, r + tap_rows[t], :] @ w16[:, t*C:(t+1)*C]^T  (rows outside a16 read as zero).
    a16 f16 [B, L, C] contiguous; w16 f16 [N, ntaps*C]; out [B, R, N] with stride(2) == 1 (rows may be strided)."""
    import ctypes
    _chk(a16, F16, "a16"); _chk(w16, F16, "w16")
    B, L, C = a16.shape
    assert a16.is_contiguous() and out.stride(2) == 1 and out.shape[0] == B
    nt = len(tap_rows)
    N = w16.shape[0]
    assert w16.shape[1] == nt * C and out.shape[2] == N
    taps = (ctypes.c_int * 3)(*(list(tap_rows) + [0] * (3 - nt)))
    call("w2v2_gemm_f16_taps", ptr(a16), out.shape[1], L, taps, C, L * C, B, nt, C, ptr(w16), w16.stride(0), N, ptr(out),
         1 if out.dtype == F32 else 0, out.stride(1), out.stride(0), stream_ptr())
    return out


def gemm_wgrad_f16_batched(dy16: torch.Tensor, x16: torch.Tensor, dw: torch.Tensor) -> torch.Tensor:
    """dw[N, K] (f32, accumulated) += sum_b dy16[b]^T @ x16[b];  dy16 [B, R, N], x16 [B, R, K] (any row / batch strides)."""
    _chk(dy16, F16, "dy16"); _chk(x16, F16, "x16"); _chk(dw, F32, "dw")
    B, R, N = dy16.shape
    K = x16.shape[2]
    assert x16.shape[:2] == (B, R) and dy16.stride(2) == 1 and x16.stride(2) == 1 and dw.shape == (N, K) and dw.stride(1) == 1
    call("w2v2_gemm_wgrad_f16_batched", ptr(dy16), dy16.stride(1), dy16.stride(0), ptr(x16), x16.stride(1), x16.stride(0), R, B,
         N, K, ptr(dw), dw.stride(0), stream_ptr())
    return dw


def groupnorm_bwd(dy16, y16, gamma, beta, scale, dgamma, dbeta, grad_scale: float) -> torch.Tensor:
    B, L, C = y16.shape
    dc = torch.empty(B, L, C, dtype=F16, device=y16.device)
    call("w2v2_groupnorm_bwd", ptr(dy16), ptr(y16), ptr(gamma), ptr(beta), ptr(scale), ptr(dc), ptr(dgamma), ptr(dbeta),
         float(grad_scale), B, L, C, stream_ptr())
    return dc


# ---- evaluation ----------------------------------------------------------------------------------


def cosine_pairs(emb: torch.Tensor, idx_a: torch.Tensor, idx_b: torch.Tensor, mean: Optional[torch.Tensor] = None,
                 std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scores[p] = cosine similarity of (optionally centred) emb[idx_a[p]] and emb[idx_b[p]]; emb f32 [N, E], idx int32 [P]."""
    _chk(emb, F32, "emb")
    assert idx_a.dtype == torch.int32 and idx_b.dtype == torch.int32 and idx_a.shape == idx_b.shape
    P = idx_a.numel()
    scores = torch.empty(P, dtype=F32, device=emb.device)
    call("w2v2_cosine_pairs", ptr(emb.contiguous()), ptr(mean), ptr(std), ptr(idx_a.contiguous()), ptr(idx_b.contiguous()),
         ptr(scores), P, emb.shape[1], stream_ptr())
    return scores


def normalize_wav(wav: torch.Tensor):
    """Per-utterance standardisation of a [B, N] batch (float32 or int16 PCM) -> (f32 [B, N], mean [B], std [B])."""
    if not wav.is_cuda:
        raise _lib.W2V2Error("normalize_wav: expected a CUDA tensor (no CPU fallback exists)")
    if wav.dtype not in (torch.float32, torch.int16):
        raise _lib.W2V2Error(f"normalize_wav: expected float32 or int16, got {wav.dtype}")
    B, N = wav.shape
    wav = wav.contiguous()
    out = torch.empty(B, N, dtype=F32, device=wav.device)
    mean = torch.empty(B, dtype=F32, device=wav.device)
    std = torch.empty(B, dtype=F32, device=wav.device)
    call("w2v2_normalize_wav", ptr(wav), 1 if wav.dtype == torch.float32 else 0, ptr(out), ptr(mean), ptr(std), B, N,
         stream_ptr())
    return out, mean, std
