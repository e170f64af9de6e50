"""Training forward / backward of the STABLE-LAYER-NORM transformer stack (the wav2vec2 "-lv60" / XLSR checkpoints:
``do_stable_layer_norm=True``, HF:731-799 encoder, HF:632-655 layers):

    h = drop(h0 + GELU(posconv(h0)))
    per layer:   h1 = h + drop(attn(LN1(h)) Wo^T + bo);   h2 = h1 + drop(W2 gelu(W1 LN2(h1) + b1) + b2)
    out = LN(h_L)

The post-LN stack (training.stack_forward_train / stack_backward) runs one native schedule call per layer with the
residual add fused into its LayerNorm kernels; the pre-LN order keeps the residual stream un-normalised, so this variant
composes the same GEMM / attention / LayerNorm / dropout kernels launch by launch from Python with the residual sums as
their own fp32 passes.  It is built for coverage of those checkpoints (no reference configuration names one), not tuned:
same numerics (fp16 operands, fp32 accumulation / residual stream / statistics, LOSS_SCALE on every gradient), same
gradient layout (training.stack_grad_order), same dropout seeds per site as csrc/schedule.cu."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

F16, F32 = torch.float16, torch.float32


def stack_forward_train_stable(eng, h0: torch.Tensor, B: int, T: int, plan, S: dict) -> torch.Tensor:
    """h0: f32 [B*T, H] (left untouched).  Fills `S` for stack_backward_stable; -> last_hidden_state f32 [B, T, H]."""
    a, w = eng.arch, eng.w
    H, M = a.hidden, B * T
    ph = plan.p_hidden if plan is not None else 0.0
    pa = plan.p_attn if plan is not None else 0.0
    pact = plan.p_act if plan is not None else 0.0
    seed = plan.seed if plan is not None else 0
    x16 = ops.cast_f16(h0)
    zpos = ops.posconv_ex(x16.view(B, T, H), w.pos_w(T), w.pos_b, a.pos_groups, a.pos_kernel, 0, 0)
    pos, zpos16 = ops.gelu_fwd(zpos.view(M, H), F32, want_x16=True)
    h, _ = ops.add2_cast(h0, pos, want16=False)                          # HF:760-761
    if ph > 0:
        ops.dropout_(h, ph, seed + 2)
    S.update(B=B, T=T, h0=h0, x16=x16, zpos16=zpos16, layers=[], stable=True)
    for l, lw in enumerate(w.layers):
        if plan is not None and plan.skip[l]:                            # LayerDrop (HF:771-778)
            S["layers"].append(None)
            continue
        _, a16 = ops.layernorm(h, lw["ln1_g"], lw["ln1_b"], a.eps, want32=False)
        qkv16 = ops.gemm_f16(a16, lw["wqkv"], lw["bqkv"], 0, F16)
        att16, lse = ops.attention(qkv16, B, T, H, a.heads, want_lse=True, drop_p=pa, drop_seed=seed + 100 + l)
        o32 = ops.gemm_f16(att16, lw["wo"], lw["bo"], 0, F32)
        if ph > 0:
            ops.dropout_(o32, ph, seed + 200 + l)
        h1, _ = ops.add2_cast(o32, h, want16=False)
        _, c16 = ops.layernorm(h1, lw["ln2_g"], lw["ln2_b"], a.eps, want32=False)
        g16, z16 = ops.gemm_f16_dual_gelu(c16, lw["w1"], lw["b1"])
        if pact > 0:
            ops.dropout_(g16, pact, seed + 400 + l)
        f2 = ops.gemm_f16(g16, lw["w2"], lw["b2"], 0, F32)
        if ph > 0:
            ops.dropout_(f2, ph, seed + 300 + l)
        h2, _ = ops.add2_cast(f2, h1, want16=False)
        S["layers"].append(dict(h_in=h, a16=a16, qkv16=qkv16, att16=att16, lse=lse, h1=h1, c16=c16, z16=z16, g16=g16))
        h = h2
    S["h_last"] = h
    out, _ = ops.layernorm(h, w.enc_ln_g, w.enc_ln_b, a.eps, want16=False)                 # HF:792
    return out.view(B, T, H)


def _branch_grad16(d32: torch.Tensor, p: float, seed: int) -> torch.Tensor:
    """Gradient entering a dropped branch: the residual gradient with the branch's dropout mask applied, in fp16."""
    if p > 0:
        c, _ = ops.add2_cast(d32, None, want16=False)                    # (never drop in place in the residual gradient)
        _, c16 = ops.dropout_(c, p, seed, want16=True)
        return c16
    return ops.cast_f16(d32)


def stack_backward_stable(eng, tw, S: dict, dh: torch.Tensor, G, on_layer_done=None):
    """Backward of stack_forward_train_stable.  dh: f32 [B,T,H], loss-scaled.  Accumulates the gradients of the final
    LayerNorm, every layer and the positional conv into G; -> the two (loss-scaled, f32 [M, H]) terms of d h0: through
    the residual stream and through the positional conv."""
    a, w = eng.arch, eng.w
    B, T = S["B"], S["T"]
    H, M = a.hidden, B * T
    plan = S.get("plan")
    ph = plan.p_hidden if plan is not None else 0.0
    pa = plan.p_attn if plan is not None else 0.0
    pact = plan.p_act if plan is not None else 0.0
    seed = plan.seed if plan is not None else 0
    qscale = float(H // a.heads) ** -0.5

    def layer_done(l):
        if on_layer_done is not None:
            lo = G.offsets[f"encoder.layers.{l}.attention.q_proj.weight"]
            hi = G.offsets[f"encoder.layers.{l + 1}.attention.q_proj.weight"] if l + 1 < a.layers else G.numel
            on_layer_done(lo, hi)

    d, _ = ops.layernorm_bwd(dh.contiguous().view(M, H), S["h_last"], w.enc_ln_g, a.eps,
                             dgamma=G.view("encoder.layer_norm.weight"), dbeta=G.view("encoder.layer_norm.bias"),
                             want16=False)
    for l in reversed(range(a.layers)):
        L = S["layers"][l]
        if L is None:
            layer_done(l)
            continue
        pre = f"encoder.layers.{l}."
        tl = tw.layers[l]
        lw = w.layers[l]
        g = lambda k: G.view(pre + k)
        # h2 = h1 + drop(g W2^T + b2)
        df2 = _branch_grad16(d, ph, seed + 300 + l)
        ops.colsum(df2, g("feed_forward.output_dense.bias"))
        ops.gemm_wgrad_f16(df2, L["g16"], g("feed_forward.output_dense.weight"))
        if pact > 0:
            dg16 = ops.gemm_f16(df2, tl["w2T"], None, 0, F16)
            ops.dropout_(dg16, pact, seed + 400 + l)
            dz16 = ops.gelu_bwd(dg16.contiguous(), L["z16"], dbias=g("feed_forward.intermediate_dense.bias"))
        else:
            dz16 = ops.gemm_f16_gelu_bwd(df2, tl["w2T"], L["z16"], g("feed_forward.intermediate_dense.bias"))
        ops.gemm_wgrad_f16(dz16, L["c16"], g("feed_forward.intermediate_dense.weight"))
        dc32 = ops.gemm_f16(dz16, tl["w1T"], None, 0, F32)
        dln, _ = ops.layernorm_bwd(dc32, L["h1"], lw["ln2_g"], a.eps, dgamma=g("final_layer_norm.weight"),
                                   dbeta=g("final_layer_norm.bias"), want16=False)
        d, _ = ops.add2_cast(dln, d, want16=False)                       # d h1
        # h1 = h + drop(att Wo^T + bo)
        do16 = _branch_grad16(d, ph, seed + 200 + l)
        ops.colsum(do16, g("attention.out_proj.bias"))
        ops.gemm_wgrad_f16(do16, L["att16"], g("attention.out_proj.weight"))
        datt16 = ops.gemm_f16(do16, tl["woT"], None, 0, F16)
        # (the q projection was used pre-scaled by d^-0.5: the attention backward multiplies dq by the same factor and
        # emits the q / k / v bias gradients; wqkvT holds the unscaled Wq^T -- as in csrc/schedule.cu)
        dqkv16 = ops.attention_bwd(L["qkv16"], L["att16"], datt16.contiguous(), L["lse"], B, T, H, a.heads, drop_p=pa,
                                   drop_seed=seed + 100 + l, qscale=qscale,
                                   dbias=G.span(pre + "attention.q_proj.bias", 1, 3 * H).view(-1))
        ops.gemm_wgrad_f16(dqkv16, L["a16"], G.span(pre + "attention.q_proj.weight", 3 * H, H))
        da32 = ops.gemm_f16(dqkv16, tl["wqkvT"], None, 0, F32)
        dln, _ = ops.layernorm_bwd(da32, L["h_in"], lw["ln1_g"], a.eps, dgamma=g("layer_norm.weight"),
                                   dbeta=g("layer_norm.bias"), want16=False)
        d, _ = ops.add2_cast(dln, d, want16=False)                       # d h_in
        S["layers"][l] = None                                            # release this layer's activations
        layer_done(l)
    # top:  h = drop(h0 + pos),  pos = GELU(zpos),  zpos = posconv(h0) + b
    if ph > 0:
        if d.data_ptr() == dh.data_ptr():
            d, _ = ops.add2_cast(d, None, want16=False)
        ops.dropout_(d, ph, seed + 2)
    d16 = ops.cast_f16(d)
    dz16 = ops.gelu_bwd(d16, S["zpos16"], dbias=G.view("encoder.pos_conv_embed.conv.bias"))
    dx_pos = ops.posconv_ex(dz16.view(B, T, H), tw.pos_dgrad_w(T), None, a.pos_groups, a.pos_kernel, 0, 1)
    I, K = H // a.pos_groups, a.pos_kernel
    dw_hki = torch.zeros(H, K * I, dtype=F32, device=dh.device)
    ops.posconv_wgrad(dz16.view(B, T, H), S["x16"].view(B, T, H), a.pos_groups, K, dw_hki)
    ops.weight_norm_bwd(dw_hki, w._pos_v, w._pos_g, 1.0,
                        G.view("encoder.pos_conv_embed.conv.parametrizations.weight.original1"),
                        G.view("encoder.pos_conv_embed.conv.parametrizations.weight.original0").view(-1))
    return d, dx_pos.view(M, H)
