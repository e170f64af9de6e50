"""Training path: forward that keeps what the backward needs, the hand-scheduled backward over the
sm_100a kernels, and the torch.autograd.Function wrappers that make ``loss.backward()`` work on the
reference-facing modules (the reference trains through Lightning's automatic optimisation,
R:src/lightning_modules/speaker/speaker_recognition_module.py:148-220).

Scope of this round: wav2vec2-base/large encoder with the CNN feature extractor frozen (the reference
default ``completely_freeze_feature_extractor: true``, R:config/network/wav2vec2_fc.yaml:16), all
stochastic regularisation at probability 0 (dropout / LayerDrop / SpecAugment), mean pooling + CE head.

Gradient convention between the Functions of this module: activation gradients carry ``LOSS_SCALE``
(their fp16 copies feed the tensor cores), parameter gradients are unscaled before they are returned
to autograd.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .engine import ArchConfig, EncoderEngine, PreparedWeights

F16, F32 = torch.float16, torch.float32

LOSS_SCALE = 4096.0          # static scale of the activation gradients (fp16 operands of dgrad / wgrad)


class TrainWeights:
    """Transposed fp16 weight copies for the data-gradient GEMMs (dX = dY W == gemm(dY, W^T))."""

    def __init__(self, w: PreparedWeights, p: Dict[str, torch.Tensor]):
        a = w.arch
        H, FF, dev = a.hidden, a.ffn, w.fp_w.device
        prep = w.prep                      # the transposes ride on the batched weight-preparation launch

        def tbuf(cols, rows):              # [cols(in), rows(out)] = W^T
            return torch.empty(cols, rows, dtype=F16, device=dev)

        self.fp_wT = tbuf(a.conv_dim, H)                                                          # [512, H]
        prep.add_transposed("feature_projection.projection.weight", self.fp_wT)
        self.layers = []
        for l in range(a.layers):
            pre = f"encoder.layers.{l}."
            L = dict(wqkvT=tbuf(H, 3 * H), woT=tbuf(H, H), w1T=tbuf(H, FF), w2T=tbuf(FF, H))
            for i, n in enumerate("qkv"):                                                      # folded, [H, 3H]
                prep.add_transposed(pre + f"attention.{n}_proj.weight", L["wqkvT"][:, i * H:(i + 1) * H])
            prep.add_transposed(pre + "attention.out_proj.weight", L["woT"])
            prep.add_transposed(pre + "feed_forward.intermediate_dense.weight", L["w1T"])
            prep.add_transposed(pre + "feed_forward.output_dense.weight", L["w2T"])
            self.layers.append(L)
        prep.run()
        self._pos_dgrad = {}
        self._w = w

    def pos_dgrad_w(self, T: int) -> torch.Tensor:
        w = self._w
        u = ops.posconv_taps_per_mma(T, w.arch.hidden, w._groups)
        if u not in self._pos_dgrad:
            self._pos_dgrad[u] = ops.posconv_fold_weight(w._pos_v, w._pos_g, w._groups, u, mode=1)
        return self._pos_dgrad[u]


# ---------------------------------------------------------------------------------------------------
# encoder forward (training) / backward


class RegPlan:
    """Per-step draw of the stochastic regularisation (HF semantics, reference defaults
    R:src/models/wav2vec2.py:83-94): dropout probabilities + counter-RNG seed, the LayerDrop decisions
    (HF:701-713) and the SpecAugment time mask (HF:101-217, 1280-1324; spans of `mask_time_length`
    frames, at least two per utterance), all drawn on the host like the reference does."""

    def __init__(self, reg, layers: int, B: int, T: int, rng, device):
        self.p_feat = float(reg.feat_proj_dropout)
        self.p_hidden = float(reg.hidden_dropout)
        self.p_attn = float(reg.attention_dropout)
        self.p_act = float(reg.activation_dropout)
        self.seed = int(rng.integers(1, 1 << 62))
        self.skip = [bool(rng.random() < reg.layerdrop) for _ in range(layers)]
        self.mask = None
        if reg.mask_time_prob > 0:
            self.mask = torch.from_numpy(compute_time_mask(B, T, reg.mask_time_prob, reg.mask_time_length, 2, rng)).to(device)

    @property
    def any(self) -> bool:
        return (self.p_feat + self.p_hidden + self.p_attn + self.p_act) > 0 or any(self.skip) or self.mask is not None


def compute_time_mask(B: int, T: int, mask_prob: float, mask_length: int, min_masks: int, rng):
    """uint8 [B*T]; own restatement of HF `_compute_mask_indices` (HF:101-217) for full-length inputs:
    num_spans = max(min_masks, int(mask_prob*T/mask_length + U[0,1))), clipped so the spans fit; span starts
    drawn without replacement from [0, T - mask_length]."""
    import numpy as np
    mask = np.zeros((B, T), dtype=np.uint8)
    if mask_length < 1 or mask_length > T:
        return mask.reshape(-1)
    for b in range(B):
        n = int(mask_prob * T / mask_length + rng.random())
        n = max(n, min_masks)
        if n * mask_length > T:
            n = T // mask_length
        starts = rng.choice(T - (mask_length - 1), size=n, replace=False)
        for s in starts:
            mask[b, s:s + mask_length] = 1
    return mask.reshape(-1)


def encoder_forward_train(eng: EncoderEngine, wav: torch.Tensor, plan: Optional[RegPlan] = None,
                          mask_embed: Optional[torch.Tensor] = None):
    """Same arithmetic as EncoderEngine.forward (the GELUs run as separate passes so the pre-activations
    can be kept) plus the train-mode regularisation of `plan`; returns (last_hidden_state f32 [B,T,H], saved)."""
    a, w = eng.arch, eng.w
    S = {"plan": plan}
    ph = plan.p_hidden if plan is not None else 0.0
    seed = plan.seed if plan is not None else 0
    feat = eng.feature_extractor(wav)                         # frozen CNN: nothing saved from inside
    B, T, C = feat.shape
    H, M = a.hidden, B * T
    feat2 = feat.contiguous().view(M, C)
    _, n16 = ops.layernorm(feat2, w.fp_ln_g, w.fp_ln_b, a.eps, want32=False)
    h0 = ops.gemm_f16(n16, w.fp_w, w.fp_b, 0, F32).contiguous()          # [M,H]
    if plan is not None and plan.p_feat > 0:
        ops.dropout_(h0, plan.p_feat, seed + 1)                          # HF:433
    if plan is not None and plan.mask is not None:
        ops.time_mask_apply_(h0, plan.mask, mask_embed)                  # HF:1301-1310
    x16 = ops.cast_f16(h0)
    zpos = ops.posconv_ex(x16.view(B, T, H), w.pos_w(T), w.pos_b, a.pos_groups, a.pos_kernel, 0, 0)
    pos, zpos16 = ops.gelu_fwd(zpos.view(M, H), F32, want_x16=True)
    h32, h16 = ops.layernorm(pos, w.enc_ln_g, w.enc_ln_b, a.eps, residual=h0)
    if ph > 0:
        h32, h16 = ops.dropout_(h32, ph, seed + 2, want16=True)          # HF:693
    S.update(B=B, T=T, feat=feat2, n16=n16, h0=h0, x16=x16, pos=pos, zpos16=zpos16, layers=[])
    for l, lw in enumerate(w.layers):
        if plan is not None and plan.skip[l]:                            # LayerDrop (HF:701-713)
            S["layers"].append(None)
            continue
        L = dict(h_in32=h32, h_in16=h16)
        L["qkv"] = ops.gemm_f16(h16, lw["wqkv"], lw["bqkv"], 0, F16)
        L["att"], L["lse"] = ops.attention(L["qkv"], B, T, H, a.heads, want_lse=True,
                                           drop_p=plan.p_attn if plan is not None else 0.0, drop_seed=seed + 100 + l)
        L["o"] = ops.gemm_f16(L["att"], lw["wo"], None, 0, F32).contiguous()
        # x1 = h_in + drop(o + bo) (HF:546-549), the dropout is generated inside the LayerNorm kernel
        h32, h16 = ops.layernorm(L["o"], lw["ln1_g"], lw["ln1_b"], a.eps, bias=lw["bo"], residual=L["h_in32"],
                                 drop_p=ph, drop_seed=seed + 200 + l)
        L["h1_32"], L["h1_16"] = h32, h16
        L["z"] = ops.gemm_f16(h16, lw["w1"], lw["b1"], 0, F16).contiguous()
        L["g"], _ = ops.gelu_fwd(L["z"], F16)
        if plan is not None and plan.p_act > 0:
            ops.dropout_(L["g"], plan.p_act, seed + 400 + l)             # HF:568
        L["f2"] = ops.gemm_f16(L["g"], lw["w2"], None, 0, F32).contiguous()
        h32, h16 = ops.layernorm(L["f2"], lw["ln2_g"], lw["ln2_b"], a.eps, bias=lw["b2"], residual=L["h1_32"],
                                 drop_p=ph, drop_seed=seed + 300 + l)                      # HF:572-574
        S["layers"].append(L)
    return h32.view(B, T, H), S


class GradBook:
    """Flat fp32 gradient buffer with named views (q/k/v of a layer are adjacent so the fused QKV
    weight gradient lands in place)."""

    def __init__(self, shapes: Dict[str, torch.Size], order: List[str], device, flat: Optional[torch.Tensor] = None):
        self.offsets = {}
        n = 0
        for k in order:
            self.offsets[k] = n
            n += int(torch.Size(shapes[k]).numel())
        self.numel = n
        # `flat` given: accumulate straight into the caller's (pre-zeroed) buffer, e.g. the trainer's
        # flat gradient -- no copy, no autograd accumulation pass
        self.flat = torch.zeros(n, dtype=F32, device=device) if flat is None else flat
        assert self.flat.numel() == n
        self.shapes = shapes

    def view(self, k: str) -> torch.Tensor:
        o = self.offsets[k]
        return self.flat[o:o + int(torch.Size(self.shapes[k]).numel())].view(self.shapes[k])

    def span(self, first: str, rows: int, cols: int) -> torch.Tensor:
        o = self.offsets[first]
        return self.flat[o:o + rows * cols].view(rows, cols)


def encoder_grad_order(arch: ArchConfig) -> List[str]:
    order = ["masked_spec_embed", "feature_projection.layer_norm.weight", "feature_projection.layer_norm.bias",
             "feature_projection.projection.weight", "feature_projection.projection.bias",
             "encoder.pos_conv_embed.conv.bias", "encoder.pos_conv_embed.conv.parametrizations.weight.original0",
             "encoder.pos_conv_embed.conv.parametrizations.weight.original1",
             "encoder.layer_norm.weight", "encoder.layer_norm.bias"]
    for l in range(arch.layers):
        pre = f"encoder.layers.{l}."
        order += [pre + "attention.q_proj.weight", pre + "attention.k_proj.weight", pre + "attention.v_proj.weight",
                  pre + "attention.q_proj.bias", pre + "attention.k_proj.bias", pre + "attention.v_proj.bias",
                  pre + "attention.out_proj.weight", pre + "attention.out_proj.bias",
                  pre + "layer_norm.weight", pre + "layer_norm.bias",
                  pre + "feed_forward.intermediate_dense.weight", pre + "feed_forward.intermediate_dense.bias",
                  pre + "feed_forward.output_dense.weight", pre + "feed_forward.output_dense.bias",
                  pre + "final_layer_norm.weight", pre + "final_layer_norm.bias"]
    return order


def encoder_backward(eng: EncoderEngine, tw: TrainWeights, params: Dict[str, torch.Tensor], S: dict,
                     dh: torch.Tensor, sink: Optional[GradBook] = None, on_layer_done=None) -> GradBook:
    """dh: f32 [B,T,H] gradient of last_hidden_state, carrying LOSS_SCALE.  Returns the (still scaled)
    parameter gradients of everything behind the frozen CNN.  `on_layer_done(lo, hi)` is called (in
    reverse layer order, also for LayerDrop-skipped layers) once flat[lo:hi] -- all gradients of one
    transformer layer -- has been enqueued: the trainer overlaps that span's all-reduce with the rest."""
    a, w = eng.arch, eng.w
    B, T = S["B"], S["T"]
    H, M, FF = a.hidden, B * T, a.ffn
    dev = dh.device
    plan = S.get("plan")
    ph = plan.p_hidden if plan is not None else 0.0
    seed = plan.seed if plan is not None else 0
    order = encoder_grad_order(a)
    G = sink if sink is not None else GradBook({k: params[k].shape for k in order}, order, dev)
    d = H // a.heads
    qscale = float(d) ** -0.5
    dy_a, dy_b = dh.contiguous().view(M, H), None

    def layer_done(l):
        if on_layer_done is not None:
            lo = G.offsets[f"encoder.layers.{l}.attention.q_proj.weight"]
            hi = G.offsets[f"encoder.layers.{l + 1}.attention.q_proj.weight"] if l + 1 < a.layers else G.numel
            on_layer_done(lo, hi)

    for l in reversed(range(a.layers)):
        L = S["layers"][l]
        if L is None:                                   # LayerDrop: identity in forward, identity in backward
            layer_done(l)
            continue
        pre = f"encoder.layers.{l}."
        lw, tl = w.layers[l], tw.layers[l]
        # LN2:  h2 = LN(drop(f2 + b2) + h1); dx2_16 is the gradient of the dropped branch, the residual keeps dx2_32
        dx2_32, dx2_16 = ops.layernorm_bwd(dy_a, L["f2"], lw["ln2_g"], a.eps, dy_b=dy_b, bias=lw["b2"],
                                           residual=L["h1_32"], dgamma=G.view(pre + "final_layer_norm.weight"),
                                           dbeta=G.view(pre + "final_layer_norm.bias"), drop_p=ph, drop_seed=seed + 300 + l,
                                           dbias=G.view(pre + "feed_forward.output_dense.bias"))
        ops.gemm_wgrad_f16(dx2_16, L["g"], G.view(pre + "feed_forward.output_dense.weight"))
        dg16 = ops.gemm_f16(dx2_16, tl["w2T"], None, 0, F16).contiguous()       # [M, FF]
        if plan is not None and plan.p_act > 0:
            ops.dropout_(dg16, plan.p_act, seed + 400 + l)
        dz16 = ops.gelu_bwd(dg16, L["z"], dbias=G.view(pre + "feed_forward.intermediate_dense.bias"))
        ops.gemm_wgrad_f16(dz16, L["h1_16"], G.view(pre + "feed_forward.intermediate_dense.weight"))
        dh1_a = ops.gemm_f16(dz16, tl["w1T"], None, 0, F32)                      # [M, H]
        # LN1:  h1 = LN(drop(o + bo) + h_in)
        dx1_32, dx1_16 = ops.layernorm_bwd(dh1_a, L["o"], lw["ln1_g"], a.eps, dy_b=dx2_32, bias=lw["bo"],
                                           residual=L["h_in32"], dgamma=G.view(pre + "layer_norm.weight"),
                                           dbeta=G.view(pre + "layer_norm.bias"), drop_p=ph, drop_seed=seed + 200 + l,
                                           dbias=G.view(pre + "attention.out_proj.bias"))
        ops.gemm_wgrad_f16(dx1_16, L["att"], G.view(pre + "attention.out_proj.weight"))
        datt16 = ops.gemm_f16(dx1_16, tl["woT"], None, 0, F16)
        dqkv16 = ops.attention_bwd(L["qkv"], L["att"], datt16, L["lse"], B, T, H, a.heads,
                                   drop_p=plan.p_attn if plan is not None else 0.0, drop_seed=seed + 100 + l)
        ops.colsum(dqkv16, G.span(pre + "attention.q_proj.bias", 1, 3 * H).view(3 * H))
        ops.gemm_wgrad_f16(dqkv16, L["h_in16"], G.span(pre + "attention.q_proj.weight", 3 * H, H))
        # the q projection was used pre-scaled by d^-0.5: chain rule for the unscaled parameters
        ops.scale_f32_(G.view(pre + "attention.q_proj.weight"), qscale)
        ops.scale_f32_(G.view(pre + "attention.q_proj.bias"), qscale)
        dy_a = ops.gemm_f16(dqkv16, tl["wqkvT"], None, 0, F32)                   # d h_in via qkv
        dy_b = dx1_32                                                            # + residual path
        layer_done(l)
    # encoder top:  h_e = drop(LN(pos + h0)),  pos = GELU(zpos),  zpos = posconv(h0) + b
    if ph > 0:
        dy_a, _ = ops.add2_cast(dy_a, dy_b, want16=False)
        dy_b = None
        ops.dropout_(dy_a, ph, seed + 2)
    dxe32, dxe16 = ops.layernorm_bwd(dy_a, S["pos"], w.enc_ln_g, a.eps, dy_b=dy_b, residual=S["h0"],
                                     dgamma=G.view("encoder.layer_norm.weight"), dbeta=G.view("encoder.layer_norm.bias"))
    dz16 = ops.gelu_bwd(dxe16, S["zpos16"], dbias=G.view("encoder.pos_conv_embed.conv.bias"))
    dx_pos = ops.posconv_ex(dz16.view(B, T, H), tw.pos_dgrad_w(T), None, a.pos_groups, a.pos_kernel, 0, 1)
    # positional-conv weight gradient (shifted-slab tensor-core kernel, no im2col), then weight-norm backward
    I = H // a.pos_groups
    K = a.pos_kernel
    dw_hki = torch.zeros(H, K * I, dtype=F32, device=dev)
    ops.posconv_wgrad(dz16.view(B, T, H), S["x16"].view(B, T, H), a.pos_groups, K, dw_hki)
    ops.weight_norm_bwd(dw_hki, w._pos_v, w._pos_g, 1.0,
                        G.view("encoder.pos_conv_embed.conv.parametrizations.weight.original1"),
                        G.view("encoder.pos_conv_embed.conv.parametrizations.weight.original0").view(-1))
    # feature projection:  h0 = timemask(drop(LN512(feat) Wp^T + bp))
    dh0_32, dh0_16 = ops.add2_cast(dxe32, dx_pos.view(M, H), want32=True, want16=True)
    if plan is not None and (plan.mask is not None or plan.p_feat > 0):
        if plan.mask is not None:
            ops.time_mask_bwd_(dh0_32, plan.mask, G.view("masked_spec_embed"))
        if plan.p_feat > 0:
            ops.dropout_(dh0_32, plan.p_feat, seed + 1)
        dh0_16 = ops.cast_f16(dh0_32)
    ops.colsum(dh0_16, G.view("feature_projection.projection.bias"))
    ops.gemm_wgrad_f16(dh0_16, S["n16"], G.view("feature_projection.projection.weight"))
    dn32 = ops.gemm_f16(dh0_16, tw.fp_wT, None, 0, F32)                           # [M, 512]
    ops.layernorm_bwd(dn32, S["feat"], w.fp_ln_g, a.eps, dgamma=G.view("feature_projection.layer_norm.weight"),
                      dbeta=G.view("feature_projection.layer_norm.bias"), want32=False, want16=False)
    return G


# ---------------------------------------------------------------------------------------------------
# autograd glue


class EncoderFn(torch.autograd.Function):
    """last_hidden_state = encoder(wav; parameters)."""

    @staticmethod
    def forward(ctx, wav, model, names, *params):
        eng = model._engine()
        plan = model._draw_reg_plan(wav, eng)
        out, saved = encoder_forward_train(eng, wav, plan, model.masked_spec_embed.detach())
        ctx.model, ctx.names, ctx.saved, ctx.eng = model, names, saved, eng
        return out

    @staticmethod
    def backward(ctx, dh):
        model, names, eng = ctx.model, ctx.names, ctx.eng
        pd = dict(model.named_parameters())
        tw = model._train_weights(eng)
        sink = getattr(model, "_grad_sink", None)
        G = encoder_backward(eng, tw, pd, ctx.saved, dh.float(), sink,
                             getattr(model, "_grad_ready_hook", None) if sink is not None else None)
        ctx.saved = None
        if sink is not None:
            # the trainer owns the (loss-scaled) flat gradient: nothing goes back through autograd
            return (None, None, None, *([None] * len(names)))
        ops.scale_f32_(G.flat, 1.0 / LOSS_SCALE)
        grads = []
        for n in names:
            grads.append(G.view(n) if (n in G.offsets and pd[n].requires_grad) else None)
        return (None, None, None, *grads)


class MeanPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.T = x.shape[1]
        return ops.stat_pool(x.contiguous(), 0)

    @staticmethod
    def backward(ctx, demb):
        return ops.mean_pool_bwd(demb.float().contiguous(), ctx.T)


class SpeakerLinearFn(torch.autograd.Function):
    """logits = x W^T + b with error-compensated fp16 operands (forward) and plain fp16 gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_split):
        xa = ops.split3_rows(x.detach().float().contiguous(), 0)
        out = ops.gemm_f16(xa, w_split, bias.detach().float() if bias is not None else None, 0, F32)
        ctx.save_for_backward(x.detach(), weight.detach())
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dlogits):
        x, W = ctx.saved_tensors
        Bn, S = dlogits.shape
        E = W.shape[1]
        ld = (S + 63) // 64 * 64
        dl16 = ops.cast_f16_rows(dlogits.float(), ld)                          # already carries LOSS_SCALE
        x16 = ops.cast_f16(x.float().contiguous())
        dW = torch.zeros(S, E, dtype=F32, device=W.device)
        ops.gemm_wgrad_f16(dl16[:, :S], x16, dW)
        ops.scale_f32_(dW, 1.0 / LOSS_SCALE)
        db = None
        if ctx.has_bias:
            db = torch.zeros(S, dtype=F32, device=W.device)
            ops.colsum(dl16[:, :S], db, 1.0 / LOSS_SCALE)
        wT = ops.cast_f16_transpose(W.float(), ld)                              # [E, ld]
        dx = ops.gemm_f16(dl16, wT, None, 0, F32)                               # scaled, flows on
        return dx, dW, db, None


class CrossEntropyFn(torch.autograd.Function):
    """(loss, softmax) = CE(logits, labels); the gradient that leaves this node carries LOSS_SCALE."""

    @staticmethod
    def forward(ctx, logits, labels):
        logits = logits.float()
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        prob, loss_rows, _ = ops.softmax_ce(logits, labels)
        ctx.save_for_backward(prob, labels)
        ctx.mark_non_differentiable(prob)
        return ops.mean_rows(loss_rows), prob

    @staticmethod
    def backward(ctx, dloss, _dprob):
        prob, labels = ctx.saved_tensors
        dl = ops.softmax_ce_bwd_f32(prob, labels, dloss.float().contiguous().view(1), LOSS_SCALE / prob.shape[0])
        return dl, None


class MeanStdPoolFn(torch.autograd.Function):
    """[std (unbiased) || mean] over time (R:src/layers/pooling.py:38-44)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.stat_pool(x, 1)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        return ops.meanstd_pool_bwd(x, dout.float())


class AamSoftmaxFn(torch.autograd.Function):
    """(loss, softmax) of AAM-softmax (R:src/optim/loss/aam_softmax.py:50-74); the gradient that leaves this
    node towards the embedding carries LOSS_SCALE, the classifier gradient is unscaled."""

    @staticmethod
    def forward(ctx, x, fc_weights, labels, margin, scale, easy_margin, w_split):
        x = x.detach().float().contiguous()
        xa = ops.l2norm_rows_split3(x, 0)
        cosine = ops.gemm_f16(xa, w_split, None, 0, F32)
        prob, loss_rows, _, cos_label = ops.aam_softmax_ce_train(cosine, labels, margin, scale, easy_margin)
        ctx.save_for_backward(x, fc_weights.detach(), labels, prob, cos_label)
        ctx.cfg = (float(margin), float(scale), bool(easy_margin))
        ctx.mark_non_differentiable(prob)
        return ops.mean_rows(loss_rows), prob

    @staticmethod
    def backward(ctx, dloss, _dprob):
        x, W, labels, prob, cos_label = ctx.saved_tensors
        margin, scale, easy = ctx.cfg
        Bn, S = prob.shape
        E = W.shape[1]
        ld = (S + 63) // 64 * 64
        dc16 = ops.aam_bwd_dcos(prob, cos_label, labels, dloss.float().contiguous().view(1), LOSS_SCALE / Bn, margin,
                                scale, easy, ld)                                  # [B, ld], loss-scaled
        Wf = W.float().contiguous()
        inv_w = ops.row_inv_norm(Wf)
        whT = ops.cast_f16_transpose(Wf, ld, inv_w)                               # normalised W, transposed [E, ld]
        dxh = ops.gemm_f16(dc16, whT, None, 0, F32)                               # d(x_hat) [B, E]
        dx = ops.l2norm_rows_bwd(x, dxh, 1.0)                                     # stays loss-scaled
        xh16 = ops.l2norm_rows_f16(x)
        dwh = torch.zeros(S, E, dtype=F32, device=W.device)
        ops.gemm_wgrad_f16(dc16[:, :S], xh16, dwh)                                # d(W_hat) [S, E]
        dW = ops.l2norm_rows_bwd(Wf, dwh, 1.0 / LOSS_SCALE)
        return dx, dW, None, None, None, None, None
