"""Training path: forward that keeps what the backward needs, the hand-scheduled backward over the
sm_100a kernels, and the torch.autograd.Function wrappers that make ``loss.backward()`` work on the
reference-facing modules (the reference trains through Lightning's automatic optimisation,
R:src/lightning_modules/speaker/speaker_recognition_module.py:148-220).

Scope: wav2vec2-base/large encoder with the CNN feature extractor frozen (the reference default
``completely_freeze_feature_extractor: true``, R:config/network/wav2vec2_fc.yaml:16) or trained, the
reference's train-mode regularisation (dropout / LayerDrop / SpecAugment), mean / mean+std / attentive
pooling, CE and AAM-softmax heads.

Gradient convention: every gradient that crosses a Function boundary -- in or out, activation or parameter -- is a
plain UNSCALED fp32 tensor, so these Functions compose with any torch op, head or loss around them (the paired-input
model scores its CLS token with an ``nn.Linear`` and torch's BCE).  INSIDE a backward the activation gradients are
multiplied by ``LOSS_SCALE`` before they are rounded to the fp16 operands of the dgrad / wgrad GEMMs and the results
are divided by it again (one extra pass over [B, T, H] per step, ~10 us).  The one exception is the trainer's gradient
sink (``model._grad_sink``): what is accumulated there stays loss-scaled and Adam's gradient scale undoes it.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from . import schedule as sched
from .engine import ArchConfig, EncoderEngine, PreparedWeights

F16, F32 = torch.float16, torch.float32

_PLAN_EARLY = __import__("os").environ.get("W2V2_PLAN_EARLY", "0") == "1"      # A/B: draw the regularisation before the CNN

LOSS_SCALE = 4096.0          # static scale of the activation gradients (fp16 operands of dgrad / wgrad)


def _entry(g: torch.Tensor, static: bool = False, hi: float = 2.0 ** 10, mid: float = 1.0):
    """Incoming gradient of a Function -> (fp32 working copy times LOSS_SCALE [times k], inv).  k is a power of two chosen
    ON THE DEVICE (no synchronisation) that is 1 for gradients of ordinary magnitude and brings the copy back into fp16
    range when something outside multiplied the loss -- Lightning's native-AMP GradScaler (`precision: 16`, the paper
    setting, R:config/experiment/speaker_wav2vec2_aam.yaml:17) starts at 2^16.  inv (device scalar 1/k, or None when
    `static`: the trainer's flat gradient keeps the plain LOSS_SCALE that its Adam undoes) goes to `_leave_`."""
    g = g.float().contiguous()
    if static:
        return ops.scaled_copy_f32(g, LOSS_SCALE), None
    # window of amax|g| * LOSS_SCALE inside which nothing changes: everything the static scale was validated on lies in it
    return ops.grad_entry_scale(g, LOSS_SCALE, 2.0 ** -10, hi, mid)


def _leave_(t: torch.Tensor, inv, s: float = 1.0 / LOSS_SCALE) -> torch.Tensor:
    """Outgoing gradient: undo the loss scale (and the entry normalisation) in place."""
    if t is None:
        return t
    return ops.scale_f32_(t, s) if inv is None else ops.scale_f32_dev_(t, s, inv)


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
            # [H, 3H]; the q block UNSCALED: the attention backward already multiplies dq by d^-0.5 (chain rule of the
            # folded scale), so dX = dq_scaled Wq and dWq / dbq need no further correction
            for i, n in enumerate("qkv"):
                prep.add_transposed(pre + f"attention.{n}_proj.weight", L["wqkvT"][:, i * H:(i + 1) * H], scale=1.0)
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

    def __init__(self, reg, layers: int, B: int, T: int, rng, device, pinned: Optional[torch.Tensor] = None,
                 hidden: int = 0):
        self.p_feat = float(reg.feat_proj_dropout)
        self.p_hidden = float(reg.hidden_dropout)
        self.p_attn = float(reg.attention_dropout)
        self.p_act = float(reg.activation_dropout)
        self.seed = int(rng.integers(1, 1 << 62))
        self.skip = [bool(rng.random() < reg.layerdrop) for _ in range(layers)]
        self.mask = None
        if reg.mask_time_prob > 0:
            m = torch.from_numpy(compute_time_mask(B, T, reg.mask_time_prob, reg.mask_time_length, 2, rng))
            if pinned is not None:               # staging buffer owned by the model: asynchronous upload
                pinned.copy_(m)
                self.mask = pinned.to(device, non_blocking=True)
            else:
                self.mask = m.to(device)
        # SpecAugment along the feature axis (HF:1312-1322; off in every reference configuration): spans of
        # `mask_feature_length` hidden units per utterance, drawn like the time spans over the [B, hidden] grid, no minimum
        # (`hidden` = 0: the split call path, where HF applies no SpecAugment at all)
        self.fmask = None
        if getattr(reg, "mask_feature_prob", 0.0) > 0 and hidden > 0:
            f = compute_time_mask(B, hidden, reg.mask_feature_prob, reg.mask_feature_length, 0, rng)
            self.fmask = torch.from_numpy(f).to(device)

    @property
    def any(self) -> bool:
        return ((self.p_feat + self.p_hidden + self.p_attn + self.p_act) > 0 or any(self.skip) or self.mask is not None
                or self.fmask is not None)


def compute_time_mask(B: int, T: int, mask_prob: float, mask_length: int, min_masks: int, rng):
    """uint8 [B*T]; own restatement of HF `_compute_mask_indices` (HF:101-217) for full-length inputs:
    num_spans = max(min_masks, int(mask_prob*T/mask_length + U[0,1))), clipped so the spans fit; span starts
    drawn without replacement from [0, T - mask_length].  Vectorised over the batch (no per-utterance python loop: the
    draw sits on the host's critical path of a step): the first n_b entries of a random permutation of the admissible
    starts are n_b starts drawn uniformly without replacement."""
    import numpy as np
    mask = np.zeros((B, T), dtype=np.uint8)
    if mask_length < 1 or mask_length > T:
        return mask.reshape(-1)
    n = (mask_prob * T / mask_length + rng.random(B)).astype(np.int64)
    n = np.maximum(n, min_masks)
    n = np.where(n * mask_length > T, T // mask_length, n)
    nmax = int(n.max())
    if nmax == 0:
        return mask.reshape(-1)
    starts = np.argsort(rng.random((B, T - (mask_length - 1))), axis=1)[:, :nmax]        # [B, nmax] distinct per row
    live = np.arange(nmax)[None, :] < n[:, None]
    rows = np.repeat(np.arange(B), nmax).reshape(B, nmax)
    for off in range(mask_length):
        mask[rows[live], starts[live] + off] = 1
    return mask.reshape(-1)


def cnn_forward_train(eng: EncoderEngine, wav: torch.Tensor):
    """Feature extractor (HF:409-419) keeping what its backward needs: the pre-GELU activation of every layer
    (the GELUs run as separate passes) and the GroupNorm affine / im2col operand of layer 0.
    -> (features f32 [B, T, C], saved)."""
    a, w = eng.arch, eng.w
    y0, ws, scale, im2col = ops.conv0_gn(wav, w.conv0_w, w.gn_g, w.gn_b, a.eps, act=0)      # GroupNorm output, f16
    h, _ = ops.gelu_fwd(y0, F16)
    saved = dict(ws=ws, scale=scale, im2col=im2col, zs=[y0], outs=[h])
    n = len(a.conv_kernel)
    for i in range(1, n):
        last = i == n - 1
        z = ops.conv1d_cl_f16(h, w.conv_w[i], a.conv_kernel[i], a.conv_stride[i], act=0, out_dtype=F32 if last else F16)
        if last:
            h, z16 = ops.gelu_fwd(z, F32, want_x16=True)          # f32 features (they feed a LayerNorm), f16 copy of z
        else:
            h, z16 = ops.gelu_fwd(z, F16)[0], z
        saved["zs"].append(z16)
        saved["outs"].append(h)
    return h, saved


def cnn_dgrad_weights(params: Dict[str, torch.Tensor], arch: ArchConfig):
    """fp16 operands of the data-gradient tap-GEMMs of conv layers 1..6 (stride 2): for the even input rows
    [W_0^T | W_2^T] (k = 3) or W_0^T (k = 2), for the odd rows W_1^T; W_j = weight[:, :, j] ([Cout, Cin]).
    A re-layout of 6 x 0.8 M parameters per optimizer step (torch indexing: plumbing, no arithmetic)."""
    out = [None]
    for i in range(1, len(arch.conv_kernel)):
        wt = params[f"feature_extractor.conv_layers.{i}.conv.weight"].detach()      # [Cout, Cin, k]
        k = arch.conv_kernel[i]
        assert arch.conv_stride[i] == 2 and k in (2, 3), "data gradient built for the stride-2, k in {2, 3} layers"
        t = [wt[:, :, j].t() for j in range(k)]                                      # [Cin, Cout] each
        even = torch.cat([t[0], t[2]], dim=1) if k == 3 else t[0]
        out.append((even.contiguous().to(F16), t[1].contiguous().to(F16)))
    return out


def cnn_backward(eng: EncoderEngine, params: Dict[str, torch.Tensor], S: dict, dfeat: torch.Tensor) -> Dict[str, torch.Tensor]:
    """dfeat: f32 [B, T, C] gradient of the extracted features (carrying LOSS_SCALE).  Returns the UNSCALED
    gradients of the feature-extractor parameters under their HF names."""
    a, w = eng.arch, eng.w
    C = a.conv_dim
    dev = dfeat.device
    inv_ls = 1.0 / LOSS_SCALE
    grads: Dict[str, torch.Tensor] = {}
    wd = cnn_dgrad_weights(params, a)
    d = ops.cast_f16(dfeat.contiguous())                                            # d out_6
    B = d.shape[0]
    for i in reversed(range(1, len(a.conv_kernel))):
        k, s = a.conv_kernel[i], a.conv_stride[i]
        dz = ops.gelu_bwd(d, S["zs"][i])                                            # [B, Lout, C]
        x_in = S["outs"][i - 1]                                                     # [B, Lin, C] f16
        Lin, Lout = x_in.shape[1], dz.shape[1]
        # weight gradient, tap by tap, into the tap-major [Cout, k*Cin] layout
        dwt = torch.zeros(C, k * C, dtype=F32, device=dev)
        for j in range(k):
            ops.gemm_wgrad_f16_batched(dz, x_in[:, j:j + s * (Lout - 1) + 1:s, :], dwt[:, j * C:(j + 1) * C])
        ops.scale_f32_(dwt, inv_ls)
        grads[f"feature_extractor.conv_layers.{i}.conv.weight"] = dwt.view(C, k, C).permute(0, 2, 1).contiguous()
        # data gradient: input row u receives dz[(u - j) / 2] W_j for every tap j of u's parity
        d_in = torch.empty(B, Lin, C, dtype=F16, device=dev)
        ops.gemm_f16_taps(dz, [0, -1] if k == 3 else [0], wd[i][0], d_in[:, 0::2, :])
        ops.gemm_f16_taps(dz, [0], wd[i][1], d_in[:, 1::2, :])
        d = d_in
    # layer 0: GELU, GroupNorm, conv (no data gradient: the input is the waveform)
    dy = ops.gelu_bwd(d, S["zs"][0])
    dgam = torch.zeros(C, dtype=F32, device=dev)
    dbet = torch.zeros(C, dtype=F32, device=dev)
    dc = ops.groupnorm_bwd(dy, S["zs"][0], w.gn_g, w.gn_b, S["scale"], dgam, dbet, inv_ls)
    o = torch.zeros(C, 64, dtype=F32, device=dev)
    ops.gemm_wgrad_f16_batched(dc, S["im2col"], o)                                  # columns [x_hi | x_lo | x_hi | 0]
    ops.scale_f32_(o, inv_ls)
    grads["feature_extractor.conv_layers.0.conv.weight"] = (o[:, :10] + o[:, 10:20]).reshape(C, 1, 10)
    grads["feature_extractor.conv_layers.0.layer_norm.weight"] = dgam
    grads["feature_extractor.conv_layers.0.layer_norm.bias"] = dbet
    return grads


def encoder_forward_train(eng: EncoderEngine, wav: torch.Tensor, plan: Optional[RegPlan] = None,
                          mask_embed: Optional[torch.Tensor] = None, train_cnn: bool = False, pre_encoder_hook=None,
                          normalize: bool = False):
    """Same arithmetic as EncoderEngine.forward (the GELUs run as separate passes so the pre-activations
    can be kept) plus the train-mode regularisation of `plan`; returns (last_hidden_state f32 [B,T,H], saved).
    `plan` may be a zero-argument callable: it is then drawn AFTER the feature extractor's kernels have been enqueued, so
    the host-side draw (seeds, LayerDrop, SpecAugment spans: ~0.3 ms) runs under ~1 ms of CNN work instead of in front of
    the step's first launch -- what a loop that synchronises every step (bench.py's e2e) would otherwise wait for."""
    S = {}
    if train_cnn:
        feat, S["cnn"] = cnn_forward_train(eng, wav)          # unfrozen CNN: pre-activations kept
    else:
        feat = eng.feature_extractor(wav, None, None, normalize)   # frozen CNN: nothing saved from inside (raw input: the
                                                                   # normaliser is folded into conv layer 0)
    if callable(plan):
        plan = plan()
    S["plan"] = plan
    if pre_encoder_hook is not None:
        pre_encoder_hook()                                    # e.g. join the optimizer stream (trainer.py)
    B, T, C = feat.shape
    feat2 = feat.contiguous().view(B * T, C)
    h0, n16 = projection_forward_train(eng, feat2, plan)
    if plan is not None and plan.mask is not None:
        ops.time_mask_apply_(h0, plan.mask, mask_embed)                  # HF:1301-1310
    if plan is not None and plan.fmask is not None:
        ops.feature_mask_(h0, plan.fmask, B, T)                          # HF:1312-1322
    S.update(feat=feat2, n16=n16)
    return stack_forward_train(eng, h0, B, T, plan, S), S


def projection_forward_train(eng: EncoderEngine, feat2: torch.Tensor, plan: Optional[RegPlan]):
    """Feature projection (HF:429-434): LayerNorm(C) -> Linear C -> H -> dropout.  feat2: f32 [M, C].
    -> (h0 f32 [M, H] (a fresh tensor), the fp16 LayerNorm output the weight gradient needs)."""
    a, w = eng.arch, eng.w
    _, n16 = ops.layernorm(feat2, w.fp_ln_g, w.fp_ln_b, a.eps, want32=False)
    h0 = ops.gemm_f16(n16, w.fp_w, w.fp_b, 0, F32).contiguous()          # [M,H]
    if plan is not None and plan.p_feat > 0:
        ops.dropout_(h0, plan.p_feat, plan.seed + 1)                     # HF:433
    return h0, n16


def stack_forward_train(eng: EncoderEngine, h0: torch.Tensor, B: int, T: int, plan: Optional[RegPlan], S: dict):
    """The transformer stack (HF:668-727) on an arbitrary sequence h0 (f32 [B*T, H], contiguous, left untouched):
    positional conv + GELU, LayerNorm, dropout, the encoder layers.  Fills `S` with what stack_backward needs;
    -> last_hidden_state f32 [B, T, H]."""
    a, w = eng.arch, eng.w
    if a.stable_layer_norm:                      # -lv60 / XLSR checkpoints: pre-LN layers (training_stable.py)
        from .training_stable import stack_forward_train_stable
        return stack_forward_train_stable(eng, h0, B, T, plan, S)
    H, M = a.hidden, B * T
    ph = plan.p_hidden if plan is not None else 0.0
    seed = plan.seed if plan is not None else 0
    x16 = ops.cast_f16(h0)
    zpos = ops.posconv_ex(x16.view(B, T, H), w.pos_w(T), w.pos_b, a.pos_groups, a.pos_kernel, 0, 0)
    pos, zpos16 = ops.gelu_fwd(zpos.view(M, H), F32, want_x16=True)
    h32, h16 = ops.layernorm(pos, w.enc_ln_g, w.enc_ln_b, a.eps, residual=h0)
    if ph > 0:
        h32, h16 = ops.dropout_(h32, ph, seed + 2, want16=True)          # HF:693
    S.update(B=B, T=T, h0=h0, x16=x16, pos=pos, zpos16=zpos16, layers=[], top=(h32, h16))
    # transformer layers: one native schedule call per layer (csrc/schedule.cu), buffers from one arena each
    sizes = sched.layer_buffer_sizes(B, T, H, a.heads, a.ffn, True)
    p32, p16 = h32.data_ptr(), h16.data_ptr()
    last = None
    for l, lw in enumerate(w.layers):
        if plan is not None and plan.skip[l]:                            # LayerDrop (HF:701-713)
            S["layers"].append(None)
            continue
        ar = sched.Arena(sizes, h32.device)
        sched.run_layer_fwd(sched.fwd_args(a, B, T, l, lw, p32, p16, ar, True, ph,
                                           plan.p_attn if plan is not None else 0.0,
                                           plan.p_act if plan is not None else 0.0, seed))
        S["layers"].append(dict(arena=ar, h_in32=p32, h_in16=p16))
        p32, p16 = ar.ptr("h2_32"), ar.ptr("h2_16")
        last = ar
    return last.tensor("h2_32", F32, (B, T, H)) if last is not None else h32.view(B, T, H)


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


PROJECTION_GRAD_ORDER = ["feature_projection.layer_norm.weight", "feature_projection.layer_norm.bias",
                         "feature_projection.projection.weight", "feature_projection.projection.bias"]


def encoder_grad_order(arch: ArchConfig) -> List[str]:
    """Layout of the flat gradient of everything behind the CNN (the trainer's buffer follows it)."""
    return ["masked_spec_embed"] + PROJECTION_GRAD_ORDER + stack_grad_order(arch)


def stack_grad_order(arch: ArchConfig) -> List[str]:
    order = ["encoder.pos_conv_embed.conv.bias", "encoder.pos_conv_embed.conv.parametrizations.weight.original0",
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
    """dh: f32 [B,T,H] gradient of last_hidden_state, already multiplied by LOSS_SCALE.  Returns the (still scaled)
    parameter gradients of everything behind the frozen CNN.  `on_layer_done(lo, hi)` is called (in
    reverse layer order, also for LayerDrop-skipped layers) once flat[lo:hi] -- all gradients of one
    transformer layer -- has been enqueued: the trainer overlaps that span's all-reduce with the rest."""
    plan = S.get("plan")
    order = encoder_grad_order(eng.arch)
    G = sink if sink is not None else GradBook({k: params[k].shape for k in order}, order, dh.device)
    dxe32, dx_pos = stack_backward(eng, tw, S, dh, G, on_layer_done)
    # feature projection:  h0 = timemask(drop(LN512(feat) Wp^T + bp))
    dh0_32, dh0_16 = ops.add2_cast(dxe32, dx_pos, want32=True, want16=True)
    if plan is not None and (plan.mask is not None or plan.fmask is not None or plan.p_feat > 0):
        if plan.fmask is not None:
            ops.feature_mask_(dh0_32, plan.fmask, S["B"], S["T"])
        if plan.mask is not None:
            ops.time_mask_bwd_(dh0_32, plan.mask, G.view("masked_spec_embed"))
        if plan.p_feat > 0:
            ops.dropout_(dh0_32, plan.p_feat, plan.seed + 1)
        dh0_16 = ops.cast_f16(dh0_32)
    dfeat = projection_backward(eng, tw, S, dh0_16, G, want_dfeat="cnn" in S)
    if "cnn" in S:                                     # unfrozen feature extractor: its gradients are returned unscaled
        G.cnn_grads = cnn_backward(eng, params, S["cnn"], dfeat.view(S["B"], S["T"], -1))
    return G


def projection_backward(eng: EncoderEngine, tw: TrainWeights, S: dict, dh0_16: torch.Tensor, G: GradBook,
                        want_dfeat: bool):
    """dh0_16: f16 [M, H] gradient of the projection output BEFORE its dropout (mask already applied), loss-scaled.
    Accumulates the four projection gradients into G; -> d feat f32 [M, C] (loss-scaled) or None."""
    a, w = eng.arch, eng.w
    ops.colsum(dh0_16, G.view("feature_projection.projection.bias"))
    ops.gemm_wgrad_f16(dh0_16, S["n16"], G.view("feature_projection.projection.weight"))
    dn32 = ops.gemm_f16(dh0_16, tw.fp_wT, None, 0, F32)                           # [M, 512]
    dfeat, _ = ops.layernorm_bwd(dn32, S["feat"], w.fp_ln_g, a.eps, dgamma=G.view("feature_projection.layer_norm.weight"),
                                 dbeta=G.view("feature_projection.layer_norm.bias"), want32=want_dfeat, want16=False)
    return dfeat


def stack_backward(eng: EncoderEngine, tw: TrainWeights, S: dict, dh: torch.Tensor, G: GradBook, on_layer_done=None):
    """Backward of stack_forward_train.  dh: f32 [B,T,H], loss-scaled.  Accumulates the gradients of the positional
    conv, the encoder LayerNorm and every layer into G; -> the two (loss-scaled, f32 [M, H]) terms of d h0:
    through the LayerNorm residual and through the positional conv."""
    a, w = eng.arch, eng.w
    if a.stable_layer_norm:
        from .training_stable import stack_backward_stable
        return stack_backward_stable(eng, tw, S, dh, G, on_layer_done)
    B, T = S["B"], S["T"]
    H, M, FF = a.hidden, B * T, a.ffn
    dev = dh.device
    plan = S.get("plan")
    ph = plan.p_hidden if plan is not None else 0.0
    seed = plan.seed if plan is not None else 0
    d = H // a.heads
    qscale = float(d) ** -0.5
    dy_a, dy_b = dh.contiguous().view(M, H), None

    def layer_done(l):
        if on_layer_done is not None:
            lo = G.offsets[f"encoder.layers.{l}.attention.q_proj.weight"]
            hi = G.offsets[f"encoder.layers.{l + 1}.attention.q_proj.weight"] if l + 1 < a.layers else G.numel
            on_layer_done(lo, hi)

    # transformer layers in reverse: one native schedule call per layer; the scratch arena is shared by all
    # layers, the two terms of the input gradient ping-pong between two buffer pairs
    scr = sched.Arena(sched.bwd_scratch_sizes(B, T, H, FF), dev)
    # 1: the layer schedule adds the q/k/v data gradient into dx1_32 (w2v2_gemm_f16_accum): one term, not two
    accum = bool(ops._lib.load().w2v2_dgrad_accumulates())
    gbase = G.flat.data_ptr()
    pa, pb, pp = dy_a.data_ptr(), None, 0
    ran = False
    for l in reversed(range(a.layers)):
        L = S["layers"][l]
        if L is None:                                   # LayerDrop: identity in forward, identity in backward
            layer_done(l)
            continue
        pre = f"encoder.layers.{l}."
        lw, tl, ar = w.layers[l], tw.layers[l], L["arena"]
        g = lambda k: gbase + 4 * G.offsets[pre + k]
        b = sched.LayerBwdArgs()
        b.B, b.T, b.H, b.heads, b.FF, b.layer = B, T, H, a.heads, FF, l
        b.eps, b.p_hidden, b.qscale = a.eps, ph, qscale
        b.p_attn = plan.p_attn if plan is not None else 0.0
        b.p_act = plan.p_act if plan is not None else 0.0
        b.seed = seed
        b.wqkvT, b.woT, b.w1T, b.w2T = (tl[k].data_ptr() for k in ("wqkvT", "woT", "w1T", "w2T"))
        b.bo, b.b2, b.ln1_g, b.ln2_g = (lw[k].data_ptr() for k in ("bo", "b2", "ln1_g", "ln2_g"))
        b.h_in32, b.h_in16 = L["h_in32"], L["h_in16"]
        for k in ("qkv16", "att16", "lse", "o32", "h1_32", "h1_16", "z16", "g16", "f2_32", "h2_32", "rstd1", "rstd2"):
            setattr(b, k, ar.ptr(k))
        b.ln1_b, b.ln2_b = lw["ln1_b"].data_ptr(), lw["ln2_b"].data_ptr()
        b.dy_a, b.dy_b = pa, pb
        b.d_wqkv, b.d_bqkv = g("attention.q_proj.weight"), g("attention.q_proj.bias")
        b.d_wo, b.d_bo = g("attention.out_proj.weight"), g("attention.out_proj.bias")
        b.d_ln1_g, b.d_ln1_b = g("layer_norm.weight"), g("layer_norm.bias")
        b.d_w1, b.d_b1 = g("feed_forward.intermediate_dense.weight"), g("feed_forward.intermediate_dense.bias")
        b.d_w2, b.d_b2 = g("feed_forward.output_dense.weight"), g("feed_forward.output_dense.bias")
        b.d_ln2_g, b.d_ln2_b = g("final_layer_norm.weight"), g("final_layer_norm.bias")
        for k in ("dx2_32", "dx2_16", "dg16", "dz16", "dh1_32", "dx1_16", "datt16", "dqkv16"):
            setattr(b, k, scr.ptr(k))
        b.dx1_32, b.dh_in32 = scr.ptr(f"dx1_32.{pp}"), scr.ptr(f"dh_in32.{pp}")
        sched.run_layer_bwd(b)
        pa, pb = (b.dx1_32, None) if accum else (b.dh_in32, b.dx1_32)
        ran, last_pp = True, pp
        pp ^= 1
        layer_done(l)
    if ran and accum:
        dy_a, dy_b = scr.tensor(f"dx1_32.{last_pp}", F32, (M, H)), None          # residual path + q/k/v term, already summed
    elif ran:
        dy_a = scr.tensor(f"dh_in32.{last_pp}", F32, (M, H))                     # d h_in via qkv
        dy_b = scr.tensor(f"dx1_32.{last_pp}", F32, (M, H))                      # + residual path
    # encoder top:  h_e = drop(LN(pos + h0)),  pos = GELU(zpos),  zpos = posconv(h0) + b
    if ph > 0:
        if dy_b is not None or dy_a.data_ptr() == dh.data_ptr():         # (never drop in place in the caller's tensor)
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
    return dxe32, dx_pos.view(M, H)


# ---------------------------------------------------------------------------------------------------
# autograd glue


class EncoderFn(torch.autograd.Function):
    """last_hidden_state = encoder(wav; parameters)."""

    @staticmethod
    def forward(ctx, wav, model, names, *params):
        eng = model._engine()
        plan = lambda: model._draw_reg_plan(wav, eng)          # drawn once the CNN forward is in flight
        if _PLAN_EARLY:
            plan = plan()
        train_cnn = any(q.requires_grad for q in model._items()[3])
        raw = bool(getattr(model, "_raw_next", False))
        out, saved = encoder_forward_train(eng, wav, plan, model.masked_spec_embed.detach(), train_cnn,
                                           getattr(model, "_pre_encoder_hook", None), normalize=raw)
        ctx.model, ctx.names, ctx.saved, ctx.eng = model, names, saved, eng
        if getattr(model, "_keep_saved", False):      # output_hidden_states=True under grad (models/wav2vec2.py)
            model.__dict__["_last_saved"] = saved
        return out

    @staticmethod
    def backward(ctx, dh):
        model, names, eng = ctx.model, ctx.names, ctx.eng
        pd = model._items()[2]
        tw = model._train_weights(eng)
        sink = getattr(model, "_grad_sink", None)
        start_hook = getattr(model, "_backward_start_hook", None)
        if sink is not None and start_hook is not None:
            start_hook()                              # trainer: the heads' gradients are complete, send them now
        d, inv = _entry(dh, static=sink is not None)
        G = encoder_backward(eng, tw, pd, ctx.saved, d, sink,
                             getattr(model, "_grad_ready_hook", None) if sink is not None else None)
        ctx.saved = None
        cnn = getattr(G, "cnn_grads", None) or {}
        G.cnn_grads = None
        if inv is not None:
            for t in cnn.values():                # cnn_backward removed LOSS_SCALE itself; the entry normalisation remains
                _leave_(t, inv, 1.0)
        if sink is not None:
            # the trainer owns the (loss-scaled) flat gradient of everything behind the CNN: only the feature
            # extractor's (unscaled) gradients go back through autograd
            return (None, None, None, *[cnn.get(n) if pd[n].requires_grad else None for n in names])
        _leave_(G.flat, inv)
        grads = []
        for n in names:
            if not pd[n].requires_grad:
                grads.append(None)
            elif n in G.offsets:
                grads.append(G.view(n))
            else:
                grads.append(cnn.get(n))
        return (None, None, None, *grads)


# ---- the split call path: model.feature_extractor / model.feature_projection / model.encoder ---------------------
# The reference's CLS-token wrapper (R:src/models/wav2vec2.py:128-140) and its paired-input model
# (R:src/lightning_modules/speaker/wav2vec2_paired_input.py:162-207) call the three parts of the HF model one by one
# and build the encoder's input sequence themselves; each part is its own autograd node here, on the same kernels.


def _own_book(model, order: List[str], device):
    """(gradient book, owned): the trainer's flat buffer when one is installed, else a fresh one for `order`."""
    sink = getattr(model, "_grad_sink", None)
    if sink is not None:
        return sink, False
    pd = model._items()[2]
    return GradBook({k: pd[k].shape for k in order}, order, device), True


def _book_grads(model, G: GradBook, owned: bool, names: List[str], inv=None):
    """Gradients to hand back to autograd for `names`: views of an owned book (unscaled here), None when the
    trainer's sink holds them (it stays loss-scaled, Adam undoes the scale)."""
    if not owned:
        return [None] * len(names)
    _leave_(G.flat, inv)
    pd = model._items()[2]
    return [G.view(n) if pd[n].requires_grad else None for n in names]


class FeatureExtractorFn(torch.autograd.Function):
    """features f32 [B, T, C] = CNN(wav), with the CNN trained (a frozen one needs no autograd node)."""

    @staticmethod
    def forward(ctx, wav, model, names, *params):
        eng = model._engine()
        feat, saved = cnn_forward_train(eng, wav.float())
        ctx.model, ctx.names, ctx.saved, ctx.eng = model, names, saved, eng
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        model, names = ctx.model, ctx.names
        pd = model._items()[2]
        d, inv = _entry(dfeat)
        grads = cnn_backward(ctx.eng, pd, ctx.saved, d)
        ctx.saved = None
        for t in grads.values():
            _leave_(t, inv, 1.0)
        return (None, None, None, *[grads.get(n) if pd[n].requires_grad else None for n in names])


class FeatureProjectionFn(torch.autograd.Function):
    """hidden f32 [B, T, H] = dropout(Linear(LayerNorm(features))) (HF:429-434)."""

    @staticmethod
    def forward(ctx, feat, model, names, *params):
        eng = model._engine()
        hook = getattr(model, "_pre_encoder_hook", None)
        if hook is not None:
            hook()
        B, T, C = feat.shape
        plan = model._draw_split_plan(B, T, feat.device)
        feat2 = feat.detach().float().contiguous().view(B * T, C)
        h0, n16 = projection_forward_train(eng, feat2, plan)
        ctx.model, ctx.names, ctx.eng, ctx.plan = model, names, eng, plan
        ctx.saved = dict(feat=feat2, n16=n16, B=B, T=T)
        return h0.view(B, T, -1)

    @staticmethod
    def backward(ctx, dh):
        model, eng, plan, S = ctx.model, ctx.eng, ctx.plan, ctx.saved
        M = S["B"] * S["T"]
        G, owned = _own_book(model, PROJECTION_GRAD_ORDER, dh.device)
        d32, inv = _entry(dh, static=not owned)
        d32 = d32.view(M, -1)
        if plan is not None and plan.p_feat > 0:
            ops.dropout_(d32, plan.p_feat, plan.seed + 1)
        dfeat = projection_backward(eng, model._train_weights(eng), S, ops.cast_f16(d32), G, ctx.needs_input_grad[0])
        ctx.saved = None
        if dfeat is not None:
            dfeat = _leave_(dfeat, inv).view(S["B"], S["T"], -1)
        return (dfeat, None, None, *_book_grads(model, G, owned, ctx.names, inv))


class EncoderStackFn(torch.autograd.Function):
    """last_hidden_state f32 [B, T', H] = transformer stack(sequence) (HF:668-727) for a caller-built sequence."""

    @staticmethod
    def forward(ctx, seq, model, names, *params):
        eng = model._engine()
        hook = getattr(model, "_pre_encoder_hook", None)
        if hook is not None:
            hook()
        B, T, H = seq.shape
        plan = model._draw_split_plan(B, T, seq.device)
        h0 = seq.detach().float().contiguous().view(B * T, H)
        S = {"plan": plan}
        out = stack_forward_train(eng, h0, B, T, plan, S)
        ctx.model, ctx.names, ctx.eng, ctx.saved = model, names, eng, S
        return out

    @staticmethod
    def backward(ctx, dh):
        model, eng, S = ctx.model, ctx.eng, ctx.saved
        G, owned = _own_book(model, stack_grad_order(eng.arch), dh.device)
        hook = getattr(model, "_grad_ready_hook", None) if not owned else None
        start_hook = getattr(model, "_backward_start_hook", None)
        if not owned and start_hook is not None:
            start_hook()
        d, inv = _entry(dh, static=not owned)
        dxe32, dx_pos = stack_backward(eng, model._train_weights(eng), S, d, G, hook)
        ctx.saved = None
        dseq = None
        if ctx.needs_input_grad[0]:
            dseq, _ = ops.add2_cast(dxe32, dx_pos, want32=True, want16=False)
            dseq = _leave_(dseq, inv).view(S["B"], S["T"], -1)
        return (dseq, None, None, *_book_grads(model, G, owned, ctx.names, inv))


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
        # logit gradients are (softmax - onehot) / B: the small ones must stay fp16-normal, hence the higher window
        dl32, inv = _entry(dlogits, hi=2.0 ** 15, mid=2.0 ** 11)
        dl16 = ops.cast_f16_rows(dl32, ld, 1.0)
        x16 = ops.cast_f16(x.float().contiguous())
        dW = torch.zeros(S, E, dtype=F32, device=W.device)
        ops.gemm_wgrad_f16(dl16[:, :S], x16, dW)
        _leave_(dW, inv)
        db = None
        if ctx.has_bias:
            db = torch.zeros(S, dtype=F32, device=W.device)
            ops.colsum(dl16[:, :S], db, 1.0)
            _leave_(db, inv)
        wT = ops.cast_f16_transpose(W.float(), ld)                              # [E, ld]
        dx = _leave_(ops.gemm_f16(dl16, wT, None, 0, F32).contiguous(), inv)
        return dx, dW, db, None


class CrossEntropyFn(torch.autograd.Function):
    """(loss, softmax) = CE(logits, labels) (mean over the batch)."""

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
        dl = ops.softmax_ce_bwd_f32(prob, labels, dloss.float().contiguous().view(1), 1.0 / prob.shape[0])
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
    """(loss, softmax) of AAM-softmax (R:src/optim/loss/aam_softmax.py:50-74)."""

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
        # dcos = (softmax - onehot) * scale * dphi * dloss / B; the entry normalisation looks at scale * dloss / B * LOSS_SCALE
        dl, inv = ops.grad_entry_scale(dloss.float().contiguous().view(1), scale * LOSS_SCALE / Bn, 2.0 ** -8, 2.0 ** 15, 2.0 ** 12)
        dc16 = ops.aam_bwd_dcos(prob, cos_label, labels, dl, 1.0 / scale, margin, scale, easy, ld)  # [B, ld], loss-scaled
        Wf = W.float().contiguous()
        inv_w = ops.row_inv_norm(Wf)
        whT = ops.cast_f16_transpose(Wf, ld, inv_w)                               # normalised W, transposed [E, ld]
        dxh = ops.gemm_f16(dc16, whT, None, 0, F32)                               # d(x_hat) [B, E]
        dx = _leave_(ops.l2norm_rows_bwd(x, dxh, 1.0 / LOSS_SCALE), inv, 1.0)
        xh16 = ops.l2norm_rows_f16(x)
        dwh = torch.zeros(S, E, dtype=F32, device=W.device)
        ops.gemm_wgrad_f16(dc16[:, :S], xh16, dwh)                                # d(W_hat) [S, E]
        dW = _leave_(ops.l2norm_rows_bwd(Wf, dwh, 1.0 / LOSS_SCALE), inv, 1.0)
        return dx, dW, None, None, None, None, None


class AspPoolFn(torch.autograd.Function):
    """[mean || std] of attentive-statistics pooling (R:src/layers/pooling.py:87-106) with its backward.
    `layer` is the parameter holder (layers.pooling._AttentiveStatisticsPooling): training-mode BatchNorm
    uses (and updates) batch statistics exactly like torch's BatchNorm1d."""

    @staticmethod
    def forward(ctx, x, w1, b1, gamma, beta, w2, b2, layer):
        bn = layer.tdnn.norm.norm
        A = layer.attention_channels
        x = x.detach().float().contiguous()
        B, T, C = x.shape
        cat3 = ops.asp_concat_split3(x)                                              # [B*T, 9C] f16: [hi | lo | hi]
        cat16 = cat3[:, :3 * C]                                                      # the plain fp16 cat (row pitch 9C)
        w1f = w1.detach().float().reshape(A, 3 * C).contiguous()
        w2f = w2.detach().float().reshape(C, A).contiguous()
        # Conv1d k=1 with error-compensated operands: the ReLU that follows makes the TDNN gradient sensitive to z
        z = ops.gemm_f16(cat3, ops.split3_rows(w1f, 1), b1.detach().float(), 0, F32).contiguous()
        g, bt = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        batch_stats = bool(bn.training)
        if batch_stats:
            momentum = 0.1 if bn.momentum is None else float(bn.momentum)
            track = bn.track_running_stats and bn.running_mean is not None
            scale, shift, mean, rstd = ops.asp_bn_batch_stats(z, g, bt, bn.eps, momentum,
                                                              bn.running_mean if track else None,
                                                              bn.running_var if track else None)
            if track and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        else:
            rstd = torch.rsqrt(bn.running_var.float() + bn.eps)
            scale = (g * rstd).contiguous()
            mean = bn.running_mean.float().contiguous()
            shift = (bt - mean * scale).contiguous()
        y16 = ops.asp_relu_bn_tanh(z, scale, shift)                                  # tanh(BN(ReLU(.)))
        logits = ops.gemm_f16(y16, ops.cast_f16(w2f), b2.detach().float(), 0, F32).contiguous()
        out = ops.asp_pool(x, logits.view(B, T, C))
        ctx.save_for_backward(x, cat16, z, y16, logits, out, scale, shift, mean, rstd, w1f, w2f)
        ctx.batch_stats = batch_stats
        ctx.shapes = (w1.shape, w2.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, cat16, z, y16, logits, out, scale, shift, mean, rstd, w1f, w2f = ctx.saved_tensors
        B, T, C = x.shape
        A = z.shape[1]
        dev = x.device
        inv_ls = 1.0 / LOSS_SCALE
        dscaled, inv = _entry(dout)
        dlg16, dx = ops.asp_pool_bwd(x, logits.view(B, T, C), out, dscaled)
        dW2 = torch.zeros(C, A, dtype=F32, device=dev)
        ops.gemm_wgrad_f16(dlg16, y16, dW2)
        db2 = torch.zeros(C, dtype=F32, device=dev)
        ops.colsum(dlg16, db2, inv_ls)
        dh = ops.gemm_f16(dlg16, ops.cast_f16_transpose(w2f, C), None, 0, F32).contiguous()       # [B*T, A]
        dgamma = torch.empty(A, dtype=F32, device=dev)
        dbeta = torch.empty(A, dtype=F32, device=dev)
        dz16 = ops.asp_act_bwd(dh, z, scale, shift, mean, rstd, ctx.batch_stats, dgamma, dbeta, inv_ls)
        dW1 = torch.zeros(A, 3 * C, dtype=F32, device=dev)
        ops.gemm_wgrad_f16(dz16, cat16, dW1)
        db1 = torch.zeros(A, dtype=F32, device=dev)
        ops.colsum(dz16, db1, inv_ls)
        dcat = ops.gemm_f16(dz16, ops.cast_f16_transpose(w1f, A), None, 0, F32)                    # [B*T, 3C]
        ops.asp_front_bwd_(x, dcat, dx)
        _leave_(dx, inv)
        _leave_(dW1, inv)
        _leave_(dW2, inv)
        for t in (db1, db2, dgamma, dbeta):           # already divided by LOSS_SCALE in their kernels
            _leave_(t, inv, 1.0)
        s1, s2 = ctx.shapes
        return dx, dW1.view(s1), db1, dgamma, dbeta, dW2.view(s2), db2, None
