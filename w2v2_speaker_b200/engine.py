"""Forward engine of the wav2vec2 encoder on the sm_100a kernels.

This is the host-side schedule of the hot path (one launch sequence per batch):

    wav f32 [B,N]
      -> conv0+GroupNorm+GELU (fused, HBM-bound)                       -> f16 [B,L0,512] channels-last
      -> conv1..6 as tap-GEMMs on tcgen05 (+GELU epilogue)             -> f16 ... f32 [B,T,512]
      -> LayerNorm(512) -> proj GEMM(+bias) -> h0 f32 [B*T,H]
      -> posconv (tcgen05, shifted-slab) ; LN(h0 + pos)                -> h f32 / f16
      -> L x { QKV GEMM -> attention (tcgen05) -> out GEMM -> LN(+bias+res) -> FFN1 GEMM(+bias+GELU)
               -> FFN2 GEMM -> LN(+bias+res) }
      -> last_hidden_state f32 [B,T,H]

GEMM operands are fp16 (rounded to nearest once, by the producing kernel), accumulation, residual
stream, LayerNorm / GroupNorm / softmax statistics are fp32.  Follows HF:1327-1383 in eval mode
(dropout, LayerDrop, SpecAugment are identity; R:src/models/wav2vec2.py:62-76).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops

F16, F32 = torch.float16, torch.float32


class WeightPrep:
    """Job table of ``w2v2_prepare_weights``: every trainable matrix -> fp16 copy (+ transposed fp16 copy,
    + folded scalar), every fused bias -> assembled fp32 vector, all in ONE launch per optimizer step."""

    DT = np.dtype([("src", "<u8"), ("dst16", "<u8"), ("dstT16", "<u8"), ("dst32", "<u8"), ("R", "<i4"), ("C", "<i4"),
                   ("ld", "<i4"), ("ldt", "<i4"), ("scale", "<f4"), ("scale_t", "<f4"), ("tile_begin", "<i8")])

    def __init__(self):
        self.jobs: Dict[str, dict] = {}
        self._table = None
        self._tiles = 0
        self._keep = []

    def add(self, key: str, src: torch.Tensor, dst16=None, dst32=None, scale: float = 1.0):
        src2 = src if src.dim() == 2 else src.view(1, -1)
        assert src2.dtype == F32 and src2.is_contiguous()
        d = dst16 if dst16 is not None else dst32
        d2 = d if d.dim() == 2 else d.view(1, -1)
        assert d2.shape == src2.shape and d2.stride(-1) == 1
        self.jobs[key] = dict(src=src2, dst16=d2 if dst16 is not None else None, dst32=d2 if dst32 is not None else None,
                              dstT=None, scale=float(scale), scale_t=float(scale))
        self._table = None

    def add_transposed(self, key: str, dstT: torch.Tensor, scale: Optional[float] = None):
        """dstT: f16 view [C, R] (row pitch = ldt) receiving the transpose of job `key` (x `scale` of the source
        if given, else the job's own scale)."""
        j = self.jobs[key]
        if scale is not None:
            j["scale_t"] = float(scale)
        assert dstT.shape == (j["src"].shape[1], j["src"].shape[0]) and dstT.stride(1) == 1
        j["dstT"] = dstT
        self._table = None

    def run(self):
        if self._table is None:
            rec = np.zeros(len(self.jobs), dtype=self.DT)
            edge = ops.prepare_tile_edge()                # tiles the kernel walks: 64 x 64 (32 x 32 for the older forms)
            t = 0
            for i, j in enumerate(self.jobs.values()):
                R, C = j["src"].shape
                rec[i] = (j["src"].data_ptr(), j["dst16"].data_ptr() if j["dst16"] is not None else 0,
                          j["dstT"].data_ptr() if j["dstT"] is not None else 0,
                          j["dst32"].data_ptr() if j["dst32"] is not None else 0, R, C,
                          (j["dst16"] if j["dst16"] is not None else j["dst32"]).stride(0) if R > 1 else C,
                          j["dstT"].stride(0) if j["dstT"] is not None else 0, j["scale"], j["scale_t"], t)
                t += ((R + edge - 1) // edge) * ((C + edge - 1) // edge)
            dev = next(iter(self.jobs.values()))["src"].device
            self._table = torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
            self._tiles = t
        ops.prepare_weights(self._table, len(self.jobs), self._tiles)


@dataclass(frozen=True)
class ArchConfig:
    """Architecture numbers (HF Wav2Vec2Config fields the kernels depend on)."""
    name: str = "base"
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    ffn: int = 3072
    conv_dim: int = 512
    conv_kernel: tuple = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: tuple = (5, 2, 2, 2, 2, 2, 2)
    pos_kernel: int = 128
    pos_groups: int = 16
    eps: float = 1e-5
    # the "-lv60" / XLSR checkpoints (HF: feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True)
    feat_extract_norm: str = "group"      # "layer": every conv layer is Conv1d(+bias) -> LayerNorm(C) -> GELU (HF:275-299)
    conv_bias: bool = False
    stable_layer_norm: bool = False       # pre-LN encoder layers + final encoder LayerNorm (HF:632-655, HF:731-799)

    def conv_lengths(self, n: int) -> List[int]:
        out = []
        for k, s in zip(self.conv_kernel, self.conv_stride):
            n = (n - k) // s + 1                 # HF:1012-1018
            out.append(n)
        return out


BASE = ArchConfig()
LARGE = ArchConfig(name="large", hidden=1024, layers=24, heads=16, ffn=4096)
LARGE_LV60 = ArchConfig(name="large-lv60", hidden=1024, layers=24, heads=16, ffn=4096, feat_extract_norm="layer",
                        conv_bias=True, stable_layer_norm=True)


def arch_from_id(huggingface_id: str) -> ArchConfig:
    """Size detection by substring of the HF id, as the reference does (R:src/models/wav2vec2.py:112-117: 768 / 1024
    features); among the large checkpoints the "-lv60" and XLSR ones are the stable-layer-norm variant (their HF
    configs: feat_extract_norm="layer", conv_bias, do_stable_layer_norm), which ``from_pretrained`` would build."""
    if "base" in huggingface_id:
        return BASE
    if "large" in huggingface_id:
        return LARGE_LV60 if ("lv60" in huggingface_id or "xlsr" in huggingface_id) else LARGE
    raise ValueError("cannot determine num features")


class PreparedWeights:
    """fp16 / re-laid-out copies of the encoder parameters in the form the kernels consume.

    Built from a dict keyed by HF state_dict names (fp32, on the GPU).  Must be rebuilt after the
    parameters change (optimizer step / load_state_dict)."""

    def __init__(self, p: Dict[str, torch.Tensor], arch: ArchConfig):
        self.arch = arch
        H = arch.hidden
        f = lambda k: p[k].detach().to(F32).contiguous()
        self.conv0_w = f("feature_extractor.conv_layers.0.conv.weight").view(arch.conv_dim, arch.conv_kernel[0])
        self.gn_g = f("feature_extractor.conv_layers.0.layer_norm.weight")
        self.gn_b = f("feature_extractor.conv_layers.0.layer_norm.bias")
        self.layer_mode = arch.feat_extract_norm == "layer"
        if self.layer_mode:
            # LayerNorm conv layers (HF:275-299): per-layer conv bias and LayerNorm affine; conv layer 0 (1 -> C, k, stride s
            # with k = 2s) runs on the same tap-GEMM as the others by viewing the waveform as [N / s, s] "channels": a k = 2,
            # stride-1 conv over s input channels, zero-padded to the 64-channel granule of the GEMM's operand tiles
            k0, s0 = arch.conv_kernel[0], arch.conv_stride[0]
            if k0 != 2 * s0 or s0 > 64:
                raise NotImplementedError("layer-norm feature extractor: conv layer 0 must have kernel = 2 x stride")
            w0 = torch.zeros(arch.conv_dim, 64, 2, dtype=F32, device=self.conv0_w.device)
            w0[:, :s0, :] = self.conv0_w.view(arch.conv_dim, 2, s0).transpose(1, 2)
            self.conv0_w_taps = ops.conv_weight_tapmajor(w0)
            n = len(arch.conv_kernel)
            self.conv_b = [f(f"feature_extractor.conv_layers.{i}.conv.bias") if arch.conv_bias else None for i in range(n)]
            self.conv_ln_g = [f(f"feature_extractor.conv_layers.{i}.layer_norm.weight") for i in range(n)]
            self.conv_ln_b = [f(f"feature_extractor.conv_layers.{i}.layer_norm.bias") for i in range(n)]
        self.conv_w = [None] + [ops.conv_weight_tapmajor(f(f"feature_extractor.conv_layers.{i}.conv.weight"))
                                for i in range(1, len(arch.conv_kernel))]
        # trainable feature extractor: remember the live fp32 sources so update() can re-derive conv_w
        # (conv0_w / gn_g / gn_b alias the parameters and follow them without a copy)
        names = [f"feature_extractor.conv_layers.{i}.conv.weight" for i in range(1, len(arch.conv_kernel))]
        self._cnn_sources = ([None] + [f(n) for n in names]) if any(p[n].requires_grad for n in names) else None
        self.fp_ln_g = f("feature_projection.layer_norm.weight")
        self.fp_ln_b = f("feature_projection.layer_norm.bias")
        self.fp_b = f("feature_projection.projection.bias")
        self._pos_v = f("encoder.pos_conv_embed.conv.parametrizations.weight.original1")
        self._pos_g = f("encoder.pos_conv_embed.conv.parametrizations.weight.original0").view(-1)
        self._pos_w = {}          # folded weight per "taps per MMA" layout (depends on the sequence length)
        self.pos_b = f("encoder.pos_conv_embed.conv.bias")
        self.enc_ln_g = f("encoder.layer_norm.weight")
        self._groups = arch.pos_groups
        self.enc_ln_b = f("encoder.layer_norm.bias")
        d = H // arch.heads
        scale = float(d) ** -0.5
        dev = self.enc_ln_g.device
        # fp16 operand copies, filled (and re-filled after every optimizer step) by one batched launch
        self.prep = WeightPrep()
        self._aliased = True          # do the job sources alias the live parameters (fp32, contiguous)?

        def src(k):
            t = f(k)
            self._aliased = self._aliased and t.data_ptr() == p[k].data_ptr()
            return t

        self.fp_w = torch.empty(H, arch.conv_dim, dtype=F16, device=dev)
        self.prep.add("feature_projection.projection.weight", src("feature_projection.projection.weight"), dst16=self.fp_w)
        self.layers = []
        for l in range(arch.layers):
            pre = f"encoder.layers.{l}."
            # the softmax scale d^-0.5 is folded into the q projection (HF:528 scales q)
            wqkv = torch.empty(3 * H, H, dtype=F16, device=dev)
            bqkv = torch.empty(3 * H, dtype=F32, device=dev)
            for i, n in enumerate("qkv"):
                self.prep.add(pre + f"attention.{n}_proj.weight", src(pre + f"attention.{n}_proj.weight"),
                              dst16=wqkv[i * H:(i + 1) * H], scale=scale if n == "q" else 1.0)
                self.prep.add(pre + f"attention.{n}_proj.bias", src(pre + f"attention.{n}_proj.bias"),
                              dst32=bqkv[i * H:(i + 1) * H], scale=scale if n == "q" else 1.0)
            L = dict(wqkv=wqkv, bqkv=bqkv, bo=f(pre + "attention.out_proj.bias"),
                     ln1_g=f(pre + "layer_norm.weight"), ln1_b=f(pre + "layer_norm.bias"),
                     b1=f(pre + "feed_forward.intermediate_dense.bias"), b2=f(pre + "feed_forward.output_dense.bias"),
                     ln2_g=f(pre + "final_layer_norm.weight"), ln2_b=f(pre + "final_layer_norm.bias"))
            for name, key in (("wo", "attention.out_proj.weight"), ("w1", "feed_forward.intermediate_dense.weight"),
                              ("w2", "feed_forward.output_dense.weight")):
                w = src(pre + key)
                L[name] = torch.empty(w.shape, dtype=F16, device=dev)
                self.prep.add(pre + key, w, dst16=L[name])
            self.layers.append(L)
        self.prep.run()

    def update(self) -> bool:
        """Re-derive the kernel-form copies of the (non-CNN) parameters after an in-place parameter update.
        False if the sources do not alias the parameters (then the caller rebuilds from scratch)."""
        if not self._aliased:
            return False
        self.prep.run()
        # the folded positional-conv weights that were in use are re-derived right here (the caller runs this on the
        # optimizer stream, under the next step's CNN forward) instead of lazily inside the next forward
        for u in list(self._pos_w.keys()):
            self._pos_w[u] = ops.posconv_fold_weight(self._pos_v, self._pos_g, self._groups, u)
        if self._cnn_sources is not None:      # unfrozen feature extractor: its tap-major fp16 weights change too
            for i, src in enumerate(self._cnn_sources):
                if src is not None:
                    self.conv_w[i] = ops.conv_weight_tapmajor(src)
        return True


def _pos_w(self, T: int) -> torch.Tensor:
    u = ops.posconv_taps_per_mma(T, self.arch.hidden, self._groups)
    if u not in self._pos_w:
        self._pos_w[u] = ops.posconv_fold_weight(self._pos_v, self._pos_g, self._groups, u)
    return self._pos_w[u]


PreparedWeights.pos_w = _pos_w


class EncoderEngine:
    """Runs the eval-mode forward of the encoder for one (weights, arch) pair."""

    def __init__(self, weights: PreparedWeights):
        self.w = weights
        self.arch = weights.arch

    # -- HF:409-419 ----------------------------------------------------------------------------
    def feature_extractor(self, wav: torch.Tensor, stages: Optional[list] = None,
                          lens: Optional[torch.Tensor] = None, normalize: bool = False) -> torch.Tensor:
        """wav [B,N] (f32, or raw int16 PCM) -> channels-last f32 [B,T,C] (HF returns its transpose [B,C,T]).  lens: int32
        [B] sample counts of a zero-padded ragged batch (frames behind an utterance's end are computed but meaningless).
        normalize: the waveform is NOT yet standardised -- the reference's input normaliser is folded into conv layer 0."""
        a, w = self.arch, self.w
        if wav.dim() != 2:
            raise ValueError(f"expected wav_input of shape [BATCH_SIZE, NUM_SAMPLES], got {tuple(wav.shape)}")
        if a.conv_lengths(wav.shape[1])[-1] < 1:
            # HF fails inside the conv stack here ("kernel size can't be greater than actual input size")
            raise ValueError(f"utterances of {wav.shape[1]} samples are shorter than the receptive field of the feature "
                             f"extractor (no output frame)")
        if w.layer_mode:
            # LayerNorm conv layers normalise every frame on its own: a zero-padded ragged batch needs no length-aware
            # statistics here (frames of an utterance that exist see only its own samples; the others are never used).
            # Raw PCM / un-normalised input: the standardisation is its own pass (nothing to fold it into).
            if normalize:
                if lens is not None:
                    raise NotImplementedError("raw / un-normalised input of a RAGGED batch is not built for the layer-norm "
                                              "feature extractor (-lv60 / XLSR checkpoints): normalise per utterance first")
                wav = ops.normalize_wav(wav)[0]
            return self._feature_extractor_layer_norm(wav.float(), stages)
        h = ops.conv0_gn_gelu(wav, w.conv0_w, w.gn_g, w.gn_b, a.eps, lens, normalize)
        if stages is not None:
            stages.append(h)
        n = len(a.conv_kernel)
        for i in range(1, n):
            h = ops.conv1d_cl_f16(h, w.conv_w[i], a.conv_kernel[i], a.conv_stride[i], act=1,
                                  out_dtype=F32 if i == n - 1 else F16)
            if stages is not None:
                stages.append(h)
        return h

    def _feature_extractor_layer_norm(self, wav: torch.Tensor, stages: Optional[list]) -> torch.Tensor:
        """The "-lv60" / XLSR feature extractor (HF:275-299, every layer): Conv1d(+bias) -> LayerNorm over the channels ->
        GELU.  Each layer is the tap-GEMM (fp32 out), the row LayerNorm kernel (which adds the conv bias on its way in)
        and the GELU pass: three launches per layer instead of one -- this variant is built for coverage (no reference
        configuration names it), not tuned."""
        a, w = self.arch, self.w
        B, N = wav.shape
        s0 = a.conv_stride[0]
        rows = N // s0
        x = torch.zeros(B, rows, 64, dtype=F16, device=wav.device)
        x[:, :, :s0] = wav[:, :rows * s0].view(B, rows, s0)
        n = len(a.conv_kernel)
        h = x
        for i in range(n):
            if i == 0:
                z = ops.conv1d_cl_f16(h, w.conv0_w_taps, 2, 1, act=0, out_dtype=F32)                # [B, rows - 1, C]
            else:
                z = ops.conv1d_cl_f16(h, w.conv_w[i], a.conv_kernel[i], a.conv_stride[i], act=0, out_dtype=F32)
            y32, _ = ops.layernorm(z, w.conv_ln_g[i], w.conv_ln_b[i], a.eps, bias=w.conv_b[i], want16=False)
            h, _ = ops.gelu_fwd(y32, F32 if i == n - 1 else F16)
            if stages is not None:
                stages.append(h)
        return h

    # -- HF:429-434 ----------------------------------------------------------------------------
    def feature_projection(self, feat_btc: torch.Tensor) -> torch.Tensor:
        a, w = self.arch, self.w
        B, T, C = feat_btc.shape
        _, n16 = ops.layernorm(feat_btc.contiguous().view(B * T, C), w.fp_ln_g, w.fp_ln_b, a.eps, want32=False)
        h0 = ops.gemm_f16(n16, w.fp_w, w.fp_b, 0, F32)
        return h0.view(B, T, a.hidden)

    POS_SINGLE_SLAB = 256      # frames the shifted-slab positional-conv kernel holds in shared memory at once

    def _posconv_long(self, x16: torch.Tensor) -> torch.Tensor:
        """Positional conv embedding (HF:326-379) of a sequence longer than one slab (full-utterance evaluation):
        the time axis is cut into chunks of 128 frames, each handed to the slab kernel as its own 'utterance'
        together with a real-data halo of K/2 = 64 frames on both sides (zeros beyond the ends of the sequence,
        which is exactly the conv's own padding); only the middle 128 outputs of a chunk are kept -- they depend on
        nothing outside it.  The gather is strided-copy plumbing; the arithmetic is the same tcgen05 kernel."""
        a, w = self.arch, self.w
        B, T, H = x16.shape
        K = a.pos_kernel
        halo, tc = K // 2, self.POS_SINGLE_SLAB - K
        n = (T + tc - 1) // tc
        xp = torch.zeros(B, n * tc + K, H, dtype=F16, device=x16.device)
        xp[:, halo:halo + T] = x16
        chunks = xp.as_strided((B, n, tc + K, H), (xp.stride(0), tc * H, H, 1)).contiguous().view(B * n, tc + K, H)
        out = ops.posconv(chunks, w.pos_w(tc + K), w.pos_b, a.pos_groups, K)              # [B*n, tc+K, H] f32
        return out.view(B, n, tc + K, H)[:, :, halo:halo + tc].reshape(B, n * tc, H)[:, :T].contiguous()

    # -- HF:668-727 ----------------------------------------------------------------------------
    def encoder(self, h0: torch.Tensor, hidden_states: Optional[list] = None,
                frame_lens: Optional[torch.Tensor] = None) -> torch.Tensor:
        """h0 f32 [B,T',H] (any sequence, e.g. with a CLS frame prepended) -> last_hidden_state f32.
        frame_lens: int32 [B] frames per utterance of a padded ragged batch -- the positional conv sees zeros behind each
        utterance's end and attention only its own keys, so the valid rows equal what a batch of one computes."""
        a, w = self.arch, self.w
        B, T, H = h0.shape
        M = B * T
        h0 = h0.contiguous()
        x16 = ops.cast_f16(h0) if frame_lens is None else ops.cast_f16_rowmask(h0, frame_lens)
        if T <= self.POS_SINGLE_SLAB:
            pos = ops.posconv(x16, w.pos_w(T), w.pos_b, a.pos_groups, a.pos_kernel)
        else:
            pos = self._posconv_long(x16)
        if a.stable_layer_norm:
            return self._encoder_stable(h0, pos, hidden_states, frame_lens)
        h32, h16 = ops.layernorm(pos.view(M, H), w.enc_ln_g, w.enc_ln_b, a.eps, residual=h0.view(M, H))
        if hidden_states is not None:
            hidden_states.append(h32.view(B, T, H))
        # one native schedule call per layer (csrc/schedule.cu); the intermediates are shared by all layers,
        # the residual stream ping-pongs between two buffer pairs (kept per layer when hidden states are traced)
        from . import schedule as sched
        sizes = sched.layer_buffer_sizes(B, T, H, a.heads, a.ffn, False)
        keep = hidden_states is not None
        n_out = len(w.layers) if keep else 2
        for i in range(n_out):
            sizes[f"h32.{i}"] = M * H * 4
            sizes[f"h16.{i}"] = M * H * 2
        ar = sched.Arena(sizes, h0.device)
        p32, p16 = h32.data_ptr(), h16.data_ptr()
        out = h32.view(B, T, H)
        for l, lw in enumerate(w.layers):
            i = l if keep else (l & 1)
            sched.run_layer_fwd(sched.fwd_args(a, B, T, l, lw, p32, p16, ar, False,
                                               out32=ar.ptr(f"h32.{i}"), out16=ar.ptr(f"h16.{i}"),
                                               key_lens=frame_lens.data_ptr() if frame_lens is not None else None))
            p32, p16 = ar.ptr(f"h32.{i}"), ar.ptr(f"h16.{i}")
            out = ar.tensor(f"h32.{i}", F32, (B, T, H))
            if keep:
                hidden_states.append(out)
        return out

    def _encoder_stable(self, h0: torch.Tensor, pos: torch.Tensor, hidden_states: Optional[list],
                        frame_lens: Optional[torch.Tensor]) -> torch.Tensor:
        """Stable-layer-norm encoder (HF:731-799 / HF:632-655), eval mode:  h = h0 + pos;  per layer
        h = h + attn(LN1(h)), h = h + FFN(LN2(h));  out = LN(h).  The same GEMM / attention / LayerNorm kernels as the
        post-LN schedule, composed per layer from here (the residual sums are their own fp32 passes): built for
        coverage of the -lv60 / XLSR checkpoints, not tuned."""
        a, w = self.arch, self.w
        B, T, H = h0.shape
        M = B * T
        h, _ = ops.add2_cast(h0.view(M, H), pos.view(M, H), want16=False)
        for lw in w.layers:
            if hidden_states is not None:
                hidden_states.append(h.view(B, T, H))
            _, a16 = ops.layernorm(h, lw["ln1_g"], lw["ln1_b"], a.eps, want32=False)
            qkv = ops.gemm_f16(a16, lw["wqkv"], lw["bqkv"], 0, F16)
            if frame_lens is not None:
                att = ops.attention_lens(qkv, B, T, H, a.heads, frame_lens)
            else:
                att = ops.attention(qkv, B, T, H, a.heads)
            o = ops.gemm_f16(att, lw["wo"], lw["bo"], 0, F32)
            h, _ = ops.add2_cast(o, h, want16=False)
            _, c16 = ops.layernorm(h, lw["ln2_g"], lw["ln2_b"], a.eps, want32=False)
            g16 = ops.gemm_f16(c16, lw["w1"], lw["b1"], 1, F16)
            f2 = ops.gemm_f16(g16, lw["w2"], lw["b2"], 0, F32)
            h, _ = ops.add2_cast(f2, h, want16=False)
        out, _ = ops.layernorm(h, w.enc_ln_g, w.enc_ln_b, a.eps, want16=False)
        out = out.view(B, T, H)
        if hidden_states is not None:
            hidden_states.append(out)
        return out

    # -- HF:1327-1383 --------------------------------------------------------------------------
    def frame_lengths(self, sample_lengths) -> List[int]:
        return [self.arch.conv_lengths(int(n))[-1] for n in sample_lengths]

    def forward(self, wav: torch.Tensor, trace: Optional[dict] = None, lengths=None, normalize: bool = False) -> torch.Tensor:
        """wav f32 [B,N] -> last_hidden_state f32 [B,T,H].  lengths: host sequence of B sample counts when `wav` is a
        zero-padded ragged batch; rows t >= frame_lengths(lengths)[b] of the result are padding."""
        stages = [] if trace is not None else None
        lens_s = lens_f = None
        if lengths is not None:
            lengths = [int(n) for n in lengths]
            if len(lengths) != wav.shape[0] or max(lengths) > wav.shape[1] or min(self.frame_lengths(lengths)) < 1:
                raise ValueError("lengths must hold one sample count per utterance, each within the padded length and "
                                 "long enough for one output frame")
            lens_s = torch.tensor(lengths, dtype=torch.int32).to(wav.device, non_blocking=True)
            lens_f = torch.tensor(self.frame_lengths(lengths), dtype=torch.int32).to(wav.device, non_blocking=True)
        feat = self.feature_extractor(wav, stages, lens_s, normalize)
        h0 = self.feature_projection(feat)
        hs = [] if trace is not None else None
        out = self.encoder(h0, hs, lens_f)
        if trace is not None:
            trace["conv"] = stages
            trace["proj"] = h0
            trace["hidden_states"] = hs
        return out
