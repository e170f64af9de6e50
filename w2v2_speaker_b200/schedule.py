"""Host side of the native per-layer launch schedules (csrc/schedule.cu): argument blocks of
``w2v2_encoder_layer_fwd`` / ``w2v2_encoder_layer_bwd`` and the device arenas their buffers live in.

One ctypes call per transformer layer replaces ~10 (forward) / ~17 (backward) python -> C round trips; the
buffers of a layer come out of ONE torch allocation (an arena) that is sliced by pointer arithmetic, so the
interpreter does a constant amount of work per layer."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_uint64, c_void_p
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_ALIGN = 256


class LayerFwdArgs(ctypes.Structure):
    """w2v2_layer_fwd_args (include/w2v2_b200.h)."""
    _fields_ = ([(n, c_int) for n in ("B", "T", "H", "heads", "FF", "layer")] +
                [(n, c_float) for n in ("eps", "p_hidden", "p_attn", "p_act")] +
                [("seed", c_uint64)] +
                [(n, c_void_p) for n in ("wqkv", "bqkv", "wo", "bo", "ln1_g", "ln1_b", "w1", "b1", "w2", "b2", "ln2_g", "ln2_b",
                                         "h_in32", "h_in16", "qkv16", "att16", "lse", "o32", "h1_32", "h1_16", "z16", "g16",
                                         "f2_32", "h2_32", "h2_16", "rstd1", "rstd2", "key_lens")])


class LayerBwdArgs(ctypes.Structure):
    """w2v2_layer_bwd_args (include/w2v2_b200.h)."""
    _fields_ = ([(n, c_int) for n in ("B", "T", "H", "heads", "FF", "layer")] +
                [(n, c_float) for n in ("eps", "p_hidden", "p_attn", "p_act", "qscale")] +
                [("seed", c_uint64)] +
                [(n, c_void_p) for n in ("wqkvT", "woT", "w1T", "w2T", "bo", "b2", "ln1_g", "ln2_g",
                                         "h_in32", "h_in16", "qkv16", "att16", "lse", "o32", "h1_32", "h1_16", "z16", "g16",
                                         "f2_32", "dy_a", "dy_b",
                                         "d_wqkv", "d_bqkv", "d_wo", "d_bo", "d_ln1_g", "d_ln1_b", "d_w1", "d_b1", "d_w2", "d_b2",
                                         "d_ln2_g", "d_ln2_b",
                                         "dx2_32", "dx2_16", "dg16", "dz16", "dh1_32", "dx1_16", "datt16", "dqkv16",
                                         "dx1_32", "dh_in32", "h2_32", "rstd1", "rstd2", "ln1_b", "ln2_b")])


class Arena:
    """One device allocation carved into named, 256-byte aligned buffers."""

    def __init__(self, sizes: Dict[str, int], device):
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        for name, nbytes in sizes.items():
            self.offsets[name] = (off, nbytes)
            off += (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        total = max(off, _ALIGN)
        raw = torch.empty(total + _ALIGN, dtype=torch.uint8, device=device)
        pad = -raw.data_ptr() % _ALIGN            # 0 with torch's CUDA allocator (512-byte blocks)
        self.buf = raw[pad:pad + total]
        self.base = self.buf.data_ptr()

    def ptr(self, name: str) -> int:
        return self.base + self.offsets[name][0]

    def tensor(self, name: str, dtype, shape) -> torch.Tensor:
        off, nbytes = self.offsets[name]
        return self.buf[off:off + nbytes].view(dtype).view(shape)


def layer_buffer_sizes(B: int, T: int, H: int, heads: int, FF: int, train: bool) -> Dict[str, int]:
    M = B * T
    s = {"qkv16": M * 3 * H * 2, "att16": M * H * 2, "o32": M * H * 4, "h1_32": M * H * 4, "h1_16": M * H * 2,
         "g16": M * FF * 2, "f2_32": M * H * 4, "h2_32": M * H * 4, "h2_16": M * H * 2}
    if train:
        s["lse"] = B * heads * T * 4
        s["z16"] = M * FF * 2
        s["rstd1"] = M * 4
        s["rstd2"] = M * 4
    return s


def bwd_scratch_sizes(B: int, T: int, H: int, FF: int) -> Dict[str, int]:
    M = B * T
    return {"dx2_32": M * H * 4, "dx2_16": M * H * 2, "dg16": M * FF * 2, "dz16": M * FF * 2, "dh1_32": M * H * 4,
            "dx1_16": M * H * 2, "datt16": M * H * 2, "dqkv16": M * 3 * H * 2,
            "dx1_32.0": M * H * 4, "dh_in32.0": M * H * 4, "dx1_32.1": M * H * 4, "dh_in32.1": M * H * 4}


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def run_layer_fwd(args: LayerFwdArgs) -> None:
    lib = _lib.load()
    rc = lib.w2v2_encoder_layer_fwd(ctypes.byref(args), _lib.stream_ptr())
    if rc != 0:
        raise _lib.W2V2Error(f"w2v2_encoder_layer_fwd failed ({rc}): {lib.w2v2_last_error().decode()}")


def run_layer_bwd(args: LayerBwdArgs) -> None:
    lib = _lib.load()
    rc = lib.w2v2_encoder_layer_bwd(ctypes.byref(args), _lib.stream_ptr())
    if rc != 0:
        raise _lib.W2V2Error(f"w2v2_encoder_layer_bwd failed ({rc}): {lib.w2v2_last_error().decode()}")


def fwd_args(arch, B: int, T: int, layer: int, lw: dict, h_in32: int, h_in16: int, arena: Arena, train: bool,
             p_hidden: float = 0.0, p_attn: float = 0.0, p_act: float = 0.0, seed: int = 0,
             out32: Optional[int] = None, out16: Optional[int] = None, key_lens: Optional[int] = None) -> LayerFwdArgs:
    """Argument block of one layer.  `lw`: engine.PreparedWeights.layers[layer]; h_in*: device pointers;
    out32 / out16 override where the layer output goes (inference ping-pong buffers)."""
    a = LayerFwdArgs()
    a.B, a.T, a.H, a.heads, a.FF, a.layer = B, T, arch.hidden, arch.heads, arch.ffn, layer
    a.eps, a.p_hidden, a.p_attn, a.p_act = arch.eps, p_hidden, p_attn, p_act
    a.seed = seed
    a.wqkv, a.bqkv, a.wo, a.bo = _p(lw["wqkv"]), _p(lw["bqkv"]), _p(lw["wo"]), _p(lw["bo"])
    a.ln1_g, a.ln1_b, a.w1, a.b1 = _p(lw["ln1_g"]), _p(lw["ln1_b"]), _p(lw["w1"]), _p(lw["b1"])
    a.w2, a.b2, a.ln2_g, a.ln2_b = _p(lw["w2"]), _p(lw["b2"]), _p(lw["ln2_g"]), _p(lw["ln2_b"])
    a.h_in32, a.h_in16 = h_in32, h_in16
    for k in ("qkv16", "att16", "o32", "h1_32", "h1_16", "g16", "f2_32"):
        setattr(a, k, arena.ptr(k))
    a.lse = arena.ptr("lse") if train else None
    a.z16 = arena.ptr("z16") if train else None
    a.rstd1 = arena.ptr("rstd1") if train else None
    a.rstd2 = arena.ptr("rstd2") if train else None
    a.h2_32 = out32 if out32 is not None else arena.ptr("h2_32")
    a.h2_16 = out16 if out16 is not None else arena.ptr("h2_16")
    a.key_lens = key_lens          # device pointer of int32 [B] (ragged evaluation batch) or None
    return a
