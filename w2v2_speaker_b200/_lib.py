"""ctypes binding of libw2v2_b200.so (the C ABI of include/w2v2_b200.h).

The library is the product: if it is missing the import of any compute entry point fails
loudly -- there is no PyTorch/CPU fallback path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libw2v2_b200.so")

# name -> (restype, argtypes); must list every symbol declared in include/w2v2_b200.h
SIGNATURES = {
    "w2v2_last_error": (c_char_p, []),
    "w2v2_layernorm_ex2": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                   c_void_p, c_int64, c_int, c_float, c_uint64, c_void_p]),
    "w2v2_layernorm_bwd_from_output": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_uint64, c_void_p]),
    "w2v2_normalize_wav": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "w2v2_cosine_pairs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "w2v2_gemm_f16_dual_gelu": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                        c_void_p, c_int64, c_void_p]),
    "w2v2_scale_copy_f32": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "w2v2_dgrad_accumulates": (c_int, []),
    "w2v2_nvls_allreduce_f32": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "w2v2_conv0_gn_lens": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                   c_int, c_int, c_void_p]),
    "w2v2_conv0_raw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                               c_void_p, c_int, c_int, c_void_p]),
    "w2v2_cast_f16_rowmask": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "w2v2_attention_lens": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "w2v2_stat_pool_lens": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "w2v2_asp_concat_split3_lens": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "w2v2_asp_concat_lens": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "w2v2_asp_pool_lens": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "w2v2_gemm_f16_accum": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "w2v2_grad_entry_scale": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_void_p, c_void_p]),
    "w2v2_scale_f32_dev": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_void_p]),
    "w2v2_gemm_f16_dual_gelu_grad": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                             c_void_p, c_int64, c_void_p]),
    "w2v2_gemm_f16_mul_colsum": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                         c_void_p, c_int64, c_void_p, c_void_p]),
    "w2v2_gemm_f16_gelu_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                       c_void_p, c_int64, c_void_p, c_void_p]),
    "w2v2_attention_bwd_ex2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                       c_uint64, c_float, c_void_p, c_void_p]),
    "w2v2_adam_step_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int,
                                  c_float, c_int, c_void_p]),
    "w2v2_conv0_gn_ex": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                                 c_int, c_void_p]),
    "w2v2_conv0_workspace_offsets": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "w2v2_gemm_f16_taps": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p,
                                   c_int64, c_int, c_void_p, c_int, c_int64, c_int64, c_void_p]),
    "w2v2_gemm_wgrad_f16_batched": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_int,
                                            c_int, c_void_p, c_int64, c_void_p]),
    "w2v2_groupnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                   c_int, c_int, c_int, c_void_p]),
    "w2v2_gemm_profile_start": (c_int, []),
    "w2v2_gemm_profile_stop": (c_int, [c_void_p, c_void_p, c_void_p]),
    "w2v2_encoder_layer_fwd": (c_int, [c_void_p, c_void_p]),
    "w2v2_encoder_layer_bwd": (c_int, [c_void_p, c_void_p]),
    "w2v2_asp_bn_batch_stats": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "w2v2_asp_pool_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_asp_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_float, c_int64, c_int, c_void_p]),
    "w2v2_asp_front_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_abi_version": (c_int, []),
    "w2v2_sm_count": (c_int, []),
    "w2v2_launch_count": (c_int64, []),
    "w2v2_gemm_f16": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_int, c_void_p, c_int64,
                              c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_void_p]),
    "w2v2_conv0_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "w2v2_conv0_gn_gelu": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                   c_int, c_void_p]),
    "w2v2_layernorm": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                               c_int64, c_int, c_void_p]),
    "w2v2_layernorm_ex": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                  c_int64, c_int, c_float, c_uint64, c_void_p]),
    "w2v2_layernorm_bwd_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_uint64, c_void_p]),
    "w2v2_posconv_taps_per_mma": (c_int, [c_int, c_int, c_int]),
    "w2v2_posconv_fold_weight": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "w2v2_posconv_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_void_p]),
    "w2v2_gelu_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "w2v2_prepare_weights": (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    "w2v2_prepare_tile_edge": (c_int, []),
    "w2v2_posconv_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "w2v2_posconv_im2col": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "w2v2_weight_norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int,
                                     c_int, c_void_p]),
    "w2v2_posconv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "w2v2_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "w2v2_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "w2v2_gemm_wgrad_f16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "w2v2_cast_f16_transpose": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "w2v2_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "w2v2_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "w2v2_gelu_bwd_colsum": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "w2v2_colsum": (c_int, [c_void_p, c_int, c_int64, c_int, c_int64, c_float, c_void_p, c_void_p]),
    "w2v2_softmax_ce_bwd": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_mean_pool_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_meanstd_pool_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_aam_softmax_ce_ex": (c_int, [c_void_p, c_int64, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "w2v2_aam_bwd_dcos": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_int, c_void_p,
                                  c_int, c_int, c_int, c_void_p]),
    "w2v2_row_inv_norm": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "w2v2_l2norm_rows_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]),
    "w2v2_dropout": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_float, c_uint64, c_void_p]),
    "w2v2_attention_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_uint64, c_void_p]),
    "w2v2_attention_bwd_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_float, c_uint64, c_void_p]),
    "w2v2_time_mask_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "w2v2_feature_mask": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_time_mask_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "w2v2_add2_cast": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "w2v2_cast_f16_rows": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p]),
    "w2v2_scale_f32": (c_int, [c_void_p, c_int64, c_float, c_void_p]),
    "w2v2_softmax_ce_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p]),
    "w2v2_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int,
                               c_float, c_void_p]),
    "w2v2_stat_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "w2v2_asp_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_asp_concat": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_asp_concat_split3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "w2v2_asp_relu_bn_tanh": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "w2v2_asp_relu_bn_tanh_ubias": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "w2v2_softmax_ce": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "w2v2_aam_softmax_ce": (c_int, [c_void_p, c_int64, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p,
                                    c_void_p, c_int, c_int, c_void_p]),
    "w2v2_l2norm_rows_f16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "w2v2_l2norm_rows_split3": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "w2v2_split3_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "w2v2_mean_rows": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "w2v2_cast_f16": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "w2v2_conv_weight_tapmajor": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

_lib = None


class W2V2Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise W2V2Error(
                f"{LIB_PATH} is missing: build it with `python -m w2v2_speaker_b200.build` "
                "(there is no fallback path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """torch's current stream of the current device as a raw cudaStream_t (the fast C accessor: the
    python-level torch.cuda.current_stream() costs several microseconds per kernel launch)."""
    if _raw_stream is not None:
        return c_void_p(_raw_stream(torch.cuda.current_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name: str, *args):
    """Call an int-returning entry point and raise W2V2Error with the library message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise W2V2Error(f"{name} failed ({rc}): {lib.w2v2_last_error().decode()}")
    return rc


# ---- NVTX ranges (tracing) ------------------------------------------------------------------------------
# W2V2_NVTX=1 brackets the phases of a step (forward / backward / all-reduce / optimizer, and every encoder layer) with
# NVTX ranges, which `ncu --nvtx --nvtx-include "backward/"` and Nsight Systems pick up; off by default (two C calls
# per range).
_NVTX = os.environ.get("W2V2_NVTX", "0") == "1"


class nvtx_range:
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False
