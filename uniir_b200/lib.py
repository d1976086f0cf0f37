"""ctypes binding of libuniir_b200.so (include/uniir_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an
exception is raised.  Nothing here imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuniir_b200.so")

BF16, F16, F32 = 0, 1, 2
ACT_NONE, ACT_QUICKGELU, ACT_GELU, ACT_TANH, ACT_RELU = 0, 1, 2, 3, 4


class UniirError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn", C.c_int32),
        ("in_dtype", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_dtype", C.c_int32),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int64), ("res_dtype", C.c_int32),
        ("aux_out", C.c_void_p), ("ld_aux", C.c_int64),
        ("act_grad_src", C.c_void_p), ("ld_ags", C.c_int64),
        ("alpha", C.c_float),
        ("accumulate", C.c_int32),
        ("split_k", C.c_int32),
        ("row_scale", C.c_void_p), ("row_scale_rows", C.c_int32),
    ]


VP, I32, I64 = C.c_void_p, C.c_int32, C.c_int64

_SIGS = {
    "uniir_abi_version": (C.c_int, []),
    "uniir_last_error": (C.c_char_p, []),
    "uniir_launch_count": (C.c_int64, []),
    "uniir_gemm_epi": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "uniir_row_inv_norm": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int64,
                                     C.c_void_p, C.c_void_p]),
    "uniir_ip_topk_plan": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_size_t)]),
    "uniir_ip_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uniir_topk_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                   C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "uniir_layernorm_fwd": (C.c_int, [VP, I64, VP, VP, VP, C.c_float, I64, I32, VP, I64, I32, VP, VP, VP]),
    "uniir_layernorm_bwd": (C.c_int, [VP, I64, I32, VP, I64, VP, VP, VP, VP, I64, I32, VP, I64, I32, VP, VP, VP, VP, VP]),
    "uniir_patch_im2col": (C.c_int, [VP, I32, I32, I32, I32, I32, VP, VP]),
    "uniir_vision_embed_fwd": (C.c_int, [VP, VP, VP, I32, I32, I32, VP, VP]),
    "uniir_vision_embed_bwd": (C.c_int, [VP, I32, I32, I32, VP, VP, VP, VP]),
    "uniir_text_embed_fwd": (C.c_int, [VP, VP, VP, I32, I32, I32, VP, VP, VP]),
    "uniir_text_embed_bwd": (C.c_int, [VP, VP, I32, I32, I32, VP, VP, VP]),
    "uniir_score_fuse_fwd": (C.c_int, [VP, VP, VP, VP, I64, I32, VP, VP]),
    "uniir_score_fuse_bwd": (C.c_int, [VP, VP, VP, I64, I32, VP, VP, VP]),
    "uniir_score_fuse_scatter": (C.c_int, [VP, VP, VP, I32, I32, VP, VP]),
    "uniir_score_fuse_gather_bwd": (C.c_int, [VP, VP, VP, I32, I32, I32, VP, VP]),
    "uniir_colsum_accum": (C.c_int, [VP, I64, I32, I64, VP, VP]),
    "uniir_pad_cast_bf16": (C.c_int, [VP, I64, I32, I32, VP, VP]),
    "uniir_unpad_accum": (C.c_int, [VP, I64, I32, I32, VP, VP]),
    "uniir_cast_bf16": (C.c_int, [VP, I64, VP, VP]),
    "uniir_adamw_step": (C.c_int, [VP, VP, VP, VP, VP, I64, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_float, I32, C.c_float, VP]),
    "uniir_attn_fwd": (C.c_int, [VP, I32, I32, I32, I32, VP, VP, VP]),
    "uniir_attn_fwd_ex": (C.c_int, [VP, I64, VP, VP, I64, I32, I32, I32, I32, I32, VP, VP, I64, VP, VP]),
    "uniir_attn_bwd": (C.c_int, [VP, VP, VP, VP, I32, I32, I32, I32, VP, VP, VP]),
    "uniir_attn_fwd_bias": (C.c_int, [VP, I64, VP, VP, I64, I32, I32, I32, I32, C.c_float, VP, VP, I64, VP, VP]),
    "uniir_rmsnorm_fwd": (C.c_int, [VP, I64, VP, C.c_float, I64, I32, VP, I64, I32, VP]),
    "uniir_token_mean": (C.c_int, [VP, I32, I32, I32, VP, VP]),
    "uniir_rmsnorm_bwd": (C.c_int, [VP, VP, VP, C.c_float, I64, I32, VP, I32, VP, VP]),
    "uniir_token_mean_bwd": (C.c_int, [VP, I32, I32, I32, VP, VP]),
    "uniir_attn_bwd_bias": (C.c_int, [VP, I64, VP, VP, I64, VP, I64, VP, I64, VP, I32, I32, I32, I32, C.c_float, VP, VP, VP,
                                      VP, I64, VP, VP, I64, VP]),
    "uniir_attn_set_mode": (C.c_int, [I32]),
    "uniir_attn_bwd_workspace_floats": (C.c_int64, [I32, I32, I32, I32]),
    "uniir_attn_set_trace": (C.c_int, [VP]),
    "uniir_infonce_workspace_bytes": (C.c_size_t, [I32, I32]),
    "uniir_infonce_normalize": (C.c_int, [VP, VP, VP, I32, I32, VP, VP, VP, VP, VP]),
    "uniir_infonce_fwd": (C.c_int, [VP, VP, I32, I32, I32, VP, I32, VP, VP, VP, VP, C.c_size_t, VP]),
    "uniir_infonce_bwd": (C.c_int, [VP, VP, I32, I32, I32, VP, I32, VP, VP, VP, VP, VP, VP, VP]),
    "uniir_attn_bwd_ex": (C.c_int, [VP, I64, VP, VP, I64, VP, I64, VP, I64, VP, I32, I32, I32, I32, I32, VP, VP, VP, I64, VP, VP, I64, VP]),
    "uniir_ema_update": (C.c_int, [VP, VP, I64, C.c_float, VP, VP]),
    "uniir_tanh_bwd": (C.c_int, [VP, VP, I64, VP, VP]),
    "uniir_sgemm_temp": (C.c_int, [VP, VP, I32, I32, I32, I32, VP, VP, VP]),
    "uniir_blip_softce": (C.c_int, [VP, VP, VP, VP, I32, I32, C.c_float, C.c_float, VP, VP, VP, VP, VP]),
    "uniir_hardneg_loss": (C.c_int, [VP, VP, VP, VP, I32, I32, I32, I32, VP, VP, VP, VP, VP, VP]),
    "uniir_infonce_normalize_bwd": (C.c_int, [VP, VP, VP, VP, VP, VP, VP, VP, I32, I32, VP, VP]),
    "uniir_dropout": (C.c_int, [VP, I32, VP, VP, I32, I64, VP, C.c_uint32, I64, C.c_float, VP]),
    "uniir_dropout_mask": (C.c_int, [VP, C.c_uint32, I64, I64, C.c_float, VP, VP]),
    "uniir_droppath_scales": (C.c_int, [VP, C.c_uint32, I32, I32, VP, VP, VP]),
    "uniir_row_scale": (C.c_int, [VP, I64, I32, I64, VP, I32, VP]),
    "uniir_attn_fwd_drop": (C.c_int, [VP, I64, VP, VP, I64, I32, I32, I32, I32, I32, VP, C.c_float, VP, VP, I64, VP, VP,
                                      C.c_uint32, C.c_float, VP]),
    "uniir_attn_bwd_drop": (C.c_int, [VP, I64, VP, VP, I64, VP, I64, VP, I64, VP, I32, I32, I32, I32, I32, VP, C.c_float,
                                      VP, VP, VP, VP, I64, VP, VP, I64, VP, C.c_uint32, C.c_float, VP]),
}

_lib = None


def exported_symbols():
    """Names every entry point include/uniir_b200.h declares (kept in sync by a CPU test)."""
    return sorted(_SIGS)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UniirError(
                f"{LIB_PATH} is missing: build it with `python -m uniir_b200.build` "
                "(there is no CPU or PyTorch fallback for the uniir_b200 hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI drifted
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().uniir_last_error()
        raise UniirError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().uniir_launch_count())
