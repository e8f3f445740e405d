"""ctypes binding of libstk.so (the C ABI declared in include/stk.h).

The library is the product's only compute path: if it is missing or fails to load this module
raises — there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong, c_size_t, c_uint32, c_uint64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
# STK_LIB: bring-up override (A/B runs of two builds of the kernel library); the product path is the in-tree build
LIB_PATH = os.environ.get("STK_LIB") or os.path.join(_HERE, "libstk.so")

STK_VERSION = 103

# epilogue ids (include/stk.h)
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GELU_SAVE, EPI_BIAS_RESID, EPI_BIAS_TANH_F32 = 0, 1, 2, 3, 4
EPI_DGELU, EPI_F32_ADD, EPI_F32, EPI_CE_STATS, EPI_CE_DLOGIT, EPI_BIAS_RESID_LN = 5, 6, 7, 8, 9, 10
EPI_BIAS_GELU_SAVE_GRAD, EPI_MUL, EPI_BIAS_DROP_RESID_LN = 11, 12, 13
WS_ATTN_BWD, WS_LINEAR_CE_FWD, WS_LINEAR_CE_BWD = 0, 1, 2
ERR_BAD_ID, ERR_BAD_LABEL, ERR_LABEL_CAPACITY = 1, 2, 4   # bits of the device-side err_flag


class StkError(RuntimeError):
    pass


class AdamSeg(Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("w16", c_void_p), ("p32_copy", c_void_p),
                ("n", c_int64), ("g16", c_void_p)]


class GemmEpilogue(Structure):
    _fields_ = [
        ("bias", c_void_p),
        ("resid", c_void_p),
        ("ldr", c_int64),
        ("c2", c_void_p),
        ("ldc2", c_int64),
        ("labels", c_void_p),
        ("lse", c_void_p),
        ("scale_dev", c_void_p),
        ("ce_partial", c_void_p),
        ("ce_pitch", c_int64),
        ("tgt_logit", c_void_p),
        ("n_offset", c_int32),
        ("ln_gamma", c_void_p),
        ("ln_beta", c_void_p),
        ("ln_mean", c_void_p),
        ("ln_rstd", c_void_p),
        ("drop_seed", c_uint32),
        ("drop_site", c_uint32),
        ("drop_thr", c_uint32),
    ]


_P = c_void_p
_SIGNATURES = {
    "stk_version": (c_int, []),
    "stk_last_error": (c_int, [c_char_p, c_size_t]),
    "stk_launch_count": (c_longlong, []),
    "stk_set_sm_reserve": (c_int, [c_int, c_int]),
    "stk_set_gemm_dynamic": (c_int, [c_int]),
    "stk_embed_text_ln_fwd": (c_int, [c_int, _P, _P, c_int64, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "stk_embed_joint_ln_fwd": (c_int, [c_int, _P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "stk_embed_joint_ln_bwd": (c_int, [c_int, _P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "stk_embed_joint_ln_fwd_shape": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P,
                                             _P, _P, _P, _P]),
    "stk_embed_joint_ln_bwd_shape": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P,
                                             _P, _P, _P, _P, _P]),
    "stk_layernorm_fwd": (c_int, [c_int, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "stk_layernorm_bwd": (c_int, [c_int, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "stk_layernorm_bwd_fused": (c_int, [c_int, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_uint32, c_uint32, c_uint32]),
    "stk_gemm": (c_int, [c_int, _P, c_int, c_int, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int64,
                         POINTER(GemmEpilogue), c_int]),
    "stk_attn_fwd": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P, _P]),
    "stk_attn_fwd_qrows": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "stk_attn_bwd": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    "stk_attn_fwd_dropout": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P, _P, c_uint32, c_uint32, c_uint32]),
    "stk_attn_bwd_dropout": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, c_uint32, c_uint32, c_uint32]),
    "stk_dropout_fwd": (c_int, [c_int, _P, _P, c_int, c_uint32, c_uint32, c_uint32, _P]),
    "stk_dropout_resid_ln_fwd": (c_int, [c_int, _P, _P, _P, c_int, _P, _P, c_uint32, c_uint32, c_uint32, _P, _P, _P, _P]),
    "stk_mask_to_bias": (c_int, [c_int, _P, _P, c_int64, _P]),
    "stk_cast_f32_to_bf16": (c_int, [c_int, _P, _P, _P, c_int64]),
    "stk_gather_rows": (c_int, [c_int, _P, _P, _P, c_int, _P]),
    "stk_scatter_add_rows": (c_int, [c_int, _P, _P, _P, c_int, _P]),
    "stk_colsum": (c_int, [c_int, _P, _P, c_int64, c_int, c_int, _P, c_int]),
    "stk_ce_finalize": (c_int, [c_int, _P, _P, c_int64, _P, c_int, _P, _P]),
    "stk_nsp_head_fwd": (c_int, [c_int, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "stk_scale_heads": (c_int, [c_int, _P, _P, c_int, _P, _P]),
    "stk_masked_mean_pool": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_int, _P]),
    "stk_query_workspace": (c_int64, [c_int, c_int64, c_int64]),
    "stk_compact_labels": (c_int, [c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "stk_linear_ce_fwd": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P, _P, c_int64, _P, _P, _P]),
    "stk_linear_ce_bwd": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, c_int64, _P, _P]),
    "stk_gelu_bwd": (c_int, [c_int, _P, _P, _P, c_int64, _P]),
    "stk_unpack_scale": (c_int, [c_int, _P, _P, _P, c_int64, c_float]),
    "stk_sumsq": (c_int, [c_int, _P, _P, c_int64, _P]),
    "stk_sumsq_bf16": (c_int, [c_int, _P, _P, c_int64, c_float, _P]),
    "stk_adamw_step": (c_int, [c_int, _P, _P, _P, _P, c_int, c_float, c_float, c_float, c_float, c_float, c_float,
                               c_float, _P, c_float, c_float]),
    "stk_nsp_pool_bwd": (c_int, [c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "stk_assemble_pairs": (c_int, [c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "stk_mask_tokens": (c_int, [c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_uint64, c_uint32, c_int64]),
    "stk_cls_head_fwd": (c_int, [c_int, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "stk_cls_pool_bwd": (c_int, [c_int, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def declared_symbols():
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load libstk.so once; raise StkError (never fall back) when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StkError(
            f"{LIB_PATH} not found: build it with `python -m stonkgs_b200.build` "
            "(stonkgs_b200 has no CPU / PyTorch fallback path)")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise StkError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise StkError(f"libstk.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.stk_version() != STK_VERSION:
        raise StkError(f"libstk.so version {lib.stk_version()} != binding version {STK_VERSION}")
    _lib = lib
    return lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().stk_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise StkError(f"{what} failed (code {rc}): {last_error()}")
