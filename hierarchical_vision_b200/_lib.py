"""ctypes binding of libhv_swin.so -- the C ABI declared in include/hv_swin.h.

There is deliberately no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` carrying ``hv_last_error()`` is raised.  The product path never computes
on the CPU and never routes through ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# HV_SWIN_LIB: another build of the same library (e.g. an instrumented -DHV_TC_TRACE build); never a different backend
LIB_PATH = os.environ.get("HV_SWIN_LIB") or os.path.join(PKG_DIR, "libhv_swin.so")

HV_F32, HV_BF16, HV_U8 = 0, 1, 2

# name -> (restype, argtypes); mirrors include/hv_swin.h one to one
_I, _P, _F, _L, _S = c_int, c_void_p, c_float, c_int64, c_size_t
SIGNATURES = {
    "hv_abi_version": (_I, []),
    "hv_last_error": (c_char_p, []),
    "hv_compiled_arch": (_I, []),
    "hv_window_attn_kernel_kind": (_I, [_I, _I, _I, _I]),
    "hv_window_attn_fwd_variant": (_I, [_I]),
    "hv_window_attn_bwd_variant": (_I, [_I]),
    "hv_window_attn_tc256_variant": (_I, [_I]),
    "hv_window_attn_stats_floats": (_S, [_I, _I, _I, _I, _I, _I, _I]),
    "hv_window_attn_kernel_name": (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _I, c_char_p, _I]),
    "hv_relative_position_index": (_I, [_I, _P]),
    "hv_shift_window_mask": (_I, [_I, _I, _I, _I, _P]),
    "hv_window_token_index": (_I, [_I, _I, _I, _I, _I, _P]),
    "hv_window16_tile_token_index": (_I, [_I, _I, _I, _I, _P]),
    "hv_merge_token_index": (_I, [_I, _I, _I, _P]),
    "hv_window_attn_fwd": (_I, [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "hv_window_attn_bwd_workspace_bytes": (_S, [_I, _I, _I, _I, _I, _I, _I]),
    "hv_window_attn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _S,
                                _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "hv_dq_colsum_workspace_bytes": (_S, [_I]),
    "hv_dq_colsum": (_I, [_P, _P, _P, _S, _L, _I, _I, _P]),
    "hv_ln_residual_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _L, _F, _I, _I, _P]),
    "hv_ln_residual_bwd_workspace_bytes": (_S, [_L, _I]),
    "hv_ln_residual_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _S, _L, _I, _L, _I, _I, _P]),
    "hv_mlp_dgelu_gemm_workspace_bytes": (_S, [_L, _I, _I]),
    "hv_sgdw_step": (_I, [_P, _P, _I, _P, _P, _P, _F, _P]),
    "hv_mlp_fc1_gelu_gemm": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _I, _P]),
    "hv_mlp_dgelu_gemm": (_I, [_P, _P, _P, _P, _P, _P, _P, _S, _L, _I, _I, _I, _P]),
    "hv_bias_gelu_fwd": (_I, [_P, _P, _P, _L, _I, _I, _P]),
    "hv_bias_gelu_bwd_workspace_bytes": (_S, [_L, _I]),
    "hv_bias_gelu_bwd": (_I, [_P, _P, _P, _P, _P, _P, _S, _L, _I, _I, _P]),
    "hv_patch_merge_gather_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "hv_patch_merge_gather_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "hv_cpb_bias_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "hv_cpb_bias_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "hv_cross_entropy_fwd_grad": (_I, [_P, _P, _P, _P, _L, _I, _F, _F, _I, _P]),
    "hv_patch_rows": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found. Build it with `python -m hierarchical_vision_b200.build` "
            "(nvcc, sm_100a). hierarchical_vision_b200 has no CPU or eager fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.hv_abi_version() != 4:
        raise RuntimeError(f"libhv_swin.so ABI version {lib.hv_abi_version()} != 4")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hv_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
