"""ctypes binding of libhamt_b200.so (the C ABI in include/hamt_b200.h).

There is NO fallback: if the shared library is missing the import of the compute path raises, and
every op raises RuntimeError(hamt_last_error()) on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libhamt_b200.so")

_lib = None

vp, ll, i32, u32, f32, f64 = C.c_void_p, C.c_longlong, C.c_int, C.c_uint, C.c_float, C.c_double


class EmbedFeatDesc(C.Structure):
    _fields_ = [("t", vp), ("ang", vp), ("A", i32),
                ("w_ang", vp), ("b_ang", vp), ("g_img", vp), ("b_img", vp), ("g_ang", vp), ("be_ang", vp),
                ("add_vec", vp), ("nav_table", vp), ("nav_ids", vp), ("extra", vp),
                ("pos_table", vp), ("pos_ids", vp), ("pos_mod", i32),
                ("g_f", vp), ("b_f", vp),
                ("out", vp), ("M", i32), ("H", i32), ("eps", f32),
                ("seed_ptr", vp), ("site", u32), ("p", f32)]


class EmbedFeatGrads(C.Structure):
    _fields_ = [("dy", vp), ("dt", vp),
                ("dw_ang", vp), ("db_ang", vp), ("dg_img", vp), ("db_img", vp), ("dg_ang", vp), ("dbe_ang", vp),
                ("dadd_vec", vp), ("dnav_table", vp), ("dextra", vp), ("dpos_table", vp), ("dg_f", vp), ("db_f", vp),
                ("db_lin", vp)]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/hamt_b200.h one to one
SIGNATURES = {
    "hamt_abi_version": [],
    "hamt_last_error": [],
    "hamt_launch_count": [],
    "hamt_gemm_bf16": [vp, i32, ll, vp, i32, ll, vp, ll, i32, i32, i32, i32, i32, vp, i32, i32, vp, ll, f32, i32, i32, vp, vp],
    "hamt_gemm_set_auto_pair": [i32],
    "hamt_gemm_set_sm_limit": [i32],
    "hamt_gemm_set_wide_epilogue": [i32],
    "hamt_attn_set_impl": [i32],
    "hamt_ln_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp, u32, f32, vp],
    "hamt_ln_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, u32, f32, vp],
    "hamt_ln_bwd_prenorm": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, u32, f32, vp],
    "hamt_ln_fwd_prenorm": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp, u32, f32, vp],
    "hamt_patchify_bf16": [vp, vp, i32, i32, i32, i32, i32, vp],
    "hamt_vit_embed_fwd": [vp, vp, vp, vp, vp, i32, i32, i32, vp, u32, f32, vp],
    "hamt_vit_embed_bwd": [vp, vp, vp, i32, i32, i32, vp, u32, f32, vp],
    "hamt_attn_fwd": [vp, vp, vp, ll, ll, ll, ll, vp, vp, ll, ll, vp, i32, i32, i32, i32, f32, vp, u32, f32, vp],
    "hamt_attn_bwd": [vp, vp, vp, ll, ll, ll, ll, vp, vp, ll, ll, vp, vp, ll, ll, vp, vp, vp, i32, i32, i32, i32, f32, vp, u32, f32, vp, vp, vp, vp],
    "hamt_embed_text_fwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, u32, f32, vp],
    "hamt_embed_text_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, u32, f32, vp],
    "hamt_embed_feat_fwd": [C.POINTER(EmbedFeatDesc), vp],
    "hamt_embed_feat_bwd": [C.POINTER(EmbedFeatDesc), C.POINTER(EmbedFeatGrads), vp],
    "hamt_cast_f32_to_bf16": [vp, vp, ll, vp],
    "hamt_colsum_bf16": [vp, ll, vp, i32, i32, vp],
    "hamt_mean_pool_fwd": [vp, vp, i32, i32, i32, vp],
    "hamt_mean_pool_bwd": [vp, vp, i32, i32, i32, vp],
    "hamt_add_bf16": [vp, vp, vp, ll, vp],
    "hamt_mul_rows_bf16": [vp, vp, vp, i32, i32, i32, vp],
    "hamt_adamw_workspace_floats": [],
    "hamt_adamw_step": [vp, vp, vp, vp, vp, ll, vp, vp, i32, vp, vp, vp, vp, vp, f64, f64, f64, i32, f32, i32, i32, vp, vp],
    "hamt_rowdot_fwd": [vp, vp, vp, vp, i32, i32, i32, vp],
    "hamt_rowdot_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "hamt_ce_fwd": [vp, ll, vp, vp, vp, i32, i32, vp],
    "hamt_ce_bwd": [vp, ll, vp, vp, vp, vp, vp, ll, i32, i32, vp],
    "hamt_gather_rows_bf16": [vp, vp, vp, i32, i32, vp],
    "hamt_scatter_rows_bf16": [vp, vp, vp, i32, i32, vp],
    "hamt_gather_rows_pad_bf16": [vp, ll, vp, vp, ll, i32, vp],
}
_RESTYPES = {"hamt_last_error": C.c_char_p, "hamt_launch_count": ll}


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"hamt_b200: native library not found at {LIB_PATH}; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / eager fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export a declared symbol
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, i32)
    if lib.hamt_abi_version() != 2:
        raise RuntimeError("hamt_b200: ABI version mismatch between _lib.py and libhamt_b200.so")
    _lib = lib
    return lib


def last_error() -> str:
    return load().hamt_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"hamt_b200.{what} failed (status {rc}): {last_error()}")


def launch_count() -> int:
    return int(load().hamt_launch_count())
