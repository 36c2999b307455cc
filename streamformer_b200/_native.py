"""ctypes binding of the C ABI declared in include/streamformer_b200.h.

This is the only place the Python host touches native code.  There is no CPU or eager fallback: if
the shared library is missing (not built) importing the op wrappers raises, and every call checks
the returned sf_status and raises ``NativeError`` with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstreamformer_b200.so")

SF_BF16, SF_F16, SF_F32, SF_U8, SF_U8_HWC = 0, 1, 2, 3, 4
SF_ACT_NONE, SF_ACT_GELU, SF_ACT_GELU_TANH = 0, 1, 2
SF_ROW_IDENTITY, SF_ROW_BTN_TO_BNT, SF_ROW_BNT_TO_BTN = 0, 1, 2

# every symbol include/streamformer_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "sf_last_error", "sf_version", "sf_launch_count", "sf_set_option", "sf_profile", "sf_profile_collect", "sf_profile_collect_phases",
    "sf_create", "sf_destroy", "sf_bind_weights", "sf_set_pos_embed", "sf_set_pixel_norm",
    "sf_workspace_bytes", "sf_forward",
    "sf_kv_create", "sf_kv_reset", "sf_kv_destroy", "sf_kv_seq_len", "sf_kv_capacity", "sf_kv_advance", "sf_kv_graph_launches", "sf_forward_stream",
    "sf_embed_forward", "sf_layer_forward", "sf_encoder_forward", "sf_final_norm", "sf_head_forward",
    "sf_op_gemm", "sf_op_layernorm", "sf_op_im2col", "sf_op_temporal_attention", "sf_op_temporal_decode", "sf_op_kv_append",
    "sf_export_packed", "sf_op_transpose", "sf_op_colsum", "sf_op_ln_backward", "sf_op_ln_affine_backward", "sf_op_gelu", "sf_op_gelu_backward",
    "sf_op_gate_backward", "sf_op_wfold_finish", "sf_op_wgrad", "sf_op_wgrad_splits", "sf_op_wfold_finish_partials", "sf_op_embed_table_grad", "sf_op_rowperm", "sf_op_attention_backward",
    "sf_op_pool_attention_backward",
    "sf_op_spatial_attention", "sf_op_siglip_head", "sf_op_l2norm_backward", "sf_op_pool_attention", "sf_op_pool_probe", "sf_op_rowstats", "sf_op_gemm_stats_parts",
]


class NativeError(RuntimeError):
    pass


class SfConfig(C.Structure):
    _fields_ = [
        ("image_size", C.c_int), ("patch_size", C.c_int), ("num_channels", C.c_int), ("num_frames", C.c_int),
        ("hidden_size", C.c_int), ("num_hidden_layers", C.c_int), ("num_attention_heads", C.c_int),
        ("intermediate_size", C.c_int), ("hidden_act", C.c_int), ("layer_norm_eps", C.c_float),
        ("causal_temporal", C.c_int), ("dtype", C.c_int), ("fold_temporal_proj", C.c_int),
    ]


class SfWeightDesc(C.Structure):
    _fields_ = [
        ("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int), ("ndim", C.c_int),
        ("shape", C.c_int64 * 4),
    ]


class SfGemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("act", C.c_int), ("residual", C.c_void_p), ("ldr", C.c_int), ("gate", C.c_void_p),
        ("row_map", C.c_int), ("T", C.c_int), ("S", C.c_int), ("pos", C.c_void_p), ("time_emb", C.c_void_p),
        ("time_len", C.c_int), ("time_total", C.c_int), ("time_off", C.c_int),
        ("ln_stats", C.c_void_p), ("ln_parts", C.c_int), ("ln_colsum", C.c_void_p), ("ln_eps", C.c_float),
        ("stats_out", C.c_void_p),
    ]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (building it is `python -m streamformer_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build the sm_100a extension first (python -m streamformer_b200.build). "
            "There is no CPU / eager fallback for the encoder."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    lib.sf_last_error.restype = C.c_char_p
    lib.sf_version.restype = C.c_char_p
    lib.sf_launch_count.restype = C.c_uint64
    lib.sf_set_option.argtypes = [C.c_char_p, i]
    lib.sf_profile.argtypes = [i]
    lib.sf_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_longlong), i]
    lib.sf_profile_collect_phases.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), i]
    lib.sf_create.argtypes = [C.POINTER(SfConfig), i, C.POINTER(vp)]
    lib.sf_destroy.argtypes = [vp]
    lib.sf_bind_weights.argtypes = [vp, vp, C.POINTER(SfWeightDesc), i]
    lib.sf_set_pos_embed.argtypes = [vp, vp, vp, i]
    lib.sf_set_pixel_norm.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), i]
    lib.sf_workspace_bytes.argtypes = [vp, i, i, i, i, C.POINTER(C.c_size_t)]
    lib.sf_forward.argtypes = [vp, vp, vp, i, i, i, i, i, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, C.c_size_t]
    lib.sf_kv_create.argtypes = [vp, i, i, i, i, C.POINTER(vp)]
    lib.sf_kv_reset.argtypes = [vp]
    lib.sf_kv_destroy.argtypes = [vp]
    lib.sf_kv_seq_len.argtypes = [vp]
    lib.sf_kv_capacity.argtypes = [vp]
    lib.sf_kv_advance.argtypes = [vp, i]
    lib.sf_kv_graph_launches.argtypes = [vp]
    lib.sf_forward_stream.argtypes = [vp, vp, vp, vp, i, i, i, i, i, vp, vp, C.POINTER(vp), vp, C.c_size_t]
    lib.sf_embed_forward.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, vp, vp, C.c_size_t]
    lib.sf_layer_forward.argtypes = [vp, vp, i, vp, vp, i, i, i, vp, vp, vp, C.c_size_t]
    lib.sf_encoder_forward.argtypes = [vp, vp, vp, i, i, i, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, C.c_size_t]
    lib.sf_final_norm.argtypes = [vp, vp, vp, i, i, i, vp]
    lib.sf_head_forward.argtypes = [vp, vp, vp, i, i, vp, vp, C.c_size_t]
    lib.sf_op_gemm.argtypes = [vp, i, vp, i, vp, i, vp, i, i, i, i, C.POINTER(SfGemmEpilogue)]
    lib.sf_op_layernorm.argtypes = [vp, i, vp, i, vp, vp, f, vp, i, i, i, i, i, i]
    lib.sf_op_im2col.argtypes = [vp, i, vp, i, vp, i, i, i, i, i]
    lib.sf_op_temporal_attention.argtypes = [vp, i, vp, i, vp, vp, i, vp, i, i, i, i, i, i, i, f]
    lib.sf_op_temporal_decode.argtypes = [vp, i, vp, i, vp, vp, i, vp, i, i, i, i, f]
    lib.sf_op_kv_append.argtypes = [vp, i, vp, i, vp, vp, i, i, i, i, i]
    lib.sf_op_spatial_attention.argtypes = [vp, i, vp, i, vp, i, i, i, i, i, f, vp]
    lib.sf_op_siglip_head.argtypes = [vp, i, vp, i, vp, i, i, i, i, vp, vp, i, i, vp, i, f, vp, i, vp, vp, i, vp]
    lib.sf_op_l2norm_backward.argtypes = [vp, i, vp, i, vp, i, vp, vp, i, i, i]
    ll = C.c_longlong
    lib.sf_export_packed.argtypes = [vp, vp, i, C.c_char_p, i, vp, C.c_size_t]
    lib.sf_op_transpose.argtypes = [vp, i, vp, i, vp, i, i, i]
    lib.sf_op_colsum.argtypes = [vp, i, vp, i, i, i, vp]
    lib.sf_op_ln_backward.argtypes = [vp, i, vp, i, vp, i, f, vp, i, vp, i, i, i]
    lib.sf_op_ln_affine_backward.argtypes = [vp, i, vp, i, vp, i, vp, f, vp, i, i, i, i, i, i, vp, vp]
    lib.sf_op_gelu.argtypes = [vp, i, vp, vp, ll, i]
    lib.sf_op_gelu_backward.argtypes = [vp, i, vp, vp, ll, i]
    lib.sf_op_gate_backward.argtypes = [vp, i, vp, vp, vp, vp, ll, vp]
    lib.sf_op_wfold_finish.argtypes = [vp, i, vp, i, vp, i, vp, vp, vp, vp, i, i, i, i, vp, vp]
    lib.sf_op_wgrad_splits.argtypes = [i, i, i]
    lib.sf_op_wgrad.argtypes = [vp, i, vp, i, vp, i, i, i, i, vp]
    lib.sf_op_wfold_finish_partials.argtypes = [vp, i, vp, i, vp, i, vp, vp, vp, vp, i, i, i, i, vp, vp]
    lib.sf_op_embed_table_grad.argtypes = [vp, i, vp, i, i, i, i, i, i, vp, vp]
    lib.sf_op_rowperm.argtypes = [vp, vp, vp, ll, i, i, i, i]
    lib.sf_op_attention_backward.argtypes = [vp, i, i, vp, i, vp, i, vp, i, vp, i, i, i, i, i, i, f]
    lib.sf_op_pool_attention_backward.argtypes = [vp, i, vp, i, vp, vp, i, vp, i, vp, i, i, i]
    lib.sf_op_rowstats.argtypes = [vp, i, vp, i, i, i, vp]
    lib.sf_op_gemm_stats_parts.argtypes = [i, i]
    lib.sf_op_pool_attention.argtypes = [vp, i, vp, i, vp, vp, i, i, i, i]
    lib.sf_op_pool_probe.argtypes = [vp, i, vp, i, vp, vp, vp, vp, i, i, i, i]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("sf_last_error", "sf_version", "sf_launch_count", "sf_kv_graph_launches"):
            fn.restype = C.c_int
    lib.sf_kv_graph_launches.restype = C.c_longlong
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sf_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'streamformer_b200'} failed (status {rc}): {msg}")


def set_option(name: str, value: int) -> None:
    check(load().sf_set_option(name.encode(), int(value)), "sf_set_option")


def launch_count() -> int:
    return int(load().sf_launch_count())


def version() -> str:
    return load().sf_version().decode()


PROFILE_CLASSES = ["gemm", "layernorm", "im2col", "temporal_attention", "spatial_attention", "pool_attention",
                   "kv_append", "other"]


def profile(enable) -> None:
    """False/0 off, True/1 per-kernel events, 2 per-phase events (kernels back to back)."""
    load().sf_profile(int(enable))


PROFILE_PHASES = ["embed", "attention_block", "mlp", "head"]


def profile_collect_phases() -> dict:
    n = len(PROFILE_PHASES)
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(load().sf_profile_collect_phases(ms, cnt, n), "sf_profile_collect_phases")
    return {PROFILE_PHASES[k]: {"ms": ms[k], "count": int(cnt[k])} for k in range(n)}


def profile_collect() -> dict:
    n = len(PROFILE_CLASSES)
    ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(load().sf_profile_collect(ms, fl, by, cnt, n), "sf_profile_collect")
    return {PROFILE_CLASSES[k]: {"ms": ms[k], "flops": fl[k], "bytes": by[k], "launches": int(cnt[k])} for k in range(n)}


def ptr_array(ptrs: Optional[Sequence[int]]):
    if ptrs is None:
        return None
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return arr
