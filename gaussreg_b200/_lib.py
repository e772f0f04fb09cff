"""ctypes loader for the in-tree sm_100a library (include/gaussreg_b200.h).

The product path has no CPU fallback: if the shared library is missing or the call fails, the
caller gets a RuntimeError.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgaussreg_b200.so")

_lib = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int
_f32 = ctypes.c_float
_sz = ctypes.c_size_t

class LayerWeights(ctypes.Structure):
    """gr_layer_weights (include/gaussreg_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("wq", "bq", "wk", "bk", "wv", "bv", "wp", "bp", "wo", "bo", "ln1_g", "ln1_b",
                                               "w1", "b1", "w2", "b2", "ln2_g", "ln2_b", "wqkv", "bqkv", "w1t", "w2t")] + [("is_self", ctypes.c_int)]


class UnaryWeights(ctypes.Structure):
    """gr_unary_weights (include/gaussreg_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("weight", "weight_packed", "bias", "gn_weight", "gn_bias")] + \
               [(n, ctypes.c_int) for n in ("in_channels", "out_channels", "leaky_relu", "split_k")] + \
               [(n, ctypes.c_void_p) for n in ("weight_packed_lo", "weight_packed_hi", "weight_packed16", "weight_packed16_lo",
                                               "weight_packed16_hi")] + \
               [(n, ctypes.c_float) for n in ("inv_scale16", "inv_scale16_lo", "inv_scale16_hi")]


class PyramidSearch(ctypes.Structure):
    """gr_pyramid_search."""
    _fields_ = [("query_stage", ctypes.c_int), ("support_stage", ctypes.c_int), ("radius", ctypes.c_float),
                ("limit", ctypes.c_int64), ("out_idx", ctypes.c_void_p), ("out_max_count", ctypes.c_void_p)]


class KPConvWeights(ctypes.Structure):
    """gr_kpconv_weights."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("weights", "weights_kmajor", "weights_kmajor_packed", "bias", "kernel_points")] + \
               [("sigma", ctypes.c_float), ("in_channels", ctypes.c_int), ("out_channels", ctypes.c_int),
                ("weights_kmajor_packed16", ctypes.c_void_p), ("inv_scale16", ctypes.c_float)]


class BlockWeights(ctypes.Structure):
    """gr_block_weights."""
    _fields_ = [("kind", ctypes.c_int), ("strided", ctypes.c_int), ("unary1", UnaryWeights), ("conv", KPConvWeights),
                ("gn_conv_weight", ctypes.c_void_p), ("gn_conv_bias", ctypes.c_void_p), ("unary2", UnaryWeights),
                ("shortcut", UnaryWeights)]


FPN_STAGES, FPN_BLOCKS = 5, 14


class FpnWeights(ctypes.Structure):
    """gr_fpn_weights."""
    _fields_ = [("blocks", BlockWeights * FPN_BLOCKS), ("decoder4", UnaryWeights), ("decoder3", UnaryWeights),
                ("decoder2", UnaryWeights), ("group_norm", ctypes.c_int), ("eps", ctypes.c_float)]


class Pyramid(ctypes.Structure):
    """gr_pyramid."""
    _fields_ = [("points", ctypes.c_void_p * FPN_STAGES), ("n_points", ctypes.c_int * FPN_STAGES)]
    for _k in ("neighbors", "subsampling", "upsampling"):
        _fields_ += [(_k, ctypes.c_void_p * FPN_STAGES), (_k + "_w", ctypes.c_int * FPN_STAGES), (_k + "_ld", ctypes.c_int64 * FPN_STAGES)]
    del _k


_SIGNATURES = {
    "gr_version": (ctypes.c_char_p, []),
    "gr_last_error": (ctypes.c_char_p, []),
    "gr_launch_count": (_i64, []),
    "gr_grid_subsample_workspace_size": (_sz, [_i64, _i32]),
    "gr_grid_subsample_chain": (_i32, [_vp, _vp, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gr_grid_subsample": (_i32, [_vp, _vp, _i32, _i64, _f32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_radius_neighbors_workspace_size": (_sz, [_i64, _i64, _i32]),
    "gr_radius_neighbors": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _f32, _vp, _i64, _vp, _vp, _sz, _vp]),
    "gr_radius_neighbors_cached": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _f32, _vp, _i64, _vp, _vp, _sz, _i32, _vp]),
    "gr_gemm": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _f32, _vp, _vp,
                       _vp, _i64, _i64, _i32, _vp]),
    "gr_linear_packed": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "gr_linear_packed16": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _f32, _vp, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "gr_set_gemm_mode": (None, [_i32]),
    "gr_get_gemm_mode": (_i32, []),
    "gr_last_gemm_path": (_i32, []),
    "gr_kpconv_aggregate_workspace_size": (_sz, [_i64]),
    "gr_kpconv_aggregate": (_i32, [_vp, _i32, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _i32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "gr_unary_block_workspace_size": (_sz, [_i64, _i32, _i32]),
    "gr_unary_block": (_i32, [_vp, _vp, _i64, _i32, _f32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gr_kpconv_block_workspace_size": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "gr_kpconv_block": (_i32, [_vp, _vp, _vp, _i32, _f32, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "gr_kpconv_fpn_workspace_size": (_sz, [_vp, _vp]),
    "gr_kpconv_fpn": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_kpconv_fpn_from": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_group_norm_workspace_size": (_sz, [_i64, _i32]),
    "gr_group_norm": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gr_layer_norm_add": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp]),
    "gr_maxpool": (_i32, [_vp, _i32, _i32, _vp, _i32, _i64, _i32, _vp, _vp]),
    "gr_upsample_concat": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _i32, _i32, _vp, _vp]),
    "gr_gather_rows": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "gr_point_to_node_workspace_size": (_sz, [_i64, _i64]),
    "gr_point_to_node_partition": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_embedding_indices": (_i32, [_vp, _i32, _f32, _f32, _i32, _vp, _vp, _vp, _vp]),
    "gr_sinusoid_rows": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "gr_embedding_combine": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "gr_pack_weight_tf32x3": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "gr_structure_embedding_fused": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gr_radius_pyramid": (_i32, [_vp, _vp, _i32, _i32, _i64, _vp, ctypes.c_size_t, _vp, _vp, _i32, ctypes.c_uint32, _vp]),
    "gr_packed_weight_f16_bytes": (ctypes.c_size_t, [_i32, _i32]),
    "gr_pack_weight_f16x3": (_i32, [_vp, _i32, _i32, _f32, _vp, _vp]),
    "gr_gemm_f16_overflow_ptr": (_i32, [_vp]),
    "gr_pack_weight_f16x2": (_i32, [_vp, _i32, _i32, _f32, _vp, _vp]),
    "gr_structure_embedding_fused_f16": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp]),
    "gr_structure_embedding_table_floats": (_i64, [_i32, _f32]),
    "gr_structure_embedding_build_table": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _f32, _vp, _vp]),
    "gr_structure_embedding_tabulated": (_i32, [_vp, _vp, _i64, _i32, _vp, _f32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gr_structure_embedding_points": (_i32, [_vp, _i32, _f32, _f32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gr_rpe_attention_probs": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "gr_rpe_attention_probs_ld": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "gr_softmax_rows": (_i32, [_vp, _i64, _i32, _vp]),
    "gr_l2_normalize_rows": (_i32, [_vp, _i64, _i32, _f32, _vp, _vp]),
    "gr_conditional_transformer_workspace_size": (_sz, [_i32, _i32, _i32, _i32]),
    "gr_conditional_transformer": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    "gr_superpoint_matching_workspace_size": (_sz, [_i32, _i32, _i32]),
    "gr_superpoint_matching": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_sinkhorn": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp]),
    "gr_lgr_workspace_size": (_sz, [_i32, _i32, _i32]),
    "gr_local_global_registration": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _i32, _f32, _i32, _i32,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_gaussian_transform": (_i32, [_vp, _i64, _i64, _vp, _f32, _f32, _vp, _vp, _vp, _i64, _vp]),
    "gr_farthest_point_sample_workspace_size": (_sz, [_i64]),
    "gr_farthest_point_sample": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _sz, _vp]),
    "gr_similarity_ransac_workspace_size": (_sz, [_i32]),
    "gr_similarity_ransac": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, ctypes.c_uint64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_weighted_procrustes": (_i32, [_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp]),
    "gr_column_order_stats_workspace_size": (_sz, [_i32]),
    "gr_column_order_stats": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gr_gaussian_select_workspace_size": (_sz, [_i64]),
    "gr_gaussian_select": (_i32, [_vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gr_gather_points_stats": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "gr_gaussian_features": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    "gr_points_normalize": (_i32, [_vp, _i64, _vp, _f32, _i32, _vp]),
    "gr_pair_major_rows": (_i32, [_vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp]),
    "gr_pair_major_table": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
}

_STATUS = {-1: "bad argument", -2: "workspace too small", -3: "capacity overflow", -4: "CUDA error"}


def exported_symbols():
    """Every symbol include/gaussreg_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def register(name, restype, argtypes):
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m gaussreg_b200.build` "
                "(gaussreg_b200 has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = restype, argtypes
    return _lib


def check(status, what):
    if status != 0:
        msg = _STATUS.get(status, f"status {status}")
        detail = lib().gr_last_error().decode() if status == -4 else ""
        raise RuntimeError(f"gaussreg_b200.{what} failed: {msg} {detail}".strip())


def launch_count():
    return int(lib().gr_launch_count())
