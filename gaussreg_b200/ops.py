"""Functional front-ends of the C ABI (include/gaussreg_b200.h) on torch CUDA tensors.

PyTorch is used for device memory and the current stream only; every function below launches the
library's own sm_100a kernels and raises RuntimeError if the call fails.  Mirrors
geotransformer/modules/ops/*.py where the reference has an equivalent.
"""
import math
import os

import torch

from . import _lib
from .ext import _stream, _workspace, radius_neighbors_device, grid_subsample_device  # noqa: F401

_F32 = torch.float32


def _req(t, dtype=_F32):
    if not t.is_cuda:
        raise RuntimeError("gaussreg_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _u8(mask):
    """bool / uint8 mask tensor -> contiguous uint8 view (no copy for bool)."""
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return _req(mask, torch.uint8)


ACT = {None: 0, "none": 0, "relu": 1, "leaky_relu": 2}


def gemm(a, b, trans_b=True, bias=None, alpha=1.0, row_div=None, residual=None, act=None, out=None):
    """2-D product with fused epilogue.  a (M,K); b (N,K) if trans_b else (K,N).  Rows of a / b / out may be
    strided views (last dim contiguous)."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    assert (b.shape[1] if trans_b else b.shape[0]) == K
    if out is None:
        out = torch.empty((M, N), dtype=_F32, device=a.device)
    assert out.stride(1) == 1
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1
    st = _lib.lib().gr_gemm(a.data_ptr(), a.stride(0), 0, b.data_ptr(), b.stride(0), 0, int(trans_b), out.data_ptr(),
                            out.stride(0), 0, M, N, K, 1, float(alpha), _ptr(bias), _ptr(row_div), _ptr(residual),
                            residual.stride(0) if residual is not None else 0, 0, ACT[act], _stream())
    _lib.check(st, "gemm")
    return out


def gemm_batched(a_ptr, lda, sa, b_ptr, ldb, sb, trans_b, c_ptr, ldc, sc, M, N, K, batch, alpha=1.0, bias=None):
    """Raw strided-batched product (pointers + element strides); used for per-head attention products."""
    st = _lib.lib().gr_gemm(a_ptr, lda, sa, b_ptr, ldb, sb, int(trans_b), c_ptr, ldc, sc, M, N, K, batch, float(alpha),
                            _ptr(bias), None, None, 0, 0, 0, _stream())
    _lib.check(st, "gemm_batched")


def linear(x, weight, bias=None, act=None, residual=None, out=None, row_div=None):
    """nn.Linear forward: x (rows, in) @ weight (out, in)^T + bias.  Large products take the tensor-core kernel
    with the weight pre-packed (cached on the parameter)."""
    M, K = x.shape
    N = weight.shape[0]
    if M < 2048 or not weight.is_contiguous() or _lib.lib().gr_get_gemm_mode() != 1:
        return gemm(x, weight, True, bias=bias, act=act, residual=residual, out=out, row_div=row_div)
    assert x.stride(1) == 1 and weight.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=_F32, device=x.device)
    if _gemm_f16():
        pk16, inv16 = packed_weight_f16x3(weight)
        st = _lib.lib().gr_linear_packed16(x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0),
                                           packed_weight_tf32x3(weight).data_ptr(), pk16.data_ptr(), inv16, out.data_ptr(),
                                           out.stride(0), M, N, K, 1.0, _ptr(bias), _ptr(row_div), _ptr(residual),
                                           residual.stride(0) if residual is not None else 0, ACT[act], _stream())
        _lib.check(st, "linear_packed16")
        return out
    st = _lib.lib().gr_linear_packed(x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0),
                                     packed_weight_tf32x3(weight).data_ptr(), out.data_ptr(), out.stride(0), M, N, K, 1.0,
                                     _ptr(bias), _ptr(row_div), _ptr(residual),
                                     residual.stride(0) if residual is not None else 0, ACT[act], _stream())
    _lib.check(st, "linear_packed")
    return out


def kpconv_aggregate(s_feats, q_points, s_points, neighbor_indices, kernel_points, sigma):
    """K1 first half -> (A (M, 15*C), row_div (M))."""
    s_feats, q_points, s_points = _req(s_feats), _req(q_points), _req(s_points)
    kernel_points = _req(kernel_points)
    assert neighbor_indices.dtype == torch.int64 and neighbor_indices.stride(1) == 1
    M, H = neighbor_indices.shape
    Ns, C = s_feats.shape
    L = _lib.lib()
    A = torch.empty((M, kernel_points.shape[0] * C), dtype=_F32, device=s_feats.device)
    row_div = torch.empty((M,), dtype=_F32, device=s_feats.device)
    ws = _workspace(L.gr_kpconv_aggregate_workspace_size(Ns), s_feats.device)
    st = L.gr_kpconv_aggregate(s_feats.data_ptr(), C, q_points.data_ptr(), s_points.data_ptr(), neighbor_indices.data_ptr(),
                               H, neighbor_indices.stride(0), M, Ns, kernel_points.data_ptr(), kernel_points.shape[0],
                               float(sigma), A.data_ptr(), row_div.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "kpconv_aggregate")
    return A, row_div


def _kmajor_weights(weights):
    """(K, C, C_out) KPConv weights -> (C_out, K*C) K-major copy for the tensor-core GEMM.  Static weights: one
    transpose per parameter version (load-time plumbing), cached ON the parameter object so that the cache can
    never outlive or alias the tensor it was made from."""
    cached = getattr(weights, "_gr_kmajor", None)
    if cached is None or cached[0] != weights._version or cached[1].device != weights.device:
        K, C, Co = weights.shape
        wt = weights.detach().reshape(K * C, Co).t().contiguous()
        cached = (weights._version, wt)
        try:
            weights._gr_kmajor = cached
        except AttributeError:
            pass
    return cached[1]


def kpconv(s_feats, q_points, s_points, neighbor_indices, weights, bias, kernel_points, sigma):
    """KPConv.forward (kpconv.py:79-122): (M, C_out)."""
    A, row_div = kpconv_aggregate(s_feats, q_points, s_points, neighbor_indices, kernel_points, sigma)
    K, C, Co = weights.shape
    if _lib.lib().gr_get_gemm_mode() == 1 and (K * C) % 4 == 0:
        return linear(A, _kmajor_weights(weights), bias=bias, row_div=row_div)
    return gemm(A, weights.view(K * C, Co), False, bias=bias, row_div=row_div)


# ------------------------------------------------------------------------------------------------
# native KPConv blocks / FPN: parameter structs for the C ABI (include/gaussreg_b200.h, csrc/backbone.cu)
# ------------------------------------------------------------------------------------------------
def packed_weight_f16x3(weight):
    """(N,K) static weight -> fp16 hi/lo tiles for the kind::f16 tensor-core GEMM, pre-scaled by a power of two so that the
    largest |w| lands in [4, 8).  Returns (buffer, 1/scale); cached on the tensor."""
    cached = getattr(weight, "_gr_packed16g", None)
    if cached is None or cached[0] != weight._version or cached[1].device != weight.device:
        N, K = weight.shape
        wmax = float(weight.detach().abs().max())
        exp = 0 if not (wmax > 0.0 and math.isfinite(wmax)) else max(-14, min(14, math.floor(math.log2(8.0 / wmax))))
        scale = 2.0 ** exp
        L = _lib.lib()
        out = torch.empty((L.gr_packed_weight_f16_bytes(N, K),), dtype=torch.uint8, device=weight.device)
        st = L.gr_pack_weight_f16x3(weight.detach().contiguous().data_ptr(), N, K, scale, out.data_ptr(), _stream())
        _lib.check(st, "pack_weight_f16x3")
        cached = (weight._version, out, 1.0 / scale)
        try:
            weight._gr_packed16g = cached
        except AttributeError:
            pass
    return cached[1], cached[2]


def _gemm_f16():
    """GAUSSREG_GEMM_F16=1: backbone products on the kind::f16 persistent kernel (fp16-split operands).  Off by default: it
    passes every parity test but buys only ~5 % on the GEMM shapes of the backbone (0.04 ms per pair) -- with 64-wide
    k-blocks the operand rings that fit in shared memory are too shallow to hide the TMA latency -- and fp16's range is a
    constraint the TF32 path does not have."""
    return os.environ.get("GAUSSREG_GEMM_F16", "0") == "1"


def _fill_unary(dst, mlp, norm, leaky, keep, split_k=0):
    """gr_unary_weights from an nn.Linear (+ optional GroupNorm wrapper).  split_k > 0 (decoder blocks): also pack the
    two column slices of the weight, see gr_unary_weights.split_k."""
    w = mlp.weight
    dst.weight = w.data_ptr()
    dst.split_k, dst.weight_packed_lo, dst.weight_packed_hi = 0, None, None
    dst.weight_packed16, dst.weight_packed16_lo, dst.weight_packed16_hi = None, None, None
    dst.inv_scale16 = dst.inv_scale16_lo = dst.inv_scale16_hi = 1.0
    if 0 < split_k < w.shape[1] and split_k % 32 == 0 and (w.shape[1] - split_k) % 32 == 0 and os.environ.get("GAUSSREG_DECODER_SPLIT", "1") != "0":
        lo, hi = w.detach()[:, :split_k].contiguous(), w.detach()[:, split_k:].contiguous()
        plo, phi = packed_weight_tf32x3(lo), packed_weight_tf32x3(hi)
        keep.extend([lo, hi, plo, phi])
        dst.split_k, dst.weight_packed_lo, dst.weight_packed_hi = split_k, plo.data_ptr(), phi.data_ptr()
        if _gemm_f16():
            (qlo, ilo), (qhi, ihi) = packed_weight_f16x3(lo), packed_weight_f16x3(hi)
            keep.extend([qlo, qhi])
            dst.weight_packed16_lo, dst.weight_packed16_hi, dst.inv_scale16_lo, dst.inv_scale16_hi = qlo.data_ptr(), qhi.data_ptr(), ilo, ihi
    if w.is_contiguous() and w.shape[1] % 4 == 0:
        pk = packed_weight_tf32x3(w)
        keep.append(pk)
        dst.weight_packed = pk.data_ptr()
        if _gemm_f16():
            pk16, inv16 = packed_weight_f16x3(w)
            keep.append(pk16)
            dst.weight_packed16, dst.inv_scale16 = pk16.data_ptr(), inv16
    else:
        dst.weight_packed = None
    dst.bias = mlp.bias.data_ptr() if mlp.bias is not None else None
    if norm is not None:
        dst.gn_weight, dst.gn_bias = norm.norm.weight.data_ptr(), norm.norm.bias.data_ptr()
    else:
        dst.gn_weight, dst.gn_bias = None, None
    dst.in_channels, dst.out_channels, dst.leaky_relu = w.shape[1], w.shape[0], int(bool(leaky))


def _fill_kpconv(dst, conv, keep):
    K, C, Co = conv.weights.shape
    dst.weights = conv.weights.data_ptr()
    if (K * C) % 4 == 0:
        wk = _kmajor_weights(conv.weights)
        pk = packed_weight_tf32x3(wk)
        keep += [wk, pk]
        dst.weights_kmajor, dst.weights_kmajor_packed = wk.data_ptr(), pk.data_ptr()
        dst.weights_kmajor_packed16, dst.inv_scale16 = None, 1.0
        if _gemm_f16():
            pk16, inv16 = packed_weight_f16x3(wk)
            keep.append(pk16)
            dst.weights_kmajor_packed16, dst.inv_scale16 = pk16.data_ptr(), inv16
    else:
        dst.weights_kmajor, dst.weights_kmajor_packed = None, None
        dst.weights_kmajor_packed16, dst.inv_scale16 = None, 1.0
    dst.bias = conv.bias.data_ptr() if conv.bias is not None else None
    dst.kernel_points = conv.kernel_points.data_ptr()
    dst.sigma, dst.in_channels, dst.out_channels = float(conv.sigma), C, Co


def _drop_tensor_list(module, *_):
    module.__dict__.pop("_gr_tensors", None)


def _module_key(module):
    """(version, storage) of every parameter / buffer under `module`.  Walking the module tree costs 0.5 ms for the
    backbone (204 tensors) -- per step, on the host's critical path -- so the flat tensor list is kept on the module;
    load_state_dict (which may re-assign Parameters) and invalidate_weight_caches drop it."""
    ts = module.__dict__.get("_gr_tensors")
    if ts is None:
        ts = list(module.parameters()) + list(module.buffers())
        if "_gr_tensors_hooked" not in module.__dict__:
            module.__dict__["_gr_tensors_hooked"] = True
            module.register_load_state_dict_post_hook(_drop_tensor_list)
        module.__dict__["_gr_tensors"] = ts
    return tuple([(t._version, t.data_ptr()) for t in ts])


def _cached_struct(module, build):
    """C parameter struct of `module`, rebuilt when any parameter / buffer changes version or storage."""
    key = _module_key(module)
    cached = module.__dict__.get("_gr_native")
    if cached is None or cached[0] != key:
        keep = []
        cached = (key, build(keep), keep)
        module.__dict__["_gr_native"] = cached
    return cached[1]


def invalidate_weight_caches(module):
    """Drop every derived weight image (packed / transposed / fused copies and native parameter structs) under
    `module`.  Needed only after IN-PLACE edits through `param.data`, which do not bump the version counter the
    caches are keyed on; load_state_dict / optimizer steps are detected automatically."""
    for m in module.modules():
        m.__dict__.pop("_gr_native", None)
        m.__dict__.pop("_gr_tensors", None)
    for p in list(module.parameters()) + list(module.buffers()):
        for attr in ("_gr_packed", "_gr_packed16", "_gr_packed16g", "_gr_kmajor", "_gr_qkv", "_gr_t", "_gr_t1_table"):
            if hasattr(p, attr):
                try:
                    delattr(p, attr)
                except AttributeError:
                    pass


def unary_block(block, x, add=None, act_after_add=None):
    """UnaryBlock / LastUnaryBlock forward (modules.py:53-101) in one C-ABI call: the GroupNorm statistics come out
    of the Linear's epilogue."""
    import ctypes
    x = _req(x)
    norm = getattr(block, "norm", None)
    leaky = getattr(block, "leaky_relu", None) is not None

    def build(keep):
        w = _lib.UnaryWeights()
        _fill_unary(w, block.mlp, norm, leaky, keep)
        return w

    w = _cached_struct(block, build)
    rows = x.shape[0]
    out = torch.empty((rows, w.out_channels), dtype=_F32, device=x.device)
    groups = norm.num_groups if norm is not None else 1
    eps = norm.norm.eps if norm is not None else 1e-5
    if add is not None:
        add = _req(add)
    L = _lib.lib()
    ws = _workspace(L.gr_unary_block_workspace_size(rows, w.out_channels, groups), x.device)
    st = L.gr_unary_block(ctypes.byref(w), x.data_ptr(), rows, groups, float(eps), _ptr(add), ACT[act_after_add], out.data_ptr(),
                          ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "unary_block")
    return out


def kpconv_block(conv, norm, s_feats, q_points, s_points, neighbor_indices):
    """KPConv [+ GroupNorm + LeakyReLU] (kpconv.py:79-122, modules.py:104-146) in one C-ABI call."""
    import ctypes
    s_feats, q_points, s_points = _req(s_feats), _req(q_points), _req(s_points)
    assert neighbor_indices.dtype == torch.int64 and neighbor_indices.stride(1) == 1

    def build(keep):
        w = _lib.KPConvWeights()
        _fill_kpconv(w, conv, keep)
        return w

    w = _cached_struct(conv, build)
    M, H = neighbor_indices.shape
    Ns = s_feats.shape[0]
    out = torch.empty((M, w.out_channels), dtype=_F32, device=s_feats.device)
    groups = norm.num_groups if norm is not None else 1
    L = _lib.lib()
    ws = _workspace(L.gr_kpconv_block_workspace_size(M, Ns, w.in_channels, w.out_channels, groups), s_feats.device)
    st = L.gr_kpconv_block(ctypes.byref(w), norm.norm.weight.data_ptr() if norm is not None else None,
                           norm.norm.bias.data_ptr() if norm is not None else None, groups,
                           float(norm.norm.eps) if norm is not None else 1e-5, s_feats.data_ptr(), q_points.data_ptr(),
                           s_points.data_ptr(), neighbor_indices.data_ptr(), H, neighbor_indices.stride(0), M, Ns, out.data_ptr(),
                           ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "kpconv_block")
    return out


_FPN_BLOCK_NAMES = ("encoder1_1", "encoder1_2", "encoder2_1", "encoder2_2", "encoder2_3", "encoder3_1", "encoder3_2", "encoder3_3",
                    "encoder4_1", "encoder4_2", "encoder4_3", "encoder5_1", "encoder5_2", "encoder5_3")


def kpconv_fpn(backbone, feats, data_dict, start_block=0):
    """KPConvFPN.forward (backbone.py:164-212) as ONE C-ABI call -> [l2, l3, l4, f5]."""
    import ctypes
    import torch.nn as nn
    feats = _req(feats)

    def build(keep):
        W = _lib.FpnWeights()
        groups, eps = None, 1e-5
        for i, name in enumerate(_FPN_BLOCK_NAMES):
            m, b = getattr(backbone, name), W.blocks[i]
            conv = m.KPConv
            _fill_kpconv(b.conv, conv, keep)
            if hasattr(m, "unary2"):  # ResidualBlock
                b.kind, b.strided = 1, int(m.strided)
                if isinstance(m.unary1, nn.Identity):
                    b.unary1.in_channels = 0
                else:
                    _fill_unary(b.unary1, m.unary1.mlp, m.unary1.norm, True, keep)
                b.gn_conv_weight, b.gn_conv_bias = m.norm_conv.norm.weight.data_ptr(), m.norm_conv.norm.bias.data_ptr()
                _fill_unary(b.unary2, m.unary2.mlp, m.unary2.norm, False, keep)
                if isinstance(m.unary_shortcut, nn.Identity):
                    b.shortcut.in_channels = 0
                else:
                    _fill_unary(b.shortcut, m.unary_shortcut.mlp, m.unary_shortcut.norm, False, keep)
                groups, eps = m.norm_conv.num_groups, m.norm_conv.norm.eps
            else:                     # ConvBlock
                b.kind, b.strided = 0, 0
                b.gn_conv_weight, b.gn_conv_bias = m.norm.norm.weight.data_ptr(), m.norm.norm.bias.data_ptr()
        # decoder inputs are cat[upsampled coarser output, skip]: the coarse part has the previous decoder's (or the last
        # encoder's) width
        c5 = backbone.encoder5_3.unary2.mlp.weight.shape[0] if hasattr(backbone.encoder5_3, "unary2") else 0
        _fill_unary(W.decoder4, backbone.decoder4.mlp, backbone.decoder4.norm, True, keep, split_k=c5)
        _fill_unary(W.decoder3, backbone.decoder3.mlp, backbone.decoder3.norm, True, keep, split_k=backbone.decoder4.mlp.weight.shape[0])
        _fill_unary(W.decoder2, backbone.decoder2.mlp, None, False, keep, split_k=backbone.decoder3.mlp.weight.shape[0])
        W.group_norm, W.eps = int(groups), float(eps)
        return W

    W = _cached_struct(backbone, build)
    P = _lib.Pyramid()
    pts, nb, sub, up = data_dict["points"], data_dict["neighbors"], data_dict["subsampling"], data_dict["upsampling"]
    keep = []
    fields = [(tabs, len(tabs), getattr(P, key), getattr(P, key + "_w"), getattr(P, key + "_ld"))
              for key, tabs in (("neighbors", nb), ("subsampling", sub), ("upsampling", up))]
    P_points, P_n = P.points, P.n_points
    for s in range(_lib.FPN_STAGES):
        pt = _req(pts[s])
        keep.append(pt)
        P_points[s], P_n[s] = pt.data_ptr(), pt.shape[0]
        for tabs, n_tabs, f_ptr, f_w, f_ld in fields:
            if s < n_tabs:
                t = tabs[s]
                assert t.dtype == torch.int64 and t.stride(1) == 1
                f_ptr[s], f_w[s], f_ld[s] = t.data_ptr(), t.shape[1], t.stride(0)
    dev = feats.device
    n = [p.shape[0] for p in pts]
    outs = [torch.empty((n[1], W.decoder2.out_channels), dtype=_F32, device=dev),
            torch.empty((n[2], W.decoder3.out_channels), dtype=_F32, device=dev),
            torch.empty((n[3], W.decoder4.out_channels), dtype=_F32, device=dev),
            torch.empty((n[4], W.blocks[13].unary2.out_channels), dtype=_F32, device=dev)]
    L = _lib.lib()
    ws = _workspace(L.gr_kpconv_fpn_workspace_size(ctypes.byref(W), ctypes.byref(P)), dev)
    st = L.gr_kpconv_fpn_from(ctypes.byref(W), ctypes.byref(P), int(start_block), feats.data_ptr(), outs[0].data_ptr(),
                              outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "kpconv_fpn")
    return outs


def group_norm(x, groups, gamma, beta, eps=1e-5, add=None, act=None, out=None):
    x = _req(x)
    n, C = x.shape
    if out is None:
        out = torch.empty_like(x)
    if add is not None:
        add = _req(add)
    L = _lib.lib()
    ws = _workspace(L.gr_group_norm_workspace_size(n, groups), x.device)
    st = L.gr_group_norm(x.data_ptr(), n, C, groups, gamma.data_ptr(), beta.data_ptr(), float(eps), _ptr(add), ACT[act],
                         out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "group_norm")
    return out


def layer_norm_add(a, b, gamma, beta, eps=1e-5):
    a = _req(a)
    if b is not None:
        b = _req(b)
    out = torch.empty_like(a)
    st = _lib.lib().gr_layer_norm_add(a.data_ptr(), _ptr(b), a.shape[0], a.shape[1], gamma.data_ptr(), beta.data_ptr(),
                                      float(eps), out.data_ptr(), _stream())
    _lib.check(st, "layer_norm_add")
    return out


def maxpool(x, neighbor_indices):
    """kpconv/functional.py:54-67."""
    x = _req(x)
    assert neighbor_indices.dtype == torch.int64 and neighbor_indices.stride(1) == 1
    M, H = neighbor_indices.shape
    out = torch.empty((M, x.shape[1]), dtype=_F32, device=x.device)
    st = _lib.lib().gr_maxpool(x.data_ptr(), x.shape[0], x.shape[1], neighbor_indices.data_ptr(), H,
                               neighbor_indices.stride(0), M, out.data_ptr(), _stream())
    _lib.check(st, "maxpool")
    return out


def upsample_concat(coarse, upsample_indices, skip):
    """cat([nearest_upsample(coarse, idx), skip], dim=1) (functional.py:6-22, backbone.py:195-197)."""
    coarse, skip = _req(coarse), _req(skip)
    assert upsample_indices.dtype == torch.int64
    M = upsample_indices.shape[0]
    out = torch.empty((M, coarse.shape[1] + skip.shape[1]), dtype=_F32, device=coarse.device)
    st = _lib.lib().gr_upsample_concat(coarse.data_ptr(), coarse.shape[0], coarse.shape[1], upsample_indices.data_ptr(),
                                       upsample_indices.stride(0), skip.data_ptr(), skip.shape[1], M, out.data_ptr(), _stream())
    _lib.check(st, "upsample_concat")
    return out


def nearest_upsample(x, upsample_indices):
    """functional.py:6-22."""
    idx = upsample_indices[:, 0].contiguous()
    return gather_rows(x, idx)


def gather_rows(x, index):
    """index_select(cat([x, 0]), index, dim=0) for an index of any shape; rows >= len(x) read as zeros."""
    x = _req(x.view(x.shape[0], -1))
    index = _req(index, torch.int64)
    rows = index.numel()
    out = torch.empty((rows, x.shape[1]), dtype=_F32, device=x.device)
    st = _lib.lib().gr_gather_rows(x.data_ptr(), x.shape[0], x.shape[1], index.data_ptr(), rows, out.data_ptr(), _stream())
    _lib.check(st, "gather_rows")
    return out.view(*index.shape, x.shape[1])


def point_to_node_partition(points, nodes, point_limit):
    """modules/ops/pointcloud_partition.py:61-111 -> (point_to_node i64, node_masks bool, knn_indices i64, knn_masks bool)."""
    points, nodes = _req(points), _req(nodes)
    N, M = points.shape[0], nodes.shape[0]
    dev = points.device
    p2n = torch.empty((N,), dtype=torch.int32, device=dev)
    node_masks = torch.empty((M,), dtype=torch.uint8, device=dev)
    knn_idx = torch.empty((M, point_limit), dtype=torch.int64, device=dev)
    knn_masks = torch.empty((M, point_limit), dtype=torch.uint8, device=dev)
    L = _lib.lib()
    ws = _workspace(L.gr_point_to_node_workspace_size(N, M), dev)
    st = L.gr_point_to_node_partition(points.data_ptr(), N, nodes.data_ptr(), M, point_limit, p2n.data_ptr(),
                                      node_masks.data_ptr(), knn_idx.data_ptr(), knn_masks.data_ptr(), ws.data_ptr(),
                                      ws.numel(), _stream())
    _lib.check(st, "point_to_node_partition")
    return p2n.long(), node_masks.view(torch.bool), knn_idx, knn_masks.view(torch.bool)


def embedding_indices(points, sigma_d, sigma_a, angle_k):
    points = _req(points)
    N = points.shape[0]
    dev = points.device
    d_idx = torch.empty((N, N), dtype=_F32, device=dev)
    a_idx = torch.empty((N, N, angle_k), dtype=_F32, device=dev)
    knn = torch.empty((N, angle_k), dtype=torch.int32, device=dev)
    st = _lib.lib().gr_embedding_indices(points.data_ptr(), N, float(sigma_d), float(sigma_a), angle_k, d_idx.data_ptr(),
                                         a_idx.data_ptr(), knn.data_ptr(), _stream())
    _lib.check(st, "embedding_indices")
    return d_idx, a_idx, knn


def sinusoid_rows(x, div_term, out=None):
    x = _req(x)
    rows = x.numel()
    if out is None:
        out = torch.empty((rows, 2 * div_term.numel()), dtype=_F32, device=x.device)
    st = _lib.lib().gr_sinusoid_rows(x.data_ptr(), rows, div_term.data_ptr(), div_term.numel(), out.data_ptr(), _stream())
    _lib.check(st, "sinusoid_rows")
    return out


def embedding_combine(D, A, k, out):
    st = _lib.lib().gr_embedding_combine(D.data_ptr(), A.data_ptr(), D.shape[0], D.shape[1], k, out.data_ptr(), _stream())
    _lib.check(st, "embedding_combine")
    return out


def packed_weight_tf32x3(weight):
    """(N,K) Linear weight -> tensor-core operand format (hi/lo TF32 tiles, pre-swizzled); cached on the parameter."""
    cached = getattr(weight, "_gr_packed", None)
    if cached is None or cached[0] != weight._version or cached[1].device != weight.device:
        N, K = weight.shape
        out = torch.empty((2 * ((N + 255) // 256 * 256) * ((K + 31) // 32 * 32),), dtype=_F32, device=weight.device)
        st = _lib.lib().gr_pack_weight_tf32x3(weight.detach().contiguous().data_ptr(), N, K, out.data_ptr(), _stream())
        _lib.check(st, "pack_weight_tf32x3")
        cached = (weight._version, out)
        try:
            weight._gr_packed = cached
        except AttributeError:
            pass
    return cached[1]


def packed_weight_f16x2(weight):
    """(256,256) projection weight -> fp16 hi/lo tiles for the fp16-split structure-embedding kernel, pre-scaled by a
    power of two so that the largest |w| lands in [4, 8) (lo parts stay fp16 normals; fp16 tops out at 65504).
    Returns (buffer, 1/scale); cached on the parameter."""
    cached = getattr(weight, "_gr_packed16", None)
    if cached is None or cached[0] != weight._version or cached[1].device != weight.device:
        N, K = weight.shape
        wmax = float(weight.detach().abs().max())
        exp = 0 if not (wmax > 0.0 and math.isfinite(wmax)) else max(-14, min(14, math.floor(math.log2(8.0 / wmax))))
        scale = 2.0 ** exp
        out = torch.empty((4 * 2 * 256 * 64 * 2,), dtype=torch.uint8, device=weight.device)
        st = _lib.lib().gr_pack_weight_f16x2(weight.detach().contiguous().data_ptr(), N, K, scale, out.data_ptr(), _stream())
        _lib.check(st, "pack_weight_f16x2")
        cached = (weight._version, out, 1.0 / scale)
        try:
            weight._gr_packed16 = cached
        except AttributeError:
            pass
    return cached[1], cached[2]


def _t1_f16():
    v = os.environ.get("GAUSSREG_T1_F16")
    return True if v is None else v not in ("0", "")


def _t1_mode():
    """GAUSSREG_T1: 'table' (default: Hermite tables, no projection in the hot path) | 'tc' (tcgen05 projections)."""
    return os.environ.get("GAUSSREG_T1", "table")


def embedding_table(div_term, proj_d_w, proj_d_b, proj_a_w, proj_a_b, sigma_a):
    """Exact node tables of proj_d(sinusoid(.)) / proj_a(sinusoid(.)), cached on proj_d.weight by parameter versions."""
    key = (proj_d_w._version, proj_d_b._version, proj_a_w._version, proj_a_b._version, proj_a_w.data_ptr(), float(sigma_a),
           proj_d_w.device)
    cached = getattr(proj_d_w, "_gr_t1_table", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    C = proj_d_w.shape[0]
    n = _lib.lib().gr_structure_embedding_table_floats(C, float(sigma_a))
    if n <= 0:
        return None
    tab = torch.empty((n,), dtype=_F32, device=proj_d_w.device)
    st = _lib.lib().gr_structure_embedding_build_table(div_term.data_ptr(), C, proj_d_w.detach().contiguous().data_ptr(),
                                                       proj_d_b.data_ptr(), proj_a_w.detach().contiguous().data_ptr(),
                                                       proj_a_b.data_ptr(), float(sigma_a), tab.data_ptr(), _stream())
    _lib.check(st, "structure_embedding_build_table")
    try:
        proj_d_w._gr_t1_table = (key, tab)
    except AttributeError:
        pass
    return tab


def structure_embedding_tabulated(d_idx, a_idx, div_term, proj_d_w, proj_d_b, proj_a_w, proj_a_b, sigma_a):
    """geotransformer.py:57-72 as f_d(d) + max_k f_a(a_k) from Hermite tables; None when the table cannot be used."""
    N = d_idx.shape[0]
    k = a_idx.shape[-1]
    C = proj_d_w.shape[0]
    tab = embedding_table(div_term, proj_d_w, proj_d_b, proj_a_w, proj_a_b, sigma_a)
    if tab is None:
        return None
    out = torch.empty((N, N, C), dtype=_F32, device=d_idx.device)
    st = _lib.lib().gr_structure_embedding_tabulated(d_idx.data_ptr(), a_idx.data_ptr(), N * N, k, tab.data_ptr(), float(sigma_a),
                                                     div_term.data_ptr(), C, proj_d_w.detach().contiguous().data_ptr(),
                                                     proj_d_b.data_ptr(), proj_a_w.detach().contiguous().data_ptr(),
                                                     proj_a_b.data_ptr(), out.data_ptr(), _stream())
    if st == -3:  # GR_ERR_CAPACITY: this sigma_a's table does not fit one SM's shared memory
        return None
    _lib.check(st, "structure_embedding_tabulated")
    return out


def structure_embedding_points(points, div_term, proj_d_w, proj_d_b, proj_a_w, proj_a_b, sigma_d, sigma_a, angle_k):
    """geotransformer.py:26-72 from the superpoint coordinates in one C-ABI call (indices + tabulated embedding); None when
    the table cannot be used."""
    points = _req(points)
    N = points.shape[0]
    C = proj_d_w.shape[0]
    if not (proj_d_w.is_contiguous() and proj_a_w.is_contiguous()):
        return None
    tab = embedding_table(div_term, proj_d_w, proj_d_b, proj_a_w, proj_a_b, sigma_a)
    if tab is None:
        return None
    dev = points.device
    scratch = torch.empty((N * N * (1 + angle_k) + N * angle_k,), dtype=_F32, device=dev)  # d_idx | a_idx | knn (int32)
    out = torch.empty((N, N, C), dtype=_F32, device=dev)
    p0 = scratch.data_ptr()
    st = _lib.lib().gr_structure_embedding_points(points.data_ptr(), N, float(sigma_d), float(sigma_a), angle_k, tab.data_ptr(),
                                                  div_term.data_ptr(), C, proj_d_w.data_ptr(), proj_d_b.data_ptr(),
                                                  proj_a_w.data_ptr(), proj_a_b.data_ptr(), p0, p0 + 4 * N * N,
                                                  p0 + 4 * N * N * (1 + angle_k), out.data_ptr(), _stream())
    if st == -3:  # GR_ERR_CAPACITY: this sigma_a's table does not fit one SM's shared memory
        return None
    _lib.check(st, "structure_embedding_points")
    return out


def structure_embedding_fused(d_idx, a_idx, div_term, proj_d_w, proj_d_b, proj_a_w, proj_a_b):
    """geotransformer.py:57-72 in one tensor-core kernel: (N,N), (N,N,k) indices -> (N,N,C)."""
    N = d_idx.shape[0]
    k = a_idx.shape[-1]
    C = proj_d_w.shape[0]
    out = torch.empty((N, N, C), dtype=_F32, device=d_idx.device)
    if _t1_f16() and C == 256 and proj_d_w.shape[1] == 256:
        wd, isd = packed_weight_f16x2(proj_d_w)
        wa, isa = packed_weight_f16x2(proj_a_w)
        st = _lib.lib().gr_structure_embedding_fused_f16(d_idx.data_ptr(), a_idx.data_ptr(), N * N, k, div_term.data_ptr(), C,
                                                         wd.data_ptr(), wa.data_ptr(), isd, isa, proj_d_b.data_ptr(),
                                                         proj_a_b.data_ptr(), out.data_ptr(), _stream())
        _lib.check(st, "structure_embedding_fused_f16")
        return out
    st = _lib.lib().gr_structure_embedding_fused(d_idx.data_ptr(), a_idx.data_ptr(), N * N, k, div_term.data_ptr(), C,
                                                 packed_weight_tf32x3(proj_d_w).data_ptr(), packed_weight_tf32x3(proj_a_w).data_ptr(),
                                                 proj_d_b.data_ptr(), proj_a_b.data_ptr(), out.data_ptr(), _stream())
    _lib.check(st, "structure_embedding_fused")
    return out


def rpe_attention_probs(q, k, U, qb, emb, num_heads):
    N, C = q.shape
    P = torch.empty((num_heads, N, N), dtype=_F32, device=q.device)
    st = _lib.lib().gr_rpe_attention_probs(q.data_ptr(), k.data_ptr(), U.data_ptr(), qb.data_ptr(), emb.data_ptr(), N, C,
                                           num_heads, P.data_ptr(), _stream())
    _lib.check(st, "rpe_attention_probs")
    return P


def softmax_rows_(x):
    x2 = x.view(-1, x.shape[-1])
    st = _lib.lib().gr_softmax_rows(x2.data_ptr(), x2.shape[0], x2.shape[1], _stream())
    _lib.check(st, "softmax_rows")
    return x


def l2_normalize_rows(x, eps=1e-12):
    x = _req(x)
    out = torch.empty_like(x)
    st = _lib.lib().gr_l2_normalize_rows(x.data_ptr(), x.shape[0], x.shape[1], float(eps), out.data_ptr(), _stream())
    _lib.check(st, "l2_normalize_rows")
    return out


def superpoint_matching(ref_feats, src_feats, ref_masks, src_masks, k, dual_normalization=True):
    """superpoint_matching.py:13-50 -> (ref_idx (k) i64, src_idx (k) i64, scores (k), count device i32)."""
    ref_feats, src_feats = _req(ref_feats), _req(src_feats)
    Nr, Ns = ref_feats.shape[0], src_feats.shape[0]
    dev = ref_feats.device
    xy = gemm(ref_feats, src_feats, True)
    ref_idx = torch.empty((k,), dtype=torch.int64, device=dev)
    src_idx = torch.empty((k,), dtype=torch.int64, device=dev)
    scores = torch.empty((k,), dtype=_F32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws = _workspace(L.gr_superpoint_matching_workspace_size(Nr, Ns, k), dev)
    st = L.gr_superpoint_matching(xy.data_ptr(), Nr, Ns, _u8(ref_masks).data_ptr(), _u8(src_masks).data_ptr(), k,
                                  int(dual_normalization), ref_idx.data_ptr(), src_idx.data_ptr(), scores.data_ptr(),
                                  count.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "superpoint_matching")
    return ref_idx, src_idx, scores, count


def sinkhorn(scores, row_masks, col_masks, alpha, num_iterations, inf=1e12):
    """learnable_sinkhorn.py:20-66: (P,K,K) -> (P,K+1,K+1)."""
    scores = _req(scores)
    P, K, K2 = scores.shape
    assert K == K2
    out = torch.empty((P, K + 1, K + 1), dtype=_F32, device=scores.device)
    rm, cm = _u8(row_masks), _u8(col_masks)
    alpha = alpha.detach().reshape(1)
    st = _lib.lib().gr_sinkhorn(scores.data_ptr(), rm.data_ptr(), cm.data_ptr(), alpha.data_ptr(), P, K, num_iterations,
                                float(inf), out.data_ptr(), _stream())
    _lib.check(st, "sinkhorn")
    return out


def local_global_registration(matching_scores, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, topk,
                              acceptance_radius, mutual, confidence_threshold, correspondence_threshold, num_refinement_steps):
    """local_global_registration.py:196-235 -> (ref_corr (cap,3), src_corr (cap,3), scores (cap), num_corr i32[1], T (4,4))."""
    ms = _req(matching_scores)
    P, ld, _ = ms.shape
    K = ref_knn_points.shape[1]
    dev = ms.device
    cap = P * K * topk
    ref_corr = torch.empty((cap, 3), dtype=_F32, device=dev)
    src_corr = torch.empty((cap, 3), dtype=_F32, device=dev)
    scores = torch.empty((cap,), dtype=_F32, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    T = torch.empty((4, 4), dtype=_F32, device=dev)
    L = _lib.lib()
    ws = _workspace(L.gr_lgr_workspace_size(P, K, topk), dev)
    st = L.gr_local_global_registration(ms.data_ptr(), P, K, ld, _req(ref_knn_points).data_ptr(), _req(src_knn_points).data_ptr(),
                                        _u8(ref_knn_masks).data_ptr(), _u8(src_knn_masks).data_ptr(), topk,
                                        float(acceptance_radius), int(mutual), float(confidence_threshold),
                                        int(correspondence_threshold), int(num_refinement_steps), ref_corr.data_ptr(),
                                        src_corr.data_ptr(), scores.data_ptr(), num.data_ptr(), T.data_ptr(), ws.data_ptr(),
                                        ws.numel(), _stream())
    _lib.check(st, "local_global_registration")
    return ref_corr, src_corr, scores, num, T


def similarity_ransac(ref_corr, src_corr, num_corr=None, num_hypotheses=10000, sample_size=5, distance_threshold=0.05, seed=0,
                      refit=False, fallback=None):
    """registration_with_ransac_from_correspondences (utils/open3d.py:169-198) on the device ->
    (T (4,4) similarity transform, info int32[2] = {inliers, hypothesis index}).  `num_corr`: device int32 count of valid
    rows (as LocalGlobalRegistration.forward_device leaves it) or None when every row is valid."""
    ref_corr, src_corr = _req(ref_corr), _req(src_corr)
    dev = ref_corr.device
    T = torch.empty((4, 4), dtype=_F32, device=dev)
    info = torch.empty((2,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws = _workspace(L.gr_similarity_ransac_workspace_size(num_hypotheses), dev)
    st = L.gr_similarity_ransac(ref_corr.data_ptr(), src_corr.data_ptr(), _ptr(num_corr), ref_corr.shape[0], int(num_hypotheses),
                                int(sample_size), float(distance_threshold), int(seed) & ((1 << 64) - 1), int(bool(refit)),
                                _ptr(_req(fallback)) if fallback is not None else None, T.data_ptr(), info.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "similarity_ransac")
    return T, info


def weighted_procrustes(src_points, ref_points, weights, eps=1e-5):
    """procrustes.py:6-82 with return_transform=True."""
    squeeze = src_points.dim() == 2
    if squeeze:
        src_points, ref_points, weights = src_points[None], ref_points[None], weights[None]
    src_points, ref_points, weights = _req(src_points), _req(ref_points), _req(weights)
    B, n = weights.shape
    T = torch.empty((B, 4, 4), dtype=_F32, device=src_points.device)
    st = _lib.lib().gr_weighted_procrustes(src_points.data_ptr(), ref_points.data_ptr(), weights.data_ptr(), B, n, float(eps),
                                           T.data_ptr(), _stream())
    _lib.check(st, "weighted_procrustes")
    return T[0] if squeeze else T


def apply_transform(points, transform):
    """modules/ops/transformation.py:7-60 for (*,3) points and a (4,4) transform (3x3 product through gr_gemm)."""
    shape = points.shape
    p = _req(points.reshape(-1, 3))
    R = transform[:3, :3].contiguous()
    t = transform[:3, 3].contiguous()
    return gemm(p, R, True, bias=t).view(shape)


def _fused_qkv(att):
    """proj_q | proj_k | proj_v stacked row-wise -> ((3C,C) weight, (3C) bias); cached on proj_q.weight and rebuilt
    when any of the six parameters changes version or device (load_state_dict copies in place -> version bump)."""
    params = (att.proj_q.weight, att.proj_k.weight, att.proj_v.weight, att.proj_q.bias, att.proj_k.bias, att.proj_v.bias)
    key = tuple((t._version, t.data_ptr()) for t in params)
    cached = getattr(att.proj_q.weight, "_gr_qkv", None)
    if cached is None or cached[0] != key:
        w = torch.cat([t.detach() for t in params[:3]], dim=0).contiguous()
        b = torch.cat([t.detach() for t in params[3:]], dim=0).contiguous()
        cached = (key, w, b)
        try:
            att.proj_q.weight._gr_qkv = cached
        except AttributeError:
            pass
    return cached[1], cached[2]


def _transposed(weight):
    """(out, in) Linear weight -> contiguous (in, out) copy, cached on the parameter per version / device."""
    cached = getattr(weight, "_gr_t", None)
    key = (weight._version, weight.data_ptr())
    if cached is None or cached[0] != key:
        cached = (key, weight.detach().t().contiguous())
        try:
            weight._gr_t = cached
        except AttributeError:
            pass
    return cached[1]


def stacked_rows(a, b):
    """The (Na+Nb, ...) tensor whose row blocks `a` and `b` are, if they are adjacent contiguous views of one storage
    (e.g. x[:n] and x[n:]); else None.  Row-wise ops then run once on both clouds."""
    if (a.dim() >= 1 and a.dim() == b.dim() and a.shape[1:] == b.shape[1:] and a.dtype == b.dtype and a.is_contiguous()
            and b.is_contiguous() and a.numel() > 0 and b.numel() > 0
            and a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
            and a.data_ptr() + a.numel() * a.element_size() == b.data_ptr()):
        return torch.as_strided(a, (a.shape[0] + b.shape[0],) + tuple(a.shape[1:]), a.stride())
    return None


def conditional_transformer(layer_modules, blocks, feats0, feats1, emb0, emb1, num_heads):
    """RPEConditionalTransformer.forward in one C-ABI call; feats are updated in place and returned."""
    import ctypes
    n = len(layer_modules)
    arr = (_lib.LayerWeights * n)()
    keep = []
    for i, (m, block) in enumerate(zip(layer_modules, blocks)):
        a, o = m.attention, m.output
        att = a.attention
        w = arr[i]
        w.wq, w.bq = att.proj_q.weight.data_ptr(), att.proj_q.bias.data_ptr()
        w.wk, w.bk = att.proj_k.weight.data_ptr(), att.proj_k.bias.data_ptr()
        w.wv, w.bv = att.proj_v.weight.data_ptr(), att.proj_v.bias.data_ptr()
        if block == "self":
            w.wp, w.bp = att.proj_p.weight.data_ptr(), att.proj_p.bias.data_ptr()
        else:
            w.wp, w.bp = None, None
        w.wo, w.bo = a.linear.weight.data_ptr(), a.linear.bias.data_ptr()
        w.ln1_g, w.ln1_b = a.norm.weight.data_ptr(), a.norm.bias.data_ptr()
        w.w1, w.b1 = o.expand.weight.data_ptr(), o.expand.bias.data_ptr()
        w.w2, w.b2 = o.squeeze.weight.data_ptr(), o.squeeze.bias.data_ptr()
        w.ln2_g, w.ln2_b = o.norm.weight.data_ptr(), o.norm.bias.data_ptr()
        w.is_self = 1 if block == "self" else 0
        wqkv, bqkv = _fused_qkv(att)
        keep.append((wqkv, bqkv))
        w.wqkv, w.bqkv = wqkv.data_ptr(), bqkv.data_ptr()
        w1t, w2t = _transposed(o.expand.weight), _transposed(o.squeeze.weight)
        keep.append((w1t, w2t))
        w.w1t, w.w2t = w1t.data_ptr(), w2t.data_ptr()
    both = stacked_rows(feats0, feats1) if feats0.dtype == _F32 and feats0.is_cuda else None
    if both is not None:  # one copy; the library then works in place on the stacked rows (no staging copies either)
        both = both.clone()
        f0, f1 = both[:feats0.shape[0]], both[feats0.shape[0]:]
    else:
        f0, f1 = _req(feats0).clone(), _req(feats1).clone()
    N0, C = f0.shape
    N1 = f1.shape[0]
    L = _lib.lib()
    ws = _workspace(L.gr_conditional_transformer_workspace_size(N0, N1, C, num_heads), f0.device)
    st = L.gr_conditional_transformer(ctypes.cast(arr, ctypes.c_void_p), n, f0.data_ptr(), f1.data_ptr(), _req(emb0).data_ptr(),
                                      _req(emb1).data_ptr(), N0, N1, C, num_heads, ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "conditional_transformer")
    return f0, f1


# ------------------------------------------------------------------------------------------------
# device guard: every op runs on the device (and that device's current stream) of its first tensor argument
# ------------------------------------------------------------------------------------------------
def _first_tensor_device(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            return a.device if a.is_cuda else None
        if isinstance(a, torch.nn.Module):
            for p in a.parameters():
                return p.device if p.is_cuda else None
    return None


def _device_guarded(fn):
    """Tensors on a GPU that is not the current one (threads per GPU, an explicit `cuda:1` model in a `cuda:0` process)
    would otherwise be launched on the wrong device: `_stream()` and the workspaces follow torch's CURRENT device.  The
    guard switches to the tensor's device for the duration of the call; on the common single-device path it costs one
    integer comparison."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _first_tensor_device(args, kwargs)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


for _name, _obj in list(globals().items()):
    if callable(_obj) and getattr(_obj, "__module__", None) == __name__ and not _name.startswith("_") and \
            _name not in ("invalidate_weight_caches", "stacked_rows"):
        globals()[_name] = _device_guarded(_obj)
del _name, _obj
