"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends for the neighbour-pyramid oracles.

* ``port``  : oracle/neighbors.c (plain-C restatement)            -> _build/liboracle_neighbors.so
* ``ref``   : the unmodified reference sources behind ref_shim.cpp -> _ref/libgaussreg_ref.so
              (compiled from /root/reference by oracle/Makefile; travels to the GPU box prebuilt)

Both expose the reference's operator interface
(geotransformer/extensions/pybind.cpp:8-17): ``grid_subsampling(points, lengths, voxel)`` and
``radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius)`` on numpy arrays, plus the
pyramid driver ``precompute_data_stack_mode`` (geotransformer/utils/data.py:13-77).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "_build", "liboracle_neighbors.so")
_REF_SO = os.path.join(_HERE, "_ref", "libgaussreg_ref.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C restatement and (when /root/reference is present) the reference shim."""
    if force or not os.path.exists(_PORT_SO) or os.path.getmtime(_PORT_SO) < os.path.getmtime(
        os.path.join(_HERE, "neighbors.c")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    if os.path.isdir("/root/reference/geotransformer/extensions") and (force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def have_ref():
    return os.path.exists(_REF_SO)


def _as(arr, dtype):
    return np.ascontiguousarray(arr, dtype=dtype)


class _Port:
    kind = "port"

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(_PORT_SO)
        self.lib.oracle_grid_subsample.restype = ctypes.c_int64
        self.lib.oracle_grid_subsample.argtypes = [_f32p, _i64p, ctypes.c_int, ctypes.c_float, _f32p, _i64p]
        for name in ("oracle_radius_neighbors", "oracle_radius_neighbors_brute"):
            fn = getattr(self.lib, name)
            fn.restype = ctypes.c_int64
            fn.argtypes = [_f32p, _f32p, _i64p, _i64p, ctypes.c_int, ctypes.c_float, _i64p, ctypes.c_int64]
        self.lib.oracle_ladder.restype = ctypes.c_int
        self.lib.oracle_ladder.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]

    def ladder(self):
        buf = (ctypes.c_uint64 * 64)()
        n = self.lib.oracle_ladder(buf, 64)
        return [int(buf[i]) for i in range(n)]

    def grid_subsampling(self, points, lengths, voxel):
        points = _as(points, np.float32).reshape(-1, 3)
        lengths = _as(lengths, np.int64)
        out = np.empty_like(points)
        out_len = np.empty_like(lengths)
        m = self.lib.oracle_grid_subsample(
            points.ctypes.data_as(_f32p), lengths.ctypes.data_as(_i64p), len(lengths), float(voxel),
            out.ctypes.data_as(_f32p), out_len.ctypes.data_as(_i64p))
        if m < 0:
            raise RuntimeError("oracle_grid_subsample failed")
        return out[:m].copy(), out_len

    def radius_neighbors(self, q, s, q_len, s_len, radius, brute=False):
        q = _as(q, np.float32).reshape(-1, 3)
        s = _as(s, np.float32).reshape(-1, 3)
        q_len = _as(q_len, np.int64)
        s_len = _as(s_len, np.int64)
        fn = self.lib.oracle_radius_neighbors_brute if brute else self.lib.oracle_radius_neighbors
        args = (q.ctypes.data_as(_f32p), s.ctypes.data_as(_f32p), q_len.ctypes.data_as(_i64p),
                s_len.ctypes.data_as(_i64p), len(q_len), float(radius))
        w = fn(*args, None, 0)
        if w < 0:
            raise RuntimeError("oracle_radius_neighbors failed")
        out = np.empty((q.shape[0], w), dtype=np.int64)
        w2 = fn(*args, out.ctypes.data_as(_i64p), w)
        assert w2 == w
        return out


class _Ref:
    kind = "reference"

    def __init__(self):
        build()
        if not have_ref():
            raise RuntimeError("oracle/_ref/libgaussreg_ref.so is missing (build it where /root/reference exists)")
        self.lib = ctypes.CDLL(_REF_SO)
        self.lib.ref_grid_subsampling.restype = ctypes.c_int64
        self.lib.ref_grid_subsampling.argtypes = [_f32p, _i64p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, _f32p, _i64p]
        self.lib.ref_radius_neighbors.restype = ctypes.c_int64
        self.lib.ref_radius_neighbors.argtypes = [_f32p, _f32p, _i64p, _i64p, ctypes.c_int, ctypes.c_int64,
                                                  ctypes.c_int64, ctypes.c_float, _i64p]

    def grid_subsampling(self, points, lengths, voxel):
        points = _as(points, np.float32).reshape(-1, 3)
        lengths = _as(lengths, np.int64)
        out = np.empty_like(points)
        out_len = np.empty_like(lengths)
        m = self.lib.ref_grid_subsampling(
            points.ctypes.data_as(_f32p), lengths.ctypes.data_as(_i64p), len(lengths), points.shape[0],
            float(voxel), out.ctypes.data_as(_f32p), out_len.ctypes.data_as(_i64p))
        return out[:m].copy(), out_len

    def radius_neighbors(self, q, s, q_len, s_len, radius):
        q = _as(q, np.float32).reshape(-1, 3)
        s = _as(s, np.float32).reshape(-1, 3)
        q_len = _as(q_len, np.int64)
        s_len = _as(s_len, np.int64)
        args = (q.ctypes.data_as(_f32p), s.ctypes.data_as(_f32p), q_len.ctypes.data_as(_i64p),
                s_len.ctypes.data_as(_i64p), len(q_len), q.shape[0], s.shape[0], float(radius))
        w = self.lib.ref_radius_neighbors(*args, None)
        out = np.empty((q.shape[0], w), dtype=np.int64)
        self.lib.ref_radius_neighbors(*args, out.ctypes.data_as(_i64p))
        return out


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref():
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref


def best():
    """The reference build when it exists, else the port (bench.py cpu_baseline)."""
    return ref() if (have_ref() or os.path.isdir("/root/reference")) else port()


def radius_search(impl, q, s, q_len, s_len, radius, limit):
    """geotransformer/modules/ops/radius_search.py:7-27 (truncate to the first `limit` columns)."""
    idx = impl.radius_neighbors(q, s, q_len, s_len, radius)
    if limit > 0:
        idx = idx[:, :limit]
    return np.ascontiguousarray(idx)


def precompute_data_stack_mode(impl, points, lengths, num_stages, voxel_size, radius, neighbor_limits):
    """geotransformer/utils/data.py:13-77, on numpy arrays."""
    assert num_stages == len(neighbor_limits)
    points_list, lengths_list, neighbors_list, subsampling_list, upsampling_list = [], [], [], [], []
    for i in range(num_stages):
        if i > 0:
            points, lengths = impl.grid_subsampling(points, lengths, voxel_size)
        points_list.append(points)
        lengths_list.append(lengths)
        voxel_size *= 2
    for i in range(num_stages):
        cur_p, cur_l = points_list[i], lengths_list[i]
        neighbors_list.append(radius_search(impl, cur_p, cur_p, cur_l, cur_l, radius, neighbor_limits[i]))
        if i < num_stages - 1:
            sub_p, sub_l = points_list[i + 1], lengths_list[i + 1]
            subsampling_list.append(radius_search(impl, sub_p, cur_p, sub_l, cur_l, radius, neighbor_limits[i]))
            upsampling_list.append(radius_search(impl, cur_p, sub_p, cur_l, sub_l, radius * 2, neighbor_limits[i + 1]))
        radius *= 2
    return {"points": points_list, "lengths": lengths_list, "neighbors": neighbors_list,
            "subsampling": subsampling_list, "upsampling": upsampling_list}


def canonicalize_ties(idx, q, s, pad):
    """Sort equal-distance runs of each row by index (the reference's std::sort is unstable on ties).

    Returns (canonical_idx, n_rows_with_ties).  Distances are recomputed with the reference formula.
    """
    idx = idx.copy()
    q = np.asarray(q, np.float32)
    s = np.asarray(s, np.float32)
    sp = np.concatenate([s, np.full((1, 3), np.inf, np.float32)], 0)
    safe = np.where(idx == pad, s.shape[0], idx)
    diff = q[:, None, :] - sp[safe]
    with np.errstate(invalid="ignore", over="ignore"):
        d = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    d = np.where(idx == pad, np.float32(np.inf), d.astype(np.float32))
    tie_rows = np.nonzero((d[:, 1:] == d[:, :-1]).any(1) & np.isfinite(d[:, 1:]).any(1))[0]
    n = 0
    for r in tie_rows:
        row_d, row_i = d[r], idx[r]
        finite = np.isfinite(row_d)
        if not (row_d[1:][finite[1:]] == row_d[:-1][finite[1:]]).any():
            continue
        order = np.lexsort((row_i, row_d))
        idx[r] = row_i[order]
        n += 1
    return idx, n
