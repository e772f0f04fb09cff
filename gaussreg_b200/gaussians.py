"""Gaussian-splat cloud -> network input on the GPU, and the demo's file formats either side of the path
(SURVEY.md section 8(f) rows N1 / N2).

Mirrors experiments/geotransformer.gaussian_splatting.indoor/demo.py: `_read_ply_by_opacity` (:30-75),
`load_data` (:81-124) and the un-normalisation + `estimated_transform.npz` writer (:173-180).  The O(n) work
(order statistics for the percentile crop, opacity/crop selection, ordered compaction, bounding box, the ordered
float32 column sums behind numpy's `mean(0)`, spherical-harmonics colour in float64, centring/rescaling) runs in
csrc/gaussians.cu through the C ABI; the O(1) scalar glue is evaluated here with numpy exactly as the reference
evaluates it, so that the results are bit-identical, not just close.

A cloud is an (N, 59) float32 array/tensor in 3DGS property order without normals (gs_fusion.py:172-184).
There is no CPU implementation: every function below needs a CUDA device.
"""
import os

import numpy as np
import torch

from . import _lib
from .ext import _device, _stream, _workspace

ATTR_DIM = 59
COL_OPACITY = 51
PLY_PROPERTIES = (["x", "y", "z"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)] + ["opacity"] +
                  [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])


# ------------------------------------------------------------------------------------------------
# N2: 3DGS point_cloud.ply <-> (N,59) array  (what plyfile does for demo.py:32-61)
# ------------------------------------------------------------------------------------------------
_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


def read_gaussian_ply(path):
    """binary_little_endian / ascii PLY with a `vertex` element -> (N,59) float32 in PLY_PROPERTIES order.
    Extra properties (nx, ny, nz, ...) are ignored; `f_rest_*` are ordered by their numeric suffix (demo.py:54-56)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise RuntimeError(f"{path}: not a PLY file")
        fmt, n, props, in_vertex = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise RuntimeError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
                elif n is None:
                    raise RuntimeError(f"{path}: elements before `vertex` are not supported")
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise RuntimeError(f"{path}: list properties in the vertex element are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if n is None:
            raise RuntimeError(f"{path}: no vertex element")
        if fmt == "ascii":
            raw = np.loadtxt(f, max_rows=n, ndmin=2)
            table = {name: raw[:, j] for j, (name, _) in enumerate(props)}
        else:
            order = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(name, order + t) for name, t in props])
            table = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    names = [p for p, _ in props]
    n_rest = len([p for p in names if p.startswith("f_rest_")])
    if n_rest != 45:  # demo.py:56 asserts 3*(3+1)**2 - 3
        raise RuntimeError(f"{path}: expected 45 f_rest_* properties (SH degree 3), found {n_rest}")
    missing = [p for p in PLY_PROPERTIES if p not in names]
    if missing:
        raise RuntimeError(f"{path}: missing properties {missing[:4]}...")
    out = np.empty((n, ATTR_DIM), dtype=np.float32)
    for j, p in enumerate(PLY_PROPERTIES):
        out[:, j] = table[p]
    return out


def write_gaussian_ply(path, cloud):
    """(N,59) float32 -> binary_little_endian PLY in the 3DGS layout (with zero normals, as 3DGS writes them)."""
    cloud = np.ascontiguousarray(np.asarray(cloud, dtype=np.float32))
    names = ["x", "y", "z", "nx", "ny", "nz"] + PLY_PROPERTIES[3:]
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % cloud.shape[0]
    header += "".join(f"property float {p}\n" for p in names) + "end_header\n"
    rows = np.zeros((cloud.shape[0], len(names)), dtype="<f4")
    rows[:, :3] = cloud[:, :3]
    rows[:, 6:] = cloud[:, 3:]
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rows.tobytes())


# ------------------------------------------------------------------------------------------------
# numpy's percentile (method='linear') split into "which order statistics" and "interpolate"
# ------------------------------------------------------------------------------------------------
def percentile_plan(n, q_percent, dtype=np.float32):
    """(previous_index, next_index, gamma) that np.percentile(a, q_percent) uses for a 1-D `dtype` array of n
    finite values: numpy/lib/_function_base_impl.py `percentile` -> `_quantile` with
    `_QuantileMethods['linear']` (virtual index (n - 1) * q), `_get_indexes`, `_get_gamma` (numpy 2.x: the quantile
    takes the array's dtype, so float32 data interpolates in float32).  Checked bit-for-bit against np.percentile
    in tests/test_gaussians_cpu.py."""
    dt = np.dtype(dtype)
    q = np.asanyarray(np.true_divide(q_percent, dt.type(100)))
    virtual = np.asanyarray((n - 1) * q)
    prev = int(np.floor(virtual).astype(np.intp))
    nxt = prev + 1
    if virtual >= n - 1:
        prev = nxt = n - 1
    if virtual < 0:
        prev = nxt = 0
    # numpy: gamma = asanyarray(virtual - previous) [float64], then cast back to the quantile's dtype.  At the
    # clipped ends previous == next, the interpolation returns that element whatever gamma is.
    gamma = np.asanyarray(np.asanyarray(virtual - np.intp(prev)), dtype=virtual.dtype)
    return prev, nxt, gamma


def percentile_lerp(a, b, gamma):
    """numpy's `_lerp` on two order statistics (same dtype as the data)."""
    a, b = np.asanyarray(a), np.asanyarray(b)
    diff = np.subtract(b, a)
    out = np.asanyarray(np.add(a, diff * gamma))
    if gamma >= 0.5:
        out = np.asanyarray(np.subtract(b, diff * (1 - gamma)))
    return out[()]


# ------------------------------------------------------------------------------------------------
# N1 on the device
# ------------------------------------------------------------------------------------------------
def _as_device_cloud(cloud):
    if isinstance(cloud, np.ndarray):
        cloud = torch.from_numpy(np.ascontiguousarray(cloud, dtype=np.float32))
    if cloud.dtype != torch.float32 or cloud.dim() != 2 or cloud.shape[1] < ATTR_DIM:
        raise RuntimeError("a Gaussian cloud must be an (N, >=59) float32 array")
    return cloud.to(_device(), non_blocking=True).contiguous()


def column_order_stats(cloud, cols, ranks):
    """Exact order statistics: value of rank ranks[j] (0-based) in column cols[j].  Returns a host float32 array."""
    import ctypes
    L = _lib.lib()
    q = len(cols)
    c_cols = (ctypes.c_int32 * q)(*[int(c) for c in cols])
    c_ranks = (ctypes.c_int64 * q)(*[int(r) for r in ranks])
    out = torch.empty((q,), dtype=torch.float32, device=cloud.device)
    ws = _workspace(L.gr_column_order_stats_workspace_size(q), cloud.device)
    st = L.gr_column_order_stats(cloud.data_ptr(), cloud.shape[0], cloud.stride(0), ctypes.cast(c_cols, ctypes.c_void_p),
                                 ctypes.cast(c_ranks, ctypes.c_void_p), q, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "column_order_stats")
    return out.cpu().numpy()


def farthest_point_sample(points, k, start_idx=0):
    """Exact farthest-point sampling on the device: (n,3) f32 points -> (k,) int64 indices in selection order, first pick
    `start_idx` (csrc/fps.cu).  Stands in for `fpsample.bucket_fps_kdline_sampling(points, k, h=9)` (demo.py:46), which is
    an accelerated exact FPS with a random start; third party, parity unpinned."""
    L = _lib.lib()
    points = points.to(torch.float32).contiguous()
    if not points.is_cuda:
        points = points.to(_device())
    n = points.shape[0]
    out = torch.empty((k,), dtype=torch.int64, device=points.device)
    ws = _workspace(L.gr_farthest_point_sample_workspace_size(n), points.device)
    st = L.gr_farthest_point_sample(points.data_ptr(), n, int(k), int(start_idx), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "farthest_point_sample")
    return out


def read_cloud_by_opacity(cloud, point_limit=None, opacity_min=0.7, crop_percent=(5, 95), fps_start_idx=0):
    """demo.py:30-75 `_read_ply_by_opacity` for a cloud already in memory.

    Returns device tensors ``(points (M,3) f32, point_features (M,4) f32, index (M,) i64)``.  When more than
    `point_limit` rows survive the filter the reference subsamples them with `fpsample` (demo.py:45-48, random start);
    here that is exact farthest-point sampling on the device started at survivor `fps_start_idx`."""
    import ctypes
    L = _lib.lib()
    cloud = _as_device_cloud(cloud)
    n, ld = cloud.shape[0], cloud.stride(0)
    if n == 0:
        raise RuntimeError("empty Gaussian cloud")
    dev = cloud.device
    # --- np.percentile(x|y|z, 5|95): order statistics on the device, interpolation as numpy does it
    plans = [percentile_plan(n, q) for q in crop_percent]
    cols, ranks = [], []
    for axis in range(3):
        for prev, nxt, _ in plans:
            cols += [axis, axis]
            ranks += [prev, nxt]
    stats = column_order_stats(cloud, cols, ranks)
    lo, hi = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
    for axis in range(3):
        v = stats[4 * axis:4 * axis + 4]
        lo[axis] = float(percentile_lerp(v[0], v[1], plans[0][2]))
        hi[axis] = float(percentile_lerp(v[2], v[3], plans[1][2]))
    # --- (opacity > 0.7) * index_x * index_y * index_z, np.where
    index = torch.empty((n,), dtype=torch.int64, device=dev)
    count = torch.empty((1,), dtype=torch.int64, device=dev)
    ws = _workspace(L.gr_gaussian_select_workspace_size(n), dev)
    st = L.gr_gaussian_select(cloud.data_ptr(), n, ld, COL_OPACITY, float(np.float32(opacity_min)), ctypes.cast(lo, ctypes.c_void_p),
                              ctypes.cast(hi, ctypes.c_void_p), index.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                              _stream())
    _lib.check(st, "gaussian_select")
    m = int(count.item())
    if m == 0:
        raise RuntimeError("no Gaussian survives the opacity / percentile filter")
    index = index[:m]
    if point_limit is not None and m > point_limit:
        # demo.py:44-47: farthest-point subsample of the survivors, kept in selection order
        cand = cloud[:, 0:3].index_select(0, index).contiguous()
        index = index[farthest_point_sample(cand, point_limit, start_idx=fps_start_idx)]
        m = point_limit
    # --- points, their float32 mean (ordered sum), bounding box
    points = torch.empty((m, 3), dtype=torch.float32, device=dev)
    stats_dev = torch.empty((9,), dtype=torch.float32, device=dev)
    ws = _workspace(64, dev)
    st = L.gr_gather_points_stats(cloud.data_ptr(), ld, index.data_ptr(), m, points.data_ptr(), stats_dev.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "gather_points_stats")
    s = stats_dev.cpu().numpy()
    center_point = np.true_divide(s[0:3], np.float32(m)).astype(np.float32)       # points.mean(0)
    max_length = np.linalg.norm(s[6:9] - s[3:6])                                 # float32
    center_point = center_point + np.array([0, 2 * max_length, 0])               # float64, demo.py:64
    view = (ctypes.c_double * 3)(*[float(v) for v in center_point])
    feats = torch.empty((m, 4), dtype=torch.float32, device=dev)
    st = L.gr_gaussian_features(cloud.data_ptr(), ld, index.data_ptr(), m, ctypes.cast(view, ctypes.c_void_p),
                                feats.data_ptr(), _stream())
    _lib.check(st, "gaussian_features")
    return points, feats, index


def center_and_scale(points, bbox=None):
    """demo.py:83-110 for one cloud: returns (points', adjust_scale, center).  `points` (M,3) f32 on the device is
    modified in place."""
    import ctypes
    L = _lib.lib()
    m = points.shape[0]
    if bbox is None:
        stats_dev = torch.empty((9,), dtype=torch.float32, device=points.device)
        ws = _workspace(64, points.device)
        tmp = torch.empty_like(points)
        st = L.gr_gather_points_stats(points.data_ptr(), 3, None, m, tmp.data_ptr(), stats_dev.data_ptr(), ws.data_ptr(),
                                      ws.numel(), _stream())
        _lib.check(st, "gather_points_stats")
        s = stats_dev.cpu().numpy()
        bbox = (s[3:6].copy(), s[6:9].copy())
    mn, mx = bbox
    volume = (mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2])
    center = (mx + mn) / 2
    adjust_scale, apply = 1., False
    if volume > 50:
        adjust_scale, apply = (50 / volume) ** (1 / 3), True
    elif volume < 10:
        adjust_scale, apply = (30 / volume) ** (1 / 3), True
    c = (ctypes.c_float * 3)(*[float(v) for v in center])
    st = L.gr_points_normalize(points.data_ptr(), m, ctypes.cast(c, ctypes.c_void_p), float(adjust_scale), int(apply), _stream())
    _lib.check(st, "points_normalize")
    return points, adjust_scale, center


def load_data(ref_cloud, src_cloud, num_sample=30000):
    """demo.py:81-124 with the clouds given as arrays (or paths to 3DGS .ply files): the dict that
    `registration_collate_fn_stack_mode` takes, points / features resident on the GPU."""
    out = {}
    for key, cloud in (("ref", ref_cloud), ("src", src_cloud)):
        if isinstance(cloud, (str, os.PathLike)):
            cloud = read_gaussian_ply(cloud)
        points, feats, _ = read_cloud_by_opacity(cloud, num_sample)
        points, scale, center = center_and_scale(points)
        out[f"{key}_points"], out[f"{key}_feats"] = points, feats
        out[f"{key}_adjust_scale"], out[f"{key}_center"] = scale, center
    return out


def unnormalize_transform(estimated_transform, ref_adjust_scale, src_adjust_scale, ref_center, src_center):
    """demo.py:173-178 (host numpy, as in the reference): similarity transform between the original clouds."""
    if isinstance(estimated_transform, torch.Tensor):
        estimated_transform = estimated_transform.detach().cpu().numpy()
    T = np.zeros_like(estimated_transform)
    T[:3, :3] = estimated_transform[:3, :3] / ref_adjust_scale * src_adjust_scale
    T[:3, 3] = estimated_transform[:3, 3] / ref_adjust_scale + ref_center - np.matmul(T[:3, :3], src_center)
    T[3, 3] = 1.
    return T


def save_estimated_transform(output_path, transform_scale):
    """demo.py:180: <output_path>/estimated_transform.npz with the key `estimated_transform`."""
    os.makedirs(output_path, exist_ok=True)
    path = os.path.join(output_path, "estimated_transform.npz")
    np.savez(path, estimated_transform=transform_scale)
    return path


def write_point_cloud_ply(path, points, colors=None):
    """The coloured point clouds demo.py:166-179 writes through Open3D (binary PLY: double xyz, uchar rgb)."""
    points = np.asarray(points, dtype=np.float64)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty double x\nproperty double y\nproperty double z\n" % points.shape[0]
    if colors is not None:
        header += "property uchar red\nproperty uchar green\nproperty uchar blue\n"
        rgb = np.clip(np.asarray(colors) * 255.0, 0, 255).astype(np.uint8)
    header += "end_header\n"
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")] + ([("r", "u1"), ("g", "u1"), ("b", "u1")] if colors is not None else [])
    rows = np.zeros(points.shape[0], dtype=fields)
    rows["x"], rows["y"], rows["z"] = points[:, 0], points[:, 1], points[:, 2]
    if colors is not None:
        rows["r"], rows["g"], rows["b"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rows.tobytes())
