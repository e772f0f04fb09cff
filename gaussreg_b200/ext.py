"""Drop-in for the reference's native module ``geotransformer.ext`` (geotransformer/extensions/pybind.cpp:6-18).

Same two functions, same argument meaning, same dtype/contiguity checks and RuntimeError behaviour
(geotransformer/extensions/common/torch_helper.h:6-35) -- but the work runs on the current CUDA
device through the C ABI in include/gaussreg_b200.h.  Tensors may live on the CPU (the reference's
convention: they are moved to the current CUDA device and the results are returned on the CPU) or
already on the GPU (results stay there).  There is no CPU implementation behind this module.

`install_as_geotransformer_ext()` registers this module under the name the reference's
`geotransformer/modules/ops/{grid_subsample,radius_search}.py` import.
"""
import sys

import torch

from . import _lib


def _check(t, name, dtype, what):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {what} tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("gaussreg_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    # raw handle of the current stream (torch.cuda.current_stream() costs ~10 us per call)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


_ws_cache = {}


def _workspace(nbytes, dev):
    """Grow-only scratch buffer per device (stream-ordered reuse on the current stream)."""
    key = (dev.index, _stream())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    return buf


def release_workspaces(device=None):
    """Drop the grow-only scratch buffers (all devices, or one): they are keyed by (device, stream) and otherwise live
    for the life of the process.  Call it after a batch of unusually large clouds, or before tearing streams down; the
    next op simply allocates again."""
    idx = None if device is None else torch.device(device).index
    for key in [k for k in _ws_cache if idx is None or k[0] == idx]:
        del _ws_cache[key]


def grid_subsample_outputs(n, batch, dev):
    """The three output buffers of grid_subsample_device (capacity n)."""
    return (torch.empty((n, 3), dtype=torch.float32, device=dev), torch.empty((batch,), dtype=torch.int64, device=dev),
            torch.empty((1,), dtype=torch.int64, device=dev))


def grid_subsample_device(points, lengths, voxel_size, n_points=None, out=None):
    """Sync-free core: returns (out_points[capacity n], out_lengths, out_total) all on the device.

    ``points`` may hold more rows than sum(lengths); only the first sum(lengths) are used.  ``out``: buffers from
    grid_subsample_outputs (e.g. allocated on the caller's stream when this call runs on another one).
    """
    L = _lib.lib()
    dev = points.device
    n = points.shape[0] if n_points is None else int(n_points)
    batch = lengths.shape[0]
    out, out_len, out_total = out if out is not None else grid_subsample_outputs(n, batch, dev)
    nbytes = L.gr_grid_subsample_workspace_size(n, batch)
    ws = _workspace(nbytes, dev)
    st = L.gr_grid_subsample(points.data_ptr(), lengths.data_ptr(), batch, n, float(voxel_size), out.data_ptr(),
                             out_len.data_ptr(), out_total.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "grid_subsample")
    return out, out_len, out_total


def grid_subsampling(points, lengths, voxel_size):
    """ext.grid_subsampling(points (N,3) f32, lengths (B,) i64, voxel_size) -> [s_points (M,3), s_lengths (B,)]"""
    _check(points, "points", torch.float32, "a float")
    _check(lengths, "lengths", torch.int64, "an long")
    src_dev = points.device
    dev = src_dev if src_dev.type == "cuda" else _device()
    p = points.to(dev, non_blocking=True)
    l = lengths.to(dev, non_blocking=True)
    if p.shape[0] == 0:
        return [points.new_zeros((0, 3)), torch.zeros_like(lengths)]
    with torch.cuda.device(dev):
        out, out_len, out_total = grid_subsample_device(p, l, voxel_size)
        m = int(out_total.item())  # the reference API returns a data-dependent shape
    s_points = out[:m]
    if src_dev.type != "cuda":
        return [s_points.cpu(), out_len.cpu()]
    return [s_points.contiguous() if m != out.shape[0] else s_points, out_len]


def radius_grid_workspace(s_points, s_lengths):
    """A dedicated workspace for one support cloud, so that its cell grid can be reused by later searches."""
    nbytes = _lib.lib().gr_radius_neighbors_workspace_size(0, s_points.shape[0], s_lengths.shape[0])
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=s_points.device)


_side_streams = {}


def side_stream(dev):
    """A helper stream per (device, current stream): work that does not depend on the caller's latest kernels (the radius
    searches of a pyramid, while the grid-subsample chain is still running) is queued there.  Kept for the life of the
    process: the per-stream workspaces are keyed by it."""
    key = (dev.index, _stream())
    s = _side_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _side_streams[key] = s
    return s


_chain_streams = {}


def chain_stream(dev):
    """A HIGH-PRIORITY stream per (device, current stream) for the grid-subsample chain: the chain is the step's serial
    prefix and consists of small kernels, the radius searches on the helper stream are wide (7500 CTAs) and not urgent.
    At equal priority the chain's kernels queue behind the searches' CTAs (measured: the two small-stage calls take 0.17
    and 0.19 ms next to the searches, 0.09 and 0.07 ms alone)."""
    key = (dev.index, _stream())
    s = _chain_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev, priority=-1)
        _chain_streams[key] = s
    return s


_copy_streams = {}


def copy_stream(dev):
    """A host->device copy stream per (device, current stream): inputs that are first needed late in the step (the
    network's input features) are uploaded there, next to the kernels of the neighbour pyramid."""
    key = (dev.index, _stream())
    s = _copy_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _copy_streams[key] = s
    return s


def h2d_rows(tensors, dev, dtype=torch.float32):
    """torch.cat(tensors).to(dev) without the pageable intermediate: every piece is copied from where it lies (pinned
    memory: an asynchronous DMA) into its rows of one device buffer on the current stream."""
    n = sum(int(t.shape[0]) for t in tensors)
    out = torch.empty((n,) + tuple(tensors[0].shape[1:]), dtype=dtype, device=dev)
    o = 0
    for t in tensors:
        out[o:o + t.shape[0]].copy_(t, non_blocking=True)
        o += int(t.shape[0])
    return out


def radius_neighbors_device(q_points, s_points, q_lengths, s_lengths, radius, ld, out=None, grid_ws=None, reuse_grid=False,
                            max_count=None):
    """Sync-free core: fills a (Nq, ld) int64 table (first min(count, ld) sorted neighbours per row,
    padded with Ns) and returns (table, max_count device scalar).  `grid_ws` (radius_grid_workspace) keeps the
    support cloud's cell grid; pass reuse_grid=True on later searches of the same (support, radius)."""
    L = _lib.lib()
    dev = q_points.device
    nq, ns, batch = q_points.shape[0], s_points.shape[0], q_lengths.shape[0]
    if max_count is None:
        max_count = torch.empty((1,), dtype=torch.int32, device=dev)
    if out is None and ld > 0:
        out = torch.empty((nq, ld), dtype=torch.int64, device=dev)
    if grid_ws is None:
        if reuse_grid:
            raise RuntimeError("reuse_grid needs the grid_ws of the call that built the grid")
        grid_ws = _workspace(L.gr_radius_neighbors_workspace_size(nq, ns, batch), dev)
    st = L.gr_radius_neighbors_cached(q_points.data_ptr(), s_points.data_ptr(), q_lengths.data_ptr(), s_lengths.data_ptr(),
                                      batch, nq, ns, float(radius), out.data_ptr() if out is not None else None,
                                      int(ld), max_count.data_ptr(), grid_ws.data_ptr(), grid_ws.numel(), int(reuse_grid),
                                      _stream())
    _lib.check(st, "radius_neighbors")
    return out, max_count


def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius):
    """ext.radius_neighbors(q (N,3), s (M,3), q_lengths, s_lengths, radius) -> (N, max_count) i64, padded with M"""
    _check(q_points, "q_points", torch.float32, "a float")
    _check(s_points, "s_points", torch.float32, "a float")
    _check(q_lengths, "q_lengths", torch.int64, "an long")
    _check(s_lengths, "s_lengths", torch.int64, "an long")
    src_dev = q_points.device
    dev = src_dev if src_dev.type == "cuda" else _device()
    q = q_points.to(dev, non_blocking=True)
    s = s_points.to(dev, non_blocking=True)
    ql = q_lengths.to(dev, non_blocking=True)
    sl = s_lengths.to(dev, non_blocking=True)
    with torch.cuda.device(dev):
        _, max_count = radius_neighbors_device(q, s, ql, sl, radius, 0)  # count pass
        w = int(max_count.item())
        if w == 0 or q.shape[0] == 0:
            out = torch.empty((q.shape[0], w), dtype=torch.int64, device=dev)
        else:
            out, _ = radius_neighbors_device(q, s, ql, sl, radius, w)
    return out.cpu() if src_dev.type != "cuda" else out


def install_as_geotransformer_ext():
    """Make ``importlib.import_module('geotransformer.ext')`` resolve to this module."""
    mod = sys.modules[__name__]
    sys.modules["geotransformer.ext"] = mod
    parent = sys.modules.get("geotransformer")
    if parent is not None:
        setattr(parent, "ext", mod)
    return mod
