"""TEST INFRASTRUCTURE: imports the UNMODIFIED Python reference from /root/reference in this
container (no GPU) so that golden vectors can be generated from it.

Only usable where /root/reference exists.  Nothing here is imported by the product or by the
GPU-side tests; the fixtures it produces are committed under tests/golden/.

What is stubbed:
  * absent third-party modules that the reference imports but the forward path does not need:
    open3d (only `o3d.io.read_point_cloud` of the 15x3 kernel disposition is emulated), matplotlib,
    IPython, ipdb, easydict;
  * `geotransformer.ext` -> oracle/_ref (the reference's own C++ sources behind a ctypes shim);
  * `geotransformer.utils.common.ensure_dir` (config.py:26-31 mkdirs at import time);
  * `Tensor.cuda()/Module.cuda()` -> no-ops (there is no GPU in this container);
  * the Open3D RANSAC call at model.py:209-215 -> returns the LGR transform unchanged.
"""
import os
import struct
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
EXP = os.path.join(REF, "experiments", "geotransformer.gaussian_splatting.indoor")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def read_disposition_ply(path):
    """Minimal binary_little_endian PLY reader for the kernel disposition (double x,y,z)."""
    with open(path, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            header += f.readline()
        lines = header.decode().splitlines()
        n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
        props = [l.split()[1] for l in lines if l.startswith("property")]
        fmt = {"double": "d", "float64": "d", "float": "f", "float32": "f"}
        rec = struct.Struct("<" + "".join(fmt[p] for p in props))
        data = [rec.unpack(f.read(rec.size)) for _ in range(n)]
    return np.asarray(data, dtype=np.float64)[:, :3]


def install():
    if "geotransformer" in sys.modules and getattr(sys.modules["geotransformer"], "_gr_harness", False):
        return
    sys.path.insert(0, ROOT)
    from oracle import neighbors as on

    # --- third-party stubs
    class _EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    sys.modules["easydict"] = types.SimpleNamespace(EasyDict=_EasyDict)
    sys.modules["ipdb"] = types.ModuleType("ipdb")
    ipy = types.ModuleType("IPython")
    ipy.embed = lambda *a, **k: None
    sys.modules["IPython"] = ipy
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    sys.modules["matplotlib"] = mpl
    for sub in ("pyplot", "cm", "colors"):
        m = types.ModuleType("matplotlib." + sub)
        sys.modules["matplotlib." + sub] = m
        setattr(mpl, sub, m)

    o3d = types.ModuleType("open3d")

    class _PCD:
        def __init__(self, pts):
            self.points = pts

    o3d.io = types.SimpleNamespace(read_point_cloud=lambda p: _PCD(read_disposition_ply(p)))
    o3d.geometry = types.SimpleNamespace()
    o3d.utility = types.SimpleNamespace()
    sys.modules["open3d"] = o3d

    # --- geotransformer.ext -> the reference's own C++ (oracle/_ref)
    R = on.ref()
    extm = types.ModuleType("geotransformer.ext")

    def grid_subsampling(points, lengths, voxel_size):
        sp, sl = R.grid_subsampling(points.numpy(), lengths.numpy(), voxel_size)
        return [torch.from_numpy(sp), torch.from_numpy(sl)]

    def radius_neighbors(q, s, ql, sl, radius):
        return torch.from_numpy(R.radius_neighbors(q.numpy(), s.numpy(), ql.numpy(), sl.numpy(), radius))

    extm.grid_subsampling = grid_subsampling
    extm.radius_neighbors = radius_neighbors

    sys.path.insert(0, REF)
    sys.path.insert(0, EXP)
    import geotransformer  # noqa: E402

    geotransformer._gr_harness = True
    sys.modules["geotransformer.ext"] = extm
    geotransformer.ext = extm
    import geotransformer.utils.common as common  # noqa: E402

    common.ensure_dir = lambda p: None

    # --- no GPU here
    torch.Tensor.cuda = lambda self, *a, **k: self.contiguous()
    torch.nn.Module.cuda = lambda self, *a, **k: self


def create_reference_model(seed=0):
    """`create_model(make_cfg())` of the reference under fixed torch/numpy seeds, eval mode, with the
    RANSAC post-step replaced by a pass-through."""
    install()
    import config as ref_config  # noqa: E402
    import model as ref_model  # noqa: E402

    ref_model.registration_with_ransac_from_correspondences = None
    cfg = ref_config.make_cfg()
    torch.manual_seed(seed)
    np.random.seed(seed)
    net = ref_model.create_model(cfg)
    net.eval()
    return net, cfg
