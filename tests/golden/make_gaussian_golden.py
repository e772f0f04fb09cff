"""TEST INFRASTRUCTURE: golden vectors for the Gaussian-cloud preparation (N1/N2), produced by the UNMODIFIED
reference functions `_read_ply_by_opacity` / `load_data` of
/root/reference/experiments/geotransformer.gaussian_splatting.indoor/demo.py and the un-normalisation lines
demo.py:173-178, executed in this container.

Stand-ins (third-party modules that are not installed here; the reference code itself is untouched):
  * `plyfile.PlyData.read(path)` -> an in-memory object exposing `elements[0].data[name]`, `elements[0][name]` and
    `elements[0].properties` over a synthetic (N,59) cloud registered under `path`;
  * `fpsample` -> raises if called (the goldens keep `count <= num_sample`, so demo.py:45-48 never runs).

    python tests/golden/make_gaussian_golden.py     # writes tests/golden/gaussian_golden.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

PROPS = (["x", "y", "z"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)] + ["opacity"] +
         [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])

_REGISTRY = {}


class _Prop:
    def __init__(self, name):
        self.name = name


class _Element:
    def __init__(self, cloud):
        self.data = np.zeros(cloud.shape[0], dtype=[(p, "<f4") for p in PROPS])
        for j, p in enumerate(PROPS):
            self.data[p] = cloud[:, j]
        # the reference sorts f_rest_* itself (demo.py:54-55): hand them over shuffled
        self.properties = [_Prop(p) for p in sorted(PROPS, key=lambda s: (len(s), s[::-1]))]

    def __getitem__(self, name):
        return self.data[name]


class _PlyData:
    def __init__(self, cloud):
        self.elements = [_Element(cloud)]

    @staticmethod
    def read(path):
        return _PlyData(_REGISTRY[path])


def install_stubs():
    import ref_harness
    ref_harness.install()
    ply = types.ModuleType("plyfile")
    ply.PlyData = _PlyData
    ply.PlyElement = object
    sys.modules["plyfile"] = ply
    fps = types.ModuleType("fpsample")

    def _no_fps(*a, **k):
        raise RuntimeError("fpsample is third-party and absent: goldens must not reach demo.py:45-48")

    fps.bucket_fps_kdline_sampling = _no_fps
    sys.modules["fpsample"] = fps
    o3d = sys.modules["open3d"]
    for name in ("geometry", "utility", "visualization", "pipelines"):
        if not hasattr(o3d, name):
            setattr(o3d, name, types.SimpleNamespace())


def test_cloud(seed, n, scale=1.0, spread=(4.0, 3.0, 2.5)):
    """Synthetic Gaussian cloud whose opacity filter and percentile crop both bite (not every row survives)."""
    rng = np.random.default_rng(seed)
    xyz = rng.normal(0.0, 1.0, size=(n, 3)) * (np.asarray(spread) * scale / 3.0) + rng.normal(0, 3, size=3)
    f_dc = rng.normal(0.0, 0.6, size=(n, 3))
    f_rest = rng.normal(0.0, 0.25, size=(n, 45))
    opacity = rng.uniform(-1.0, 4.0, size=(n, 1))
    scale_ = rng.normal(-4.0, 0.5, size=(n, 3))
    rot = rng.normal(0.0, 1.0, size=(n, 4))
    return np.concatenate([xyz, f_dc, f_rest, opacity, scale_, rot], axis=1).astype(np.float32)


CASES = {
    # name: (seed_ref, seed_src, n, scale)  -- volumes in / below / above the [10, 50] no-rescale band of demo.py:96-110
    "mid": (11, 12, 6000, 1.0),
    "small": (21, 22, 4000, 0.35),
    "large": (31, 32, 5000, 2.2),
}


def main():
    install_stubs()
    import demo as ref_demo  # the reference's demo.py

    out = {}
    for name, (s0, s1, n, scale) in CASES.items():
        ref_cloud, src_cloud = test_cloud(s0, n, scale), test_cloud(s1, n, scale)
        _REGISTRY["ref.ply"], _REGISTRY["src.ply"] = ref_cloud, src_cloud
        pts, feats = ref_demo._read_ply_by_opacity("ref.ply", 30000)
        args = types.SimpleNamespace(ref_file="ref.ply", src_file="src.ply", num_sample=30000)
        d = ref_demo.load_data(args)
        out[f"{name}/read_points"] = pts
        out[f"{name}/read_feats"] = feats
        for k, v in d.items():
            out[f"{name}/{k}"] = np.asarray(v)
        # demo.py:173-178 applied to a fixed rigid transform (executed from the reference source text itself)
        ang = 0.3 + 0.1 * len(name)
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=np.float32)
        T[:3, 3] = np.array([0.3, -0.2, 0.1], dtype=np.float32)
        src = open(ref_demo.__file__).read().splitlines()
        first = next(i for i, l in enumerate(src) if "estimated_transform_scale = np.zeros_like" in l)
        block = "\n".join(l.strip() for l in src[first:first + 4])
        env = {"np": np, "estimated_transform": T, "ref_adjust_scale": d["ref_adjust_scale"],
               "src_adjust_scale": d["src_adjust_scale"], "ref_center": d["ref_center"].copy(), "src_center": d["src_center"].copy()}
        exec(block, env)
        out[f"{name}/transform_in"] = T
        out[f"{name}/transform_scale"] = env["estimated_transform_scale"]
        out[f"{name}/cloud_checksum"] = np.array([ref_cloud.astype(np.float64).sum(), src_cloud.astype(np.float64).sum()])
        print(name, "kept", pts.shape[0], "of", n, "scales", d["ref_adjust_scale"], d["src_adjust_scale"])
    np.savez_compressed(os.path.join(HERE, "gaussian_golden.npz"), **out)
    print("wrote gaussian_golden.npz,", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
