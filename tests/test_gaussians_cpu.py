"""CPU tests for the Gaussian-cloud preparation (N1/N2): the numpy oracle against goldens produced by the
unmodified reference (tests/golden/make_gaussian_golden.py), and the host-side glue of the product
(numpy-percentile plan, PLY reader/writer) -- no GPU work."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from gaussreg_b200 import gaussians as G  # noqa: E402
from oracle import gaussians as og  # noqa: E402
from make_gaussian_golden import CASES, test_cloud as make_test_cloud  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "gaussian_golden.npz"))


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_reference_golden_bit_for_bit(case):
    s0, s1, n, scale = CASES[case]
    ref_cloud, src_cloud = make_test_cloud(s0, n, scale), make_test_cloud(s1, n, scale)
    chk = np.array([ref_cloud.astype(np.float64).sum(), src_cloud.astype(np.float64).sum()])
    assert np.array_equal(chk, GOLD[f"{case}/cloud_checksum"]), "synthetic cloud generator drifted"
    pts, feats, _ = og.read_cloud_by_opacity(ref_cloud, 30000)
    assert np.array_equal(_bits(pts), _bits(GOLD[f"{case}/read_points"]))
    assert np.array_equal(_bits(feats), _bits(GOLD[f"{case}/read_feats"]))
    d = og.load_data(ref_cloud, src_cloud, 30000)
    for k in ("ref_points", "src_points", "ref_feats", "src_feats"):
        assert d[k].dtype == np.float32 and np.array_equal(_bits(d[k]), _bits(GOLD[f"{case}/{k}"])), k
    for k in ("ref_adjust_scale", "src_adjust_scale", "ref_center", "src_center"):
        assert np.array_equal(np.asarray(d[k]), GOLD[f"{case}/{k}"]), k
    T = og.unnormalize_transform(GOLD[f"{case}/transform_in"], d["ref_adjust_scale"], d["src_adjust_scale"], d["ref_center"],
                                 d["src_center"])
    assert np.array_equal(T, GOLD[f"{case}/transform_scale"])
    T2 = G.unnormalize_transform(GOLD[f"{case}/transform_in"], d["ref_adjust_scale"], d["src_adjust_scale"], d["ref_center"],
                                 d["src_center"])
    assert np.array_equal(T2, GOLD[f"{case}/transform_scale"])


def test_percentile_plan_reproduces_numpy():
    rng = np.random.default_rng(5)
    for n in [1, 2, 3, 7, 20, 21, 100, 101, 999, 4000, 30000, 65537, 200001]:
        a = rng.normal(size=n).astype(np.float32)
        srt = np.sort(a)
        for q in (0, 5, 50, 95, 99.5, 100):
            prev, nxt, gamma = G.percentile_plan(n, q)
            got = G.percentile_lerp(srt[prev], srt[nxt], gamma)
            want = np.percentile(a, q)
            assert got.dtype == want.dtype == np.float32
            assert got.tobytes() == want.tobytes(), (n, q, got, want)


def test_gaussian_ply_round_trip(tmp_path):
    cloud = make_test_cloud(3, 257)
    path = os.path.join(tmp_path, "point_cloud.ply")
    G.write_gaussian_ply(path, cloud)
    back = G.read_gaussian_ply(path)
    assert back.dtype == np.float32 and np.array_equal(_bits(back), _bits(cloud))
    with open(path, "rb") as f:
        head = f.read(400).decode("ascii", "replace")
    assert "property float nx" in head and "property float f_rest_6\n" in head
    # a file without SH degree 3 is rejected like demo.py:56 does
    bad = os.path.join(tmp_path, "bad.ply")
    with open(bad, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 0\nproperty float x\nend_header\n")
    with pytest.raises(RuntimeError):
        G.read_gaussian_ply(bad)


def test_estimated_transform_npz_writer(tmp_path):
    T = np.arange(16, dtype=np.float32).reshape(4, 4)
    p = G.save_estimated_transform(os.path.join(tmp_path, "out"), T)
    assert os.path.basename(p) == "estimated_transform.npz"
    assert np.array_equal(np.load(p)["estimated_transform"], T)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        G.read_cloud_by_opacity(make_test_cloud(1, 64))
