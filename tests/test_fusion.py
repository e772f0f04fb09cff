"""N4: Gaussian merge, reference gs_fusion.py.  The CPU oracle (oracle/fusion.py) is pinned to goldens
produced by the UNMODIFIED reference `gaussian_fuse` (tests/golden/make_fusion_golden.py); the device path
(gaussreg_b200/fusion.py, csrc/fusion.cu) is compared with both."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import fusion as ofu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = os.path.join(HERE, "golden", "fusion_golden.npz")
SEED = 1234  # numpy seed under which the reference drew its 15 SH probe directions


def _case(gold, name):
    from make_gaussian_golden import test_cloud
    s1, s2, n1, n2 = [int(v) for v in gold[f"{name}/spec"]]
    return test_cloud(s1, n1), test_cloud(s2, n2), gold[f"{name}/transform"], gold[f"{name}/fused"]


@pytest.mark.parametrize("name", ["rigid", "similarity"])
def test_oracle_matches_reference_golden(name):
    c1, c2, T, want = _case(np.load(GOLD), name)
    np.random.seed(SEED)
    got = ofu.gaussian_fuse(c1, c2, T)
    assert got.shape == want.shape
    # positions, SH coefficients, opacities and log-scales bit-identical; quaternions within 2 ulp (torch vs numpy fp32)
    assert np.array_equal(got[:, :55].view(np.uint32), want[:, :55].view(np.uint32))
    assert np.abs(got[:, 55:] - want[:, 55:]).max() < 3e-7


def test_oracle_sh_band_matrices_are_rotations_of_the_basis():
    """Independent of the probe directions: M_l is the exact band rotation, so two different direction sets agree, M(I) = I and
    M(R1 R2) follows the composition of rotations."""
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3))
    R, _ = np.linalg.qr(A)
    R *= np.sign(np.linalg.det(R))
    d1 = rng.normal(size=(15, 3)); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 = rng.normal(size=(15, 3)); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    for a, b in zip(ofu.sh_band_transforms(R, d1), ofu.sh_band_transforms(R, d2)):
        assert np.abs(a - b).max() < 1e-9
    for m in ofu.sh_band_transforms(np.eye(3), d1):
        assert np.abs(m - np.eye(m.shape[0])).max() < 1e-10
    for m in ofu.sh_band_transforms(R, d1):
        assert np.abs(m @ m.T - np.eye(m.shape[0])).max() < 1e-9  # orthogonal: rotations preserve the band's energy


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rigid", "similarity"])
def test_device_fusion_vs_reference_golden(name):
    from gaussreg_b200 import fusion
    c1, c2, T, want = _case(np.load(GOLD), name)
    np.random.seed(SEED)
    got = fusion.gaussian_fuse(c1, c2, T).cpu().numpy()
    assert got.shape == want.shape                       # identical selection of Gaussians from both clouds
    n1 = int(np.isin(want[:, 51], c1[:, 51]).sum())
    assert np.array_equal(got[:, 3:6], want[:, 3:6]) and np.array_equal(got[:, 51], want[:, 51])  # untouched attributes
    err = np.abs(got - want)
    assert err[:, 0:3].max() <= 4e-6 * max(1.0, np.abs(want[:, 0:3]).max())      # positions (fp32 matmul order)
    assert err[:, 6:51].max() <= 1e-6 * max(1.0, np.abs(want[:, 6:51]).max())    # SH bands (double precision inside)
    assert err[:, 52:55].max() <= 1e-6 * max(1.0, np.abs(want[:, 52:55]).max())  # log-scales
    assert err[:, 55:59].max() <= 1e-6                                             # quaternions
    # the transformed second cloud alone against the CPU oracle
    np.random.seed(SEED)
    t_gpu = fusion.transform_gaussians(c2, T).cpu().numpy()
    np.random.seed(SEED)
    t_cpu = ofu.transform_cloud(c2, T)
    assert np.abs(t_gpu - t_cpu).max() <= 4e-6 * max(1.0, np.abs(t_cpu).max())
    q = t_gpu[:, 55:59]
    assert n1 > 0 and np.all(np.abs(np.linalg.norm(q, axis=1) - 1.0) < 1e-5)  # matrix_to_quaternion returns unit quaternions


@pytest.mark.gpu
def test_fusion_files_round_trip(tmp_path):
    from gaussreg_b200 import fusion, gaussians
    c1, c2, T, want = _case(np.load(GOLD), "similarity")
    root = tmp_path / "scene"
    for tag, cloud in (("A", c1), ("B", c2)):
        d = root / tag / "output" / "point_cloud" / "iteration_30000"
        d.mkdir(parents=True)
        gaussians.write_gaussian_ply(str(d / "point_cloud.ply"), cloud)
    (root / "A" / "output" / "cameras.json").write_text("[]")
    (root / "A" / "output" / "cfg_args").write_text("Namespace()")
    tp = tmp_path / "estimated_transform.npz"
    np.savez(tp, estimated_transform=T)
    np.random.seed(SEED)
    fusion.main(["--root_path", str(root), "--transform_path", str(tp)])
    out = root / "fuse" / "output" / "point_cloud" / "iteration_30000" / "point_cloud.ply"
    fused = gaussians.read_gaussian_ply(str(out))
    assert fused.shape == want.shape and np.abs(fused - want).max() < 1e-4
    assert (root / "fuse" / "output" / "cameras.json").exists() and (root / "fuse" / "output" / "cfg_args").exists()
