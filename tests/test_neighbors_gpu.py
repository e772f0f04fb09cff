"""GPU parity tests for G1/G2: the CUDA path, called through the C ABI behind
the reference's `geotransformer.ext` interface, against the reference-generated golden fixtures,
the plain-C oracle, and (when the prebuilt oracle/_ref travelled) the reference itself.
Bar: bit-exact values, order and indices (equal-distance ties canonicalised by index)."""
import os

import numpy as np
import pytest
import torch

from gaussreg_b200 import ext
from gaussreg_b200.synthetic import make_pair_inputs
from oracle import neighbors as on

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "neighbors_golden.npz")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _gpu_grid(pts, lens, voxel):
    sp, sl = ext.grid_subsampling(_t(pts), _t(lens), voxel)
    return sp.numpy(), sl.numpy()


def _gpu_radius(q, s, ql, sl, r):
    return ext.radius_neighbors(_t(q), _t(s), _t(ql), _t(sl), r).numpy()


@pytest.mark.parametrize("name", ["room_1500", "room_1500_coarse", "box_1200"])
def test_golden(name):
    gold = np.load(GOLD)
    seed, n, voxel, radius = gold[f"{name}/meta"]
    d = make_pair_inputs(int(seed), int(n), geometry=str(gold[f"{name}/geom"]))
    pts = np.concatenate([d["ref_points"], d["src_points"]]).astype(np.float32)
    lens = np.array([int(n), int(n)], np.int64)
    sp, sl = _gpu_grid(pts, lens, float(voxel))
    assert np.array_equal(sl, gold[f"{name}/sub_lengths"])
    assert np.array_equal(sp.view(np.uint32), gold[f"{name}/sub_points"].view(np.uint32))
    assert np.array_equal(_gpu_radius(pts, pts, lens, lens, float(radius)), gold[f"{name}/self"])
    assert np.array_equal(_gpu_radius(sp, pts, sl, lens, float(radius)), gold[f"{name}/down"])
    assert np.array_equal(_gpu_radius(pts, sp, lens, sl, float(radius) * 2), gold[f"{name}/up"])


@pytest.mark.parametrize("n,geom,seed", [(5000, "room", 0), (30000, "room", 0), (30000, "box", 1), (12000, "room", 7)])
def test_pyramid_vs_oracle(n, geom, seed):
    """The whole 5-stage pyramid at BASELINE sizes: 4 grid subsamples + 13 radius searches."""
    d = make_pair_inputs(seed, n, geometry=geom)
    pts = np.concatenate([d["ref_points"], d["src_points"]]).astype(np.float32)
    lens = np.array([n, n], np.int64)
    limits = [89, 30, 43, 49, 49]
    impl = on.ref() if on.have_ref() else on.port()
    want = on.precompute_data_stack_mode(impl, pts, lens, 5, 0.025, 0.0625, limits)

    class Gpu:
        grid_subsampling = staticmethod(_gpu_grid)
        radius_neighbors = staticmethod(_gpu_radius)

    got = on.precompute_data_stack_mode(Gpu, pts, lens, 5, 0.025, 0.0625, limits)
    for i in range(5):
        assert np.array_equal(want["lengths"][i], got["lengths"][i]), i
        assert np.array_equal(want["points"][i].view(np.uint32), got["points"][i].view(np.uint32)), i
    qs = {"neighbors": lambda i: (i, i), "subsampling": lambda i: (i + 1, i), "upsampling": lambda i: (i, i + 1)}
    for key, f in qs.items():
        for i, (x, y) in enumerate(zip(want[key], got[key])):
            qi, si = f(i)
            assert x.shape == y.shape, (key, i, x.shape, y.shape)
            xc, _ = on.canonicalize_ties(x, want["points"][qi], want["points"][si], want["points"][si].shape[0])
            bad = (xc != y).any(1)
            assert not bad.any(), (key, i, int(bad.sum()))


def test_ragged_batch_and_ties():
    rng = np.random.default_rng(3)
    lens = np.array([1, 700, 33, 1500], np.int64)
    pts = rng.normal(scale=0.4, size=(int(lens.sum()), 3)).astype(np.float32)
    pts[701:705] = pts[700]  # exact duplicates: zero-distance ties inside a cloud
    P = on.port()
    for voxel in (0.05, 0.3):
        a, al = P.grid_subsampling(pts, lens, voxel)
        b, bl = _gpu_grid(pts, lens, voxel)
        assert np.array_equal(al, bl)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for r in (0.05, 0.2, 0.9):
        a = P.radius_neighbors(pts, pts, lens, lens, r)
        b = _gpu_radius(pts, pts, lens, lens, r)
        assert a.shape == b.shape
        assert np.array_equal(a, b), r
    # different query / support clouds, queries far outside the support bounding box
    q = (rng.normal(scale=0.4, size=(500, 3)) + np.array([0.0, 0.0, 3.0])).astype(np.float32)
    ql = np.array([100, 100, 100, 200], np.int64)
    a = P.radius_neighbors(q, pts, ql, lens, 2.5)
    b = _gpu_radius(q, pts, ql, lens, 2.5)
    assert a.shape == b.shape and np.array_equal(a, b)


def test_dense_rows_overflow_path():
    """More hits per row than the shared-memory hit buffer (256): the rescan fallback."""
    rng = np.random.default_rng(9)
    pts = rng.random((3000, 3)).astype(np.float32) * 0.2
    lens = np.array([3000], np.int64)
    P = on.port()
    a = P.radius_neighbors(pts, pts, lens, lens, 0.08)
    assert a.shape[1] > 256
    b = _gpu_radius(pts, pts, lens, lens, 0.08)
    assert a.shape == b.shape and np.array_equal(a, b)


def test_large_voxels_single_cell_and_device_inputs():
    rng = np.random.default_rng(4)
    pts = rng.random((5000, 3)).astype(np.float32)
    lens = np.array([2500, 2500], np.int64)
    P = on.port()
    a, al = P.grid_subsampling(pts, lens, 5.0)  # everything in one voxel: long sequential sums
    b, bl = _gpu_grid(pts, lens, 5.0)
    assert al.tolist() == [1, 1] and np.array_equal(al, bl)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # CUDA tensors in -> CUDA tensors out
    sp, sl = ext.grid_subsampling(_t(pts).cuda(), _t(lens).cuda(), 0.1)
    assert sp.is_cuda and sl.is_cuda
    c, cl = P.grid_subsampling(pts, lens, 0.1)
    assert np.array_equal(sp.cpu().numpy().view(np.uint32), c.view(np.uint32))
    nb = ext.radius_neighbors(sp, sp, sl, sl, 0.25)
    assert nb.is_cuda and np.array_equal(nb.cpu().numpy(), P.radius_neighbors(c, c, cl, cl, 0.25))


def test_reference_error_behaviour():
    pts = torch.zeros((4, 3), dtype=torch.float64)
    lens = torch.tensor([4])
    with pytest.raises(RuntimeError):
        ext.grid_subsampling(pts, lens, 0.1)  # CHECK_IS_FLOAT
    with pytest.raises(RuntimeError):
        ext.grid_subsampling(pts.float(), lens.int(), 0.1)  # CHECK_IS_LONG
    with pytest.raises(RuntimeError):
        ext.radius_neighbors(torch.zeros((3, 4)).t(), torch.zeros((4, 3)), lens, lens, 0.1)  # CHECK_CONTIGUOUS
