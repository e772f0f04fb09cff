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


# ---------------------------------------------------------------------------------------------- property tests
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@st.composite
def _clouds(draw):
    """A stacked batch of 1-4 clouds: ragged lengths, three geometries (blob / shell / lattice with exact duplicates and
    points exactly on voxel boundaries), a random offset and scale."""
    seed = draw(st.integers(0, 2 ** 31 - 1))
    nb = draw(st.integers(1, 4))
    lens = [draw(st.integers(1, 900)) for _ in range(nb)]
    kind = draw(st.sampled_from(["blob", "shell", "lattice"]))
    scale = draw(st.sampled_from([0.05, 0.4, 3.0]))
    offset = draw(st.sampled_from([0.0, -7.5, 123.25]))
    rng = np.random.default_rng(seed)
    n = sum(lens)
    if kind == "blob":
        pts = rng.normal(scale=scale, size=(n, 3))
    elif kind == "shell":
        v = rng.normal(size=(n, 3))
        pts = scale * v / np.linalg.norm(v, axis=1, keepdims=True) + rng.normal(scale=0.01 * scale, size=(n, 3))
    else:
        pts = rng.integers(-6, 7, size=(n, 3)) * (scale / 4.0)  # many exact duplicates / boundary points
    return (pts + offset).astype(np.float32), np.array(lens, np.int64), scale


@settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(_clouds(), st.sampled_from([0.25, 0.5, 1.0, 2.5]))
def test_property_grid_subsample_bit_exact(cloud, rel_voxel):
    """G1 for arbitrary ragged batches: values AND emission order bit-identical to the reference's C++ (oracle/_ref),
    idempotence of the per-cloud lengths, and every output point inside the bounding box of its cloud."""
    pts, lens, scale = cloud
    voxel = float(np.float32(rel_voxel * scale / 4.0))
    impl = on.ref() if on.have_ref() else on.port()
    want, wl = impl.grid_subsampling(pts, lens, voxel)
    got, gl = _gpu_grid(pts, lens, voxel)
    assert np.array_equal(wl, gl)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert int(gl.sum()) == got.shape[0] and np.all(gl >= 1) and np.all(gl <= lens)
    o = 0
    s = 0
    for nb, ns in zip(lens, gl):
        c, sub = pts[o:o + nb], got[s:s + ns]
        assert np.all(sub >= c.min(0) - 1e-6 * (1 + np.abs(c.min(0)))) and np.all(sub <= c.max(0) + 1e-6 * (1 + np.abs(c.max(0))))
        o += nb
        s += ns


@settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(_clouds(), st.sampled_from([0.1, 0.3, 0.8, 2.0]), st.booleans())
def test_property_radius_neighbors_bit_exact(cloud, rel_radius, cross):
    """G2 for arbitrary ragged batches: indices bit-identical to the reference after canonicalising equal-distance ties,
    plus the invariants of radius_neighbors_cpu.cpp:3-91 -- ascending distance, strict `d < r^2` in float32, neighbours
    only from the query's own batch element, padding = number of support points."""
    pts, lens, scale = cloud
    radius = float(np.float32(rel_radius * scale / 2.0))
    impl = on.ref() if on.have_ref() else on.port()
    if cross:  # queries = a jittered subset of the support clouds (same batch layout)
        rng = np.random.default_rng(int(lens.sum()))
        keep = [max(1, int(n) // 3) for n in lens]
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
        q = np.concatenate([pts[s:s + k] for s, k in zip(starts, keep)]) + rng.normal(scale=0.05 * scale, size=(sum(keep), 3)).astype(np.float32)
        q, ql = q.astype(np.float32), np.array(keep, np.int64)
    else:
        q, ql = pts, lens
    want = impl.radius_neighbors(q, pts, ql, lens, radius)
    got = _gpu_radius(q, pts, ql, lens, radius)
    assert want.shape == got.shape
    n_s = pts.shape[0]
    wc, _ = on.canonicalize_ties(want, q, pts, n_s)
    assert np.array_equal(wc, got)
    if got.shape[1]:
        valid = got < n_s
        d = ((q[:, None, :] - pts[np.minimum(got, n_s - 1)]) ** 2)
        d2 = (d[..., 0] + d[..., 1]) + d[..., 2]
        assert np.all(d2[valid] < np.float32(radius) * np.float32(radius))
        dd = np.where(valid, d2, np.inf)
        assert np.all(dd[:, 1:] >= dd[:, :-1])                       # ascending, padding last
        qb = np.repeat(np.arange(len(ql)), ql)
        sb = np.repeat(np.arange(len(lens)), lens)
        assert np.all(sb[np.minimum(got, n_s - 1)][valid] == np.broadcast_to(qb[:, None], got.shape)[valid])
