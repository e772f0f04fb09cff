"""GPU parity tests for the Gaussian-cloud preparation (N1): csrc/gaussians.cu through the C ABI against the
goldens of the unmodified reference and against the numpy oracle on larger clouds."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from gaussreg_b200 import gaussians as G  # noqa: E402
from oracle import gaussians as og  # noqa: E402
from make_gaussian_golden import CASES, test_cloud as make_test_cloud  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "gaussian_golden.npz"))


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32)


def _ulp_diff(a, b):
    return int(np.abs(_bits(a).astype(np.int64) - _bits(b).astype(np.int64)).max()) if a.size else 0


def test_column_order_stats_exact():
    cloud = make_test_cloud(7, 50001)
    dev = G._as_device_cloud(cloud)
    cols = [0, 0, 1, 2, 2, 51, 58]
    ranks = [0, 50000, 25000, 2500, 47500, 12345, 1]
    got = G.column_order_stats(dev, cols, ranks)
    want = np.array([np.sort(cloud[:, c])[r] for c, r in zip(cols, ranks)], dtype=np.float32)
    assert np.array_equal(_bits(got), _bits(want))
    # duplicates and negative zeros
    dup = cloud.copy()
    dup[:, 0] = np.round(dup[:, 0])
    dup[::7, 0] = -0.0
    got = G.column_order_stats(G._as_device_cloud(dup), [0] * 5, [0, 10, 20000, 40000, 50000])
    want = np.sort(dup[:, 0])[[0, 10, 20000, 40000, 50000]]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case", sorted(CASES))
def test_load_data_matches_reference_golden(case):
    s0, s1, n, scale = CASES[case]
    ref_cloud, src_cloud = make_test_cloud(s0, n, scale), make_test_cloud(s1, n, scale)
    pts, feats, idx = G.read_cloud_by_opacity(ref_cloud, 30000)
    assert np.array_equal(_bits(pts.cpu().numpy()), _bits(GOLD[f"{case}/read_points"]))
    f = feats.cpu().numpy()
    # colours: float64 arithmetic in the reference's operation order -> identical bits; opacity: 1 / (1 + exp(-o)) with
    # numpy's float32 exp (a SIMD polynomial documented at up to 2.52 ulp) vs a correctly rounded exp here: the two
    # sigmoids may differ by a few units in the last place (tolerance: 3 ulp = 1.8e-7 absolute)
    assert np.array_equal(_bits(f[:, 1:]), _bits(GOLD[f"{case}/read_feats"][:, 1:]))
    assert _ulp_diff(f[:, 0], GOLD[f"{case}/read_feats"][:, 0]) <= 3
    d = G.load_data(ref_cloud, src_cloud, 30000)
    for k in ("ref_points", "src_points"):
        assert np.array_equal(_bits(d[k].cpu().numpy()), _bits(GOLD[f"{case}/{k}"])), k
    for k in ("ref_feats", "src_feats"):
        g = d[k].cpu().numpy()
        assert np.array_equal(_bits(g[:, 1:]), _bits(GOLD[f"{case}/{k}"][:, 1:])), k
        assert _ulp_diff(g[:, 0], GOLD[f"{case}/{k}"][:, 0]) <= 3
    for k in ("ref_adjust_scale", "src_adjust_scale", "ref_center", "src_center"):
        assert np.array_equal(np.asarray(d[k]), GOLD[f"{case}/{k}"]), k
    T = G.unnormalize_transform(torch.from_numpy(GOLD[f"{case}/transform_in"]).cuda(), d["ref_adjust_scale"], d["src_adjust_scale"],
                                d["ref_center"], d["src_center"])
    assert np.array_equal(T, GOLD[f"{case}/transform_scale"])


def test_full_size_cloud_vs_oracle():
    """BASELINE size (30k kept of a 60k cloud): device path vs the numpy oracle, same bits."""
    cloud = make_test_cloud(99, 60000, 1.3)
    pts, feats, idx = G.read_cloud_by_opacity(cloud, None)
    wp, wf, wi = og.read_cloud_by_opacity(cloud, None)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(_bits(pts.cpu().numpy()), _bits(wp))
    f = feats.cpu().numpy()
    assert np.array_equal(_bits(f[:, 1:]), _bits(wf[:, 1:]))
    assert _ulp_diff(f[:, 0], wf[:, 0]) <= 3
    # idempotence property: selecting from the already selected cloud with an open crop keeps every row
    sel = torch.from_numpy(cloud).cuda()[idx]
    p2, _, i2 = G.read_cloud_by_opacity(sel, None, opacity_min=0.0, crop_percent=(0, 100))
    assert i2.numel() <= sel.shape[0]


def test_point_limit_and_edge_cases():
    cloud = make_test_cloud(5, 3000)
    with pytest.raises(RuntimeError):
        G.read_cloud_by_opacity(cloud[:, :10])   # not a 59-attribute cloud
    dead = cloud.copy()
    dead[:, 51] = -5.0                            # nothing passes opacity > 0.7
    with pytest.raises(RuntimeError):
        G.read_cloud_by_opacity(dead)


def _fps_numpy(points, k, start):
    """Exact farthest-point sampling, float32 squared distances (dx*dx + dy*dy + dz*dz), ties to the lowest index."""
    pts = points.astype(np.float32)
    sel = [start]
    d = None
    for _ in range(k - 1):
        diff = pts - pts[sel[-1]]
        cur = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        d = cur if d is None else np.minimum(d, cur)
        sel.append(int(np.argmax(d)))
    return np.array(sel, np.int64)


@pytest.mark.parametrize("n,k,start", [(5000, 700, 0), (100000, 300, 17), (257, 257, 5), (1, 1, 0)])
def test_farthest_point_sample_exact(n, k, start):
    rng = np.random.default_rng(n + k)
    pts = rng.normal(size=(n, 3)).astype(np.float32)
    pts[n // 3] = pts[0]  # duplicates: ties resolved to the lowest index
    got = G.farthest_point_sample(torch.from_numpy(pts), k, start_idx=start).cpu().numpy()
    assert np.array_equal(got, _fps_numpy(pts, k, start))
    assert len(set(got.tolist())) == min(k, len(np.unique(pts, axis=0))) or k > len(np.unique(pts, axis=0))


def test_point_limit_subsamples_like_the_reference_demo():
    """demo.py:44-47: more survivors than point_limit -> farthest-point subsample of the survivors, selection order."""
    cloud = make_test_cloud(5, 6000)
    p_all, f_all, i_all = G.read_cloud_by_opacity(cloud, None)
    m = i_all.numel()
    limit = m // 3
    p, f, i = G.read_cloud_by_opacity(cloud, limit)
    assert p.shape == (limit, 3) and f.shape == (limit, 4) and i.shape == (limit,)
    want = i_all.cpu().numpy()[_fps_numpy(p_all.cpu().numpy(), limit, 0)]
    assert np.array_equal(i.cpu().numpy(), want)
    pos = {int(v): j for j, v in enumerate(i_all.tolist())}
    rows = [pos[int(v)] for v in i.tolist()]
    # same Gaussians (opacity column); the colours differ because the SH view point follows the selected set's centroid
    assert torch.equal(p.cpu(), p_all.cpu()[rows]) and torch.equal(f.cpu()[:, 0], f_all.cpu()[rows, 0])
    # fills space: the subsample's nearest-neighbour spacing is far larger than the full set's
    d_sub = torch.cdist(p[:500], p).topk(2, largest=False).values[:, 1].median()
    d_all = torch.cdist(p_all[:500], p_all).topk(2, largest=False).values[:, 1].median()
    assert d_sub > 1.3 * d_all


def test_ply_file_to_registration_input(tmp_path):
    """N2 + N1 + G3: 3DGS .ply files -> load_data -> collate, all through the public API."""
    from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
    from gaussreg_b200.data import registration_collate_fn_stack_mode
    ref_cloud, src_cloud = make_test_cloud(41, 5000), make_test_cloud(42, 5000)
    rp, sp = os.path.join(tmp_path, "ref.ply"), os.path.join(tmp_path, "src.ply")
    G.write_gaussian_ply(rp, ref_cloud)
    G.write_gaussian_ply(sp, src_cloud)
    d = G.load_data(rp, sp, 30000)
    want = og.load_data(ref_cloud, src_cloud, 30000)
    assert np.array_equal(_bits(d["ref_points"].cpu().numpy()), _bits(want["ref_points"]))
    cfg = make_cfg()
    data = registration_collate_fn_stack_mode([d], cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius,
                                              NEIGHBOR_LIMITS)
    assert data["features"].shape == (d["ref_points"].shape[0] + d["src_points"].shape[0], 4)
    assert len(data["points"]) == cfg.backbone.num_stages


def test_demo_end_to_end(tmp_path):
    """demo.py equivalent: two 3DGS .ply files in, estimated_transform.npz (un-normalised) + point clouds out."""
    import types
    from gaussreg_b200 import demo
    from gaussreg_b200.synthetic import make_gaussian_pair
    ref, src, _ = make_gaussian_pair(3, 6000)
    ref[:, :3] *= 2.0   # volume 240 m^3 -> the > 50 rescale branch of demo.py:96-98
    src[:, :3] *= 2.0
    rp, sp = os.path.join(tmp_path, "ref.ply"), os.path.join(tmp_path, "src.ply")
    G.write_gaussian_ply(rp, ref)
    G.write_gaussian_ply(sp, src)
    out_dir = os.path.join(tmp_path, "out")
    args = types.SimpleNamespace(ref_file=rp, src_file=sp, output_path=out_dir, weights="random:0", num_sample=30000)
    T, path = demo.run(args)
    saved = np.load(path)["estimated_transform"]
    assert saved.shape == (4, 4) and np.array_equal(saved, T) and T[3, 3] == 1 and np.isfinite(T).all()
    for name in ("point_cloud_ref.ply", "point_cloud_src.ply", "point_cloud_src_org.ply"):
        assert os.path.getsize(os.path.join(out_dir, name)) > 1000
    # the linear part is rotation * (src_scale / ref_scale): its determinant is that ratio cubed
    d = og.load_data(ref, src, 30000)
    ratio = float(d["src_adjust_scale"]) / float(d["ref_adjust_scale"])
    assert abs(np.linalg.det(T[:3, :3].astype(np.float64)) - ratio ** 3) < 1e-3
