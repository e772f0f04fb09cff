"""CPU tests: the plain-C oracle (oracle/neighbors.c) against the reference-generated golden
fixtures, against the compiled reference itself when oracle/_ref exists, and self-consistency."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from gaussreg_b200.synthetic import make_pair_inputs
from oracle import neighbors as on

GOLD = os.path.join(os.path.dirname(__file__), "golden", "neighbors_golden.npz")
CASES = ["room_1500", "room_1500_coarse", "box_1200"]


def _case(gold, name):
    seed, n, voxel, radius = gold[f"{name}/meta"]
    d = make_pair_inputs(int(seed), int(n), geometry=str(gold[f"{name}/geom"]))
    pts = np.concatenate([d["ref_points"], d["src_points"]]).astype(np.float32)
    lens = np.array([int(n), int(n)], np.int64)
    return pts, lens, float(voxel), float(radius)


@pytest.mark.parametrize("name", CASES)
def test_port_matches_golden(name):
    gold = np.load(GOLD)
    pts, lens, voxel, radius = _case(gold, name)
    P = on.port()
    sp, sl = P.grid_subsampling(pts, lens, voxel)
    assert np.array_equal(sl, gold[f"{name}/sub_lengths"])
    # bit-identical values AND emission order
    assert np.array_equal(sp.view(np.uint32), gold[f"{name}/sub_points"].view(np.uint32))
    assert np.array_equal(P.radius_neighbors(pts, pts, lens, lens, radius), gold[f"{name}/self"])
    assert np.array_equal(P.radius_neighbors(sp, pts, sl, lens, radius), gold[f"{name}/down"])
    assert np.array_equal(P.radius_neighbors(pts, sp, lens, sl, radius * 2), gold[f"{name}/up"])


def test_grid_and_brute_agree():
    rng = np.random.default_rng(5)
    q = rng.normal(size=(700, 3)).astype(np.float32)
    s = rng.normal(size=(900, 3)).astype(np.float32)
    ql = np.array([300, 400], np.int64)
    sl = np.array([500, 400], np.int64)
    P = on.port()
    for r in (0.05, 0.3, 1.0, 10.0):
        a = P.radius_neighbors(q, s, ql, sl, r)
        b = P.radius_neighbors(q, s, ql, sl, r, brute=True)
        assert np.array_equal(a, b)


def test_bucket_ladder():
    """k_ladder in oracle/neighbors.c (and c_ladder in csrc/neighbors.cu) == this box's libstdc++."""
    src = os.path.join(os.path.dirname(on.__file__), "probe_ladder.cpp")
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "probe")
        subprocess.check_call(["g++", "-O2", "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    measured = [int(line.split("nb=")[1]) for line in out.splitlines() if "nb=" in line]
    assert measured == on.port().ladder()[: len(measured)]
    cu = open(os.path.join(os.path.dirname(on.__file__), "..", "gaussreg_b200", "csrc", "neighbors.cu")).read()
    body = cu.split("c_ladder[23] = {")[1].split("}")[0]
    assert [int(t.strip().rstrip("u")) for t in body.split(",")] == on.port().ladder()


@pytest.mark.skipif(not (on.have_ref() or os.path.isdir("/root/reference")), reason="reference build unavailable")
@pytest.mark.parametrize("n,geom", [(4000, "room"), (3000, "box")])
def test_port_matches_reference_pyramid(n, geom):
    d = make_pair_inputs(21, n, geometry=geom)
    pts = np.concatenate([d["ref_points"], d["src_points"]]).astype(np.float32)
    lens = np.array([n, n], np.int64)
    limits = [89, 30, 43, 49, 49]
    a = on.precompute_data_stack_mode(on.ref(), pts, lens, 5, 0.025, 0.0625, limits)
    b = on.precompute_data_stack_mode(on.port(), pts, lens, 5, 0.025, 0.0625, limits)
    for i in range(5):
        assert np.array_equal(a["lengths"][i], b["lengths"][i])
        assert np.array_equal(a["points"][i].view(np.uint32), b["points"][i].view(np.uint32))
    for key, qs in (("neighbors", lambda i: (i, i)), ("subsampling", lambda i: (i + 1, i)), ("upsampling", lambda i: (i, i + 1))):
        for i, (x, y) in enumerate(zip(a[key], b[key])):
            qi, si = qs(i)
            xc, _ = on.canonicalize_ties(x, a["points"][qi], a["points"][si], a["points"][si].shape[0])
            # rows whose tie straddles the truncation column may legitimately differ: none expected here
            assert np.array_equal(xc, y), (key, i)


def test_edge_cases():
    P = on.port()
    # single point, duplicates (distance ties at d = 0), three ragged clouds
    pts = np.array([[0, 0, 0], [1, 1, 1], [1, 1, 1], [1, 1, 1], [5, 5, 5], [5.01, 5, 5]], np.float32)
    lens = np.array([1, 3, 2], np.int64)
    sp, sl = P.grid_subsampling(pts, lens, 0.5)
    assert sl.tolist() == [1, 1, 1]
    nb = P.radius_neighbors(pts, pts, lens, lens, 0.1)
    assert nb.shape == (6, 3)
    assert nb[0].tolist() == [0, 6, 6]
    assert nb[1].tolist() == [1, 2, 3] and nb[3].tolist() == [1, 2, 3]  # ties by ascending index
    assert nb[4].tolist() == [4, 5, 6] and nb[5].tolist() == [5, 4, 6]
    # a query outside every support ball -> width 0 table
    q = np.array([[100, 100, 100]], np.float32)
    assert P.radius_neighbors(q, pts[:1], np.array([1]), np.array([1]), 0.1).shape == (1, 0)
