"""Generates tests/golden/neighbors_golden.npz from the UNMODIFIED reference CPU extension
(oracle/_ref/libgaussreg_ref.so, built from /root/reference by oracle/Makefile).

Run where /root/reference exists:   python tests/golden/make_neighbor_golden.py
The fixture pins oracle/neighbors.c (CPU tests) and the CUDA path (GPU tests) to the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import neighbors as on  # noqa: E402
from gaussreg_b200.synthetic import make_pair_inputs  # noqa: E402

CASES = [
    # name, seed, n_points, geometry, voxel, radius
    ("room_1500", 11, 1500, "room", 0.05, 0.0625),
    ("room_1500_coarse", 12, 1500, "room", 0.2, 0.25),
    ("box_1200", 13, 1200, "box", 0.1, 0.125),
]


def main():
    R = on.ref()
    out = {}
    for name, seed, n, geom, voxel, radius in CASES:
        d = make_pair_inputs(seed, n, geometry=geom)
        pts = np.concatenate([d["ref_points"], d["src_points"]]).astype(np.float32)
        lens = np.array([n, n], np.int64)
        sp, sl = R.grid_subsampling(pts, lens, voxel)
        nb_self = R.radius_neighbors(pts, pts, lens, lens, radius)
        nb_down = R.radius_neighbors(sp, pts, sl, lens, radius)
        nb_up = R.radius_neighbors(pts, sp, lens, sl, radius * 2)
        nb_down_c, _ = on.canonicalize_ties(nb_down, sp, pts, pts.shape[0])
        nb_self_c, _ = on.canonicalize_ties(nb_self, pts, pts, pts.shape[0])
        nb_up_c, _ = on.canonicalize_ties(nb_up, pts, sp, sp.shape[0])
        out[f"{name}/meta"] = np.array([seed, n, voxel, radius], np.float64)
        out[f"{name}/geom"] = np.array(geom)
        out[f"{name}/sub_points"] = sp
        out[f"{name}/sub_lengths"] = sl
        out[f"{name}/self"] = nb_self_c.astype(np.int32)
        out[f"{name}/down"] = nb_down_c.astype(np.int32)
        out[f"{name}/up"] = nb_up_c.astype(np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neighbors_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
