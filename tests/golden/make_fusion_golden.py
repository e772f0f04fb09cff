"""TEST INFRASTRUCTURE: golden vectors for the Gaussian merge (N4), produced by the UNMODIFIED reference function
`gaussian_fuse` of /root/reference/gs_fusion.py (and its helpers sh_rotation / quaternion_to_matrix /
matrix_to_quaternion / load_ply), executed in this container.

Stand-ins (third-party I/O only; the reference arithmetic is untouched):
  * `plyfile.PlyData.read(path)` -> in-memory (N,59) clouds (same stand-in as make_gaussian_golden.py);
  * `gs_fusion.save_ply` is replaced by a function that captures its arguments (it only formats a PLY file);
  * `numpy.random.seed(1234)` before the call: sh_rotation draws its 15 probe directions from numpy's global RNG.

    python tests/golden/make_fusion_golden.py     # writes tests/golden/fusion_golden.npz
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_gaussian_golden as mg  # noqa: E402

SEED = 1234
CASES = {
    # name: (seed_1, seed_2, n_1, n_2, scale, angle, axis, translation)
    "rigid": (51, 52, 1500, 1300, 1.0, 0.5, (0.0, 0.0, 1.0), (0.3, -0.2, 0.1)),
    "similarity": (61, 62, 1200, 1700, 1.37, -0.9, (1.0, 2.0, -0.5), (-1.5, 0.4, 2.0)),
}


def similarity(scale, angle, axis, t):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = scale * R, t
    return T.astype(np.float32)  # demo.py saves the float32 transform


def main():
    mg.install_stubs()
    sys.path.insert(0, "/root/reference")
    import gs_fusion as gs  # the reference's gs_fusion.py

    out = {}
    for name, (s1, s2, n1, n2, scale, angle, axis, t) in CASES.items():
        c1, c2 = mg.test_cloud(s1, n1), mg.test_cloud(s2, n2)
        mg._REGISTRY["a.ply"], mg._REGISTRY["b.ply"] = c1, c2
        T = similarity(scale, angle, axis, t)
        captured = {}

        def capture(xyz, f_dc, f_rest, opacities, scale_, rotation, path):
            captured.update(xyz=xyz, f_dc=f_dc, f_rest=f_rest, opacities=opacities, scale=scale_, rotation=rotation)

        gs.save_ply = capture
        with tempfile.TemporaryDirectory() as d:
            tp = os.path.join(d, "t.npz")
            np.savez(tp, estimated_transform=T)
            np.random.seed(SEED)
            gs.gaussian_fuse("a.ply", "b.ply", tp, os.path.join(d, "out.ply"))
        fused = np.concatenate([captured["xyz"], captured["f_dc"], captured["f_rest"], captured["opacities"], captured["scale"],
                                captured["rotation"]], axis=1).astype(np.float32)  # save_ply writes 'f4'
        out[f"{name}/transform"] = T
        out[f"{name}/spec"] = np.array([s1, s2, n1, n2], np.int64)
        out[f"{name}/fused"] = fused
        print(name, "fused", fused.shape, "from", n1, "+", n2)
    np.savez_compressed(os.path.join(HERE, "fusion_golden.npz"), **out)
    print("wrote fusion_golden.npz,", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
