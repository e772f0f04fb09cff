"""Merge two 3DGS scenes under the estimated similarity transform (reference: gs_fusion.py).

    python -m gaussreg_b200.fusion --root_path scene --transform_path demo_outputs/estimated_transform.npz

Same arguments, same file layout and the same arithmetic as the reference script: the second cloud's positions, log-scales,
rotation quaternions and spherical-harmonic bands 1-3 are moved by the transform (csrc/fusion.cu, one pass over the (N,59)
rows), each cloud keeps the Gaussians that are nearer to its own centroid than to the other's, and the result is written as a
3DGS `point_cloud.ply`.  The 3x3 / 5x5 / 7x7 band matrices are built once on the host from 15 random probe directions drawn
from numpy's global generator exactly like gs_fusion.py:53-68 (seed numpy to reproduce the reference bit for bit); they do
not depend on the Gaussian, so the reference's N*3 batched pseudo-inverses are not repeated.
"""
import argparse
import ctypes
import os
import shutil

import numpy as np
import torch

from . import _lib, gaussians
from .ext import _stream

C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)


def _band_values(dirs):
    """SH basis values of bands 1 / 2 / 3 at probe directions 0:3 / 3:8 / 8:15 (gs_fusion.py:9-51): rows = directions."""
    d1, d2, d3 = dirs[0:3], dirs[3:8], dirs[8:15]
    x, y, z = d1[:, 0], d1[:, 1], d1[:, 2]
    b1 = np.stack([-C1 * y, C1 * z, -C1 * x], axis=1)
    x, y, z = d2[:, 0], d2[:, 1], d2[:, 2]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    b2 = np.stack([C2[0] * xy, C2[1] * yz, C2[2] * (2.0 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)], axis=1)
    x, y, z = d3[:, 0], d3[:, 1], d3[:, 2]
    xx, yy, zz, xy = x * x, y * y, z * z, x * y
    b3 = np.stack([C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy), C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
                   C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy)], axis=1)
    return b1, b2, b3


def sh_band_transforms(rotation, dirs=None):
    """gs_fusion.py:53-68 for all Gaussians at once: M_l = pinv(B_l(dirs)) @ B_l(dirs @ R^T), l = 1, 2, 3 (float64)."""
    if dirs is None:
        dirs = np.random.randn(15, 3)
        dirs = dirs / (np.linalg.norm(dirs, axis=1, keepdims=True) + 1e-8)
    rot = _band_values(dirs @ np.asarray(rotation).T)
    return [np.linalg.pinv(b) @ r for b, r in zip(_band_values(dirs), rot)]


def decompose_similarity(transform):
    """gs_fusion.py:236-239: (unit rotation, scale, translation) with the reference's dtypes."""
    transform = np.asarray(transform)
    rotation = transform[:3, :3]
    translation = transform[:3, 3]
    scale = (rotation @ rotation.T)[0, 0] ** 0.5
    return rotation / scale, scale, translation


def transform_gaussians(cloud, transform, dirs=None):
    """(N,59) cloud (numpy or tensor) under the similarity `transform` (4,4) -> CUDA tensor (N,59) float32."""
    cloud = gaussians._as_device_cloud(cloud)
    rotation, scale, translation = decompose_similarity(transform)
    m1, m2, m3 = sh_band_transforms(rotation, dirs)
    sh = np.ascontiguousarray(np.concatenate([m1.ravel(), m2.ravel(), m3.ravel()]).astype(np.float64))
    R = np.ascontiguousarray(np.asarray(rotation, np.float32))
    t = np.ascontiguousarray(np.asarray(translation, np.float32))
    out = torch.empty((cloud.shape[0], gaussians.ATTR_DIM), dtype=torch.float32, device=cloud.device)
    log_scale = float(np.log(scale))  # np.float32 in, np.float32 out when the transform is float32 (gs_fusion.py:241)
    st = _lib.lib().gr_gaussian_transform(cloud.data_ptr(), cloud.stride(0), cloud.shape[0], R.ctypes.data_as(ctypes.c_void_p),
                                          float(scale), log_scale, t.ctypes.data_as(ctypes.c_void_p),
                                          sh.ctypes.data_as(ctypes.c_void_p), out.data_ptr(), out.stride(0), _stream())
    _lib.check(st, "gaussian_transform")
    return out


def gaussian_fuse(cloud_1, cloud_2, transform, dirs=None):
    """gs_fusion.py:231-262 without the file I/O -> fused CUDA tensor (M,59) float32."""
    c1 = gaussians._as_device_cloud(cloud_1)
    c2 = transform_gaussians(cloud_2, transform, dirs)
    xyz_1, xyz_2 = c1[:, 0:3], c2[:, 0:3]
    ctr_1 = xyz_1.double().mean(0).float()
    ctr_2 = xyz_2.double().mean(0).float()
    keep_1 = torch.linalg.norm(xyz_1 - ctr_1, dim=1) < torch.linalg.norm(xyz_1 - ctr_2, dim=1)
    keep_2 = torch.linalg.norm(xyz_2 - ctr_2, dim=1) < torch.linalg.norm(xyz_2 - ctr_1, dim=1)
    return torch.cat([c1[keep_1][:, :gaussians.ATTR_DIM], c2[keep_2]], dim=0)


def gaussian_fuse_files(input_path_1, input_path_2, transform_path, output_path):
    """gs_fusion.gaussian_fuse: two 3DGS point_cloud.ply files + estimated_transform.npz -> fused point_cloud.ply."""
    c1 = gaussians.read_gaussian_ply(input_path_1)
    c2 = gaussians.read_gaussian_ply(input_path_2)
    transform = np.load(transform_path)["estimated_transform"]
    fused = gaussian_fuse(c1, c2, transform).cpu().numpy()
    gaussians.write_gaussian_ply(output_path, fused)
    return fused.shape[0]


def main(argv=None):
    parser = argparse.ArgumentParser(description="Fusion script parameters")
    parser.add_argument("--root_path", type=str, default="scene_name")
    parser.add_argument("--transform_path", type=str, default="demo_outputs/estimated_transform.npz")
    args, _ = parser.parse_known_args(argv)
    root = args.root_path
    in_1 = os.path.join(root, "A/output/point_cloud/iteration_30000/point_cloud.ply")
    in_2 = os.path.join(root, "B/output/point_cloud/iteration_30000/point_cloud.ply")
    os.makedirs(os.path.join(root, "fuse/output/point_cloud/iteration_30000"))
    for name in ("cameras.json", "cfg_args"):
        src = os.path.join(root, "A/output", name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(root, "fuse/output", name))
    out = os.path.join(root, "fuse/output/point_cloud/iteration_30000/point_cloud.ply")
    n = gaussian_fuse_files(in_1, in_2, args.transform_path, out)
    print(f"fused {n} Gaussians -> {out}")


if __name__ == "__main__":
    main()
