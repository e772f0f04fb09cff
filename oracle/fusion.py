"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's Gaussian merge (gs_fusion.py:53-68, 70-159, 231-262):
apply the estimated similarity transform to the second 3DGS cloud (positions, log-scales, rotation quaternions, degree-1..3
spherical-harmonic bands), keep from each cloud the Gaussians nearer to its own centroid than to the other's, concatenate.

Pinned by tests/golden/fusion_golden.npz, produced by the UNMODIFIED reference `gaussian_fuse` (make_fusion_golden.py).
Clouds are (N,59) float32 in 3DGS property order without normals: xyz(3) f_dc(3) f_rest(45) opacity(1) scale(3) rot(4).
"""
import numpy as np

C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def band_values(dirs):
    """gs_fusion.py:9-51: SH basis values of bands 1, 2, 3 at the probe directions 0:3, 3:8, 8:15 -> (3,3), (5,5), (7,7),
    rows = directions, columns = basis functions."""
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
    """gs_fusion.py:53-68: M_l = pinv(B_l(dirs)) @ B_l(dirs @ R^T) for l = 1, 2, 3; rotated coefficients = coeffs @ M_l.
    `dirs` defaults to 15 normalised draws of numpy's GLOBAL generator, exactly as the reference consumes it."""
    if dirs is None:
        dirs = np.random.randn(15, 3)
        dirs = dirs / (np.linalg.norm(dirs, axis=1, keepdims=True) + 1e-8)
    rot = band_values(dirs @ np.asarray(rotation).T)
    return [np.linalg.pinv(b) @ r for b, r in zip(band_values(dirs), rot)]


def quaternion_to_matrix(q):
    """gs_fusion.py:70-98 (real part first), float32."""
    q = np.asarray(q, np.float32)
    r, i, j, k = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    two_s = np.float32(2.0) / (q * q).sum(-1)
    o = np.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                  two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                  two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], axis=-1)
    return o.reshape(-1, 3, 3).astype(np.float32)


def matrix_to_quaternion(m):
    """gs_fusion.py:111-159: the best-conditioned of the four candidates, float32."""
    m = np.asarray(m, np.float32).reshape(-1, 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = [m[:, c] for c in range(9)]
    q_abs = np.sqrt(np.maximum(np.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22,
                                         1.0 - m00 - m11 + m22], axis=-1).astype(np.float32), 0.0)).astype(np.float32)
    cand = np.stack([np.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
                     np.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], -1),
                     np.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], -1),
                     np.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], -1)], axis=-2).astype(np.float32)
    cand = cand / (np.float32(2.0) * np.maximum(q_abs[..., None], np.float32(0.1)))
    return cand[np.arange(m.shape[0]), q_abs.argmax(-1)].astype(np.float32)


def transform_cloud(cloud, transform, dirs=None):
    """The second cloud under the similarity transform (gs_fusion.py:236-246) -> (N,59) float32."""
    cloud = np.asarray(cloud, np.float32)
    transform = np.asarray(transform)
    rotation = transform[:3, :3]
    translation = transform[None, :3, 3]
    scale = (rotation @ rotation.T)[0, 0] ** 0.5
    rotation = rotation / scale
    out = cloud.astype(np.float64)
    out[:, 0:3] = cloud[:, 0:3] @ rotation.T * scale + translation
    if scale != 1.0:
        out[:, 52:55] = cloud[:, 52:55].astype(np.float64) + np.log(scale)
    rm = np.matmul(np.asarray(rotation, np.float32)[None], quaternion_to_matrix(cloud[:, 55:59]))
    out[:, 55:59] = matrix_to_quaternion(rm)
    sh = cloud[:, 6:51].astype(np.float64).reshape(-1, 3, 15)
    m1, m2, m3 = sh_band_transforms(rotation, dirs)
    out[:, 6:51] = np.concatenate([sh[:, :, 0:3] @ m1, sh[:, :, 3:8] @ m2, sh[:, :, 8:15] @ m3], axis=2).reshape(-1, 45)
    return out.astype(np.float32)


def gaussian_fuse(cloud_1, cloud_2, transform, dirs=None):
    """gs_fusion.py:231-262 without the file I/O -> fused (M,59) float32."""
    cloud_1 = np.asarray(cloud_1, np.float32)
    c2 = transform_cloud(cloud_2, transform, dirs)
    # the selection runs on the float32 positions exactly as the reference computes them (before the 'f4' cast of save_ply)
    t = np.asarray(transform)
    rotation = t[:3, :3]
    scale = (rotation @ rotation.T)[0, 0] ** 0.5
    xyz_1, xyz_2 = cloud_1[:, 0:3], np.asarray(cloud_2, np.float32)[:, 0:3] @ (rotation / scale).T * scale + t[None, :3, 3]
    ctr_1, ctr_2 = xyz_1.mean(0), xyz_2.mean(0)
    keep_1 = np.linalg.norm(xyz_1 - ctr_1, axis=1) < np.linalg.norm(xyz_1 - ctr_2, axis=1)
    keep_2 = np.linalg.norm(xyz_2 - ctr_2, axis=1) < np.linalg.norm(xyz_2 - ctr_1, axis=1)
    return np.concatenate([cloud_1[keep_1], c2[keep_2]], axis=0).astype(np.float32)
