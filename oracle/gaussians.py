"""TEST INFRASTRUCTURE -- numpy restatement of the reference's Gaussian-cloud preparation (row N1/N2 of
SURVEY.md section 8(f)).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Follows, line by line in behaviour:
  * experiments/geotransformer.gaussian_splatting.indoor/demo.py:30-75  `_read_ply_by_opacity`
    (the PLY columns arrive here as an (N,59) float32 array in 3DGS property order without normals,
    gs_fusion.py:172-184; the farthest-point subsampling of demo.py:45-48 is the third-party `fpsample==0.3.2`
    -> parity unpinned, not restated: inputs must already satisfy `count <= point_limit`);
  * demo.py:81-124 `load_data` (bounding-box centring, volume rescale);
  * demo.py:173-178 the un-normalisation of the estimated transform;
  * geotransformer/utils/graphics_utils.py:34-89 `eval_sh` for deg = 3.

Pinned by tests/golden/gaussian_golden.npz, produced by running the UNMODIFIED reference functions in this
container (tests/golden/make_gaussian_golden.py: `plyfile` is replaced by an in-memory stand-in, nothing else).
"""
import numpy as np

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)

COL_XYZ, COL_FDC, COL_FREST, COL_OPACITY = 0, 3, 6, 51
ATTR_DIM = 59


def eval_sh_deg3(sh, dirs):
    """graphics_utils.py:34-89 with deg = 3.  sh (..., C, 16), dirs (..., 3) -> (..., C).  The expression trees are
    kept exactly as the reference writes them (numpy rounds every binary operation separately)."""
    result = C0 * sh[..., 0]
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    result = (result - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3])
    xx, yy, zz = x * x, y * y, z * z
    xy, yz, xz = x * y, y * z, x * z
    result = (result + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] +
              C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
    result = (result + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10] +
              C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
              C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14] +
              C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def read_cloud_by_opacity(cloud, point_limit=None):
    """demo.py:30-75 on an (N,59) float32 cloud.  Returns (points (M,3) f32, point_features (M,4) f32, index (M,))."""
    cloud = np.asarray(cloud, dtype=np.float32)
    opacity = cloud[:, COL_OPACITY].copy()
    opacity = 1 / (1 + np.exp(-opacity))                                       # :34 (float32)
    x, y, z = cloud[:, 0].copy(), cloud[:, 1].copy(), cloud[:, 2].copy()
    index_x = (x < np.percentile(x, 95)) * (x > np.percentile(x, 5))           # :40
    index_y = (y < np.percentile(y, 95)) * (y > np.percentile(y, 5))
    index_z = (z < np.percentile(z, 95)) * (z > np.percentile(z, 5))
    index = np.where((opacity > 0.7) * index_x * index_y * index_z)[0]         # :43
    points = np.stack([x, y, z], axis=1)
    if point_limit is not None and index.shape[0] > point_limit:
        raise NotImplementedError("farthest-point sampling (fpsample==0.3.2, demo.py:45-48) is third-party: parity unpinned")
    features_dc = np.zeros((points.shape[0], 3, 1))                            # float64, :49-52
    features_dc[:, :, 0] = cloud[:, COL_FDC:COL_FDC + 3]
    features_extra = np.zeros((points.shape[0], 45))
    features_extra[:, :] = cloud[:, COL_FREST:COL_FREST + 45]
    features_extra = features_extra.reshape((features_extra.shape[0], 3, 15))   # :60
    features = np.concatenate([features_dc, features_extra], axis=2)[index]     # (M,3,16)
    points = points[index]
    center_point = points.mean(0)                                               # float32, :63
    max_length = np.linalg.norm(points.max(axis=0) - points.min(axis=0))
    center_point = center_point + np.array([0, 2 * max_length, 0])              # float64 from here
    dir_pp = points - center_point[None, :].repeat(points.shape[0], 0)
    dir_pp_normalized = dir_pp / (np.linalg.norm(dir_pp, axis=1, keepdims=True) + 1e-6)
    sh2rgb = eval_sh_deg3(features, dir_pp_normalized)
    colors = np.clip(sh2rgb + 0.5, 0.0, 1.0) * 255
    point_features = np.concatenate([opacity[index].reshape(points.shape[0], -1), colors.astype(np.float32)], axis=1)
    return points, point_features, index


def _center_and_scale(points):
    """demo.py:85-110 for one cloud: (points', adjust_scale, center)."""
    volume = ((points[:, 0].max() - points[:, 0].min()) * (points[:, 1].max() - points[:, 1].min()) *
              (points[:, 2].max() - points[:, 2].min()))
    center = (points.max(0) + points.min(0)) / 2
    points = points - center
    adjust_scale = 1.
    if volume > 50:
        adjust_scale = (50 / volume) ** (1 / 3)
        points = points * adjust_scale
    elif volume < 10:
        adjust_scale = (30 / volume) ** (1 / 3)
        points = points * adjust_scale
    return points, adjust_scale, center


def load_data(ref_cloud, src_cloud, num_sample=30000):
    """demo.py:81-124: the dict handed to registration_collate_fn_stack_mode."""
    ref_points, ref_feats, _ = read_cloud_by_opacity(ref_cloud, num_sample)
    src_points, src_feats, _ = read_cloud_by_opacity(src_cloud, num_sample)
    ref_points, ref_adjust_scale, ref_center = _center_and_scale(ref_points)
    src_points, src_adjust_scale, src_center = _center_and_scale(src_points)
    return {
        "ref_points": ref_points.astype(np.float32), "src_points": src_points.astype(np.float32),
        "ref_feats": ref_feats.astype(np.float32), "src_feats": src_feats.astype(np.float32),
        "ref_adjust_scale": ref_adjust_scale, "src_adjust_scale": src_adjust_scale,
        "ref_center": ref_center, "src_center": src_center,
    }


def unnormalize_transform(estimated_transform, ref_adjust_scale, src_adjust_scale, ref_center, src_center):
    """demo.py:173-178: the similarity transform between the ORIGINAL (un-centred, un-scaled) clouds."""
    T = np.zeros_like(estimated_transform)
    T[:3, :3] = estimated_transform[:3, :3] / ref_adjust_scale * src_adjust_scale
    T[:3, 3] = estimated_transform[:3, 3] / ref_adjust_scale + ref_center - np.matmul(T[:3, :3], src_center)
    T[3, 3] = 1.
    return T
