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


def _sh_basis_terms(dirs):
    """The 15 direction-only factors of the degree-1..3 real SH basis, each parenthesised the way Python evaluates
    the reference's products (graphics_utils.py:60-88 multiply left to right: constant x monomial [x polynomial]),
    paired with the sign with which the reference adds the term.  dirs (..., 3) -> list of (sign, (..., 1) array)."""
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    xx, yy, zz = x * x, y * y, z * z
    xy, yz, xz = x * y, y * z, x * z
    band1 = [(-1, C1 * y), (+1, C1 * z), (-1, C1 * x)]
    band2 = [(+1, C2[0] * xy), (+1, C2[1] * yz), (+1, C2[2] * (2.0 * zz - xx - yy)), (+1, C2[3] * xz), (+1, C2[4] * (xx - yy))]
    band3 = [(+1, C3[0] * y * (3 * xx - yy)), (+1, C3[1] * xy * z), (+1, C3[2] * y * (4 * zz - xx - yy)),
             (+1, C3[3] * z * (2 * zz - 3 * xx - 3 * yy)), (+1, C3[4] * x * (4 * zz - xx - yy)), (+1, C3[5] * z * (xx - yy)),
             (+1, C3[6] * x * (xx - 3 * yy))]
    return band1 + band2 + band3


def eval_sh_deg3(sh, dirs):
    """graphics_utils.py:34-89 with deg = 3.  sh (..., C, 16), dirs (..., 3) -> (..., C).  One running sum, term k
    added (or subtracted) as `basis_k * sh[..., k]` in coefficient order: the same sequence of individually rounded
    numpy operations as the reference's three chained expressions."""
    acc = C0 * sh[..., 0]
    for k, (sign, basis) in enumerate(_sh_basis_terms(dirs), start=1):
        term = basis * sh[..., k]
        acc = acc + term if sign > 0 else acc - term
    return acc


def read_cloud_by_opacity(cloud, point_limit=None):
    """demo.py:30-75 on an (N,59) float32 cloud.  Returns (points (M,3) f32, point_features (M,4) f32, index (M,))."""
    cloud = np.asarray(cloud, dtype=np.float32)
    alpha = cloud[:, COL_OPACITY].copy()
    alpha = 1 / (1 + np.exp(-alpha))                                            # :34, float32 throughout
    axes = [cloud[:, a].copy() for a in range(3)]
    inside = [(c < np.percentile(c, 95)) * (c > np.percentile(c, 5)) for c in axes]   # :40-42
    keep = np.where((alpha > 0.7) * inside[0] * inside[1] * inside[2])[0]       # :43
    if point_limit is not None and keep.shape[0] > point_limit:
        raise NotImplementedError("farthest-point sampling (fpsample==0.3.2, demo.py:45-48) is third-party: parity unpinned")
    xyz = np.stack(axes, axis=1)[keep]                                           # float32 (M,3)
    # float64 SH table (M, 3 channels, 16 coefficients): DC first, then the 15 higher-order ones per channel (:49-61)
    coeffs = np.zeros((keep.shape[0], 3, 16))
    coeffs[:, :, 0] = cloud[keep, COL_FDC:COL_FDC + 3]
    coeffs[:, :, 1:] = cloud[keep, COL_FREST:COL_FREST + 45].astype(np.float64).reshape(-1, 3, 15)
    eye = xyz.mean(0)                                                            # float32 mean, :63
    diagonal = np.linalg.norm(xyz.max(axis=0) - xyz.min(axis=0))
    eye = eye + np.array([0, 2 * diagonal, 0])                                   # float64 from here, :64
    view = xyz - eye[None, :].repeat(xyz.shape[0], 0)
    view = view / (np.linalg.norm(view, axis=1, keepdims=True) + 1e-6)
    rgb = np.clip(eval_sh_deg3(coeffs, view) + 0.5, 0.0, 1.0) * 255
    point_features = np.concatenate([alpha[keep].reshape(xyz.shape[0], -1), rgb.astype(np.float32)], axis=1)
    return xyz, point_features, keep


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
