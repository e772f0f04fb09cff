"""Deterministic synthetic Gaussian-splat scene pairs.

A "Gaussian cloud" is ``N x 59`` float32 in 3DGS PLY property order without normals
(reference: gs_fusion.py:172-184): ``xyz(3) f_dc(3) f_rest(45) opacity(1, logit) scale(3, log) rot(4)``.
The network input is derived from it the way the reference demo does
(experiments/geotransformer.gaussian_splatting.indoor/demo.py:63-72): ``[sigmoid(opacity), RGB*255]``
with RGB = clip(SH(deg 3, view direction) + 0.5, 0, 1), the view point being the cloud centroid
shifted by twice the bounding-box diagonal along +y.

Geometry ("room-shell"): points area-weighted on the six faces of a 4.0 x 3.0 x 2.5 m box with
N(0, 1 cm) jitter; ``src`` is an independent resample of the same room moved by a known rigid
transform (0.5 rad about z, t = (0.3, -0.2, 0.1)).  Both clouds are then centred on their bounding
box centre as demo.py:85-93 does (the 30 m^3 room needs no volume rescale, demo.py:96-110).
"""
import numpy as np

ATTR_DIM = 59
ROOM = (4.0, 3.0, 2.5)
GT_ANGLE = 0.5
GT_TRANSLATION = (0.3, -0.2, 0.1)

# real SH basis constants up to degree 3 (same polynomials as 3DGS / graphics_utils.py:3-21)
_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
          -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


def sh_basis_deg3(dirs):
    """(N,3) unit directions -> (N,16) real SH basis values (degree <= 3)."""
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    xx, yy, zz = x * x, y * y, z * z
    xy, yz, xz = x * y, y * z, x * z
    b = np.empty((dirs.shape[0], 16), dtype=dirs.dtype)
    b[:, 0] = _SH_C0
    b[:, 1] = -_SH_C1 * y
    b[:, 2] = _SH_C1 * z
    b[:, 3] = -_SH_C1 * x
    b[:, 4] = _SH_C2[0] * xy
    b[:, 5] = _SH_C2[1] * yz
    b[:, 6] = _SH_C2[2] * (2.0 * zz - xx - yy)
    b[:, 7] = _SH_C2[3] * xz
    b[:, 8] = _SH_C2[4] * (xx - yy)
    b[:, 9] = _SH_C3[0] * y * (3 * xx - yy)
    b[:, 10] = _SH_C3[1] * xy * z
    b[:, 11] = _SH_C3[2] * y * (4 * zz - xx - yy)
    b[:, 12] = _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy)
    b[:, 13] = _SH_C3[4] * x * (4 * zz - xx - yy)
    b[:, 14] = _SH_C3[5] * z * (xx - yy)
    b[:, 15] = _SH_C3[6] * x * (xx - 3 * yy)
    return b


def gaussian_features(cloud):
    """(N,59) Gaussian attributes -> (N,4) float32 network features [opacity, R, G, B] (demo.py:63-72, no filtering).

    Part of the SYNTHETIC INPUT GENERATOR only (bench / tests build their host-side pairs with it).  The product's
    implementation of this step is gaussians.read_cloud_by_opacity (csrc/gaussians.cu)."""
    cloud = np.asarray(cloud)
    pts = cloud[:, 0:3].astype(np.float64)
    f_dc = cloud[:, 3:6].astype(np.float64)  # (N,3)
    f_rest = cloud[:, 6:51].astype(np.float64).reshape(-1, 3, 15)
    sh = np.concatenate([f_dc[:, :, None], f_rest], axis=2)  # (N,3,16)
    opacity = 1.0 / (1.0 + np.exp(-cloud[:, 51].astype(np.float32)))
    center = pts.mean(0)
    diag = np.linalg.norm(pts.max(0) - pts.min(0))
    center = center + np.array([0.0, 2.0 * diag, 0.0])
    d = pts - center[None, :]
    d = d / (np.linalg.norm(d, axis=1, keepdims=True) + 1e-6)
    rgb = np.einsum("ncb,nb->nc", sh, sh_basis_deg3(d))
    colors = np.clip(rgb + 0.5, 0.0, 1.0) * 255.0
    return np.concatenate([opacity.reshape(-1, 1).astype(np.float32), colors.astype(np.float32)], axis=1)


def _room_points(rng, n, room):
    lx, ly, lz = room
    areas = np.array([ly * lz, ly * lz, lx * lz, lx * lz, lx * ly, lx * ly])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    u = rng.random(n)
    v = rng.random(n)
    p = np.empty((n, 3))
    for f in range(6):
        m = face == f
        axis, side = f // 2, f % 2
        a, b = [(1, 2), (0, 2), (0, 1)][axis]
        p[m, axis] = side * room[axis]
        p[m, a] = u[m] * room[a]
        p[m, b] = v[m] * room[b]
    p += rng.normal(0.0, 0.01, size=p.shape)
    return p - np.asarray(room) / 2.0


def _attributes(rng, pts, textured):
    n = pts.shape[0]
    if textured:
        # colour is a smooth function of position so that both clouds of a pair see the same "texture"
        ph = pts @ np.array([[1.7, 0.4, -0.9], [-0.6, 2.1, 0.8], [0.5, -1.3, 1.9]])
        f_dc = 0.8 * np.sin(ph) + rng.normal(0.0, 0.02, size=(n, 3))
        f_rest = rng.normal(0.0, 0.02, size=(n, 45))
    else:
        f_dc = rng.normal(0.0, 0.3, size=(n, 3))
        f_rest = rng.normal(0.0, 0.3, size=(n, 45))
    opacity = rng.uniform(1.0, 4.0, size=(n, 1))
    scale = rng.normal(-4.0, 0.5, size=(n, 3))
    rot = rng.normal(0.0, 1.0, size=(n, 4))
    rot /= np.linalg.norm(rot, axis=1, keepdims=True)
    return f_dc, f_rest, opacity, scale, rot


def gt_transform(angle=GT_ANGLE, translation=GT_TRANSLATION):
    c, s = np.cos(angle), np.sin(angle)
    T = np.eye(4)
    T[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    T[:3, 3] = translation
    return T


def make_gaussian_pair(seed, n_points=30000, geometry="room", textured=False, angle=GT_ANGLE,
                       translation=GT_TRANSLATION, room=ROOM):
    """Returns ``(ref_cloud, src_cloud, src_to_ref)``: two (N,59) float32 Gaussian clouds (already
    bbox-centred) and the 4x4 float64 transform with ``ref ~= R @ src + t``.

    ``geometry='box'`` fills the room volume uniformly instead (denser coarse stages: neighbour
    widths hit the 43/49 limits)."""
    rng = np.random.default_rng(seed)
    clouds = []
    T_move = gt_transform(angle, translation)
    offsets = []
    for which in range(2):
        if geometry == "room":
            p = _room_points(rng, n_points, room)
        elif geometry == "box":
            p = (rng.random((n_points, 3)) - 0.5) * np.asarray(room)
        else:
            raise ValueError(f"unknown geometry {geometry!r}")
        attrs = _attributes(rng, p, textured)
        if which == 1:
            p = p @ T_move[:3, :3].T + T_move[:3, 3]
        centre = (p.max(0) + p.min(0)) / 2.0
        p = p - centre
        offsets.append(centre)
        clouds.append(np.concatenate([p, *attrs], axis=1).astype(np.float32))
    # room frame X -> ref: X - c0 ; src: R X + t - c1   =>  ref = R^T (src + c1 - t) - c0
    R = T_move[:3, :3]
    T = np.eye(4)
    T[:3, :3] = R.T
    T[:3, 3] = R.T @ (offsets[1] - T_move[:3, 3]) - offsets[0]
    return clouds[0], clouds[1], T


def make_pair_inputs(seed, n_points=30000, **kw):
    """Network-ready dict for one pair, the shape demo.py:112-123 hands to the collate function."""
    ref, src, T = make_gaussian_pair(seed, n_points, **kw)
    return {
        "ref_points": np.ascontiguousarray(ref[:, :3]),
        "src_points": np.ascontiguousarray(src[:, :3]),
        "ref_feats": gaussian_features(ref),
        "src_feats": gaussian_features(src),
        "transform": T.astype(np.float32),
    }
