"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the similarity RANSAC that the reference's model applies after
LocalGlobalRegistration (experiments/geotransformer.gaussian_splatting.indoor/model.py:209-215 ->
geotransformer/utils/open3d.py:169-198).

PARITY UNPINNED: the arithmetic lives in open3d==0.11.2 (environment.yaml:111), a third-party dependency that is not
vendored under the reference tree and is not installed here; the reference holds no golden vector for it, and Open3D
seeds its sampler from std::random_device.  What is restated is that release's published algorithm
(`RegistrationRANSACBasedOnCorrespondence` + `TransformationEstimationPointToPoint(with_scaling=True)` =
Eigen::umeyama): independent uniform draws of `ransac_n` correspondences, closed-form similarity, score on all
correspondences (fitness, then inlier RMSE), best sample returned without re-fit.  The sample stream is the same
counter-based generator as csrc/ransac.cu, so that both sides evaluate the same hypotheses.
"""
import numpy as np

_M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def sample_indices(seed, hyp, sample_n, n):
    return [splitmix64(seed ^ splitmix64((hyp << 8) | j)) % n for j in range(sample_n)]


def umeyama(src, ref):
    """Eigen::umeyama with scaling: (c, R, t) minimising sum |ref - (c R src + t)|^2.  None for degenerate input."""
    src, ref = np.asarray(src, np.float64), np.asarray(ref, np.float64)
    mu_s, mu_r = src.mean(0), ref.mean(0)
    ds, dr = src - mu_s, ref - mu_r
    var = (ds ** 2).sum() / src.shape[0]
    if not var > 1e-20:
        return None
    sigma = dr.T @ ds / src.shape[0]
    U, D, Vt = np.linalg.svd(sigma)
    if not (D[0] > 0 and D[1] > 1e-12 * D[0]):
        return None
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    c = float((D * S).sum() / var)
    t = mu_r - c * R @ mu_s
    return c, R, t


def similarity_ransac(ref_corr, src_corr, num_hypotheses=10000, sample_size=5, distance_threshold=0.05, seed=0, refit=False):
    """-> (T (4,4) float64, inliers of the best hypothesis, its index)."""
    ref_corr, src_corr = np.asarray(ref_corr, np.float32), np.asarray(src_corr, np.float32)
    n = ref_corr.shape[0]
    best = (0, np.inf, -1, np.eye(4))
    if n < 3:
        return best[3], 0, -1
    for h in range(num_hypotheses):
        idx = sample_indices(seed, h, sample_size, n)
        sol = umeyama(src_corr[idx], ref_corr[idx])
        if sol is None:
            continue
        c, R, t = sol
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = c * R, t
        T32 = T.astype(np.float32)
        d2 = ((src_corr @ T32[:3, :3].T + T32[:3, 3] - ref_corr) ** 2).sum(1)
        inl = d2 < np.float32(distance_threshold) ** 2
        k = int(inl.sum())
        if k == 0:
            continue
        rmse = float(np.sqrt(d2[inl].sum() / k))
        if k > best[0] or (k == best[0] and rmse < best[1]):
            best = (k, rmse, h, T)
    k, _, h, T = best
    if refit and h >= 0:
        T32 = T.astype(np.float32)
        d2 = ((src_corr @ T32[:3, :3].T + T32[:3, 3] - ref_corr) ** 2).sum(1)
        inl = d2 < np.float32(distance_threshold) ** 2
        sol = umeyama(src_corr[inl], ref_corr[inl]) if inl.sum() >= 3 else None
        if sol is not None:
            c, R, t = sol
            T = np.eye(4)
            T[:3, :3], T[:3, 3] = c * R, t
    return T, k, h
