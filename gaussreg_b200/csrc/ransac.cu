// N3: similarity-transform RANSAC over the LocalGlobalRegistration correspondences, on the device.
//
// Replaces the step with which the reference's model overwrites the LGR transform
// (experiments/geotransformer.gaussian_splatting.indoor/model.py:209-215 ->
//  geotransformer/utils/open3d.py:169-198): open3d==0.11.2 `registration_ransac_based_on_correspondence` with
// `TransformationEstimationPointToPoint(with_scaling=True)`, ransac_n = 5, 10 000 iterations, inlier distance 0.05.
// Open3D is a third-party dependency that is not vendored in the reference; the published algorithm of that
// release is restated here:
//   per iteration: draw ransac_n correspondences independently and uniformly (with replacement), estimate the
//   similarity transform with Umeyama's closed form (Eigen::umeyama: R = U S V^T, c = tr(D S) / var(src),
//   t = mu_ref - c R mu_src), score it on ALL correspondences (inlier: |T src - ref| < threshold;
//   fitness = #inliers / #correspondences, rmse over the inliers), keep it if the fitness is higher, or equal with
//   a lower rmse.  Release 0.11.2 returns the best sample's transform as is (no final re-fit); `refit` != 0 adds the
//   re-estimation on the best inlier set that later Open3D releases perform.
// Open3D draws from a global Mersenne twister seeded from std::random_device, so its result is not reproducible;
// this kernel uses a counter-based generator keyed by (seed, iteration, draw): deterministic for a given seed and
// independent of the launch geometry.  Parity with the reference is therefore statistical (tests compare rotation /
// translation / scale errors on well-conditioned pairs), and is documented as unpinned in DESIGN.md.
//
// One thread per hypothesis: the 5-point Umeyama solve (one-sided Jacobi SVD of the 3x3 covariance, fp64) costs a
// few hundred instructions, the scoring loop streams the correspondences from shared memory in tiles.  The number of
// correspondences is read from a DEVICE scalar (the count LocalGlobalRegistration leaves there): no host sync.
#include "common.cuh"

namespace gr {

constexpr int kRansacThreads = 128;
constexpr int kRansacTile = 1024;  // correspondences per shared-memory tile (24 KB)
constexpr int kRansacMaxSample = 8;

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// One-sided Jacobi SVD of a 3x3 matrix: on return A = U diag(sg) (columns), V orthogonal, A_in = A V^T.
__device__ void jacobi_svd3(double A[9], double V[9]) {
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int r = 0; r < 3; ++r) {
          al += A[3 * r + p] * A[3 * r + p];
          be += A[3 * r + q] * A[3 * r + q];
          ga += A[3 * r + p] * A[3 * r + q];
        }
        if (fabs(ga) <= 1e-16 * sqrt(al * be) || ga == 0.0) continue;
        off += fabs(ga);
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < 3; ++r) {
          const double ap = A[3 * r + p], aq = A[3 * r + q];
          A[3 * r + p] = c * ap - s * aq;
          A[3 * r + q] = s * ap + c * aq;
          const double vp = V[3 * r + p], vq = V[3 * r + q];
          V[3 * r + p] = c * vp - s * vq;
          V[3 * r + q] = s * vp + c * vq;
        }
      }
    if (off == 0.0) break;
  }
}

__device__ __forceinline__ double det3(const double M[9]) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// Umeyama similarity from weighted moments: mu_s, mu_r, var_s = E|s - mu_s|^2, Sigma = E (r - mu_r)(s - mu_s)^T (row-major).
// Writes T (3x4 row-major: c R | t).  Returns false for a degenerate sample (zero variance / rank < 2).
__device__ bool umeyama_from_moments(const double mu_s[3], const double mu_r[3], double var_s, const double Sigma[9], float T[12]) {
  if (!(var_s > 1e-20)) return false;
  double A[9], V[9];
  for (int i = 0; i < 9; ++i) A[i] = Sigma[i];
  jacobi_svd3(A, V);
  double sg[3];
  for (int j = 0; j < 3; ++j) sg[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  // sort the singular values descending (column permutation of U and V)
  int ord[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2 - a; ++b)
      if (sg[ord[b]] < sg[ord[b + 1]]) { const int t = ord[b]; ord[b] = ord[b + 1]; ord[b + 1] = t; }
  if (!(sg[ord[1]] > 1e-12 * sg[ord[0]]) || !(sg[ord[0]] > 0.0)) return false;  // (near) collinear sample
  double U[9], Vs[9], D[3];
  for (int j = 0; j < 3; ++j) {
    const int o = ord[j];
    D[j] = sg[o];
    for (int r = 0; r < 3; ++r) { Vs[3 * r + j] = V[3 * r + o]; U[3 * r + j] = sg[o] > 0.0 ? A[3 * r + o] / sg[o] : 0.0; }
  }
  if (!(D[2] > 1e-12 * D[0])) {
    // planar sample: third left vector = u0 x u1 (completes an orthonormal basis; its sign is fixed by S below)
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
  const double sdet = det3(U) * det3(Vs) < 0.0 ? -1.0 : 1.0;
  const double S[3] = {1.0, 1.0, sdet};
  const double c = (D[0] * S[0] + D[1] * S[1] + D[2] * S[2]) / var_s;
  double R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double a = 0.0;
      for (int k = 0; k < 3; ++k) a += U[3 * i + k] * S[k] * Vs[3 * j + k];
      R[3 * i + j] = a;
    }
  for (int i = 0; i < 3; ++i) {
    double t = mu_r[i];
    for (int j = 0; j < 3; ++j) { T[4 * i + j] = (float)(c * R[3 * i + j]); t -= c * R[3 * i + j] * mu_s[j]; }
    T[4 * i + 3] = (float)t;
  }
  return isfinite(c);
}

// keys[h] = (inliers << 32) | ~bits(rmse): larger is better (more inliers, then lower rmse); transforms[h] = 3x4
__global__ void __launch_bounds__(kRansacThreads) ransac_hypotheses_kernel(const float* __restrict__ ref, const float* __restrict__ src,
                                                                          const int* __restrict__ num_dev, int cap, int n_hyp,
                                                                          int sample_n, float thr, unsigned long long seed,
                                                                          unsigned long long* __restrict__ keys, float* __restrict__ Ts) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[kRansacTile * 6];
  int n = num_dev ? *num_dev : cap;
  n = n < 0 ? 0 : (n > cap ? cap : n);
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  float T[12];
  bool ok = false;
  if (h < n_hyp && n >= 3) {
    double ms[3] = {0, 0, 0}, mr[3] = {0, 0, 0};
    float ps[kRansacMaxSample][3], pr[kRansacMaxSample][3];
    for (int j = 0; j < sample_n; ++j) {
      const unsigned long long r = splitmix64(seed ^ splitmix64(((unsigned long long)h << 8) | (unsigned long long)j));
      const int i = (int)(r % (unsigned long long)n);
      for (int d = 0; d < 3; ++d) {
        ps[j][d] = src[3 * i + d]; pr[j][d] = ref[3 * i + d];
        ms[d] += ps[j][d]; mr[d] += pr[j][d];
      }
    }
    for (int d = 0; d < 3; ++d) { ms[d] /= sample_n; mr[d] /= sample_n; }
    double var = 0.0, Sg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < sample_n; ++j) {
      double ds[3], dr[3];
      for (int d = 0; d < 3; ++d) { ds[d] = ps[j][d] - ms[d]; dr[d] = pr[j][d] - mr[d]; var += ds[d] * ds[d]; }
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Sg[3 * a + b] += dr[a] * ds[b];
    }
    var /= sample_n;
    for (int i = 0; i < 9; ++i) Sg[i] /= sample_n;
    ok = umeyama_from_moments(ms, mr, var, Sg, T);
  }
  // score on all correspondences (tiles staged in shared memory by the whole block)
  int inl = 0;
  float sq = 0.f;
  const float thr2 = thr * thr;
  for (int t0 = 0; t0 < n; t0 += kRansacTile) {
    const int m = min(kRansacTile, n - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < m * 3; i += blockDim.x) { sh[i] = src[3 * t0 + i]; sh[kRansacTile * 3 + i] = ref[3 * t0 + i]; }
    __syncthreads();
    if (ok) {
      for (int i = 0; i < m; ++i) {
        const float x = sh[3 * i], y = sh[3 * i + 1], z = sh[3 * i + 2];
        const float dx = T[0] * x + T[1] * y + T[2] * z + T[3] - sh[kRansacTile * 3 + 3 * i];
        const float dy = T[4] * x + T[5] * y + T[6] * z + T[7] - sh[kRansacTile * 3 + 3 * i + 1];
        const float dz = T[8] * x + T[9] * y + T[10] * z + T[11] - sh[kRansacTile * 3 + 3 * i + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < thr2) { ++inl; sq += d2; }
      }
    }
  }
  if (h < n_hyp) {
    unsigned long long key = 0ull;
    if (ok && inl > 0) {
      const float rmse = sqrtf(sq / (float)inl);
      key = ((unsigned long long)(unsigned)inl << 32) | (unsigned long long)(~__float_as_uint(rmse));
    }
    keys[h] = key;
    for (int i = 0; i < 12; ++i) Ts[12 * (long long)h + i] = ok ? T[i] : 0.f;
  }
}

// best hypothesis (ties: lowest index), optional re-fit on its inlier set; one block
__global__ void __launch_bounds__(256) ransac_select_kernel(const float* __restrict__ ref, const float* __restrict__ src,
                                                            const int* __restrict__ num_dev, int cap, int n_hyp, float thr, int refit,
                                                            const unsigned long long* __restrict__ keys, const float* __restrict__ Ts,
                                                            const float* __restrict__ fallback, float* __restrict__ T_out,
                                                            int* __restrict__ info) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned long long s_key[256];
  __shared__ int s_idx[256];
  __shared__ double s_red[256];
  __shared__ float s_T[12];
  int n = num_dev ? *num_dev : cap;
  n = n < 0 ? 0 : (n > cap ? cap : n);
  const int tid = threadIdx.x;
  unsigned long long bk = 0ull;
  int bi = 0x7fffffff;
  for (int h = tid; h < n_hyp; h += 256) {
    const unsigned long long k = keys[h];
    if (k > bk) { bk = k; bi = h; }  // ascending h per thread: the first maximum wins
  }
  s_key[tid] = bk; s_idx[tid] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const unsigned long long k2 = s_key[tid + o];
      const int i2 = s_idx[tid + o];
      if (k2 > s_key[tid] || (k2 == s_key[tid] && i2 < s_idx[tid])) { s_key[tid] = k2; s_idx[tid] = i2; }
    }
    __syncthreads();
  }
  const unsigned long long best_key = s_key[0];
  const int best = s_idx[0];
  const bool found = best_key != 0ull;
  if (tid < 12) s_T[tid] = found ? Ts[12 * (long long)best + tid] : (fallback ? fallback[tid] : ((tid % 5 == 0) ? 1.f : 0.f));
  __syncthreads();
  if (found && refit) {
    // moments over the inliers of the best hypothesis: 1 + 3 + 3 + 1 + 9 sums, block tree in double (fixed order)
    float T[12];
    for (int i = 0; i < 12; ++i) T[i] = s_T[i];
    const float thr2 = thr * thr;
    auto block_sum = [&](double v) {
      __syncthreads();
      s_red[tid] = v;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) { if (tid < o) s_red[tid] += s_red[tid + o]; __syncthreads(); }
      return s_red[0];
    };
    auto is_inlier = [&](int i) {
      const float x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
      const float dx = T[0] * x + T[1] * y + T[2] * z + T[3] - ref[3 * i];
      const float dy = T[4] * x + T[5] * y + T[6] * z + T[7] - ref[3 * i + 1];
      const float dz = T[8] * x + T[9] * y + T[10] * z + T[11] - ref[3 * i + 2];
      return dx * dx + dy * dy + dz * dz < thr2;
    };
    double cnt = 0, a[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256)
      if (is_inlier(i)) { cnt += 1.0; for (int d = 0; d < 3; ++d) { a[d] += src[3 * i + d]; a[3 + d] += ref[3 * i + d]; } }
    const double N = block_sum(cnt);
    double mu[6];
    for (int d = 0; d < 6; ++d) mu[d] = block_sum(a[d]) / (N > 0 ? N : 1.0);
    double var = 0, Sg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256)
      if (is_inlier(i)) {
        double ds[3], dr[3];
        for (int d = 0; d < 3; ++d) { ds[d] = src[3 * i + d] - mu[d]; dr[d] = ref[3 * i + d] - mu[3 + d]; var += ds[d] * ds[d]; }
        for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) Sg[3 * p + q] += dr[p] * ds[q];
      }
    var = block_sum(var) / (N > 0 ? N : 1.0);
    for (int i = 0; i < 9; ++i) Sg[i] = block_sum(Sg[i]) / (N > 0 ? N : 1.0);
    if (tid == 0 && N >= 3.0) {
      float Tn[12];
      if (umeyama_from_moments(mu, mu + 3, var, Sg, Tn))
        for (int i = 0; i < 12; ++i) s_T[i] = Tn[i];
    }
    __syncthreads();
  }
  if (tid < 12) T_out[tid] = s_T[tid];
  if (tid >= 12 && tid < 16) T_out[tid] = tid == 15 ? 1.f : 0.f;
  if (tid == 0 && info) { info[0] = found ? (int)(best_key >> 32) : 0; info[1] = found ? best : -1; }
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_similarity_ransac_workspace_size(int num_hypotheses) {
  return (size_t)num_hypotheses * (sizeof(unsigned long long) + 12 * sizeof(float)) + 512;
}

/* N3 (model.py:209-215, utils/open3d.py:169-198).  ref_corr / src_corr (capacity,3) f32; d_num_corr: device int32 count (NULL: all
 * `capacity` rows are valid).  fallback: device (4,4) transform returned when no hypothesis has an inlier (e.g. fewer than 3
 * correspondences), NULL = identity.  T_out (4,4): similarity transform [c R | t; 0 0 0 1]; info[2] (may be NULL) =
 * {inliers of the best hypothesis, its index}. */
extern "C" int gr_similarity_ransac(const float* ref_corr, const float* src_corr, const int32_t* d_num_corr, int capacity,
                                    int num_hypotheses, int sample_size, float distance_threshold, uint64_t seed, int refit,
                                    const float* fallback, float* T_out, int32_t* info, void* ws, size_t ws_bytes, void* stream) {
  if (capacity < 0 || num_hypotheses <= 0 || sample_size < 3 || sample_size > kRansacMaxSample || !(distance_threshold > 0.f))
    return GR_ERR_BAD_ARG;
  if (!ref_corr || !src_corr || !T_out) return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < gr_similarity_ransac_workspace_size(num_hypotheses)) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* keys = static_cast<unsigned long long*>(ws);
  float* Ts = reinterpret_cast<float*>(keys + num_hypotheses);
  GR_CHECK_CUDA(launch_pdl(ransac_hypotheses_kernel, dim3(ceil_div(num_hypotheses, kRansacThreads)), dim3(kRansacThreads), (size_t)(0), st, ref_corr, src_corr, d_num_corr, capacity, num_hypotheses, sample_size, distance_threshold, (unsigned long long)seed, keys, Ts));
  GR_CHECK_LAUNCH("ransac_hypotheses_kernel");
  GR_CHECK_CUDA(launch_pdl(ransac_select_kernel, dim3(1), dim3(256), (size_t)(0), st, ref_corr, src_corr, d_num_corr, capacity, num_hypotheses, distance_threshold, refit, keys, Ts,
                                          fallback, T_out, info));
  GR_CHECK_LAUNCH("ransac_select_kernel");
  return GR_OK;
}
