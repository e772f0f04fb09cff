// L1/L2: local-to-global registration, entirely on the device
// (geotransformer/modules/geotransformer/local_global_registration.py:11-235,
//  geotransformer/modules/registration/procrustes.py:6-82).
//
// The reference does six 3x3 SVDs on the CPU (procrustes.py:59) and one Python-list sync
// (local_global_registration.py:159); here the weighted Kabsch solve is a one-sided Jacobi SVD in
// double precision run by one thread, and the whole refinement chain is a single CTA.
#include <cooperative_groups.h>

#include "common.cuh"

namespace gr {

constexpr int kLgrMaxPerRow = 8;   // >= topk
constexpr int kLgrThreads = 128;

// ---- 3x3 SVD based weighted Kabsch --------------------------------------------------------------
// H = sum_i w_i (s_i - sc)(r_i - rc)^T ;  R = V diag(1,1,det(V U^T)) U^T ;  t = rc - R sc
// __noinline__ on purpose: inlined into the single-thread branch of block_procrustes, nvcc 12.9 -O3 produced
// wrong rotations for 128-thread CTAs (verified on B200; the same source is correct on the host and out of line).
__device__ __noinline__ void kabsch_from_H(const double H[9], const double sc[3], const double rc[3], float T[12]) {
  double A[9], V[9];
  for (int i = 0; i < 9; ++i) { A[i] = H[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;  // a sweep that rotates nothing leaves A and V as they are: so would every later one
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int r = 0; r < 3; ++r) { al += A[3 * r + p] * A[3 * r + p]; be += A[3 * r + q] * A[3 * r + q]; ga += A[3 * r + p] * A[3 * r + q]; }
        if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
        rotated = true;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < 3; ++r) {
          const double ap = A[3 * r + p], aq = A[3 * r + q];
          A[3 * r + p] = c * ap - s * aq; A[3 * r + q] = s * ap + c * aq;
          const double vp = V[3 * r + p], vq = V[3 * r + q];
          V[3 * r + p] = c * vp - s * vq; V[3 * r + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sg[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) sg[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2 - a; ++b)
      if (sg[ord[b]] < sg[ord[b + 1]]) { const int t = ord[b]; ord[b] = ord[b + 1]; ord[b + 1] = t; }
  double U[9], Vs[9];
  for (int j = 0; j < 3; ++j) {
    const int o = ord[j];
    for (int r = 0; r < 3; ++r) { Vs[3 * r + j] = V[3 * r + o]; U[3 * r + j] = sg[o] > 0 ? A[3 * r + o] / sg[o] : 0.0; }
  }
  const double s0 = sg[ord[0]], s1 = sg[ord[1]], s2 = sg[ord[2]];
  const double tiny = 1e-14 * (s0 > 0 ? s0 : 1.0);
  if (s0 <= 0.0) {  // H == 0: identity rotation
    for (int i = 0; i < 9; ++i) { U[i] = (i % 4 == 0) ? 1.0 : 0.0; Vs[i] = U[i]; }
  } else {
    if (s1 <= tiny) {  // rank 1: any unit vector orthogonal to u0
      const double ax = fabs(U[0]), ay = fabs(U[3]), az = fabs(U[6]);
      double e[3] = {0, 0, 0};
      e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
      const double d = e[0] * U[0] + e[1] * U[3] + e[2] * U[6];
      double w[3] = {e[0] - d * U[0], e[1] - d * U[3], e[2] - d * U[6]};
      const double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
      U[1] = w[0] / nw; U[4] = w[1] / nw; U[7] = w[2] / nw;
    }
    if (s2 <= tiny) {  // rank <= 2: u2 = u0 x u1 (sign absorbed by the determinant correction)
      U[2] = U[3] * U[7] - U[6] * U[4];
      U[5] = U[6] * U[1] - U[0] * U[7];
      U[8] = U[0] * U[4] - U[3] * U[1];
    }
  }
  // M = V U^T
  double M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[3 * i + j] = Vs[3 * i] * U[3 * j] + Vs[3 * i + 1] * U[3 * j + 1] + Vs[3 * i + 2] * U[3 * j + 2];
  const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
  const double d = det > 0 ? 1.0 : (det < 0 ? -1.0 : 0.0);
  double R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = Vs[3 * i] * U[3 * j] + Vs[3 * i + 1] * U[3 * j + 1] + d * Vs[3 * i + 2] * U[3 * j + 2];
  for (int i = 0; i < 3; ++i) {
    T[4 * i] = (float)R[3 * i]; T[4 * i + 1] = (float)R[3 * i + 1]; T[4 * i + 2] = (float)R[3 * i + 2];
    T[4 * i + 3] = (float)(rc[i] - (R[3 * i] * sc[0] + R[3 * i + 1] * sc[1] + R[3 * i + 2] * sc[2]));
  }
}

constexpr int kLgrShFloats = 32 * 9 + 16;  // scratch of block_sum_n: 32 warp partials per value, then the totals

// Sums N per-thread values over the CTA (every thread receives all N totals): one shuffle butterfly per value, the warp
// partials through shared memory, a second butterfly by warp 0.  Three barriers whatever N is.
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float* sh /* kLgrShFloats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sh[k * 32 + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float t = lane < nw ? sh[k * 32 + lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) sh[32 * N + k] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = sh[32 * N + k];
}

// Sum over one CTA / over the CTAs of a thread-block cluster (fixed rank order, every thread of every CTA receives the
// same totals).  The cluster form exchanges the per-CTA totals through distributed shared memory, double-buffered so
// that one cluster barrier per reduction suffices (a buffer is rewritten two reductions later, and a CTA only gets
// there through the barrier of the reduction in between, which its peers reach after their reads).
struct BlockReducer {
  float* sh;
  template <int N>
  __device__ __forceinline__ void sum(float (&v)[N]) { block_sum_n<N>(v, sh); }
};
constexpr int kRefineCtas = 8;
struct ClusterReducer {
  float* sh;
  float* xchg;  // [2][16] in every CTA's shared memory
  int parity;
  template <int N>
  __device__ __forceinline__ void sum(float (&v)[N]) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    block_sum_n<N>(v, sh);
    float* mine = xchg + parity * 16;
    if (threadIdx.x < N) mine[threadIdx.x] = v[threadIdx.x < N ? threadIdx.x : 0];
    cluster.sync();
    if (threadIdx.x < N) {
      float t = 0.f;
      for (int r = 0; r < kRefineCtas; ++r) t += cluster.map_shared_rank(mine, r)[threadIdx.x];
      sh[threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = sh[k];
    __syncthreads();
    parity ^= 1;
  }
};

// A run of correspondences; the pointers may address shared or global memory.
struct CorrSeg {
  const float* src;
  const float* ref;
  const float* w;
  int n;
};

// Weighted Procrustes over the correspondences of segments a then b, executed by a whole CTA.  Weights w_i >= 0.
// procrustes.py:41-70: w <- w / (sum w + eps); centroids; H; SVD.  Result T (3x4 row-major) in shared memory.
//
// All sums are accumulated in fp32, as the reference does (torch.sum / bmm on float tensors).  This matters for
// DEGENERATE patches (H numerically zero or rank one: three nearly coincident correspondences, vanishing weights):
// there the reference's rotation is whatever the fp32 rounding noise in H dictates.  Accumulating H exactly (fp64)
// would instead return a clean near-identity rotation for H ~ 0, and such a hypothesis can collect spuriously many
// inliers whenever the true motion is small -- a systematic deviation from the reference's hypothesis selection
// (observed on the textured3k golden).  Only the 3x3 factorisation itself runs in fp64.
template <typename Reducer>
__device__ void block_procrustes(const CorrSeg a, const CorrSeg b, float eps, Reducer& red, float* T_out /* shared, 12 */) {
  const CorrSeg segs[2] = {a, b};
  float sw[1] = {0.f};
#pragma unroll
  for (int g = 0; g < 2; ++g)
    for (int i = threadIdx.x; i < segs[g].n; i += blockDim.x) sw[0] += segs[g].w[i];
  red.template sum<1>(sw);
  const float denom = sw[0] + eps;
  float cen[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const float* __restrict__ src = segs[g].src;
    const float* __restrict__ ref = segs[g].ref;
    const float* __restrict__ w = segs[g].w;
    for (int i = threadIdx.x; i < segs[g].n; i += blockDim.x) {
      const float wi = w[i] / denom;
      cen[0] += src[3 * i] * wi; cen[1] += src[3 * i + 1] * wi; cen[2] += src[3 * i + 2] * wi;
      cen[3] += ref[3 * i] * wi; cen[4] += ref[3 * i + 1] * wi; cen[5] += ref[3 * i + 2] * wi;
    }
  }
  red.template sum<6>(cen);
  const float scx = cen[0], scy = cen[1], scz = cen[2];
  const float rcx = cen[3], rcy = cen[4], rcz = cen[5];
  float h[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const float* __restrict__ src = segs[g].src;
    const float* __restrict__ ref = segs[g].ref;
    const float* __restrict__ w = segs[g].w;
    for (int i = threadIdx.x; i < segs[g].n; i += blockDim.x) {
      const float wi = w[i] / denom;
      const float sx = src[3 * i] - scx, sy = src[3 * i + 1] - scy, sz = src[3 * i + 2] - scz;
      const float rx = wi * (ref[3 * i] - rcx), ry = wi * (ref[3 * i + 1] - rcy), rz = wi * (ref[3 * i + 2] - rcz);
      h[0] += sx * rx; h[1] += sx * ry; h[2] += sx * rz;
      h[3] += sy * rx; h[4] += sy * ry; h[5] += sy * rz;
      h[6] += sz * rx; h[7] += sz * ry; h[8] += sz * rz;
    }
  }
  red.template sum<9>(h);
  if (threadIdx.x == 0) {
    double H[9];
    for (int k = 0; k < 9; ++k) H[k] = (double)h[k];
    const double sc[3] = {(double)scx, (double)scy, (double)scz}, rc[3] = {(double)rcx, (double)rcy, (double)rcz};
    kabsch_from_H(H, sc, rc, T_out);
  }
  __syncthreads();
}

// ---- correspondence extraction ------------------------------------------------------------------
struct PatchCorr {  // per patch: up to K*topk candidates in (row, col) order
  int count;
};

// One CTA per patch.  score = exp(log score) on the K x K block without dustbin.  Threads 0..K-1 find the top-k of
// their row and threads 128..128+K-1 the top-k of their column, each in ONE scan with a sorted insertion (value
// descending, index ascending on ties: what repeated torch.topk / argmax yields).  Mutual selection then only has to
// test the <= topk row winners of a row against the column winners (local_global_registration.py:71-93).
constexpr int kCorrThreads = 256;
constexpr int kCorrTop = kLgrMaxPerRow / 2;  // >= topk
__global__ void __launch_bounds__(kCorrThreads) lgr_correspondence_kernel(
    const float* __restrict__ ms, int K, int ld /* K or K+1 */, const unsigned char* __restrict__ ref_masks,
    const unsigned char* __restrict__ src_masks, int topk, float conf, int* __restrict__ counts,
    int* __restrict__ cand_rc /* (P, K*topk) packed row<<16|col */, float* __restrict__ cand_score) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float e[];  // K x (K+1) exp scores (padded rows: conflict-free row and column walks)
  const int KP = K + 1;
  __shared__ int row_top[kLgrThreads][kCorrTop];
  __shared__ int col_top[kLgrThreads][kCorrTop];
  __shared__ int warp_tot[kLgrThreads / 32];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* src = ms + (long long)b * ld * ld;
  for (int i = t; i < K * K; i += blockDim.x) e[(i / K) * KP + (i % K)] = expf(src[(i / K) * ld + (i % K)]);
  __syncthreads();
  {
    const int line = t & (kLgrThreads - 1);
    const bool rows = t < kLgrThreads;
    if (line < K) {
      float bv[kCorrTop];
      int bi[kCorrTop];
#pragma unroll
      for (int q = 0; q < kCorrTop; ++q) { bv[q] = -INFINITY; bi[q] = -1; }
      const float* walk = rows ? e + line * KP : e + line;
      const int step = rows ? 1 : KP;
      for (int j = 0; j < K; ++j) {
        const float v = walk[j * step];
#pragma unroll
        for (int q = kCorrTop - 1; q >= 0; --q)
          if (v > bv[q]) {
            if (q < kCorrTop - 1) { bv[q + 1] = bv[q]; bi[q + 1] = bi[q]; }
            bv[q] = v; bi[q] = j;
          }
      }
#pragma unroll
      for (int q = 0; q < kCorrTop; ++q) (rows ? row_top : col_top)[line][q] = bi[q];
    }
  }
  __syncthreads();
  int mine[kCorrTop];
  int nm = 0;
  if (t < K && ref_masks[(long long)b * K + t]) {
    for (int r = 0; r < topk; ++r) {
      const int j = row_top[t][r];
      if (j < 0) continue;
      bool in_col = false;
      for (int q = 0; q < topk; ++q) in_col |= col_top[j][q] == t;
      if (in_col && e[t * KP + j] > conf && src_masks[(long long)b * K + j]) mine[nm++] = j;
    }
    // torch.nonzero order: ascending column within the row
    for (int x = 1; x < nm; ++x)
      for (int y = x; y > 0 && mine[y - 1] > mine[y]; --y) { const int tmp = mine[y]; mine[y] = mine[y - 1]; mine[y - 1] = tmp; }
  }
  if (t < kLgrThreads) {  // exclusive scan of the row counts over the first four warps
    const int lane = t & 31, warp = t >> 5;
    int inc = nm;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) warp_tot[warp] = inc;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (t == kLgrThreads - 1) counts[b] = before + inc;
    const long long base = (long long)b * K * topk + before + inc - nm;
    for (int q = 0; q < nm; ++q) {
      cand_rc[base + q] = (t << 16) | mine[q];
      cand_score[base + q] = e[t * KP + mine[q]];
    }
  }
}

// gathers the stacked correspondences in torch.nonzero (row-major) order; block b copies patch b
__global__ void __launch_bounds__(128) lgr_compact_kernel(const int* __restrict__ counts, int P, int cap_per_patch,
                                                          const int* __restrict__ cand_rc, const float* __restrict__ cand_score,
                                                          const float* __restrict__ ref_knn_points, const float* __restrict__ src_knn_points,
                                                          int K, float* __restrict__ ref_corr, float* __restrict__ src_corr,
                                                          float* __restrict__ corr_scores, int* __restrict__ offsets, int* __restrict__ total) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_off;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < b; ++i) acc += counts[i];
    s_off = acc;
    offsets[b] = acc;
    if (b == P - 1) { offsets[P] = acc + counts[b]; *total = acc + counts[b]; }
  }
  __syncthreads();
  const int off = s_off, n = counts[b];
  for (int q = threadIdx.x; q < n; q += blockDim.x) {
    const int rc = cand_rc[(long long)b * cap_per_patch + q];
    const int i = rc >> 16, j = rc & 0xffff;
    const float* rp = ref_knn_points + ((long long)b * K + i) * 3;
    const float* sp = src_knn_points + ((long long)b * K + j) * 3;
    const long long o = (long long)(off + q);
    ref_corr[3 * o] = rp[0]; ref_corr[3 * o + 1] = rp[1]; ref_corr[3 * o + 2] = rp[2];
    src_corr[3 * o] = sp[0]; src_corr[3 * o + 1] = sp[1]; src_corr[3 * o + 2] = sp[2];
    corr_scores[o] = cand_score[(long long)b * cap_per_patch + q];
  }
}

// local hypotheses: one CTA per patch; patches with fewer than `thr` correspondences are skipped.
// Then the hypothesis is scored on ALL correspondences.  inliers[b] = -1 for skipped patches.
__global__ void __launch_bounds__(kLgrThreads) lgr_hypothesis_kernel(const float* __restrict__ ref_corr, const float* __restrict__ src_corr,
                                                                     const float* __restrict__ corr_scores, const int* __restrict__ offsets,
                                                                     int P, int thr, float radius, float eps, float* __restrict__ T_local,
                                                                     int* __restrict__ inliers) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[kLgrShFloats];
  __shared__ float T[12];
  __shared__ int s_cnt;
  const int b = blockIdx.x;
  const int off = offsets[b], n = offsets[b + 1] - off, C = offsets[P];
  if (n < thr) {
    if (threadIdx.x == 0) inliers[b] = -1;
    return;
  }
  BlockReducer red{sh};
  block_procrustes(CorrSeg{src_corr + 3ll * off, ref_corr + 3ll * off, corr_scores + off, n}, CorrSeg{nullptr, nullptr, nullptr, 0}, eps, red, T);
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float sx = src_corr[3 * i], sy = src_corr[3 * i + 1], sz = src_corr[3 * i + 2];
    // apply_transform: p R^T + t, each output = dot(row of R, p) + t
    const float ax = (sx * T[0] + sy * T[1] + sz * T[2]) + T[3];
    const float ay = (sx * T[4] + sy * T[5] + sz * T[6]) + T[7];
    const float az = (sx * T[8] + sy * T[9] + sz * T[10]) + T[11];
    const float dx = ref_corr[3 * i] - ax, dy = ref_corr[3 * i + 1] - ay, dz = ref_corr[3 * i + 2] - az;
    c += sqrtf(dx * dx + dy * dy + dz * dz) < radius ? 1 : 0;
  }
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) inliers[b] = s_cnt;
  if (threadIdx.x < 12) T_local[b * 12 + threadIdx.x] = T[threadIdx.x];
}

// ONE thread-block cluster of kRefineCtas CTAs: pick the best hypothesis (first maximum), then `steps` rounds of global
// refinement.  CTA r owns the r-th contiguous slice of the correspondences and stages up to `cap` of them in its shared
// memory once (8 x 6400 covers every case seen; whatever exceeds it is read in place), so the 4 * steps passes cost
// shared-memory latency; the three sums of each weighted Procrustes are exchanged through distributed shared memory
// (ClusterReducer), every CTA then solves the same 3x3 problem redundantly and continues with the same transform.
// (One CTA alone spent 0.21 ms here: its slice of the L2 bandwidth on ~25 k correspondences per pass.)
constexpr int kRefineThreads = 1024;
constexpr int kRefineCap = 6400;  // x 32 B = 200 KB of dynamic shared memory
__global__ void __cluster_dims__(kRefineCtas, 1, 1) __launch_bounds__(kRefineThreads)
    lgr_refine_kernel(const float* __restrict__ ref_corr, const float* __restrict__ src_corr, const float* __restrict__ corr_scores,
                      const int* __restrict__ offsets, int P, const float* __restrict__ T_local, const int* __restrict__ inliers,
                      float radius, float eps, int steps, int cap, float* __restrict__ w /* scratch, C */,
                      float* __restrict__ T_out /* 16 */, int* __restrict__ best_out) {
  pdl_wait();
  pdl_trigger();
  namespace cg = cooperative_groups;
  const int rank = (int)cg::this_cluster().block_rank();
  extern __shared__ float stage[];  // src (3 cap) | ref (3 cap) | score (cap) | w (cap)
  __shared__ float sh[kLgrShFloats];
  __shared__ float xchg[2 * 16];
  __shared__ float T[12];
  __shared__ int s_best;
  const int C = offsets[P];
  const int per = (C + kRefineCtas - 1) / kRefineCtas;
  const int lo = rank * per < C ? rank * per : C;
  const int n = C - lo < per ? C - lo : per;
  const int n0 = n < cap ? n : cap, n1 = n - n0;
  const float* g_src = src_corr + 3ll * lo;
  const float* g_ref = ref_corr + 3ll * lo;
  const float* g_score = corr_scores + lo;
  float* g_w = w + lo;
  float* s_src = stage;
  float* s_ref = s_src + 3 * (size_t)cap;
  float* s_score = s_ref + 3 * (size_t)cap;
  float* s_w = s_score + cap;
  for (int i = threadIdx.x; i < 3 * n0; i += blockDim.x) { s_src[i] = g_src[i]; s_ref[i] = g_ref[i]; }
  for (int i = threadIdx.x; i < n0; i += blockDim.x) s_score[i] = g_score[i];
  if (threadIdx.x == 0) {
    int best = -1, bc = -1;
    for (int b = 0; b < P; ++b)
      if (inliers[b] > bc) { bc = inliers[b]; best = b; }
    s_best = best;
    if (rank == 0) *best_out = best;
  }
  __syncthreads();
  ClusterReducer red{sh, xchg, 0};
  const CorrSeg near_scores{s_src, s_ref, s_score, n0}, far_scores{g_src + 3ll * n0, g_ref + 3ll * n0, g_score + n0, n1};
  const CorrSeg near_w{s_src, s_ref, s_w, n0}, far_w{g_src + 3ll * n0, g_ref + 3ll * n0, g_w + n0, n1};
  auto rescore = [&]() {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const CorrSeg& sg = g == 0 ? near_scores : far_scores;
      float* wout = g == 0 ? s_w : g_w + n0;
      for (int i = threadIdx.x; i < sg.n; i += blockDim.x) {
        const float sx = sg.src[3 * i], sy = sg.src[3 * i + 1], sz = sg.src[3 * i + 2];
        const float ax = (sx * T[0] + sy * T[1] + sz * T[2]) + T[3];
        const float ay = (sx * T[4] + sy * T[5] + sz * T[6]) + T[7];
        const float az = (sx * T[8] + sy * T[9] + sz * T[10]) + T[11];
        const float dx = sg.ref[3 * i] - ax, dy = sg.ref[3 * i + 1] - ay, dz = sg.ref[3 * i + 2] - az;
        wout[i] = sqrtf(dx * dx + dy * dy + dz * dz) < radius ? sg.w[i] : 0.f;
      }
    }
    __syncthreads();
  };
  if (s_best >= 0) {
    if (threadIdx.x < 12) T[threadIdx.x] = T_local[s_best * 12 + threadIdx.x];
    __syncthreads();
    rescore();
  } else {
    // degenerate: initialise with all correspondences (local_global_registration.py:180-185)
    block_procrustes(near_scores, far_scores, eps, red, T);
    rescore();
  }
  block_procrustes(near_w, far_w, eps, red, T);
  for (int s = 0; s < steps - 1; ++s) {
    rescore();
    block_procrustes(near_w, far_w, eps, red, T);
  }
  if (rank == 0) {
    if (threadIdx.x < 12) T_out[threadIdx.x] = T[threadIdx.x];
    if (threadIdx.x >= 12 && threadIdx.x < 16) T_out[threadIdx.x] = threadIdx.x == 15 ? 1.f : 0.f;
  }
  cg::this_cluster().sync();  // no CTA may exit while a peer can still read its exchange buffer
}

// batched weighted Procrustes: one CTA per problem (drop-in for WeightedProcrustes.forward)
__global__ void __launch_bounds__(kLgrThreads) procrustes_batched_kernel(const float* __restrict__ src, const float* __restrict__ ref,
                                                                         const float* __restrict__ w, int n, float eps,
                                                                         float* __restrict__ T_out /* (B,4,4) */) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[kLgrShFloats];
  __shared__ float T[12];
  const long long b = blockIdx.x;
  BlockReducer red{sh};
  block_procrustes(CorrSeg{src + b * n * 3, ref + b * n * 3, w + b * n, n}, CorrSeg{nullptr, nullptr, nullptr, 0}, eps, red, T);
  if (threadIdx.x < 12) T_out[b * 16 + threadIdx.x] = T[threadIdx.x];
  if (threadIdx.x >= 12 && threadIdx.x < 16) T_out[b * 16 + threadIdx.x] = threadIdx.x == 15 ? 1.f : 0.f;
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_lgr_workspace_size(int P, int K, int topk) {
  Carver c(nullptr, 0);
  const size_t cap = (size_t)P * K * topk * 2;
  c.take<int>(P + 1);      // counts
  c.take<int>(P + 2);      // offsets
  c.take<int>(cap);        // cand_rc
  c.take<float>(cap);      // cand_score
  c.take<float>(12 * (size_t)P);
  c.take<int>(P);
  c.take<float>(cap);      // refinement weights
  return c.off;
}

/* L1.  matching_scores (P, ld, ld) log-scores with ld = K or K+1 (dustbin row/col ignored);
 * knn points (P,K,3), masks (P,K) u8.  Outputs: ref/src_corr_points (cap,3), corr_scores (cap) with
 * cap = P*K*topk (x2 when not mutual), num_corr device scalar, transform (4,4). */
extern "C" int gr_local_global_registration(const float* matching_scores, int P, int K, int ld, const float* ref_knn_points,
                                            const float* src_knn_points, const uint8_t* ref_knn_masks,
                                            const uint8_t* src_knn_masks, int topk, float acceptance_radius, int mutual,
                                            float confidence_threshold, int correspondence_threshold, int num_refinement_steps,
                                            float* ref_corr_points, float* src_corr_points, float* corr_scores,
                                            int32_t* num_corr, float* transform, void* ws, size_t ws_bytes, void* stream) {
  if (P <= 0 || K <= 0 || K > kLgrThreads || (ld != K && ld != K + 1) || topk <= 0 || topk > kLgrMaxPerRow / 2 ||
      num_refinement_steps < 1 || !mutual /* only the mutual mode of config.py:119 is implemented */)
    return GR_ERR_BAD_ARG;
  if (!matching_scores || !ref_knn_points || !src_knn_points || !ref_knn_masks || !src_knn_masks || !ref_corr_points ||
      !src_corr_points || !corr_scores || !num_corr || !transform)
    return GR_ERR_BAD_ARG;
  Carver c(ws, ws_bytes);
  const int cap_pp = K * topk * (mutual ? 1 : 2);
  const size_t cap = (size_t)P * K * topk * 2;
  int* counts = c.take<int>(P + 1);
  int* offsets = c.take<int>(P + 2);
  int* cand_rc = c.take<int>(cap);
  float* cand_score = c.take<float>(cap);
  float* T_local = c.take<float>(12 * (size_t)P);
  int* inliers = c.take<int>(P);
  float* w = c.take<float>(cap);
  if (!ws || !c.ok) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)K * (K + 1) * sizeof(float);
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(lgr_correspondence_kernel), (int)smem));
  GR_CHECK_CUDA(launch_pdl(lgr_correspondence_kernel, dim3(P), dim3(kCorrThreads), (size_t)(smem), st, matching_scores, K, ld, ref_knn_masks, src_knn_masks, topk,
                                                           confidence_threshold, counts, cand_rc, cand_score));
  GR_CHECK_LAUNCH("lgr_correspondence_kernel");
  GR_CHECK_CUDA(launch_pdl(lgr_compact_kernel, dim3(P), dim3(128), (size_t)(0), st, counts, P, cap_pp, cand_rc, cand_score, ref_knn_points, src_knn_points, K, ref_corr_points,
                                        src_corr_points, corr_scores, offsets, num_corr));
  GR_CHECK_LAUNCH("lgr_compact_kernel");
  GR_CHECK_CUDA(launch_pdl(lgr_hypothesis_kernel, dim3(P), dim3(kLgrThreads), (size_t)(0), st, ref_corr_points, src_corr_points, corr_scores, offsets, P, correspondence_threshold,
                                                   acceptance_radius, 1e-5f, T_local, inliers));
  GR_CHECK_LAUNCH("lgr_hypothesis_kernel");
  const size_t refine_smem = (size_t)kRefineCap * 8 * sizeof(float);
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(lgr_refine_kernel), (int)refine_smem));
  GR_CHECK_CUDA(launch_pdl(lgr_refine_kernel, dim3(kRefineCtas), dim3(kRefineThreads), refine_smem, st, ref_corr_points, src_corr_points, corr_scores, offsets, P, T_local, inliers,
                           acceptance_radius, 1e-5f, num_refinement_steps, kRefineCap, w, transform, counts + P));
  GR_CHECK_LAUNCH("lgr_refine_kernel");
  return GR_OK;
}

/* L2.  src/ref (B,n,3), weights (B,n) -> transforms (B,4,4). */
extern "C" int gr_weighted_procrustes(const float* src_points, const float* ref_points, const float* weights, int B, int n, float eps,
                                      float* transforms, void* stream) {
  if (B < 0 || n <= 0) return GR_ERR_BAD_ARG;
  if (B == 0) return GR_OK;
  if (!src_points || !ref_points || !weights || !transforms) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(procrustes_batched_kernel, dim3(B), dim3(kLgrThreads), (size_t)(0), static_cast<cudaStream_t>(stream), src_points, ref_points, weights, n, eps, transforms));
  GR_CHECK_LAUNCH("procrustes_batched_kernel");
  return GR_OK;
}
