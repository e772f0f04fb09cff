// T1: geometric structure embedding (geotransformer/modules/geotransformer/geotransformer.py:9-72).
//
//   gr_embedding_indices : pairwise distances, 3 nearest neighbours (self excluded), triplet angles
//                          -> d_idx (N,N) = dist / sigma_d ; a_idx (N,N,k) = atan2(|r x a|, r.a) * 180/(sigma_a*pi)
//   gr_sinusoid_rows     : E[r, 2i] = sin(x_r * div_i), E[r, 2i+1] = cos(x_r * div_i)   (positional_embedding.py:19-35)
//   gr_embedding_combine : emb[r, c] = D[r, c] + max_k A[r*k + k', c]                     (geotransformer.py:65-70)
// The two 256x256 projections in between run in the GEMM.
#include "common.cuh"

namespace gr {

__device__ __forceinline__ float sq_norm3f(float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}

// one CTA per anchor n: row of distances into shared memory, k+1 smallest by (d, index), drop the first
__global__ void __launch_bounds__(128) pairdist_knn_kernel(const float* __restrict__ pts, int N, float sigma_d, int k,
                                                           float* __restrict__ d_idx, int* __restrict__ knn) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float row[];  // N distances
  __shared__ unsigned long long sh_best[4];
  __shared__ unsigned long long sh_prev;
  const int n = blockIdx.x;
  const float x0 = pts[3 * n], x1 = pts[3 * n + 1], x2 = pts[3 * n + 2];
  const float xx = sq_norm3f(x0, x1, x2);
  for (int m = threadIdx.x; m < N; m += blockDim.x) {
    const float y0 = pts[3 * m], y1 = pts[3 * m + 1], y2 = pts[3 * m + 2];
    const float xy = fmaf(x2, y2, fmaf(x1, y1, __fmul_rn(x0, y0)));
    const float sq = fmaxf(__fadd_rn(__fsub_rn(xx, __fmul_rn(2.0f, xy)), sq_norm3f(y0, y1, y2)), 0.0f);
    const float d = sqrtf(sq);
    row[m] = d;
    d_idx[(long long)n * N + m] = __fdiv_rn(d, sigma_d);
  }
  __syncthreads();
  // selection: k+1 rounds of "smallest key greater than the previous one"
  unsigned long long prev = 0ull;
  for (int r = 0; r <= k; ++r) {
    unsigned long long best = ~0ull;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(row[m]) << 32) | (unsigned int)m;
      if ((r == 0 || key > prev) && key < best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) sh_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = sh_best[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) b = sh_best[w] < b ? sh_best[w] : b;
      sh_prev = b;
      if (r > 0) knn[n * k + (r - 1)] = b == ~0ull ? n : (int)(b & 0xffffffffull);
    }
    __syncthreads();
    prev = sh_prev;
  }
}

// a_idx[n, m, j] for j < k
__global__ void __launch_bounds__(256) angle_index_kernel(const float* __restrict__ pts, int N, int k, const int* __restrict__ knn,
                                                          float factor_a, float* __restrict__ a_idx) {
  pdl_wait();
  pdl_trigger();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * N * k;
  if (t >= total) return;
  const int j = (int)(t % k);
  const int m = (int)((t / k) % N);
  const int n = (int)(t / ((long long)k * N));
  const float px = pts[3 * n], py = pts[3 * n + 1], pz = pts[3 * n + 2];
  const int q = knn[n * k + j];
  const float rx = __fsub_rn(pts[3 * q], px), ry = __fsub_rn(pts[3 * q + 1], py), rz = __fsub_rn(pts[3 * q + 2], pz);
  const float ax = __fsub_rn(pts[3 * m], px), ay = __fsub_rn(pts[3 * m + 1], py), az = __fsub_rn(pts[3 * m + 2], pz);
  // torch.cross / linalg.norm / sum(ref * anc): separately rounded products
  const float cx = __fsub_rn(__fmul_rn(ry, az), __fmul_rn(rz, ay));
  const float cy = __fsub_rn(__fmul_rn(rz, ax), __fmul_rn(rx, az));
  const float cz = __fsub_rn(__fmul_rn(rx, ay), __fmul_rn(ry, ax));
  const float sinv = sqrtf(sq_norm3f(cx, cy, cz));
  // torch.sum starts from +0: (+0) + (-0) = +0, so a zero anchor vector gives atan2(0, +0) = 0 (not pi)
  const float cosv = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(rx, ax)), __fmul_rn(ry, ay)), __fmul_rn(rz, az));
  a_idx[t] = __fmul_rn(atan2f(sinv, cosv), factor_a);
}

// E[r, 2i] = sin(x[r] * div[i]); E[r, 2i+1] = cos(x[r] * div[i]);  C = 2 * n_div
__global__ void __launch_bounds__(256) sinusoid_rows_kernel(const float* __restrict__ x, long long rows,
                                                            const float* __restrict__ div, int n_div, float* __restrict__ E) {
  pdl_wait();
  pdl_trigger();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * n_div;
  if (t >= total) return;
  const int i = (int)(t % n_div);
  const long long r = t / n_div;
  const float om = __fmul_rn(x[r], div[i]);
  float s, c;
  sincosf(om, &s, &c);
  reinterpret_cast<float2*>(E)[t] = make_float2(s, c);
}

// out[r, c] = D[r, c] + max_j A[(r*k + j), c]
__global__ void __launch_bounds__(256) embedding_combine_kernel(const float* __restrict__ D, const float* __restrict__ A,
                                                                long long rows, int C, int k, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C) return;
  const int c = (int)(t % C);
  const long long r = t / C;
  float mx = A[(r * k) * C + c];
  for (int j = 1; j < k; ++j) mx = fmaxf(mx, A[(r * k + j) * C + c]);
  out[t] = D[t] + mx;
}

}  // namespace gr

using namespace gr;

/* T1a: d_idx (N,N) f32, a_idx (N,N,k) f32, knn (N,k) i32 scratch/output. */
extern "C" int gr_embedding_indices(const float* points, int N, float sigma_d, float sigma_a, int angle_k, float* d_idx,
                                    float* a_idx, int32_t* knn, void* stream) {
  if (N <= 0 || angle_k <= 0 || angle_k > 8 || !(sigma_d > 0.f) || !(sigma_a > 0.f)) return GR_ERR_BAD_ARG;
  if (!points || !d_idx || !a_idx || !knn) return GR_ERR_BAD_ARG;
  if ((size_t)N * sizeof(float) > 160 * 1024) return GR_ERR_CAPACITY;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)N * sizeof(float);
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(pairdist_knn_kernel), (int)smem));
  GR_CHECK_CUDA(launch_pdl(pairdist_knn_kernel, dim3(N), dim3(128), (size_t)(smem), st, points, N, sigma_d, angle_k, d_idx, knn));
  GR_CHECK_LAUNCH("pairdist_knn_kernel");
  const float factor_a = (float)(180.0 / ((double)sigma_a * 3.141592653589793));  // geotransformer.py:14
  const long long total = (long long)N * N * angle_k;
  GR_CHECK_CUDA(launch_pdl(angle_index_kernel, dim3(ceil_div(total, 256)), dim3(256), (size_t)(0), st, points, N, angle_k, knn, factor_a, a_idx));
  GR_CHECK_LAUNCH("angle_index_kernel");
  return GR_OK;
}

extern "C" int gr_sinusoid_rows(const float* x, int64_t rows, const float* div_term, int n_div, float* E, void* stream) {
  if (rows < 0 || n_div <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x || !div_term || !E) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(sinusoid_rows_kernel, dim3(ceil_div(rows * n_div, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), x, rows, div_term, n_div, E));
  GR_CHECK_LAUNCH("sinusoid_rows_kernel");
  return GR_OK;
}

extern "C" int gr_embedding_combine(const float* D, const float* A, int64_t rows, int C, int k, float* out, void* stream) {
  if (rows < 0 || C <= 0 || k <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!D || !A || !out) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(embedding_combine_kernel, dim3(ceil_div(rows * C, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), D, A, rows, C, k, out));
  GR_CHECK_LAUNCH("embedding_combine_kernel");
  return GR_OK;
}
