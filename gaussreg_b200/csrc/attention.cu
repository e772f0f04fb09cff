// T2/T3: attention score kernels.
//
// RPE self-attention (geotransformer/modules/transformer/rpe_transformer.py:50-70):
//     score[h,n,m] = ( q_h[n].k_h[m] + sum_c q_h[n,c] * (W_p e[n,m] + b_p)_{h,c} ) / sqrt(d_head)
// The reference materialises p = proj_p(e) as (1,H,N,N,64).  Here the second term is reassociated to
//     (W_p,h^T q_h[n]) . e[n,m] + q_h[n].b_p,h  =  U[h,n,:] . e[n,m,:] + qb[h,n]
// (same mathematics, different rounding: SURVEY.md section 8(a) row T2), so e is streamed once per layer
// and nothing of size N*N*C is written.  U and qb come out of two small GEMMs.
#include "common.cuh"

namespace gr {

// One CTA per query row n.  P[h, n, :] = softmax_m(score[h, n, m]);  C = H * DH, H <= 8.
template <int H>
__global__ void __launch_bounds__(256) rpe_scores_softmax_kernel(const float* __restrict__ q, const float* __restrict__ kmat,
                                                                 const float* __restrict__ U, const float* __restrict__ qb,
                                                                 const float* __restrict__ emb, int N, int C, float scale,
                                                                 float* __restrict__ P) {
  extern __shared__ float sm[];
  float* sU = sm;                 // [H][C]
  float* sq = sU + H * C;         // [C]
  float* ss = sq + C;             // [H][N] scores
  __shared__ float red[H][8];
  const int n = blockIdx.x;
  const int DH = C / H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < H * C; i += blockDim.x) sU[i] = U[((long long)(i / C) * N + n) * C + (i % C)];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sq[i] = q[(long long)n * C + i];
  __syncthreads();
  float qbh[H];
#pragma unroll
  for (int h = 0; h < H; ++h) qbh[h] = qb[(long long)h * N + n];

  // one thread per key m: the thread walks its own embedding row e[n,m,:] and key row k[m,:] sequentially
  // (float4), U / q come from shared memory as warp broadcasts -> no cross-lane reduction at all
  const float* erow = emb + (long long)n * N * C;
  const int c4_per_head = DH >> 2;
  for (int m = threadIdx.x; m < N; m += blockDim.x) {
    const float4* e4 = reinterpret_cast<const float4*>(erow + (long long)m * C);
    const float4* k4 = reinterpret_cast<const float4*>(kmat + (long long)m * C);
    float accp[H], acce[H];
#pragma unroll
    for (int h = 0; h < H; ++h) { accp[h] = 0.f; acce[h] = 0.f; }
#pragma unroll
    for (int hc = 0; hc < H; ++hc) {
      for (int i = 0; i < c4_per_head; ++i) {
        const int c4 = hc * c4_per_head + i;
        const float4 e = __ldg(e4 + c4);
        const float4 kv = __ldg(k4 + c4);
        const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * c4);
        acce[hc] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, fmaf(qv.z, kv.z, fmaf(qv.w, kv.w, acce[hc]))));
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float4 u = *reinterpret_cast<const float4*>(sU + h * C + 4 * c4);
          accp[h] = fmaf(u.x, e.x, fmaf(u.y, e.y, fmaf(u.z, e.z, fmaf(u.w, e.w, accp[h]))));
        }
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) ss[h * N + m] = (acce[h] + (accp[h] + qbh[h])) * scale;
  }
  __syncthreads();
  // softmax over m for each head
  for (int h = 0; h < H; ++h) {
    float mx = -INFINITY;
    for (int m = threadIdx.x; m < N; m += blockDim.x) mx = fmaxf(mx, ss[h * N + m]);
    mx = warp_max(mx);
    if (lane == 0) red[h][warp] = mx;
  }
  __syncthreads();
  float hmax[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float mx = red[h][0];
    for (int w = 1; w < nwarp; ++w) mx = fmaxf(mx, red[h][w]);
    hmax[h] = mx;
  }
  __syncthreads();
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
      const float e = expf(ss[h * N + m] - hmax[h]);
      ss[h * N + m] = e;
      s += e;
    }
    s = warp_sum(s);
    if (lane == 0) red[h][warp] = s;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += red[h][w];
    const float inv = 1.0f / s;
    float* out = P + ((long long)h * N + n) * N;
    for (int m = threadIdx.x; m < N; m += blockDim.x) out[m] = ss[h * N + m] * inv;
  }
}

// in-place row softmax, one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, long long rows, int cols) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float* p = x + r * cols;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, p[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { const float e = expf(p[c] - mx); p[c] = e; s += e; }
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int c = lane; c < cols; c += 32) p[c] *= inv;
}

// F.normalize(x, p=2, dim=1): x / max(|x|, eps)
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const float* __restrict__ x, long long rows, int C, float eps,
                                                                float* __restrict__ y) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = x[r * C + c]; s = fmaf(v, v, s); }
  const float nrm = fmaxf(sqrtf(warp_sum(s)), eps);
  for (int c = lane; c < C; c += 32) y[r * C + c] = x[r * C + c] / nrm;
}

}  // namespace gr

using namespace gr;

/* T2: q,k (N,C) ; U (H,N,C) ; qb (H,N) ; emb (N,N,C) -> P (H,N,N) softmax probabilities. */
extern "C" int gr_rpe_attention_probs(const float* q, const float* k, const float* U, const float* qb, const float* emb, int N,
                                      int C, int num_heads, float* P, void* stream) {
  if (N <= 0 || C <= 0 || num_heads != 4 || C % num_heads != 0) return GR_ERR_BAD_ARG;
  if (!q || !k || !U || !qb || !emb || !P) return GR_ERR_BAD_ARG;
  const size_t smem = ((size_t)num_heads * C + C + (size_t)num_heads * N) * sizeof(float);
  if (smem > 200 * 1024) return GR_ERR_CAPACITY;
  auto kern = rpe_scores_softmax_kernel<4>;
  if (smem > 48 * 1024) GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float scale = 1.0f / sqrtf((float)(C / num_heads));
  kern<<<N, 256, smem, static_cast<cudaStream_t>(stream)>>>(q, k, U, qb, emb, N, C, scale, P);
  GR_CHECK_LAUNCH("rpe_scores_softmax_kernel");
  return GR_OK;
}

extern "C" int gr_softmax_rows(float* x, int64_t rows, int cols, void* stream) {
  if (rows < 0 || cols <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x) return GR_ERR_BAD_ARG;
  softmax_rows_kernel<<<ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, cols);
  GR_CHECK_LAUNCH("softmax_rows_kernel");
  return GR_OK;
}

extern "C" int gr_l2_normalize_rows(const float* x, int64_t rows, int C, float eps, float* y, void* stream) {
  if (rows < 0 || C <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x || !y) return GR_ERR_BAD_ARG;
  l2_normalize_rows_kernel<<<ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, C, eps, y);
  GR_CHECK_LAUNCH("l2_normalize_rows_kernel");
  return GR_OK;
}
