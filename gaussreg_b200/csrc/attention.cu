// T2/T3: attention score kernels.
//
// RPE self-attention (geotransformer/modules/transformer/rpe_transformer.py:50-70):
//     score[h,n,m] = ( q_h[n].k_h[m] + sum_c q_h[n,c] * (W_p e[n,m] + b_p)_{h,c} ) / sqrt(d_head)
// The reference materialises p = proj_p(e) as (1,H,N,N,64).  Here the second term is reassociated to
//     (W_p,h^T q_h[n]) . e[n,m] + q_h[n].b_p,h  =  U[h,n,:] . e[n,m,:] + qb[h,n]
// (same mathematics, different rounding: SURVEY.md section 8(a) row T2), so e is streamed once per layer
// and nothing of size N*N*C is written.  U and qb come out of two small GEMMs.
#include <stdlib.h>

#include "common.cuh"

namespace gr {

// One CTA per query row n.  P[h, n, :] = softmax_m(score[h, n, m]);  C = H * DH, H <= 8.
template <int H>
__global__ void __launch_bounds__(256) rpe_scores_softmax_kernel(const float* __restrict__ q, const float* __restrict__ kmat,
                                                                 const float* __restrict__ U, const float* __restrict__ qb,
                                                                 const float* __restrict__ emb, int N, int C, float scale,
                                                                 float* __restrict__ P, long long ldq, long long ldk) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  float* sU = sm;                 // [H][C]
  float* sq = sU + H * C;         // [C]
  float* ss = sq + C;             // [H][N] scores
  __shared__ float red[H][8];
  const int n = blockIdx.x;
  const int DH = C / H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < H * C; i += blockDim.x) sU[i] = U[((long long)(i / C) * N + n) * C + (i % C)];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sq[i] = q[(long long)n * ldq + i];
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
    const float4* k4 = reinterpret_cast<const float4*>(kmat + (long long)m * ldk);
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

// ---- v2 (C = 256, H = 4): warp-cooperative, coalesced stream of the embedding ---------------------------------
// One CTA per query row n, one warp per group of 8 keys.  Lane l owns the channels of float4 #l and #(l+32) of a
// 256-channel row, holds U[h, n, those 8 channels] for the four heads in registers (32 floats), and reads the
// eight embedding rows of the group with sixteen independent, fully coalesced 512-byte warp loads (256 B in
// flight per lane, streamed past L1).  The 8 keys x 4 heads = 32 per-lane partial sums are folded across the
// warp by a transposing butterfly (31 shuffles for all 32 sums instead of 5 per sum); lane L ends up with the
// total of key L/4, head L%4.  The q.k^T term arrives as raw batched-GEMM output in P (gr_rpe_attention_probs
// issues that product first) and is added before the scale, in the reference's order (qk + (p-term)).
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}

template <int OFF, int CNT>
__device__ __forceinline__ void butterfly_step(float (&v)[32], int lane) {
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const float send = up ? v[i] : v[i + CNT];
    const float keep = up ? v[i + CNT] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

constexpr int kRpeKeys = 8;  // keys per warp iteration

__global__ void __launch_bounds__(256, 2) rpe_scores_softmax_v2_kernel(const float* __restrict__ U, const float* __restrict__ qb,
                                                                       const float* __restrict__ emb, int N, float scale,
                                                                       float* __restrict__ P) {
  pdl_wait();
  pdl_trigger();
  constexpr int H = 4, C = 256, C4 = C / 4;
  extern __shared__ float ss[];  // [H][N] scores of this query row
  __shared__ float red[H][8];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;

  float4 u[H][2];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float4* up = reinterpret_cast<const float4*>(U + ((long long)h * N + n) * C);
    u[h][0] = __ldg(up + lane);
    u[h][1] = __ldg(up + 32 + lane);
  }
  const float qbl = qb[(long long)(lane & 3) * N + n];  // this lane's head after the butterfly
  const float4* erow = reinterpret_cast<const float4*>(emb) + (long long)n * N * C4;
  const int ngroups = (N + kRpeKeys - 1) / kRpeKeys;
  for (int g = warp; g < ngroups; g += nwarp) {
    const int m0 = g * kRpeKeys;
    float4 e[kRpeKeys][2];
#pragma unroll
    for (int kk = 0; kk < kRpeKeys; ++kk) {
      const int m = m0 + kk;
      if (m < N) {
        e[kk][0] = ld_stream4(erow + (long long)m * C4 + lane);
        e[kk][1] = ld_stream4(erow + (long long)m * C4 + 32 + lane);
      } else {
        e[kk][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        e[kk][1] = e[kk][0];
      }
    }
    float v[32];
#pragma unroll
    for (int kk = 0; kk < kRpeKeys; ++kk)
#pragma unroll
      for (int h = 0; h < H; ++h) v[kk * H + h] = dot4(u[h][1], e[kk][1], dot4(u[h][0], e[kk][0], 0.f));
    butterfly_step<16, 16>(v, lane);
    butterfly_step<8, 8>(v, lane);
    butterfly_step<4, 4>(v, lane);
    butterfly_step<2, 2>(v, lane);
    butterfly_step<1, 1>(v, lane);
    const int m = m0 + (lane >> 2);
    if (m < N) ss[(lane & 3) * N + m] = v[0] + qbl;
  }
  __syncthreads();
  // score = (q.k + (U.e + q.b_p)) * scale, then softmax over m per head
  for (int h = 0; h < H; ++h) {
    const float* qk = P + ((long long)h * N + n) * N;
    float mx = -INFINITY;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
      const float s = (qk[m] + ss[h * N + m]) * scale;
      ss[h * N + m] = s;
      mx = fmaxf(mx, s);
    }
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
      const float ev = expf(ss[h * N + m] - hmax[h]);
      ss[h * N + m] = ev;
      s += ev;
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
  pdl_wait();
  pdl_trigger();
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
  pdl_wait();
  pdl_trigger();
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

extern "C" int gr_gemm(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB, int trans_b,
                       float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch, float alpha, const float* bias,
                       const float* row_div, const float* residual, int64_t ldr, int64_t strideR, int act, void* stream);

static int rpe_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_RPE"); v = e ? atoi(e) : 2; }  // 1: thread-per-key kernel, 2: warp-cooperative
  return v;
}

/* T2: q,k (N,C) with row pitches ldq / ldk ; U (H,N,C) ; qb (H,N) ; emb (N,N,C) -> P (H,N,N) softmax probabilities. */
extern "C" int gr_rpe_attention_probs_ld(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* U, const float* qb,
                                         const float* emb, int N, int C, int num_heads, float* P, void* stream) {
  if (N <= 0 || C <= 0 || num_heads != 4 || C % num_heads != 0) return GR_ERR_BAD_ARG;
  if (!q || !k || !U || !qb || !emb || !P) return GR_ERR_BAD_ARG;
  const float scale = 1.0f / sqrtf((float)(C / num_heads));
  if (C == 256 && rpe_variant() == 2 && (size_t)num_heads * N * sizeof(float) <= 100 * 1024) {
    const int dh = C / num_heads;
    // raw q_h . k_h^T into P, then the streaming kernel adds the position term and normalises in place
    int rc = gr_gemm(q, ldq, dh, k, ldk, dh, 1, P, N, (int64_t)N * N, N, N, dh, num_heads, 1.f, nullptr, nullptr, nullptr, 0, 0, 0,
                     stream);
    if (rc != GR_OK) return rc;
    const size_t smem = (size_t)num_heads * N * sizeof(float);
    auto kern = rpe_scores_softmax_v2_kernel;
    if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), (int)smem));
    GR_CHECK_CUDA(launch_pdl(kern, dim3(N), dim3(256), (size_t)(smem), static_cast<cudaStream_t>(stream), U, qb, emb, N, scale, P));
    GR_CHECK_LAUNCH("rpe_scores_softmax_v2_kernel");
    return GR_OK;
  }
  if (ldk % 4 != 0) return GR_ERR_BAD_ARG;
  const size_t smem = ((size_t)num_heads * C + C + (size_t)num_heads * N) * sizeof(float);
  if (smem > 200 * 1024) return GR_ERR_CAPACITY;
  auto kern = rpe_scores_softmax_kernel<4>;
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), (int)smem));
  GR_CHECK_CUDA(launch_pdl(kern, dim3(N), dim3(256), (size_t)(smem), static_cast<cudaStream_t>(stream), q, k, U, qb, emb, N, C, scale, P, ldq, ldk));
  GR_CHECK_LAUNCH("rpe_scores_softmax_kernel");
  return GR_OK;
}

extern "C" int gr_rpe_attention_probs(const float* q, const float* k, const float* U, const float* qb, const float* emb, int N,
                                      int C, int num_heads, float* P, void* stream) {
  return gr_rpe_attention_probs_ld(q, C, k, C, U, qb, emb, N, C, num_heads, P, stream);
}

extern "C" int gr_softmax_rows(float* x, int64_t rows, int cols, void* stream) {
  if (rows < 0 || cols <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(softmax_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, rows, cols));
  GR_CHECK_LAUNCH("softmax_rows_kernel");
  return GR_OK;
}

extern "C" int gr_l2_normalize_rows(const float* x, int64_t rows, int C, float eps, float* y, void* stream) {
  if (rows < 0 || C <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x || !y) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(l2_normalize_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), x, rows, C, eps, y));
  GR_CHECK_LAUNCH("l2_normalize_rows_kernel");
  return GR_OK;
}
