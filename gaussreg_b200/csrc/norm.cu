// K2 normalisation kernels: GroupNorm over a stacked pair (statistics across ALL rows, as
// geotransformer/modules/kpconv/modules.py:33-50 feeds (1, C, N) to nn.GroupNorm) with fused
// residual add + LeakyReLU, and LayerNorm(a + b) for the transformer (rpe_transformer.py:101-103,
// output_layer.py:14-21).  Statistics are accumulated in double in a fixed order (deterministic).
#include <stdlib.h>

#include "common.cuh"

namespace gr {

constexpr int kGnMaxBlocks = 148 * 4;

__host__ __device__ inline int gn_rows_per_block(long long n_rows) {
  long long r = (n_rows + kGnMaxBlocks - 1) / kGnMaxBlocks;
  r = (r + 7) / 8 * 8;  // multiple of the widest row-lane count
  return (int)(r < 8 ? 8 : r);
}

// partial[blk][g] = (sum, sumsq) over the block's rows and the group's channels.
// Thread layout: consecutive threads walk consecutive channels (coalesced); for C <= 256 the remaining
// thread bits walk rows.  Per-thread accumulation in fp32 over at most a few dozen values, then double.
__global__ void __launch_bounds__(256) groupnorm_partial_kernel(const float* __restrict__ x, int N, int C, int G,
                                                                int rows_per_block, double2* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double2 sh[];  // [max(C, 256)]
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  const int tid = threadIdx.x;
  if (C <= 256) {
    const int R = 256 / C;  // row lanes (C is a multiple of G=32, typically a power of two)
    const int c = tid % C, rs = tid / C;
    double s = 0.0, q = 0.0;
    if (rs < R) {
      float fs = 0.f, fq = 0.f;
      int cnt = 0;
      for (int r = r0 + rs; r < r1; r += R) {
        const float v = x[(long long)r * C + c];
        fs += v; fq = fmaf(v, v, fq);
        if (++cnt == 32) { s += (double)fs; q += (double)fq; fs = 0.f; fq = 0.f; cnt = 0; }
      }
      s += (double)fs; q += (double)fq;
    }
    sh[tid] = make_double2(s, q);
    __syncthreads();
    if (tid < C) {  // fold the row lanes (fixed order)
      double2 a = sh[tid];
      for (int k = 1; k < R; ++k) { const double2 b = sh[tid + k * C]; a.x += b.x; a.y += b.y; }
      sh[tid] = a;
    }
    __syncthreads();
  } else {
    for (int c = tid; c < C; c += 256) {
      double s = 0.0, q = 0.0;
      for (int r = r0; r < r1; ++r) {
        const double v = (double)x[(long long)r * C + c];
        s += v; q += v * v;
      }
      sh[c] = make_double2(s, q);
    }
    __syncthreads();
  }
  const int cg = C / G;
  for (int g = tid; g < G; g += 256) {
    double s = 0.0, q = 0.0;
    for (int c = g * cg; c < (g + 1) * cg; ++c) { s += sh[c].x; q += sh[c].y; }
    partial[(long long)blockIdx.x * G + g] = make_double2(s, q);
  }
}

// stats[g] = (mean, rstd); one CTA per group, fixed-order tree over the block partials
__global__ void __launch_bounds__(256) groupnorm_finalize_kernel(const double2* __restrict__ partial, int nblk, int G, long long count,
                                                                 float eps, float2* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  __shared__ double shs[8], shq[8];
  const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = 0.0, q = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) { const double2 p = partial[(long long)b * G + g]; s += p.x; q += p.y; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if (lane == 0) { shs[warp] = s; shq[warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.0; q = 0.0;
    for (int w = 0; w < 8; ++w) { s += shs[w]; q += shq[w]; }
    const double mean = s / (double)count;
    double var = q / (double)count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

// partial[blk][g] = (sum, sumsq) over the block's rows and the group's channels (float4 form of the kernel above;
// measured on B200: 1.14 ms vs 1.27 ms per pair for the 47 norms.  Folding the partials in the last-finishing CTA
// instead of groupnorm_finalize_kernel was tried and is SLOWER (1.75 ms): one CTA reading 300 KB of partials costs
// more than a 32-CTA launch).
// Thread layout: a thread owns 4 consecutive channels (one 16-byte load per row); `lanes` = min(C/4, 256) threads
// span a row, the remaining thread bits walk rows.  fp32 accumulation over at most 32 rows, then double.
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ x, int N, int C, int G, int rows_per_block,
                                                              double2* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double2 sh[];  // chs[C] channel sums, then stage[R * C] when several row lanes share a channel
  double2* chs = sh;
  double2* stage = sh + C;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  const int tid = threadIdx.x;
  const int c4n = C >> 2;
  const int lanes = c4n < 256 ? c4n : 256;
  const int R = 256 / lanes;
  const int rs = tid / lanes;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
  if (rs < R) {
    for (int cc = tid % lanes; cc < c4n; cc += lanes) {
      double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
      float4 fs = make_float4(0.f, 0.f, 0.f, 0.f), fq = fs;
      int cnt = 0;
      for (int r = r0 + rs; r < r1; r += R) {
        const float4 v = __ldg(x4 + (long long)r * c4n + cc);
        fs.x += v.x; fs.y += v.y; fs.z += v.z; fs.w += v.w;
        fq.x = fmaf(v.x, v.x, fq.x); fq.y = fmaf(v.y, v.y, fq.y); fq.z = fmaf(v.z, v.z, fq.z); fq.w = fmaf(v.w, v.w, fq.w);
        if (++cnt == 32) {
          s[0] += (double)fs.x; s[1] += (double)fs.y; s[2] += (double)fs.z; s[3] += (double)fs.w;
          q[0] += (double)fq.x; q[1] += (double)fq.y; q[2] += (double)fq.z; q[3] += (double)fq.w;
          fs = make_float4(0.f, 0.f, 0.f, 0.f); fq = fs; cnt = 0;
        }
      }
      s[0] += (double)fs.x; s[1] += (double)fs.y; s[2] += (double)fs.z; s[3] += (double)fs.w;
      q[0] += (double)fq.x; q[1] += (double)fq.y; q[2] += (double)fq.z; q[3] += (double)fq.w;
      double2* dst = (R == 1) ? chs + 4 * cc : stage + (size_t)rs * C + 4 * cc;
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k] = make_double2(s[k], q[k]);
    }
  }
  __syncthreads();
  if (R > 1) {
    for (int c = tid; c < C; c += 256) {  // fold the row lanes (fixed order)
      double2 a = stage[c];
      for (int k = 1; k < R; ++k) { const double2 b = stage[(size_t)k * C + c]; a.x += b.x; a.y += b.y; }
      chs[c] = a;
    }
    __syncthreads();
  }
  const int cg = C / G;
  for (int g = tid; g < G; g += 256) {
    double s = 0.0, q = 0.0;
    for (int c = g * cg; c < (g + 1) * cg; ++c) { s += chs[c].x; q += chs[c].y; }
    partial[(long long)blockIdx.x * G + g] = make_double2(s, q);
  }
}

// y = act( (x - mean) * rstd * gamma + beta  [+ add] ); one thread per 4 consecutive channels (C % 4 == 0)
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ x, int n_rows, int C, int G,
                                                              const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ add,
                                                              int act, float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  const int c4 = C >> 2, cg = C / G;
  const long long total4 = (long long)n_rows * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
    float v[4] = {xv.x, xv.y, xv.z, xv.w};
    const float g4[4] = {gm.x, gm.y, gm.z, gm.w}, b4[4] = {bt.x, bt.y, bt.z, bt.w};
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (add) { const float4 av = reinterpret_cast<const float4*>(add)[i]; a4[0] = av.x; a4[1] = av.y; a4[2] = av.z; a4[3] = av.w; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 st = stats[(c + k) / cg];
      float t = (v[k] - st.x) * st.y * g4[k] + b4[k];
      if (add) t += a4[k];
      if (act == 2) t = t > 0.f ? t : 0.1f * t;
      else if (act == 1) t = fmaxf(t, 0.f);
      v[k] = t;
    }
    reinterpret_cast<float4*>(y)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Small inputs (a few MB: the coarse pyramid stages): one CTA per group walks all rows of its C/G channels twice --
// statistics, then apply -- in ONE launch; the second pass is served by L1/L2.  Three launches cost more than the
// data movement there.  Same arithmetic as the three-kernel path: fp32 partial sums over at most 32 rows folded
// into double, fixed reduction order, float (mean, rstd), identical apply expression.
constexpr int kGnSmallThreads = 1024;

__global__ void __launch_bounds__(kGnSmallThreads) groupnorm_small_kernel(const float* __restrict__ x, int N, int C, int G,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                        float eps, const float* __restrict__ add, int act,
                                                                        float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  __shared__ double shs[32], shq[32];
  __shared__ float s_stats[2];
  const int g = blockIdx.x;
  const int cg = C / G, cg4 = cg >> 2;          // cg % 4 == 0 (checked by the caller)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cl = tid % cg4, rs = tid / cg4;      // float4 column inside the group, row lane
  const int R = kGnSmallThreads / cg4;
  const int c0 = g * cg + 4 * cl;
  const bool worker = rs < R;
  double s = 0.0, q = 0.0;
  if (worker) {
    float4 fs = make_float4(0.f, 0.f, 0.f, 0.f), fq = fs;
    int cnt = 0;
    for (int r = rs; r < N; r += R) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long long)r * C + c0));
      fs.x += v.x; fs.y += v.y; fs.z += v.z; fs.w += v.w;
      fq.x = fmaf(v.x, v.x, fq.x); fq.y = fmaf(v.y, v.y, fq.y); fq.z = fmaf(v.z, v.z, fq.z); fq.w = fmaf(v.w, v.w, fq.w);
      if (++cnt == 32) {
        s += ((double)fs.x + (double)fs.y) + ((double)fs.z + (double)fs.w);
        q += ((double)fq.x + (double)fq.y) + ((double)fq.z + (double)fq.w);
        fs = make_float4(0.f, 0.f, 0.f, 0.f); fq = fs; cnt = 0;
      }
    }
    s += ((double)fs.x + (double)fs.y) + ((double)fs.z + (double)fs.w);
    q += ((double)fq.x + (double)fq.y) + ((double)fq.z + (double)fq.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if (lane == 0) { shs[warp] = s; shq[warp] = q; }
  __syncthreads();
  if (tid == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < kGnSmallThreads / 32; ++w) { ts += shs[w]; tq += shq[w]; }
    const double count = (double)N * (double)cg;
    const double mean = ts / count;
    double var = tq / count - mean * mean;
    if (var < 0.0) var = 0.0;
    s_stats[0] = (float)mean;
    s_stats[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (!worker) return;
  const float mean = s_stats[0], rstd = s_stats[1];
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c0), bt = *reinterpret_cast<const float4*>(beta + c0);
  const float g4[4] = {gm.x, gm.y, gm.z, gm.w}, b4[4] = {bt.x, bt.y, bt.z, bt.w};
  for (int r = rs; r < N; r += R) {
    const long long o = (long long)r * C + c0;
    const float4 xv = *reinterpret_cast<const float4*>(x + o);
    float v[4] = {xv.x, xv.y, xv.z, xv.w};
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (add) { const float4 av = *reinterpret_cast<const float4*>(add + o); a4[0] = av.x; a4[1] = av.y; a4[2] = av.z; a4[3] = av.w; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = (v[k] - mean) * rstd * g4[k] + b4[k];
      if (add) t += a4[k];
      if (act == 2) t = t > 0.f ? t : 0.1f * t;
      else if (act == 1) t = fmaxf(t, 0.f);
      v[k] = t;
    }
    *reinterpret_cast<float4*>(y + o) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// y[row] = LayerNorm(a[row] + b[row]) * gamma + beta ; one warp per row, C <= 1024, C % 32 == 0
__global__ void __launch_bounds__(256) layernorm_add_kernel(const float* __restrict__ a, const float* __restrict__ b, int rows,
                                                            int C, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int per = C / 32;
  float v[32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i < per) {
      const long long o = (long long)row * C + lane + 32 * i;
      v[i] = a[o] + (b ? b[o] : 0.f);
      s += v[i];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) {
      const int c = lane + 32 * i;
      y[(long long)row * C + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_group_norm_workspace_size(int64_t n_rows, int groups) {
  const size_t nblk = (size_t)kGnMaxBlocks + 1;
  (void)n_rows;
  return nblk * groups * sizeof(double2) + groups * sizeof(float2) + 512;
}

/* K2: y = act(GroupNorm_G(x over all n_rows) [+ add]); act: 0 none, 1 relu, 2 leaky(0.1).  y may alias x. */
extern "C" int gr_group_norm(const float* x, int64_t n_rows, int C, int groups, const float* gamma, const float* beta,
                             float eps, const float* add, int act, float* y, void* ws, size_t ws_bytes, void* stream) {
  if (n_rows < 0 || C <= 0 || groups <= 0 || C % groups != 0 || C > 4096 || C % 4 != 0) return GR_ERR_BAD_ARG;
  if (n_rows == 0) return GR_OK;
  if (!x || !y || !gamma || !beta) return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < gr_group_norm_workspace_size(n_rows, groups)) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rpb = gn_rows_per_block(n_rows);
  const int nblk = ceil_div(n_rows, rpb);
  double2* partial = static_cast<double2*>(ws);
  float2* stats = reinterpret_cast<float2*>(static_cast<char*>(ws) + (((size_t)nblk * groups * sizeof(double2) + 255) & ~size_t(255)));
  static int small_knob = -1;  // GAUSSREG_GN_SMALL: 1 = single-launch kernel for inputs of at most 16 MB (C/G a multiple of 8)
  if (small_knob < 0) { const char* e = getenv("GAUSSREG_GN_SMALL"); small_knob = e ? atoi(e) : 0; }
  {
    const int cg = C / groups;
    if (small_knob && cg % 8 == 0 && cg / 4 <= kGnSmallThreads && (long long)n_rows * C * 4 <= (16ll << 20)) {
      GR_CHECK_CUDA(launch_pdl(groupnorm_small_kernel, dim3(groups), dim3(kGnSmallThreads), (size_t)(0), st, x, (int)n_rows, C, groups, gamma, beta, eps, add, act, y));
      GR_CHECK_LAUNCH("groupnorm_small_kernel");
      return GR_OK;
    }
  }
  static int variant = -1;  // 0: scalar partial kernel, 1: float4 stats kernel (default)
  if (variant < 0) { const char* e = getenv("GAUSSREG_GN"); variant = e ? atoi(e) : 1; }
  if (variant == 0) {
    const size_t smem = (size_t)(C > 256 ? C : 256) * sizeof(double2);
    if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(groupnorm_partial_kernel), (int)smem));
    GR_CHECK_CUDA(launch_pdl(groupnorm_partial_kernel, dim3(nblk), dim3(256), (size_t)(smem), st, x, (int)n_rows, C, groups, rpb, partial));
    GR_CHECK_LAUNCH("groupnorm_partial_kernel");
  } else {
    const int lanes = (C / 4) < 256 ? (C / 4) : 256;
    const int R = 256 / lanes;
    const size_t smem = ((size_t)C + (R > 1 ? (size_t)R * C : 0)) * sizeof(double2);
    GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(groupnorm_stats_kernel), 96 * 1024));
    GR_CHECK_CUDA(launch_pdl(groupnorm_stats_kernel, dim3(nblk), dim3(256), (size_t)(smem), st, x, (int)n_rows, C, groups, rpb, partial));
    GR_CHECK_LAUNCH("groupnorm_stats_kernel");
  }
  GR_CHECK_CUDA(launch_pdl(groupnorm_finalize_kernel, dim3(groups), dim3(256), 0, st, partial, nblk, groups, (long long)n_rows * (C / groups), eps, stats));
  GR_CHECK_LAUNCH("groupnorm_finalize_kernel");
  const long long total4 = (long long)n_rows * (C / 4);
  const int blocks = (int)min((long long)148 * 16, (total4 + 255) / 256);
  GR_CHECK_CUDA(launch_pdl(groupnorm_apply_kernel, dim3(blocks), dim3(256), 0, st, x, (int)n_rows, C, groups, stats, gamma, beta, add, act, y));
  GR_CHECK_LAUNCH("groupnorm_apply_kernel");
  return GR_OK;
}

namespace gr {
// y = act(GN_a(x) + GN_b(x2)): the ResidualBlock tail LeakyReLU(norm(unary2) + norm(shortcut)) in ONE pass over the two raw
// products (kpconv/modules.py:205-224) -- the shortcut's own apply pass (a write and a re-read of (M, out)) disappears.
__global__ void __launch_bounds__(256) groupnorm_apply2_kernel(const float* __restrict__ x, const float* __restrict__ x2, int n_rows, int C,
                                                               int G, const float2* __restrict__ stats, const float2* __restrict__ stats2,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ gamma2, const float* __restrict__ beta2, int act,
                                                               float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  const int c4 = C >> 2, cg = C / G;
  const long long total4 = (long long)n_rows * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const float4 xv = reinterpret_cast<const float4*>(x)[i], zv = reinterpret_cast<const float4*>(x2)[i];
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
    const float4 gm2 = *reinterpret_cast<const float4*>(gamma2 + c), bt2 = *reinterpret_cast<const float4*>(beta2 + c);
    float v[4] = {xv.x, xv.y, xv.z, xv.w};
    const float z[4] = {zv.x, zv.y, zv.z, zv.w};
    const float g4[4] = {gm.x, gm.y, gm.z, gm.w}, b4[4] = {bt.x, bt.y, bt.z, bt.w};
    const float h4[4] = {gm2.x, gm2.y, gm2.z, gm2.w}, d4[4] = {bt2.x, bt2.y, bt2.z, bt2.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 st = stats[(c + k) / cg], st2 = stats2[(c + k) / cg];
      float t = (v[k] - st.x) * st.y * g4[k] + b4[k];          // same expression order as groupnorm_apply_kernel ...
      const float a = (z[k] - st2.x) * st2.y * h4[k] + d4[k];  // ... for both terms, so the result is bit-identical to
      t += a;                                                  // apply(shortcut) followed by apply(unary2, add)
      if (act == 2) t = t > 0.f ? t : 0.1f * t;
      else if (act == 1) t = fmaxf(t, 0.f);
      v[k] = t;
    }
    reinterpret_cast<float4*>(y)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

int group_norm_finalize(const double2* partial, int nblk, long long n_rows, int C, int groups, float eps, float2* stats, void* stream) {
  if (n_rows <= 0) return GR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GR_CHECK_CUDA(launch_pdl(groupnorm_finalize_kernel, dim3(groups), dim3(256), 0, st, partial, nblk, groups, n_rows * (long long)(C / groups), eps, stats));
  GR_CHECK_LAUNCH("groupnorm_finalize_kernel");
  return GR_OK;
}

int group_norm_apply2(const float* x, const float* x2, long long n_rows, int C, int groups, const float2* stats, const float2* stats2,
                      const float* gamma, const float* beta, const float* gamma2, const float* beta2, int act, float* y, void* stream) {
  if (n_rows <= 0) return GR_OK;
  if (C % 4 != 0) return GR_ERR_BAD_ARG;
  const long long total4 = n_rows * (long long)(C / 4);
  const int blocks = (int)min((long long)148 * 16, (total4 + 255) / 256);
  GR_CHECK_CUDA(launch_pdl(groupnorm_apply2_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x, x2, (int)n_rows, C, groups, stats,
                           stats2, gamma, beta, gamma2, beta2, act, y));
  GR_CHECK_LAUNCH("groupnorm_apply2_kernel");
  return GR_OK;
}

int group_norm_from_partial(const float* x, long long n_rows, int C, int groups, const double2* partial, int nblk,
                            const float* gamma, const float* beta, float eps, const float* add, int act, float* y,
                            float2* stats, void* stream) {
  if (n_rows <= 0) return GR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GR_CHECK_CUDA(launch_pdl(groupnorm_finalize_kernel, dim3(groups), dim3(256), 0, st, partial, nblk, groups, n_rows * (long long)(C / groups), eps, stats));
  GR_CHECK_LAUNCH("groupnorm_finalize_kernel");
  const long long total4 = n_rows * (long long)(C / 4);
  const int blocks = (int)min((long long)148 * 16, (total4 + 255) / 256);
  GR_CHECK_CUDA(launch_pdl(groupnorm_apply_kernel, dim3(blocks), dim3(256), 0, st, x, (int)n_rows, C, groups, stats, gamma, beta, add, act, y));
  GR_CHECK_LAUNCH("groupnorm_apply_kernel");
  return GR_OK;
}
}  // namespace gr

/* T2/T3: y = LayerNorm(a + b) (b may be NULL). */
extern "C" int gr_layer_norm_add(const float* a, const float* b, int64_t rows, int C, const float* gamma, const float* beta,
                                 float eps, float* y, void* stream) {
  if (rows < 0 || C <= 0 || C % 32 != 0 || C > 1024) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!a || !y || !gamma || !beta) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(layernorm_add_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), a, b, (int)rows, C, gamma, beta, eps, y));
  GR_CHECK_LAUNCH("layernorm_add_kernel");
  return GR_OK;
}
