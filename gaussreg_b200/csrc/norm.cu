// K2 normalisation kernels: GroupNorm over a stacked pair (statistics across ALL rows, as
// geotransformer/modules/kpconv/modules.py:33-50 feeds (1, C, N) to nn.GroupNorm) with fused
// residual add + LeakyReLU, and LayerNorm(a + b) for the transformer (rpe_transformer.py:101-103,
// output_layer.py:14-21).  Statistics are accumulated in double in a fixed order (deterministic).
#include "common.cuh"

namespace gr {

constexpr int kGnRowsPerBlock = 256;

// partial[blk][g] = (sum, sumsq) over the block's rows and the group's channels
__global__ void __launch_bounds__(256) groupnorm_partial_kernel(const float* __restrict__ x, int N, int C, int G,
                                                                double2* __restrict__ partial) {
  extern __shared__ double2 sh[];  // [C]
  const int r0 = blockIdx.x * kGnRowsPerBlock, r1 = min(N, r0 + kGnRowsPerBlock);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int r = r0; r < r1; ++r) {
      const double v = (double)x[(long long)r * C + c];
      s += v; q += v * v;
    }
    sh[c] = make_double2(s, q);
  }
  __syncthreads();
  const int cg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int c = g * cg; c < (g + 1) * cg; ++c) { s += sh[c].x; q += sh[c].y; }
    partial[(long long)blockIdx.x * G + g] = make_double2(s, q);
  }
}

// stats[g] = (mean, rstd)
__global__ void groupnorm_finalize_kernel(const double2* __restrict__ partial, int nblk, int G, long long count, float eps,
                                          float2* __restrict__ stats) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nblk; ++b) { const double2 p = partial[(long long)b * G + g]; s += p.x; q += p.y; }
  const double mean = s / (double)count;
  double var = q / (double)count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// y = act( (x - mean) * rstd * gamma + beta  [+ add] )
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ x, long long total, int C, int G,
                                                              const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ add,
                                                              int act, float* __restrict__ y) {
  const int cg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float2 st = stats[c / cg];
    float v = (x[i] - st.x) * st.y * gamma[c] + beta[c];
    if (add) v += add[i];
    if (act == 2) v = v > 0.f ? v : 0.1f * v;
    else if (act == 1) v = fmaxf(v, 0.f);
    y[i] = v;
  }
}

// y[row] = LayerNorm(a[row] + b[row]) * gamma + beta ; one warp per row, C <= 1024, C % 32 == 0
__global__ void __launch_bounds__(256) layernorm_add_kernel(const float* __restrict__ a, const float* __restrict__ b, int rows,
                                                            int C, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ y) {
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
  const size_t nblk = (size_t)((n_rows + kGnRowsPerBlock - 1) / kGnRowsPerBlock);
  return nblk * groups * sizeof(double2) + groups * sizeof(float2) + 512;
}

/* K2: y = act(GroupNorm_G(x over all n_rows) [+ add]); act: 0 none, 1 relu, 2 leaky(0.1).  y may alias x. */
extern "C" int gr_group_norm(const float* x, int64_t n_rows, int C, int groups, const float* gamma, const float* beta,
                             float eps, const float* add, int act, float* y, void* ws, size_t ws_bytes, void* stream) {
  if (n_rows < 0 || C <= 0 || groups <= 0 || C % groups != 0 || C > 4096) return GR_ERR_BAD_ARG;
  if (n_rows == 0) return GR_OK;
  if (!x || !y || !gamma || !beta) return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < gr_group_norm_workspace_size(n_rows, groups)) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = ceil_div(n_rows, kGnRowsPerBlock);
  double2* partial = static_cast<double2*>(ws);
  float2* stats = reinterpret_cast<float2*>(static_cast<char*>(ws) + (((size_t)nblk * groups * sizeof(double2) + 255) & ~size_t(255)));
  const size_t smem = (size_t)C * sizeof(double2);
  if (smem > 48 * 1024) GR_CHECK_CUDA(cudaFuncSetAttribute(groupnorm_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  groupnorm_partial_kernel<<<nblk, 256, smem, st>>>(x, (int)n_rows, C, groups, partial);
  GR_CHECK_LAUNCH("groupnorm_partial_kernel");
  groupnorm_finalize_kernel<<<ceil_div(groups, 64), 64, 0, st>>>(partial, nblk, groups, (long long)n_rows * (C / groups), eps, stats);
  GR_CHECK_LAUNCH("groupnorm_finalize_kernel");
  const long long total = (long long)n_rows * C;
  const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
  groupnorm_apply_kernel<<<blocks, 256, 0, st>>>(x, total, C, groups, stats, gamma, beta, add, act, y);
  GR_CHECK_LAUNCH("groupnorm_apply_kernel");
  return GR_OK;
}

/* T2/T3: y = LayerNorm(a + b) (b may be NULL). */
extern "C" int gr_layer_norm_add(const float* a, const float* b, int64_t rows, int C, const float* gamma, const float* beta,
                                 float eps, float* y, void* stream) {
  if (rows < 0 || C <= 0 || C % 32 != 0 || C > 1024) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!a || !y || !gamma || !beta) return GR_ERR_BAD_ARG;
  layernorm_add_kernel<<<ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, (int)rows, C, gamma, beta, eps, y);
  GR_CHECK_LAUNCH("layernorm_add_kernel");
  return GR_OK;
}
