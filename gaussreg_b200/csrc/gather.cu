// K3 gather kernels: neighbourhood max-pool (kpconv/functional.py:54-67), nearest upsample fused with
// the skip concatenation (functional.py:6-22 + backbone.py:195-208), generic padded row gather
// (modules/ops/index_select.py with the zero sentinel row used by model.py:106-109,178-181).
#include "common.cuh"

namespace gr {

// out[m, c] = max_h padded(x)[idx[m,h], c], padded row = 0
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ x, int Ns, int C,
                                                      const long long* __restrict__ idx, int H, long long ldi, int M,
                                                      float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int m = blockIdx.x;
  extern __shared__ int sh_idx[];
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const long long j = idx[(long long)m * ldi + h];
    sh_idx[h] = (j >= Ns || j < 0) ? -1 : (int)j;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = -INFINITY;
    for (int h = 0; h < H; ++h) {
      const int j = sh_idx[h];
      v = fmaxf(v, j < 0 ? 0.f : x[(long long)j * C + c]);
    }
    out[(long long)m * C + c] = v;
  }
}

// out[m, 0:C1] = padded(coarse)[idx[m,0]] ; out[m, C1:C1+C2] = skip[m]
__global__ void __launch_bounds__(256) upsample_concat_kernel(const float* __restrict__ coarse, int Nc, int C1,
                                                              const long long* __restrict__ idx, long long ldi,
                                                              const float* __restrict__ skip, int C2, int M,
                                                              float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int m = blockIdx.x;
  const long long j = idx[(long long)m * ldi];
  const bool pad = j >= Nc || j < 0;
  float* o = out + (long long)m * (C1 + C2);
  for (int c = threadIdx.x; c < C1; c += blockDim.x) o[c] = pad ? 0.f : coarse[j * C1 + c];
  for (int c = threadIdx.x; c < C2; c += blockDim.x) o[C1 + c] = skip[(long long)m * C2 + c];
}

// out[r, :] = idx[r] in [0,n) ? x[idx[r], :] : 0
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ x, int n, int C,
                                                          const long long* __restrict__ idx, long long rows,
                                                          float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long j = idx[r];
  const bool pad = j >= n || j < 0;
  for (int c = lane; c < C; c += 32) out[r * C + c] = pad ? 0.f : x[j * C + c];
}

}  // namespace gr

using namespace gr;

extern "C" int gr_maxpool(const float* x, int Ns, int C, const int64_t* idx, int H, int64_t ld_idx, int M, float* out,
                          void* stream) {
  if (C <= 0 || H <= 0 || M < 0 || Ns < 0) return GR_ERR_BAD_ARG;
  if (M == 0) return GR_OK;
  if (!x || !idx || !out) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(maxpool_kernel, dim3(M), dim3(C >= 256 ? 256 : (C >= 128 ? 128 : 64)), (size_t)(H * sizeof(int)), static_cast<cudaStream_t>(stream), x, Ns, C, reinterpret_cast<const long long*>(idx), H, ld_idx, M, out));
  GR_CHECK_LAUNCH("maxpool_kernel");
  return GR_OK;
}

extern "C" int gr_upsample_concat(const float* coarse, int Nc, int C1, const int64_t* idx, int64_t ld_idx, const float* skip,
                                  int C2, int M, float* out, void* stream) {
  if (C1 <= 0 || C2 < 0 || M < 0 || Nc < 0) return GR_ERR_BAD_ARG;
  if (M == 0) return GR_OK;
  if (!coarse || !idx || !out || (C2 > 0 && !skip)) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(upsample_concat_kernel, dim3(M), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), coarse, Nc, C1, reinterpret_cast<const long long*>(idx),
                                                                           ld_idx, skip, C2, M, out));
  GR_CHECK_LAUNCH("upsample_concat_kernel");
  return GR_OK;
}

extern "C" int gr_gather_rows(const float* x, int n, int C, const int64_t* idx, int64_t rows, float* out, void* stream) {
  if (C <= 0 || rows < 0 || n < 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x || !idx || !out) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(gather_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), x, n, C, reinterpret_cast<const long long*>(idx),
                                                                                        rows, out));
  GR_CHECK_LAUNCH("gather_rows_kernel");
  return GR_OK;
}
