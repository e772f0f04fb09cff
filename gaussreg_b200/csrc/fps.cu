// Farthest-point subsampling of the surviving Gaussians down to `point_limit`
// (reference: experiments/geotransformer.gaussian_splatting.indoor/demo.py:44-47 calls the third-party
// `fpsample.bucket_fps_kdline_sampling(points[index], point_limit, h=9)`; fpsample==0.3.2 is not vendored, its
// bucket/KD-line variant is an acceleration structure around EXACT farthest-point sampling, whose result for a
// given start index is what is restated here: repeatedly pick the point with the largest squared distance to the
// selected set, squared distances in float32 (dx*dx + dy*dy + dz*dz, no FMA), ties to the lowest index).
//
// One launch per selected point: every block first folds the previous launch's per-block maxima (<= 592 packed
// 64-bit keys) to learn which point was selected, then relaxes the distances of its slice of the cloud against it
// and leaves its own maximum for the next launch.  For 1e6 candidates that is one 4 MB sweep (mostly L2 hits) per
// launch; the K launches are issued from C++ in one C-ABI call.
#include "common.cuh"

namespace gr {

constexpr int kFpsThreads = 256;
constexpr int kFpsMaxBlocks = 148 * 4;

__device__ __forceinline__ unsigned long long fps_key(float d, int i) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
}

__global__ void __launch_bounds__(kFpsThreads) fps_step_kernel(const float* __restrict__ pts, int n, float* __restrict__ dist, int iter,
                                                               int start, long long* __restrict__ sel,
                                                               unsigned long long* __restrict__ partial, int nblk) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned long long sh[kFpsThreads];
  __shared__ int s_cur;
  const int tid = threadIdx.x;
  // ---- which point does this iteration select?
  if (iter == 0) {
    if (tid == 0) s_cur = start;
  } else {
    const unsigned long long* prev = partial + (size_t)((iter - 1) & 1) * kFpsMaxBlocks;
    unsigned long long best = 0ull;
    for (int b = tid; b < nblk; b += kFpsThreads) { const unsigned long long k = prev[b]; if (k > best) best = k; }
    sh[tid] = best;
    __syncthreads();
    for (int o = kFpsThreads / 2; o > 0; o >>= 1) {
      if (tid < o && sh[tid + o] > sh[tid]) sh[tid] = sh[tid + o];
      __syncthreads();
    }
    if (tid == 0) s_cur = (int)(0xffffffffu - (unsigned)(sh[0] & 0xffffffffull));
  }
  __syncthreads();
  const int cur = s_cur;
  if (blockIdx.x == 0 && tid == 0) sel[iter] = cur;
  // ---- relax this block's slice against the new point, keep the slice's farthest candidate
  const float cx = pts[3 * cur], cy = pts[3 * cur + 1], cz = pts[3 * cur + 2];
  unsigned long long best = 0ull;
  for (int i = blockIdx.x * kFpsThreads + tid; i < n; i += gridDim.x * kFpsThreads) {
    const float dx = __fsub_rn(pts[3 * i], cx), dy = __fsub_rn(pts[3 * i + 1], cy), dz = __fsub_rn(pts[3 * i + 2], cz);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float old = iter == 0 ? d : dist[i];
    const float m = fminf(old, d);
    dist[i] = m;
    const unsigned long long k = fps_key(m, i);
    if (k > best) best = k;
  }
  __syncthreads();
  sh[tid] = best;
  __syncthreads();
  for (int o = kFpsThreads / 2; o > 0; o >>= 1) {
    if (tid < o && sh[tid + o] > sh[tid]) sh[tid] = sh[tid + o];
    __syncthreads();
  }
  if (tid == 0) partial[(size_t)(iter & 1) * kFpsMaxBlocks + blockIdx.x] = sh[0];
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_farthest_point_sample_workspace_size(int64_t n_points) {
  return (((size_t)n_points * sizeof(float) + 255) & ~size_t(255)) + 2 * kFpsMaxBlocks * sizeof(unsigned long long) + 256;
}

/* points (n,3) f32 -> out_idx (k) i64: exact farthest-point sampling in selection order, first pick = start_idx. */
extern "C" int gr_farthest_point_sample(const float* points, int64_t n_points, int k, int64_t start_idx, int64_t* out_idx, void* ws,
                                        size_t ws_bytes, void* stream) {
  if (n_points <= 0 || n_points >= (1ll << 31) || k <= 0 || k > n_points || start_idx < 0 || start_idx >= n_points) return GR_ERR_BAD_ARG;
  if (!points || !out_idx) return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < gr_farthest_point_sample_workspace_size(n_points)) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws, ws_bytes);
  float* dist = c.take<float>((size_t)n_points);
  unsigned long long* partial = c.take<unsigned long long>(2 * kFpsMaxBlocks);
  int nblk = ceil_div(n_points, kFpsThreads * 4);
  if (nblk > kFpsMaxBlocks) nblk = kFpsMaxBlocks;
  if (nblk < 1) nblk = 1;
  for (int it = 0; it < k; ++it) {
    GR_CHECK_CUDA(launch_pdl(fps_step_kernel, dim3(nblk), dim3(kFpsThreads), (size_t)(0), st, points, (int)n_points, dist, it, (int)start_idx, reinterpret_cast<long long*>(out_idx),
                                                  partial, nblk));
    if (it == 0 || it == k - 1) GR_CHECK_LAUNCH("fps_step_kernel");
  }
  count_launch(k > 2 ? k - 2 : 0);
  return GR_OK;
}
