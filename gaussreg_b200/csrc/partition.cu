// P1: point-to-node partition (geotransformer/modules/ops/pointcloud_partition.py:61-111) without the
// (M_c x N_f) distance matrix: nearest node per fine point, then per node the `point_limit` closest of
// its OWN points, ascending by squared distance (ties by index), sentinel N_f / mask elsewhere.
#include "common.cuh"
#include "scan.cuh"

namespace gr {

// pairwise_distance.py:19-30:  d = (|x|^2 - 2 x.y) + |y|^2, clamped at 0  (x = node, y = point)
__device__ __forceinline__ float pd_sqdist(float xx, float x0, float x1, float x2, float yy, float y0, float y1, float y2) {
  const float xy = fmaf(x2, y2, fmaf(x1, y1, __fmul_rn(x0, y0)));
  const float d = __fadd_rn(__fsub_rn(xx, __fmul_rn(2.0f, xy)), yy);
  return fmaxf(d, 0.0f);
}
__device__ __forceinline__ float sq_norm3(float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}

constexpr int kNodeTile = 1024;

__global__ void __launch_bounds__(256) nearest_node_kernel(const float* __restrict__ pts, int N, const float* __restrict__ nodes,
                                                           int M, int* __restrict__ p2n, float* __restrict__ dmin,
                                                           uint32_t* __restrict__ node_cnt) {
  pdl_wait();
  pdl_trigger();
  __shared__ float4 sh[kNodeTile];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float y0 = 0.f, y1 = 0.f, y2 = 0.f, yy = 0.f;
  if (n < N) { y0 = pts[3 * n]; y1 = pts[3 * n + 1]; y2 = pts[3 * n + 2]; yy = sq_norm3(y0, y1, y2); }
  float best = INFINITY;
  int arg = 0;
  for (int base = 0; base < M; base += kNodeTile) {
    const int cnt = min(kNodeTile, M - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float a = nodes[3 * (base + i)], b = nodes[3 * (base + i) + 1], c = nodes[3 * (base + i) + 2];
      sh[i] = make_float4(a, b, c, sq_norm3(a, b, c));
    }
    __syncthreads();
    if (n < N) {
      for (int i = 0; i < cnt; ++i) {
        const float4 x = sh[i];
        const float d = pd_sqdist(x.w, x.x, x.y, x.z, yy, y0, y1, y2);
        if (d < best) { best = d; arg = base + i; }  // first minimum wins
      }
    }
  }
  if (n < N) {
    p2n[n] = arg;
    dmin[n] = best;
    atomicAdd(&node_cnt[arg], 1u);
  }
}

__global__ void __launch_bounds__(256) node_scatter_kernel(const int* __restrict__ p2n, int N, const uint32_t* __restrict__ node_off,
                                                           uint32_t* __restrict__ node_fill, int* __restrict__ plist) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m = p2n[n];
  plist[node_off[m] + atomicAdd(&node_fill[m], 1u)] = n;
}

// one CTA per node: rank its points by (distance, index), keep the first K
__global__ void __launch_bounds__(128) node_knn_kernel(const float* __restrict__ dmin, const uint32_t* __restrict__ node_off,
                                                       const int* __restrict__ plist, int N, int M, int K,
                                                       long long* __restrict__ knn_idx, unsigned char* __restrict__ knn_mask,
                                                       unsigned char* __restrict__ node_mask) {
  pdl_wait();
  pdl_trigger();
  const int m = blockIdx.x;
  const uint32_t s = node_off[m], e = node_off[m + 1];
  const int cnt = (int)(e - s);
  if (threadIdx.x == 0) node_mask[m] = cnt > 0 ? 1 : 0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (k >= cnt) { knn_idx[(long long)m * K + k] = N; knn_mask[(long long)m * K + k] = 0; }
  }
  for (int a = threadIdx.x; a < cnt; a += blockDim.x) {
    const int ia = plist[s + a];
    const unsigned long long ka = ((unsigned long long)__float_as_uint(dmin[ia]) << 32) | (unsigned int)ia;
    int rank = 0;
    for (int b = 0; b < cnt; ++b) {
      const int ib = plist[s + b];
      const unsigned long long kb = ((unsigned long long)__float_as_uint(dmin[ib]) << 32) | (unsigned int)ib;
      rank += kb < ka ? 1 : 0;
    }
    if (rank < K) { knn_idx[(long long)m * K + rank] = ia; knn_mask[(long long)m * K + rank] = 1; }
  }
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_point_to_node_workspace_size(int64_t n_points, int64_t n_nodes) {
  Carver c(nullptr, 0);
  c.take<float>(n_points);
  c.take<uint32_t>(n_nodes + 1);
  c.take<uint32_t>(n_nodes + 1);
  c.take<int>(n_points);
  c.take<uint32_t>(scan_workspace_elems(n_nodes + 1));
  return c.off;
}

/* P1.  points (N,3), nodes (M,3) -> point_to_node (N) i32, node_masks (M) u8, knn_indices (M,K) i64 (sentinel N),
 * knn_masks (M,K) u8. */
extern "C" int gr_point_to_node_partition(const float* points, int N, const float* nodes, int M, int point_limit,
                                          int32_t* point_to_node, uint8_t* node_masks, int64_t* knn_indices,
                                          uint8_t* knn_masks, void* ws, size_t ws_bytes, void* stream) {
  if (N <= 0 || M <= 0 || point_limit <= 0) return GR_ERR_BAD_ARG;
  if (!points || !nodes || !point_to_node || !node_masks || !knn_indices || !knn_masks) return GR_ERR_BAD_ARG;
  Carver c(ws, ws_bytes);
  float* dmin = c.take<float>(N);
  uint32_t* cnt = c.take<uint32_t>(M + 1);
  uint32_t* fill = c.take<uint32_t>(M + 1);
  int* plist = c.take<int>(N);
  uint32_t* sws = c.take<uint32_t>(scan_workspace_elems(M + 1));
  if (!ws || !c.ok) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GR_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(M + 1) * sizeof(uint32_t), st));
  GR_CHECK_CUDA(cudaMemsetAsync(fill, 0, (size_t)(M + 1) * sizeof(uint32_t), st));
  GR_CHECK_CUDA(launch_pdl(nearest_node_kernel, dim3(ceil_div(N, 256)), dim3(256), (size_t)(0), st, points, N, nodes, M, point_to_node, dmin, cnt));
  GR_CHECK_LAUNCH("nearest_node_kernel");
  int rc = exclusive_scan_u32(cnt, cnt, M + 1, sws, st);
  if (rc != GR_OK) return rc;
  GR_CHECK_CUDA(launch_pdl(node_scatter_kernel, dim3(ceil_div(N, 256)), dim3(256), (size_t)(0), st, point_to_node, N, cnt, fill, plist));
  GR_CHECK_LAUNCH("node_scatter_kernel");
  GR_CHECK_CUDA(launch_pdl(node_knn_kernel, dim3(M), dim3(128), (size_t)(0), st, dmin, cnt, plist, N, M, point_limit, reinterpret_cast<long long*>(knn_indices), knn_masks,
                                     node_masks));
  GR_CHECK_LAUNCH("node_knn_kernel");
  return GR_OK;
}
