// Device-wide exclusive prefix sum over uint32 (three launches: tile reduce, tile-sum scan, tile scan).
// Used by the cell binning (G2) and voxel compaction (G1) passes.  In-place (in == out) is allowed.
#pragma once
#include "common.cuh"

namespace gr {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total, uint32_t* sh /*>=9*/) {
  // inclusive warp scan
  uint32_t x = v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = lane < (kScanThreads / 32) ? sh[lane] : 0u;
    uint32_t t = s;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    if (lane < (kScanThreads / 32)) sh[lane] = t - s;  // exclusive warp offsets
    if (lane == (kScanThreads / 32) - 1) sh[8] = t;    // block total
  }
  __syncthreads();
  uint32_t r = sh[warp] + x - v;
  *total = sh[8];
  __syncthreads();
  return r;
}

static __global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in,
                                                                    uint32_t* __restrict__ tile_sums, int64_t n) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t sh[8];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t t = threadIdx.x < 8 ? sh[threadIdx.x] : 0u;
    t = warp_sum(t);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = t;
  }
}

// single CTA: exclusive scan of tile_sums[0..m) in place
static __global__ void __launch_bounds__(kScanThreads) scan_tilesums_kernel(uint32_t* __restrict__ tile_sums, int m) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t sh[9];
  uint32_t carry = 0;
  for (int base = 0; base < m; base += kScanThreads) {
    int i = base + threadIdx.x;
    uint32_t v = i < m ? tile_sums[i] : 0u;
    uint32_t total;
    uint32_t ex = block_exclusive_scan_256(v, &total, sh);
    if (i < m) tile_sums[i] = carry + ex;
    carry += total;
  }
}

static __global__ void __launch_bounds__(kScanThreads) scan_final_kernel(const uint32_t* in, uint32_t* out,
                                                                   const uint32_t* __restrict__ tile_sums, int64_t n) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t sh[9];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    v[k] = i < n ? in[i] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan_256(s, &total, sh) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}

// Small inputs (<= 8192 elements): ONE CTA walks the array in tiles of 4096 with a running carry -- one launch instead
// of three dependent ones.
constexpr int kScan1Threads = 1024;
static __global__ void __launch_bounds__(kScan1Threads) scan_single_cta_kernel(const uint32_t* in, uint32_t* out, int64_t n) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t sh_warp[32];
  __shared__ uint32_t sh_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sh_carry = 0u;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 4 * kScan1Threads) {
    const int64_t i0 = base + 4 * (int64_t)tid;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
    const uint32_t s = v[0] + v[1] + v[2] + v[3];
    uint32_t x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) sh_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = sh_warp[lane];
      uint32_t t = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      sh_warp[lane] = t - w;  // exclusive warp offsets; lane 31's inclusive total is the tile total
    }
    __syncthreads();
    const uint32_t carry = sh_carry;
    uint32_t ex = carry + sh_warp[warp] + (x - s);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) out[i0 + k] = ex;
      ex += v[k];
    }
    __syncthreads();
    if (tid == kScan1Threads - 1) sh_carry = ex;  // the last thread's running value = carry + tile total
    __syncthreads();
  }
}

inline size_t scan_workspace_elems(int64_t n) { return static_cast<size_t>((n + kScanTile - 1) / kScanTile) + 1; }

// out[i] = sum_{j<i} in[j], i in [0,n).  tile_ws: scan_workspace_elems(n) uint32.
inline int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* tile_ws, cudaStream_t st) {
  if (n <= 0) return GR_OK;
  if (n <= 8192) {  // (measured: at 60k elements the serial tiles of one CTA lose to the three-kernel form)
    GR_CHECK_CUDA(launch_pdl(scan_single_cta_kernel, dim3(1), dim3(kScan1Threads), (size_t)0, st, in, out, n));
    GR_CHECK_LAUNCH("scan_single_cta_kernel");
    return GR_OK;
  }
  int tiles = static_cast<int>((n + kScanTile - 1) / kScanTile);
  GR_CHECK_CUDA(launch_pdl(scan_reduce_kernel, dim3(tiles), dim3(kScanThreads), (size_t)0, st, in, tile_ws, n));
  GR_CHECK_LAUNCH("scan_reduce_kernel");
  GR_CHECK_CUDA(launch_pdl(scan_tilesums_kernel, dim3(1), dim3(kScanThreads), (size_t)0, st, tile_ws, tiles));
  GR_CHECK_LAUNCH("scan_tilesums_kernel");
  GR_CHECK_CUDA(launch_pdl(scan_final_kernel, dim3(tiles), dim3(kScanThreads), (size_t)0, st, in, out, tile_ws, n));
  GR_CHECK_LAUNCH("scan_final_kernel");
  return GR_OK;
}

}  // namespace gr
