// Shared helpers for the gaussreg_b200 sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gaussreg_b200.h"

#define GR_STR_(x) #x
#define GR_STR(x) GR_STR_(x)

namespace gr {

// ---- error plumbing -------------------------------------------------------------------------
void set_last_error(const char* what, cudaError_t e);
void count_launch(int n = 1);

#define GR_CHECK_LAUNCH(name)                                 \
  do {                                                        \
    ::gr::count_launch();                                     \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) {                                 \
      ::gr::set_last_error(name, e__);                        \
      return GR_ERR_CUDA;                                     \
    }                                                         \
  } while (0)

#define GR_CHECK_CUDA(expr)                                   \
  do {                                                        \
    cudaError_t e__ = (expr);                                 \
    if (e__ != cudaSuccess) {                                 \
      ::gr::set_last_error(#expr, e__);                       \
      return GR_ERR_CUDA;                                     \
    }                                                         \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember it per (kernel, device) so that a
// process driving several GPUs opts in on each of them (runtime.cu)
cudaError_t ensure_smem_attr(const void* kernel, int bytes);

// optional request to the tensor-core GEMM: also produce GroupNorm partial statistics of the output.  On return
// nblk > 0 is the number of row blocks written to partial[nblk][groups]; nblk == 0: the product ran without them.
struct GnStatsOut {
  double2* partial;
  size_t capacity_blocks;
  int groups;
  int nblk;
};
// gemm.cu: C = act(alpha A op(B) / row_div + bias + residual) with optional packed weights / GroupNorm statistics
int gemm_ex(const float* A, long long lda, const float* B, long long ldb, int trans_b, float* C, long long ldc, int M, int N, int K,
            float alpha, const float* bias, const float* row_div, const float* residual, long long ldr, int act, void* stream,
            const float* B_packed, GnStatsOut* gn, const void* B_packed16 = nullptr, float inv_scale16 = 1.f);
// norm.cu: y = act(GroupNorm(x) [+ add]) from partial statistics (finalize + apply); ws holds `groups` float2
int group_norm_from_partial(const float* x, long long n_rows, int C, int groups, const double2* partial, int nblk,
                            const float* gamma, const float* beta, float eps, const float* add, int act, float* y,
                            float2* stats, void* stream);
// the two halves of group_norm_from_partial, and the two-input apply y = act(GN(x) + GN2(x2))
int group_norm_finalize(const double2* partial, int nblk, long long n_rows, int C, int groups, float eps, float2* stats, void* stream);
int group_norm_apply2(const float* x, const float* x2, long long n_rows, int C, int groups, const float2* stats, const float2* stats2,
                      const float* gamma, const float* beta, const float* gamma2, const float* beta2, int act, float* y, void* stream);
// rows covered by one partial block of the split-K reduction (gemm_tc.cu)
constexpr int kGnReduceRows = 32;

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Most of this library's kernels are short (2-20 us) links of long dependent chains; measured on B200, the step's wall
// time exceeds the sum of its kernel times by ~1.6 us per launch.  A kernel launched with the programmatic-stream-
// serialization attribute may be scheduled while its predecessor is still running: everything before its
// `pdl_wait()` (parameter loads, shared-memory carving, barrier / TMEM set-up) overlaps the predecessor's tail, and
// `pdl_wait()` returns once the predecessor has completed and its memory is visible.  `pdl_trigger()` lets the NEXT
// kernel in the stream start being scheduled.  Rules kept by every kernel that is launched through launch_pdl():
// no global memory access before pdl_wait(); pdl_trigger() right after it.  GAUSSREG_PDL=0 turns the attribute off
// (the two instructions are then no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- workspace carving ----------------------------------------------------------------------
struct Carver {
  char* base;
  size_t off;
  size_t cap;
  bool ok;
  __host__ Carver(void* p, size_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes), ok(true) {}
  template <typename T>
  __host__ T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    if (base != nullptr && off > cap) ok = false;
    return r;
  }
};

// ---- device helpers -------------------------------------------------------------------------
constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// order-preserving float <-> uint mapping (for atomicMin/Max on floats)
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// largest b with off[b] <= i, off has nb+1 ascending entries (off[0] = 0)
__device__ __forceinline__ int find_segment(const int* __restrict__ off, int nb, int i) {
  int lo = 0, hi = nb;  // invariant: off[lo] <= i < off[hi] (caller guarantees i < off[nb])
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace gr
