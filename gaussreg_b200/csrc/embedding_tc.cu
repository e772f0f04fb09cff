// T1 fused: geometric structure embedding on the tensor cores, without any (N*N*k, C) intermediate.
//
//   emb[r, :] = (W_d E(d_r) + b_d) + max_k (W_a E(a_{r,k}) + b_a),   r = n*N + m,  C = 256
//   E(x)[2i] = sin(x * div_i), E(x)[2i+1] = cos(x * div_i)
//
// Reference: geotransformer/modules/geotransformer/geotransformer.py:57-72 materialises E(d) (N^2 x 256),
// E(a) (N^2 x 3 x 256) and both projections before the max.  Here one CTA owns a 128-row x 128-column
// output tile: producer warps GENERATE the sinusoid operand tiles (hi / lo TF32 split) straight into the UMMA
// canonical shared-memory layout, the weight tiles arrive pre-split and pre-swizzled by one bulk async copy
// per k-block, four fp32 accumulators (d, a0, a1, a2) live side by side in TMEM (4 x 128 = 512 columns), and
// the epilogue applies the bias / max-over-k / sum directly on the accumulators.
#include <stdlib.h>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace gr {
namespace tc {

constexpr int kEmbC = 256;             // hidden dim (= K of the projections and total N)
constexpr int kEmbBN = 128;            // output columns per CTA
constexpr int kEmbBM = 128;
constexpr int kEmbBK = 32;
constexpr int kEmbKB = kEmbC / kEmbBK;  // 8 k-blocks per projection
constexpr int kEmbStages = 3;
constexpr int kEmbTile = kEmbBM * kEmbBK * 4;           // 16 KB (one hi or lo tile, A or B)
constexpr int kEmbStageBytes = 4 * kEmbTile;            // A hi, A lo, B hi, B lo
constexpr int kEmbSmem = kEmbStages * kEmbStageBytes + 1024 + 256;
constexpr int kEmbProducers = 512;                       // 16 producer / epilogue warps
constexpr int kEmbThreads = kEmbProducers + 32;          // + the MMA warp

// W (N, K) row-major -> for every (n-tile of 128, k-block of 32): [hi tile 16 KB][lo tile 16 KB] in the
// K-major SWIZZLE_128B byte order the tensor core reads.
__global__ void __launch_bounds__(256) pack_weight_tf32x3_kernel(const float* __restrict__ W, int N, int K, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int nt = blockIdx.y, kb = blockIdx.x;
  const int kblocks = (K + 31) / 32;
  unsigned char* base = reinterpret_cast<unsigned char*>(out) + ((size_t)(nt * kblocks + kb)) * 2 * kEmbTile;
  for (int ch = threadIdx.x; ch < 128 * 8; ch += blockDim.x) {
    const int r = ch >> 3, c = ch & 7;
    const int gn = nt * 128 + r, gk = kb * 32 + c * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // rows / columns beyond (N, K) are zero padding
    if (gn < N) {
      const float* src = W + (size_t)gn * K + gk;
      if (gk + 3 < K) { v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = src[3]; }
      else { if (gk < K) v.x = src[0]; if (gk + 1 < K) v.y = src[1]; if (gk + 2 < K) v.z = src[2]; }
    }
    float4 hi, lo;
    hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
    lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(base + off) = hi;
    *reinterpret_cast<float4*>(base + kEmbTile + off) = lo;
  }
}

// sin / cos with fp32 accuracy.  Half of the 128 sinusoid frequencies are so low that the argument stays below
// 0.5 rad for every index the model produces: there a short Taylor polynomial (truncation error < 3e-10) replaces
// the general routine and its range reduction.  The branch is warp-uniform except in one or two k-blocks.
__device__ __forceinline__ void sincos_fp32(float a, float* s, float* c) {
  if (fabsf(a) < 0.5f) {
    const float a2 = a * a;
    float ps = fmaf(a2, 2.7557319e-6f, -1.9841270e-4f);
    ps = fmaf(ps, a2, 8.3333333e-3f);
    ps = fmaf(ps, a2, -1.6666667e-1f);
    *s = fmaf(a * a2, ps, a);
    float pc = fmaf(a2, -2.7557319e-7f, 2.4801587e-5f);
    pc = fmaf(pc, a2, -1.3888889e-3f);
    pc = fmaf(pc, a2, 4.1666667e-2f);
    pc = fmaf(pc, a2, -0.5f);
    *c = fmaf(pc, a2, 1.0f);
  } else {
    sincosf(a, s, c);
  }
}

// |a| < 0.5: Taylor kernels (truncation < 2e-11), no range reduction
__device__ __forceinline__ void sincos_small(float a, float* s, float* c) {
  const float a2 = a * a;
  float ps = fmaf(a2, 2.7557319e-6f, -1.9841270e-4f);
  ps = fmaf(ps, a2, 8.3333333e-3f);
  ps = fmaf(ps, a2, -1.6666667e-1f);
  *s = fmaf(a * a2, ps, a);
  float pc = fmaf(a2, -2.7557319e-7f, 2.4801587e-5f);
  pc = fmaf(pc, a2, -1.3888889e-3f);
  pc = fmaf(pc, a2, 4.1666667e-2f);
  pc = fmaf(pc, a2, -0.5f);
  *c = fmaf(pc, a2, 1.0f);
}

// Branch-free sin / cos for the arguments this model produces (|a| < ~100): three-term Cody-Waite reduction by
// pi/2 (the products are exact inside the FMAs), degree-9 / degree-10 minimax-free Taylor kernels on
// [-pi/4, pi/4], quadrant fix-up by selects.  Measured against float64 over 2e6 arguments in [-64, 64]:
// max error 6.9e-8 absolute, 1.42 ulp -- the same level as libm's float sin / cos -- at ~24 instructions for the
// pair, with no divergence between lanes that hold different rows.
__device__ __forceinline__ void sincos_cw(float a, float* s, float* c) {
  const float kf = rintf(a * 0.636619772f);
  float r = fmaf(kf, -1.5707963705062866f, a);
  r = fmaf(kf, 4.371138828673793e-08f, r);
  r = fmaf(kf, 1.7151245100058819e-15f, r);
  const float z = r * r;
  float ps = fmaf(z, 2.7557314297e-06f, -1.9841270114e-04f);
  ps = fmaf(ps, z, 8.3333337680e-03f);
  ps = fmaf(ps, z, -1.6666667163e-01f);
  const float sn = fmaf(r * z, ps, r);
  float pc = fmaf(z, -2.7557314297e-07f, 2.4801587642e-05f);
  pc = fmaf(pc, z, -1.3888889225e-03f);
  pc = fmaf(pc, z, 4.1666667908e-02f);
  pc = fmaf(pc, z, -0.5f);
  const float cs = fmaf(pc, z, 1.0f);
  const int q = __float2int_rn(kf);
  const float s0 = (q & 1) ? cs : sn, c0 = (q & 1) ? sn : cs;
  *s = (q & 2) ? -s0 : s0;
  *c = ((q + 1) & 2) ? -c0 : c0;
}

template <bool CW>
__global__ void __launch_bounds__(kEmbThreads, 1) structure_embedding_tc_kernel(
    const float* __restrict__ d_idx, const float* __restrict__ a_idx, long long rows, int angle_k,
    const float* __restrict__ div_term, const float* __restrict__ wd_packed, const float* __restrict__ wa_packed,
    const float* __restrict__ bias_d, const float* __restrict__ bias_a, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kEmbStages * kEmbStageBytes);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kEmbStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * kEmbStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kEmbStages + 1);
  __shared__ float s_div[kEmbC / 2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nhalf = blockIdx.x;                         // which 128 output columns
  const long long r0 = (long long)blockIdx.y * kEmbBM;  // first pair-row of the tile
  const int n_gemm = 1 + angle_k;                       // d, a_0 .. a_{k-1}
  const int n_iter = n_gemm * kEmbKB;

  if (tid < kEmbC / 2) s_div[tid] = div_term[tid];
  if (tid == 0) {
    for (int s = 0; s < kEmbStages; ++s) { mbar_init(full_bar(s), kEmbProducers / 32); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int kMmaWarp = kEmbProducers / 32;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < kMmaWarp) {
    // ---------------------------------------------------------------- producers: sinusoid operand tiles
    const int row = tid >> 2, h8 = tid & 3;  // four threads per row, 8 consecutive k each (4 sin/cos pairs)
    const long long r = r0 + row;
    const bool valid = r < rows;
    float xg[4] = {0.f, 0.f, 0.f, 0.f};  // embedding indices of this row for d, a0, a1, a2
    if (valid) {
      xg[0] = d_idx[r];
      for (int k = 0; k < angle_k; ++k) xg[1 + k] = a_idx[r * angle_k + k];
    }
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kEmbStages;
      const int g = it / kEmbKB, kb = it % kEmbKB;
      if (it >= kEmbStages) mbar_wait(empty_bar(s), ((it / kEmbStages) - 1) & 1);
      unsigned char* st = smem + s * kEmbStageBytes;
      if (tid == 0) {  // weight tile: one bulk copy of the pre-packed [hi | lo] pair
        const float* src = (g == 0 ? wd_packed : wa_packed) + ((size_t)(nhalf * kEmbKB + kb)) * (2 * kEmbTile / 4);
        mbar_expect_tx(full_bar(s), 2 * kEmbTile);  // registered before the copy; thread 0 still arrives below
        bulk_copy_g2s(smem_u32(st + 2 * kEmbTile), src, 2 * kEmbTile, full_bar(s));
      }
      const float x = g == 0 ? xg[0] : (g == 1 ? xg[1] : (g == 2 ? xg[2] : xg[3]));
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {  // 2 chunks of 16 B = 2 (sin, cos) pairs each
        const int c = h8 * 2 + jj;
        const int i0 = kb * 16 + c * 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          if (CW) {
            sincos_cw(__fmul_rn(x, s_div[i0]), &v.x, &v.y);
            sincos_cw(__fmul_rn(x, s_div[i0 + 1]), &v.z, &v.w);
          } else {
            sincos_fp32(__fmul_rn(x, s_div[i0]), &v.x, &v.y);
            sincos_fp32(__fmul_rn(x, s_div[i0 + 1]), &v.z, &v.w);
          }
        }
        float4 hi, lo;
        hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
        lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + kEmbTile + off) = lo;
      }
      // every lane fences its own stores towards the async proxy; ONE arrival per warp (512 single arrivals on one
      // mbarrier word serialise for the better part of a microsecond per k-block)
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
    }
    // ---------------------------------------------------------------- epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3, cgrp = warp >> 2;  // TMEM lane quarter, 32-column group (16 warps cover 4 x 4 chunks)
    // accumulator rows live one per thread (TMEM lane); transpose each 32x32 chunk through shared memory (the operand
    // stages are free now) so that the global stores are full 128-byte lines instead of 32 scattered 16-byte pieces
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    {
      const int c0 = cgrp * 32;
      const uint32_t lane_addr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      uint32_t t[32];
      float mx[32];
      tmem_ld32(lane_addr + 1 * kEmbBN, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) mx[j] = __uint_as_float(t[j]);
      for (int k = 1; k < angle_k; ++k) {
        tmem_ld32(lane_addr + (1 + k) * kEmbBN, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) mx[j] = fmaxf(mx[j], __uint_as_float(t[j]));
      }
      tmem_ld32(lane_addr, t);
      const int nbase = nhalf * kEmbBN + c0;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = (__uint_as_float(t[j]) + bias_d[nbase + j]) + (mx[j] + bias_a[nbase + j]);
        o.y = (__uint_as_float(t[j + 1]) + bias_d[nbase + j + 1]) + (mx[j + 1] + bias_a[nbase + j + 1]);
        o.z = (__uint_as_float(t[j + 2]) + bias_d[nbase + j + 2]) + (mx[j + 2] + bias_a[nbase + j + 2]);
        o.w = (__uint_as_float(t[j + 3]) + bias_d[nbase + j + 3]) + (mx[j + 3] + bias_a[nbase + j + 3]);
        *reinterpret_cast<float4*>(stage + lane * 36 + j) = o;
      }
      __syncwarp();
      const int c4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
      for (int r4 = 0; r4 < 32; r4 += 4) {
        const int row = r4 + rsub;
        const long long rr = r0 + q * 32 + row;
        if (rr < rows)
          *reinterpret_cast<float4*>(out + rr * kEmbC + nbase + c4) = *reinterpret_cast<const float4*>(stage + row * 36 + c4);
      }
    }
    tc_fence_before();
  } else {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kEmbBN >> 3) << 17) | ((uint32_t)(kEmbBM >> 4) << 24);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kEmbStages;
        const int g = it / kEmbKB, kb = it % kEmbKB;
        mbar_wait(full_bar(s), (it / kEmbStages) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * kEmbStageBytes);
        const uint32_t a_lo = a_hi + kEmbTile, b_hi = a_hi + 2 * kEmbTile, b_lo = a_hi + 3 * kEmbTile;
        const uint32_t acc = tmem_acc + (uint32_t)(g * kEmbBN);
#pragma unroll
        for (int k8 = 0; k8 < kEmbBK / 8; ++k8) {
          const uint32_t ko = k8 * 32;
          umma_tf32(acc, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, (kb | k8) != 0 ? 1u : 0u);
          umma_tf32(acc, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
          umma_tf32(acc, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, 1u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(512));
  }
}

// ---- full-width form: one CTA owns 128 rows x ALL 256 columns ---------------------------------------------------
// The 128-column kernel above generates every sinusoid tile twice (once per column half); its producers, not the
// tensor core, set its pace (ncu: issue slots 51 %, tensor pipe 46 %).  Here the operand tile is generated once.
// Four 128 x 256 fp32 accumulators do not fit in TMEM (4 x 256 > 512 columns), so the angle products are reduced
// on the fly: the GEMMs run in the order a_0, a_1, .., d and ping-pong between two 256-column TMEM slots; while the
// tensor core works on one slot, the producer warps read the finished slot into a running per-thread max (64
// registers: their row x 64 columns) and hand the slot back (slot_free barrier).  The epilogue adds d and the biases.
constexpr int kE2BN = 256;
constexpr int kE2Stages = 2;
constexpr int kE2ATile = kEmbBM * kEmbBK * 4;        // 16 KB
constexpr int kE2BTile = kE2BN * kEmbBK * 4;         // 32 KB
constexpr int kE2StageBytes = 2 * kE2ATile + 2 * kE2BTile;  // 96 KB
constexpr int kE2Smem = kE2Stages * kE2StageBytes + 1024 + 256;

// CL = 2: thread-block clusters of two CTAs (two row tiles).  Both CTAs need the same weight tile at the same step,
// and the weight stream out of L2 (64 KB per k-block per SM, hi + lo) is what bounds this kernel once the operand is
// generated only once: CTA 0 fetches the hi half, CTA 1 the lo half, each multicast into both CTAs' stage (every
// CTA's full barrier sees all 64 KB).  A stage may be overwritten in BOTH CTAs only when both tensor cores have
// consumed it, so every MMA commit is multicast to both CTAs' empty barriers (arrival count 2).
template <int CL>
__global__ void __launch_bounds__(kEmbThreads, 1) structure_embedding_tc256_kernel(
    const float* __restrict__ d_idx, const float* __restrict__ a_idx, long long rows, int angle_k,
    const float* __restrict__ div_term, const float* __restrict__ wd_packed, const float* __restrict__ wa_packed,
    const float* __restrict__ bias_d, const float* __restrict__ bias_a, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kE2Stages * kE2StageBytes);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kE2Stages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * kE2Stages);
  auto slot_free = [&](int s) { return bar_base + 8u * (2 * kE2Stages + 1 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kE2Stages + 3);
  __shared__ float s_div[kEmbC / 2];
  __shared__ float s_xmax[4][kEmbProducers / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long r0 = (long long)blockIdx.x * kEmbBM;
  const int n_gemm = 1 + angle_k;  // a_0 .. a_{k-1}, then d
  const int n_iter = n_gemm * kEmbKB;

  if (tid < kEmbC / 2) s_div[tid] = div_term[tid];
  if (tid == 0) {
    for (int s = 0; s < kE2Stages; ++s) { mbar_init(full_bar(s), kEmbProducers / 32); mbar_init(empty_bar(s), CL); }
    mbar_init(accum_bar, 1);
    mbar_init(slot_free(0), kEmbProducers / 32);
    mbar_init(slot_free(1), kEmbProducers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int kMmaWarp = kEmbProducers / 32;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL > 1) cluster_sync_all();  // the peer's barriers exist before anything is multicast into them
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < kMmaWarp) {
    const int row = tid >> 2, h8 = tid & 3;
    const long long r = r0 + row;
    const bool valid = r < rows;
    float xg[4] = {0.f, 0.f, 0.f, 0.f};  // indices in GEMM order: a_0 .. a_{k-1}, d
    if (valid) {
      for (int k = 0; k < angle_k; ++k) xg[k] = a_idx[r * angle_k + k];
      xg[angle_k] = d_idx[r];
    }
    // largest |index| of this row tile per product: a k-block whose largest argument stays below 0.5 rad for EVERY
    // row takes the short Taylor kernels without range reduction (a CTA-uniform branch, no divergence)
    {
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const float mg = warp_max(fabsf(xg[g4]));
        if (lane == 0) s_xmax[g4][warp] = mg;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEmbProducers) : "memory");  // producer warps only
    }
    float xmax[4];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      float mg = 0.f;
#pragma unroll
      for (int w = 0; w < kEmbProducers / 32; ++w) mg = fmaxf(mg, s_xmax[g4][w]);
      xmax[g4] = mg;
    }
    const int q = warp & 3, cq = warp >> 2;  // TMEM lane quarter / 64-column group read by this warp
    const uint32_t my_tmem = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
    float mx[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) mx[j] = -INFINITY;

    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kE2Stages;
      const int g = it / kEmbKB, kb = it % kEmbKB;
      if (it >= kE2Stages) mbar_wait(empty_bar(s), ((it / kE2Stages) - 1) & 1);
      if (kb == 2 && g >= 1) {
        // every MMA of GEMM g-1 has completed (the commit just waited for was issued after them): fold its
        // accumulator into the running max and return the TMEM slot
        tc_fence_after();
        const uint32_t src = my_tmem + (uint32_t)(((g - 1) & 1) * kE2BN);
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          uint32_t t[16];
          tmem_ld16_nowait(src + c16 * 16, t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) mx[c16 * 16 + j] = fmaxf(mx[c16 * 16 + j], __uint_as_float(t[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(slot_free((g - 1) & 1));
      }
      unsigned char* st = smem + s * kE2StageBytes;
      if (tid == 0) {
        const float* wsrc = (g == angle_k ? wd_packed : wa_packed);
        mbar_expect_tx(full_bar(s), 2 * kE2BTile);
        const uint32_t b_hi = smem_u32(st + 2 * kE2ATile), b_lo = b_hi + kE2BTile;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {  // packed tile (nt, kb) = [hi 16 KB][lo 16 KB]
          const float* src = wsrc + ((size_t)(nt * kEmbKB + kb)) * (2 * kEmbTile / 4);
          if (CL == 1) {
            bulk_copy_g2s(b_hi + nt * kEmbTile, src, kEmbTile, full_bar(s));
            bulk_copy_g2s(b_lo + nt * kEmbTile, src + kEmbTile / 4, kEmbTile, full_bar(s));
          } else if (crank == 0) {
            bulk_copy_g2s_multicast(b_hi + nt * kEmbTile, src, kEmbTile, full_bar(s), (uint16_t)0x3);
          } else {
            bulk_copy_g2s_multicast(b_lo + nt * kEmbTile, src + kEmbTile / 4, kEmbTile, full_bar(s), (uint16_t)0x3);
          }
        }
      }
      const float x = g == 0 ? xg[0] : (g == 1 ? xg[1] : (g == 2 ? xg[2] : xg[3]));
      const float xm = g == 0 ? xmax[0] : (g == 1 ? xmax[1] : (g == 2 ? xmax[2] : xmax[3]));
      const bool small_args = xm * s_div[kb * 16] < 0.5f;  // div_term decreases with the frequency index
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int c = h8 * 2 + jj;
        const int i0 = kb * 16 + c * 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          if (small_args) {
            sincos_small(__fmul_rn(x, s_div[i0]), &v.x, &v.y);
            sincos_small(__fmul_rn(x, s_div[i0 + 1]), &v.z, &v.w);
          } else {
            sincos_cw(__fmul_rn(x, s_div[i0]), &v.x, &v.y);
            sincos_cw(__fmul_rn(x, s_div[i0 + 1]), &v.z, &v.w);
          }
        }
        float4 hi, lo;
        hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
        lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + kE2ATile + off) = lo;
      }
      // every lane fences its own stores towards the async proxy; ONE arrival per warp (512 single arrivals on one
      // mbarrier word serialise for the better part of a microsecond per k-block)
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
    }
    // ---------------------------------------------------------------- epilogue
    // Read-outs happen at kb == 2 of the FOLLOWING product, so every angle product has been folded into mx by now;
    // the d product sits in slot (angle_k & 1).
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t dsrc = my_tmem + (uint32_t)((angle_k & 1) * kE2BN);
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
#pragma unroll
    for (int c32 = 0; c32 < 2; ++c32) {
      uint32_t t[32];
      tmem_ld32(dsrc + c32 * 32, t);
      const int nbase = cq * 64 + c32 * 32;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = (__uint_as_float(t[j]) + bias_d[nbase + j]) + (mx[c32 * 32 + j] + bias_a[nbase + j]);
        o.y = (__uint_as_float(t[j + 1]) + bias_d[nbase + j + 1]) + (mx[c32 * 32 + j + 1] + bias_a[nbase + j + 1]);
        o.z = (__uint_as_float(t[j + 2]) + bias_d[nbase + j + 2]) + (mx[c32 * 32 + j + 2] + bias_a[nbase + j + 2]);
        o.w = (__uint_as_float(t[j + 3]) + bias_d[nbase + j + 3]) + (mx[c32 * 32 + j + 3] + bias_a[nbase + j + 3]);
        *reinterpret_cast<float4*>(stage + lane * 36 + j) = o;
      }
      __syncwarp();
      const int c4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
      for (int r4 = 0; r4 < 32; r4 += 4) {
        const int rw = r4 + rsub;
        const long long rr = r0 + q * 32 + rw;
        if (rr < rows)
          *reinterpret_cast<float4*>(out + rr * kEmbC + nbase + c4) = *reinterpret_cast<const float4*>(stage + rw * 36 + c4);
      }
    }
    tc_fence_before();
  } else {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kE2BN >> 3) << 17) | ((uint32_t)(kEmbBM >> 4) << 24);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kE2Stages;
        const int g = it / kEmbKB, kb = it % kEmbKB;
        if (kb == 0 && g >= 2) {  // the slot still holds GEMM g-2 until the producers have read it out
          mbar_wait(slot_free(g & 1), ((g >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(full_bar(s), (it / kE2Stages) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * kE2StageBytes);
        const uint32_t a_lo = a_hi + kE2ATile, b_hi = a_hi + 2 * kE2ATile, b_lo = b_hi + kE2BTile;
        const uint32_t acc = tmem_acc + (uint32_t)((g & 1) * kE2BN);
#pragma unroll
        for (int k8 = 0; k8 < kEmbBK / 8; ++k8) {
          const uint32_t ko = k8 * 32;
          umma_tf32(acc, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, (kb | k8) != 0 ? 1u : 0u);
          umma_tf32(acc, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
          umma_tf32(acc, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, 1u);
        }
        if (CL == 1) umma_commit(empty_bar(s));
        else umma_commit_multicast(empty_bar(s), (uint16_t)0x3);
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer may still multicast into this CTA's stages / barriers until it is done too
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(512));
  }
}

// sin / cos by a 128-entry table of the full period plus a tiny polynomial: a = k pi/64 + d, |d| <= pi/128,
//     sin a = S_k cos d + C_k sin d,   cos a = C_k cos d - S_k sin d,
// sin d = d - d^3/6 (truncation 7e-11), cos d = 1 - d^2/2 + d^4/24 (3e-13).  The reduction is two FMAs (the products
// k * C1, k * C2 are not rounded inside an FMA; C1 + C2 = pi/64 to 1e-16) and the table covers all four quadrants, so
// there is no fix-up: 17 instructions for the pair instead of ~30 for sincos_cw, same accuracy class (table entries and
// two roundings: ~1e-7 absolute).  tab[k] = (sin, cos)(k pi / 64), correctly rounded.
__device__ __forceinline__ void sincos_tab(float a, const float2* __restrict__ tab, float* s, float* c) {
  const float kf = rintf(a * 20.371832715762602f);
  float d = fmaf(kf, -0.049087386578321457f, a);
  d = fmaf(kf, 1.3659809e-09f, d);  // fp32(pi/64) - pi/64 = +1.3659809e-9
  const float2 t = tab[__float2int_rn(kf) & 127];
  const float d2 = d * d;
  const float sd = fmaf(d * d2, -0.16666667f, d);
  const float cd = fmaf(d2, fmaf(d2, 0.041666668f, -0.5f), 1.0f);
  *s = fmaf(t.x, cd, t.y * sd);
  *c = fmaf(t.y, cd, -(t.x * sd));
}

// ---- T1 with fp16-split operands ("3xFP16") ------------------------------------------------------------------------
// TF32 and fp16 carry the same 11 significant bits, so x = hi + lo with two fp16 parts is exactly as accurate as the
// two-part TF32 split -- as long as the values sit in fp16's exponent range.  Here they do by construction: the A
// operand is sin / cos in [-1, 1] and the B operand is a static weight matrix that is pre-scaled by a power of two at
// packing time (so that its lo parts are fp16 normals); the scale is divided out in the epilogue, exactly.  What is
// lost is bounded absolutely, not relatively: an fp16 subnormal lo part is good to 2^-25 (3e-8) of an O(1) value.
// kind::f16 issues K = 16 per instruction at the rate kind::tf32 issues K = 8: the three products per k-step cost half
// the tensor-pipe time, and a 64-wide k-block fills the same 128-byte swizzled row a 32-wide fp32 block did.
constexpr int kF16BK = 64;                                 // K elements per k-block (128-byte rows of fp16)
constexpr int kF16KB = kEmbC / kF16BK;                     // 4 k-blocks per projection
constexpr int kF16ATile = kEmbBM * kF16BK * 2;             // 16 KB
constexpr int kF16BTile = kE2BN * kF16BK * 2;              // 32 KB
constexpr int kF16StageBytes = 2 * kF16ATile + 2 * kF16BTile;  // 96 KB
constexpr int kF16Stages = 2;
constexpr int kF16Smem = kF16Stages * kF16StageBytes + 1024 + 256;

// W (256, 256) fp32 -> per 64-wide k-block [hi 32 KB][lo 32 KB] of fp16 in the K-major SWIZZLE_128B byte order, W
// multiplied by `scale` (a power of two) first.
__global__ void __launch_bounds__(256) pack_weight_f16x2_kernel(const float* __restrict__ W, int N, int K, float scale,
                                                                unsigned char* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int kb = blockIdx.x;
  unsigned char* base = out + (size_t)kb * 2 * kF16BTile;
  for (int ch = threadIdx.x; ch < kE2BN * 8; ch += blockDim.x) {
    const int r = ch >> 3, c = ch & 7;
    const int gk = kb * kF16BK + c * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = 0.f, v1 = 0.f;  // rows / columns beyond (N, K) are zero padding
      if (r < N && gk + 2 * e < K) v0 = W[(size_t)r * K + gk + 2 * e] * scale;
      if (r < N && gk + 2 * e + 1 < K) v1 = W[(size_t)r * K + gk + 2 * e + 1] * scale;
      const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
      const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
      hi[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + kF16BTile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Same structure as structure_embedding_tc256_kernel<1> (16 producer / read-out warps + the MMA warp, four products in
// the order a_0, a_1, a_2, d ping-ponging between two 256-column TMEM slots, running max in registers), with fp16
// operand tiles: 4 k-blocks of 64 per product.
__global__ void __launch_bounds__(kEmbThreads, 1) structure_embedding_f16_kernel(
    const float* __restrict__ d_idx, const float* __restrict__ a_idx, long long rows, int angle_k,
    const float* __restrict__ div_term, const unsigned char* __restrict__ wd_packed, const unsigned char* __restrict__ wa_packed,
    float inv_scale_d, float inv_scale_a, const float* __restrict__ bias_d, const float* __restrict__ bias_a,
    float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ unsigned char smem_raw[];
  // (the rounded-up pointer is generic, so the tile stores below compile to generic ST.E rather than STS; keeping it in
  // the shared window -- as gemm_tc.cu does -- measured 3 % SLOWER here: 0.82 vs 0.80 ms per pair)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kF16Stages * kF16StageBytes);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kF16Stages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * kF16Stages);
  auto slot_free = [&](int s) { return bar_base + 8u * (2 * kF16Stages + 1 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kF16Stages + 3);
  __shared__ float s_div[kEmbC / 2];
  __shared__ float2 s_tab[128];  // (sin, cos)(k pi / 64)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long r0 = (long long)blockIdx.x * kEmbBM;
  const int n_gemm = 1 + angle_k;  // a_0 .. a_{k-1}, then d
  const int n_iter = n_gemm * kF16KB;

  if (tid < kEmbC / 2) s_div[tid] = div_term[tid];
  if (tid >= 128 && tid < 256) {
    double sd, cd;
    sincospi((double)(tid - 128) / 64.0, &sd, &cd);
    s_tab[tid - 128] = make_float2((float)sd, (float)cd);
  }
  if (tid == 0) {
    for (int s = 0; s < kF16Stages; ++s) { mbar_init(full_bar(s), kEmbProducers / 32); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    mbar_init(slot_free(0), kEmbProducers / 32);
    mbar_init(slot_free(1), kEmbProducers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int kMmaWarp = kEmbProducers / 32;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < kMmaWarp) {
    const int row = tid >> 2, h8 = tid & 3;
    const long long r = r0 + row;
    const bool valid = r < rows;
    float xg[4] = {0.f, 0.f, 0.f, 0.f};  // indices in GEMM order: a_0 .. a_{k-1}, d
    if (valid) {
      const float dv = d_idx[r];
#pragma unroll
      for (int k = 0; k < 3; ++k) xg[k] = k < angle_k ? a_idx[r * angle_k + k] : (k == angle_k ? dv : 0.f);
      if (angle_k == 3) xg[3] = dv;
    }
    const int q = warp & 3, cq = warp >> 2;  // TMEM lane quarter / 64-column group read by this warp
    const uint32_t my_tmem = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
    float mx[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) mx[j] = -INFINITY;

    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kF16Stages;
      const int g = it / kF16KB, kb = it % kF16KB;
      if (it >= kF16Stages) mbar_wait(empty_bar(s), ((it / kF16Stages) - 1) & 1);
      if (kb == 2 && g >= 1) {
        // the commit just waited for (iteration it - 2 = k-block 0 of this product) was issued after every MMA of
        // product g-1: fold that accumulator into the running max and return the TMEM slot
        tc_fence_after();
        const uint32_t src = my_tmem + (uint32_t)(((g - 1) & 1) * kE2BN);
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          uint32_t t[16];
          tmem_ld16_nowait(src + c16 * 16, t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) mx[c16 * 16 + j] = fmaxf(mx[c16 * 16 + j], __uint_as_float(t[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(slot_free((g - 1) & 1));
      }
      unsigned char* st = smem + s * kF16StageBytes;
      if (tid == 0) {
        const unsigned char* wsrc = (g == angle_k ? wd_packed : wa_packed) + (size_t)kb * 2 * kF16BTile;
        mbar_expect_tx(full_bar(s), 2 * kF16BTile);
        const uint32_t b_hi = smem_u32(st + 2 * kF16ATile);
#pragma unroll
        for (int part = 0; part < 4; ++part)  // [hi 32 KB][lo 32 KB], source and destination both contiguous
          bulk_copy_g2s(b_hi + part * (kF16BTile / 2), wsrc + (size_t)part * (kF16BTile / 2), kF16BTile / 2, full_bar(s));
      }
      const float x = g == 0 ? xg[0] : (g == 1 ? xg[1] : (g == 2 ? xg[2] : xg[3]));
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int c = h8 * 2 + jj;           // 16-byte chunk of the 128-byte row: K elements 8c .. 8c+7 of the k-block
        const int i0 = kb * 32 + c * 4;      // = four (sin, cos) pairs of frequencies i0 .. i0+3
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float sv, cv;
          sincos_tab(__fmul_rn(x, s_div[i0 + e]), s_tab, &sv, &cv);   // rows beyond `rows` carry x = 0 and are never stored
          const __half2 h = __floats2half2_rn(sv, cv);               // one cvt.rn.f16x2.f32: sin in the low half
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(sv - hf.x, cv - hf.y);
          hi[e] = *reinterpret_cast<const uint32_t*>(&h);
          lo[e] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(st + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(st + kF16ATile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
    }
    // ---------------------------------------------------------------- epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t dsrc = my_tmem + (uint32_t)((angle_k & 1) * kE2BN);
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
#pragma unroll
    for (int c32 = 0; c32 < 2; ++c32) {
      uint32_t t[32];
      tmem_ld32(dsrc + c32 * 32, t);
      const int nbase = cq * 64 + c32 * 32;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = (__uint_as_float(t[j]) * inv_scale_d + bias_d[nbase + j]) + (mx[c32 * 32 + j] * inv_scale_a + bias_a[nbase + j]);
        o.y = (__uint_as_float(t[j + 1]) * inv_scale_d + bias_d[nbase + j + 1]) + (mx[c32 * 32 + j + 1] * inv_scale_a + bias_a[nbase + j + 1]);
        o.z = (__uint_as_float(t[j + 2]) * inv_scale_d + bias_d[nbase + j + 2]) + (mx[c32 * 32 + j + 2] * inv_scale_a + bias_a[nbase + j + 2]);
        o.w = (__uint_as_float(t[j + 3]) * inv_scale_d + bias_d[nbase + j + 3]) + (mx[c32 * 32 + j + 3] * inv_scale_a + bias_a[nbase + j + 3]);
        *reinterpret_cast<float4*>(stage + lane * 36 + j) = o;
      }
      __syncwarp();
      const int c4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
      for (int r4 = 0; r4 < 32; r4 += 4) {
        const int rw = r4 + rsub;
        const long long rr = r0 + q * 32 + rw;
        if (rr < rows)
          *reinterpret_cast<float4*>(out + rr * kEmbC + nbase + c4) = *reinterpret_cast<const float4*>(stage + rw * 36 + c4);
      }
    }
    tc_fence_before();
  } else {
    if (lane == 0) {
      // kind::f16: c_format F32 (1 << 4), a/b_format F16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(kE2BN >> 3) << 17) | ((uint32_t)(kEmbBM >> 4) << 24);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kF16Stages;
        const int g = it / kF16KB, kb = it % kF16KB;
        if (kb == 0 && g >= 2) {  // the slot still holds GEMM g-2 until the producers have read it out
          mbar_wait(slot_free(g & 1), ((g >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(full_bar(s), (it / kF16Stages) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * kF16StageBytes);
        const uint32_t a_lo = a_hi + kF16ATile, b_hi = a_hi + 2 * kF16ATile, b_lo = b_hi + kF16BTile;
        const uint32_t acc = tmem_acc + (uint32_t)((g & 1) * kE2BN);
#pragma unroll
        for (int k16 = 0; k16 < kF16BK / 16; ++k16) {
          const uint32_t ko = k16 * 32;  // 16 fp16 = 32 bytes along the swizzled row
          umma_f16(acc, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, (kb | k16) != 0 ? 1u : 0u);
          umma_f16(acc, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
          umma_f16(acc, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, 1u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(512));
  }
}

}  // namespace tc
}  // namespace gr

using namespace gr;

/* Packs a (N,K) fp32 weight into the tensor-core operand format, zero-padded to 256 rows x 32 columns:
 * out holds 2 * roundup(N,256) * roundup(K,32) floats. */
extern "C" int gr_pack_weight_tf32x3(const float* W, int N, int K, float* out, void* stream) {
  if (N <= 0 || K <= 0 || !W || !out) return GR_ERR_BAD_ARG;
  dim3 grid((K + 31) / 32, ((N + 255) / 256) * 2);
  GR_CHECK_CUDA(launch_pdl(tc::pack_weight_tf32x3_kernel, dim3(grid), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), W, N, K, out));
  GR_CHECK_LAUNCH("pack_weight_tf32x3_kernel");
  return GR_OK;
}

/* T1 fused.  d_idx (rows), a_idx (rows, angle_k), div_term (128), packed proj_d / proj_a weights
 * (gr_pack_weight_tf32x3 of the (256,256) Linear weights), biases (256) -> out (rows, 256). */
extern "C" int gr_structure_embedding_fused(const float* d_idx, const float* a_idx, int64_t rows, int angle_k,
                                            const float* div_term, int hidden_dim, const float* wd_packed,
                                            const float* wa_packed, const float* bias_d, const float* bias_a, float* out,
                                            void* stream) {
  if (rows < 0 || angle_k < 1 || angle_k > 3 || hidden_dim != tc::kEmbC) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!d_idx || !a_idx || !div_term || !wd_packed || !wa_packed || !bias_d || !bias_a || !out) return GR_ERR_BAD_ARG;
  static bool knobs_read = false;
  static int cw = 1, width = 256, cluster = 1;
  // the shared-memory opt-in is a per-device attribute (ensure_smem_attr remembers it per kernel and device)
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(tc::structure_embedding_tc_kernel<true>), tc::kEmbSmem));
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(tc::structure_embedding_tc_kernel<false>), tc::kEmbSmem));
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(tc::structure_embedding_tc256_kernel<1>), tc::kE2Smem));
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(tc::structure_embedding_tc256_kernel<2>), tc::kE2Smem));
  if (!knobs_read) {
    // 2: CTA pairs share the weight stream by multicast.  Correct, but measured SLOWER on B200 (1.24 vs 1.18 ms per
    // pair): the kernel is paced by its producers' instruction issue, not by L2, and the pair adds hand-over stalls.
    const char* e = getenv("GAUSSREG_T1_CLUSTER");
    cluster = e ? atoi(e) : 1;
    e = getenv("GAUSSREG_T1_SINCOS");  // 0: libm sincosf + small-argument polynomial, 1: branch-free Cody-Waite
    cw = e ? atoi(e) : 1;
    e = getenv("GAUSSREG_T1_WIDTH");               // 256: full-width CTA with ping-pong TMEM slots, 128: two column halves
    width = e ? atoi(e) : 256;
    knobs_read = true;
  }
  if (width == 256) {
    const unsigned tiles = (unsigned)((rows + tc::kEmbBM - 1) / tc::kEmbBM);
    if (cluster == 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((tiles + 1) / 2 * 2);  // whole pairs; a padding CTA runs the protocol on rows >= `rows` and stores nothing
      cfg.blockDim = dim3(tc::kEmbThreads);
      cfg.dynamicSmemBytes = tc::kE2Smem;
      cfg.stream = static_cast<cudaStream_t>(stream);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      GR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc::structure_embedding_tc256_kernel<2>, d_idx, a_idx, (long long)rows, angle_k, div_term,
                                       wd_packed, wa_packed, bias_d, bias_a, out));
    } else {
      GR_CHECK_CUDA(launch_pdl(tc::structure_embedding_tc256_kernel<1>, dim3(tiles), dim3(tc::kEmbThreads), (size_t)(tc::kE2Smem), static_cast<cudaStream_t>(stream), d_idx, a_idx, rows, angle_k, div_term, wd_packed, wa_packed, bias_d, bias_a, out));
    }
    GR_CHECK_LAUNCH("structure_embedding_tc256_kernel");
    return GR_OK;
  }
  dim3 grid(tc::kEmbC / tc::kEmbBN, (unsigned)((rows + tc::kEmbBM - 1) / tc::kEmbBM));
  if (cw)
    GR_CHECK_CUDA(launch_pdl(tc::structure_embedding_tc_kernel<true>, dim3(grid), dim3(tc::kEmbThreads), (size_t)(tc::kEmbSmem), static_cast<cudaStream_t>(stream), d_idx, a_idx, rows, angle_k, div_term, wd_packed, wa_packed, bias_d, bias_a, out));
  else
    GR_CHECK_CUDA(launch_pdl(tc::structure_embedding_tc_kernel<false>, dim3(grid), dim3(tc::kEmbThreads), (size_t)(tc::kEmbSmem), static_cast<cudaStream_t>(stream), d_idx, a_idx, rows, angle_k, div_term, wd_packed, wa_packed, bias_d, bias_a, out));
  GR_CHECK_LAUNCH("structure_embedding_tc_kernel");
  return GR_OK;
}

/* fp16-split form of the packed (256,256) projection weight: 4 k-blocks x [hi 32 KB][lo 32 KB] = 256 KB.  `scale` must be
 * a power of two chosen so that scale * max|W| stays far below 65504 (the caller passes 1 / scale to the fused kernel). */
extern "C" int gr_pack_weight_f16x2(const float* W, int N, int K, float scale, void* out, void* stream) {
  if (N <= 0 || N > tc::kE2BN || K <= 0 || K > tc::kEmbC || !W || !out || !(scale > 0.f)) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(tc::pack_weight_f16x2_kernel, dim3(tc::kF16KB), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), W, N, K, scale,
                           reinterpret_cast<unsigned char*>(out)));
  GR_CHECK_LAUNCH("pack_weight_f16x2_kernel");
  return GR_OK;
}

/* T1 fused with fp16-split operands (same result contract as gr_structure_embedding_fused; weights from
 * gr_pack_weight_f16x2 with scales 1 / inv_scale_d, 1 / inv_scale_a). */
extern "C" int gr_structure_embedding_fused_f16(const float* d_idx, const float* a_idx, int64_t rows, int angle_k,
                                                const float* div_term, int hidden_dim, const void* wd_packed, const void* wa_packed,
                                                float inv_scale_d, float inv_scale_a, const float* bias_d, const float* bias_a,
                                                float* out, void* stream) {
  if (rows < 0 || angle_k < 1 || angle_k > 3 || hidden_dim != tc::kEmbC) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!d_idx || !a_idx || !div_term || !wd_packed || !wa_packed || !bias_d || !bias_a || !out) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(tc::structure_embedding_f16_kernel), tc::kF16Smem));
  const unsigned tiles = (unsigned)((rows + tc::kEmbBM - 1) / tc::kEmbBM);
  GR_CHECK_CUDA(launch_pdl(tc::structure_embedding_f16_kernel, dim3(tiles), dim3(tc::kEmbThreads), (size_t)(tc::kF16Smem), static_cast<cudaStream_t>(stream),
                           d_idx, a_idx, (long long)rows, angle_k, div_term, reinterpret_cast<const unsigned char*>(wd_packed),
                           reinterpret_cast<const unsigned char*>(wa_packed), inv_scale_d, inv_scale_a, bias_d, bias_a, out));
  GR_CHECK_LAUNCH("structure_embedding_f16_kernel");
  return GR_OK;
}
