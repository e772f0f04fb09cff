// K1: KPConv neighbour gather + kernel-point correlation.
//
// Reference: geotransformer/modules/kpconv/kpconv.py:79-122.  For every query m and kernel point k
//     A[m, k, c] = sum_h max(0, 1 - |(s_h - q_m) - kp_k| / sigma) * feat[idx[m,h], c]
// followed by the dense contraction  out[m, :] = (A[m, :] . W[(k,c), :]) / max(1, #{h : sum_c feat > 0}) + bias
// which runs in the GEMM (gemm.cu, "NN" form, row_div + bias epilogue).
//
// The shadow neighbour (idx == Ns, a point at +1e6) has zero influence and zero features, so it is
// skipped rather than gathered.
#include <stdlib.h>

#include "common.cuh"

namespace gr {

constexpr int kKP = 15;

// flag[j] = (sum_c feat[j, c] > 0)   (kpconv.py:113-114); one warp per row
__global__ void __launch_bounds__(256) row_positive_kernel(const float* __restrict__ feats, int n, int C,
                                                           unsigned char* __restrict__ flag) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += feats[(long long)row * C + c];
  s = warp_sum(s);
  if (lane == 0) flag[row] = s > 0.f ? 1 : 0;
}

// GROUP lanes cooperate on one query, each lane owns CPL consecutive channels per pass.
// The gather is what bounds this kernel (one L2 round trip per neighbour row), so the neighbour loop is unrolled by
// four with all four row loads issued before the FMAs; rows are padded to a multiple of four with zero-weight
// entries (the shadow neighbour contributes exactly 0 in the reference as well), weights are read as float4.
template <int GROUP, int CPL>
__global__ void __launch_bounds__(128) kpconv_aggregate_kernel(
    const float* __restrict__ feats, int C, const float* __restrict__ q_pts, const float* __restrict__ s_pts,
    const long long* __restrict__ idx, int H, long long ldi, int M, int Ns, const float* __restrict__ kp, float sigma,
    const unsigned char* __restrict__ pos_flag, float* __restrict__ A, float* __restrict__ row_div) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float smem[];
  constexpr int QPW = 32 / GROUP;  // queries per warp
  const int Hp = (H + 3) & ~3;
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / GROUP, gl = lane % GROUP;
  const int slot = warp * QPW + sub;
  float* w = smem + (size_t)slot * Hp * 17;                           // [Hp][16] weights (k = 15 is padding)
  int* nidx = reinterpret_cast<int*>(w + (size_t)Hp * 16);            // [Hp] neighbour index (or -1)
  const int m = (blockIdx.x * warps + warp) * QPW + sub;
  const bool active = m < M;

  float kx[kKP], ky[kKP], kz[kKP];
#pragma unroll
  for (int k = 0; k < kKP; ++k) { kx[k] = kp[3 * k]; ky[k] = kp[3 * k + 1]; kz[k] = kp[3 * k + 2]; }

  int cnt = 0;
  if (active) {
    const float qx = q_pts[3 * m], qy = q_pts[3 * m + 1], qz = q_pts[3 * m + 2];
    for (int h = gl; h < Hp; h += GROUP) {
      const long long j = h < H ? idx[(long long)m * ldi + h] : -1;
      float4* w4 = reinterpret_cast<float4*>(w + h * 16);
      if (j >= Ns || j < 0) {
        nidx[h] = -1;
        w4[0] = w4[1] = w4[2] = w4[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      nidx[h] = (int)j;
      cnt += pos_flag[j];
      // (s - q) - kp, squared norm, each op rounded separately like the ATen elementwise chain
      const float dx = __fsub_rn(s_pts[3 * j], qx), dy = __fsub_rn(s_pts[3 * j + 1], qy), dz = __fsub_rn(s_pts[3 * j + 2], qz);
      float wk[16];
#pragma unroll
      for (int k = 0; k < kKP; ++k) {
        const float ex = __fsub_rn(dx, kx[k]), ey = __fsub_rn(dy, ky[k]), ez = __fsub_rn(dz, kz[k]);
        const float sq = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
        wk[k] = fmaxf(__fsub_rn(1.0f, __fdiv_rn(sqrtf(sq), sigma)), 0.0f);
      }
      wk[15] = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) w4[k4] = make_float4(wk[4 * k4], wk[4 * k4 + 1], wk[4 * k4 + 2], wk[4 * k4 + 3]);
    }
  }
  // neighbour count of the group
#pragma unroll
  for (int o = GROUP / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  __syncwarp();
  if (active && gl == 0) row_div[m] = (float)max(cnt, 1);

  if (!active) return;
  const int passes = C / (GROUP * CPL);
  for (int ps = 0; ps < passes; ++ps) {
    const int c0 = ps * GROUP * CPL + gl * CPL;
    float acc[kKP][CPL];
#pragma unroll
    for (int k = 0; k < kKP; ++k)
#pragma unroll
      for (int v = 0; v < CPL; ++v) acc[k][v] = 0.f;
    for (int h0 = 0; h0 < Hp; h0 += 4) {
      float f[4][CPL];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = nidx[h0 + u];
#pragma unroll
        for (int v = 0; v < CPL; ++v) f[u][v] = 0.f;
        if (j >= 0) {
          const float* src = feats + (long long)j * C + c0;
          if constexpr (CPL == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src));
            f[u][0] = t.x; f[u][1] = t.y; f[u][2] = t.z; f[u][3] = t.w;
          } else if constexpr (CPL == 2) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(src));
            f[u][0] = t.x; f[u][1] = t.y;
          } else {
#pragma unroll
            for (int v = 0; v < CPL; ++v) f[u][v] = __ldg(src + v);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4* w4 = reinterpret_cast<const float4*>(w + (h0 + u) * 16);
        const float4 wa = w4[0], wb = w4[1], wc = w4[2], wd = w4[3];
        const float wk[kKP] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w, wd.x, wd.y, wd.z};
#pragma unroll
        for (int k = 0; k < kKP; ++k)
#pragma unroll
          for (int v = 0; v < CPL; ++v) acc[k][v] = fmaf(wk[k], f[u][v], acc[k][v]);
      }
    }
    float* out = A + (long long)m * kKP * C + c0;
#pragma unroll
    for (int k = 0; k < kKP; ++k) {
      if constexpr (CPL == 4) {
        *reinterpret_cast<float4*>(out + (long long)k * C) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
      } else if constexpr (CPL == 2) {
        *reinterpret_cast<float2*>(out + (long long)k * C) = make_float2(acc[k][0], acc[k][1]);
      } else {
#pragma unroll
        for (int v = 0; v < CPL; ++v) out[(long long)k * C + v] = acc[k][v];
      }
    }
  }
}


// ---- tensor-core form of the aggregation (C % 32 == 0) -----------------------------------------------------------
// Per query the aggregation is a small dense product  D (16 x C) = Wgt (16 x H) . F (H x C)  (kernel points x
// neighbours x channels; row 15 of Wgt is padding).  One warp owns one query and issues it as mma.sync m16n8k8 TF32
// products with the 3xTF32 split (lo*hi + hi*lo + hi*hi, fp32 accumulate), i.e. 3 warp instructions per 16x8x8 block
// where the scalar form needed 32 FFMA.  The influence weights are evaluated DIRECTLY in the A-fragment layout
// (lane (g,t) owns kernel points g, g+8 and neighbours 8s+t, 8s+t+4 of step s), so they never touch shared memory
// and are computed exactly once per (query, neighbour, kernel point).  The B fragments are gathered straight from
// the feature rows: the four n-tiles of a 32-channel block are interleaved (column n of tile i <-> channel 4n+i) so
// that one 16-byte load per neighbour feeds all four tiles and the eight lanes of a neighbour read one full 128-byte
// line; the same interleave makes every lane's results 8 consecutive channels of A.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 3xTF32 split without cvt.rna (which ptxas expands into four instructions on sm_100a): the tensor core reads only
// the upper 19 bits of an fp32 word, so the raw word already acts as the TRUNCATED hi part; lo = x - trunc(x) is exact
// in fp32 and is itself truncated to TF32 by the hardware (relative error of the pair < 2^-20).
__device__ __forceinline__ uint32_t tf32_hi_bits(float x) { return __float_as_uint(x) & 0xffffe000u; }
// max(0, 1 - |d - kp| / sigma); sqrt.approx (<= 1 ulp) and a multiply by 1/sigma instead of the IEEE sqrt/divide
// chain: the weights differ from the ATen chain by ~1e-7 absolute, far inside the 1e-5 rel-L2 budget of K1
__device__ __forceinline__ float kp_influence(float dx, float dy, float dz, float kx, float ky, float kz, float inv_sigma) {
  const float ex = dx - kx, ey = dy - ky, ez = dz - kz;
  const float sq = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(sq));
  return fmaxf(fmaf(-d, inv_sigma, 1.0f), 0.0f);
}

struct AggQuery {
  float qx, qy, qz, k0x, k0y, k0z, k1x, k1y, k1z, inv_sigma;
  bool pad1;  // lane's second kernel point is the padding row 15
};

// neighbour slots 8s+t and 8s+t+4 of a query row: support indices, -1 for the shadow point / padding
__device__ __forceinline__ void agg_load_idx(const long long* __restrict__ idx_row, int H, int Ns, int s, int t, int& j0, int& j1) {
  const int h0 = 8 * s + t, h1 = h0 + 4;
  const long long i0 = h0 < H ? __ldg(idx_row + h0) : -1, i1 = h1 < H ? __ldg(idx_row + h1) : -1;
  j0 = (i0 >= 0 && i0 < Ns) ? (int)i0 : -1;
  j1 = (i1 >= 0 && i1 < Ns) ? (int)i1 : -1;
}

// A fragments (hi, lo) of one neighbour step
__device__ __forceinline__ void agg_step_weights(const AggQuery& Q, const float* __restrict__ s_pts, int j0, int j1,
                                                 uint32_t (&ahi)[4], uint32_t (&alo)[4]) {
  float w[4] = {0.f, 0.f, 0.f, 0.f};
  if (j0 >= 0) {
    const float dx = __ldg(s_pts + 3 * j0) - Q.qx, dy = __ldg(s_pts + 3 * j0 + 1) - Q.qy, dz = __ldg(s_pts + 3 * j0 + 2) - Q.qz;
    w[0] = kp_influence(dx, dy, dz, Q.k0x, Q.k0y, Q.k0z, Q.inv_sigma);
    w[1] = Q.pad1 ? 0.f : kp_influence(dx, dy, dz, Q.k1x, Q.k1y, Q.k1z, Q.inv_sigma);
  }
  if (j1 >= 0) {
    const float dx = __ldg(s_pts + 3 * j1) - Q.qx, dy = __ldg(s_pts + 3 * j1 + 1) - Q.qy, dz = __ldg(s_pts + 3 * j1 + 2) - Q.qz;
    w[2] = kp_influence(dx, dy, dz, Q.k0x, Q.k0y, Q.k0z, Q.inv_sigma);
    w[3] = Q.pad1 ? 0.f : kp_influence(dx, dy, dz, Q.k1x, Q.k1y, Q.k1z, Q.inv_sigma);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    ahi[e] = tf32_hi_bits(w[e]);
    alo[e] = __float_as_uint(w[e] - __uint_as_float(ahi[e]));  // the tensor core ignores the 13 low bits of lo
  }
}

__device__ __forceinline__ float4 agg_gather(const float* __restrict__ feats, int C, int c0, int j) {
  return j >= 0 ? __ldg(reinterpret_cast<const float4*>(feats + (long long)j * C + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// one neighbour step of one 32-channel block: 4 n-tiles x 3 MMAs, issued term-major so that consecutive MMAs feed
// four independent accumulators
__device__ __forceinline__ void agg_step_mma(float (&acc)[4][4], const float4 f0, const float4 f1, const uint32_t (&ahi)[4],
                                             const uint32_t (&alo)[4]) {
  const float v0[4] = {f0.x, f0.y, f0.z, f0.w}, v1[4] = {f1.x, f1.y, f1.z, f1.w};
  uint32_t b0h[4], b1h[4], b0l[4], b1l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    b0h[i] = tf32_hi_bits(v0[i]); b1h[i] = tf32_hi_bits(v1[i]);
    b0l[i] = __float_as_uint(v0[i] - __uint_as_float(b0h[i])); b1l[i] = __float_as_uint(v1[i] - __uint_as_float(b1h[i]));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) mma_tf32_16x8x8(acc[i], alo, b0h[i], b1h[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) mma_tf32_16x8x8(acc[i], ahi, b0l[i], b1l[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) mma_tf32_16x8x8(acc[i], ahi, b0h[i], b1h[i]);
}

__device__ __forceinline__ void agg_store(float* __restrict__ o0, int C, bool pad1, const float (&acc)[4][4]) {
  // lane (g,t): kernel points g (acc[.][0..1]) and g+8 (acc[.][2..3]), 8 consecutive channels
  *reinterpret_cast<float4*>(o0) = make_float4(acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
  *reinterpret_cast<float4*>(o0 + 4) = make_float4(acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
  if (!pad1) {
    float* o1 = o0 + (long long)8 * C;
    *reinterpret_cast<float4*>(o1) = make_float4(acc[0][2], acc[1][2], acc[2][2], acc[3][2]);
    *reinterpret_cast<float4*>(o1 + 4) = make_float4(acc[0][3], acc[1][3], acc[2][3], acc[3][3]);
  }
}

constexpr int kAggThreads = 128;

// S > 0: H <= 8 S, the query's weight fragments stay in registers across all channel blocks; the gathers of channel
//        block cb+1 are in flight while block cb is multiplied, and the next query's index row is prefetched.
// S == 0: streaming form for any H (weights re-evaluated per 32-channel block; used for H > 56, which the model only
//         meets on its 32-channel stage-0 layers, where there is a single block anyway).
template <int S>
__global__ void __launch_bounds__(kAggThreads, S == 0 ? 4 : (S <= 4 ? 4 : 3)) kpconv_aggregate_mma_kernel(
    const float* __restrict__ feats, int C, const float* __restrict__ q_pts, const float* __restrict__ s_pts,
    const long long* __restrict__ idx, int H, long long ldi, int M, int Ns, const float* __restrict__ kp, float sigma,
    const unsigned char* __restrict__ pos_flag, float* __restrict__ A, float* __restrict__ row_div, int dbg_wrap) {
  pdl_wait();
  pdl_trigger();
  const int wpb = kAggThreads / 32;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  AggQuery Q;
  Q.k0x = kp[3 * g]; Q.k0y = kp[3 * g + 1]; Q.k0z = kp[3 * g + 2];
  Q.pad1 = g == 7;
  const int g1 = Q.pad1 ? g : g + 8;
  Q.k1x = kp[3 * g1]; Q.k1y = kp[3 * g1 + 1]; Q.k1z = kp[3 * g1 + 2];
  Q.inv_sigma = 1.0f / sigma;
  const int stride = gridDim.x * wpb;
  int m = blockIdx.x * wpb + (threadIdx.x >> 5);
  if (m >= M) return;

  if constexpr (S > 0) {
    int jn0[S], jn1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) agg_load_idx(idx + (long long)m * ldi, H, Ns, s, t, jn0[s], jn1[s]);
    for (; m < M; m += stride) {
      int j0[S], j1[S];
      float4 f0[S], f1[S];
      uint32_t ahi[S][4], alo[S][4];
      int cnt = 0;
#pragma unroll
      for (int s = 0; s < S; ++s) { j0[s] = jn0[s]; j1[s] = jn1[s]; }
      // first channel block's gathers go out before the geometry is touched
#pragma unroll
      for (int s = 0; s < S; ++s) { f0[s] = agg_gather(feats, C, 4 * g, j0[s]); f1[s] = agg_gather(feats, C, 4 * g, j1[s]); }
      Q.qx = __ldg(q_pts + 3 * m); Q.qy = __ldg(q_pts + 3 * m + 1); Q.qz = __ldg(q_pts + 3 * m + 2);
      if (g == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          if (j0[s] >= 0) cnt += pos_flag[j0[s]];
          if (j1[s] >= 0) cnt += pos_flag[j1[s]];
        }
      }
      if (m + stride < M) {
#pragma unroll
        for (int s = 0; s < S; ++s) agg_load_idx(idx + (long long)(m + stride) * ldi, H, Ns, s, t, jn0[s], jn1[s]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s) agg_step_weights(Q, s_pts, j0[s], j1[s], ahi[s], alo[s]);
      cnt = warp_sum(cnt);
      if (lane == 0) row_div[m] = (float)max(cnt, 1);
      float* out_row = A + (long long)(dbg_wrap > 0 ? m % dbg_wrap : m) * kKP * C + (long long)g * C + 8 * t;
      for (int cb = 0; cb < C; cb += 32) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
        const bool more = cb + 32 < C;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float4 a = f0[s], b = f1[s];
          if (more) { f0[s] = agg_gather(feats, C, cb + 32 + 4 * g, j0[s]); f1[s] = agg_gather(feats, C, cb + 32 + 4 * g, j1[s]); }
          agg_step_mma(acc, a, b, ahi[s], alo[s]);
        }
        if (dbg_wrap >= 0 || acc[0][0] == 12345.678f) agg_store(out_row + cb, C, Q.pad1, acc);
      }
    }
  } else {
    const int steps = (H + 7) >> 3;
    for (; m < M; m += stride) {
      const long long* idx_row = idx + (long long)m * ldi;
      Q.qx = __ldg(q_pts + 3 * m); Q.qy = __ldg(q_pts + 3 * m + 1); Q.qz = __ldg(q_pts + 3 * m + 2);
      int cnt = 0;
      if (g == 0) {
        for (int h = t; h < H; h += 4) {
          const long long j = idx_row[h];
          if (j >= 0 && j < Ns) cnt += pos_flag[j];
        }
      }
      cnt = warp_sum(cnt);
      if (lane == 0) row_div[m] = (float)max(cnt, 1);
      float* out_row = A + (long long)m * kKP * C + (long long)g * C + 8 * t;
      for (int cb = 0; cb < C; cb += 32) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
#pragma unroll 2
        for (int s = 0; s < steps; ++s) {
          int j0, j1;
          uint32_t ahi[4], alo[4];
          agg_load_idx(idx_row, H, Ns, s, t, j0, j1);
          const float4 a = agg_gather(feats, C, cb + 4 * g, j0), b = agg_gather(feats, C, cb + 4 * g, j1);
          agg_step_weights(Q, s_pts, j0, j1, ahi, alo);
          agg_step_mma(acc, a, b, ahi, alo);
        }
        agg_store(out_row + cb, C, Q.pad1, acc);
      }
    }
  }
}

// ---- cp.async form: the gathers are decoupled from the register file ---------------------------------------------
// Measured on B200 (tools/agg_bench.py): the FFMA kernel, the register-prefetching mma.sync kernel and the same
// kernel with its stores removed all take the same time per layer -- the aggregation is bound by the LATENCY of
// its dependent gathers (index row -> neighbour rows) at 12-20 resident warps per SM, not by arithmetic or by the
// A store.  Here every warp owns a private shared-memory ring: the 32-channel tile of the NEXT (query, channel
// block) and the next query's neighbour coordinates are fetched with cp.async (zero-filled for shadow neighbours)
// while the current tile is multiplied, and the index rows are read two queries ahead.  B fragments come from
// shared memory (row pitch 40 floats: conflict-free for the (g,t) fragment pattern).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kAggPitch = 40;  // floats per staged neighbour row (32 used)

template <int S>
struct AggCp {
  static constexpr int Hp = 8 * S;
  static constexpr int U = 2 * S;  // issue instructions per tile: 4 rows x 8 lanes x 16 bytes each
  static constexpr int kWarpFloats = 2 * Hp * kAggPitch + 2 * Hp * 4;
  static constexpr int kSmemBytes = (kAggThreads / 32) * kWarpFloats * 4;
};

template <int S>
__global__ void __launch_bounds__(kAggThreads, S <= 4 ? 4 : (S <= 6 ? 3 : 2)) kpconv_aggregate_cp_kernel(
    const float* __restrict__ feats, int C, const float* __restrict__ q_pts, const float* __restrict__ s_pts,
    const long long* __restrict__ idx, int H, long long ldi, int M, int Ns, const float* __restrict__ kp, float sigma,
    const unsigned char* __restrict__ pos_flag, float* __restrict__ A, float* __restrict__ row_div) {
  pdl_wait();
  pdl_trigger();
  using L = AggCp<S>;
  constexpr int Hp = L::Hp, U = L::U;
  extern __shared__ __align__(16) float agg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rsub = lane >> 3, ch = lane & 7;  // issue mapping: row 4u + rsub, 16-byte chunk ch of the 128-byte tile row
  float* tbuf = agg_smem + warp * L::kWarpFloats;  // [2][Hp][kAggPitch]
  float* pbuf = tbuf + 2 * Hp * kAggPitch;         // [2][Hp][4]
  const uint32_t tbuf_s = static_cast<uint32_t>(__cvta_generic_to_shared(tbuf));
  const uint32_t pbuf_s = static_cast<uint32_t>(__cvta_generic_to_shared(pbuf));

  AggQuery Q;
  Q.k0x = kp[3 * g]; Q.k0y = kp[3 * g + 1]; Q.k0z = kp[3 * g + 2];
  Q.pad1 = g == 7;
  const int g1 = Q.pad1 ? g : g + 8;
  Q.k1x = kp[3 * g1]; Q.k1y = kp[3 * g1 + 1]; Q.k1z = kp[3 * g1 + 2];
  Q.inv_sigma = 1.0f / sigma;
  const int wpb = kAggThreads / 32;
  const int stride = gridDim.x * wpb;
  int m = blockIdx.x * wpb + warp;
  if (m >= M) return;
  const int nb = C >> 5;

  // index rows are read as (low, high) words; the validity test is applied when the row is USED, two queries after
  // its load was issued, so that no instruction waits on it
  auto load_raw = [&](int mq, int2 (&r)[U]) {
    const int2* row = reinterpret_cast<const int2*>(idx + (long long)mq * ldi);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int h = 4 * u + rsub;
      r[u] = h < H ? __ldg(row + h) : make_int2(-1, -1);
    }
  };
  auto to_j = [&](const int2 (&r)[U], int (&j)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) j[u] = (r[u].y == 0 && (unsigned)r[u].x < (unsigned)Ns) ? r[u].x : -1;
  };
  auto issue_tile = [&](int buf, const int (&j)[U], int cb) {
    const float* src0 = feats + cb * 32 + 4 * ch;
    const uint32_t dst0 = tbuf_s + (uint32_t)(((buf * Hp + rsub) * kAggPitch + 4 * ch) * 4);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int jj = j[u];
      cp_async16(dst0 + (uint32_t)(4 * u * kAggPitch * 4), src0 + (long long)(jj < 0 ? 0 : jj) * C, jj < 0 ? 0 : 16);
    }
  };
  auto issue_pos = [&](int buf, const int (&j)[U]) {
    if (ch < 3) {
      const uint32_t dst0 = pbuf_s + (uint32_t)(((buf * Hp + rsub) * 4 + ch) * 4);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j[u];
        cp_async4(dst0 + (uint32_t)(4 * u * 4 * 4), s_pts + 3ll * (jj < 0 ? 0 : jj) + ch, jj < 0 ? 0 : 4);
      }
    }
  };

  int jr[U], jn[U];
  int2 raw[U];
  load_raw(m, raw);
  to_j(raw, jr);
#pragma unroll
  for (int u = 0; u < U; ++u) jn[u] = -1;
  if (m + stride < M) { load_raw(m + stride, raw); to_j(raw, jn); }
  issue_pos(0, jr);
  issue_tile(0, jr, 0);
  cp_async_commit();
  int p = 0, pp = 0;
  uint32_t ahi[S][4], alo[S][4];
  for (; m < M; m += stride) {
    const bool has_next = m + stride < M;
    const bool has_nn = m + 2 * stride < M;
    if (has_nn) load_raw(m + 2 * stride, raw);
    unsigned char fl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) fl[u] = (ch == 0 && jr[u] >= 0) ? __ldg(pos_flag + jr[u]) : (unsigned char)0;
    Q.qx = __ldg(q_pts + 3 * m); Q.qy = __ldg(q_pts + 3 * m + 1); Q.qz = __ldg(q_pts + 3 * m + 2);
    float* out_row = A + (long long)m * kKP * C + (long long)g * C + 8 * t;
    for (int cbi = 0; cbi < nb; ++cbi) {
      // next tile in flight before this one is consumed
      if (cbi + 1 < nb) {
        issue_tile(p ^ 1, jr, cbi + 1);
        cp_async_commit();
        cp_async_wait<1>();
      } else if (has_next) {
        issue_pos(pp ^ 1, jn);
        issue_tile(p ^ 1, jn, 0);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();  // every lane's copies of this tile have landed and are visible to the warp
      if (cbi == 0) {
        const float4* pq = reinterpret_cast<const float4*>(pbuf + pp * Hp * 4);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float4 a = pq[8 * s + t], b = pq[8 * s + t + 4];
          // shadow / padding rows are zero-filled: their weights are finite and multiply zero feature rows
          float w[4];
          w[0] = kp_influence(a.x - Q.qx, a.y - Q.qy, a.z - Q.qz, Q.k0x, Q.k0y, Q.k0z, Q.inv_sigma);
          w[1] = Q.pad1 ? 0.f : kp_influence(a.x - Q.qx, a.y - Q.qy, a.z - Q.qz, Q.k1x, Q.k1y, Q.k1z, Q.inv_sigma);
          w[2] = kp_influence(b.x - Q.qx, b.y - Q.qy, b.z - Q.qz, Q.k0x, Q.k0y, Q.k0z, Q.inv_sigma);
          w[3] = Q.pad1 ? 0.f : kp_influence(b.x - Q.qx, b.y - Q.qy, b.z - Q.qz, Q.k1x, Q.k1y, Q.k1z, Q.inv_sigma);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ahi[s][e] = tf32_hi_bits(w[e]);
            alo[s][e] = __float_as_uint(w[e] - __uint_as_float(ahi[s][e]));
          }
        }
      }
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
      const float* tb = tbuf + p * Hp * kAggPitch + 4 * g;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float4 f0 = *reinterpret_cast<const float4*>(tb + (8 * s + t) * kAggPitch);
        const float4 f1 = *reinterpret_cast<const float4*>(tb + (8 * s + t + 4) * kAggPitch);
        agg_step_mma(acc, f0, f1, ahi[s], alo[s]);
      }
      agg_store(out_row + cbi * 32, C, Q.pad1, acc);
      __syncwarp();  // the tile is free again before the next iteration's copies overwrite it
      p ^= 1;
    }
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) cnt += fl[u];
    cnt = warp_sum(cnt);
    if (lane == 0) row_div[m] = (float)max(cnt, 1);
    pp ^= 1;
#pragma unroll
    for (int u = 0; u < U; ++u) jr[u] = jn[u];
    if (has_nn) to_j(raw, jn);
  }
}

template <int S>
static int launch_aggregate_cp(const float* feats, int C, const float* q, const float* s, const long long* idx, int H,
                               long long ldi, int M, int Ns, const float* kp, float sigma, const unsigned char* flag, float* A,
                               float* row_div, cudaStream_t st) {
  using L = AggCp<S>;
  auto kern = kpconv_aggregate_cp_kernel<S>;
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), L::kSmemBytes));  // per (kernel, device)
  int per_sm = 0;
  GR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kAggThreads, L::kSmemBytes));
  if (per_sm < 1) return GR_ERR_CAPACITY;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int wpb = kAggThreads / 32;
  int blocks = ceil_div(M, wpb);
  if (blocks > sms * per_sm) blocks = sms * per_sm;
  GR_CHECK_CUDA(launch_pdl(kern, dim3(blocks), dim3(kAggThreads), (size_t)(L::kSmemBytes), st, feats, C, q, s, idx, H, ldi, M, Ns, kp, sigma, flag, A, row_div));
  GR_CHECK_LAUNCH("kpconv_aggregate_cp_kernel");
  return GR_OK;
}

template <int S>
static int launch_aggregate_mma(const float* feats, int C, const float* q, const float* s, const long long* idx, int H,
                                long long ldi, int M, int Ns, const float* kp, float sigma, const unsigned char* flag, float* A,
                                float* row_div, cudaStream_t st) {
  // persistent warps: a few queries per warp so that the next query's index row can be prefetched
  const int wpb = kAggThreads / 32;
  int blocks = ceil_div(M, wpb);
  const int cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  static int dbg = -2;
  if (dbg == -2) { const char* e = getenv("GAUSSREG_AGG_DEBUG_WRAP"); dbg = e ? atoi(e) : 0; }
  GR_CHECK_CUDA(launch_pdl(kpconv_aggregate_mma_kernel<S>, dim3(blocks), dim3(kAggThreads), (size_t)(0), st, feats, C, q, s, idx, H, ldi, M, Ns, kp, sigma, flag, A, row_div, dbg));
  GR_CHECK_LAUNCH("kpconv_aggregate_mma_kernel");
  return GR_OK;
}

template <int GROUP, int CPL>
static int launch_aggregate(const float* feats, int C, const float* q, const float* s, const long long* idx, int H,
                            long long ldi, int M, int Ns, const float* kp, float sigma, const unsigned char* flag, float* A,
                            float* row_div, cudaStream_t st) {
  constexpr int QPW = 32 / GROUP;
  const size_t per_slot = (size_t)((H + 3) & ~3) * 17 * sizeof(float);
  int warps = 4;
  while (warps > 1 && per_slot * QPW * warps > 96 * 1024) warps >>= 1;
  const size_t smem = per_slot * QPW * warps;
  if (smem > 200 * 1024) return GR_ERR_CAPACITY;
  auto kern = kpconv_aggregate_kernel<GROUP, CPL>;
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), (int)smem));
  const int qpb = QPW * warps;
  GR_CHECK_CUDA(launch_pdl(kern, dim3(ceil_div(M, qpb)), dim3(warps * 32), (size_t)(smem), st, feats, C, q, s, idx, H, ldi, M, Ns, kp, sigma, flag, A, row_div));
  GR_CHECK_LAUNCH("kpconv_aggregate_kernel");
  return GR_OK;
}

}  // namespace gr

using namespace gr;

// workspace: Ns bytes of flags
extern "C" size_t gr_kpconv_aggregate_workspace_size(int64_t n_support) { return ((size_t)n_support + 255) & ~size_t(255); }

/* K1 first half.  A: (M, 15*C) f32, row_div: (M) f32 = max(1, #neighbours whose feature row sums > 0). */
extern "C" int gr_kpconv_aggregate(const float* s_feats, int C, const float* q_points, const float* s_points,
                                   const int64_t* neighbor_idx, int H, int64_t ld_idx, int M, int Ns,
                                   const float* kernel_points, int n_kernel_points, float sigma, float* A, float* row_div,
                                   void* ws, size_t ws_bytes, void* stream) {
  if (n_kernel_points != kKP || C <= 0 || H <= 0 || M < 0 || Ns < 0 || !(sigma > 0.f)) return GR_ERR_BAD_ARG;
  if (M == 0) return GR_OK;
  if (!s_feats || !q_points || !s_points || !neighbor_idx || !kernel_points || !A || !row_div) return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < (size_t)Ns) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* flag = static_cast<unsigned char*>(ws);
  if (Ns > 0) {
    GR_CHECK_CUDA(launch_pdl(row_positive_kernel, dim3(ceil_div(Ns, 8)), dim3(256), (size_t)(0), st, s_feats, Ns, C, flag));
    GR_CHECK_LAUNCH("row_positive_kernel");
  }
  const long long* idx = reinterpret_cast<const long long*>(neighbor_idx);
  static int use_mma = -1;
  if (use_mma < 0) { const char* e = getenv("GAUSSREG_KPCONV_MMA"); use_mma = e ? atoi(e) : 1; }
  if (use_mma && C % 32 == 0 && (reinterpret_cast<uintptr_t>(s_feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
#define GR_AGG_MMA(S) return launch_aggregate_mma<S>(s_feats, C, q_points, s_points, idx, H, ld_idx, M, Ns, kernel_points, sigma, flag, A, row_div, st)
    static int use_cp = -1;
    if (use_cp < 0) { const char* e = getenv("GAUSSREG_KPCONV_CP"); use_cp = e ? atoi(e) : 1; }
    if (use_cp) {
#define GR_AGG_CP(S) return launch_aggregate_cp<S>(s_feats, C, q_points, s_points, idx, H, ld_idx, M, Ns, kernel_points, sigma, flag, A, row_div, st)
      switch ((H + 7) >> 3) {
        case 1: GR_AGG_CP(1);
        case 2: GR_AGG_CP(2);
        case 3: GR_AGG_CP(3);
        case 4: GR_AGG_CP(4);
        case 5: GR_AGG_CP(5);
        case 6: GR_AGG_CP(6);
        case 7: GR_AGG_CP(7);
        default: break;  // H > 56: streaming mma kernel below
      }
#undef GR_AGG_CP
    }
    switch ((H + 7) >> 3) {
      case 1: GR_AGG_MMA(1);
      case 2: GR_AGG_MMA(2);
      case 3: GR_AGG_MMA(3);
      case 4: GR_AGG_MMA(4);
      case 5: GR_AGG_MMA(5);
      case 6: GR_AGG_MMA(6);
      case 7: GR_AGG_MMA(7);
      default: GR_AGG_MMA(0);
    }
#undef GR_AGG_MMA
  }
#define GR_AGG(G, V) return launch_aggregate<G, V>(s_feats, C, q_points, s_points, idx, H, ld_idx, M, Ns, kernel_points, sigma, flag, A, row_div, st)
  // the kernel is bound by instruction issue, not bandwidth: four channels per lane where C allows it (two queries
  // per warp for C = 64 / 32) keeps the FMA share of the instruction stream high
  if (C % 128 == 0) GR_AGG(32, 4);
  if (C % 64 == 0) GR_AGG(16, 4);
  if (C % 32 == 0) GR_AGG(16, 2);
  if (C == 16) GR_AGG(16, 1);
  if (C == 8) GR_AGG(8, 1);
  if (C == 4) GR_AGG(4, 1);
  if (C == 2) GR_AGG(2, 1);
  if (C == 1) GR_AGG(1, 1);
#undef GR_AGG
  return GR_ERR_BAD_ARG;
}
