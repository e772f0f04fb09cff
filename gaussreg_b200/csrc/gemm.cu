// fp32 GEMM with fused epilogue (SIMT FFMA path).
//
//   C[b] = act( alpha * A[b] * op(B[b]) / row_div[:,None] + bias[None,:] + residual[b] )
//
// A is (M,K) row-major with leading dimension lda; B is either (N,K) row-major ("NT": torch Linear
// weights, Q K^T) or (K,N) row-major ("NN": KPConv weights, P V).  Batched through element strides, so
// "b n (h c)" head slices are addressed without copies.  fp32 accumulate in K order: this is the
// selection-safe path (top-k / argmin steps follow most of these products, SURVEY.md section 7 hard
// part 2); the tcgen05 3xTF32 path for the large products lives in gemm_tc.cu.
#include <stdlib.h>

#include "common.cuh"

extern "C" int gr_get_gemm_mode(void);

namespace gr {

struct GemmParams {
  const float* A; const float* B; float* C;
  const float* bias; const float* row_div; const float* residual;
  long long lda, ldb, ldc, ldr;
  long long sA, sB, sC, sR;  // batch strides (elements)
  int M, N, K;
  float alpha;
  int act;  // 0 none, 1 relu, 2 leaky relu 0.1
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.1f * v;
  return v;
}

template <int BM, int BN, int BK, int TM, int TN, bool TRANSB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) sgemm_kernel(GemmParams p) {
  pdl_wait();
  pdl_trigger();
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* __restrict__ A = p.A + (long long)blockIdx.z * p.sA;
  const float* __restrict__ B = p.B + (long long)blockIdx.z * p.sB;
  float* __restrict__ C = p.C + (long long)blockIdx.z * p.sC;

  // global -> register staging, one element per (thread, step)
  constexpr int A_ELEMS = BM * BK / NT;
  constexpr int B_ELEMS = BN * BK / NT;
  float ra[A_ELEMS], rb[B_ELEMS];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_ELEMS; ++i) {
      const int e = tid + i * NT;          // consecutive threads walk K first (K-contiguous rows)
      const int kk = e % BK, mm = e / BK;
      const int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < p.M && gk < p.K) ? A[(long long)gm * p.lda + gk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < B_ELEMS; ++i) {
      const int e = tid + i * NT;
      if (TRANSB) {
        const int kk = e % BK, nn = e / BK;
        const int gn = n0 + nn, gk = k0 + kk;
        rb[i] = (gn < p.N && gk < p.K) ? B[(long long)gn * p.ldb + gk] : 0.f;
      } else {
        const int nn = e % BN, kk = e / BN;  // N-contiguous rows
        const int gn = n0 + nn, gk = k0 + kk;
        rb[i] = (gn < p.N && gk < p.K) ? B[(long long)gk * p.ldb + gn] : 0.f;
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_ELEMS; ++i) {
      const int e = tid + i * NT;
      As[buf][e % BK][e / BK] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_ELEMS; ++i) {
      const int e = tid + i * NT;
      if (TRANSB) Bs[buf][e % BK][e / BK] = rb[i];
      else Bs[buf][e / BN][e % BN] = rb[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int t = 0; t < nk; ++t) {
    const int buf = t & 1;
    if (t + 1 < nk) load_tiles((t + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  const float* __restrict__ R = p.residual ? p.residual + (long long)blockIdx.z * p.sR : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= p.M) continue;
    const float rd = p.row_div ? p.row_div[gm] : 1.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.row_div) v = v / rd;
      if (p.bias) v += p.bias[gn];
      if (R) v += R[(long long)gm * p.ldr + gn];
      C[(long long)gm * p.ldc + gn] = apply_act(v, p.act);
    }
  }
}

// Superpoint-sized products (M ~ 500): a 32x32 output tile per CTA, the K loop split over four 64-thread groups that
// each own a private shared-memory tile, partial sums folded in a fixed order at the end.  Four times the loads in
// flight per CTA: these products are pure latency (a few dozen CTAs of work), not throughput.
template <bool TRANSB>
__global__ void __launch_bounds__(256) sgemm_small_kernel(GemmParams p) {
  pdl_wait();
  pdl_trigger();
  constexpr int BM = 32, BN = 32, BK = 32, G = 4, NT = 64;
  __shared__ __align__(16) float As[G][BK][BM + 4];
  __shared__ __align__(16) float Bs[G][BK][BN + 4];
  const int tid = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int tx = tid % 8, ty = tid / 8;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* __restrict__ A = p.A + (long long)blockIdx.z * p.sA;
  const float* __restrict__ B = p.B + (long long)blockIdx.z * p.sB;
  float* __restrict__ C = p.C + (long long)blockIdx.z * p.sC;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (p.K + BK - 1) / BK;
  for (int t = grp; t < nk; t += G) {
    const int k0 = t * BK;
#pragma unroll
    for (int i = 0; i < BM * BK / NT; ++i) {
      const int e = tid + i * NT;
      const int kk = e % BK, mm = e / BK;
      const int gm = m0 + mm, gk = k0 + kk;
      As[grp][kk][mm] = (gm < p.M && gk < p.K) ? A[(long long)gm * p.lda + gk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < BN * BK / NT; ++i) {
      const int e = tid + i * NT;
      if (TRANSB) {
        const int kk = e % BK, nn = e / BK;
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[grp][kk][nn] = (gn < p.N && gk < p.K) ? B[(long long)gn * p.ldb + gk] : 0.f;
      } else {
        const int nn = e % BN, kk = e / BN;
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[grp][kk][nn] = (gn < p.N && gk < p.K) ? B[(long long)gk * p.ldb + gn] : 0.f;
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");  // named barrier of this 64-thread group
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[grp][kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[grp][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
  }
  // fold the four partial tiles: groups 1..3 park their accumulators in (their own) shared memory
  __syncthreads();
  float* park = &As[0][0][0];  // G*BK*(BM+4) floats >= 3 * 64 * 16
  if (grp > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) park[((grp - 1) * 64 + tid) * 16 + i * 4 + j] = acc[i][j];
  }
  __syncthreads();
  if (grp != 0) return;
#pragma unroll
  for (int g = 0; g < G - 1; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += park[(g * 64 + tid) * 16 + i * 4 + j];
  const float* __restrict__ R = p.residual ? p.residual + (long long)blockIdx.z * p.sR : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
    const float rd = p.row_div ? p.row_div[gm] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.row_div) v = v / rd;
      if (p.bias) v += p.bias[gn];
      if (R) v += R[(long long)gm * p.ldr + gn];
      C[(long long)gm * p.ldc + gn] = apply_act(v, p.act);
    }
  }
}

// ---- superpoint-sized products on the (legacy-path) tensor cores --------------------------------------------------
// The ~480-row transformer products are far too small for 128-row tcgen05 tiles (a handful of CTAs), and the FFMA
// kernel above spends 14-20 us on each of them (one L2 round trip per 32-deep k-slab, no overlap).  This kernel keeps
// them on 64x64 tiles (dozens of CTAs) but runs the inner product as mma.sync m16n8k8 TF32 with the 3xTF32 split
// (fp32-level accuracy: the raw fp32 word serves as the truncated hi part, lo = x - trunc(x), products lo*hi + hi*lo +
// hi*hi in fp32 accumulators) and double-buffers the k-slabs with cp.async.  4 warps, warp tile 32x32.
// Requirements (checked by the dispatcher): 16-byte aligned operands, lda/ldb/strides % 4 == 0, K % 4 == 0
// (and N % 4 == 0 for the (K,N) "NN" form).
__device__ __forceinline__ void sm_cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void sm_mma_tf32(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned sm_lo_bits(float x) { return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u)); }

template <bool TRANSB>
__global__ void __launch_bounds__(128) gemm_mma_small_kernel(GemmParams p) {
  pdl_wait();
  pdl_trigger();
  constexpr int BM = 64, BN = 64, BK = 32;
  constexpr int PA = BK + 4;                       // A tile (and the (N,K) B tile): [64][36], conflict-free fragment reads
  constexpr int PB = TRANSB ? BK + 4 : BN + 8;     // (K,N) B tile: [32][72]
  constexpr int B_ELEMS = TRANSB ? BN * PB : BK * PB;
  __shared__ __align__(16) float As[2][BM * PA];
  __shared__ __align__(16) float Bs[2][B_ELEMS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* __restrict__ A = p.A + (long long)blockIdx.z * p.sA;
  const float* __restrict__ B = p.B + (long long)blockIdx.z * p.sB;
  float* __restrict__ C = p.C + (long long)blockIdx.z * p.sC;

  auto load_slab = [&](int buf, int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // A: 64 rows x 8 chunks of 16 bytes
      const int e = tid + i * 128, r = e >> 3, c = e & 7;
      const int gm = m0 + r, gk = k0 + 4 * c;
      const bool ok = gm < p.M && gk < p.K;
      sm_cp_async16(&As[buf][r * PA + 4 * c], A + (ok ? (long long)gm * p.lda + gk : 0), ok ? 16 : 0);
    }
    if (TRANSB) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 128, r = e >> 3, c = e & 7;
        const int gn = n0 + r, gk = k0 + 4 * c;
        const bool ok = gn < p.N && gk < p.K;
        sm_cp_async16(&Bs[buf][r * PB + 4 * c], B + (ok ? (long long)gn * p.ldb + gk : 0), ok ? 16 : 0);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // 32 k-rows x 16 chunks
        const int e = tid + i * 128, r = e >> 4, c = e & 15;
        const int gk = k0 + r, gn = n0 + 4 * c;
        const bool ok = gk < p.K && gn < p.N;  // N % 4 == 0: a chunk is entirely inside or outside
        sm_cp_async16(&Bs[buf][r * PB + 4 * c], B + (ok ? (long long)gk * p.ldb + gn : 0), ok ? 16 : 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  load_slab(0, 0);
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      load_slab(buf ^ 1, (kt + 1) * BK);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* as = As[buf];
    const float* bs = Bs[buf];
#pragma unroll
    for (int k8 = 0; k8 < BK; k8 += 8) {
      unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float* ap = as + (wm + 16 * i + g) * PA + k8 + t;
        const float a0 = ap[0], a1 = ap[8 * PA], a2 = ap[4], a3 = ap[8 * PA + 4];
        ah[i][0] = __float_as_uint(a0); ah[i][1] = __float_as_uint(a1); ah[i][2] = __float_as_uint(a2); ah[i][3] = __float_as_uint(a3);
        al[i][0] = sm_lo_bits(a0); al[i][1] = sm_lo_bits(a1); al[i][2] = sm_lo_bits(a2); al[i][3] = sm_lo_bits(a3);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float b0, b1;
        if (TRANSB) { const float* bp = bs + (wn + 8 * j + g) * PB + k8 + t; b0 = bp[0]; b1 = bp[4]; }
        else { const float* bp = bs + (k8 + t) * PB + wn + 8 * j + g; b0 = bp[0]; b1 = bp[4 * PB]; }
        bh[j][0] = __float_as_uint(b0); bh[j][1] = __float_as_uint(b1);
        bl[j][0] = sm_lo_bits(b0); bl[j][1] = sm_lo_bits(b1);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sm_mma_tf32(acc[i][j], al[i][0], al[i][1], al[i][2], al[i][3], bh[j][0], bh[j][1]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sm_mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bl[j][0], bl[j][1]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sm_mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bh[j][0], bh[j][1]);
    }
    __syncthreads();
  }
  // epilogue: lane (g,t) holds rows g, g+8 and columns 2t, 2t+1 of every 16x8 tile
  const float* __restrict__ R = p.residual ? p.residual + (long long)blockIdx.z * p.sR : nullptr;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int gm = m0 + wm + 16 * i + g + 8 * hrow;
      if (gm >= p.M) continue;
      const float rd = p.row_div ? p.row_div[gm] : 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gn = n0 + wn + 8 * j + 2 * t + e;
          if (gn >= p.N) continue;
          float v = acc[i][j][2 * hrow + e] * p.alpha;
          if (p.row_div) v = v / rd;
          if (p.bias) v += p.bias[gn];
          if (R) v += R[(long long)gm * p.ldr + gn];
          C[(long long)gm * p.ldc + gn] = apply_act(v, p.act);
        }
    }
}

// ---- latency-optimised form for the superpoint-sized products (M of a few hundred rows) ---------------------------
// Measured (profiles/r02a_launches_by_kernel.txt): the double-buffered 64x64 kernel above spends 22 us on a
// 480 x 256 x 256 product -- 32 CTAs, each walking its 8 k-slabs one L2 round trip after the other.  These products
// are pure latency, so here (a) tiles are 32 x 32 (120+ CTAs), (b) the FOUR WARPS SPLIT K: each warp fetches its own
// quarter of the A and B panels with cp.async (two commit groups, no block barrier in the main loop), multiplies it
// into a full 32 x 32 partial tile, and (c) the four partials are folded through shared memory in a fixed order before
// the fused epilogue.  One exposed memory latency per CTA instead of K / 32.
constexpr int kTinyKq = 64;   // K handled per warp and pass (K <= 256 in one pass; 66 KB of panels -> 3 CTAs per SM)

__device__ __forceinline__ void sm_cp_async4(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}

// A_VEC = false: A rows are not 16-byte addressable (odd pitch or K, e.g. the (N,N) attention probabilities with odd N):
// its panel is fetched with 4-byte cp.async instead
template <bool TRANSB, bool A_VEC, int KQ>
__global__ void __launch_bounds__(128) gemm_mma_tiny_kernel(GemmParams p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float tiny_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  // KQ (the K range of one warp per pass) is a template parameter: every panel index below is then a compile-time
  // constant and the load loops unroll into plain cp.async sequences -- with a run-time kq the index arithmetic
  // (divisions by the chunk count) was two thirds of the instructions of this latency-bound kernel.
  constexpr int PA = KQ + 4;                 // A panel [32][KQ + 4] (and the (N,K) B panel)
  constexpr int PBn = 40;                    // (K,N) B panel [KQ][40]
  constexpr int a_elems = 32 * PA, b_elems = TRANSB ? 32 * PA : KQ * PBn;
  constexpr int HALF = KQ / 2;               // first commit group covers [0, HALF), the second [HALF, KQ)
  constexpr int C4 = HALF / 4;               // 16-byte chunks per row and part: 2, 4 or 8
  float* As = tiny_smem + warp * (a_elems + b_elems);
  float* Bs = As + a_elems;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const float* __restrict__ A = p.A + (long long)blockIdx.z * p.sA;
  const float* __restrict__ B = p.B + (long long)blockIdx.z * p.sB;
  float* __restrict__ C = p.C + (long long)blockIdx.z * p.sC;

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  for (int kpass = 0; kpass < p.K; kpass += 4 * KQ) {
    const int kw = kpass + warp * KQ;         // this warp's K range [kw, kw + KQ)
    if (kpass > 0) __syncwarp();
    // ---- issue this warp's loads: two groups
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      constexpr int kb0 = 0;
      const int kb = part == 0 ? kb0 : HALF;
      if (A_VEC) {
#pragma unroll
        for (int i = 0; i < C4; ++i) {
          const int e = lane + 32 * i;
          const int r = e / C4, c = e % C4;
          const int gm = m0 + r, gk = kw + kb + 4 * c;
          const bool ok = gm < p.M && gk < p.K;
          const int rest = (p.K - gk) * 4;  // a row may end inside the chunk (K % 4 != 0 on a padded pitch): zero fill
          sm_cp_async16(As + r * PA + kb + 4 * c, A + (ok ? (long long)gm * p.lda + gk : 0), ok ? (rest < 16 ? rest : 16) : 0);
        }
      } else {
#pragma unroll 4
        for (int i = 0; i < HALF; ++i) {
          const int e = lane + 32 * i;
          const int r = e / HALF, c = e % HALF;
          const int gm = m0 + r, gk = kw + kb + c;
          const bool ok = gm < p.M && gk < p.K;
          sm_cp_async4(As + r * PA + kb + c, A + (ok ? (long long)gm * p.lda + gk : 0), ok ? 4 : 0);
        }
      }
      if (TRANSB) {
#pragma unroll
        for (int i = 0; i < C4; ++i) {
          const int e = lane + 32 * i;
          const int r = e / C4, c = e % C4;
          const int gn = n0 + r, gk = kw + kb + 4 * c;
          const bool ok = gn < p.N && gk < p.K;
          sm_cp_async16(Bs + r * PA + kb + 4 * c, B + (ok ? (long long)gn * p.ldb + gk : 0), ok ? 16 : 0);
        }
      } else {
#pragma unroll
        for (int i = 0; i < HALF / 4; ++i) {  // HALF rows of eight 16-byte chunks
          const int e = lane + 32 * i;
          const int r = kb + (e >> 3), c = e & 7;
          const int gk = kw + r, gn = n0 + 4 * c;
          const bool ok = gk < p.K && gn < p.N;
          sm_cp_async16(Bs + r * PBn + 4 * c, B + (ok ? (long long)gk * p.ldb + gn : 0), ok ? 16 : 0);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- multiply: first half while the second is still in flight
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      if (part == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      const int kb = part == 0 ? 0 : HALF;
#pragma unroll
      for (int k8 = 0; k8 < HALF; k8 += 8) {
        if (kw + kb + k8 >= p.K) break;        // zero-filled tail
        unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float* ap = As + (16 * i + g) * PA + kb + k8 + t;
          const float a0 = ap[0], a1 = ap[8 * PA], a2 = ap[4], a3 = ap[8 * PA + 4];
          ah[i][0] = __float_as_uint(a0); ah[i][1] = __float_as_uint(a1); ah[i][2] = __float_as_uint(a2); ah[i][3] = __float_as_uint(a3);
          al[i][0] = sm_lo_bits(a0); al[i][1] = sm_lo_bits(a1); al[i][2] = sm_lo_bits(a2); al[i][3] = sm_lo_bits(a3);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float b0, b1;
          if (TRANSB) { const float* bp = Bs + (8 * j + g) * PA + kb + k8 + t; b0 = bp[0]; b1 = bp[4]; }
          else { const float* bp = Bs + (kb + k8 + t) * PBn + 8 * j + g; b0 = bp[0]; b1 = bp[4 * PBn]; }
          bh[j][0] = __float_as_uint(b0); bh[j][1] = __float_as_uint(b1);
          bl[j][0] = sm_lo_bits(b0); bl[j][1] = sm_lo_bits(b1);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) sm_mma_tf32(acc[i][j], al[i][0], al[i][1], al[i][2], al[i][3], bh[j][0], bh[j][1]);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) sm_mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bl[j][0], bl[j][1]);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) sm_mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bh[j][0], bh[j][1]);
      }
    }
  }
  // ---- fold the four K-quarters (fixed order 0+1+2+3), then the epilogue on 8 outputs per thread
  __syncthreads();
  float* red = tiny_smem;  // [4][32][33]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = 16 * i + g + 8 * (e >> 1), c = 8 * j + 2 * t + (e & 1);
        red[(warp * 32 + r) * 33 + c] = acc[i][j][e];
      }
  __syncthreads();
  const float* __restrict__ R = p.residual ? p.residual + (long long)blockIdx.z * p.sR : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = tid + i * 128, r = e >> 5, c = e & 31;
    const int gm = m0 + r, gn = n0 + c;
    if (gm >= p.M || gn >= p.N) continue;
    float v = ((red[r * 33 + c] + red[(32 + r) * 33 + c]) + red[(64 + r) * 33 + c]) + red[(96 + r) * 33 + c];
    v *= p.alpha;
    if (p.row_div) v = v / p.row_div[gm];
    if (p.bias) v += p.bias[gn];
    if (R) v += R[(long long)gm * p.ldr + gn];
    C[(long long)gm * p.ldc + gn] = apply_act(v, p.act);
  }
}

template <bool TB, bool AV, int KQ>
static int launch_tiny_kq(const GemmParams& p, int batch, cudaStream_t st) {
  constexpr int PA = KQ + 4;
  constexpr size_t per_warp = (size_t)32 * PA + (TB ? (size_t)32 * PA : (size_t)KQ * 40);
  constexpr size_t red = 4 * 32 * 33 * sizeof(float);
  constexpr size_t smem = 4 * per_warp * sizeof(float) > red ? 4 * per_warp * sizeof(float) : red;
  dim3 grid((p.N + 31) / 32, (p.M + 31) / 32, batch);
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(gemm_mma_tiny_kernel<TB, AV, KQ>), (int)smem));
  GR_CHECK_CUDA(launch_pdl(gemm_mma_tiny_kernel<TB, AV, KQ>, grid, dim3(128), smem, st, p));
  GR_CHECK_LAUNCH("gemm_mma_tiny_kernel");
  return GR_OK;
}

template <bool TB, bool AV>
static int launch_tiny_av(const GemmParams& p, int batch, cudaStream_t st) {
  // K range per warp and pass: a quarter of K, rounded up to 16 / 32 / 64 (K <= 256 runs in one pass)
  if (p.K <= 64) return launch_tiny_kq<TB, AV, 16>(p, batch, st);
  if (p.K <= 128) return launch_tiny_kq<TB, AV, 32>(p, batch, st);
  return launch_tiny_kq<TB, AV, kTinyKq>(p, batch, st);
}

static int launch_tiny(const GemmParams& p, int batch, bool transb, cudaStream_t st) {
  const bool a_vec = (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && p.lda % 4 == 0 && p.sA % 4 == 0;
  if (transb) return a_vec ? launch_tiny_av<true, true>(p, batch, st) : launch_tiny_av<true, false>(p, batch, st);
  return a_vec ? launch_tiny_av<false, true>(p, batch, st) : launch_tiny_av<false, false>(p, batch, st);
}

static bool mma_knob() {
  static int knob = -1;
  if (knob < 0) { const char* e = getenv("GAUSSREG_GEMM_SMALL_MMA"); knob = e ? atoi(e) : 1; }
  return knob != 0;
}
static bool mma_small_ok(const GemmParams& p, bool transb) {
  if (!mma_knob()) return false;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(p.A) || !al16(p.B) || p.lda % 4 != 0 || p.ldb % 4 != 0 || p.sA % 4 != 0 || p.sB % 4 != 0 || p.K % 4 != 0) return false;
  if (!transb && p.N % 4 != 0) return false;
  return p.K >= 16;
}
// the tiny kernel fetches A element-wise when needed; B must be 16-byte addressable ((N,K): K % 4 == 0; (K,N): N % 4 == 0)
static bool mma_tiny_ok(const GemmParams& p, bool transb) {
  if (!mma_knob()) return false;
  if ((reinterpret_cast<uintptr_t>(p.B) & 15) != 0 || p.ldb % 4 != 0 || p.sB % 4 != 0) return false;
  if (transb ? (p.K % 4 != 0) : (p.N % 4 != 0)) return false;
  return p.K >= 8;
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_sgemm(const GemmParams& p, int batch, bool transb, cudaStream_t st) {
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, batch);
  constexpr int NT = (BM / TM) * (BN / TN);
  if (transb) GR_CHECK_CUDA(launch_pdl(sgemm_kernel<BM, BN, BK, TM, TN, true>, dim3(grid), dim3(NT), (size_t)(0), st, p));
  else GR_CHECK_CUDA(launch_pdl(sgemm_kernel<BM, BN, BK, TM, TN, false>, dim3(grid), dim3(NT), (size_t)(0), st, p));
  GR_CHECK_LAUNCH("sgemm_kernel");
  return GR_OK;
}

int sgemm(const GemmParams& p, int batch, bool transb, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || batch <= 0) return GR_OK;
  if (gr_get_gemm_mode() == 1) {
    static int tiny = -1;
    if (tiny < 0) { const char* e = getenv("GAUSSREG_GEMM_TINY"); tiny = e ? atoi(e) : 1; }
    const long long tiles32 = (long long)((p.M + 31) / 32) * ((p.N + 31) / 32) * batch;
    // a few hundred rows: latency-optimised 32x32 tiles with K split over the warps
    if (tiny && tiles32 <= 148 * 8 && mma_tiny_ok(p, transb)) return launch_tiny(p, batch, transb, st);
    // everything else that is 16-byte addressable: 64x64 tiles on mma.sync (3xTF32), cp.async double buffering
    if (mma_small_ok(p, transb)) {
      dim3 grid((p.N + 63) / 64, (p.M + 63) / 64, batch);
      if (transb) GR_CHECK_CUDA(launch_pdl(gemm_mma_small_kernel<true>, grid, dim3(128), 0, st, p));
      else GR_CHECK_CUDA(launch_pdl(gemm_mma_small_kernel<false>, grid, dim3(128), 0, st, p));
      GR_CHECK_LAUNCH("gemm_mma_small_kernel");
      return GR_OK;
    }
    if (tiny && mma_tiny_ok(p, transb) && tiles32 <= 148 * 64) return launch_tiny(p, batch, transb, st);
  }
  // fp32 FFMA kernels: SIMT mode, odd shapes; pick the largest tile that still yields ~1.5 waves of CTAs on 148 SMs
  const long long ctas128 = (long long)((p.M + 127) / 128) * ((p.N + 127) / 128) * batch;
  const long long ctas64 = (long long)((p.M + 63) / 64) * ((p.N + 63) / 64) * batch;
  if (ctas128 >= 222) return launch_sgemm<128, 128, 8, 8, 8>(p, batch, transb, st);
  if (ctas64 >= 148) return launch_sgemm<64, 64, 8, 4, 4>(p, batch, transb, st);
  {
    dim3 grid((p.N + 31) / 32, (p.M + 31) / 32, batch);
    if (transb) GR_CHECK_CUDA(launch_pdl(sgemm_small_kernel<true>, dim3(grid), dim3(256), (size_t)(0), st, p));
    else GR_CHECK_CUDA(launch_pdl(sgemm_small_kernel<false>, dim3(grid), dim3(256), (size_t)(0), st, p));
    GR_CHECK_LAUNCH("sgemm_small_kernel");
    return GR_OK;
  }
}

// gemm_tc.cu
int gemm_tf32x3(const float* A, long long lda, long long sA, const float* B, long long ldb, long long sB, float* C, long long ldc,
                long long sC, int M, int N, int K, int batch, float alpha, const float* bias, const float* row_div,
                const float* residual, long long ldr, long long sR, int act, cudaStream_t st, const float* B_packed,
                GnStatsOut* gn, const void* B_packed16, float inv_scale16);

static int g_gemm_mode = -1;  // 0: SIMT only, 1: tensor cores where the problem qualifies
static thread_local int g_last_path = 0;

}  // namespace gr

using namespace gr;

/* mode 0 = fp32 FFMA kernels only, 1 = tcgen05 3xTF32 for large K-major products (default; env GAUSSREG_GEMM=simt|tc). */
extern "C" void gr_set_gemm_mode(int mode) { g_gemm_mode = mode ? 1 : 0; }
extern "C" int gr_get_gemm_mode(void) {
  if (g_gemm_mode < 0) {
    const char* e = getenv("GAUSSREG_GEMM");
    g_gemm_mode = (e && (e[0] == 's' || e[0] == 'S')) ? 0 : 1;
  }
  return g_gemm_mode;
}

/* 1 if the calling thread's last gr_gemm ran on the tensor cores (tcgen05), 0 if on the FFMA kernel. */
extern "C" int gr_last_gemm_path(void) { return g_last_path; }

static int gemm_dispatch(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB,
                         int trans_b, float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch, float alpha,
                         const float* bias, const float* row_div, const float* residual, int64_t ldr, int64_t strideR,
                         int act, void* stream, const float* B_packed, GnStatsOut* gn = nullptr, const void* B_packed16 = nullptr,
                         float inv_scale16 = 1.f) {
  if (gn) gn->nblk = 0;
  if (M < 0 || N < 0 || K < 0 || batch < 0 || act < 0 || act > 2) return GR_ERR_BAD_ARG;
  if (M == 0 || N == 0 || batch == 0) return GR_OK;
  if (!A || !B || !C) return GR_ERR_BAD_ARG;
  if (trans_b && gr_get_gemm_mode() == 1) {
    const int rc = gemm_tf32x3(A, lda, strideA, B, ldb, strideB, C, ldc, strideC, M, N, K, batch, alpha, bias, row_div, residual,
                               ldr, strideR, act, static_cast<cudaStream_t>(stream), B_packed, gn, B_packed16, inv_scale16);
    if (rc <= 0) { g_last_path = 1; return rc; }
  }
  if (gn) gn->nblk = 0;
  g_last_path = 0;
  GemmParams p;
  p.A = A; p.B = B; p.C = C; p.bias = bias; p.row_div = row_div; p.residual = residual;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldr = ldr;
  p.sA = strideA; p.sB = strideB; p.sC = strideC; p.sR = strideR;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.act = act;
  return sgemm(p, batch, trans_b != 0, static_cast<cudaStream_t>(stream));
}

namespace gr {
int gemm_ex(const float* A, long long lda, const float* B, long long ldb, int trans_b, float* C, long long ldc, int M, int N, int K,
            float alpha, const float* bias, const float* row_div, const float* residual, long long ldr, int act, void* stream,
            const float* B_packed, GnStatsOut* gn, const void* B_packed16, float inv_scale16) {
  return gemm_dispatch(A, lda, 0, B, ldb, 0, trans_b, C, ldc, 0, M, N, K, 1, alpha, bias, row_div, residual, ldr, 0, act, stream,
                       B_packed, gn, B_packed16, inv_scale16);
}
}  // namespace gr

extern "C" int gr_gemm(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB,
                       int trans_b, float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch, float alpha,
                       const float* bias, const float* row_div, const float* residual, int64_t ldr, int64_t strideR,
                       int act, void* stream) {
  return gemm_dispatch(A, lda, strideA, B, ldb, strideB, trans_b, C, ldc, strideC, M, N, K, batch, alpha, bias, row_div, residual,
                       ldr, strideR, act, stream, nullptr);
}

/* C = act(alpha * A . W^T / row_div + bias + residual) for a static weight W (N,K): `W_packed` is its
 * gr_pack_weight_tf32x3 image (tensor-core operand format), `W` the plain matrix (used when the problem is routed
 * to the FFMA kernel). */
extern "C" int gr_linear_packed(const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_packed, float* C,
                                int64_t ldc, int M, int N, int K, float alpha, const float* bias, const float* row_div,
                                const float* residual, int64_t ldr, int act, void* stream) {
  return gemm_dispatch(A, lda, 0, W, ldw, 0, 1, C, ldc, 0, M, N, K, 1, alpha, bias, row_div, residual, ldr, 0, act, stream, W_packed);
}

/* gr_linear_packed with the fp16-split image of W as well (gr_pack_weight_f16x3, packing scale 1 / inv_scale16): products the
 * tensor-core path accepts then run on kind::f16 (see gemm_tc.cu). */
extern "C" int gr_linear_packed16(const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_packed,
                                  const void* W_packed16, float inv_scale16, float* C, int64_t ldc, int M, int N, int K, float alpha,
                                  const float* bias, const float* row_div, const float* residual, int64_t ldr, int act, void* stream) {
  return gemm_dispatch(A, lda, 0, W, ldw, 0, 1, C, ldc, 0, M, N, K, 1, alpha, bias, row_div, residual, ldr, 0, act, stream, W_packed,
                       nullptr, W_packed16, inv_scale16);
}
