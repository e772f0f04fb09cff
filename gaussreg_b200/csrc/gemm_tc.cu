// Tensor-core GEMM for sm_100a: tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-accurate).
//
//   C = act(alpha * A . B^T / row_div + bias + residual),  A (M,K) and B (N,K) both K-major fp32.
//
// Every fp32 operand x is split on the fly into  hi = tf32(x)  and  lo = tf32(x - hi); the product is
// accumulated as  lo*hi + hi*lo + hi*hi  in a TMEM fp32 accumulator, which restores ~fp32 accuracy
// (the top-k / argmin selections that follow these products do not tolerate single-pass TF32, see
// DESIGN.md).  Warp roles per CTA (one 128 x BN output tile):
//   warps 0-7  producers: ld.global (float4) -> split -> st.shared into the UMMA canonical K-major
//              SWIZZLE_128B layout (hi and lo tiles), then the epilogue (tcgen05.ld -> registers ->
//              fused epilogue -> st.global);
//   warp 8     one elected thread issues the tcgen05.mma chain and tcgen05.commit's.
// Stage hand-over is by mbarriers: full[s] (256 producer arrivals) / empty[s] (tcgen05.commit).
#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no link-time libcuda)
#include <stdlib.h>

#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace gr {

struct GemmParams;  // gemm.cu


namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                       // 32 fp32 = 128 bytes = one SWIZZLE_128B row
constexpr int kThreads = kProducerThreads + 32;

struct Params {
  const float* A; const float* B; float* C;
  const float* bias; const float* row_div; const float* residual;
  long long lda, ldb, ldc, ldr;
  long long sA, sB, sC, sR;
  int M, N, K;
  float alpha;
  int act;
  int k_split;  // 0: blockIdx.z is a batch index; > 0: blockIdx.z selects the K slice [z*k_split, (z+1)*k_split)
  // optional: B pre-split into hi/lo TF32 tiles in the shared-memory byte order (gr_pack_weight_tf32x3, padded to
  // 128 rows x 32 k): the B tiles then arrive by bulk async copies instead of being converted by the producers
  const float* B_packed;
  int packed_kblocks;  // ceil(K / 32) of the packed matrix
  // optional GroupNorm statistics of the OUTPUT (K2 fused into its producer): per 128-row tile and group the
  // (sum, sum of squares) of the final values, folded in a fixed order -> gn_partial[blockIdx.y * gn_groups + g]
  double2* gn_partial;
  int gn_groups;
  // optional: B pre-split into fp16 hi/lo tiles (gr_pack_weight_f16x3: per 128-row tile and 64-wide k-block [hi 16 KB][lo 16 KB],
  // B multiplied by a power of two whose inverse the launcher folds into alpha) for the kind::f16 persistent kernel
  const unsigned char* B_packed16;
  int packed_kblocks64;
  float inv_scale16;
};

template <int BN>
struct Cfg {
  static constexpr int kStages = BN == 256 ? 2 : (BN == 128 ? 3 : 2);  // BN = 64: 2 stages so that two CTAs share an SM
  static constexpr int kABytes = BM * BK * 4;      // one A tile (hi or lo)
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Operand tile = ROWS x 32 fp32 (K-major, ld elements).  The producers first issue all their 16-byte global loads
// (tile_load), and convert / store them one k-block later (tile_store), so the L2/HBM latency of k-block kb+1
// overlaps the TF32 split of k-block kb.
template <int ROWS>
struct TileRegs { float4 v[ROWS * 8 / kProducerThreads]; };

template <int ROWS>
__device__ __forceinline__ void tile_load(TileRegs<ROWS>& t, const float* __restrict__ g, long long ld, int row0, int row_limit,
                                          int k0, int K, int tid) {
  constexpr int kPer = ROWS * 8 / kProducerThreads;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int ch = tid + i * kProducerThreads;
    const int r = ch >> 3, c = ch & 7;
    const int gr = row0 + r, gk = k0 + c * 4;
    t.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < row_limit && gk < K) t.v[i] = __ldg(reinterpret_cast<const float4*>(g + (long long)gr * ld + gk));
  }
}

template <int ROWS>
__device__ __forceinline__ void tile_store(const TileRegs<ROWS>& t, unsigned char* s_hi, unsigned char* s_lo, int tid) {
  constexpr int kPer = ROWS * 8 / kProducerThreads;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int ch = tid + i * kProducerThreads;
    const int r = ch >> 3, c = ch & 7;
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
    const float4 v = t.v[i];
    float4 hi, lo;
    hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
    // lo = x - hi is exact in fp32; the tensor core drops its 13 low mantissa bits itself (error 2^-22 |x|), so the
    // second cvt.rna (four instructions on sm_100a) is not spent on it
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
    *reinterpret_cast<float4*>(s_hi + off) = hi;
    *reinterpret_cast<float4*>(s_lo + off) = lo;
  }
}

// Epilogue of one 128 x BN output tile, executed by the 8 producer warps (256 threads): TMEM -> registers -> shared
// memory transpose -> fused epilogue -> global, plus the optional GroupNorm column statistics.
template <int BN>
__device__ __forceinline__ void tile_epilogue(const Params& p, unsigned char* smem, uint32_t accum_bar, uint32_t tmem_acc, int tid,
                                              int warp, int lane, int m0, int n0) {
  // ------------------------------------------------------------------ epilogue
  mbar_wait(accum_bar, 0);
  tc_fence_after();
  // The operand stages are reused as the transposition buffer below.  The mbarrier chain (producer stores -> full ->
  // MMA -> commit -> accum) already orders every earlier write / tensor-core read before this point; the CTA-scope
  // barrier among the 256 epilogue threads states the same hand-over in a form compute-sanitizer's racecheck can see.
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const int q = warp & 3, half = warp >> 2;
  float* __restrict__ Cp = p.C + (long long)blockIdx.z * p.sC;  // split-K: sC = M*N, raw partial sums
  const float* __restrict__ R = p.residual ? p.residual + (long long)blockIdx.z * p.sR : nullptr;
  const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cp) & 15) == 0) &&
                      (!R || ((p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0)));
  // Each thread holds one accumulator ROW (TMEM lane); storing rows directly would make every warp store hit 32
  // different lines.  The 32x32 chunk is therefore transposed through shared memory (the operand stages are free
  // once accum_bar has fired): rows are written with a 36-float pitch, read back as 4 rows x 32 columns per warp
  // instruction, so that global stores / residual loads are full 128-byte lines.
  float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
#pragma unroll 1
  for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
    uint32_t r[32], rc[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), rc);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 v;
      v.x = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
      v.y = __uint_as_float(r[j + 1]) + __uint_as_float(rc[j + 1]);
      v.z = __uint_as_float(r[j + 2]) + __uint_as_float(rc[j + 2]);
      v.w = __uint_as_float(r[j + 3]) + __uint_as_float(rc[j + 3]);
      *reinterpret_cast<float4*>(stage + lane * 36 + j) = v;
    }
    __syncwarp();
    const int nbase = n0 + c0;
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    const int n = nbase + c4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) {
      if (n + 3 < p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) bv = *reinterpret_cast<const float4*>(p.bias + n);
      else { if (n < p.N) bv.x = p.bias[n]; if (n + 1 < p.N) bv.y = p.bias[n + 1]; if (n + 2 < p.N) bv.z = p.bias[n + 2]; if (n + 3 < p.N) bv.w = p.bias[n + 3]; }
    }
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int rr = 0; rr < 32; rr += 4) {
      const int row = rr + rsub;
      const int m = m0 + q * 32 + row;
      if (m >= p.M) continue;
      const float4 a = *reinterpret_cast<const float4*>(stage + row * 36 + c4);
      float x[4] = {a.x * p.alpha, a.y * p.alpha, a.z * p.alpha, a.w * p.alpha};
      if (p.row_div) { const float rd = p.row_div[m]; x[0] /= rd; x[1] /= rd; x[2] /= rd; x[3] /= rd; }
      x[0] += bv.x; x[1] += bv.y; x[2] += bv.z; x[3] += bv.w;
      if (R) {
        const float* rp = R + (long long)m * p.ldr + n;
        if (vec_ok && n + 3 < p.N) { const float4 t = *reinterpret_cast<const float4*>(rp); x[0] += t.x; x[1] += t.y; x[2] += t.z; x[3] += t.w; }
        else { for (int e = 0; e < 4; ++e) if (n + e < p.N) x[e] += rp[e]; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (p.act == 1) x[e] = fmaxf(x[e], 0.f);
        else if (p.act == 2) x[e] = x[e] > 0.f ? x[e] : 0.1f * x[e];
      }
      float* dst = Cp + (long long)m * p.ldc + n;
      if (vec_ok && n + 3 < p.N) *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
      else { for (int e = 0; e < 4; ++e) if (n + e < p.N) dst[e] = x[e]; }
      if (p.gn_partial) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < p.N) { cs[e] += x[e]; cq[e] = fmaf(x[e], x[e], cq[e]); }
      }
    }
    if (p.gn_partial) {  // column sums over this warp's 32 rows (fp32, at most 8 terms per lane, then 4 lanes)
      float2* gn_col = reinterpret_cast<float2*>(smem + 40 * 1024);  // [4 row quarters][BN]
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);  cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 8);
        cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16); cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 16);
      }
      if (rsub == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) gn_col[q * BN + c0 + c4 + e] = make_float2(cs[e], cq[e]);
      }
    }
  }
  if (p.gn_partial) {
    // fold the four row quarters per column (double, fixed order), then the columns of each group
    const float2* gn_col = reinterpret_cast<const float2*>(smem + 40 * 1024);
    double2* gn_cold = reinterpret_cast<double2*>(smem + 48 * 1024);  // [BN]
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid < BN) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) { const float2 v = gn_col[qq * BN + tid]; a += (double)v.x; b += (double)v.y; }
      gn_cold[tid] = make_double2(a, b);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int cg = p.N / p.gn_groups;  // host guarantees cg | BN and N % BN-aligned groups
    if (tid < BN / cg) {
      const int gidx = n0 / cg + tid;
      if (gidx < p.gn_groups) {
        double a = 0.0, b = 0.0;
        for (int c = tid * cg; c < (tid + 1) * cg; ++c) { a += gn_cold[c].x; b += gn_cold[c].y; }
        p.gn_partial[(long long)blockIdx.y * p.gn_groups + gidx] = make_double2(a, b);
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, BN == 64 ? 2 : 1) gemm_tf32x3_kernel(Params p) {
  pdl_wait();
  pdl_trigger();
  using C = Cfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);  // full[st], empty[st], accum, tmem slot
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * C::kStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool split = p.k_split > 0;
  const int kbeg = split ? blockIdx.z * p.k_split : 0;
  const int kend = split ? min(p.K, kbeg + p.k_split) : p.K;
  const float* __restrict__ A = p.A + (split ? 0ll : (long long)blockIdx.z * p.sA);
  const float* __restrict__ B = p.B + (split ? 0ll : (long long)blockIdx.z * p.sB);
  const int nkb = (kend - kbeg + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar(s), kProducerThreads / 32); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {  // TMEM allocation (power of two >= 32 columns), owned by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ producers
    TileRegs<BM> ra;
    TileRegs<BN> rb;
    const bool packed = p.B_packed != nullptr;
    // CTAs of different M-tiles walk the k-blocks in rotated order: at any instant they pull DIFFERENT weight tiles
    // out of L2 instead of all hammering the same few L2 slices (the sum over k-blocks is order-independent per tile
    // up to fp32 rounding, and stays deterministic)
    const int rot = (int)(blockIdx.y % (unsigned)nkb);
    auto kbr = [&](int kb) { int r = kb + rot; return r >= nkb ? r - nkb : r; };
    tile_load<BM>(ra, A, p.lda, m0, p.M, kbeg + kbr(0) * BK, kend, tid);
    if (!packed) tile_load<BN>(rb, B, p.ldb, n0, p.N, kbeg + kbr(0) * BK, kend, tid);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % C::kStages;
      if (kb >= C::kStages) mbar_wait(empty_bar(s), ((kb / C::kStages) - 1) & 1);
      unsigned char* st = smem + s * C::kStageBytes;
      if (packed && tid == 0) {
        // packed layout: tile (nt, kblock) -> [hi 16 KB][lo 16 KB], nt over 128-row groups; this CTA needs BN rows
        constexpr uint32_t kRowBytes = BN < 128 ? BN * 128 : 128 * 128;  // bytes per copy (rows x 128 B)
        constexpr int kCopies = BN / 128 > 0 ? BN / 128 : 1;
        mbar_expect_tx(full_bar(s), 2u * kCopies * kRowBytes);
        const int kblock = (kbeg / BK) + kbr(kb);
        const uint32_t b_hi_s = smem_u32(st + 2 * C::kABytes), b_lo_s = b_hi_s + C::kBBytes;
#pragma unroll
        for (int c = 0; c < kCopies; ++c) {
          const int row0 = n0 + c * 128;               // first B row of this copy
          const int nt = row0 >> 7, rin = row0 & 127;   // packed tile and row offset inside it (multiple of 8)
          const unsigned char* src = reinterpret_cast<const unsigned char*>(p.B_packed) +
                                     ((size_t)nt * p.packed_kblocks + kblock) * (2 * 16384) + (size_t)rin * 128;
          bulk_copy_g2s(b_hi_s + c * 16384, src, kRowBytes, full_bar(s));
          bulk_copy_g2s(b_lo_s + c * 16384, src + 16384, kRowBytes, full_bar(s));
        }
      }
      TileRegs<BM> ra_next;
      TileRegs<BN> rb_next;
      if (kb + 1 < nkb) {  // prefetch the next k-block before converting this one
        tile_load<BM>(ra_next, A, p.lda, m0, p.M, kbeg + kbr(kb + 1) * BK, kend, tid);
        if (!packed) tile_load<BN>(rb_next, B, p.ldb, n0, p.N, kbeg + kbr(kb + 1) * BK, kend, tid);
      }
      tile_store<BM>(ra, st, st + C::kABytes, tid);
      if (!packed) tile_store<BN>(rb, st + 2 * C::kABytes, st + 2 * C::kABytes + C::kBBytes, tid);
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));  // one arrival per warp: 256 single arrivals on one word serialise
      if (kb + 1 < nkb) {
        ra = ra_next;
        if (!packed) rb = rb_next;
      }
    }
    tile_epilogue<BN>(p, smem, accum_bar, tmem_acc, tid, warp, lane, m0, n0);
    tc_fence_before();
  } else {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32 (bit 4), A=B=tf32 (2 at bits 7 and 10), K-major both, N>>3 at 17, M>>4 at 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::kStages;
        mbar_wait(full_bar(s), (kb / C::kStages) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * C::kStageBytes);
        const uint32_t a_lo = a_hi + C::kABytes;
        const uint32_t b_hi = a_hi + 2 * C::kABytes;
        const uint32_t b_lo = b_hi + C::kBBytes;
#pragma unroll
        for (int k8 = 0; k8 < BK / 8; ++k8) {
          const uint32_t ko = k8 * 32;  // 8 tf32 = 32 bytes along K inside the 128-byte swizzle row
          // two accumulators: columns [0,BN) take the hi*hi products, columns [BN,2BN) the small correction terms.
          // The tensor core truncates on every accumulation, so keeping the large partial sum on a chain of K/8
          // steps (instead of 3K/8) cuts its rounding error threefold; the correction sum is ~2^-11 smaller and
          // its truncation is irrelevant.  The epilogue adds the two.
          const uint32_t first = (kb | k8) != 0 ? 1u : 0u;
          umma_tf32(tmem_acc + BN, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, first);
          umma_tf32(tmem_acc + BN, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
          umma_tf32(tmem_acc, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, first);
        }
        umma_commit(empty_bar(s));  // frees the stage when these MMAs have read it
      }
      umma_commit(accum_bar);  // accumulator complete
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(2 * BN));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-fed variant for products with a PRE-PACKED static B (every Linear / KPConv contraction of the backbone).
//
// Measured on B200 (round 1, profiles/r01h): the A-streaming products of the backbone (e.g. 41907 x 64 x 960, a
// 161 MB operand) ran at ~1.6 TB/s of A traffic -- a quarter of HBM -- because each k-block of A travelled
// ld.global -> registers -> cvt -> st.shared with a single k-block of look-ahead.  Here
//   * warp 9 (one thread) streams A with tensor-map TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a ring of
//     kRaw raw fp32 tiles that runs up to kRaw k-blocks ahead of the tensor core, and the packed B tiles with plain
//     bulk copies into the operand ring;
//   * the raw tile is used AS IS for the hi operand: the tensor core reads only the upper 19 bits of each fp32 word
//     (sign, exponent, 10 mantissa bits), i.e. it sees trunc_tf32(x).  Warps 0-7 only derive the low part
//     lo = x - trunc_tf32(x) (exact in fp32) into the operand ring -- one LDS.128 / STS.128 pair per 4 elements at
//     the SAME byte offset (TMA already wrote the UMMA canonical swizzled layout), no address arithmetic;
//   * warp 8 issues the same three products per k-step as gemm_tf32x3_kernel and releases both rings with
//     tcgen05.commit.
// Out-of-range rows / k are zero-filled by TMA, so edge tiles need no predicates.
template <int BN>
struct CfgTma {
  static constexpr int kABytes = BM * BK * 4;
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kOpBytes = kABytes + 2 * kBBytes;              // a_lo | b_hi | b_lo
  // Ring depths.  ncu (profiles/r02c): the producers of a 15-k-block tile spend most of their time waiting for an
  // operand slot, i.e. for the MMAs of k-block kb - kOps to retire -- the tensor core's issue-to-commit latency, not
  // memory, paces narrow tiles.  BN = 64: two CTAs per SM (3 + 2 stages each, 112 KB) keep four k-blocks in flight
  // per SM and overlap one CTA's epilogue with the other's main loop; BN = 128: 5 + 3 stages; BN = 256: 4 + 2.
  static constexpr int kOps = BN == 128 ? 3 : 2;                       // operand-ring depth (a_lo | b_hi | b_lo)
  static constexpr int kRaw = BN == 256 ? 4 : (BN == 128 ? 5 : 3);     // raw A ring depth (16 KB each)
  static constexpr int kSmemBytes = kRaw * kABytes + kOps * kOpBytes + 1024 /*align*/ + 512 /*barriers*/;
};
constexpr int kThreadsTma = kProducerThreads + 64;

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreadsTma, BN == 64 ? 2 : 1) gemm_tf32x3_tma_kernel(const __grid_constant__ CUtensorMap tmap_a, Params p) {
  using C = CfgTma<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  unsigned char* raw_ring = smem;                                   // [kRaw][16 KB]
  unsigned char* op_ring = smem + C::kRaw * C::kABytes;             // [kOps][a_lo | b_hi | b_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(op_ring + C::kOps * C::kOpBytes);
  const uint32_t bar_base = smem_u32(bars);
  auto raw_full = [&](int r) { return bar_base + 8u * r; };                          // TMA bytes of raw tile r
  auto raw_empty = [&](int r) { return bar_base + 8u * (C::kRaw + r); };             // tcgen05.commit
  auto op_full = [&](int s) { return bar_base + 8u * (2 * C::kRaw + s); };           // 8 splitter warps + B bytes
  auto op_empty = [&](int s) { return bar_base + 8u * (2 * C::kRaw + C::kOps + s); };  // tcgen05.commit
  const uint32_t accum_bar = bar_base + 8u * (2 * C::kRaw + 2 * C::kOps);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kRaw + 2 * C::kOps + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool split = p.k_split > 0;
  const int kbeg = split ? blockIdx.z * p.k_split : 0;
  const int kend = split ? min(p.K, kbeg + p.k_split) : p.K;
  const int nkb = (kend - kbeg + BK - 1) / BK;
  const int rot = (int)(blockIdx.y % (unsigned)nkb);  // see gemm_tf32x3_kernel: de-synchronise the weight stream
  auto kbr = [&](int kb) { int r = kb + rot; return r >= nkb ? r - nkb : r; };

  if (tid == 0) {
    for (int r = 0; r < C::kRaw; ++r) { mbar_init(raw_full(r), 1); mbar_init(raw_empty(r), 1); }
    for (int s = 0; s < C::kOps; ++s) { mbar_init(op_full(s), kProducerThreads / 32 + 1); mbar_init(op_empty(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  // barriers, TMEM and the tensor-map prefetch are set up: only now wait for the producer of A / the residual
  pdl_wait();
  pdl_trigger();

  if (warp < 8) {
    // ------------------------------------------------------------------ low-part derivation, then the epilogue
    for (int kb = 0; kb < nkb; ++kb) {
      const int r = kb % C::kRaw, s = kb % C::kOps;
      if (kb >= C::kOps) mbar_wait(op_empty(s), ((kb / C::kOps) - 1) & 1);  // a_lo slot free again
      mbar_wait(raw_full(r), (kb / C::kRaw) & 1);
      const float4* src = reinterpret_cast<const float4*>(raw_ring + r * C::kABytes);
      float4* dst = reinterpret_cast<float4*>(op_ring + s * C::kOpBytes);
#pragma unroll
      for (int i = 0; i < BM * BK / 4 / kProducerThreads; ++i) {
        const float4 v = src[tid + i * kProducerThreads];
        float4 lo;
        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        dst[tid + i * kProducerThreads] = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(op_full(s));
    }
    tile_epilogue<BN>(p, smem, accum_bar, tmem_acc, tid, warp, lane, m0, n0);
    tc_fence_before();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int r = kb % C::kRaw, s = kb % C::kOps;
        mbar_wait(op_full(s), (kb / C::kOps) & 1);  // implies raw tile r has landed (the splitters read it)
        tc_fence_after();
        const uint32_t a_hi = smem_u32(raw_ring + r * C::kABytes);
        const uint32_t a_lo = smem_u32(op_ring + s * C::kOpBytes);
        const uint32_t b_hi = a_lo + C::kABytes;
        const uint32_t b_lo = b_hi + C::kBBytes;
#pragma unroll
        for (int k8 = 0; k8 < BK / 8; ++k8) {
          const uint32_t ko = k8 * 32;
          const uint32_t first = (kb | k8) != 0 ? 1u : 0u;
          umma_tf32(tmem_acc + BN, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, first);
          umma_tf32(tmem_acc + BN, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
          umma_tf32(tmem_acc, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, first);
        }
        umma_commit(raw_empty(r));
        umma_commit(op_empty(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ loader (one thread): TMA for A, bulk copies for B
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      constexpr uint32_t kRowBytes = BN < 128 ? BN * 128 : 128 * 128;
      constexpr int kCopies = BN / 128 > 0 ? BN / 128 : 1;
      // A runs kRaw - kOps k-blocks ahead of B.  Every wait below is on the MMA of a k-block whose A and B tiles were
      // issued earlier in this same sequence, so the single loader thread can never block itself.
      auto issue_a = [&](int kb) {
        const int r = kb % C::kRaw;
        if (kb >= C::kRaw) mbar_wait(raw_empty(r), ((kb / C::kRaw) - 1) & 1);
        mbar_arrive_expect_tx(raw_full(r), (uint32_t)C::kABytes);
        tma_load_2d(smem_u32(raw_ring + r * C::kABytes), &tmap_a, kbeg + kbr(kb) * BK, m0, raw_full(r));
      };
      auto issue_b = [&](int kb) {
        const int s = kb % C::kOps;
        if (kb >= C::kOps) mbar_wait(op_empty(s), ((kb / C::kOps) - 1) & 1);
        mbar_arrive_expect_tx(op_full(s), 2u * kCopies * kRowBytes);
        const int kblock = (kbeg / BK) + kbr(kb);
        const uint32_t b_hi_s = smem_u32(op_ring + s * C::kOpBytes + C::kABytes), b_lo_s = b_hi_s + C::kBBytes;
#pragma unroll
        for (int c = 0; c < kCopies; ++c) {
          const int row0 = n0 + c * 128;
          const int nt = row0 >> 7, rin = row0 & 127;
          const unsigned char* src = reinterpret_cast<const unsigned char*>(p.B_packed) +
                                     ((size_t)nt * p.packed_kblocks + kblock) * (2 * 16384) + (size_t)rin * 128;
          bulk_copy_g2s(b_hi_s + c * 16384, src, kRowBytes, op_full(s));
          bulk_copy_g2s(b_lo_s + c * 16384, src + 16384, kRowBytes, op_full(s));
        }
      };
      constexpr int kLead = C::kRaw - C::kOps;
      for (int kb = 0; kb < kLead && kb < nkb; ++kb) issue_a(kb);
      for (int kb = 0; kb < nkb; ++kb) {
        issue_b(kb);
        if (kb + kLead < nkb) issue_a(kb + kLead);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(2 * BN));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent form of the TMA-fed kernel: one CTA per SM walks the output tiles of the whole problem.
//
// Measured (tools/gemm_bench.py, round 2): with one tile per CTA a 128 x 256 x 64 product spends ~12 us per tile in
// serial prologue -> first loads -> MMA -> epilogue -> exit, i.e. the small-K products of the backbone (a third of its
// GEMM launches) move their operands at 1.5 TB/s.  Here the roles never stop:
//   warp 9        loader   streams A (tensor-map TMA) and the packed B tiles of tile after tile into the rings,
//   warps 4-7     splitters derive the low parts,
//   warp 8        MMA      accumulates tile i into TMEM buffer i & 1,
//   warps 0-3     epilogue drains buffer (i-1) & 1 (TMEM -> transpose -> fused epilogue / GroupNorm statistics -> global)
// so the epilogue of one tile overlaps the loads and MMAs of the next, and barrier / TMEM / tensor-map set-up happens
// once per SM.  Two accumulator sets of 2 BN columns each need 4 BN <= 512 TMEM columns: BN <= 128 (wider outputs are
// walked as several column tiles; their A tiles are re-read from L2).
template <int BN>
struct CfgPersist {
  static constexpr int kABytes = BM * BK * 4;
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kOpBytes = kABytes + 2 * kBBytes;
  static constexpr int kOps = BN == 128 ? 2 : 3;
  static constexpr int kRaw = BN == 128 ? 5 : 4;  // (a 7-deep raw ring with 2 operand stages measured the same: not load-latency bound)
  static constexpr int kEpiBytes = 8 * 32 * 36 * 4 + 4 * BN * 8 + BN * 16;  // 8 transposition tiles | gn_col | gn_cold
  static constexpr int kSmemBytes = kRaw * kABytes + kOps * kOpBytes + kEpiBytes + 1024 /*align*/ + 512 /*barriers*/;
};
constexpr int kThreadsPersist = 448;  // 8 epilogue warps, 4 splitter warps, MMA warp, loader warp

template <int BN>
__global__ void __launch_bounds__(kThreadsPersist, 1) gemm_tf32x3_persist_kernel(const __grid_constant__ CUtensorMap tmap_a, Params p,
                                                                                int tiles_m, int tiles_n, int n_splits) {
  using C = CfgPersist<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  unsigned char* raw_ring = smem;
  unsigned char* op_ring = smem + C::kRaw * C::kABytes;
  unsigned char* epi = op_ring + C::kOps * C::kOpBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + C::kEpiBytes);
  const uint32_t bar_base = smem_u32(bars);
  auto raw_full = [&](int r) { return bar_base + 8u * r; };
  auto raw_empty = [&](int r) { return bar_base + 8u * (C::kRaw + r); };
  auto op_full = [&](int s) { return bar_base + 8u * (2 * C::kRaw + s); };
  auto op_empty = [&](int s) { return bar_base + 8u * (2 * C::kRaw + C::kOps + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (2 * C::kRaw + 2 * C::kOps + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (2 * C::kRaw + 2 * C::kOps + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kRaw + 2 * C::kOps + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = tiles_m * tiles_n * n_splits;
  const bool split = p.k_split > 0;

  if (tid == 0) {
    for (int r = 0; r < C::kRaw; ++r) { mbar_init(raw_full(r), 1); mbar_init(raw_empty(r), 1); }
    for (int s = 0; s < C::kOps; ++s) { mbar_init(op_full(s), 4 + 1); mbar_init(op_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(4 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  // tile t -> (m tile, n tile, k slice): n fastest, so that the CTAs of one wave share their A rows through L2
  auto tile_coords = [&](int t, int& mt, int& nt, int& z) { nt = t % tiles_n; mt = (t / tiles_n) % tiles_m; z = t / (tiles_n * tiles_m); };
  auto k_range = [&](int z, int& kbeg, int& nkb) {
    kbeg = split ? z * p.k_split : 0;
    const int kend = split ? min(p.K, kbeg + p.k_split) : p.K;
    nkb = (kend - kbeg + BK - 1) / BK;
  };

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue warps: TMEM lane quarter q = warp & 3,
    // column half = warp >> 2 (a warp may only touch the 32 TMEM lanes of its quarter)
    const int q = warp & 3, half = warp >> 2;
    float* stage = reinterpret_cast<float*>(epi) + warp * (32 * 36);
    float2* gn_col = reinterpret_cast<float2*>(epi + 8 * 32 * 36 * 4);   // [4][BN]
    double2* gn_cold = reinterpret_cast<double2*>(epi + 8 * 32 * 36 * 4 + 4 * BN * 8);  // [BN]
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      int mt, nt, z;
      tile_coords(t, mt, nt, z);
      const int m0 = mt * BM, n0 = nt * BN, b = it & 1;
      mbar_wait(acc_full(b), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * 2 * BN);
      float* __restrict__ Cp = p.C + (long long)z * p.sC;
      const float* __restrict__ R = p.residual ? p.residual + (long long)z * p.sR : nullptr;
      const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cp) & 15) == 0) &&
                          (!R || ((p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0)));
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t r[32], rc[32];
        tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), rc);
        if (c0 + 32 >= (half + 1) * (BN / 2)) {  // last TMEM read of this warp: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(b));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v;
          v.x = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
          v.y = __uint_as_float(r[j + 1]) + __uint_as_float(rc[j + 1]);
          v.z = __uint_as_float(r[j + 2]) + __uint_as_float(rc[j + 2]);
          v.w = __uint_as_float(r[j + 3]) + __uint_as_float(rc[j + 3]);
          *reinterpret_cast<float4*>(stage + lane * 36 + j) = v;
        }
        __syncwarp();
        const int c4 = (lane & 7) * 4, rsub = lane >> 3;
        const int n = n0 + c0 + c4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          if (n + 3 < p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) bv = *reinterpret_cast<const float4*>(p.bias + n);
          else { if (n < p.N) bv.x = p.bias[n]; if (n + 1 < p.N) bv.y = p.bias[n + 1]; if (n + 2 < p.N) bv.z = p.bias[n + 2]; if (n + 3 < p.N) bv.w = p.bias[n + 3]; }
        }
        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 32; rr += 4) {
          const int row = rr + rsub;
          const int m = m0 + q * 32 + row;
          if (m >= p.M) continue;
          const float4 a = *reinterpret_cast<const float4*>(stage + row * 36 + c4);
          float x[4] = {a.x * p.alpha, a.y * p.alpha, a.z * p.alpha, a.w * p.alpha};
          if (p.row_div) { const float rd = p.row_div[m]; x[0] /= rd; x[1] /= rd; x[2] /= rd; x[3] /= rd; }
          x[0] += bv.x; x[1] += bv.y; x[2] += bv.z; x[3] += bv.w;
          if (R) {
            const float* rp = R + (long long)m * p.ldr + n;
            if (vec_ok && n + 3 < p.N) { const float4 tt = *reinterpret_cast<const float4*>(rp); x[0] += tt.x; x[1] += tt.y; x[2] += tt.z; x[3] += tt.w; }
            else { for (int e = 0; e < 4; ++e) if (n + e < p.N) x[e] += rp[e]; }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.act == 1) x[e] = fmaxf(x[e], 0.f);
            else if (p.act == 2) x[e] = x[e] > 0.f ? x[e] : 0.1f * x[e];
          }
          float* dst = Cp + (long long)m * p.ldc + n;
          if (vec_ok && n + 3 < p.N) *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
          else { for (int e = 0; e < 4; ++e) if (n + e < p.N) dst[e] = x[e]; }
          if (p.gn_partial) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < p.N) { cs[e] += x[e]; cq[e] = fmaf(x[e], x[e], cq[e]); }
          }
        }
        if (p.gn_partial) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);  cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 8);
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16); cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 16);
          }
          if (rsub == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) gn_col[q * BN + c0 + c4 + e] = make_float2(cs[e], cq[e]);
          }
        }
        __syncwarp();  // the transposition tile is rewritten by the next chunk
      }
      if (p.gn_partial) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (tid < BN) {
          double a = 0.0, bsum = 0.0;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) { const float2 v = gn_col[qq * BN + tid]; a += (double)v.x; bsum += (double)v.y; }
          gn_cold[tid] = make_double2(a, bsum);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const int cg = p.N / p.gn_groups;
        if (tid < BN / cg) {
          const int gidx = n0 / cg + tid;
          if (gidx < p.gn_groups) {
            double a = 0.0, bsum = 0.0;
            for (int c = tid * cg; c < (tid + 1) * cg; ++c) { a += gn_cold[c].x; bsum += gn_cold[c].y; }
            p.gn_partial[(long long)mt * p.gn_groups + gidx] = make_double2(a, bsum);
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");  // gn_col / gn_cold are free for the next tile
      }
    }
    tc_fence_before();
  } else if (warp < 12) {
    // ------------------------------------------------------------------ splitters: lo = x - trunc_tf32(x)
    const int stid = tid - 256;
    int kit = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int mt, nt, z, kbeg, nkb;
      tile_coords(t, mt, nt, z);
      k_range(z, kbeg, nkb);
      for (int kb = 0; kb < nkb; ++kb, ++kit) {
        const int r = kit % C::kRaw, s = kit % C::kOps;
        if (kit >= C::kOps) mbar_wait(op_empty(s), ((kit / C::kOps) - 1) & 1);
        mbar_wait(raw_full(r), (kit / C::kRaw) & 1);
        const float4* src = reinterpret_cast<const float4*>(raw_ring + r * C::kABytes);
        float4* dst = reinterpret_cast<float4*>(op_ring + s * C::kOpBytes);
#pragma unroll
        for (int i = 0; i < BM * BK / 4 / 128; ++i) {
          const float4 v = src[stid + i * 128];
          float4 lo;
          lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          dst[stid + i * 128] = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(op_full(s));
      }
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int kit = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int mt, nt, z, kbeg, nkb;
        tile_coords(t, mt, nt, z);
        k_range(z, kbeg, nkb);
        const int b = it & 1;
        if (it >= 2) mbar_wait(acc_empty(b), ((it >> 1) - 1) & 1);  // the epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(b * 2 * BN);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int r = kit % C::kRaw, s = kit % C::kOps;
          mbar_wait(op_full(s), (kit / C::kOps) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(raw_ring + r * C::kABytes);
          const uint32_t a_lo = smem_u32(op_ring + s * C::kOpBytes);
          const uint32_t b_hi = a_lo + C::kABytes;
          const uint32_t b_lo = b_hi + C::kBBytes;
#pragma unroll
          for (int k8 = 0; k8 < BK / 8; ++k8) {
            const uint32_t ko = k8 * 32;
            const uint32_t first = (kb | k8) != 0 ? 1u : 0u;
            umma_tf32(tacc + BN, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, first);
            umma_tf32(tacc + BN, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
            umma_tf32(tacc, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, first);
          }
          umma_commit(raw_empty(r));
          umma_commit(op_empty(s));
        }
        umma_commit(acc_full(b));
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ loader
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      constexpr uint32_t kRowBytes = BN < 128 ? BN * 128 : 128 * 128;
      // flattened k-block stream over this CTA's tiles; A runs kLead k-blocks ahead of B (see gemm_tf32x3_tma_kernel)
      constexpr int kLead = C::kRaw - C::kOps >= 1 ? C::kRaw - C::kOps : 0;
      struct Cursor { int t, kb, nkb, kbeg, m0, n0, rot; };
      auto open_tile = [&](Cursor& c) {
        if (c.t >= total_tiles) { c.nkb = 0; return; }
        int mt, nt, z;
        tile_coords(c.t, mt, nt, z);
        k_range(z, c.kbeg, c.nkb);
        c.m0 = mt * BM; c.n0 = nt * BN; c.kb = 0;
        c.rot = mt % c.nkb;
      };
      auto advance = [&](Cursor& c) {
        if (++c.kb >= c.nkb) { c.t += gridDim.x; open_tile(c); }
      };
      auto kbr = [&](const Cursor& c) { int r = c.kb + c.rot; return r >= c.nkb ? r - c.nkb : r; };
      Cursor ca, cb;
      ca.t = cb.t = blockIdx.x;
      open_tile(ca); open_tile(cb);
      int ia = 0, ib = 0;  // issued k-blocks
      auto issue_a = [&]() {
        const int r = ia % C::kRaw;
        if (ia >= C::kRaw) mbar_wait(raw_empty(r), ((ia / C::kRaw) - 1) & 1);
        mbar_arrive_expect_tx(raw_full(r), (uint32_t)C::kABytes);
        tma_load_2d(smem_u32(raw_ring + r * C::kABytes), &tmap_a, ca.kbeg + kbr(ca) * BK, ca.m0, raw_full(r));
        ++ia; advance(ca);
      };
      auto issue_b = [&]() {
        const int s = ib % C::kOps;
        if (ib >= C::kOps) mbar_wait(op_empty(s), ((ib / C::kOps) - 1) & 1);
        mbar_arrive_expect_tx(op_full(s), 2u * kRowBytes);
        const int kblock = (cb.kbeg / BK) + kbr(cb);
        const uint32_t b_hi_s = smem_u32(op_ring + s * C::kOpBytes + C::kABytes), b_lo_s = b_hi_s + C::kBBytes;
        const int nt = cb.n0 >> 7, rin = cb.n0 & 127;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.B_packed) +
                                   ((size_t)nt * p.packed_kblocks + kblock) * (2 * 16384) + (size_t)rin * 128;
        bulk_copy_g2s(b_hi_s, src, kRowBytes, op_full(s));
        bulk_copy_g2s(b_lo_s, src + 16384, kRowBytes, op_full(s));
        ++ib; advance(cb);
      };
      for (int i = 0; i < kLead && ca.nkb > 0; ++i) issue_a();
      while (cb.nkb > 0) {
        issue_b();
        if (ca.nkb > 0) issue_a();
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(4 * BN));
  }
}

// ---- the same persistent kernel with fp16-split operands ("3xFP16") ---------------------------------------------------
// A tcgen05.mma dispatch costs ~80 cycles on these tiles whatever its kind (measured with GAUSSREG_GEMM_DEBUG: the
// operand fetch from shared memory is 32 bytes per row either way), and kind::f16 covers K = 16 per dispatch where
// kind::tf32 covers 8.  fp16 and TF32 carry the same 11 significant bits, so x = hi + lo in two fp16 parts is as
// accurate as the TF32 split provided the values sit in fp16's range: weights are pre-scaled by a power of two at
// packing time (divided out through alpha), activations are O(1..1e3) after GroupNorm / neighbourhood sums, and the
// splitter raises a device flag (gr_gemm_f16_overflow_ptr) should it ever meet |x| > 6e4.
// k-block = 64 K elements = two TMA boxes of 128 x 32 fp32 (raw ring of boxes) -> the splitter warps convert them into
// one fp16 hi tile and one fp16 lo tile (128-byte rows, SWIZZLE_128B) -> 4 k-steps x 3 MMAs.  The raw boxes are released
// by the splitters (the tensor core never reads them), the operand stages by the MMA commits.
template <int BN>
struct CfgPersist16 {
  static constexpr int kBoxBytes = BM * BK * 4;          // one raw TMA box: 128 rows x 32 fp32 = 16 KB
  static constexpr int kABytes = BM * 128;               // one fp16 operand tile of A: 128 rows x 64 halves = 16 KB
  static constexpr int kBBytes = BN * 128;
  static constexpr int kOpBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kOps = 2;
  static constexpr int kRaw = BN == 128 ? 3 : 4;         // boxes; a k-block takes two consecutive ones
  static constexpr int kEpiBytes = 8 * 32 * 36 * 4 + 4 * BN * 8 + BN * 16;
  static constexpr int kSmemBytes = kRaw * kBoxBytes + kOps * kOpBytes + kEpiBytes + 1024 /*align*/ + 512 /*barriers*/;
};
__device__ int g_f16_overflow;  // set when a splitter met |x| > 6e4 (fp16 tops out at 65504)


template <int BN>
__global__ void __launch_bounds__(kThreadsPersist, 1) gemm_f16x3_persist_kernel(const __grid_constant__ CUtensorMap tmap_a, Params p,
                                                                                int tiles_m, int tiles_n, int n_splits) {
  using C = CfgPersist16<BN>;
  constexpr int BK16 = 64;  // K elements per k-block
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  unsigned char* raw_ring = smem;
  unsigned char* op_ring = smem + C::kRaw * C::kBoxBytes;
  unsigned char* epi = op_ring + C::kOps * C::kOpBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + C::kEpiBytes);
  const uint32_t bar_base = smem_u32(bars);
  auto raw_full = [&](int r) { return bar_base + 8u * r; };
  auto raw_empty = [&](int r) { return bar_base + 8u * (C::kRaw + r); };
  auto op_full = [&](int s) { return bar_base + 8u * (2 * C::kRaw + s); };
  auto op_empty = [&](int s) { return bar_base + 8u * (2 * C::kRaw + C::kOps + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (2 * C::kRaw + 2 * C::kOps + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (2 * C::kRaw + 2 * C::kOps + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kRaw + 2 * C::kOps + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = tiles_m * tiles_n * n_splits;
  const bool split = p.k_split > 0;

  if (tid == 0) {
    for (int r = 0; r < C::kRaw; ++r) { mbar_init(raw_full(r), 1); mbar_init(raw_empty(r), 4); }
    for (int s = 0; s < C::kOps; ++s) { mbar_init(op_full(s), 4 + 1); mbar_init(op_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(4 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  // tile t -> (m tile, n tile, k slice): n fastest, so that the CTAs of one wave share their A rows through L2
  auto tile_coords = [&](int t, int& mt, int& nt, int& z) { nt = t % tiles_n; mt = (t / tiles_n) % tiles_m; z = t / (tiles_n * tiles_m); };
  auto k_range = [&](int z, int& kbeg, int& nkb) {
    kbeg = split ? z * p.k_split : 0;
    const int kend = split ? min(p.K, kbeg + p.k_split) : p.K;
    nkb = (kend - kbeg + BK16 - 1) / BK16;
  };

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue warps: TMEM lane quarter q = warp & 3,
    // column half = warp >> 2 (a warp may only touch the 32 TMEM lanes of its quarter)
    const int q = warp & 3, half = warp >> 2;
    float* stage = reinterpret_cast<float*>(epi) + warp * (32 * 36);
    float2* gn_col = reinterpret_cast<float2*>(epi + 8 * 32 * 36 * 4);   // [4][BN]
    double2* gn_cold = reinterpret_cast<double2*>(epi + 8 * 32 * 36 * 4 + 4 * BN * 8);  // [BN]
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      int mt, nt, z;
      tile_coords(t, mt, nt, z);
      const int m0 = mt * BM, n0 = nt * BN, b = it & 1;
      mbar_wait(acc_full(b), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * 2 * BN);
      float* __restrict__ Cp = p.C + (long long)z * p.sC;
      const float* __restrict__ R = p.residual ? p.residual + (long long)z * p.sR : nullptr;
      const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cp) & 15) == 0) &&
                          (!R || ((p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0)));
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t r[32], rc[32];
        tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), rc);
        if (c0 + 32 >= (half + 1) * (BN / 2)) {  // last TMEM read of this warp: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(b));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v;
          v.x = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
          v.y = __uint_as_float(r[j + 1]) + __uint_as_float(rc[j + 1]);
          v.z = __uint_as_float(r[j + 2]) + __uint_as_float(rc[j + 2]);
          v.w = __uint_as_float(r[j + 3]) + __uint_as_float(rc[j + 3]);
          *reinterpret_cast<float4*>(stage + lane * 36 + j) = v;
        }
        __syncwarp();
        const int c4 = (lane & 7) * 4, rsub = lane >> 3;
        const int n = n0 + c0 + c4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          if (n + 3 < p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) bv = *reinterpret_cast<const float4*>(p.bias + n);
          else { if (n < p.N) bv.x = p.bias[n]; if (n + 1 < p.N) bv.y = p.bias[n + 1]; if (n + 2 < p.N) bv.z = p.bias[n + 2]; if (n + 3 < p.N) bv.w = p.bias[n + 3]; }
        }
        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 32; rr += 4) {
          const int row = rr + rsub;
          const int m = m0 + q * 32 + row;
          if (m >= p.M) continue;
          const float4 a = *reinterpret_cast<const float4*>(stage + row * 36 + c4);
          float x[4] = {a.x * p.alpha, a.y * p.alpha, a.z * p.alpha, a.w * p.alpha};
          if (p.row_div) { const float rd = p.row_div[m]; x[0] /= rd; x[1] /= rd; x[2] /= rd; x[3] /= rd; }
          x[0] += bv.x; x[1] += bv.y; x[2] += bv.z; x[3] += bv.w;
          if (R) {
            const float* rp = R + (long long)m * p.ldr + n;
            if (vec_ok && n + 3 < p.N) { const float4 tt = *reinterpret_cast<const float4*>(rp); x[0] += tt.x; x[1] += tt.y; x[2] += tt.z; x[3] += tt.w; }
            else { for (int e = 0; e < 4; ++e) if (n + e < p.N) x[e] += rp[e]; }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.act == 1) x[e] = fmaxf(x[e], 0.f);
            else if (p.act == 2) x[e] = x[e] > 0.f ? x[e] : 0.1f * x[e];
          }
          float* dst = Cp + (long long)m * p.ldc + n;
          if (vec_ok && n + 3 < p.N) *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
          else { for (int e = 0; e < 4; ++e) if (n + e < p.N) dst[e] = x[e]; }
          if (p.gn_partial) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < p.N) { cs[e] += x[e]; cq[e] = fmaf(x[e], x[e], cq[e]); }
          }
        }
        if (p.gn_partial) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);  cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 8);
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16); cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 16);
          }
          if (rsub == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) gn_col[q * BN + c0 + c4 + e] = make_float2(cs[e], cq[e]);
          }
        }
        __syncwarp();  // the transposition tile is rewritten by the next chunk
      }
      if (p.gn_partial) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (tid < BN) {
          double a = 0.0, bsum = 0.0;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) { const float2 v = gn_col[qq * BN + tid]; a += (double)v.x; bsum += (double)v.y; }
          gn_cold[tid] = make_double2(a, bsum);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const int cg = p.N / p.gn_groups;
        if (tid < BN / cg) {
          const int gidx = n0 / cg + tid;
          if (gidx < p.gn_groups) {
            double a = 0.0, bsum = 0.0;
            for (int c = tid * cg; c < (tid + 1) * cg; ++c) { a += gn_cold[c].x; bsum += gn_cold[c].y; }
            p.gn_partial[(long long)mt * p.gn_groups + gidx] = make_double2(a, bsum);
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");  // gn_col / gn_cold are free for the next tile
      }
    }
    tc_fence_before();
  } else if (warp < 12) {
    // ------------------------------------------------------------------ splitters: two fp32 boxes -> fp16 hi / lo tiles
    const int stid = tid - 256;
    int kit = 0;
    bool bad = false;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int mt, nt, z, kbeg, nkb;
      tile_coords(t, mt, nt, z);
      k_range(z, kbeg, nkb);
      for (int kb = 0; kb < nkb; ++kb, ++kit) {
        const int s = kit % C::kOps;
        const int b0 = 2 * kit, b1 = 2 * kit + 1;  // box counters
        const int r0 = b0 % C::kRaw, r1 = b1 % C::kRaw;
        if (kit >= C::kOps) mbar_wait(op_empty(s), ((kit / C::kOps) - 1) & 1);
        mbar_wait(raw_full(r0), (b0 / C::kRaw) & 1);
        mbar_wait(raw_full(r1), (b1 / C::kRaw) & 1);
        unsigned char* hi_tile = op_ring + s * C::kOpBytes;
        unsigned char* lo_tile = hi_tile + C::kABytes;
#pragma unroll
        for (int i = 0; i < BM * 8 / 128; ++i) {
          // task = (row, output chunk oc of 8 halves): K elements 8 oc .. 8 oc + 7 = fp32 chunks 2q, 2q+1 of box oc >> 2
          const int e = stid + i * 128, row = e >> 3, oc = e & 7, q = oc & 3;
          const unsigned char* box = raw_ring + ((oc >> 2) ? r1 : r0) * C::kBoxBytes + (row >> 3) * 1024 + (row & 7) * 128;
          const float4 v0 = *reinterpret_cast<const float4*>(box + (((2 * q) ^ (row & 7)) << 4));
          const float4 v1 = *reinterpret_cast<const float4*>(box + (((2 * q + 1) ^ (row & 7)) << 4));
          const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const __half2 h = __floats2half2_rn(x[2 * c], x[2 * c + 1]);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(x[2 * c] - hf.x, x[2 * c + 1] - hf.y);
            hi[c] = *reinterpret_cast<const uint32_t*>(&h);
            lo[c] = *reinterpret_cast<const uint32_t*>(&l);
          }
          bad |= fmaxf(fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))),
                       fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7])))) > 6.0e4f;
          const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((oc ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(op_full(s)); mbar_arrive(raw_empty(r0)); mbar_arrive(raw_empty(r1)); }
      }
    }
    if (bad) atomicExch(&g_f16_overflow, 1);
  } else if (warp == 12) {
    // ------------------------------------------------------------------ MMA issuer: the WHOLE warp walks the loop with
    // warp-uniform values (descriptors live in uniform registers), one elected lane issues the tcgen05 instructions
    {
      // kind::f16: c_format F32 (1 << 4), a/b_format F16 (0), K-major both
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int kit = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int mt, nt, z, kbeg, nkb;
        tile_coords(t, mt, nt, z);
        k_range(z, kbeg, nkb);
        const int b = it & 1;
        if (it >= 2) mbar_wait(acc_empty(b), ((it >> 1) - 1) & 1);  // the epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(b * 2 * BN);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % C::kOps;
          mbar_wait(op_full(s), (kit / C::kOps) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(op_ring + s * C::kOpBytes);
          const uint32_t a_lo = a_hi + C::kABytes, b_hi = a_lo + C::kABytes, b_lo = b_hi + C::kBBytes;
          // the start-address field is the low 14 bits (address >> 4): a k-step of 16 halves = 32 bytes adds 2
          const uint64_t d_ah = make_desc(a_hi), d_al = make_desc(a_lo), d_bh = make_desc(b_hi), d_bl = make_desc(b_lo);
          if (elect_one_sync()) {
#pragma unroll
            for (int k16 = 0; k16 < BK16 / 16; ++k16) {
              const uint32_t first = (kb | k16) != 0 ? 1u : 0u;
              umma_f16(tacc + BN, d_al + 2 * k16, d_bh + 2 * k16, idesc, first);
              umma_f16(tacc + BN, d_ah + 2 * k16, d_bl + 2 * k16, idesc, 1u);
              umma_f16(tacc, d_ah + 2 * k16, d_bh + 2 * k16, idesc, first);
            }
            umma_commit(op_empty(s));
          }
          __syncwarp();
        }
        if (elect_one_sync()) umma_commit(acc_full(b));
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ loader
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      constexpr uint32_t kRowBytes = BN < 128 ? BN * 128 : 128 * 128;
      // flattened k-block stream over this CTA's tiles; A runs kLead k-blocks ahead of B (see gemm_tf32x3_tma_kernel)
      constexpr int kLead = 1;  // A (raw boxes, released early by the splitters) runs one k-block ahead of B
      struct Cursor { int t, kb, nkb, kbeg, m0, n0, rot; };
      auto open_tile = [&](Cursor& c) {
        if (c.t >= total_tiles) { c.nkb = 0; return; }
        int mt, nt, z;
        tile_coords(c.t, mt, nt, z);
        k_range(z, c.kbeg, c.nkb);
        c.m0 = mt * BM; c.n0 = nt * BN; c.kb = 0;
        c.rot = mt % c.nkb;
      };
      auto advance = [&](Cursor& c) {
        if (++c.kb >= c.nkb) { c.t += gridDim.x; open_tile(c); }
      };
      auto kbr = [&](const Cursor& c) { int r = c.kb + c.rot; return r >= c.nkb ? r - c.nkb : r; };
      Cursor ca, cb;
      ca.t = cb.t = blockIdx.x;
      open_tile(ca); open_tile(cb);
      int ia = 0, ib = 0;  // issued k-blocks
      auto issue_a = [&]() {  // one k-block = two boxes
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int bi = 2 * ia + j, r = bi % C::kRaw;
          if (bi >= C::kRaw) mbar_wait(raw_empty(r), ((bi / C::kRaw) - 1) & 1);
          mbar_arrive_expect_tx(raw_full(r), (uint32_t)C::kBoxBytes);
          tma_load_2d(smem_u32(raw_ring + r * C::kBoxBytes), &tmap_a, ca.kbeg + kbr(ca) * BK16 + BK * j, ca.m0, raw_full(r));
        }
        ++ia; advance(ca);
      };
      auto issue_b = [&]() {
        const int s = ib % C::kOps;
        if (ib >= C::kOps) mbar_wait(op_empty(s), ((ib / C::kOps) - 1) & 1);
        mbar_arrive_expect_tx(op_full(s), 2u * kRowBytes);
        const int kblock = (cb.kbeg / BK16) + kbr(cb);
        const uint32_t b_hi_s = smem_u32(op_ring + s * C::kOpBytes + 2 * C::kABytes), b_lo_s = b_hi_s + C::kBBytes;
        const int nt = cb.n0 >> 7, rin = cb.n0 & 127;
        const unsigned char* src = p.B_packed16 + ((size_t)nt * p.packed_kblocks64 + kblock) * (2 * 16384) + (size_t)rin * 128;
        bulk_copy_g2s(b_hi_s, src, kRowBytes, op_full(s));
        bulk_copy_g2s(b_lo_s, src + 16384, kRowBytes, op_full(s));
        ++ib; advance(cb);
      };
      for (int i = 0; i < kLead && ca.nkb > 0; ++i) issue_a();
      while (cb.nkb > 0) {
        issue_b();
        if (ca.nkb > 0) issue_a();
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(4 * BN));
  }
}


// C = epilogue(sum_z partial[z]) in a fixed order; one thread per 4 consecutive columns (N % 4 == 0)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, Params p) {
  pdl_wait();
  pdl_trigger();
  const int n4 = p.N >> 2;
  const long long total = (long long)p.M * n4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int m = (int)(i / n4), n = (int)(i % n4) * 4;
  const long long plane = (long long)p.M * p.N;
  float4 acc = reinterpret_cast<const float4*>(partial)[i];
  for (int z = 1; z < splits; ++z) {
    const float4 v = reinterpret_cast<const float4*>(partial + z * plane)[i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  const float rd = p.row_div ? p.row_div[m] : 1.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = v[j] * p.alpha;
    if (p.row_div) x = x / rd;
    if (p.bias) x += p.bias[n + j];
    if (p.residual) x += p.residual[(long long)m * p.ldr + n + j];
    if (p.act == 1) x = fmaxf(x, 0.f);
    else if (p.act == 2) x = x > 0.f ? x : 0.1f * x;
    p.C[(long long)m * p.ldc + n + j] = x;
  }
}

// Split-K reduction that also carries the GroupNorm statistics of its output: a block owns kRedRows rows and all N
// columns (thread = 4 consecutive columns, remaining thread bits walk rows); per-thread fp32 sums over at most 32
// rows, folded in double in a fixed order -> gn_partial[blockIdx.x * G + g].
constexpr int kRedRows = kGnReduceRows;
__global__ void __launch_bounds__(256) splitk_reduce_gn_kernel(const float* __restrict__ partial, int splits, Params p, int col_chunk) {
  pdl_wait();
  pdl_trigger();
  // block (x, y): rows [32 x, 32 x + 32), columns [y * col_chunk, (y + 1) * col_chunk); col_chunk is a multiple of the
  // group width, so every group's statistics come from exactly one column block
  extern __shared__ double2 red_sh[];  // chs[col_chunk], then stage[R * col_chunk] when several row lanes share a column
  const int cbeg = blockIdx.y * col_chunk;
  const int ncol = min(col_chunk, p.N - cbeg);
  double2* chs = red_sh;
  double2* stage = red_sh + col_chunk;
  const int r0 = blockIdx.x * kRedRows, r1 = min(p.M, r0 + kRedRows);
  const int tid = threadIdx.x;
  const int c4n = ncol >> 2;
  const int lanes = c4n < 256 ? c4n : 256;
  const int R = 256 / lanes;
  const int rs = tid / lanes;
  const int n4 = p.N >> 2;
  const long long plane = (long long)p.M * p.N;
  if (rs < R) {
    for (int cc = tid % lanes; cc < c4n; cc += lanes) {
      const int n = cbeg + cc * 4;
      float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) bv = make_float4(p.bias[n], p.bias[n + 1], p.bias[n + 2], p.bias[n + 3]);
      for (int r = r0 + rs; r < r1; r += R) {
        const long long i = (long long)r * n4 + (n >> 2);
        float4 acc = reinterpret_cast<const float4*>(partial)[i];
        for (int z = 1; z < splits; ++z) {
          const float4 v = reinterpret_cast<const float4*>(partial + z * plane)[i];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float v[4] = {acc.x, acc.y, acc.z, acc.w};
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
        const float rd = p.row_div ? p.row_div[r] : 1.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x = v[j] * p.alpha;
          if (p.row_div) x = x / rd;
          if (p.bias) x += b4[j];
          if (p.residual) x += p.residual[(long long)r * p.ldr + n + j];
          if (p.act == 1) x = fmaxf(x, 0.f);
          else if (p.act == 2) x = x > 0.f ? x : 0.1f * x;
          v[j] = x;
          fs[j] += x; fq[j] = fmaf(x, x, fq[j]);
        }
        float* dst = p.C + (long long)r * p.ldc + n;
        if ((p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        else { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; dst[3] = v[3]; }
      }
      double2* dst = (R == 1) ? chs + 4 * cc : stage + (size_t)rs * col_chunk + 4 * cc;
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k] = make_double2((double)fs[k], (double)fq[k]);
    }
  }
  __syncthreads();
  if (R > 1) {
    for (int c = tid; c < ncol; c += 256) {
      double2 a = stage[c];
      for (int k = 1; k < R; ++k) { const double2 b = stage[(size_t)k * col_chunk + c]; a.x += b.x; a.y += b.y; }
      chs[c] = a;
    }
    __syncthreads();
  }
  const int cg = p.N / p.gn_groups;
  for (int gl = tid; gl < ncol / cg; gl += 256) {
    double a = 0.0, b = 0.0;
    for (int c = gl * cg; c < (gl + 1) * cg; ++c) { a += chs[c].x; b += chs[c].y; }
    p.gn_partial[(long long)blockIdx.x * p.gn_groups + cbeg / cg + gl] = make_double2(a, b);
  }
}

template <int BN>
static int launch(const Params& p, int batch, cudaStream_t st) {
  using C = Cfg<BN>;
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(gemm_tf32x3_kernel<BN>), C::kSmemBytes));
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, batch);
  GR_CHECK_CUDA(launch_pdl(gemm_tf32x3_kernel<BN>, dim3(grid), dim3(kThreads), (size_t)(C::kSmemBytes), st, p));
  GR_CHECK_LAUNCH("gemm_tf32x3_kernel");
  return GR_OK;
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeFn>(ptr);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// > 0: the TMA path cannot serve this call (caller falls back to the ld.global producers)
template <int BN>
static int launch_tma(const Params& p, int zdim, cudaStream_t st) {
  using C = CfgTma<BN>;
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return 1;
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.lda * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  static int promo = -1;
  if (promo < 0) { const char* e = getenv("GAUSSREG_TMA_L2PROMO"); promo = e ? atoi(e) : 256; }
  const CUtensorMapL2promotion l2p = promo >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                   : (promo >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.A), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(gemm_tf32x3_tma_kernel<BN>), C::kSmemBytes));
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, zdim);
  GR_CHECK_CUDA(launch_pdl(gemm_tf32x3_tma_kernel<BN>, grid, dim3(kThreadsTma), (size_t)C::kSmemBytes, st, tm, p));
  GR_CHECK_LAUNCH("gemm_tf32x3_tma_kernel");
  return GR_OK;
}

// > 0: the persistent path cannot serve this call
template <int BN>
static int launch_persist(const Params& p, int zdim, cudaStream_t st) {
  using C = CfgPersist<BN>;
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return 1;
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.lda * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.A), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(gemm_tf32x3_persist_kernel<BN>), C::kSmemBytes));
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const long long total = (long long)tiles_m * tiles_n * zdim;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(total < sms ? total : sms);
  GR_CHECK_CUDA(launch_pdl(gemm_tf32x3_persist_kernel<BN>, dim3(grid), dim3(kThreadsPersist), (size_t)C::kSmemBytes, st, tm, p, tiles_m,
                           tiles_n, zdim));
  GR_CHECK_LAUNCH("gemm_tf32x3_persist_kernel");
  return GR_OK;
}

// B (N, K) row-major fp32 -> per (128-row tile, 64-wide k-block): [hi 16 KB][lo 16 KB] of fp16 (B * scale), rows of 128 bytes
// in the K-major SWIZZLE_128B order; zero padding beyond (N, K).
__global__ void __launch_bounds__(256) pack_weight_f16x3_kernel(const float* __restrict__ W, int N, int K, float scale,
                                                                unsigned char* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int nt = blockIdx.y, kb = blockIdx.x;
  const int kblocks = (K + 63) / 64;
  unsigned char* base = out + ((size_t)nt * kblocks + kb) * (2 * 16384);
  for (int ch = threadIdx.x; ch < 128 * 8; ch += blockDim.x) {
    const int r = ch >> 3, c = ch & 7;
    const int gn = nt * 128 + r, gk = kb * 64 + c * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = 0.f, v1 = 0.f;
      if (gn < N && gk + 2 * e < K) v0 = W[(size_t)gn * K + gk + 2 * e] * scale;
      if (gn < N && gk + 2 * e + 1 < K) v1 = W[(size_t)gn * K + gk + 2 * e + 1] * scale;
      const __half2 h = __floats2half2_rn(v0, v1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&h);
      lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// > 0: the fp16 persistent path cannot serve this call
template <int BN>
static int launch_persist16(const Params& p, int zdim, cudaStream_t st) {
  using C = CfgPersist16<BN>;
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return 1;
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.lda * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.A), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(gemm_f16x3_persist_kernel<BN>), C::kSmemBytes));
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const long long total = (long long)tiles_m * tiles_n * zdim;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(total < sms ? total : sms);
  Params pp = p;
  pp.alpha = p.alpha * p.inv_scale16;  // a power of two: exact
  GR_CHECK_CUDA(launch_pdl(gemm_f16x3_persist_kernel<BN>, dim3(grid), dim3(kThreadsPersist), (size_t)C::kSmemBytes, st, tm, pp, tiles_m,
                           tiles_n, zdim));
  GR_CHECK_LAUNCH("gemm_f16x3_persist_kernel");
  return GR_OK;
}

static bool use_f16() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_GEMM_F16"); v = e ? atoi(e) : 0; }  // opt-in, see ops._gemm_f16
  return v != 0;
}

static bool use_tma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_GEMM_TMA"); v = e ? atoi(e) : 1; }
  return v != 0;
}

// one launch of the tile kernel: TMA-fed when B is pre-packed, ld.global producers otherwise
static int launch_any(int bn, const Params& p, int zdim, cudaStream_t st) {
  if (p.B_packed16 && use_f16() && use_tma() && p.K >= 64 && (p.k_split == 0 || p.k_split % 64 == 0)) {
    const int rc16 = p.N <= 64 ? launch_persist16<64>(p, zdim, st) : launch_persist16<128>(p, zdim, st);
    if (rc16 <= 0) return rc16;
  }
  if (p.B_packed && use_tma() && p.K >= 2 * BK) {
    static int persist = -1;
    if (persist < 0) { const char* e = getenv("GAUSSREG_GEMM_PERSIST"); persist = e ? atoi(e) : 1; }
    // The persistent kernel overlaps one tile's epilogue with the next tile's MMAs, which pays when a CTA gets
    // several short tiles (narrow N or small K: 49 vs 61 us on 60000x32x480, 30 vs 34 us on 41907x256x64).  Wide
    // products with a long K loop stay on the 128x256 tiles (twice the flops per A byte read from shared memory),
    // and grids of at most ~one tile per SM have nothing to overlap.  persist=2 forces it for every shape.
    const long long tiles = (long long)((p.M + BM - 1) / BM) * ((p.N + (bn > 128 ? 128 : bn) - 1) / (bn > 128 ? 128 : bn)) * zdim;
    const bool pays = tiles > 222 && (p.N <= 128 || p.K <= 128);
    if (persist > 1 || (persist == 1 && pays)) {
      const int rcp = bn == 64 ? launch_persist<64>(p, zdim, st) : launch_persist<128>(p, zdim, st);
      if (rcp <= 0) return rcp;
    }
    const int rc = bn == 256 ? launch_tma<256>(p, zdim, st) : (bn == 128 ? launch_tma<128>(p, zdim, st) : launch_tma<64>(p, zdim, st));
    if (rc <= 0) return rc;
  }
  return bn == 256 ? launch<256>(p, zdim, st) : (bn == 128 ? launch<128>(p, zdim, st) : launch<64>(p, zdim, st));
}

}  // namespace tc

// Grow-only scratch for the split-K partial sums, one buffer per (device, stream): reuse is stream-ordered, so
// successive calls on the same stream may share it without synchronisation.
static float* splitk_scratch(cudaStream_t st, size_t bytes) {
  struct Entry { int dev; cudaStream_t st; float* ptr; size_t bytes; };
  static std::mutex mu;
  static std::vector<Entry> entries;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : entries) {
    if (e.dev == dev && e.st == st) {
      if (e.bytes >= bytes) return e.ptr;
      cudaStreamSynchronize(st);
      cudaFree(e.ptr);
      e.ptr = nullptr; e.bytes = 0;
      const size_t want = bytes + bytes / 2;
      if (cudaMalloc(&e.ptr, want) != cudaSuccess) { set_last_error("split-K scratch cudaMalloc", cudaGetLastError()); return nullptr; }
      e.bytes = want;
      return e.ptr;
    }
  }
  Entry e{dev, st, nullptr, 0};
  const size_t want = bytes < (size_t(64) << 20) ? (size_t(64) << 20) : bytes + bytes / 2;
  if (cudaMalloc(&e.ptr, want) != cudaSuccess) { set_last_error("split-K scratch cudaMalloc", cudaGetLastError()); return nullptr; }
  e.bytes = want;
  entries.push_back(e);
  return e.ptr;
}

static bool use_bn256() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_BN256"); v = e ? atoi(e) : 1; }
  return v != 0;
}

// Returns GR_OK when the tensor-core path ran, 1 when the problem does not qualify (caller falls back to SIMT).
int gemm_tf32x3(const float* A, long long lda, long long sA, const float* B, long long ldb, long long sB, float* C, long long ldc,
                long long sC, int M, int N, int K, int batch, float alpha, const float* bias, const float* row_div,
                const float* residual, long long ldr, long long sR, int act, cudaStream_t st, const float* B_packed,
                GnStatsOut* gn, const void* B_packed16, float inv_scale16) {
  if (gn) gn->nblk = 0;
  const bool aligned = (lda % 4 == 0) && (ldb % 4 == 0) && (K % 4 == 0) && (sA % 4 == 0) && (sB % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  if (!aligned || N < 32 || (long long)M * N * K < (1ll << 22)) return 1;
  // superpoint-sized products without GroupNorm statistics (the transformer): a dozen 128-row tiles leave most SMs
  // idle; the latency-optimised mma.sync kernels of gemm.cu (32x32 / 64x64 tiles) are faster there
  if (!gn && M < 1536 && K <= 1024) return 1;
  static int kSlice = 0, kSched = 1;
  if (kSlice == 0) {
    const char* e = getenv("GAUSSREG_KSLICE");  // longest K chained into one TMEM accumulator (multiple of 32); default 768
    kSlice = e ? atoi(e) : 768;
    if (kSlice < 32 || kSlice % 32 != 0) kSlice = 768;
    e = getenv("GAUSSREG_SPLITK_SCHED");        // 1: wave-aware slice count (default), 0: slice only for accuracy
    kSched = e ? atoi(e) : 1;
  }
  const int bn = (N > 128 && use_bn256()) ? 256 : (N > 64 ? 128 : 64);
  const long long tiles = (long long)((M + 127) / 128) * ((N + bn - 1) / bn) * batch;
  // ---- how many K slices?
  // (a) accuracy: the tensor core's fp32 accumulation truncates, so the error grows with the number of MMA steps
  //     chained into one accumulator; K beyond 1.5 x kSlice is cut into slices of at most kSlice, each accumulated
  //     in its own TMEM tile, and the slices are summed in fp32 (round-to-nearest) by splitk_reduce_kernel.
  // (b) occupancy: a 128-row tile grid rarely matches 148 SMs.  Among the slice counts allowed by (a), take the
  //     one that minimises  waves x (slice length + fixed cost per CTA) + cost of the reduction pass
  //     (calibrated on B200: ~0.06 us per unit of K per wave, ~6 us fixed, partial sums at ~3 TB/s).
  int splits = 1, slice = K;
  const bool may_split = batch == 1 && N % 4 == 0;
  if (may_split && K > kSlice + kSlice / 2) { splits = (K + kSlice - 1) / kSlice; slice = kSlice; }
  if (may_split && kSched && K >= 256) {
    const long long slots = 148ll * (bn == 64 ? 2 : 1);
    double best = 1e30;
    int best_s = splits, best_slice = slice;
    const int s_max = K / 128 < 64 ? K / 128 : 64;
    for (int sc = 1; sc <= s_max; ++sc) {
      const int gran = (B_packed16 && tc::use_f16()) ? 64 : 32;  // whole k-blocks (the fp16 kernel's are 64 wide)
      int sl = ((K + sc - 1) / sc + gran - 1) / gran * gran;
      if (sl > kSlice + kSlice / 2) continue;  // rule (a)
      const int se = (K + sl - 1) / sl;
      const long long waves = (tiles * se + slots - 1) / slots;
      double cost = (double)waves * (0.06 * sl + 6.0);
      if (se > 1) cost += 4.0 + 2.0 * se * (double)M * N * 4.0 / 3.0e6;
      if (cost < best) { best = cost; best_s = se; best_slice = sl; }
    }
    splits = best_s; slice = best_slice;
  }
  // superpoint-sized products (M ~ 500) that still yield only a handful of CTAs: the FFMA kernel with 32x32 tiles
  // fills the machine better
  if (tiles * splits < 24) return 1;
  tc::Params p;
  p.A = A; p.B = B; p.C = C; p.bias = bias; p.row_div = row_div; p.residual = residual;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldr = ldr; p.sA = sA; p.sB = sB; p.sC = sC; p.sR = sR;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.act = act; p.k_split = 0;
  p.B_packed = (batch == 1) ? B_packed : nullptr;
  p.packed_kblocks = (K + 31) / 32;
  p.B_packed16 = (batch == 1) ? static_cast<const unsigned char*>(B_packed16) : nullptr;
  p.packed_kblocks64 = (K + 63) / 64;
  p.inv_scale16 = inv_scale16;
  p.gn_partial = nullptr; p.gn_groups = 0;
  // GroupNorm statistics ride in the epilogue when every group lies inside one column tile
  const bool gn_ok = gn && gn->partial && batch == 1 && gn->groups > 0 && N % gn->groups == 0 && N % 4 == 0 &&
                     bn % (N / gn->groups) == 0 && (N / gn->groups) <= bn;
  if (splits > 1) {
    float* partial = splitk_scratch(st, (size_t)splits * M * N * sizeof(float));
    if (partial == nullptr) return GR_ERR_CUDA;
    tc::Params q = p;
    q.C = partial; q.ldc = N; q.sC = (long long)M * N; q.bias = nullptr; q.row_div = nullptr; q.residual = nullptr;
    q.alpha = 1.f; q.act = 0; q.k_split = slice;
    int rc = tc::launch_any(bn, q, splits, st);
    if (rc == GR_OK) {
      const int nblk = (M + tc::kRedRows - 1) / tc::kRedRows;
      if (gn_ok && (size_t)nblk <= gn->capacity_blocks) {
        p.gn_partial = gn->partial; p.gn_groups = gn->groups;
        // column chunks of at most 256 (a multiple of the group width): small-M products still fill the machine
        const int cgw = N / gn->groups;
        int col_chunk = N;
        if (N > 256 && 256 % cgw == 0) col_chunk = 256;
        const int c4n = col_chunk / 4, lanes = c4n < 256 ? c4n : 256, R = 256 / lanes;
        const size_t smem = ((size_t)col_chunk + (R > 1 ? (size_t)R * col_chunk : 0)) * sizeof(double2);
        if (smem > 48 * 1024) {
          if (ensure_smem_attr(reinterpret_cast<const void*>(tc::splitk_reduce_gn_kernel), (int)smem) != cudaSuccess) return GR_ERR_CUDA;
        }
        dim3 rgrid(nblk, (N + col_chunk - 1) / col_chunk);
        if (launch_pdl(tc::splitk_reduce_gn_kernel, rgrid, dim3(256), smem, st, partial, splits, p, col_chunk) != cudaSuccess) rc = GR_ERR_CUDA;
        gn->nblk = nblk;
      } else {
        const long long total = (long long)M * (N / 4);
        if (launch_pdl(tc::splitk_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, partial, splits, p) != cudaSuccess) rc = GR_ERR_CUDA;
      }
      count_launch();
      if (cudaGetLastError() != cudaSuccess) rc = GR_ERR_CUDA;
    }
    return rc;
  }
  if (gn_ok && (size_t)((M + 127) / 128) <= gn->capacity_blocks) {
    p.gn_partial = gn->partial; p.gn_groups = gn->groups;
    gn->nblk = (M + 127) / 128;
  }
  return tc::launch_any(bn, p, batch, st);
}

}  // namespace gr

/* fp16-split image of a static (N, K) weight for the kind::f16 tensor-core kernels: `scale` must be a power of two with
 * scale * max|W| far below 65504 (callers pass 1 / scale to the products).  out: gr_packed_weight_f16_bytes(N, K) bytes. */
extern "C" size_t gr_packed_weight_f16_bytes(int N, int K) {
  if (N <= 0 || K <= 0) return 0;
  return (size_t)((N + 127) / 128) * ((K + 63) / 64) * (2 * 16384);
}
extern "C" int gr_pack_weight_f16x3(const float* W, int N, int K, float scale, void* out, void* stream) {
  if (N <= 0 || K <= 0 || !W || !out || !(scale > 0.f)) return GR_ERR_BAD_ARG;
  dim3 grid((K + 63) / 64, (N + 127) / 128);
  GR_CHECK_CUDA(gr::launch_pdl(gr::tc::pack_weight_f16x3_kernel, grid, dim3(256), (size_t)0, static_cast<cudaStream_t>(stream), W, N, K, scale,
                           static_cast<unsigned char*>(out)));
  GR_CHECK_LAUNCH("pack_weight_f16x3_kernel");
  return GR_OK;
}
/* device address of the flag the fp16 kernels raise when an activation exceeded fp16's range (|x| > 6e4) */
extern "C" int gr_gemm_f16_overflow_ptr(int** dev_ptr) {
  if (!dev_ptr) return GR_ERR_BAD_ARG;
  void* q = nullptr;
  if (cudaGetSymbolAddress(&q, gr::tc::g_f16_overflow) != cudaSuccess) return GR_ERR_CUDA;
  *dev_ptr = static_cast<int*>(q);
  return GR_OK;
}
