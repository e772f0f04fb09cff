// M1: superpoint matching (geotransformer/modules/geotransformer/superpoint_matching.py:13-50) and
// S1: log-domain Sinkhorn with learnable dustbin (modules/sinkhorn/learnable_sinkhorn.py:5-66).
#include <stdlib.h>

#include "common.cuh"

namespace gr {

// ---------------------------------------------------------------------------------------------
// M1
// ---------------------------------------------------------------------------------------------

// S[i,j] = mask ? exp(-max(2 - 2 xy, 0)) : 0 (in place on xy); rowsum[i] in fixed order (warp per row)
__global__ void __launch_bounds__(256) match_exp_rows_kernel(float* __restrict__ S, int Nr, int Ns,
                                                             const unsigned char* __restrict__ rmask,
                                                             const unsigned char* __restrict__ smask, float* __restrict__ rowsum) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= Nr) return;
  const int lane = threadIdx.x & 31;
  const bool rv = rmask[i] != 0;
  float s = 0.f;
  for (int j = lane; j < Ns; j += 32) {
    float v = 0.f;
    if (rv && smask[j]) v = expf(-fmaxf(2.0f - 2.0f * S[(long long)i * Ns + j], 0.0f));
    S[(long long)i * Ns + j] = v;
    s += v;
  }
  s = warp_sum(s);
  if (lane == 0) rowsum[i] = s;
}

__global__ void __launch_bounds__(256) match_colsum_kernel(const float* __restrict__ S, int Nr, int Ns, float* __restrict__ colsum) {
  pdl_wait();
  pdl_trigger();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Ns) return;
  float s = 0.f;
  for (int i = 0; i < Nr; ++i) s += S[(long long)i * Ns + j];
  colsum[j] = s;
}

// key = (score bits << 32) | ~flat  -> descending key order == descending score, ascending flat index
__global__ void __launch_bounds__(256) match_keys_kernel(const float* __restrict__ S, int Nr, int Ns,
                                                         const float* __restrict__ rowsum, const float* __restrict__ colsum,
                                                         const unsigned char* __restrict__ rmask, const unsigned char* __restrict__ smask,
                                                         int dual, unsigned long long* __restrict__ keys, int* __restrict__ n_valid) {
  pdl_wait();
  pdl_trigger();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {
    int a = 0, b = 0;
    for (int i = 0; i < Nr; ++i) a += rmask[i] != 0;
    for (int j = 0; j < Ns; ++j) b += smask[j] != 0;
    *n_valid = a * b > 0 ? (int)min((long long)a * b, (long long)0x7fffffff) : 0;
  }
  if (t >= (long long)Nr * Ns) return;
  const int i = (int)(t / Ns), j = (int)(t % Ns);
  unsigned long long key = 0ull;
  if (rmask[i] && smask[j]) {
    const float v = S[t];
    const float sc = dual ? (v / rowsum[i]) * (v / colsum[j]) : v;
    key = ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xffffffffu - (unsigned int)t);
    if (key == 0ull) key = 1ull;
  }
  keys[t] = key;
}

constexpr int kTopChunk = 2048;

// each CTA sorts a chunk of 2048 keys (descending) in shared memory and keeps the first `keep`
__global__ void __launch_bounds__(1024) topk_chunk_kernel(const unsigned long long* __restrict__ in, long long n, int keep,
                                                          unsigned long long* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned long long sh[kTopChunk];
  const long long base = (long long)blockIdx.x * kTopChunk;
  for (int i = threadIdx.x; i < kTopChunk; i += blockDim.x) sh[i] = (base + i < n) ? in[base + i] : 0ull;
  __syncthreads();
  for (int size = 2; size <= kTopChunk; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < kTopChunk / 2; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = sh[lo], b = sh[hi];
        if ((a < b) == desc) { sh[lo] = b; sh[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < keep; i += blockDim.x) out[(long long)blockIdx.x * keep + i] = sh[i];
}

__global__ void match_decode_kernel(const unsigned long long* __restrict__ keys, int k, int Ns, const int* __restrict__ n_valid,
                                    long long* __restrict__ ref_idx, long long* __restrict__ src_idx, float* __restrict__ scores,
                                    int* __restrict__ count) {
  pdl_wait();
  pdl_trigger();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = min(k, *n_valid);
  if (t == 0) *count = c;
  if (t >= k) return;
  if (t < c) {
    const unsigned long long key = keys[t];
    const unsigned int flat = 0xffffffffu - (unsigned int)(key & 0xffffffffull);
    ref_idx[t] = flat / Ns;
    src_idx[t] = flat % Ns;
    scores[t] = __uint_as_float((unsigned int)(key >> 32));
  } else {
    ref_idx[t] = 0; src_idx[t] = 0; scores[t] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// S1
// ---------------------------------------------------------------------------------------------

// one CTA per patch pair; the (K+1)x(K+1) padded score matrix stays in shared memory for all iterations.
// 16 warps; every warp owns rows (columns) w, w+16, ... and reduces four of them at a time so that the
// shuffle / exp latencies of independent rows overlap.
constexpr int kSinkThreads = 512;
constexpr int kSinkIlp = 4;

// out[i] = bias[i] - logsumexp_j(ps[i*si + j*sj] + add[j]) for the lines owned by this warp.
// The log-sum-exp shift is the line's PREVIOUS log-sum-exp (bias[i] - out[i]) instead of a freshly computed
// maximum: the iterates move slowly and all scores are O(10), so exp(x - shift) can neither overflow nor flush a
// whole line to zero, and the max pass (half of the work) disappears.  The first iteration uses the true maximum.
// Fully masked lines (bias = -inf marker): every term equals -inf in fp32, the reference's result is exactly 0.
template <bool FIRST>
__device__ __forceinline__ void sinkhorn_lse_pass(const float* __restrict__ ps, int K1, int si, int sj,
                                                  const float* __restrict__ add, const float* __restrict__ bias,
                                                  float* __restrict__ out, float masked_below, int warp, int nwarp, int lane) {
  for (int i0 = warp; i0 < K1; i0 += nwarp * kSinkIlp) {
    float shift[kSinkIlp], sm[kSinkIlp];
    bool live[kSinkIlp];
#pragma unroll
    for (int r = 0; r < kSinkIlp; ++r) {
      const int i = i0 + r * nwarp;
      live[r] = i < K1 && bias[i] > masked_below;
      shift[r] = 0.f;
      if (live[r]) {
        if (FIRST) {
          float m = -INFINITY;
          for (int j = lane; j < K1; j += 32) m = fmaxf(m, ps[i * si + j * sj] + add[j]);
          shift[r] = m;
        } else {
          shift[r] = bias[i] - out[i];
        }
      }
    }
    if (FIRST) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < kSinkIlp; ++r) shift[r] = fmaxf(shift[r], __shfl_xor_sync(0xffffffffu, shift[r], o));
    }
#pragma unroll
    for (int r = 0; r < kSinkIlp; ++r) {
      const int i = i0 + r * nwarp;
      float s = 0.f;
      if (live[r])
        for (int j = lane; j < K1; j += 32) s += __expf(ps[i * si + j * sj] + add[j] - shift[r]);  // ex2.approx: |rel err| ~1e-6
      sm[r] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < kSinkIlp; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], o);
    if (!FIRST) {
#pragma unroll
      for (int r = 0; r < kSinkIlp; ++r) {
        if (live[r] && !(sm[r] > 1e-30f && sm[r] < 1e30f)) {  // see sinkhorn_lse_pass128: redo with the true maximum
          const int i = i0 + r * nwarp;
          float m = -INFINITY;
          for (int j = lane; j < K1; j += 32) m = fmaxf(m, ps[i * si + j * sj] + add[j]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          float s2 = 0.f;
          for (int j = lane; j < K1; j += 32) s2 += __expf(ps[i * si + j * sj] + add[j] - m);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          sm[r] = s2; shift[r] = m;
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < kSinkIlp; ++r) {
        const int i = i0 + r * nwarp;
        if (i < K1) out[i] = live[r] ? bias[i] - (logf(sm[r]) + shift[r]) : 0.f;
      }
    }
  }
}

// K = 128 form of the pass: 17 warps.  Warps 0-15 own the 128 regular lines (8 each, all eight reduced together so
// that 32 independent exp chains per lane are in flight), warp 16 owns the dustbin line; the dustbin COLUMN term
// (j = 128) is added by lane 0 after its four strided terms, which is exactly the order of the generic loop
// (j = lane, lane+32, ...), so both forms produce identical bits.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// EXP2: everything (scores, potentials, marginals) is kept in units of log2, so a term costs add, add, ex2 instead of
// add, add, mul, ex2 (__expf multiplies by log2 e first); converted back once at the end.
template <bool FIRST, bool EXP2>
__device__ __forceinline__ void sinkhorn_lse_pass128(const float* __restrict__ ps, int si, int sj, const float* __restrict__ add,
                                                     const float* __restrict__ bias, float* __restrict__ out,
                                                     float masked_below, int warp, int lane) {
  constexpr int K = 128;
  if (warp > 16) return;
  const int nl = warp < 16 ? 8 : 1;          // lines of this warp
  const int ibase = warp < 16 ? warp : K;    // line r -> ibase + 16 r
  float a[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) a[t] = add[lane + 32 * t];
  const float ad = add[K];
  float shift[8], sm[8];
  bool live[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = ibase + 16 * r;
    live[r] = r < nl && bias[i] > masked_below;
    shift[r] = 0.f;
    if (live[r]) {
      if (FIRST) {
        float m = ps[i * si + K * sj] + ad;
#pragma unroll
        for (int t = 0; t < 4; ++t) m = fmaxf(m, ps[i * si + (lane + 32 * t) * sj] + a[t]);
        shift[r] = m;
      } else {
        shift[r] = bias[i] - out[i];
      }
    }
  }
  if (FIRST) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < 8; ++r) shift[r] = fmaxf(shift[r], __shfl_xor_sync(0xffffffffu, shift[r], o));
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = ibase + 16 * r;
    float s = 0.f;
    if (live[r]) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float arg = ps[i * si + (lane + 32 * t) * sj] + a[t] - shift[r];
        s += EXP2 ? ex2_approx(arg) : __expf(arg);
      }
      if (lane == 0) {
        const float arg = ps[i * si + K * sj] + ad - shift[r];
        s += EXP2 ? ex2_approx(arg) : __expf(arg);
      }
    }
    sm[r] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], o);
  // lane r finishes line r (every lane holds all eight totals)
  float my_sum = sm[0], my_shift = shift[0];
  bool my_live = live[0];
#pragma unroll
  for (int r = 1; r < 8; ++r)
    if (lane == r) { my_sum = sm[r]; my_shift = shift[r]; my_live = live[r]; }
  if (!FIRST) {
    // The shift is the previous iterate's log-sum-exp.  Should every term of a line flush to zero (or the sum leave the
    // fp32 range) -- possible for trained weights with |scores| >> 10 -- that line is redone with its true maximum by the
    // whole warp.  One ballot per pass; it never fires on O(10) scores.
    unsigned bad = __ballot_sync(0xffffffffu, lane < nl && my_live && !(my_sum > 1e-30f && my_sum < 1e30f));
    while (bad) {
      const int r = __ffs(bad) - 1;
      bad &= bad - 1;
      const int i = ibase + 16 * r;
      float m = lane == 0 ? ps[i * si + K * sj] + ad : -INFINITY;
#pragma unroll
      for (int t = 0; t < 4; ++t) m = fmaxf(m, ps[i * si + (lane + 32 * t) * sj] + a[t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s2 = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float arg = ps[i * si + (lane + 32 * t) * sj] + a[t] - m;
        s2 += EXP2 ? ex2_approx(arg) : __expf(arg);
      }
      if (lane == 0) {
        const float arg = ps[i * si + K * sj] + ad - m;
        s2 += EXP2 ? ex2_approx(arg) : __expf(arg);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      if (lane == r) { my_sum = s2; my_shift = m; }
    }
  }
  if (lane < nl) {
    const int i = ibase + 16 * lane;
    out[i] = my_live ? bias[i] - ((EXP2 ? log2f(my_sum) : logf(my_sum)) + my_shift) : 0.f;
  }
}

// Steady-state form of the K = 128 pass (every iteration but the first).  Same arithmetic per term as
// sinkhorn_lse_pass128<false>, restructured for throughput: no branch inside the term loops (masked lines are computed and
// discarded; a branch per line serialised the eight exp chains), compile-time strides so that every shared-memory
// address is base + immediate, the dustbin-column terms of the warp's eight lines evaluated by eight lanes at once, and
// the eight line totals reduced with a transposing butterfly (9 shuffles instead of 40).  The summation tree differs
// from the generic pass in the last bits; it is fixed, so results are reproducible run to run.
template <bool EXP2, bool ROWPASS>
__device__ __forceinline__ void sinkhorn_lse_steady128(const float* __restrict__ ps, const float* __restrict__ add,
                                                       const float* __restrict__ bias, float* __restrict__ out,
                                                       float masked_below, int warp, int lane) {
  constexpr int K = 128, K1 = K + 1;
  constexpr int si = ROWPASS ? K1 : 1, sj = ROWPASS ? 1 : K1;
  if (warp > 16) return;
  const bool reg = warp < 16;              // warps 0-15: lines warp + 16 r; warp 16: the dustbin line (eight copies of it)
  const int nl = reg ? 8 : 1;
  const int ibase = reg ? warp : K;
  const int ls = reg ? 16 * si : 0;        // stride between the warp's lines
  const float* base = ps + ibase * si + lane * sj;
  float a[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) a[t] = add[lane + 32 * t];
  const float ad = add[K];
  float shift[8], sm[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = ibase + (reg ? 16 * r : 0);
    shift[r] = bias[i] - out[i];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float arg = base[r * ls + 32 * t * sj] + a[t] - shift[r];
      s += EXP2 ? ex2_approx(arg) : __expf(arg);
    }
    sm[r] = s;
  }
  // line L = (lane >> 2) & 7 ends up on lanes 4L .. 4L+3
  const int L = (lane >> 2) & 7;
  const int iL = ibase + (reg ? 16 * L : 0);
  const float my_bias = bias[iL];
  const float my_shift = my_bias - out[iL];
  const bool my_live = L < nl && my_bias > masked_below;
  const float dust_arg = ps[iL * si + K * sj] + ad - my_shift;
  const float dust = EXP2 ? ex2_approx(dust_arg) : __expf(dust_arg);
  float v4[4], v2[2], my_sum;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float send = hi ? sm[k] : sm[k + 4], keep = hi ? sm[k + 4] : sm[k];
      v4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float send = hi ? v4[k] : v4[k + 2], keep = hi ? v4[k + 2] : v4[k];
      v2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
    const float send = hi ? v2[0] : v2[1], keep = hi ? v2[1] : v2[0];
    my_sum = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  my_sum += __shfl_xor_sync(0xffffffffu, my_sum, 2);
  my_sum += __shfl_xor_sync(0xffffffffu, my_sum, 1);
  my_sum += dust;
  float fin_shift = my_shift;
  // Underflow / overflow repair, as in sinkhorn_lse_pass128: redo the line with its true maximum (whole warp).
  unsigned bad = __ballot_sync(0xffffffffu, (lane & 3) == 0 && my_live && !(my_sum > 1e-30f && my_sum < 1e30f));
  while (bad) {
    const int src = __ffs(bad) - 1;
    bad &= bad - 1;
    const int r = src >> 2;
    const int i = ibase + (reg ? 16 * r : 0);
    float m = lane == 0 ? ps[i * si + K * sj] + ad : -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) m = fmaxf(m, ps[i * si + (lane + 32 * t) * sj] + a[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s2 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float arg = ps[i * si + (lane + 32 * t) * sj] + a[t] - m;
      s2 += EXP2 ? ex2_approx(arg) : __expf(arg);
    }
    if (lane == 0) {
      const float arg = ps[i * si + K * sj] + ad - m;
      s2 += EXP2 ? ex2_approx(arg) : __expf(arg);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == src) { my_sum = s2; fin_shift = m; }
  }
  if ((lane & 3) == 0 && L < nl)
    out[iL] = my_live ? my_bias - ((EXP2 ? log2f(my_sum) : logf(my_sum)) + fin_shift) : 0.f;
}

constexpr int kSink128Threads = 17 * 32;

template <bool FAST128, bool EXP2 = false>
__global__ void __launch_bounds__(FAST128 ? kSink128Threads : kSinkThreads, FAST128 ? 2 : 1) sinkhorn_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ row_masks,
                                                       const unsigned char* __restrict__ col_masks, const float* __restrict__ alpha_p,
                                                       int K, int iters, float inf, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int K1 = K + 1;
  float* ps = sm;               // K1*K1
  float* u = ps + K1 * K1;      // K1
  float* v = u + K1;            // K1
  float* lmu = v + K1;          // K1
  float* lnu = lmu + K1;        // K1
  __shared__ int s_cnt[2];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const unsigned char* rm = row_masks + (long long)b * K;
  const unsigned char* cm = col_masks + (long long)b * K;
  const float alpha = *alpha_p;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  {
    int a = 0, c = 0;
    for (int i = threadIdx.x; i < K; i += blockDim.x) { a += rm[i] != 0; c += cm[i] != 0; }
    a = warp_sum(a); c = warp_sum(c);
    if (lane == 0) { atomicAdd(&s_cnt[0], a); atomicAdd(&s_cnt[1], c); }
  }
  for (int t = threadIdx.x; t < K1 * K1; t += blockDim.x) {
    const int i = t / K1, j = t % K1;
    const bool masked = (i < K && !rm[i]) || (j < K && !cm[j]);
    float val = (i < K && j < K) ? scores[((long long)b * K + i) * K + j] : alpha;
    val = masked ? -inf : val;
    ps[t] = EXP2 ? val * 1.4426950408889634f : val;
  }
  __syncthreads();
  const float nvr = (float)s_cnt[0], nvc = (float)s_cnt[1];
  const float norm = -logf(nvr + nvc);
  for (int i = threadIdx.x; i < K1; i += blockDim.x) {
    float mu = i < K ? norm : logf(nvc) + norm;
    float nu = i < K ? norm : logf(nvr) + norm;
    if (i < K && !rm[i]) mu = -inf;
    if (i < K && !cm[i]) nu = -inf;
    if (EXP2) { mu *= 1.4426950408889634f; nu *= 1.4426950408889634f; }
    lmu[i] = mu; lnu[i] = nu; u[i] = 0.f; v[i] = 0.f;
  }
  __syncthreads();
  const float masked_below = -0.5f * inf;  // (a masked marginal scaled by log2 e is still far below this)
  for (int it = 0; it < iters; ++it) {
    // u_i = log_mu_i - logsumexp_j(ps_ij + v_j) ; v_j = log_nu_j - logsumexp_i(ps_ij + u_i)
    if (FAST128) {
      if (it == 0) sinkhorn_lse_pass128<true, EXP2>(ps, K1, 1, v, lmu, u, masked_below, warp, lane);
      else sinkhorn_lse_steady128<EXP2, true>(ps, v, lmu, u, masked_below, warp, lane);
      __syncthreads();
      if (it == 0) sinkhorn_lse_pass128<true, EXP2>(ps, 1, K1, u, lnu, v, masked_below, warp, lane);
      else sinkhorn_lse_steady128<EXP2, false>(ps, u, lnu, v, masked_below, warp, lane);
      __syncthreads();
      continue;
    }
    if (it == 0) sinkhorn_lse_pass<true>(ps, K1, K1, 1, v, lmu, u, masked_below, warp, nwarp, lane);
    else sinkhorn_lse_pass<false>(ps, K1, K1, 1, v, lmu, u, masked_below, warp, nwarp, lane);
    __syncthreads();
    if (it == 0) sinkhorn_lse_pass<true>(ps, K1, 1, K1, u, lnu, v, masked_below, warp, nwarp, lane);
    else sinkhorn_lse_pass<false>(ps, K1, 1, K1, u, lnu, v, masked_below, warp, nwarp, lane);
    __syncthreads();
  }
  float* o = out + (long long)b * K1 * K1;
  for (int t = threadIdx.x; t < K1 * K1; t += blockDim.x) {
    const int i = t / K1, j = t % K1;
    if (EXP2) {
      // back to natural-log units; the score itself is re-read unscaled so that it carries no extra rounding
      const bool masked = (i < K && !rm[i]) || (j < K && !cm[j]);
      float val = (i < K && j < K) ? scores[((long long)b * K + i) * K + j] : alpha;
      val = masked ? -inf : val;
      o[t] = (val + u[i] * 0.6931471805599453f + v[j] * 0.6931471805599453f) - norm;
    } else {
      o[t] = (ps[t] + u[i] + v[j]) - norm;
    }
  }
}

// ---- S1, K = 128: scaling form ---------------------------------------------------------------------------------
// The log-domain iteration u_i = log mu_i - LSE_j(ps_ij + v_j), v_j = log nu_j - LSE_i(ps_ij + u_i) spends one exp per
// matrix entry per half step: 2 * 100 * 129^2 exps per patch, SFU-bound (0.47 ms for 256 patches even with both
// passes branch-free).  With the potentials split as u = U + log2 a, v = V + log2 b and Kt_ij = 2^(ps_ij + U_i + V_j)
// held fixed for a block of iterations, the same recurrence is
//     a_i = mu_i / sum_j Kt_ij b_j ,   b_j = nu_j / sum_i Kt_ij a_i          (one FMA per entry per half step)
// and every kScaleBlock iterations a, b are absorbed into U, V and Kt is rebuilt from ps (so rounding does not
// accumulate in Kt and its entries stay <= ~1).  The first iteration runs in the log domain with the true row/column
// maxima, which is what bounds Kt.  Kt lives in registers: warp w < 16 owns rows w + 16 r (r < 8), lane l the columns
// l + 32 t (t < 4); warp 16 owns the dustbin row; the dustbin column entry of row L sits on lane 4 L.  Row sums are
// reduced inside the warp (transposing butterfly), column sums through 17 per-warp partial vectors in shared memory,
// added in a fixed order.  Should any sum leave [1e-30, 1e30] (scores far outside the trained range), the CTA starts
// over with the plain log-domain kernel body, so the result is always the robust one.
constexpr int kScaleBlock = 11;

__global__ void __launch_bounds__(kSink128Threads, 2) sinkhorn_scaling128_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ row_masks,
                                                                                 const unsigned char* __restrict__ col_masks, const float* __restrict__ alpha_p,
                                                                                 int iters, float inf, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  constexpr int K = 128, K1 = K + 1;
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  extern __shared__ float sm[];
  float* ps = sm;               // K1*K1, log2 units
  float* u = ps + K1 * K1;      // potentials, log2 units
  float* v = u + K1;
  float* lmu = v + K1;          // log2 marginals
  float* lnu = lmu + K1;
  float* mu = lnu + K1;         // marginals
  float* nu = mu + K1;
  float* a_s = nu + K1;         // scalings of the current block
  float* b_s = a_s + K1;
  float* part = b_s + K1;       // 17 x K1 column partials
  __shared__ int s_cnt[2];
  __shared__ int s_trouble;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned char* rm = row_masks + (long long)b * K;
  const unsigned char* cm = col_masks + (long long)b * K;
  const float alpha = *alpha_p;
  if (tid < 2) s_cnt[tid] = 0;
  if (tid == 2) s_trouble = 0;
  __syncthreads();
  {
    int a = 0, c = 0;
    for (int i = tid; i < K; i += blockDim.x) { a += rm[i] != 0; c += cm[i] != 0; }
    a = warp_sum(a); c = warp_sum(c);
    if (lane == 0) { atomicAdd(&s_cnt[0], a); atomicAdd(&s_cnt[1], c); }
  }
  for (int t = tid; t < K1 * K1; t += blockDim.x) {
    const int i = t / K1, j = t % K1;
    const bool masked = (i < K && !rm[i]) || (j < K && !cm[j]);
    float val = (i < K && j < K) ? scores[((long long)b * K + i) * K + j] : alpha;
    val = masked ? -inf : val;
    ps[t] = val * kLog2e;
  }
  __syncthreads();
  const float nvr = (float)s_cnt[0], nvc = (float)s_cnt[1];
  const float norm = -logf(nvr + nvc);
  const float masked_below = -0.5f * inf;
  for (int i = tid; i < K1; i += blockDim.x) {
    float m = i < K ? norm : logf(nvc) + norm;
    float n = i < K ? norm : logf(nvr) + norm;
    if (i < K && !rm[i]) m = -inf;
    if (i < K && !cm[i]) n = -inf;
    m *= kLog2e; n *= kLog2e;
    lmu[i] = m; lnu[i] = n; u[i] = 0.f; v[i] = 0.f;
    mu[i] = m > masked_below ? ex2_approx(m) : 0.f;
    nu[i] = n > masked_below ? ex2_approx(n) : 0.f;
  }
  __syncthreads();
  int done = 0;
  if (iters > 0) {
    sinkhorn_lse_pass128<true, true>(ps, K1, 1, v, lmu, u, masked_below, warp, lane);
    __syncthreads();
    sinkhorn_lse_pass128<true, true>(ps, 1, K1, u, lnu, v, masked_below, warp, lane);
    __syncthreads();
    done = 1;
  }
  // roles
  const bool reg = warp < 16;
  const int nl = reg ? 8 : 1;
  const int ibase = reg ? warp : K;
  const int L = (lane >> 2) & 7;
  const int iL = ibase + (reg ? 16 * L : 0);
  const bool fin = (lane & 3) == 0 && L < nl;        // this lane finishes row iL
  const bool live_row = fin && lmu[iL] > masked_below;
  const float my_mu = fin ? mu[iL] : 0.f;
  const bool col_owner = tid < K1;
  const bool live_col = col_owner && lnu[col_owner ? tid : 0] > masked_below;
  const float my_nu = col_owner ? nu[tid] : 0.f;
  bool trouble = false;
  while (done < iters) {
    const int m = iters - done < kScaleBlock ? iters - done : kScaleBlock;
    float k[8][4];
    {
      float vj[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) vj[t] = v[lane + 32 * t];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = ibase + (reg ? 16 * r : 0);
        const float ui = u[i];
#pragma unroll
        for (int t = 0; t < 4; ++t) k[r][t] = r < nl ? ex2_approx(ps[i * K1 + lane + 32 * t] + ui + vj[t]) : 0.f;
      }
    }
    const float kd = fin ? ex2_approx(ps[iL * K1 + K] + u[iL] + v[K]) : 0.f;
    if (col_owner) b_s[tid] = 1.f;
    __syncthreads();
    float my_a = 0.f, my_b = 0.f;
    for (int it = 0; it < m; ++it) {
      // ---- a_i = mu_i / sum_j Kt_ij b_j
      float sr[8];
      {
        float bj[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) bj[t] = b_s[lane + 32 * t];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float acc = k[r][0] * bj[0];
#pragma unroll
          for (int t = 1; t < 4; ++t) acc = fmaf(k[r][t], bj[t], acc);
          sr[r] = acc;
        }
      }
      float v4[4], v2[2], my_s;
      {
        const bool hi = lane & 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float send = hi ? sr[q] : sr[q + 4], keep = hi ? sr[q + 4] : sr[q];
          v4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool hi = lane & 8;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float send = hi ? v4[q] : v4[q + 2], keep = hi ? v4[q + 2] : v4[q];
          v2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool hi = lane & 4;
        const float send = hi ? v2[0] : v2[1], keep = hi ? v2[1] : v2[0];
        my_s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      my_s += __shfl_xor_sync(0xffffffffu, my_s, 2);
      my_s += __shfl_xor_sync(0xffffffffu, my_s, 1);
      my_s = fmaf(kd, b_s[K], my_s);
      if (fin) {
        my_a = live_row ? my_mu / my_s : 0.f;
        if (live_row && !(my_s > 1e-30f && my_s < 1e30f)) s_trouble = 1;
        a_s[iL] = my_a;
      }
      __syncthreads();
      // ---- b_j = nu_j / sum_i Kt_ij a_i
      {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float ar = a_s[ibase + (reg ? 16 * r : 0)];
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[t] = fmaf(k[r][t], ar, acc[t]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) part[warp * K1 + lane + 32 * t] = acc[t];
        float d = kd * my_a;  // zero off the finishing lanes
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 8);
        d += __shfl_xor_sync(0xffffffffu, d, 16);
        if (lane == 0) part[warp * K1 + K] = d;
      }
      __syncthreads();
      if (col_owner) {
        float tj = part[tid];
#pragma unroll
        for (int w = 1; w < 17; ++w) tj += part[w * K1 + tid];
        my_b = live_col ? my_nu / tj : 0.f;
        if (live_col && !(tj > 1e-30f && tj < 1e30f)) s_trouble = 1;
        b_s[tid] = my_b;
      }
      __syncthreads();
    }
    if (live_row) u[iL] += log2f(my_a);
    if (live_col) v[tid] += log2f(my_b);
    done += m;
    __syncthreads();
    trouble = s_trouble != 0;
    if (trouble) break;
  }
  if (trouble) {  // out-of-range sums: redo the patch with the log-domain recurrence
    for (int i = tid; i < K1; i += blockDim.x) { u[i] = 0.f; v[i] = 0.f; }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
      if (it == 0) sinkhorn_lse_pass128<true, true>(ps, K1, 1, v, lmu, u, masked_below, warp, lane);
      else sinkhorn_lse_steady128<true, true>(ps, v, lmu, u, masked_below, warp, lane);
      __syncthreads();
      if (it == 0) sinkhorn_lse_pass128<true, true>(ps, 1, K1, u, lnu, v, masked_below, warp, lane);
      else sinkhorn_lse_steady128<true, false>(ps, u, lnu, v, masked_below, warp, lane);
      __syncthreads();
    }
  }
  float* o = out + (long long)b * K1 * K1;
  for (int t = tid; t < K1 * K1; t += blockDim.x) {
    const int i = t / K1, j = t % K1;
    const bool masked = (i < K && !rm[i]) || (j < K && !cm[j]);
    float val = (i < K && j < K) ? scores[((long long)b * K + i) * K + j] : alpha;
    val = masked ? -inf : val;
    o[t] = (val + u[i] * kLn2 + v[j] * kLn2) - norm;
  }
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_superpoint_matching_workspace_size(int Nr, int Ns, int k) {
  Carver c(nullptr, 0);
  c.take<float>(Nr);
  c.take<float>(Ns);
  const long long n = (long long)Nr * Ns;
  c.take<unsigned long long>(n);
  long long m = n;
  while (m > kTopChunk) { m = ((m + kTopChunk - 1) / kTopChunk) * k; c.take<unsigned long long>(m); }
  c.take<unsigned long long>(k);
  c.take<int>(4);
  return c.off;
}

/* M1.  xy: (Nr,Ns) f32 = ref_feats . src_feats^T, overwritten.  Outputs k entries (padded with 0 beyond *count). */
extern "C" int gr_superpoint_matching(float* xy, int Nr, int Ns, const uint8_t* ref_masks, const uint8_t* src_masks, int k,
                                      int dual_normalization, int64_t* ref_idx, int64_t* src_idx, float* scores, int32_t* count,
                                      void* ws, size_t ws_bytes, void* stream) {
  if (Nr <= 0 || Ns <= 0 || k <= 0 || k > 1024 || (long long)Nr * Ns >= (1ll << 32)) return GR_ERR_BAD_ARG;
  if (!xy || !ref_masks || !src_masks || !ref_idx || !src_idx || !scores || !count) return GR_ERR_BAD_ARG;
  Carver c(ws, ws_bytes);
  float* rowsum = c.take<float>(Nr);
  float* colsum = c.take<float>(Ns);
  const long long n = (long long)Nr * Ns;
  unsigned long long* keys = c.take<unsigned long long>(n);
  unsigned long long* level[8];
  long long level_n[8];
  int nl = 0;
  long long m = n;
  while (m > kTopChunk) { m = ((m + kTopChunk - 1) / kTopChunk) * k; level[nl] = c.take<unsigned long long>(m); level_n[nl] = m; ++nl; }
  unsigned long long* fin = c.take<unsigned long long>(k);
  int* n_valid = c.take<int>(4);
  if (!ws || !c.ok) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GR_CHECK_CUDA(launch_pdl(match_exp_rows_kernel, dim3(ceil_div(Nr, 8)), dim3(256), (size_t)(0), st, xy, Nr, Ns, ref_masks, src_masks, rowsum));
  GR_CHECK_LAUNCH("match_exp_rows_kernel");
  GR_CHECK_CUDA(launch_pdl(match_colsum_kernel, dim3(ceil_div(Ns, 256)), dim3(256), (size_t)(0), st, xy, Nr, Ns, colsum));
  GR_CHECK_LAUNCH("match_colsum_kernel");
  GR_CHECK_CUDA(launch_pdl(match_keys_kernel, dim3(ceil_div(n, 256)), dim3(256), (size_t)(0), st, xy, Nr, Ns, rowsum, colsum, ref_masks, src_masks, dual_normalization, keys, n_valid));
  GR_CHECK_LAUNCH("match_keys_kernel");
  const unsigned long long* cur = keys;
  long long cur_n = n;
  for (int l = 0; l < nl; ++l) {
    GR_CHECK_CUDA(launch_pdl(topk_chunk_kernel, dim3(ceil_div(cur_n, kTopChunk)), dim3(1024), (size_t)(0), st, cur, cur_n, k, level[l]));
    GR_CHECK_LAUNCH("topk_chunk_kernel");
    cur = level[l]; cur_n = level_n[l];
  }
  GR_CHECK_CUDA(launch_pdl(topk_chunk_kernel, dim3(1), dim3(1024), (size_t)(0), st, cur, cur_n, k, fin));
  GR_CHECK_LAUNCH("topk_chunk_kernel");
  GR_CHECK_CUDA(launch_pdl(match_decode_kernel, dim3(ceil_div(k, 256)), dim3(256), (size_t)(0), st, fin, k, Ns, n_valid, reinterpret_cast<long long*>(ref_idx),
                                                        reinterpret_cast<long long*>(src_idx), scores, count));
  GR_CHECK_LAUNCH("match_decode_kernel");
  return GR_OK;
}

/* S1.  scores (P,K,K), masks (P,K) u8 (1 = valid), alpha device scalar -> out (P,K+1,K+1). */
extern "C" int gr_sinkhorn(const float* scores, const uint8_t* row_masks, const uint8_t* col_masks, const float* alpha, int P,
                           int K, int num_iterations, float inf, float* out, void* stream) {
  if (P < 0 || K <= 0 || num_iterations < 0) return GR_ERR_BAD_ARG;
  if (P == 0) return GR_OK;
  if (!scores || !row_masks || !col_masks || !alpha || !out) return GR_ERR_BAD_ARG;
  const size_t smem = ((size_t)(K + 1) * (K + 1) + 4 * (size_t)(K + 1)) * sizeof(float);
  if (smem > 220 * 1024) return GR_ERR_CAPACITY;
  static int fast_knob = -1;
  if (fast_knob < 0) { const char* e = getenv("GAUSSREG_SINKHORN128"); fast_knob = e ? atoi(e) : 1; }
  static int exp2_knob = -1;
  if (exp2_knob < 0) { const char* e = getenv("GAUSSREG_SINKHORN_EXP2"); exp2_knob = e ? atoi(e) : 1; }  // measured: 0.63 -> 0.52 ms
  static int scaling_knob = -1;
  if (scaling_knob < 0) { const char* e = getenv("GAUSSREG_SINKHORN_SCALING"); scaling_knob = e ? atoi(e) : 1; }
  if (K == 128 && fast_knob && exp2_knob && scaling_knob) {
    const size_t smem_sc = ((size_t)(K + 1) * (K + 1) + (8 + 17) * (size_t)(K + 1)) * sizeof(float);
    GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(sinkhorn_scaling128_kernel), (int)smem_sc));
    GR_CHECK_CUDA(launch_pdl(sinkhorn_scaling128_kernel, dim3(P), dim3(kSink128Threads), smem_sc, static_cast<cudaStream_t>(stream), scores, row_masks, col_masks, alpha,
                             num_iterations, inf, out));
  } else if (K == 128 && fast_knob && exp2_knob) {
    GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(sinkhorn_kernel<true, true>), (int)smem));
    GR_CHECK_CUDA(launch_pdl(sinkhorn_kernel<true, true>, dim3(P), dim3(kSink128Threads), (size_t)(smem), static_cast<cudaStream_t>(stream), scores, row_masks, col_masks, alpha,
                                                                                                K, num_iterations, inf, out));
  } else if (K == 128 && fast_knob) {
    GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(sinkhorn_kernel<true>), (int)smem));
    GR_CHECK_CUDA(launch_pdl(sinkhorn_kernel<true>, dim3(P), dim3(kSink128Threads), (size_t)(smem), static_cast<cudaStream_t>(stream), scores, row_masks, col_masks, alpha, K,
                                                                                          num_iterations, inf, out));
  } else {
    if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(sinkhorn_kernel<false>), (int)smem));
    GR_CHECK_CUDA(launch_pdl(sinkhorn_kernel<false>, dim3(P), dim3(kSinkThreads), (size_t)(smem), static_cast<cudaStream_t>(stream), scores, row_masks, col_masks, alpha, K,
                                                                                         num_iterations, inf, out));
  }
  GR_CHECK_LAUNCH("sinkhorn_kernel");
  return GR_OK;
}
