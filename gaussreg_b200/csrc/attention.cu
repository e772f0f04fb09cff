// T2/T3: attention score kernels.
//
// RPE self-attention (geotransformer/modules/transformer/rpe_transformer.py:50-70):
//     score[h,n,m] = ( q_h[n].k_h[m] + sum_c q_h[n,c] * (W_p e[n,m] + b_p)_{h,c} ) / sqrt(d_head)
// The reference materialises p = proj_p(e) as (1,H,N,N,64).  Here the second term is reassociated to
//     (W_p,h^T q_h[n]) . e[n,m] + q_h[n].b_p,h  =  U[h,n,:] . e[n,m,:] + qb[h,n]
// (same mathematics, different rounding: SURVEY.md section 8(a) row T2), so e is streamed once per layer
// and nothing of size N*N*C is written.  U and qb come out of two small GEMMs.
#include <stdlib.h>

#include "common.cuh"

namespace gr {

// One CTA per query row n.  P[h, n, :] = softmax_m(score[h, n, m]);  C = H * DH, H <= 8.
template <int H>
__global__ void __launch_bounds__(256) rpe_scores_softmax_kernel(const float* __restrict__ q, const float* __restrict__ kmat,
                                                                 const float* __restrict__ U, const float* __restrict__ qb,
                                                                 const float* __restrict__ emb, int N, int C, float scale,
                                                                 float* __restrict__ P, long long ldq, long long ldk) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  float* sU = sm;                 // [H][C]
  float* sq = sU + H * C;         // [C]
  float* ss = sq + C;             // [H][N] scores
  __shared__ float red[H][8];
  const int n = blockIdx.x;
  const int DH = C / H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < H * C; i += blockDim.x) sU[i] = U[((long long)(i / C) * N + n) * C + (i % C)];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sq[i] = q[(long long)n * ldq + i];
  __syncthreads();
  float qbh[H];
#pragma unroll
  for (int h = 0; h < H; ++h) qbh[h] = qb[(long long)h * N + n];

  // one thread per key m: the thread walks its own embedding row e[n,m,:] and key row k[m,:] sequentially
  // (float4), U / q come from shared memory as warp broadcasts -> no cross-lane reduction at all
  const float* erow = emb + (long long)n * N * C;
  const int c4_per_head = DH >> 2;
  for (int m = threadIdx.x; m < N; m += blockDim.x) {
    const float4* e4 = reinterpret_cast<const float4*>(erow + (long long)m * C);
    const float4* k4 = reinterpret_cast<const float4*>(kmat + (long long)m * ldk);
    float accp[H], acce[H];
#pragma unroll
    for (int h = 0; h < H; ++h) { accp[h] = 0.f; acce[h] = 0.f; }
#pragma unroll
    for (int hc = 0; hc < H; ++hc) {
      for (int i = 0; i < c4_per_head; ++i) {
        const int c4 = hc * c4_per_head + i;
        const float4 e = __ldg(e4 + c4);
        const float4 kv = __ldg(k4 + c4);
        const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * c4);
        acce[hc] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, fmaf(qv.z, kv.z, fmaf(qv.w, kv.w, acce[hc]))));
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float4 u = *reinterpret_cast<const float4*>(sU + h * C + 4 * c4);
          accp[h] = fmaf(u.x, e.x, fmaf(u.y, e.y, fmaf(u.z, e.z, fmaf(u.w, e.w, accp[h]))));
        }
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) ss[h * N + m] = (acce[h] + (accp[h] + qbh[h])) * scale;
  }
  __syncthreads();
  // softmax over m for each head
  for (int h = 0; h < H; ++h) {
    float mx = -INFINITY;
    for (int m = threadIdx.x; m < N; m += blockDim.x) mx = fmaxf(mx, ss[h * N + m]);
    mx = warp_max(mx);
    if (lane == 0) red[h][warp] = mx;
  }
  __syncthreads();
  float hmax[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float mx = red[h][0];
    for (int w = 1; w < nwarp; ++w) mx = fmaxf(mx, red[h][w]);
    hmax[h] = mx;
  }
  __syncthreads();
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
      const float e = expf(ss[h * N + m] - hmax[h]);
      ss[h * N + m] = e;
      s += e;
    }
    s = warp_sum(s);
    if (lane == 0) red[h][warp] = s;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += red[h][w];
    const float inv = 1.0f / s;
    float* out = P + ((long long)h * N + n) * N;
    for (int m = threadIdx.x; m < N; m += blockDim.x) out[m] = ss[h * N + m] * inv;
  }
}

// ---- T3 fused: vanilla multi-head cross-attention (vanilla_transformer.py:43-60), H = 4 heads of 64 channels ----
//     out[n, h*64 + d] = sum_m softmax_m( q_h[n].k_h[m] / 8 ) v_h[m, d]
// One CTA per (16 query rows, head).  The keys and then the values of the head flow through ONE ring of kXaDepth
// 64-row chunks fed by cp.async (chunk ci < nchunks is a key chunk, the rest value chunks; kXaDepth - 1 in flight, so
// the value stream also runs ahead through the softmax); the 16 x M score block never leaves shared memory.  Both
// products run as mma.sync m16n8k8 TF32 with the 3xTF32 split of gemm.cu (fp32-level accuracy): per 64-row chunk warp w
// owns keys 8w..8w+7 of the score block, then output channels 8w..8w+7 of P.V.  (An FFMA version of this kernel
// issued 90 K warp instructions per CTA and took 29 us; three separate launches -- scores, softmax, P.V -- 27 us.)
constexpr int kXaRows = 16, kXaChunk = 64, kXaPitchK = 68, kXaPitchV = 72, kXaThreads = 256, kXaDepth = 6;
constexpr int kXaBuf = kXaChunk * kXaPitchV;  // floats per ring buffer (key chunks use pitch 68, value chunks 72:
                                              // each makes its mma fragment loads bank-conflict free)

__device__ __forceinline__ void xa_cp_async16(void* dst, const void* src, int src_bytes) {  // src_bytes < 16: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void xa_mma(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// the tensor core reads the top 19 bits of an fp32 word: the raw word is the hi part, x - trunc(x) the lo part
__device__ __forceinline__ unsigned xa_lo(float x) { return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u)); }

__host__ __device__ inline int xa_score_pitch(int M) {  // >= M + 8 (the last 8-wide k step may overhang), = 4 (mod 32):
  const int mp = (M + 11) & ~3;                         // rows g, g+8 and columns t, t+4 then hit 32 different banks
  return mp + ((36 - (mp & 31)) & 31);
}

__global__ void __launch_bounds__(kXaThreads) cross_attention_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k,
                                                                     long long ldk, const float* __restrict__ v, long long ldv, int N, int M,
                                                                     float scale, float* __restrict__ out, long long ldo) {
  pdl_wait();
  pdl_trigger();
  constexpr int DH = 64;
  extern __shared__ __align__(16) float xa[];
  const int SP = xa_score_pitch(M);
  float* qs = xa;                                          // [16][68]
  float* ring = qs + kXaRows * kXaPitchK;                  // [kXaDepth][kXaBuf]
  float* sc = ring + kXaDepth * kXaBuf;                    // [16][SP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * kXaRows, h = blockIdx.y;
  const int nchunks = (M + kXaChunk - 1) / kXaChunk;
  auto issue = [&](int ci) {  // always commits a group, possibly empty, so that group counting stays uniform
    if (ci < 2 * nchunks) {
      const bool keys = ci < nchunks;
      const float* base = keys ? k : v;
      const long long ld = keys ? ldk : ldv;
      const int c = keys ? ci : ci - nchunks;
      const int pitch = keys ? kXaPitchK : kXaPitchV;
      float* dst = ring + (ci % kXaDepth) * kXaBuf;
#pragma unroll
      for (int i = 0; i < kXaChunk * (DH / 4) / kXaThreads; ++i) {
        const int e = tid + i * kXaThreads;
        const int r = e >> 4, c4 = e & 15;
        const int m = c * kXaChunk + r;
        const bool ok = m < M;
        xa_cp_async16(dst + r * pitch + 4 * c4, base + (ok ? (long long)m * ld + h * DH + 4 * c4 : 0), ok ? 16 : 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // q first: behind kXaDepth - 1 chunks of prefetch these 4 KB would wait for 80 KB of ring traffic
  float qreg[kXaRows * DH / kXaThreads];
#pragma unroll
  for (int i = 0; i < kXaRows * DH / kXaThreads; ++i) {
    const int e = tid + i * kXaThreads, r = e >> 6, d = e & 63;
    qreg[i] = n0 + r < N ? q[(long long)(n0 + r) * ldq + h * DH + d] : 0.f;
  }
#pragma unroll
  for (int ci = 0; ci < kXaDepth - 1; ++ci) issue(ci);
#pragma unroll
  for (int i = 0; i < kXaRows * DH / kXaThreads; ++i) {
    const int e = tid + i * kXaThreads, r = e >> 6, d = e & 63;
    qs[r * kXaPitchK + d] = qreg[i];
  }
  for (int e = tid; e < kXaRows * (SP - M); e += kXaThreads) {  // zero probabilities in the padding columns
    const int r = e / (SP - M), c = e % (SP - M);
    sc[r * SP + M + c] = 0.f;
  }
  __syncthreads();
  // A fragments of q (16 x 64): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4) per 8-wide k step
  unsigned qh[8][4], ql[8][4];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const float* ap = qs + g * kXaPitchK + 8 * ks + t;
    const float a0 = ap[0], a1 = ap[8 * kXaPitchK], a2 = ap[4], a3 = ap[8 * kXaPitchK + 4];
    qh[ks][0] = __float_as_uint(a0); qh[ks][1] = __float_as_uint(a1); qh[ks][2] = __float_as_uint(a2); qh[ks][3] = __float_as_uint(a3);
    ql[ks][0] = xa_lo(a0); ql[ks][1] = xa_lo(a1); ql[ks][2] = xa_lo(a2); ql[ks][3] = xa_lo(a3);
  }
  float o[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f}, o2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ci = 0; ci < 2 * nchunks; ++ci) {
    if (ci == nchunks) {
      // ---- softmax: warp w owns rows 2 w, 2 w + 1 (every score was written before the barrier that closed chunk nchunks-1)
      for (int r = 2 * warp; r < 2 * warp + 2; ++r) {
        float* row = sc + r * SP;
        float mx = -INFINITY;
        for (int m = lane; m < M; m += 32) mx = fmaxf(mx, row[m]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int m = lane; m < M; m += 32) { const float e = expf(row[m] - mx); row[m] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int m = lane; m < M; m += 32) row[m] *= inv;
      }
    }
    issue(ci + kXaDepth - 1);                                       // into the buffer chunk ci - 1 just released
    asm volatile("cp.async.wait_group %0;" ::"n"(kXaDepth - 1) : "memory");  // chunk ci has landed (this thread's part)
    __syncthreads();                                                // ... everybody's part; also publishes the softmax
    const float* buf = ring + (ci % kXaDepth) * kXaBuf;
    if (ci < nchunks) {
      // ---- scores of keys 8 warp .. 8 warp + 7 of this chunk: B fragment b0 (k = t, n = g), b1 (k = t + 4, n = g)
      const float* kr = buf + (8 * warp + g) * kXaPitchK + t;
      // three independent accumulation chains (lo.hi, hi.lo, hi.hi): with two warps per scheduler a single chain of
      // 24 dependent HMMAs per chunk was the critical path
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const float b0 = kr[8 * ks], b1 = kr[8 * ks + 4];
        xa_mma(acc1, ql[ks], __float_as_uint(b0), __float_as_uint(b1));
        xa_mma(acc2, qh[ks], xa_lo(b0), xa_lo(b1));
        xa_mma(acc, qh[ks], __float_as_uint(b0), __float_as_uint(b1));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += acc1[e] + acc2[e];
      // C fragment: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
      const int m = ci * kXaChunk + 8 * warp + 2 * t;
      if (m < M) { sc[g * SP + m] = acc[0] * scale; sc[(g + 8) * SP + m] = acc[2] * scale; }
      if (m + 1 < M) { sc[g * SP + m + 1] = acc[1] * scale; sc[(g + 8) * SP + m + 1] = acc[3] * scale; }
    } else {
      // ---- P.V over 64 values, output channels 8 warp .. 8 warp + 7
      const int mbase = (ci - nchunks) * kXaChunk;
      const float* pr = sc + g * SP + mbase + t;
      const float* vr = buf + t * kXaPitchV + 8 * warp + g;
      const int ksteps = M - mbase >= kXaChunk ? 8 : (M - mbase + 7) >> 3;  // padding probabilities and values are zero
#pragma unroll 8
      for (int ks = 0; ks < ksteps; ++ks) {
        const float a0 = pr[8 * ks], a1 = pr[8 * SP + 8 * ks], a2 = pr[8 * ks + 4], a3 = pr[8 * SP + 8 * ks + 4];
        const unsigned ah[4] = {__float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3)};
        const unsigned al[4] = {xa_lo(a0), xa_lo(a1), xa_lo(a2), xa_lo(a3)};
        const float b0 = vr[8 * ks * kXaPitchV], b1 = vr[(8 * ks + 4) * kXaPitchV];
        xa_mma(o1, al, __float_as_uint(b0), __float_as_uint(b1));
        xa_mma(o2, ah, xa_lo(b0), xa_lo(b1));
        xa_mma(o, ah, __float_as_uint(b0), __float_as_uint(b1));
      }
    }
    __syncthreads();  // buffer ci % depth is free again
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) o[e] += o1[e] + o2[e];
  {
    float* op = out + h * DH + 8 * warp + 2 * t;
    if (n0 + g < N) { op[(long long)(n0 + g) * ldo] = o[0]; op[(long long)(n0 + g) * ldo + 1] = o[1]; }
    if (n0 + g + 8 < N) { op[(long long)(n0 + g + 8) * ldo] = o[2]; op[(long long)(n0 + g + 8) * ldo + 1] = o[3]; }
  }
}

// ---- v2 (C = 256, H = 4): warp-cooperative, coalesced stream of the embedding ---------------------------------
// One CTA per query row n, one warp per group of 8 keys.  Lane l owns the channels of float4 #l and #(l+32) of a
// 256-channel row, holds U[h, n, those 8 channels] for the four heads in registers (32 floats), and reads the
// eight embedding rows of the group with sixteen independent, fully coalesced 512-byte warp loads (256 B in
// flight per lane, streamed past L1).  The 8 keys x 4 heads = 32 per-lane partial sums are folded across the
// warp by a transposing butterfly (31 shuffles for all 32 sums instead of 5 per sum); lane L ends up with the
// total of key L/4, head L%4.  The q.k^T term arrives as raw batched-GEMM output in P (gr_rpe_attention_probs
// issues that product first) and is added before the scale, in the reference's order (qk + (p-term)).
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}

template <int OFF, int CNT>
__device__ __forceinline__ void butterfly_step(float (&v)[32], int lane) {
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const float send = up ? v[i] : v[i + CNT];
    const float keep = up ? v[i + CNT] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

constexpr int kRpeKeys = 8;  // keys per warp iteration

// U: head h, row n at U + h * u_head + n * C.  qb == nullptr: the q.b_p term is computed here from q (row pitch ldq) and
// b_p -- 256 products per query row, not worth a launch.  P: row (h, n) at P + (h * N + n) * ldp.
__global__ void __launch_bounds__(256, 2) rpe_scores_softmax_v2_kernel(const float* __restrict__ U, long long u_head,
                                                                       const float* __restrict__ qb, const float* __restrict__ q,
                                                                       long long ldq, const float* __restrict__ bp,
                                                                       const float* __restrict__ emb, int N, float scale,
                                                                       float* __restrict__ P, long long ldp) {
  pdl_wait();
  pdl_trigger();
  constexpr int H = 4, C = 256, C4 = C / 4;
  extern __shared__ float ss[];  // [H][N] scores of this query row
  __shared__ float red[H][8];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;

  float4 u[H][2];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float4* up = reinterpret_cast<const float4*>(U + (long long)h * u_head + (long long)n * C);
    u[h][0] = __ldg(up + lane);
    u[h][1] = __ldg(up + 32 + lane);
  }
  float qbl;  // q_h[n] . b_p,h for this lane's head (lane & 3) after the butterfly
  if (qb) {
    qbl = qb[(long long)(lane & 3) * N + n];
  } else {
    // lane l covers channels 8 l .. 8 l + 7, i.e. head l / 8; eight lanes fold to one head
    const float4* q4 = reinterpret_cast<const float4*>(q + (long long)n * ldq) + 2 * lane;
    const float4* b4 = reinterpret_cast<const float4*>(bp) + 2 * lane;
    float part = dot4(__ldg(q4 + 1), __ldg(b4 + 1), dot4(__ldg(q4), __ldg(b4), 0.f));
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    qbl = __shfl_sync(0xffffffffu, part, 8 * (lane & 3));
  }
  const float4* erow = reinterpret_cast<const float4*>(emb) + (long long)n * N * C4;
  const int ngroups = (N + kRpeKeys - 1) / kRpeKeys;
  for (int g = warp; g < ngroups; g += nwarp) {
    const int m0 = g * kRpeKeys;
    float4 e[kRpeKeys][2];
#pragma unroll
    for (int kk = 0; kk < kRpeKeys; ++kk) {
      const int m = m0 + kk;
      if (m < N) {
        e[kk][0] = ld_stream4(erow + (long long)m * C4 + lane);
        e[kk][1] = ld_stream4(erow + (long long)m * C4 + 32 + lane);
      } else {
        e[kk][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        e[kk][1] = e[kk][0];
      }
    }
    float v[32];
#pragma unroll
    for (int kk = 0; kk < kRpeKeys; ++kk)
#pragma unroll
      for (int h = 0; h < H; ++h) v[kk * H + h] = dot4(u[h][1], e[kk][1], dot4(u[h][0], e[kk][0], 0.f));
    butterfly_step<16, 16>(v, lane);
    butterfly_step<8, 8>(v, lane);
    butterfly_step<4, 4>(v, lane);
    butterfly_step<2, 2>(v, lane);
    butterfly_step<1, 1>(v, lane);
    const int m = m0 + (lane >> 2);
    if (m < N) ss[(lane & 3) * N + m] = v[0] + qbl;
  }
  __syncthreads();
  // score = (q.k + (U.e + q.b_p)) * scale, then softmax over m per head
  for (int h = 0; h < H; ++h) {
    const float* qk = P + ((long long)h * N + n) * ldp;
    float mx = -INFINITY;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
      const float s = (qk[m] + ss[h * N + m]) * scale;
      ss[h * N + m] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) red[h][warp] = mx;
  }
  __syncthreads();
  float hmax[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float mx = red[h][0];
    for (int w = 1; w < nwarp; ++w) mx = fmaxf(mx, red[h][w]);
    hmax[h] = mx;
  }
  __syncthreads();
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
      const float ev = expf(ss[h * N + m] - hmax[h]);
      ss[h * N + m] = ev;
      s += ev;
    }
    s = warp_sum(s);
    if (lane == 0) red[h][warp] = s;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += red[h][w];
    const float inv = 1.0f / s;
    float* out = P + ((long long)h * N + n) * ldp;
    for (int m = threadIdx.x; m < N; m += blockDim.x) out[m] = ss[h * N + m] * inv;
  }
}

// in-place row softmax, one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, long long rows, int cols) {
  pdl_wait();
  pdl_trigger();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float* p = x + r * cols;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, p[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { const float e = expf(p[c] - mx); p[c] = e; s += e; }
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int c = lane; c < cols; c += 32) p[c] *= inv;
}

// F.normalize(x, p=2, dim=1): x / max(|x|, eps)
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const float* __restrict__ x, long long rows, int C, float eps,
                                                                float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = x[r * C + c]; s = fmaf(v, v, s); }
  const float nrm = fmaxf(sqrtf(warp_sum(s)), eps);
  for (int c = lane; c < C; c += 32) y[r * C + c] = x[r * C + c] / nrm;
}

}  // namespace gr

using namespace gr;

extern "C" int gr_gemm(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB, int trans_b,
                       float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch, float alpha, const float* bias,
                       const float* row_div, const float* residual, int64_t ldr, int64_t strideR, int act, void* stream);

static int rpe_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_RPE"); v = e ? atoi(e) : 2; }  // 1: thread-per-key kernel, 2: warp-cooperative
  return v;
}

static int rpe_threads() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_RPE_THREADS"); v = e ? atoi(e) : 256; }
  return v == 128 ? 128 : 256;
}

namespace gr {
/* The streaming form with everything the transformer driver wants to pass: U with its own head stride (so that one
 * product can serve both stacked clouds), the q.b_p term either precomputed (qb) or fused (qb == nullptr, bp given),
 * and P rows on a pitch ldp >= N (a multiple of 4 keeps the P.V product on 16-byte loads).  C = 256, H = 4. */
int rpe_attention_probs_ex(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* U, int64_t u_head, const float* qb,
                           const float* bp, const float* emb, int N, float* P, int64_t ldp, void* stream) {
  constexpr int C = 256, H = 4, dh = C / H;
  if (N <= 0 || ldp < N || (size_t)H * N * sizeof(float) > 100 * 1024) return GR_ERR_BAD_ARG;
  if (!q || !k || !U || (!qb && !bp) || !emb || !P) return GR_ERR_BAD_ARG;
  if (!qb && ((ldq & 3) != 0 || (reinterpret_cast<uintptr_t>(q) & 15) != 0 || (reinterpret_cast<uintptr_t>(bp) & 15) != 0))
    return GR_ERR_BAD_ARG;
  const float scale = 1.0f / sqrtf((float)dh);
  // raw q_h . k_h^T into P, then the streaming kernel adds the position term and normalises in place
  int rc = gr_gemm(q, ldq, dh, k, ldk, dh, 1, P, ldp, (int64_t)N * ldp, N, N, dh, H, 1.f, nullptr, nullptr, nullptr, 0, 0, 0, stream);
  if (rc != GR_OK) return rc;
  const size_t smem = (size_t)H * N * sizeof(float);
  auto kern = rpe_scores_softmax_v2_kernel;
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), (int)smem));
  GR_CHECK_CUDA(launch_pdl(kern, dim3(N), dim3(rpe_threads()), (size_t)(smem), static_cast<cudaStream_t>(stream), U, (long long)u_head, qb, q,
                           (long long)ldq, bp, emb, N, scale, P, (long long)ldp));
  GR_CHECK_LAUNCH("rpe_scores_softmax_v2_kernel");
  return GR_OK;
}

static size_t cross_attention_smem(int M) {
  return ((size_t)kXaRows * kXaPitchK + (size_t)kXaDepth * kXaBuf + (size_t)kXaRows * xa_score_pitch(M)) * sizeof(float);
}
bool cross_attention_fits(int M) { return cross_attention_smem(M) <= 200 * 1024; }  // M <= ~1500 keys

/* out (N, 4*64 on pitch ldo) = multi-head softmax(q k^T / 8) v, heads side by side in the 256 columns of q / k / v. */
int cross_attention(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, int N, int M, float* out,
                    int64_t ldo, void* stream) {
  if (N <= 0 || M <= 0 || !q || !k || !v || !out) return GR_ERR_BAD_ARG;
  if (((ldk | ldv) & 3) != 0 || ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) != 0) return GR_ERR_BAD_ARG;
  const size_t smem = cross_attention_smem(M);
  if (!cross_attention_fits(M)) return GR_ERR_CAPACITY;
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(cross_attention_kernel), (int)smem));
  GR_CHECK_CUDA(launch_pdl(cross_attention_kernel, dim3((N + kXaRows - 1) / kXaRows, 4), dim3(kXaThreads), smem, static_cast<cudaStream_t>(stream), q,
                           (long long)ldq, k, (long long)ldk, v, (long long)ldv, N, M, 0.125f, out, (long long)ldo));
  GR_CHECK_LAUNCH("cross_attention_kernel");
  return GR_OK;
}
}  // namespace gr

/* T2: q,k (N,C) with row pitches ldq / ldk ; U (H,N,C) ; qb (H,N) ; emb (N,N,C) -> P (H,N,N) softmax probabilities. */
extern "C" int gr_rpe_attention_probs_ld(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* U, const float* qb,
                                         const float* emb, int N, int C, int num_heads, float* P, void* stream) {
  if (N <= 0 || C <= 0 || num_heads != 4 || C % num_heads != 0) return GR_ERR_BAD_ARG;
  if (!q || !k || !U || !qb || !emb || !P) return GR_ERR_BAD_ARG;
  const float scale = 1.0f / sqrtf((float)(C / num_heads));
  if (C == 256 && rpe_variant() == 2 && (size_t)num_heads * N * sizeof(float) <= 100 * 1024)
    return gr::rpe_attention_probs_ex(q, ldq, k, ldk, U, (int64_t)N * C, qb, nullptr, emb, N, P, N, stream);
  if (ldk % 4 != 0) return GR_ERR_BAD_ARG;
  const size_t smem = ((size_t)num_heads * C + C + (size_t)num_heads * N) * sizeof(float);
  if (smem > 200 * 1024) return GR_ERR_CAPACITY;
  auto kern = rpe_scores_softmax_kernel<4>;
  if (smem > 48 * 1024) GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(kern), (int)smem));
  GR_CHECK_CUDA(launch_pdl(kern, dim3(N), dim3(256), (size_t)(smem), static_cast<cudaStream_t>(stream), q, k, U, qb, emb, N, C, scale, P, ldq, ldk));
  GR_CHECK_LAUNCH("rpe_scores_softmax_kernel");
  return GR_OK;
}

extern "C" int gr_rpe_attention_probs(const float* q, const float* k, const float* U, const float* qb, const float* emb, int N,
                                      int C, int num_heads, float* P, void* stream) {
  return gr_rpe_attention_probs_ld(q, C, k, C, U, qb, emb, N, C, num_heads, P, stream);
}

extern "C" int gr_softmax_rows(float* x, int64_t rows, int cols, void* stream) {
  if (rows < 0 || cols <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(softmax_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, rows, cols));
  GR_CHECK_LAUNCH("softmax_rows_kernel");
  return GR_OK;
}

extern "C" int gr_l2_normalize_rows(const float* x, int64_t rows, int C, float eps, float* y, void* stream) {
  if (rows < 0 || C <= 0) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x || !y) return GR_ERR_BAD_ARG;
  GR_CHECK_CUDA(launch_pdl(l2_normalize_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), x, rows, C, eps, y));
  GR_CHECK_LAUNCH("l2_normalize_rows_kernel");
  return GR_OK;
}
