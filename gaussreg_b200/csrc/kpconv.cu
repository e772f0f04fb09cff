// K1: KPConv neighbour gather + kernel-point correlation.
//
// Reference: geotransformer/modules/kpconv/kpconv.py:79-122.  For every query m and kernel point k
//     A[m, k, c] = sum_h max(0, 1 - |(s_h - q_m) - kp_k| / sigma) * feat[idx[m,h], c]
// followed by the dense contraction  out[m, :] = (A[m, :] . W[(k,c), :]) / max(1, #{h : sum_c feat > 0}) + bias
// which runs in the GEMM (gemm.cu, "NN" form, row_div + bias epilogue).
//
// The shadow neighbour (idx == Ns, a point at +1e6) has zero influence and zero features, so it is
// skipped rather than gathered.
#include "common.cuh"

namespace gr {

constexpr int kKP = 15;

// flag[j] = (sum_c feat[j, c] > 0)   (kpconv.py:113-114); one warp per row
__global__ void __launch_bounds__(256) row_positive_kernel(const float* __restrict__ feats, int n, int C,
                                                           unsigned char* __restrict__ flag) {
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
  if (smem > 48 * 1024) GR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int qpb = QPW * warps;
  kern<<<ceil_div(M, qpb), warps * 32, smem, st>>>(feats, C, q, s, idx, H, ldi, M, Ns, kp, sigma, flag, A, row_div);
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
    row_positive_kernel<<<ceil_div(Ns, 8), 256, 0, st>>>(s_feats, Ns, C, flag);
    GR_CHECK_LAUNCH("row_positive_kernel");
  }
  const long long* idx = reinterpret_cast<const long long*>(neighbor_idx);
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
