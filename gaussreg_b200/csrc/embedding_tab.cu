// T1 by tabulation: geometric structure embedding (geotransformer/modules/geotransformer/geotransformer.py:57-72,
// modules/transformer/positional_embedding.py:19-35) without the two 256x256 projections in the hot path.
//
// proj(sinusoid(x)) is, channel by channel, a function of ONE scalar:
//     f_c(x) = b_c + sum_i W[c,2i] sin(x w_i) + W[c,2i+1] cos(x w_i),   w_i = 10000^(-2i/C) <= 1,
// i.e. band-limited with highest angular frequency 1.  Such a function is reproduced below fp32 rounding by Hermite
// interpolation from a small table of exact node values: the angle index lives in [0, 180/sigma_a] (= [0, 12] for
// sigma_a = 15) and takes a cubic Hermite table with step 1/8 (99 nodes); the distance index dist/sigma_d takes a
// quintic Hermite table with step 1/2 over [0, 1024] (204.8 m at sigma_d = 0.2), whose first 129 nodes ([0, 64],
// 12.8 m) sit in shared memory and the rest (6 MB, L2-resident) is read from global memory.  Interpolation error
// h^4/384 |f''''| resp. h^6/46080 |f^(6)|: measured 7e-8 relative to the exact function, where the reference's own
// fp32 evaluation (sin of an fp32-rounded phase, fp32 GEMM) sits at 3e-7 — the table is the more accurate of the two.
// An index outside its table (or NaN) is evaluated directly from W (slow, exact).
//
// The tables are built ONCE per weight set in fp64 (gr_structure_embedding_build_table), laid out per 64-channel slice.
// The evaluation kernel is persistent: one CTA per SM, its slice of both tables (146 KB) in shared memory, one half-warp per
// (i, j) pair and 64 channels (four per lane), 18 conflict-free LDS.128 + 84 FMA per lane per pair.  Work per cloud: N^2 C outputs x
// 72 B of shared-memory reads — the kernel is bound by shared-memory bandwidth (and behind it by the N^2 C x 4 B HBM
// write), not by 2 N^2 (1+k) C^2 tensor FLOPs.
#include "common.cuh"

namespace gr {

constexpr int kTabSlice = 64;      // channels per CTA
constexpr int kTabInvHA = 8;       // 1 / step of the angle table
constexpr int kTabInvHD = 2;       // 1 / step of the distance table
constexpr int kTabND = 64 * kTabInvHD + 1;     // distance nodes held in shared memory: [0, 64]
constexpr int kTabNDG = 1024 * kTabInvHD + 1;  // distance nodes of the full table in global memory: [0, 1024]
constexpr int kTabThreads = 1024;

// nodes of the angle table: indices reach fl(pi_f32 * factor_a) (atan2f <= float(pi)); node n+1 must exist
static inline int tab_nodes_a(float sigma_a) {
  const float factor_a = (float)(180.0 / ((double)sigma_a * 3.141592653589793));
  const float amax = 3.14159274101257324f * factor_a;
  return (int)(amax * (float)kTabInvHA) + 3;
}
// rows (of 64 channels) per slice: angle nodes (f, h f'), then distance nodes (f, h f', h^2 f''); the first
// tab_rows_smem rows are the image every CTA copies into shared memory
__host__ __device__ static inline size_t tab_rows(int nA) { return (size_t)nA * 2 + (size_t)kTabNDG * 3; }
__host__ __device__ static inline size_t tab_rows_smem(int nA) { return (size_t)nA * 2 + (size_t)kTabND * 3; }

// one thread per (table row group, channel): exact node values in fp64
__global__ void __launch_bounds__(128) embedding_table_build_kernel(const float* __restrict__ div, int C, const float* __restrict__ Wd,
                                                                    const float* __restrict__ bd, const float* __restrict__ Wa,
                                                                    const float* __restrict__ ba, int nA, float* __restrict__ tab) {
  pdl_wait();
  pdl_trigger();
  const int node = blockIdx.x;  // [0, nA): angle nodes, then distance nodes
  const bool is_a = node < nA;
  const int n = is_a ? node : node - nA;
  const double h = is_a ? 1.0 / kTabInvHA : 1.0 / kTabInvHD;
  const double x = n * h;
  const float* W = is_a ? Wa : Wd;
  const float* b = is_a ? ba : bd;
  const size_t rows = tab_rows(nA);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double f = 0.0, f1 = 0.0, f2 = 0.0;
    for (int i = 0; i < C / 2; ++i) {
      const double om = (double)div[i];
      double s, co;
      sincos(x * om, &s, &co);
      const double w0 = (double)W[(size_t)c * C + 2 * i], w1 = (double)W[(size_t)c * C + 2 * i + 1];
      f += w0 * s + w1 * co;
      f1 += om * (w0 * co - w1 * s);
      f2 -= om * om * (w0 * s + w1 * co);
    }
    f += (double)b[c];
    float* base = tab + ((size_t)(c / kTabSlice) * rows) * kTabSlice + (c % kTabSlice);
    if (is_a) {
      base[(size_t)(2 * n) * kTabSlice] = (float)f;
      base[(size_t)(2 * n + 1) * kTabSlice] = (float)(h * f1);
    } else {
      float* d = base + (size_t)(2 * nA) * kTabSlice;
      d[(size_t)(3 * n) * kTabSlice] = (float)f;
      d[(size_t)(3 * n + 1) * kTabSlice] = (float)(h * f1);
      d[(size_t)(3 * n + 2) * kTabSlice] = (float)(h * h * f2);
    }
  }
}

// completes only after everything before it in the stream: a later kernel's pre-pdl_wait prologue may then read the table
__global__ void embedding_table_fence_kernel() {
  pdl_wait();
  pdl_trigger();
}

// exact evaluation of four channels for an index outside its table
__device__ __noinline__ float4 embedding_direct(float x, const float* __restrict__ W, const float* __restrict__ b,
                                                const float* __restrict__ div, int C, int c0) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < C / 2; ++i) {
    float s, c;
    sincosf(__fmul_rn(x, div[i]), &s, &c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] = fmaf(W[(size_t)(c0 + j) * C + 2 * i], s, acc[j]);
      acc[j] = fmaf(W[(size_t)(c0 + j) * C + 2 * i + 1], c, acc[j]);
    }
  }
  return make_float4(acc[0] + b[c0], acc[1] + b[c0 + 1], acc[2] + b[c0 + 2], acc[3] + b[c0 + 3]);
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// r = fma(w, a, r) / r = fma(w, a - b, b) on four channels
__device__ __forceinline__ void fma4(float4& r, float w, const float4& a) {
  r.x = fmaf(w, a.x, r.x); r.y = fmaf(w, a.y, r.y); r.z = fmaf(w, a.z, r.z); r.w = fmaf(w, a.w, r.w);
}
__device__ __forceinline__ float4 lerp4(float w, const float4& f0, const float4& f1) {
  return make_float4(fmaf(w, f1.x - f0.x, f0.x), fmaf(w, f1.y - f0.y, f0.y), fmaf(w, f1.z - f0.z, f0.z), fmaf(w, f1.w - f0.w, f0.w));
}

// cubic Hermite, node rows (f, h f'); sA already offset by the lane's channel
__device__ __forceinline__ float4 hermite3(const float* __restrict__ sA, int n, float t) {
  const float* p = sA + (size_t)n * (2 * kTabSlice);
  const float4 f0 = lds4(p), g0 = lds4(p + kTabSlice), f1 = lds4(p + 2 * kTabSlice), g1 = lds4(p + 3 * kTabSlice);
  const float t2 = t * t, t3 = t2 * t;
  const float h01 = fmaf(-2.f, t3, 3.f * t2);
  const float h10 = fmaf(-2.f, t2, t3) + t;
  const float h11 = t3 - t2;
  float4 r = lerp4(h01, f0, f1);
  fma4(r, h10, g0);
  fma4(r, h11, g1);
  return r;
}

// quintic Hermite, node rows (f, h f', h^2 f'')
__device__ __forceinline__ float4 hermite5(const float* __restrict__ p, float t) {
  const float t2 = t * t, t3 = t2 * t;
  // H3 = t^3 (10 - 15 t + 6 t^2); H1 = t - t^3 (6 - 8 t + 3 t^2); H2 = t^2/2 - t^3 (3 - 3 t + t^2)/2
  // H4 = -t^3 (4 - 7 t + 3 t^2);  H5 = t^3 (1 - 2 t + t^2)/2
  const float H3 = t3 * fmaf(t, fmaf(6.f, t, -15.f), 10.f);
  const float H1 = fmaf(-t3, fmaf(t, fmaf(3.f, t, -8.f), 6.f), t);
  const float H2 = 0.5f * fmaf(-t3, fmaf(t, t - 3.f, 3.f), t2);
  const float H4 = -t3 * fmaf(t, fmaf(3.f, t, -7.f), 4.f);
  const float H5 = 0.5f * t3 * fmaf(t, t - 2.f, 1.f);
  float4 r = lerp4(H3, lds4(p), lds4(p + 3 * kTabSlice));
  fma4(r, H1, lds4(p + kTabSlice));
  fma4(r, H2, lds4(p + 2 * kTabSlice));
  fma4(r, H4, lds4(p + 4 * kTabSlice));
  fma4(r, H5, lds4(p + 5 * kTabSlice));
  return r;
}

// torch.max propagates NaN; fmaxf does not
__device__ __forceinline__ float max_nan(float m, float v) { return (v != v) ? v : fmaxf(m, v); }

// One warp = two (i, j) pairs at a time: half-warp h takes pair 2p + h, each lane four channels (LDS.128: a quarter-warp
// reads 128 contiguous bytes of one table row, so the two halves never conflict).  The index arithmetic and the Hermite
// basis -- identical for every lane of a pair -- are thereby paid once per 128 outputs instead of once per 64.
template <int K>
__global__ void __launch_bounds__(kTabThreads, 1) structure_embedding_table_kernel(
    const float* __restrict__ d_idx, const float* __restrict__ a_idx, long long rows, const float* __restrict__ tab, int nA, int C,
    const float* __restrict__ div, const float* __restrict__ Wd, const float* __restrict__ bd, const float* __restrict__ Wa,
    const float* __restrict__ ba, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int slices = C / kTabSlice;
  const int slice = blockIdx.x % slices;
  const int cta = blockIdx.x / slices, ncta = gridDim.x / slices;
  const int trows = (int)tab_rows_smem(nA);
  const float* gD = tab + ((size_t)slice * tab_rows(nA) + (size_t)nA * 2) * kTabSlice + 4 * (threadIdx.x & 15);
  {
    // the table is static data, complete long before this launch (build + fence kernel): safe ahead of pdl_wait
    const float4* src = reinterpret_cast<const float4*>(tab + (size_t)slice * tab_rows(nA) * kTabSlice);
    float4* dst = reinterpret_cast<float4*>(sm);
    const int n4 = trows * kTabSlice / 4;
    for (int i = threadIdx.x; i < n4; i += kTabThreads) dst[i] = __ldg(src + i);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, sub = lane & 15;
  const float* sA = sm + 4 * sub;
  const float* sD = sm + (size_t)nA * 2 * kTabSlice + 4 * sub;
  const int c0 = slice * kTabSlice + 4 * sub;
  pdl_wait();
  pdl_trigger();
  __syncthreads();
  const long long nbatch = (rows + 31) >> 5;
  constexpr int kWarps = kTabThreads / 32;
  for (long long b = (long long)cta * kWarps + warp; b < nbatch; b += (long long)ncta * kWarps) {
    const long long r0 = b << 5;
    const long long r = r0 + lane;
    float xd = 0.f, xa[K];
#pragma unroll
    for (int j = 0; j < K; ++j) xa[j] = 0.f;
    if (r < rows) {
      xd = d_idx[r];
#pragma unroll
      for (int j = 0; j < K; ++j) xa[j] = a_idx[r * K + j];
    }
    const int np = (int)min(32ll, rows - r0);
#pragma unroll 2
    for (int p = 0; p < 16; ++p) {
      if (2 * p >= np) break;
      const int q = 2 * p + half;
      const float x = __shfl_sync(0xffffffffu, xd, q);
      float a[K];
#pragma unroll
      for (int j = 0; j < K; ++j) a[j] = __shfl_sync(0xffffffffu, xa[j], q);
      if (q >= np) continue;
      const float u = x * (float)kTabInvHD;
      const int n = (int)u;
      float4 acc;
      if (x >= 0.f && n < kTabND - 1) acc = hermite5(sD + (size_t)n * (3 * kTabSlice), u - (float)n);
      else if (x >= 0.f && n < kTabNDG - 1) acc = hermite5(gD + (size_t)n * (3 * kTabSlice), u - (float)n);  // L2-resident
      else acc = embedding_direct(x, Wd, bd, div, C, c0);
      float4 mx = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float ua = a[j] * (float)kTabInvHA;
        const int na = (int)ua;
        float4 v;
        if (a[j] >= 0.f && na < nA - 1) v = hermite3(sA, na, ua - (float)na);
        else v = embedding_direct(a[j], Wa, ba, div, C, c0);
        if (j == 0) mx = v;
        else mx = make_float4(max_nan(mx.x, v.x), max_nan(mx.y, v.y), max_nan(mx.z, v.z), max_nan(mx.w, v.w));
      }
      *reinterpret_cast<float4*>(out + (size_t)(r0 + q) * C + c0) = make_float4(acc.x + mx.x, acc.y + mx.y, acc.z + mx.z, acc.w + mx.w);
    }
  }
}

}  // namespace gr

using namespace gr;

/* number of floats of the table gr_structure_embedding_build_table writes for this (hidden_dim, sigma_a) */
extern "C" int64_t gr_structure_embedding_table_floats(int hidden_dim, float sigma_a) {
  if (hidden_dim <= 0 || hidden_dim % kTabSlice != 0 || !(sigma_a > 0.f)) return 0;
  return (int64_t)tab_rows(tab_nodes_a(sigma_a)) * hidden_dim;
}

/* T1 table: exact fp64 node values (and scaled derivatives) of proj_a(sinusoid(.)) on [0, 180/sigma_a] and of
 * proj_d(sinusoid(.)) on [0, 64].  Once per weight set.  W_* (hidden_dim, hidden_dim) row-major Linear weights. */
extern "C" int gr_structure_embedding_build_table(const float* div_term, int hidden_dim, const float* W_d, const float* b_d,
                                                  const float* W_a, const float* b_a, float sigma_a, float* table, void* stream) {
  if (hidden_dim <= 0 || hidden_dim % kTabSlice != 0 || !(sigma_a > 0.f)) return GR_ERR_BAD_ARG;
  if (!div_term || !W_d || !b_d || !W_a || !b_a || !table) return GR_ERR_BAD_ARG;
  const int nA = tab_nodes_a(sigma_a);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GR_CHECK_CUDA(launch_pdl(embedding_table_build_kernel, dim3(nA + kTabNDG), dim3(128), (size_t)0, st, div_term, hidden_dim, W_d, b_d,
                           W_a, b_a, nA, table));
  GR_CHECK_LAUNCH("embedding_table_build_kernel");
  GR_CHECK_CUDA(launch_pdl(embedding_table_fence_kernel, dim3(1), dim3(32), (size_t)0, st));
  GR_CHECK_LAUNCH("embedding_table_fence_kernel");
  return GR_OK;
}

/* T1 from the table: d_idx (rows), a_idx (rows, angle_k) -> out (rows, hidden_dim) = f_d(d) + max_k f_a(a_k).
 * The raw weights are only touched for indices outside the table.  GR_ERR_CAPACITY when the table of this sigma_a does
 * not fit the shared memory of one SM (callers then use gr_structure_embedding_fused*). */
extern "C" int gr_structure_embedding_tabulated(const float* d_idx, const float* a_idx, int64_t rows, int angle_k,
                                                const float* table, float sigma_a, const float* div_term, int hidden_dim,
                                                const float* W_d, const float* b_d, const float* W_a, const float* b_a, float* out,
                                                void* stream) {
  if (rows < 0 || angle_k < 1 || angle_k > 3 || hidden_dim <= 0 || hidden_dim % kTabSlice != 0 || !(sigma_a > 0.f))
    return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!d_idx || !a_idx || !table || !div_term || !W_d || !b_d || !W_a || !b_a || !out) return GR_ERR_BAD_ARG;
  const int nA = tab_nodes_a(sigma_a);
  const size_t smem = tab_rows_smem(nA) * kTabSlice * sizeof(float);
  if (smem > 227 * 1024) return GR_ERR_CAPACITY;
  int dev = 0, sms = 0;
  GR_CHECK_CUDA(cudaGetDevice(&dev));
  GR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int slices = hidden_dim / kTabSlice;
  const long long nbatch = (rows + 31) / 32;
  long long per_slice = sms / slices > 0 ? sms / slices : 1;
  const long long need = (nbatch + kTabThreads / 32 - 1) / (kTabThreads / 32);
  if (per_slice > need) per_slice = need;
  const dim3 grid((unsigned)(per_slice * slices));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define GR_TAB_LAUNCH(KK)                                                                                                      \
  GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(structure_embedding_table_kernel<KK>), (int)smem));            \
  GR_CHECK_CUDA(launch_pdl(structure_embedding_table_kernel<KK>, grid, dim3(kTabThreads), smem, st, d_idx, a_idx, (long long)rows, \
                           table, nA, hidden_dim, div_term, W_d, b_d, W_a, b_a, out))
  if (angle_k == 1) { GR_TAB_LAUNCH(1); }
  else if (angle_k == 2) { GR_TAB_LAUNCH(2); }
  else { GR_TAB_LAUNCH(3); }
#undef GR_TAB_LAUNCH
  GR_CHECK_LAUNCH("structure_embedding_table_kernel");
  return GR_OK;
}

/* T1 from the superpoint coordinates in ONE call: gr_embedding_indices followed by gr_structure_embedding_tabulated (the two
 * wrappers cost the host ~30 us between them, during which the GPU -- idle since the stage-size read -- has nothing to do).
 * d_idx (N,N), a_idx (N,N,angle_k), knn (N,angle_k) are the intermediate buffers of gr_embedding_indices. */
extern "C" int gr_structure_embedding_points(const float* points, int N, float sigma_d, float sigma_a, int angle_k, const float* table,
                                             const float* div_term, int hidden_dim, const float* W_d, const float* b_d,
                                             const float* W_a, const float* b_a, float* d_idx, float* a_idx, int32_t* knn,
                                             float* out, void* stream) {
  const int rc = gr_embedding_indices(points, N, sigma_d, sigma_a, angle_k, d_idx, a_idx, knn, stream);
  if (rc != GR_OK) return rc;
  return gr_structure_embedding_tabulated(d_idx, a_idx, (int64_t)N * N, angle_k, table, sigma_a, div_term, hidden_dim, W_d, b_d, W_a,
                                          b_a, out, stream);
}
