// N1: Gaussian-splat cloud -> network input.
//
// Reference: experiments/geotransformer.gaussian_splatting.indoor/demo.py:30-75 (_read_ply_by_opacity: opacity
// sigmoid, 5/95-percentile crop, SH degree-3 view-dependent colour) and :81-124 (load_data: bounding-box centring
// and volume rescale), geotransformer/utils/graphics_utils.py:34-89 (eval_sh).
//
// The cloud is (n, ld) fp32 rows in 3DGS property order without normals (gs_fusion.py:172-184):
//   xyz 0..2 | f_dc 3..5 | f_rest 6..50 (channel-major: 15 R, 15 G, 15 B) | opacity 51 | scale 52..54 | rot 55..58.
// Everything O(n) runs here; the O(1) scalar glue (numpy's percentile interpolation, the f32 mean division, the
// volume rule) stays in the Python host code, evaluated with numpy exactly as the reference does.
#include "common.cuh"
#include "scan.cuh"

namespace gr {

constexpr int kMaxQueries = 16;

struct SelectQueries {
  int col[kMaxQueries];
  unsigned int rank[kMaxQueries];
  int n;
};

// ---- exact order statistics by 8-bit radix select ------------------------------------------------------
// state[q] = (prefix of the key found so far, rank still to skip inside that prefix)
struct SelState { unsigned int prefix; unsigned int rank; };

__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ cloud, long long n, int ld, SelectQueries qs,
                                                          const SelState* __restrict__ state, int shift,
                                                          unsigned int* __restrict__ hist) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned int sh[kMaxQueries * 256];
  for (int i = threadIdx.x; i < qs.n * 256; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  unsigned int prefix[kMaxQueries];
#pragma unroll
  for (int q = 0; q < kMaxQueries; ++q) prefix[q] = q < qs.n ? state[q].prefix : 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* row = cloud + i * ld;
#pragma unroll
    for (int q = 0; q < kMaxQueries; ++q) {
      if (q >= qs.n) break;
      const unsigned int key = f2ord(row[qs.col[q]]);
      // bits above (shift + 8) must equal the prefix chosen by the earlier passes
      const bool match = shift == 24 || (key >> (shift + 8)) == prefix[q];
      if (match) atomicAdd(&sh[q * 256 + ((key >> shift) & 255u)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < qs.n * 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one warp per query: pick the digit holding the wanted rank, clear the histogram for the next pass
__global__ void select_pick_kernel(unsigned int* __restrict__ hist, SelState* __restrict__ state, int nq, int last,
                                   float* __restrict__ out_values) {
  pdl_wait();
  pdl_trigger();
  const int q = blockIdx.x;
  if (q >= nq) return;
  __shared__ unsigned int cnt[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { cnt[i] = hist[q * 256 + i]; hist[q * 256 + i] = 0u; }
  __syncthreads();
  if (threadIdx.x == 0) {
    SelState s = state[q];
    unsigned int cum = 0;
    int d = 0;
    for (; d < 255; ++d) {
      if (cum + cnt[d] > s.rank) break;
      cum += cnt[d];
    }
    s.prefix = (s.prefix << 8) | (unsigned int)d;
    s.rank -= cum;
    state[q] = s;
    if (last) out_values[q] = ord2f(s.prefix);
  }
}

__global__ void select_init_kernel(SelState* __restrict__ state, SelectQueries qs, unsigned int* __restrict__ hist) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < qs.n) { state[i].prefix = 0u; state[i].rank = qs.rank[i]; }
  if (i < kMaxQueries * 256) hist[i] = 0u;
}

// ---- selection mask + ordered compaction ----------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f32(float o) {
  // demo.py:34: 1 / (1 + np.exp(-opacity)) on a float32 array; exp evaluated in double and rounded once
  const float e = (float)exp(-(double)o);
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

struct CropBox { double lo[3], hi[3]; };

__global__ void __launch_bounds__(256) gaussian_flag_kernel(const float* __restrict__ cloud, long long n, int ld, int opacity_col,
                                                            float opacity_min, CropBox box, uint32_t* __restrict__ flag) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint32_t f = 0u;
  if (i < n) {
    const float* row = cloud + i * ld;
    bool keep = sigmoid_f32(row[opacity_col]) > opacity_min;  // demo.py:43 (opacity>0.7)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = (double)row[a];
      keep = keep && (v < box.hi[a]) && (v > box.lo[a]);  // demo.py:40-42, strict on both sides
    }
    f = keep ? 1u : 0u;
  }
  flag[i] = f;
}

__global__ void __launch_bounds__(256) gaussian_compact_kernel(const uint32_t* __restrict__ scan, long long n,
                                                               long long* __restrict__ out_index, long long* __restrict__ out_count) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *out_count = (long long)scan[n];
  if (i >= n) return;
  if (scan[i + 1] != scan[i]) out_index[scan[i]] = i;
}

// ---- gather xyz + statistics ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_points_kernel(const float* __restrict__ cloud, int ld, const long long* __restrict__ index,
                                                            long long m, float* __restrict__ out_points,
                                                            unsigned int* __restrict__ mnmx) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float inf = __int_as_float(0x7f800000);
  float x = inf, y = inf, z = inf;
  const bool valid = i < m;
  if (valid) {
    const float* row = cloud + (index ? index[i] : i) * ld;
    x = row[0]; y = row[1]; z = row[2];
    out_points[3 * i] = x; out_points[3 * i + 1] = y; out_points[3 * i + 2] = z;
  }
  const float mnx = warp_min(x), mny = warp_min(y), mnz = warp_min(z);
  const float mxx = warp_max(valid ? x : -inf), mxy = warp_max(valid ? y : -inf), mxz = warp_max(valid ? z : -inf);
  if (lane_id() == 0 && mnx != inf) {
    atomicMin(&mnmx[0], f2ord(mnx)); atomicMin(&mnmx[1], f2ord(mny)); atomicMin(&mnmx[2], f2ord(mnz));
    atomicMax(&mnmx[3], f2ord(mxx)); atomicMax(&mnmx[4], f2ord(mxy)); atomicMax(&mnmx[5], f2ord(mxz));
  }
}

__global__ void mnmx_init_kernel(unsigned int* __restrict__ mnmx) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x < 3) mnmx[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) mnmx[threadIdx.x] = 0u;
}

// numpy reduces an (m,3) C-ordered float32 array over axis 0 row by row: three running float32 sums in INPUT
// ORDER (demo.py:62 points.mean(0)).  The same order here: one CTA stages 1024 rows at a time in shared memory,
// threads 0..2 add them sequentially (no reassociation, so the bits match numpy's).
__global__ void __launch_bounds__(1024) sequential_colsum_kernel(const float* __restrict__ pts, long long m,
                                                                 const unsigned int* __restrict__ mnmx, float* __restrict__ out_stats) {
  pdl_wait();
  pdl_trigger();
  __shared__ float buf[2][3 * 1024];
  float acc = 0.f;
  const long long nchunk = (m + 1023) / 1024;
  auto load = [&](long long c, int b) {
    const long long base = c * 1024 * 3, lim = m * 3;
    for (int k = threadIdx.x; k < 3 * 1024; k += 1024) buf[b][k] = (base + k < lim) ? pts[base + k] : 0.f;
  };
  if (nchunk > 0) load(0, 0);
  __syncthreads();
  for (long long c = 0; c < nchunk; ++c) {
    const int b = (int)(c & 1);
    if (c + 1 < nchunk) load(c + 1, b ^ 1);
    if (threadIdx.x < 3) {
      const int rows = (int)min((long long)1024, m - c * 1024);
      for (int r = 0; r < rows; ++r) acc = __fadd_rn(acc, buf[b][3 * r + threadIdx.x]);
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) out_stats[threadIdx.x] = acc;
  else if (threadIdx.x < 9) out_stats[threadIdx.x] = ord2f(mnmx[threadIdx.x - 3]);
}

// ---- SH degree-3 colour in double, operation for operation as numpy evaluates graphics_utils.py:57-88 -----
struct ViewPoint { double c[3]; };

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

__global__ void __launch_bounds__(128) gaussian_features_kernel(const float* __restrict__ cloud, int ld, const long long* __restrict__ index,
                                                                long long m, ViewPoint view, float* __restrict__ out_feats) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float* row = cloud + (index ? index[i] : i) * ld;
  // demo.py:66-67: dir = (p - view) / (|p - view| + 1e-6), all float64
  const double dx = dsub((double)row[0], view.c[0]), dy = dsub((double)row[1], view.c[1]), dz = dsub((double)row[2], view.c[2]);
  const double nrm = dadd(__dsqrt_rn(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz))), 1e-6);
  const double x = __ddiv_rn(dx, nrm), y = __ddiv_rn(dy, nrm), z = __ddiv_rn(dz, nrm);
  const double C0 = 0.28209479177387814, C1 = 0.4886025119029199;
  const double C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396};
  const double C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                        -0.4570457994644658, 1.445305721320277, -0.5900435899266435};
  const double xx = dmul(x, x), yy = dmul(y, y), zz = dmul(z, z);
  const double xy = dmul(x, y), yz = dmul(y, z), xz = dmul(x, z);
  // direction-only factors, parenthesised as Python evaluates them (left to right)
  const double b1 = dmul(C1, y), b2 = dmul(C1, z), b3 = dmul(C1, x);
  const double b4 = dmul(C2[0], xy), b5 = dmul(C2[1], yz);
  const double b6 = dmul(C2[2], dsub(dsub(dmul(2.0, zz), xx), yy));
  const double b7 = dmul(C2[3], xz), b8 = dmul(C2[4], dsub(xx, yy));
  const double b9 = dmul(dmul(C3[0], y), dsub(dmul(3.0, xx), yy));
  const double b10 = dmul(dmul(C3[1], xy), z);
  const double b11 = dmul(dmul(C3[2], y), dsub(dsub(dmul(4.0, zz), xx), yy));
  const double b12 = dmul(dmul(C3[3], z), dsub(dsub(dmul(2.0, zz), dmul(3.0, xx)), dmul(3.0, yy)));
  const double b13 = dmul(dmul(C3[4], x), dsub(dsub(dmul(4.0, zz), xx), yy));
  const double b14 = dmul(dmul(C3[5], z), dsub(xx, yy));
  const double b15 = dmul(dmul(C3[6], x), dsub(xx, dmul(3.0, yy)));
  float4 f;
  f.x = sigmoid_f32(row[51]);
  float rgb[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float* rest = row + 6 + 15 * ch;  // sh[ch, 1..15]
    double r = dmul(C0, (double)row[3 + ch]);
    r = dsub(r, dmul(b1, (double)rest[0]));
    r = dadd(r, dmul(b2, (double)rest[1]));
    r = dsub(r, dmul(b3, (double)rest[2]));
    r = dadd(r, dmul(b4, (double)rest[3]));
    r = dadd(r, dmul(b5, (double)rest[4]));
    r = dadd(r, dmul(b6, (double)rest[5]));
    r = dadd(r, dmul(b7, (double)rest[6]));
    r = dadd(r, dmul(b8, (double)rest[7]));
    r = dadd(r, dmul(b9, (double)rest[8]));
    r = dadd(r, dmul(b10, (double)rest[9]));
    r = dadd(r, dmul(b11, (double)rest[10]));
    r = dadd(r, dmul(b12, (double)rest[11]));
    r = dadd(r, dmul(b13, (double)rest[12]));
    r = dadd(r, dmul(b14, (double)rest[13]));
    r = dadd(r, dmul(b15, (double)rest[14]));
    // demo.py:70: np.clip(sh2rgb + 0.5, 0.0, 1.0) * 255, then .astype(float32) at :71
    double c = dadd(r, 0.5);
    c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
    rgb[ch] = (float)dmul(c, 255.0);
  }
  f.y = rgb[0]; f.z = rgb[1]; f.w = rgb[2];
  reinterpret_cast<float4*>(out_feats)[i] = f;
}

struct Center3 { float c[3]; };

// demo.py:85-110: points = points - center ; [points = points * scale]  (float32, each operation rounded)
__global__ void __launch_bounds__(256) points_normalize_kernel(float* __restrict__ pts, long long m, Center3 ctr, float scale,
                                                               int apply_scale) {
  pdl_wait();
  pdl_trigger();
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 3 * m) return;
  float v = __fsub_rn(pts[k], ctr.c[k % 3]);
  if (apply_scale) v = __fmul_rn(v, scale);
  pts[k] = v;
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_column_order_stats_workspace_size(int n_queries) {
  (void)n_queries;
  return (size_t)kMaxQueries * 256 * sizeof(unsigned int) + kMaxQueries * sizeof(SelState) + kMaxQueries * sizeof(unsigned int) + 1024;
}

/* out_values[j] = the ranks[j]-th smallest (0-based) entry of column cols[j] of cloud (n, ld).  cols / ranks are HOST
 * arrays, n_queries <= 16; out_values is a DEVICE array.  Exact (radix select on the order-preserving key). */
extern "C" int gr_column_order_stats(const float* cloud, int64_t n, int ld, const int32_t* cols, const int64_t* ranks,
                                     int n_queries, float* out_values, void* ws, size_t ws_bytes, void* stream) {
  if (n <= 0 || ld <= 0 || n_queries <= 0 || n_queries > kMaxQueries || !cloud || !cols || !ranks || !out_values ||
      n >= (1ll << 32))
    return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < gr_column_order_stats_workspace_size(n_queries)) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws, ws_bytes);
  unsigned int* hist = c.take<unsigned int>((size_t)kMaxQueries * 256);
  SelState* state = c.take<SelState>(kMaxQueries);
  SelectQueries qs;  // travels as a kernel argument: no host buffer outlives the call, no hidden synchronisation
  qs.n = n_queries;
  for (int q = 0; q < n_queries; ++q) {
    if (cols[q] < 0 || cols[q] >= ld || ranks[q] < 0 || ranks[q] >= n) return GR_ERR_BAD_ARG;
    qs.col[q] = cols[q];
    qs.rank[q] = (unsigned int)ranks[q];
  }
  for (int q = n_queries; q < kMaxQueries; ++q) { qs.col[q] = 0; qs.rank[q] = 0u; }
  GR_CHECK_CUDA(launch_pdl(select_init_kernel, dim3(ceil_div(kMaxQueries * 256, 256)), dim3(256), (size_t)(0), st, state, qs, hist));
  GR_CHECK_LAUNCH("select_init_kernel");
  const int blocks = (int)min((long long)148 * 8, (long long)ceil_div(n, 256));
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    GR_CHECK_CUDA(launch_pdl(select_hist_kernel, dim3(blocks), dim3(256), (size_t)(0), st, cloud, (long long)n, ld, qs, state, shift, hist));
    GR_CHECK_LAUNCH("select_hist_kernel");
    GR_CHECK_CUDA(launch_pdl(select_pick_kernel, dim3(n_queries), dim3(256), (size_t)(0), st, hist, state, n_queries, pass == 3, out_values));
    GR_CHECK_LAUNCH("select_pick_kernel");
  }
  return GR_OK;
}

extern "C" size_t gr_gaussian_select_workspace_size(int64_t n) {
  return ((size_t)n + 2) * sizeof(uint32_t) + scan_workspace_elems(n + 2) * sizeof(uint32_t) + 1024;
}

/* keep row i iff sigmoid(cloud[i, opacity_col]) > opacity_min and lo[a] < cloud[i, a] < hi[a] (a = 0..2, compared in
 * double).  lo / hi: HOST double[3].  out_index (device, capacity n): kept rows in ascending order; out_count (device). */
extern "C" int gr_gaussian_select(const float* cloud, int64_t n, int ld, int opacity_col, float opacity_min, const double* lo,
                                  const double* hi, int64_t* out_index, int64_t* out_count, void* ws, size_t ws_bytes,
                                  void* stream) {
  if (n < 0 || ld < 3 || opacity_col < 0 || opacity_col >= ld || !lo || !hi || !out_count || n >= (1ll << 31)) return GR_ERR_BAD_ARG;
  if (n > 0 && (!cloud || !out_index)) return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < gr_gaussian_select_workspace_size(n)) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws, ws_bytes);
  uint32_t* flag = c.take<uint32_t>((size_t)n + 2);
  uint32_t* scan_ws = c.take<uint32_t>(scan_workspace_elems(n + 2));
  CropBox box;
  for (int a = 0; a < 3; ++a) { box.lo[a] = lo[a]; box.hi[a] = hi[a]; }
  GR_CHECK_CUDA(launch_pdl(gaussian_flag_kernel, dim3(ceil_div(n + 1, 256)), dim3(256), (size_t)(0), st, cloud, (long long)n, ld, opacity_col, opacity_min, box, flag));
  GR_CHECK_LAUNCH("gaussian_flag_kernel");
  const int rc = exclusive_scan_u32(flag, flag, n + 1, scan_ws, st);
  if (rc != GR_OK) return rc;
  GR_CHECK_CUDA(launch_pdl(gaussian_compact_kernel, dim3(ceil_div(n > 0 ? n : 1, 256)), dim3(256), (size_t)(0), st, flag, (long long)n, reinterpret_cast<long long*>(out_index),
                                                                       reinterpret_cast<long long*>(out_count)));
  GR_CHECK_LAUNCH("gaussian_compact_kernel");
  return GR_OK;
}

/* out_points (m,3) = cloud[index, 0:3] (index NULL = identity); out_stats (device float[9]) = {float32 column sums
 * accumulated in input order, min xyz, max xyz}.  ws: 64 bytes of device scratch. */
extern "C" int gr_gather_points_stats(const float* cloud, int ld, const int64_t* index, int64_t m, float* out_points,
                                      float* out_stats, void* ws, size_t ws_bytes, void* stream) {
  if (m <= 0 || ld < 3 || !cloud || !out_points || !out_stats) return GR_ERR_BAD_ARG;
  if (!ws || ws_bytes < 64) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int* mnmx = static_cast<unsigned int*>(ws);
  GR_CHECK_CUDA(launch_pdl(mnmx_init_kernel, dim3(1), dim3(32), (size_t)(0), st, mnmx));
  GR_CHECK_LAUNCH("mnmx_init_kernel");
  GR_CHECK_CUDA(launch_pdl(gather_points_kernel, dim3(ceil_div(m, 256)), dim3(256), (size_t)(0), st, cloud, ld, reinterpret_cast<const long long*>(index), (long long)m,
                                                         out_points, mnmx));
  GR_CHECK_LAUNCH("gather_points_kernel");
  GR_CHECK_CUDA(launch_pdl(sequential_colsum_kernel, dim3(1), dim3(1024), (size_t)(0), st, out_points, (long long)m, mnmx, out_stats));
  GR_CHECK_LAUNCH("sequential_colsum_kernel");
  return GR_OK;
}

/* out_feats (m,4) = [sigmoid(opacity), 255 * clip(SH_deg3(dir) + 0.5, 0, 1) for R,G,B], dir = (p - view) / (|p - view| + 1e-6)
 * in double (demo.py:63-72).  view_point: HOST double[3].  The cloud must use the 59-attribute layout (ld >= 59). */
extern "C" int gr_gaussian_features(const float* cloud, int ld, const int64_t* index, int64_t m, const double* view_point,
                                    float* out_feats, void* stream) {
  if (m < 0 || ld < 59 || !view_point) return GR_ERR_BAD_ARG;
  if (m == 0) return GR_OK;
  if (!cloud || !out_feats) return GR_ERR_BAD_ARG;
  ViewPoint v;
  for (int a = 0; a < 3; ++a) v.c[a] = view_point[a];
  GR_CHECK_CUDA(launch_pdl(gaussian_features_kernel, dim3(ceil_div(m, 128)), dim3(128), (size_t)(0), static_cast<cudaStream_t>(stream), cloud, ld, reinterpret_cast<const long long*>(index), (long long)m, v, out_feats));
  GR_CHECK_LAUNCH("gaussian_features_kernel");
  return GR_OK;
}

/* points (m,3) <- (points - center) [* scale]  in float32 (demo.py:87,93,99-110).  center3: HOST float[3]. */
extern "C" int gr_points_normalize(float* points, int64_t m, const float* center3, float scale, int apply_scale, void* stream) {
  if (m < 0 || !center3) return GR_ERR_BAD_ARG;
  if (m == 0) return GR_OK;
  if (!points) return GR_ERR_BAD_ARG;
  Center3 c;
  for (int a = 0; a < 3; ++a) c.c[a] = center3[a];
  GR_CHECK_CUDA(launch_pdl(points_normalize_kernel, dim3(ceil_div(3 * m, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), points, (long long)m, c, scale,
                                                                                            apply_scale));
  GR_CHECK_LAUNCH("points_normalize_kernel");
  return GR_OK;
}
