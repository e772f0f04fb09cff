// K2/K4: KPConv blocks and the whole KPConvFPN issued from ONE C-ABI call
// (reference: geotransformer/modules/kpconv/modules.py:53-225, experiments/geotransformer.gaussian_splatting.indoor/
// backbone.py:164-212).
//
// Host-side orchestration only: every tensor op below is one of this library's kernels.  Round 1 issued the ~400
// launches of the backbone from Python (about 5 ms of interpreter time per pair, as much as the kernels take); from
// C++ the same sequence costs well under a millisecond of host time, and a host thread per CUDA stream can drive
// several pairs at once (the C-ABI call releases the GIL).  The GroupNorm statistics of every Linear / KPConv
// output come out of the product's own epilogue (GnStatsOut), so no block re-reads its activations for them.
#include <stdlib.h>

#include "common.cuh"

extern "C" {
size_t gr_kpconv_aggregate_workspace_size(int64_t n_support);
int gr_kpconv_aggregate(const float* s_feats, int C, const float* q_points, const float* s_points,
                        const int64_t* neighbor_idx, int H, int64_t ld_idx, int M, int Ns, const float* kernel_points,
                        int n_kernel_points, float sigma, float* A, float* row_div, void* ws, size_t ws_bytes,
                        void* stream);
size_t gr_group_norm_workspace_size(int64_t n_rows, int groups);
int gr_group_norm(const float* x, int64_t n_rows, int C, int groups, const float* gamma, const float* beta, float eps,
                  const float* add, int act, float* y, void* ws, size_t ws_bytes, void* stream);
int gr_maxpool(const float* x, int Ns, int C, const int64_t* idx, int H, int64_t ld_idx, int M, float* out, void* stream);
int gr_upsample_concat(const float* coarse, int Nc, int C1, const int64_t* idx, int64_t ld_idx, const float* skip,
                       int C2, int M, float* out, void* stream);
}

namespace gr {

// Stack arena over the caller's workspace.  `dry` walks the same allocation sequence without a buffer (sizing).
struct Arena {
  char* base;
  size_t off, cap, peak;
  bool dry;
  Arena(void* p, size_t bytes, bool dry_run) : base(static_cast<char*>(p)), off(0), cap(bytes), peak(0), dry(dry_run) {}
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    if (off > peak) peak = off;
    return r;
  }
  bool ok() const { return dry || (base != nullptr && peak <= cap); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

#define GR_TRY(expr)                 \
  do {                               \
    const int rc__ = (expr);         \
    if (rc__ != GR_OK) return rc__;  \
  } while (0)

static size_t gn_blocks_capacity(long long rows) {
  return (size_t)((rows + kGnReduceRows - 1) / kGnReduceRows + 2);
}

// y = act(GroupNorm(x) [+ add]) where x was just produced by a product that may have left partial statistics
static int norm_after_product(const float* x, long long rows, int C, int groups, float eps, const GnStatsOut& gn,
                              const float* gamma, const float* beta, const float* add, int act, float* y, Arena& ar, void* st) {
  float2* stats = ar.take<float2>((size_t)groups);
  const size_t gws_bytes = gr_group_norm_workspace_size(rows, groups);
  char* gws = ar.take<char>(gws_bytes);
  if (ar.dry) return GR_OK;
  if (!ar.ok()) return GR_ERR_WORKSPACE;
  if (gn.nblk > 0) return group_norm_from_partial(x, rows, C, groups, gn.partial, gn.nblk, gamma, beta, eps, add, act, y, stats, st);
  return gr_group_norm(x, rows, C, groups, gamma, beta, eps, add, act, y, gws, gws_bytes, st);
}

// UnaryBlock / LastUnaryBlock (modules.py:53-101)
static int unary(const gr_unary_weights& w, const float* x, long long rows, int groups, float eps, const float* add,
                 int act_after_add, float* y, Arena& ar, void* st) {
  const int in = w.in_channels, out = w.out_channels;
  const size_t mk = ar.mark();
  int rc = GR_OK;
  if (!w.gn_weight) {
    if (!ar.dry) rc = gemm_ex(x, in, w.weight, in, 1, y, out, (int)rows, out, in, 1.f, w.bias, nullptr, nullptr, 0, 0, st,
                              w.weight_packed, nullptr, w.weight_packed16, w.inv_scale16);
  } else {
    float* tmp = ar.take<float>((size_t)rows * out);
    GnStatsOut gn{ar.take<double2>(gn_blocks_capacity(rows) * groups), gn_blocks_capacity(rows), groups, 0};
    if (!ar.dry) {
      if (!ar.ok()) return GR_ERR_WORKSPACE;
      rc = gemm_ex(x, in, w.weight, in, 1, tmp, out, (int)rows, out, in, 1.f, w.bias, nullptr, nullptr, 0, 0, st, w.weight_packed, &gn,
                   w.weight_packed16, w.inv_scale16);
    }
    if (rc == GR_OK)
      rc = norm_after_product(tmp, rows, out, groups, eps, gn, w.gn_weight, w.gn_bias, add, w.leaky_relu ? 2 : act_after_add, y, ar, st);
  }
  ar.release(mk);
  return rc;
}

// Optional row chunking of the KPConv operand A (M, 15 C) (GAUSSREG_KPCONV_CHUNK_MB > 0): the layer is cut into row
// chunks whose operand stays in the 126 MB L2 between the aggregation and the contraction.  Measured on B200 in both
// rounds and OFF by default: 40 MB chunks cost +0.5 ms per pair, 24 MB chunks +1.0 ms -- the extra launches and the
// partial waves of the smaller grids outweigh the DRAM traffic saved (the two kernels are latency-, not
// bandwidth-bound).  Chunk boundaries are multiples of 128 rows (tile and statistics granularity).
static int kpconv_chunk_rows(int M, int KC) {
  static int mb = -1;
  if (mb < 0) { const char* e = getenv("GAUSSREG_KPCONV_CHUNK_MB"); mb = e ? atoi(e) : 0; }
  if (mb <= 0) return M;
  long long rows = ((long long)mb << 20) / ((long long)KC * 4);
  rows = rows / 128 * 128;
  if (rows < 1024) rows = 1024;
  if (rows >= M) return M;
  // balance the chunks
  const int n = (int)((M + rows - 1) / rows);
  long long per = ((M + n - 1) / n + 127) / 128 * 128;
  return (int)per;
}

// KPConv (kpconv.py:79-122) [+ GroupNorm + LeakyReLU]
static int kpconv(const gr_kpconv_weights& w, const float* gn_w, const float* gn_b, int groups, float eps, const float* s_feats,
                  const float* q_pts, const float* s_pts, const int64_t* idx, int H, int64_t ld, int M, int Ns, float* y, Arena& ar,
                  void* st) {
  const int C = w.in_channels, Co = w.out_channels, KC = 15 * C;
  const size_t mk = ar.mark();
  const int chunk = kpconv_chunk_rows(M, KC);
  float* A = ar.take<float>((size_t)chunk * KC);
  float* row_div = ar.take<float>((size_t)M);
  const size_t fws_bytes = gr_kpconv_aggregate_workspace_size(Ns);
  char* fws = ar.take<char>(fws_bytes);
  float* tmp = gn_w ? ar.take<float>((size_t)M * Co) : y;
  GnStatsOut gn{nullptr, 0, groups, 0};
  size_t cap_blocks = 0;
  if (gn_w) { cap_blocks = gn_blocks_capacity(M) + (size_t)(M / chunk + 2); gn.partial = ar.take<double2>(cap_blocks * groups); }
  int rc = GR_OK;
  int nblk_total = 0;
  bool stats_ok = gn_w != nullptr;
  if (!ar.dry) {
    if (!ar.ok()) return GR_ERR_WORKSPACE;
    for (int r0 = 0; r0 < M && rc == GR_OK; r0 += chunk) {
      const int rows = M - r0 < chunk ? M - r0 : chunk;
      // the support-row flags (workspace) are recomputed per chunk by gr_kpconv_aggregate: Ns bytes, negligible
      rc = gr_kpconv_aggregate(s_feats, C, q_pts + 3ll * r0, s_pts, idx + (long long)r0 * ld, H, ld, rows, Ns, w.kernel_points, 15,
                               w.sigma, A, row_div + r0, fws, fws_bytes, st);
      if (rc != GR_OK) break;
      GnStatsOut g{gn.partial ? gn.partial + (size_t)nblk_total * groups : nullptr, cap_blocks - (size_t)nblk_total, groups, 0};
      GnStatsOut* gp = (gn_w && stats_ok) ? &g : nullptr;
      if (w.weights_kmajor && KC % 4 == 0)
        rc = gemm_ex(A, KC, w.weights_kmajor, KC, 1, tmp + (size_t)r0 * Co, Co, rows, Co, KC, 1.f, w.bias, row_div + r0, nullptr, 0, 0, st,
                     w.weights_kmajor_packed, gp, w.weights_kmajor_packed16, w.inv_scale16);
      else
        rc = gemm_ex(A, KC, w.weights, Co, 0, tmp + (size_t)r0 * Co, Co, rows, Co, KC, 1.f, w.bias, row_div + r0, nullptr, 0, 0, st, nullptr,
                     nullptr), gp = nullptr;
      if (gp && g.nblk > 0) nblk_total += g.nblk; else stats_ok = false;  // one chunk without statistics -> separate pass
    }
    gn.nblk = stats_ok ? nblk_total : 0;
  }
  if (rc == GR_OK && gn_w) rc = norm_after_product(tmp, M, Co, groups, eps, gn, gn_w, gn_b, nullptr, 2, y, ar, st);
  ar.release(mk);
  return rc;
}

// ConvBlock / ResidualBlock (modules.py:104-225); y (M, out) is caller-allocated
static int block(const gr_block_weights& b, int groups, float eps, const float* s_feats, const float* q_pts, const float* s_pts,
                 const int64_t* idx, int H, int64_t ld, int M, int Ns, float* y, Arena& ar, void* st) {
  if (b.kind == 0) return kpconv(b.conv, b.gn_conv_weight, b.gn_conv_bias, groups, eps, s_feats, q_pts, s_pts, idx, H, ld, M, Ns, y, ar, st);
  const size_t mk = ar.mark();
  const int in = b.unary1.in_channels > 0 ? b.unary1.in_channels : b.conv.in_channels;
  const int mid = b.conv.in_channels;
  const float* x = s_feats;
  if (b.unary1.in_channels > 0) {
    float* x1 = ar.take<float>((size_t)Ns * mid);
    GR_TRY(unary(b.unary1, s_feats, Ns, groups, eps, nullptr, 0, x1, ar, st));
    x = x1;
  }
  float* c = ar.take<float>((size_t)M * b.conv.out_channels);
  GR_TRY(kpconv(b.conv, b.gn_conv_weight, b.gn_conv_bias, groups, eps, x, q_pts, s_pts, idx, H, ld, M, Ns, c, ar, st));
  const float* sc = s_feats;
  if (b.strided) {
    float* pooled = ar.take<float>((size_t)M * in);
    if (!ar.dry) {
      if (!ar.ok()) return GR_ERR_WORKSPACE;
      GR_TRY(gr_maxpool(s_feats, Ns, in, idx, H, ld, M, pooled, st));
    }
    sc = pooled;
  }
  static int dual = -1;
  if (dual < 0) { const char* e = getenv("GAUSSREG_GN_DUAL"); dual = e ? atoi(e) : 1; }
  if (dual && b.shortcut.in_channels > 0 && b.shortcut.gn_weight && b.unary2.gn_weight && !b.shortcut.leaky_relu && !b.unary2.leaky_relu &&
      b.shortcut.out_channels == b.unary2.out_channels && b.unary2.out_channels % 4 == 0) {
    // both products leave their GroupNorm statistics; ONE apply pass then forms LeakyReLU(GN(unary2) + GN(shortcut))
    const int out = b.unary2.out_channels;
    float* ts = ar.take<float>((size_t)M * out);
    float* tu = ar.take<float>((size_t)M * out);
    GnStatsOut gs{ar.take<double2>(gn_blocks_capacity(M) * groups), gn_blocks_capacity(M), groups, 0};
    GnStatsOut gu{ar.take<double2>(gn_blocks_capacity(M) * groups), gn_blocks_capacity(M), groups, 0};
    float2* st_s = ar.take<float2>((size_t)groups);
    float2* st_u = ar.take<float2>((size_t)groups);
    if (!ar.dry) {
      if (!ar.ok()) return GR_ERR_WORKSPACE;
      GR_TRY(gemm_ex(sc, b.shortcut.in_channels, b.shortcut.weight, b.shortcut.in_channels, 1, ts, out, M, out, b.shortcut.in_channels, 1.f,
                     b.shortcut.bias, nullptr, nullptr, 0, 0, st, b.shortcut.weight_packed, &gs, b.shortcut.weight_packed16,
                     b.shortcut.inv_scale16));
      GR_TRY(gemm_ex(c, b.unary2.in_channels, b.unary2.weight, b.unary2.in_channels, 1, tu, out, M, out, b.unary2.in_channels, 1.f,
                     b.unary2.bias, nullptr, nullptr, 0, 0, st, b.unary2.weight_packed, &gu, b.unary2.weight_packed16,
                     b.unary2.inv_scale16));
      if (gs.nblk > 0 && gu.nblk > 0) {
        GR_TRY(group_norm_finalize(gs.partial, gs.nblk, M, out, groups, eps, st_s, st));
        GR_TRY(group_norm_finalize(gu.partial, gu.nblk, M, out, groups, eps, st_u, st));
        GR_TRY(group_norm_apply2(tu, ts, M, out, groups, st_u, st_s, b.unary2.gn_weight, b.unary2.gn_bias, b.shortcut.gn_weight,
                                 b.shortcut.gn_bias, 2, y, st));
        ar.release(mk);
        return GR_OK;
      }
      // a product ran on a path without epilogue statistics: normalise the two results the long way (ts -> ts, then tu + ts)
      GR_TRY(norm_after_product(ts, M, out, groups, eps, gs, b.shortcut.gn_weight, b.shortcut.gn_bias, nullptr, 0, ts, ar, st));
      GR_TRY(norm_after_product(tu, M, out, groups, eps, gu, b.unary2.gn_weight, b.unary2.gn_bias, ts, 2, y, ar, st));
    } else {
      norm_after_product(ts, M, out, groups, eps, gs, nullptr, nullptr, nullptr, 0, ts, ar, st);  // sizing of the fallback (two norms)
      norm_after_product(tu, M, out, groups, eps, gu, nullptr, nullptr, nullptr, 0, y, ar, st);
    }
    ar.release(mk);
    return GR_OK;
  }
  if (b.shortcut.in_channels > 0) {
    float* s2 = ar.take<float>((size_t)M * b.shortcut.out_channels);
    GR_TRY(unary(b.shortcut, sc, M, groups, eps, nullptr, 0, s2, ar, st));
    sc = s2;
  }
  // LeakyReLU(GN(Linear(c)) + shortcut): the add and the activation ride in the GroupNorm apply pass
  GR_TRY(unary(b.unary2, c, M, groups, eps, sc, 2, y, ar, st));
  ar.release(mk);
  return GR_OK;
}

static int block_out_channels(const gr_block_weights& b) { return b.kind == 0 ? b.conv.out_channels : b.unary2.out_channels; }

// stage of the QUERY points of block i and whether it is the strided block that enters that stage
static const int kBlockStage[GR_FPN_BLOCKS] = {0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4};
static const bool kBlockStrided[GR_FPN_BLOCKS] = {false, false, true, false, false, true, false, false, true, false, false, true, false, false};

// start_block > 0: `feats` are the features AFTER block start_block - 1 (on the support points of block start_block): the
// caller ran the first blocks itself, e.g. the two stage-0 blocks while the rest of the pyramid was still being built
static int fpn(const gr_fpn_weights& W, const gr_pyramid& P, const float* feats, float* out_l2, float* out_l3, float* out_l4,
               float* out_f5, Arena& ar, void* st, int start_block = 0) {
  const int G = W.group_norm;
  const float eps = W.eps;
  const float* cur = feats;      // features on the support points of the next block
  const float* stage_out[GR_FPN_STAGES] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int stage_ch[GR_FPN_STAGES] = {0, 0, 0, 0, 0};
  for (int i = start_block; i < GR_FPN_BLOCKS; ++i) {
    const gr_block_weights& b = W.blocks[i];
    const int qs = kBlockStage[i];
    const bool strided = kBlockStrided[i];
    if ((b.strided != 0) != strided && b.kind == 1) return GR_ERR_BAD_ARG;
    const int ss = strided ? qs - 1 : qs;
    const int M = P.n_points[qs], Ns = P.n_points[ss];
    const int64_t* idx = strided ? P.subsampling[ss] : P.neighbors[qs];
    const int H = strided ? P.subsampling_w[ss] : P.neighbors_w[qs];
    const int64_t ld = strided ? P.subsampling_ld[ss] : P.neighbors_ld[qs];
    const int oc = block_out_channels(b);
    const bool last = i == GR_FPN_BLOCKS - 1;
    float* y = (last && out_f5) ? out_f5 : ar.take<float>((size_t)M * oc);
    GR_TRY(block(b, G, eps, cur, P.points[qs], P.points[ss], idx, H, ld, M, Ns, y, ar, st));
    cur = y;
    stage_out[qs] = y;
    stage_ch[qs] = oc;
  }
  // decoders (backbone.py:195-208): nearest upsample (column 0 of the upsampling table) | skip -> unary
  struct Dec { const gr_unary_weights* w; int fine; float* out; };
  const Dec decs[3] = {{&W.decoder4, 3, out_l4}, {&W.decoder3, 2, out_l3}, {&W.decoder2, 1, out_l2}};
  const float* coarse = stage_out[4];
  int coarse_ch = stage_ch[4];
  for (const Dec& d : decs) {
    const int f = d.fine, M = P.n_points[f], Nc = P.n_points[f + 1];
    const size_t mk = ar.mark();
    if (!ar.dry && d.w->in_channels != coarse_ch + stage_ch[f]) return GR_ERR_BAD_ARG;
    const bool split = d.w->split_k == coarse_ch && d.w->weight_packed_lo && d.w->weight_packed_hi;
    float* y = d.out;
    if (split) {
      // cat[up(coarse), skip] W^T = up(coarse W_lo^T) + skip W_hi^T: the coarse half of the product runs on the Nc
      // coarse rows, its result is gathered to the fine rows straight into the output buffer and rides in the second
      // product's epilogue as the residual (in place), together with the bias and the GroupNorm statistics.
      const int out = d.w->out_channels, in = d.w->in_channels, C2 = stage_ch[f];
      float* yc = ar.take<float>((size_t)Nc * out);
      float* tmp = ar.take<float>((size_t)M * out);
      GnStatsOut gn{ar.take<double2>(gn_blocks_capacity(M) * G), gn_blocks_capacity(M), G, 0};
      if (!y) y = ar.take<float>((size_t)M * out);
      int rc = GR_OK;
      if (!ar.dry) {
        if (!ar.ok()) return GR_ERR_WORKSPACE;
        float* dst = d.w->gn_weight ? tmp : y;
        GR_TRY(gemm_ex(coarse, coarse_ch, d.w->weight, in, 1, yc, out, Nc, out, coarse_ch, 1.f, nullptr, nullptr, nullptr, 0, 0, st,
                       d.w->weight_packed_lo, nullptr, d.w->weight_packed16_lo, d.w->inv_scale16_lo));
        GR_TRY(gr_upsample_concat(yc, Nc, out, P.upsampling[f], P.upsampling_ld[f], nullptr, 0, M, dst, st));
        rc = gemm_ex(stage_out[f], C2, d.w->weight + coarse_ch, in, 1, dst, out, M, out, C2, 1.f, d.w->bias, nullptr, dst, out, 0, st,
                     d.w->weight_packed_hi, d.w->gn_weight ? &gn : nullptr, d.w->weight_packed16_hi, d.w->inv_scale16_hi);
      }
      if (rc == GR_OK && d.w->gn_weight)
        rc = norm_after_product(tmp, M, out, G, eps, gn, d.w->gn_weight, d.w->gn_bias, nullptr, d.w->leaky_relu ? 2 : 0, y, ar, st);
      GR_TRY(rc);
    } else {
      float* cat = ar.take<float>((size_t)M * (coarse_ch + stage_ch[f]));
      if (!ar.dry) {
        if (!ar.ok()) return GR_ERR_WORKSPACE;
        GR_TRY(gr_upsample_concat(coarse, Nc, coarse_ch, P.upsampling[f], P.upsampling_ld[f], stage_out[f], stage_ch[f], M, cat, st));
      }
      if (!y) y = ar.take<float>((size_t)M * d.w->out_channels);  // cannot persist past release: only used when dry / unwanted
      GR_TRY(unary(*d.w, cat, M, G, eps, nullptr, 0, y, ar, st));
    }
    if (d.out) ar.release(mk);
    coarse = y;
    coarse_ch = d.w->out_channels;
  }
  return ar.ok() ? GR_OK : GR_ERR_WORKSPACE;
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_unary_block_workspace_size(int64_t rows, int out_channels, int groups) {
  Arena ar(nullptr, 0, true);
  gr_unary_weights w{};
  w.in_channels = 1; w.out_channels = out_channels; w.gn_weight = reinterpret_cast<const float*>(1);
  unary(w, nullptr, rows, groups, 0.f, nullptr, 0, nullptr, ar, nullptr);
  return ar.peak + 256;
}

extern "C" int gr_unary_block(const gr_unary_weights* w, const float* x, int64_t rows, int groups, float eps, const float* add,
                              int act_after_add, float* y, void* ws, size_t ws_bytes, void* stream) {
  if (!w || rows < 0 || w->in_channels <= 0 || w->out_channels <= 0 || !w->weight) return GR_ERR_BAD_ARG;
  if (rows == 0) return GR_OK;
  if (!x || !y) return GR_ERR_BAD_ARG;
  if (w->gn_weight && (groups <= 0 || w->out_channels % groups != 0 || !w->gn_bias)) return GR_ERR_BAD_ARG;
  Arena ar(ws, ws_bytes, false);
  return unary(*w, x, rows, groups, eps, add, act_after_add, y, ar, stream);
}

extern "C" size_t gr_kpconv_block_workspace_size(int M, int Ns, int in_channels, int out_channels, int groups) {
  Arena ar(nullptr, 0, true);
  gr_kpconv_weights w{};
  w.in_channels = in_channels; w.out_channels = out_channels;
  kpconv(w, reinterpret_cast<const float*>(1), nullptr, groups, 0.f, nullptr, nullptr, nullptr, nullptr, 1, 1, M, Ns, nullptr, ar, nullptr);
  return ar.peak + 256;
}

extern "C" int gr_kpconv_block(const gr_kpconv_weights* w, const float* gn_weight, const float* gn_bias, int groups, float eps,
                               const float* s_feats, const float* q_points, const float* s_points, const int64_t* neighbor_idx,
                               int H, int64_t ld_idx, int M, int Ns, float* y, void* ws, size_t ws_bytes, void* stream) {
  if (!w || M < 0 || Ns < 0 || H <= 0 || w->in_channels <= 0 || w->out_channels <= 0 || !w->weights || !w->kernel_points)
    return GR_ERR_BAD_ARG;
  if (M == 0) return GR_OK;
  if (!s_feats || !q_points || !s_points || !neighbor_idx || !y) return GR_ERR_BAD_ARG;
  if (gn_weight && (groups <= 0 || w->out_channels % groups != 0 || !gn_bias)) return GR_ERR_BAD_ARG;
  Arena ar(ws, ws_bytes, false);
  return kpconv(*w, gn_weight, gn_bias, groups, eps, s_feats, q_points, s_points, neighbor_idx, H, ld_idx, M, Ns, y, ar, stream);
}

extern "C" size_t gr_kpconv_fpn_workspace_size(const gr_fpn_weights* w, const gr_pyramid* pyr) {
  if (!w || !pyr) return 0;
  Arena ar(nullptr, 0, true);
  fpn(*w, *pyr, nullptr, reinterpret_cast<float*>(1), reinterpret_cast<float*>(1), reinterpret_cast<float*>(1),
      reinterpret_cast<float*>(1), ar, nullptr);
  return ar.peak + 256;
}

extern "C" int gr_kpconv_fpn_from(const gr_fpn_weights* w, const gr_pyramid* pyr, int start_block, const float* feats, float* out_l2,
                                  float* out_l3, float* out_l4, float* out_f5, void* ws, size_t ws_bytes, void* stream);

extern "C" int gr_kpconv_fpn(const gr_fpn_weights* w, const gr_pyramid* pyr, const float* feats, float* out_l2, float* out_l3,
                             float* out_l4, float* out_f5, void* ws, size_t ws_bytes, void* stream) {
  return gr_kpconv_fpn_from(w, pyr, 0, feats, out_l2, out_l3, out_l4, out_f5, ws, ws_bytes, stream);
}

/* Same, entering at block `start_block` (0, or 2 = after the two stage-0 blocks encoder1_1 / encoder1_2, whose output the
 * caller passes as `feats`): stage 0 needs nothing but the input cloud and its own neighbour table, so those blocks can
 * be queued before the rest of the pyramid is known. */
extern "C" int gr_kpconv_fpn_from(const gr_fpn_weights* w, const gr_pyramid* pyr, int start_block, const float* feats, float* out_l2,
                                  float* out_l3, float* out_l4, float* out_f5, void* ws, size_t ws_bytes, void* stream) {
  if (!w || !pyr || !feats || !out_l2 || !out_l3 || !out_l4 || !out_f5 || w->group_norm <= 0) return GR_ERR_BAD_ARG;
  if (start_block != 0 && start_block != 2) return GR_ERR_BAD_ARG;
  for (int s = 0; s < GR_FPN_STAGES; ++s) {
    if (pyr->n_points[s] <= 0 || !pyr->points[s] || !pyr->neighbors[s] || pyr->neighbors_w[s] <= 0) return GR_ERR_BAD_ARG;
    if (s + 1 < GR_FPN_STAGES && (!pyr->subsampling[s] || !pyr->upsampling[s] || pyr->subsampling_w[s] <= 0 || pyr->upsampling_w[s] <= 0))
      return GR_ERR_BAD_ARG;
  }
  if (!ws) return GR_ERR_WORKSPACE;
  Arena ar(ws, ws_bytes, false);
  return fpn(*w, *pyr, feats, out_l2, out_l3, out_l4, out_f5, ar, stream, start_block);
}
