// T4: the RPE-conditional transformer (geotransformer/modules/transformer/conditional_transformer.py:97-117 with
// rpe_transformer.py:18-131, vanilla_transformer.py:15-129, output_layer.py:6-21) as ONE C-ABI call.
//
// Host-side orchestration only: every tensor op below is one of this library's kernels.  Running the ~130 launches
// of the six layers from C++ instead of from Python removes ~10 us of interpreter overhead per op, which is more
// than most of these superpoint-sized kernels take on a B200.
#include "common.cuh"

extern "C" {
int gr_gemm(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB, int trans_b,
            float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch, float alpha, const float* bias,
            const float* row_div, const float* residual, int64_t ldr, int64_t strideR, int act, void* stream);
int gr_layer_norm_add(const float* a, const float* b, int64_t rows, int C, const float* gamma, const float* beta,
                      float eps, float* y, void* stream);
int gr_rpe_attention_probs(const float* q, const float* k, const float* U, const float* qb, const float* emb, int N, int C,
                           int num_heads, float* P, void* stream);
int gr_softmax_rows(float* x, int64_t rows, int cols, void* stream);
}

namespace gr {

struct TfWs {
  float *q, *k, *v, *U, *qb, *P, *hid, *att, *ffn, *y;
  size_t bytes;
};

static TfWs carve_tf(void* ws, size_t ws_bytes, int N, int C, int H, bool* ok) {
  Carver c(ws, ws_bytes);
  TfWs w;
  w.q = c.take<float>((size_t)N * C);
  w.k = c.take<float>((size_t)N * C);
  w.v = c.take<float>((size_t)N * C);
  w.U = c.take<float>((size_t)H * N * C);
  w.qb = c.take<float>((size_t)H * N);
  w.P = c.take<float>((size_t)H * N * N);
  w.hid = c.take<float>((size_t)N * C);
  w.att = c.take<float>((size_t)N * C);
  w.ffn = c.take<float>((size_t)N * 2 * C);
  w.y = c.take<float>((size_t)N * C);
  w.bytes = c.off;
  *ok = c.ok;
  return w;
}

#define GR_TRY(expr)                 \
  do {                               \
    const int rc__ = (expr);         \
    if (rc__ != GR_OK) return rc__;  \
  } while (0)

static int linear(const float* x, int rows, int in, const float* W, const float* b, int out, float* y, int act, void* st) {
  return gr_gemm(x, in, 0, W, in, 0, 1, y, out, 0, rows, out, in, 1, 1.f, b, nullptr, nullptr, 0, 0, act, st);
}

// x (N,C) attends mem (M,C) [emb (N,N,C) when self]; result overwrites x
static int layer(const gr_layer_weights& L, float* x, int N, const float* mem, int M, const float* emb, int C, int H, TfWs& w,
                 void* st) {
  const int dh = C / H;
  GR_TRY(linear(x, N, C, L.wq, L.bq, C, w.q, 0, st));
  GR_TRY(linear(mem, M, C, L.wk, L.bk, C, w.k, 0, st));
  GR_TRY(linear(mem, M, C, L.wv, L.bv, C, w.v, 0, st));
  if (L.is_self) {
    // U[h] = q_h (N,dh) @ W_p[h*dh:(h+1)*dh, :] ; qb[h] = q_h @ b_p[h*dh:(h+1)*dh]   (see attention.cu)
    GR_TRY(gr_gemm(w.q, C, dh, L.wp, C, (int64_t)dh * C, 0, w.U, C, (int64_t)N * C, N, C, dh, H, 1.f, nullptr, nullptr, nullptr, 0,
                   0, 0, st));
    GR_TRY(gr_gemm(w.q, C, dh, L.bp, dh, dh, 1, w.qb, 1, N, N, 1, dh, H, 1.f, nullptr, nullptr, nullptr, 0, 0, 0, st));
    GR_TRY(gr_rpe_attention_probs(w.q, w.k, w.U, w.qb, emb, N, C, H, w.P, st));
  } else {
    GR_TRY(gr_gemm(w.q, C, dh, w.k, C, dh, 1, w.P, M, (int64_t)N * M, N, M, dh, H, 1.0f / sqrtf((float)dh), nullptr, nullptr,
                   nullptr, 0, 0, 0, st));
    GR_TRY(gr_softmax_rows(w.P, (int64_t)H * N, M, st));
  }
  // hidden[:, h*dh:(h+1)*dh] = P[h] @ v[:, h*dh:(h+1)*dh]
  GR_TRY(gr_gemm(w.P, M, (int64_t)N * M, w.v, C, dh, 0, w.hid, C, dh, N, dh, M, H, 1.f, nullptr, nullptr, nullptr, 0, 0, 0, st));
  GR_TRY(linear(w.hid, N, C, L.wo, L.bo, C, w.att, 0, st));
  GR_TRY(gr_layer_norm_add(w.att, x, N, C, L.ln1_g, L.ln1_b, 1e-5f, w.y, st));
  GR_TRY(linear(w.y, N, C, L.w1, L.b1, 2 * C, w.ffn, 1, st));
  GR_TRY(linear(w.ffn, N, 2 * C, L.w2, L.b2, C, w.att, 0, st));
  GR_TRY(gr_layer_norm_add(w.y, w.att, N, C, L.ln2_g, L.ln2_b, 1e-5f, x, st));
  return GR_OK;
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_conditional_transformer_workspace_size(int N0, int N1, int C, int num_heads) {
  bool ok;
  return carve_tf(nullptr, 0, N0 > N1 ? N0 : N1, C, num_heads, &ok).bytes;
}

/* feats0 (N0,C) / feats1 (N1,C) are updated in place through all layers; "self" layers use emb0 (N0,N0,C) and
 * emb1 (N1,N1,C), "cross" layers run sequentially (feats0 attends feats1, then feats1 attends the UPDATED feats0). */
extern "C" int gr_conditional_transformer(const gr_layer_weights* layers, int n_layers, float* feats0, float* feats1,
                                          const float* emb0, const float* emb1, int N0, int N1, int C, int num_heads,
                                          void* ws, size_t ws_bytes, void* stream) {
  if (!layers || n_layers <= 0 || !feats0 || !feats1 || N0 <= 0 || N1 <= 0 || C <= 0 || num_heads <= 0 || C % num_heads != 0)
    return GR_ERR_BAD_ARG;
  bool ok;
  TfWs w = carve_tf(ws, ws_bytes, N0 > N1 ? N0 : N1, C, num_heads, &ok);
  if (!ws || !ok) return GR_ERR_WORKSPACE;
  for (int i = 0; i < n_layers; ++i) {
    const gr_layer_weights& L = layers[i];
    if (L.is_self) {
      if (!emb0 || !emb1 || !L.wp || !L.bp) return GR_ERR_BAD_ARG;
      GR_TRY(layer(L, feats0, N0, feats0, N0, emb0, C, num_heads, w, stream));
      GR_TRY(layer(L, feats1, N1, feats1, N1, emb1, C, num_heads, w, stream));
    } else {
      GR_TRY(layer(L, feats0, N0, feats1, N1, nullptr, C, num_heads, w, stream));
      GR_TRY(layer(L, feats1, N1, feats0, N0, nullptr, C, num_heads, w, stream));
    }
  }
  return GR_OK;
}
