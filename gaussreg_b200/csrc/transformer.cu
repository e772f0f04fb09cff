// T4: the RPE-conditional transformer (geotransformer/modules/transformer/conditional_transformer.py:97-117 with
// rpe_transformer.py:18-131, vanilla_transformer.py:15-129, output_layer.py:6-21) as ONE C-ABI call.
//
// Host-side orchestration only: every tensor op below is one of this library's kernels.  Running the ~130 launches
// of the six layers from C++ instead of from Python removes ~10 us of interpreter overhead per op, which is more
// than most of these superpoint-sized kernels take on a B200.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

extern "C" {
int gr_gemm(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB, int trans_b,
            float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch, float alpha, const float* bias,
            const float* row_div, const float* residual, int64_t ldr, int64_t strideR, int act, void* stream);
int gr_layer_norm_add(const float* a, const float* b, int64_t rows, int C, const float* gamma, const float* beta,
                      float eps, float* y, void* stream);
int gr_rpe_attention_probs_ld(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* U, const float* qb,
                              const float* emb, int N, int C, int num_heads, float* P, void* stream);
int gr_softmax_rows(float* x, int64_t rows, int cols, void* stream);
}

namespace gr {
int rpe_attention_probs_ex(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* U, int64_t u_head, const float* qb,
                           const float* bp, const float* emb, int N, float* P, int64_t ldp, void* stream);
int cross_attention(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, int N, int M, float* out,
                    int64_t ldo, void* stream);
bool cross_attention_fits(int M);
}

namespace gr {

// ---- fused AttentionOutput (output_layer.py:14-21): y = LayerNorm(x + squeeze(relu(expand(x)))) -----------------
// Three superpoint-sized launches (two GEMMs of ~0.1 GFLOP and a LayerNorm) are pure latency; here a CTA owns 8 rows,
// keeps x, the 512 hidden units and the result in shared memory, streams the two K-major (pre-transposed) weight
// matrices out of L2 with coalesced loads, and finishes with the LayerNorm of layernorm_add_kernel (same lane
// mapping and reduction order).  C = 256 only.
constexpr int kMlpRows = 8;
constexpr int kMlpC = 256;

__global__ void __launch_bounds__(kMlpC) transformer_mlp_kernel(const float* __restrict__ x, int N, const float* __restrict__ w1t,
                                                               const float* __restrict__ b1, const float* __restrict__ w2t,
                                                               const float* __restrict__ b2, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float xs[kMlpRows][kMlpC];
  __shared__ __align__(16) float hs[kMlpRows][2 * kMlpC];
  __shared__ __align__(16) float ys[kMlpRows][kMlpC];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * kMlpRows;
#pragma unroll
  for (int r = 0; r < kMlpRows; ++r) xs[r][tid] = (row0 + r < N) ? x[(long long)(row0 + r) * kMlpC + tid] : 0.f;
  __syncthreads();
  {  // expand + ReLU: hidden units tid and tid + 256
    float a0[kMlpRows], a1[kMlpRows];
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) { a0[r] = 0.f; a1[r] = 0.f; }
#pragma unroll 2
    for (int k = 0; k < kMlpC; k += 4) {
      float w0[4], w1[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        w0[kk] = __ldg(w1t + (long long)(k + kk) * (2 * kMlpC) + tid);
        w1[kk] = __ldg(w1t + (long long)(k + kk) * (2 * kMlpC) + kMlpC + tid);
      }
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[r][k]);
        a0[r] = fmaf(xv.x, w0[0], a0[r]); a1[r] = fmaf(xv.x, w1[0], a1[r]);
        a0[r] = fmaf(xv.y, w0[1], a0[r]); a1[r] = fmaf(xv.y, w1[1], a1[r]);
        a0[r] = fmaf(xv.z, w0[2], a0[r]); a1[r] = fmaf(xv.z, w1[2], a1[r]);
        a0[r] = fmaf(xv.w, w0[3], a0[r]); a1[r] = fmaf(xv.w, w1[3], a1[r]);
      }
    }
    const float bb0 = b1[tid], bb1 = b1[kMlpC + tid];
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      hs[r][tid] = fmaxf(a0[r] + bb0, 0.f);
      hs[r][kMlpC + tid] = fmaxf(a1[r] + bb1, 0.f);
    }
  }
  __syncthreads();
  {  // squeeze: output channel tid
    float a[kMlpRows];
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) a[r] = 0.f;
#pragma unroll 2
    for (int k = 0; k < 2 * kMlpC; k += 4) {
      float w[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w[kk] = __ldg(w2t + (long long)(k + kk) * kMlpC + tid);
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(&hs[r][k]);
        a[r] = fmaf(hv.x, w[0], a[r]);
        a[r] = fmaf(hv.y, w[1], a[r]);
        a[r] = fmaf(hv.z, w[2], a[r]);
        a[r] = fmaf(hv.w, w[3], a[r]);
      }
    }
    const float bb = b2[tid];
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) ys[r][tid] = a[r] + bb;
  }
  __syncthreads();
  // LayerNorm(x + h): warp w owns row w, lane mapping and reduction order of layernorm_add_kernel
  const int row = row0 + warp;
  if (row >= N) return;
  float v[kMlpC / 32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMlpC / 32; ++i) {
    v[i] = xs[warp][lane + 32 * i] + ys[warp][lane + 32 * i];
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)kMlpC;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMlpC / 32; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)kMlpC + eps);
#pragma unroll
  for (int i = 0; i < kMlpC / 32; ++i) {
    const int c = lane + 32 * i;
    y[(long long)row * kMlpC + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

static bool mlp_fused() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_TF_MLP"); v = e ? atoi(e) : 0; }
  return v != 0;
}

struct TfWs {
  float *x, *qkv, *U, *qb, *P, *hid, *att, *ffn, *y;
};

// N = rows of the larger cloud, Nt = N0 + N1 (row-wise buffers hold the stacked clouds)
static size_t carve_tf(void* ws, size_t ws_bytes, int N, int Nt, int C, int H, TfWs* out, bool* ok) {
  Carver c(ws, ws_bytes);
  TfWs w;
  w.x = c.take<float>((size_t)Nt * C);
  w.qkv = c.take<float>((size_t)Nt * 3 * C);
  w.U = c.take<float>((size_t)H * Nt * C);
  w.qb = c.take<float>((size_t)H * N);
  w.P = c.take<float>((size_t)H * N * ((N + 3) & ~3));
  w.hid = c.take<float>((size_t)Nt * C);
  w.att = c.take<float>((size_t)Nt * C);
  w.ffn = c.take<float>((size_t)Nt * 2 * C);
  w.y = c.take<float>((size_t)Nt * C);
  if (out) *out = w;
  *ok = c.ok;
  return c.off;
}

#define GR_TRY(expr)                 \
  do {                               \
    const int rc__ = (expr);         \
    if (rc__ != GR_OK) return rc__;  \
  } while (0)

static int linear(const float* x, int rows, int in, const float* W, const float* b, int out, float* y, int act, void* st) {
  return gr_gemm(x, in, 0, W, in, 0, 1, y, out, 0, rows, out, in, 1, 1.f, b, nullptr, nullptr, 0, 0, act, st);
}

// out-projection + LayerNorm + AttentionOutput on `rows` rows (rpe_transformer.py:101-103, output_layer.py:14-21):
// x <- LN2(y + FFN(y)),  y = LN1(x + W_o hid + b_o)
static int post_attention(const gr_layer_weights& L, float* x, const float* hid, int rows, int C, TfWs& w, void* st) {
  // the residual x rides in the product's epilogue, LayerNorm then reads one operand
  GR_TRY(gr_gemm(hid, C, 0, L.wo, C, 0, 1, w.att, C, 0, rows, C, C, 1, 1.f, L.bo, nullptr, x, C, 0, 0, st));
  GR_TRY(gr_layer_norm_add(w.att, nullptr, rows, C, L.ln1_g, L.ln1_b, 1e-5f, w.y, st));
  if (L.w1t && L.w2t && C == kMlpC && mlp_fused()) {
    GR_CHECK_CUDA(launch_pdl(transformer_mlp_kernel, dim3((rows + kMlpRows - 1) / kMlpRows), dim3(kMlpC), (size_t)(0), static_cast<cudaStream_t>(st), w.y, rows, L.w1t, L.b1, L.w2t, L.b2, L.ln2_g, L.ln2_b, 1e-5f, x));
    GR_CHECK_LAUNCH("transformer_mlp_kernel");
    return GR_OK;
  }
  GR_TRY(linear(w.y, rows, C, L.w1, L.b1, 2 * C, w.ffn, 1, st));
  GR_TRY(gr_gemm(w.ffn, 2 * C, 0, L.w2, 2 * C, 0, 1, w.att, C, 0, rows, C, 2 * C, 1, 1.f, L.b2, nullptr, w.y, C, 0, 0, st));
  GR_TRY(gr_layer_norm_add(w.att, nullptr, rows, C, L.ln2_g, L.ln2_b, 1e-5f, x, st));
  return GR_OK;
}

static bool tf_fused() {  // 0: round-2a sequence (per-cloud U / qb products, three-launch cross-attention)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_TF_FUSED"); v = e ? atoi(e) : 1; }
  return v != 0;
}

// RPE self-attention of BOTH clouds (rows [0,N0) and [N0,N0+N1) of the stacked x): every row-wise product (q|k|v,
// U = q_h W_p,h, out-projection, FFN) and both LayerNorms run once on the N0+N1 stacked rows; only the N x N attention
// itself is per cloud.  The q.b_p term is computed inside the score kernel.
static int self_layer(const gr_layer_weights& L, float* x, int N0, int N1, const float* emb0, const float* emb1, int C, int H,
                      TfWs& w, void* st) {
  const int dh = C / H, Nt = N0 + N1;
  if (!(L.wqkv && L.bqkv)) return GR_ERR_BAD_ARG;
  GR_TRY(linear(x, Nt, C, L.wqkv, L.bqkv, 3 * C, w.qkv, 0, st));
  const int64_t ld = 3 * C;
  const bool fused = tf_fused() && C == 256 && H == 4 && (size_t)H * (N0 > N1 ? N0 : N1) * sizeof(float) <= 100 * 1024;
  if (fused)  // U[h] (Nt, C) = q_h (Nt, dh) @ W_p[h*dh:(h+1)*dh, :] for both clouds at once   (see attention.cu)
    GR_TRY(gr_gemm(w.qkv, ld, dh, L.wp, C, (int64_t)dh * C, 0, w.U, C, (int64_t)Nt * C, Nt, C, dh, H, 1.f, nullptr, nullptr, nullptr, 0, 0,
                   0, st));
  for (int c = 0; c < 2; ++c) {
    const int N = c == 0 ? N0 : N1, off = c == 0 ? 0 : N0;
    const float* q = w.qkv + (size_t)off * ld;
    const float *k = q + C, *v = q + 2 * C;
    const float* emb = c == 0 ? emb0 : emb1;
    const int64_t ldp = fused ? ((N + 3) & ~3) : N;
    if (fused) {
      GR_TRY(rpe_attention_probs_ex(q, ld, k, ld, w.U + (size_t)off * C, (int64_t)Nt * C, nullptr, L.bp, emb, N, w.P, ldp, st));
    } else {
      GR_TRY(gr_gemm(q, ld, dh, L.wp, C, (int64_t)dh * C, 0, w.U, C, (int64_t)N * C, N, C, dh, H, 1.f, nullptr, nullptr, nullptr, 0, 0,
                     0, st));
      GR_TRY(gr_gemm(q, ld, dh, L.bp, dh, dh, 1, w.qb, 1, N, N, 1, dh, H, 1.f, nullptr, nullptr, nullptr, 0, 0, 0, st));
      GR_TRY(gr_rpe_attention_probs_ld(q, ld, k, ld, w.U, w.qb, emb, N, C, H, w.P, st));
    }
    GR_TRY(gr_gemm(w.P, ldp, (int64_t)N * ldp, v, ld, dh, 0, w.hid + (size_t)off * C, C, dh, N, dh, N, H, 1.f, nullptr, nullptr, nullptr,
                   0, 0, 0, st));
  }
  return post_attention(L, x, w.hid, Nt, C, w, st);
}

// vanilla cross-attention: x (N rows) attends mem (M rows); result overwrites x
static int cross_layer(const gr_layer_weights& L, float* x, int N, const float* mem, int M, int C, int H, TfWs& w, void* st) {
  const int dh = C / H;
  const float *q, *k, *v;
  int64_t ldq, ldk, ldv;
  if (L.wqkv && L.bqkv) {  // q from x, k|v from mem in one product
    float* kv = w.qkv + (size_t)N * C;
    GR_TRY(linear(x, N, C, L.wqkv, L.bqkv, C, w.qkv, 0, st));
    GR_TRY(linear(mem, M, C, L.wqkv + (size_t)C * C, L.bqkv + C, 2 * C, kv, 0, st));
    q = w.qkv; ldq = C;
    k = kv; v = kv + C; ldk = ldv = 2 * C;
  } else {
    float* kb = w.qkv + (size_t)N * C;
    float* vb = kb + (size_t)M * C;
    GR_TRY(linear(x, N, C, L.wq, L.bq, C, w.qkv, 0, st));
    GR_TRY(linear(mem, M, C, L.wk, L.bk, C, kb, 0, st));
    GR_TRY(linear(mem, M, C, L.wv, L.bv, C, vb, 0, st));
    q = w.qkv; k = kb; v = vb;
    ldq = ldk = ldv = C;
  }
  GR_TRY(gr_gemm(q, ldq, dh, k, ldk, dh, 1, w.P, M, (int64_t)N * M, N, M, dh, H, 1.0f / sqrtf((float)dh), nullptr, nullptr, nullptr, 0,
                 0, 0, st));
  GR_TRY(gr_softmax_rows(w.P, (int64_t)H * N, M, st));
  GR_TRY(gr_gemm(w.P, M, (int64_t)N * M, v, ldv, dh, 0, w.hid, C, dh, N, dh, M, H, 1.f, nullptr, nullptr, nullptr, 0, 0, 0, st));
  return post_attention(L, x, w.hid, N, C, w, st);
}

// One cross layer of conditional_transformer.py:107-112 (cloud 0 attends cloud 1, then cloud 1 attends the UPDATED
// cloud 0) on the stacked x.  Everything the pre-update state determines -- q of both clouds, k|v of cloud 1 -- comes
// out of ONE product with the packed (3C, C) weight; only k|v of the updated cloud 0 needs a second one.  Each
// attention is one fused kernel.
static int cross_pair(const gr_layer_weights& L, float* x, int N0, int N1, int C, TfWs& w, void* st) {
  const int Nt = N0 + N1;
  const int64_t ld = 3 * C;
  float* x0 = x;
  float* x1 = x + (size_t)N0 * C;
  float* qkv0 = w.qkv;
  float* qkv1 = w.qkv + (size_t)N0 * ld;
  GR_TRY(linear(x, Nt, C, L.wqkv, L.bqkv, 3 * C, w.qkv, 0, st));
  GR_TRY(cross_attention(qkv0, ld, qkv1 + C, ld, qkv1 + 2 * C, ld, N0, N1, w.hid, C, st));
  GR_TRY(post_attention(L, x0, w.hid, N0, C, w, st));
  GR_TRY(gr_gemm(x0, C, 0, L.wqkv + (size_t)C * C, C, 0, 1, qkv0 + C, ld, 0, N0, 2 * C, C, 1, 1.f, L.bqkv + C, nullptr, nullptr, 0, 0, 0,
                 st));
  GR_TRY(cross_attention(qkv1, ld, qkv0 + C, ld, qkv0 + 2 * C, ld, N1, N0, w.hid + (size_t)N0 * C, C, st));
  return post_attention(L, x1, w.hid + (size_t)N0 * C, N1, C, w, st);
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_conditional_transformer_workspace_size(int N0, int N1, int C, int num_heads) {
  bool ok;
  return carve_tf(nullptr, 0, N0 > N1 ? N0 : N1, N0 + N1, C, num_heads, nullptr, &ok);
}

/* feats0 (N0,C) / feats1 (N1,C) are updated in place through all layers; "self" layers use emb0 (N0,N0,C) and
 * emb1 (N1,N1,C), "cross" layers run sequentially (feats0 attends feats1, then feats1 attends the UPDATED feats0).
 * Internally the two clouds are stacked into one (N0+N1, C) matrix: the row-wise products and LayerNorms of a self
 * layer then run once for both clouds. */
extern "C" int gr_conditional_transformer(const gr_layer_weights* layers, int n_layers, float* feats0, float* feats1,
                                          const float* emb0, const float* emb1, int N0, int N1, int C, int num_heads,
                                          void* ws, size_t ws_bytes, void* stream) {
  if (!layers || n_layers <= 0 || !feats0 || !feats1 || N0 <= 0 || N1 <= 0 || C <= 0 || num_heads <= 0 || C % num_heads != 0)
    return GR_ERR_BAD_ARG;
  bool ok;
  TfWs w;
  carve_tf(ws, ws_bytes, N0 > N1 ? N0 : N1, N0 + N1, C, num_heads, &w, &ok);
  if (!ws || !ok) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // feats1 directly behind feats0: the caller's buffer IS the stacked matrix, no staging copies
  const bool in_place = feats1 == feats0 + (size_t)N0 * C;
  float* X = in_place ? feats0 : w.x;
  float* x0 = X;
  float* x1 = X + (size_t)N0 * C;
  if (!in_place) {
    GR_CHECK_CUDA(cudaMemcpyAsync(x0, feats0, (size_t)N0 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    GR_CHECK_CUDA(cudaMemcpyAsync(x1, feats1, (size_t)N1 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  for (int i = 0; i < n_layers; ++i) {
    const gr_layer_weights& L = layers[i];
    if (L.is_self) {
      if (!emb0 || !emb1 || !L.wp || !L.bp) return GR_ERR_BAD_ARG;
      GR_TRY(self_layer(L, X, N0, N1, emb0, emb1, C, num_heads, w, stream));
    } else {
      if (tf_fused() && C == 256 && num_heads == 4 && L.wqkv && L.bqkv && cross_attention_fits(N0 > N1 ? N0 : N1)) {
        GR_TRY(cross_pair(L, X, N0, N1, C, w, stream));
      } else {
        GR_TRY(cross_layer(L, x0, N0, x1, N1, C, num_heads, w, stream));
        GR_TRY(cross_layer(L, x1, N1, x0, N0, C, num_heads, w, stream));
      }
    }
  }
  if (!in_place) {
    GR_CHECK_CUDA(cudaMemcpyAsync(feats0, x0, (size_t)N0 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    GR_CHECK_CUDA(cudaMemcpyAsync(feats1, x1, (size_t)N1 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return GR_OK;
}
