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
  float *qkv, *U, *qb, *P, *hid, *att, *ffn, *y;
};

static size_t carve_tf(void* ws, size_t ws_bytes, int N, int C, int H, TfWs* sets, int n_sets, bool* ok) {
  Carver c(ws, ws_bytes);
  for (int i = 0; i < n_sets; ++i) {
    TfWs w;
    w.qkv = c.take<float>((size_t)N * 3 * C);
    w.U = c.take<float>((size_t)H * N * C);
    w.qb = c.take<float>((size_t)H * N);
    w.P = c.take<float>((size_t)H * N * N);
    w.hid = c.take<float>((size_t)N * C);
    w.att = c.take<float>((size_t)N * C);
    w.ffn = c.take<float>((size_t)N * 2 * C);
    w.y = c.take<float>((size_t)N * C);
    if (sets) sets[i] = w;
  }
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

// x (N,C) attends mem (M,C) [emb (N,N,C) when self]; result overwrites x
static int layer(const gr_layer_weights& L, float* x, int N, const float* mem, int M, const float* emb, int C, int H, TfWs& w,
                 void* st) {
  const int dh = C / H;
  const float *q, *k, *v;
  int64_t ldq, ldk, ldv;
  if (L.wqkv && L.bqkv && L.is_self) {  // x == mem: one product for q|k|v
    GR_TRY(linear(x, N, C, L.wqkv, L.bqkv, 3 * C, w.qkv, 0, st));
    q = w.qkv; k = w.qkv + C; v = w.qkv + 2 * C;
    ldq = ldk = ldv = 3 * C;
  } else if (L.wqkv && L.bqkv) {        // q from x, k|v from mem
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
  if (L.is_self) {
    // U[h] = q_h (N,dh) @ W_p[h*dh:(h+1)*dh, :] ; qb[h] = q_h @ b_p[h*dh:(h+1)*dh]   (see attention.cu)
    GR_TRY(gr_gemm(q, ldq, dh, L.wp, C, (int64_t)dh * C, 0, w.U, C, (int64_t)N * C, N, C, dh, H, 1.f, nullptr, nullptr, nullptr, 0,
                   0, 0, st));
    GR_TRY(gr_gemm(q, ldq, dh, L.bp, dh, dh, 1, w.qb, 1, N, N, 1, dh, H, 1.f, nullptr, nullptr, nullptr, 0, 0, 0, st));
    GR_TRY(gr_rpe_attention_probs_ld(q, ldq, k, ldk, w.U, w.qb, emb, N, C, H, w.P, st));
  } else {
    GR_TRY(gr_gemm(q, ldq, dh, k, ldk, dh, 1, w.P, M, (int64_t)N * M, N, M, dh, H, 1.0f / sqrtf((float)dh), nullptr, nullptr,
                   nullptr, 0, 0, 0, st));
    GR_TRY(gr_softmax_rows(w.P, (int64_t)H * N, M, st));
  }
  // hidden[:, h*dh:(h+1)*dh] = P[h] @ v[:, h*dh:(h+1)*dh]
  GR_TRY(gr_gemm(w.P, M, (int64_t)N * M, v, ldv, dh, 0, w.hid, C, dh, N, dh, M, H, 1.f, nullptr, nullptr, nullptr, 0, 0, 0, st));
  GR_TRY(linear(w.hid, N, C, L.wo, L.bo, C, w.att, 0, st));
  GR_TRY(gr_layer_norm_add(w.att, x, N, C, L.ln1_g, L.ln1_b, 1e-5f, w.y, st));
  if (L.w1t && L.w2t && C == kMlpC && mlp_fused()) {
    transformer_mlp_kernel<<<(N + kMlpRows - 1) / kMlpRows, kMlpC, 0, static_cast<cudaStream_t>(st)>>>(
        w.y, N, L.w1t, L.b1, L.w2t, L.b2, L.ln2_g, L.ln2_b, 1e-5f, x);
    GR_CHECK_LAUNCH("transformer_mlp_kernel");
    return GR_OK;
  }
  GR_TRY(linear(w.y, N, C, L.w1, L.b1, 2 * C, w.ffn, 1, st));
  GR_TRY(linear(w.ffn, N, 2 * C, L.w2, L.b2, C, w.att, 0, st));
  GR_TRY(gr_layer_norm_add(w.y, w.att, N, C, L.ln2_g, L.ln2_b, 1e-5f, x, st));
  return GR_OK;
}

// A helper stream per (device, caller stream): the two clouds of a 'self' layer are independent, and one
// superpoint-sized kernel (a few dozen CTAs) leaves most of the 148 SMs idle, so they run side by side.
struct SideStream {
  int dev;
  cudaStream_t main, side;
  cudaEvent_t fork, join;
};

static SideStream* side_stream(cudaStream_t main) {
  static std::mutex mu;
  static std::vector<SideStream*> all;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (SideStream* s : all)
    if (s->dev == dev && s->main == main) return s;
  SideStream* s = new SideStream{dev, main, nullptr, nullptr, nullptr};
  if (cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming) != cudaSuccess) {
    set_last_error("transformer side stream", cudaGetLastError());
    delete s;
    return nullptr;
  }
  all.push_back(s);
  return s;
}

static bool two_streams() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GAUSSREG_TF_STREAMS"); v = e ? atoi(e) : 2; }
  return v >= 2;
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_conditional_transformer_workspace_size(int N0, int N1, int C, int num_heads) {
  bool ok;
  return carve_tf(nullptr, 0, N0 > N1 ? N0 : N1, C, num_heads, nullptr, 2, &ok);
}

/* feats0 (N0,C) / feats1 (N1,C) are updated in place through all layers; "self" layers use emb0 (N0,N0,C) and
 * emb1 (N1,N1,C), "cross" layers run sequentially (feats0 attends feats1, then feats1 attends the UPDATED feats0).
 * The two clouds of a self layer are issued on two streams (joined again before the next cross layer and before
 * returning: from the caller's point of view all work is ordered on `stream`). */
extern "C" int gr_conditional_transformer(const gr_layer_weights* layers, int n_layers, float* feats0, float* feats1,
                                          const float* emb0, const float* emb1, int N0, int N1, int C, int num_heads,
                                          void* ws, size_t ws_bytes, void* stream) {
  if (!layers || n_layers <= 0 || !feats0 || !feats1 || N0 <= 0 || N1 <= 0 || C <= 0 || num_heads <= 0 || C % num_heads != 0)
    return GR_ERR_BAD_ARG;
  bool ok;
  TfWs w[2];
  carve_tf(ws, ws_bytes, N0 > N1 ? N0 : N1, C, num_heads, w, 2, &ok);
  if (!ws || !ok) return GR_ERR_WORKSPACE;
  cudaStream_t main = static_cast<cudaStream_t>(stream);
  SideStream* ss = two_streams() ? side_stream(main) : nullptr;
  if (two_streams() && !ss) return GR_ERR_CUDA;
  for (int i = 0; i < n_layers; ++i) {
    const gr_layer_weights& L = layers[i];
    if (L.is_self) {
      if (!emb0 || !emb1 || !L.wp || !L.bp) return GR_ERR_BAD_ARG;
      if (ss) {
        GR_CHECK_CUDA(cudaEventRecord(ss->fork, main));
        GR_CHECK_CUDA(cudaStreamWaitEvent(ss->side, ss->fork, 0));
        const int rc0 = layer(L, feats0, N0, feats0, N0, emb0, C, num_heads, w[0], main);
        const int rc1 = layer(L, feats1, N1, feats1, N1, emb1, C, num_heads, w[1], ss->side);
        // always join, also on failure: the caller's stream must never run ahead of the helper
        cudaEventRecord(ss->join, ss->side);
        cudaStreamWaitEvent(main, ss->join, 0);
        if (rc0 != GR_OK) return rc0;
        if (rc1 != GR_OK) return rc1;
      } else {
        GR_TRY(layer(L, feats0, N0, feats0, N0, emb0, C, num_heads, w[0], stream));
        GR_TRY(layer(L, feats1, N1, feats1, N1, emb1, C, num_heads, w[0], stream));
      }
    } else {
      GR_TRY(layer(L, feats0, N0, feats1, N1, nullptr, C, num_heads, w[0], stream));
      GR_TRY(layer(L, feats1, N1, feats0, N0, nullptr, C, num_heads, w[0], stream));
    }
  }
  return GR_OK;
}
