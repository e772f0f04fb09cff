// N4: apply a similarity transform to a 3DGS Gaussian cloud (reference: gs_fusion.py:231-262 `gaussian_fuse`, the part that
// moves the second cloud: positions :240, log-scales :241, rotation quaternions :242-243 via quaternion_to_matrix :70-98 and
// matrix_to_quaternion :111-159, spherical-harmonic bands 1-3 :244 via sh_rotation :53-68).
//
// One pass over the (N,59) rows: 236 bytes read and written per Gaussian, nothing else.  The three SH band matrices
// (3x3, 5x5, 7x7; identical for every Gaussian and channel -- the reference rebuilds them N*3 times with batched pinv)
// are computed once on the host and passed by value; they are applied in double precision like the reference's float64
// arrays, everything else in float32 with the reference's operation order.
#include "common.cuh"

namespace gr {

struct FuseParams {
  double m1[9], m2[25], m3[49];  // rotated band = band (row vector) @ m
  float R[9];                    // unit rotation, row-major
  float t[3];
  float scale, log_scale;
  int apply_scale;               // gs_fusion.py:241: `if scale != 1.`
};

__global__ void __launch_bounds__(128) gaussian_transform_kernel(const float* __restrict__ in, long long ld_in, long long n, FuseParams p,
                                                                 float* __restrict__ out, long long ld_out) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = in + i * ld_in;
  float* o = out + i * ld_out;
  // xyz @ R^T * scale + t
  const float x = a[0], y = a[1], z = a[2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float d = fmaf(z, p.R[3 * r + 2], fmaf(y, p.R[3 * r + 1], __fmul_rn(x, p.R[3 * r])));
    o[r] = __fadd_rn(__fmul_rn(d, p.scale), p.t[r]);
  }
  // f_dc, opacity unchanged
  o[3] = a[3]; o[4] = a[4]; o[5] = a[5];
  o[51] = a[51];
  // SH bands: (3 channels) x (3 | 5 | 7) coefficients, channel-major f_rest layout (load_ply reshapes to (N,3,15))
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* s = a + 6 + 15 * c;
    float* d = o + 6 + 15 * c;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) v += (double)s[k] * p.m1[3 * k + j];
      d[j] = (float)v;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 5; ++k) v += (double)s[3 + k] * p.m2[5 * k + j];
      d[3 + j] = (float)v;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 7; ++k) v += (double)s[8 + k] * p.m3[7 * k + j];
      d[8 + j] = (float)v;
    }
  }
  // log-scales
#pragma unroll
  for (int k = 0; k < 3; ++k) o[52 + k] = p.apply_scale ? (float)((double)a[52 + k] + (double)p.log_scale) : a[52 + k];
  // rotation: q' = matrix_to_quaternion(R @ quaternion_to_matrix(q)), real part first
  const float qr = a[55], qi = a[56], qj = a[57], qk = a[58];
  const float two_s = __fdiv_rn(2.0f, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qr, qr), __fmul_rn(qi, qi)), __fmul_rn(qj, qj)), __fmul_rn(qk, qk)));
  float Q[9];
  Q[0] = 1.f - two_s * (qj * qj + qk * qk); Q[1] = two_s * (qi * qj - qk * qr);       Q[2] = two_s * (qi * qk + qj * qr);
  Q[3] = two_s * (qi * qj + qk * qr);       Q[4] = 1.f - two_s * (qi * qi + qk * qk); Q[5] = two_s * (qj * qk - qi * qr);
  Q[6] = two_s * (qi * qk - qj * qr);       Q[7] = two_s * (qj * qk + qi * qr);       Q[8] = 1.f - two_s * (qi * qi + qj * qj);
  float M[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) M[3 * r + c] = fmaf(p.R[3 * r + 2], Q[6 + c], fmaf(p.R[3 * r + 1], Q[3 + c], p.R[3 * r] * Q[c]));
  const float m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[3], m11 = M[4], m12 = M[5], m20 = M[6], m21 = M[7], m22 = M[8];
  float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
#pragma unroll
  for (int k = 0; k < 4; ++k) qa[k] = qa[k] > 0.f ? sqrtf(qa[k]) : 0.f;
  int best = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (qa[k] > qa[best]) best = k;  // torch.argmax: first maximum
  float cand[4];
  if (best == 0) { cand[0] = qa[0] * qa[0]; cand[1] = m21 - m12; cand[2] = m02 - m20; cand[3] = m10 - m01; }
  else if (best == 1) { cand[0] = m21 - m12; cand[1] = qa[1] * qa[1]; cand[2] = m10 + m01; cand[3] = m02 + m20; }
  else if (best == 2) { cand[0] = m02 - m20; cand[1] = m10 + m01; cand[2] = qa[2] * qa[2]; cand[3] = m12 + m21; }
  else { cand[0] = m10 - m01; cand[1] = m20 + m02; cand[2] = m21 + m12; cand[3] = qa[3] * qa[3]; }
  const float den = 2.0f * fmaxf(qa[best], 0.1f);
#pragma unroll
  for (int k = 0; k < 4; ++k) o[55 + k] = __fdiv_rn(cand[k], den);
}

}  // namespace gr

using namespace gr;

/* cloud (n,59) with row pitch ld_in -> out (n,59) with pitch ld_out (may alias cloud when the pitches agree).
 * h_rotation: host 3x3 unit rotation (row-major), scale, log_scale = log(scale) as the caller's float32 value, h_translation[3],
 * h_sh: host doubles 9 + 25 + 49 = the band matrices M_1, M_2, M_3 (rotated coefficients = coefficients @ M_l). */
extern "C" int gr_gaussian_transform(const float* cloud, int64_t ld_in, int64_t n, const float* h_rotation, float scale, float log_scale,
                                     const float* h_translation, const double* h_sh, float* out, int64_t ld_out, void* stream) {
  if (n < 0 || ld_in < 59 || ld_out < 59 || !(scale > 0.f)) return GR_ERR_BAD_ARG;
  if (n == 0) return GR_OK;
  if (!cloud || !out || !h_rotation || !h_translation || !h_sh) return GR_ERR_BAD_ARG;
  FuseParams p;
  for (int i = 0; i < 9; ++i) { p.R[i] = h_rotation[i]; p.m1[i] = h_sh[i]; }
  for (int i = 0; i < 25; ++i) p.m2[i] = h_sh[9 + i];
  for (int i = 0; i < 49; ++i) p.m3[i] = h_sh[34 + i];
  for (int i = 0; i < 3; ++i) p.t[i] = h_translation[i];
  p.scale = scale; p.log_scale = log_scale; p.apply_scale = scale != 1.0f;
  GR_CHECK_CUDA(launch_pdl(gaussian_transform_kernel, dim3(ceil_div(n, 128)), dim3(128), (size_t)(0), static_cast<cudaStream_t>(stream), cloud, ld_in, n, p, out, ld_out));
  GR_CHECK_LAUNCH("gaussian_transform_kernel");
  return GR_OK;
}
