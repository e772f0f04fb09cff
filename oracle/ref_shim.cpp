// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin extern "C" wrapper around the *unmodified* reference CPU sources, compiled
// in place from /root/reference by oracle/Makefile into oracle/_ref/libgaussreg_ref.so.
// It replaces the torch/pybind layer of the reference
// (geotransformer/extensions/cpu/grid_subsampling/grid_subsampling.cpp:5-62,
//  geotransformer/extensions/cpu/radius_neighbors/radius_neighbors.cpp:5-68) with
// plain pointers so that the reference algorithm can be called without torch
// headers, from ctypes, on any box the .so travels to.
#include <cstdint>
#include <cstring>
#include <vector>

#include "cpu/grid_subsampling/grid_subsampling_cpu.h"
#include "cpu/radius_neighbors/radius_neighbors_cpu.h"

extern "C" {

// returns total number of subsampled points; out_points must hold n_points*3 floats.
int64_t ref_grid_subsampling(const float* points, const int64_t* lengths, int batch,
                             int64_t n_points, float voxel, float* out_points,
                             int64_t* out_lengths) {
  std::vector<PointXYZ> pts(reinterpret_cast<const PointXYZ*>(points),
                            reinterpret_cast<const PointXYZ*>(points) + n_points);
  std::vector<long> len(lengths, lengths + batch);
  std::vector<PointXYZ> s_pts;
  std::vector<long> s_len;
  grid_subsampling_cpu(pts, s_pts, len, s_len, voxel);
  std::memcpy(out_points, s_pts.data(), sizeof(float) * 3 * s_pts.size());
  for (int b = 0; b < batch; ++b) out_lengths[b] = s_len[b];
  return static_cast<int64_t>(s_pts.size());
}

// Two-call protocol: first call with out_idx == nullptr returns max_count;
// the state is kept in a thread-local so the (expensive) search runs once.
static thread_local std::vector<long> g_last;

int64_t ref_radius_neighbors(const float* q, const float* s, const int64_t* q_len,
                             const int64_t* s_len, int batch, int64_t nq, int64_t ns,
                             float radius, int64_t* out_idx) {
  if (out_idx == nullptr) {
    std::vector<PointXYZ> vq(reinterpret_cast<const PointXYZ*>(q),
                             reinterpret_cast<const PointXYZ*>(q) + nq);
    std::vector<PointXYZ> vs(reinterpret_cast<const PointXYZ*>(s),
                             reinterpret_cast<const PointXYZ*>(s) + ns);
    std::vector<long> ql(q_len, q_len + batch), sl(s_len, s_len + batch);
    g_last.clear();
    radius_neighbors_cpu(vq, vs, ql, sl, g_last, radius);
    return nq ? static_cast<int64_t>(g_last.size() / nq) : 0;
  }
  std::memcpy(out_idx, g_last.data(), sizeof(long) * g_last.size());
  int64_t w = nq ? static_cast<int64_t>(g_last.size() / nq) : 0;
  g_last.clear();
  g_last.shrink_to_fit();
  return w;
}

}  // extern "C"
