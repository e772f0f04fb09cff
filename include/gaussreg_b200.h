/*
 * gaussreg_b200 -- C ABI of the B200-native coarse-registration forward path.
 *
 * This is the drop-in boundary.  Conventions, all entry points:
 *   - every data pointer is a DEVICE pointer owned by the caller (PyTorch allocates), unless the
 *     parameter name starts with `h_` (host);
 *   - `ws` / `ws_bytes` is caller-owned scratch, size from the matching *_workspace_size();
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises the stream or the
 *     device, data-dependent sizes are returned through device scalars;
 *   - return value: 0 on success, negative gr_status on error (the Python shim raises RuntimeError);
 *   - stacked ("stack mode") layout as in the reference: clouds of a batch are concatenated along
 *     dim 0 and described by an int64 `lengths[batch]` array.
 *   - there is NO CPU fallback: every function launches sm_100a kernels.
 */
#ifndef GAUSSREG_B200_H_
#define GAUSSREG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  GR_OK = 0,
  GR_ERR_BAD_ARG = -1,
  GR_ERR_WORKSPACE = -2,
  GR_ERR_CAPACITY = -3,
  GR_ERR_CUDA = -4,
} gr_status;

/* Version / build info string (static storage). */
const char* gr_version(void);
/* Text of the last CUDA error seen by this library on the calling thread. */
const char* gr_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t gr_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * G1  grid subsampling.
 * Replaces geotransformer.ext.grid_subsampling
 *   (geotransformer/extensions/pybind.cpp:13-17 -> cpu/grid_subsampling/grid_subsampling.cpp:5-62,
 *    grid_subsampling_cpu.cpp:3-75).
 * Voxel barycentres per cloud, bit-identical values AND emission order (libstdc++
 * unordered_map iteration order is replayed on the device).
 *
 *   points      (n_points,3) f32, stacked; only the first sum(lengths) rows are read
 *   lengths     (batch) i64, DEVICE; sum(lengths) <= n_points
 *   out_points  capacity (n_points,3) f32; the first *out_total rows are written
 *   out_lengths (batch) i64, DEVICE
 *   out_total   DEVICE i64 scalar = sum(out_lengths)
 * --------------------------------------------------------------------------------------------- */
size_t gr_grid_subsample_workspace_size(int64_t n_points, int batch);
int gr_grid_subsample(const float* points, const int64_t* lengths, int batch, int64_t n_points, float voxel_size,
                      float* out_points, int64_t* out_lengths, int64_t* out_total, void* ws, size_t ws_bytes,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * G2  fixed-radius neighbour search.
 * Replaces geotransformer.ext.radius_neighbors
 *   (pybind.cpp:8-12 -> cpu/radius_neighbors/radius_neighbors.cpp:5-68, radius_neighbors_cpu.cpp:3-91)
 * and the column truncation of geotransformer/modules/ops/radius_search.py:24-27.
 *
 * For every query row: all support points of the same batch element with
 *   ((qx-sx)^2 + (qy-sy)^2) + (qz-sz)^2 < radius*radius   (f32, no FMA, strict)
 * ascending by that distance, ties by ascending index (the reference's std::sort leaves ties in
 * unspecified order), indices offset to the stacked support array, rows padded with sum(s_lengths).
 *
 *   q_points (nq,3), s_points (ns,3) f32; q_lengths/s_lengths (batch) i64 DEVICE,
 *            sum(q_lengths) <= nq, sum(s_lengths) <= ns
 *   out_idx  (nq, ld) i64 row-major, or NULL to only count.  Each row receives its first
 *            min(count, ld) neighbours, the rest of the row is padding.
 *   out_max_count DEVICE i32 scalar: max over rows of the untruncated neighbour count.  The
 *            reference's result is out_idx[:, :min(max_count, limit)].
 * --------------------------------------------------------------------------------------------- */
size_t gr_radius_neighbors_workspace_size(int64_t nq, int64_t ns, int batch);
int gr_radius_neighbors(const float* q_points, const float* s_points, const int64_t* q_lengths,
                        const int64_t* s_lengths, int batch, int64_t nq, int64_t ns, float radius, int64_t* out_idx,
                        int64_t ld, int32_t* out_max_count, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAUSSREG_B200_H_ */
