// Neighbour pyramid kernels for sm_100a.
//
//   G2  gr_radius_neighbors : uniform cell binning of the support cloud (counting sort into
//       cell-major float4 records), one warp per query scanning the 3x3 runs of x-contiguous
//       cells, warp-ballot compaction of the hits into shared memory, rank-by-counting sort on the
//       packed (distance bits, index) key, coalesced row write.
//   G1  gr_grid_subsample   : 64-bit voxel keys -> open-addressing hash table (first occurrence,
//       count) -> voxels compacted in first-occurrence order -> per-voxel point lists summed in
//       input order (the reference's sequential float +=) -> replay of the libstdc++
//       unordered_map iteration order (one CTA per cloud, small rehash phases in shared memory).
//
// All float arithmetic that decides an index uses explicit round-to-nearest intrinsics so that
// nvcc cannot contract a*b+c into an FMA: the reference is x86-64 code without FMA.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace gr {

// =============================================================================================
// shared small kernels
// =============================================================================================

struct BBox {
  unsigned int mn[3];
  unsigned int mx[3];
};

// lengths (i64) -> int32 offsets off[0..batch]; also resets the per-cloud bounding boxes.
__global__ void prep_offsets_kernel(const int64_t* __restrict__ len_a, int* __restrict__ off_a,
                                    const int64_t* __restrict__ len_b, int* __restrict__ off_b, int batch,
                                    BBox* __restrict__ bbox, int* __restrict__ scalars, int n_scalars) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    off_a[0] = 0;
    for (int b = 0; b < batch; ++b) { acc += len_a[b]; off_a[b + 1] = static_cast<int>(acc); }
    if (len_b != nullptr) {
      acc = 0;
      off_b[0] = 0;
      for (int b = 0; b < batch; ++b) { acc += len_b[b]; off_b[b + 1] = static_cast<int>(acc); }
    }
  }
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    for (int a = 0; a < 3; ++a) {
      bbox[b].mn[a] = 0xffffffffu;
      bbox[b].mx[a] = 0u;
    }
  }
  for (int i = threadIdx.x; i < n_scalars; i += blockDim.x) scalars[i] = 0;
}

// per-cloud min / max corner (strict comparisons in the reference == plain min/max here)
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ pts, const int* __restrict__ off, int batch,
                                                   BBox* __restrict__ bbox) {
  pdl_wait();
  pdl_trigger();
  const int n = off[batch];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  int b = 0;
  float x = 0.f, y = 0.f, z = 0.f;
  if (valid) {
    b = find_segment(off, batch, i);
    x = pts[3 * i]; y = pts[3 * i + 1]; z = pts[3 * i + 2];
  }
  // warp-uniform cloud -> shuffle reduce, else per-lane atomics
  const unsigned int act = __ballot_sync(0xffffffffu, valid);
  if (act == 0u) return;
  const int leader = __ffs(act) - 1;
  const int b0 = __shfl_sync(0xffffffffu, b, leader);
  const bool uniform = __all_sync(0xffffffffu, !valid || b == b0);
  if (uniform) {
    const float inf = __int_as_float(0x7f800000);
    float mnx = warp_min(valid ? x : inf), mny = warp_min(valid ? y : inf), mnz = warp_min(valid ? z : inf);
    float mxx = warp_max(valid ? x : -inf), mxy = warp_max(valid ? y : -inf), mxz = warp_max(valid ? z : -inf);
    if (lane_id() == 0) {
      atomicMin(&bbox[b0].mn[0], f2ord(mnx)); atomicMin(&bbox[b0].mn[1], f2ord(mny)); atomicMin(&bbox[b0].mn[2], f2ord(mnz));
      atomicMax(&bbox[b0].mx[0], f2ord(mxx)); atomicMax(&bbox[b0].mx[1], f2ord(mxy)); atomicMax(&bbox[b0].mx[2], f2ord(mxz));
    }
  } else if (valid) {
    atomicMin(&bbox[b].mn[0], f2ord(x)); atomicMin(&bbox[b].mn[1], f2ord(y)); atomicMin(&bbox[b].mn[2], f2ord(z));
    atomicMax(&bbox[b].mx[0], f2ord(x)); atomicMax(&bbox[b].mx[1], f2ord(y)); atomicMax(&bbox[b].mx[2], f2ord(z));
  }
}

// =============================================================================================
// G2  radius neighbours
// =============================================================================================

constexpr int kCellsPerCloud = 1 << 18;  // dense cell table capacity per cloud (coarsened beyond)
constexpr int kHitCap = 256;             // per-warp shared-memory hit buffer
constexpr int kSearchWarps = 8;

struct CellGrid {
  float ox, oy, oz;  // origin (support min corner)
  float inv;         // 1 / cell edge
  int nx, ny, nz;
  int base;          // first cell of this cloud in the global cell table
};

__global__ void cell_geometry_kernel(const BBox* __restrict__ bbox, const int* __restrict__ s_off, int batch,
                                     float radius, CellGrid* __restrict__ grids) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  CellGrid g;
  g.base = b * kCellsPerCloud;
  if (s_off[b + 1] == s_off[b]) {  // empty support cloud
    g.ox = g.oy = g.oz = 0.f; g.inv = 0.f; g.nx = g.ny = g.nz = 1;
    grids[b] = g;
    return;
  }
  const double mn[3] = {(double)ord2f(bbox[b].mn[0]), (double)ord2f(bbox[b].mn[1]), (double)ord2f(bbox[b].mn[2])};
  const double mx[3] = {(double)ord2f(bbox[b].mx[0]), (double)ord2f(bbox[b].mx[1]), (double)ord2f(bbox[b].mx[2])};
  // cell edge 1% above the radius: two points closer than `radius` are at most one cell apart
  // even after float rounding of the cell coordinate.
  double cell = (double)radius * 1.01;
  if (!(cell > 0.0)) cell = 1.0;
  for (int it = 0; it < 64; ++it) {
    double cells = 1.0;
    for (int a = 0; a < 3; ++a) cells *= floor((mx[a] - mn[a]) / cell) + 2.0;
    if (cells <= (double)kCellsPerCloud) break;
    cell *= 1.26;  // coarser cells stay a superset of the ball
  }
  g.ox = (float)mn[0]; g.oy = (float)mn[1]; g.oz = (float)mn[2];
  g.inv = (float)(1.0 / cell);
  g.nx = (int)floor((mx[0] - mn[0]) / cell) + 2;
  g.ny = (int)floor((mx[1] - mn[1]) / cell) + 2;
  g.nz = (int)floor((mx[2] - mn[2]) / cell) + 2;
  grids[b] = g;
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv, int n) {
  float f = floorf((x - o) * inv);
  f = fminf(fmaxf(f, -2.0f), (float)n + 1.0f);
  return (int)f;
}

__global__ void __launch_bounds__(256) cell_count_kernel(const float* __restrict__ s, const int* __restrict__ s_off,
                                                         int batch, const CellGrid* __restrict__ grids,
                                                         uint32_t* __restrict__ cell_cnt, int* __restrict__ cell_of,
                                                         int* __restrict__ rank_of) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s_off[batch]) return;
  const int b = find_segment(s_off, batch, i);
  const CellGrid g = grids[b];
  int cx = min(max(cell_coord(s[3 * i], g.ox, g.inv, g.nx), 0), g.nx - 1);
  int cy = min(max(cell_coord(s[3 * i + 1], g.oy, g.inv, g.ny), 0), g.ny - 1);
  int cz = min(max(cell_coord(s[3 * i + 2], g.oz, g.inv, g.nz), 0), g.nz - 1);
  const int c = g.base + (cz * g.ny + cy) * g.nx + cx;
  cell_of[i] = c;
  rank_of[i] = (int)atomicAdd(&cell_cnt[c], 1u);
}

__global__ void __launch_bounds__(256) cell_scatter_kernel(const float* __restrict__ s, const int* __restrict__ s_off,
                                                           int batch, const uint32_t* __restrict__ cell_start,
                                                           const int* __restrict__ cell_of, const int* __restrict__ rank_of,
                                                           float4* __restrict__ sorted) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s_off[batch]) return;
  const uint32_t p = cell_start[cell_of[i]] + (uint32_t)rank_of[i];
  sorted[p] = make_float4(s[3 * i], s[3 * i + 1], s[3 * i + 2], __int_as_float(i));
}

// reference distance: ((dx*dx) + dy*dy) + dz*dz, every operation rounded separately
__device__ __forceinline__ float ref_sqdist(float qx, float qy, float qz, float sx, float sy, float sz) {
  const float dx = __fsub_rn(qx, sx), dy = __fsub_rn(qy, sy), dz = __fsub_rn(qz, sz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

// One warp per query row.
constexpr int kInlineBatch = 64;  // clouds per call up to which the search kernel derives the query offsets itself

__global__ void __launch_bounds__(kSearchWarps * 32) radius_search_kernel(
    const float* __restrict__ q, const float4* __restrict__ sorted, const uint32_t* __restrict__ cell_start,
    const CellGrid* __restrict__ grids, const int* __restrict__ q_off_in, const int* __restrict__ s_off, int batch, float r2,
    long long* __restrict__ out, long long ld, int* __restrict__ max_count, const long long* __restrict__ q_len) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned long long sh_hits[kSearchWarps][kHitCap];
  __shared__ int sh_max;
  __shared__ int sh_qoff[kInlineBatch + 1];
  if (threadIdx.x == 0) {
    sh_max = 0;
    if (q_len != nullptr) {  // query offsets straight from the lengths (batch <= kInlineBatch): no one-CTA prep launch
      long long acc = 0;
      sh_qoff[0] = 0;
      for (int b = 0; b < batch; ++b) { acc += q_len[b]; sh_qoff[b + 1] = (int)acc; }
    }
  }
  __syncthreads();
  const int* q_off = q_len != nullptr ? sh_qoff : q_off_in;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kSearchWarps + warp;
  const int nq = q_off[batch];
  int n_hits = 0;
  if (row < nq) {
    const int b = find_segment(q_off, batch, row);
    const CellGrid g = grids[b];
    const float qx = q[3 * row], qy = q[3 * row + 1], qz = q[3 * row + 2];
    const long long pad = (long long)s_off[batch];
    unsigned long long* hits = sh_hits[warp];

    // 9 runs of x-contiguous cells: lane r in [0,9) owns (dy, dz) = (r%3-1, r/3-1)
    uint32_t my_start = 0, my_len = 0;
    if (s_off[b + 1] > s_off[b]) {
      const int cx = cell_coord(qx, g.ox, g.inv, g.nx);
      const int cy = cell_coord(qy, g.oy, g.inv, g.ny);
      const int cz = cell_coord(qz, g.oz, g.inv, g.nz);
      const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
      if (lane < 9 && x0 <= x1) {
        const int yy = cy + (lane % 3) - 1, zz = cz + (lane / 3) - 1;
        if (yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz) {
          const int c0 = g.base + (zz * g.ny + yy) * g.nx;
          my_start = cell_start[c0 + x0];
          my_len = cell_start[c0 + x1 + 1] - my_start;
        }
      }
    }
    // inclusive scan of the 9 run lengths
    uint32_t incl = my_len;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    uint32_t run_start[9], run_end[9];  // run r covers flattened candidates [run_end[r]-len, run_end[r])
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      run_end[r] = __shfl_sync(0xffffffffu, incl, r);
      run_start[r] = __shfl_sync(0xffffffffu, my_start, r);
    }
    const uint32_t total = run_end[8];

    auto candidate = [&](uint32_t c, float& d, uint32_t& idx) {
      // locate the run holding flattened candidate c
      uint32_t pos = run_start[0] + c;
#pragma unroll
      for (int r = 1; r < 9; ++r)
        if (c >= run_end[r - 1]) pos = run_start[r] + (c - run_end[r - 1]);
      const float4 p = __ldg(&sorted[pos]);
      d = ref_sqdist(qx, qy, qz, p.x, p.y, p.z);
      idx = (uint32_t)__float_as_int(p.w);
    };

    for (uint32_t base = 0; base < total; base += 32) {
      const uint32_t c = base + lane;
      bool hit = false;
      unsigned long long key = 0ull;
      if (c < total) {
        float d; uint32_t idx;
        candidate(c, d, idx);
        hit = d < r2;  // strict, nanoflann.hpp:249-252
        key = ((unsigned long long)__float_as_uint(d) << 32) | idx;
      }
      const unsigned int m = __ballot_sync(0xffffffffu, hit);
      if (out != nullptr && hit) {
        const int o = n_hits + __popc(m & ((1u << lane) - 1u));
        if (o < kHitCap) hits[o] = key;
      }
      n_hits += __popc(m);
    }
    __syncwarp();

    if (out != nullptr) {
      long long* orow = out + (long long)row * ld;
      if (n_hits <= kHitCap) {
        for (int h = lane; h < n_hits; h += 32) {
          const unsigned long long k = hits[h];
          int rank = 0;
          for (int j = 0; j < n_hits; ++j) rank += (hits[j] < k) ? 1 : 0;
          if (rank < ld) orow[rank] = (long long)(k & 0xffffffffull);
        }
      } else {
        // rare: more hits than the shared buffer holds -> rank every hit by rescanning the candidates
        for (uint32_t base = 0; base < total; base += 32) {
          const uint32_t c = base + lane;
          bool hit = false;
          unsigned long long key = ~0ull;
          if (c < total) {
            float d; uint32_t idx;
            candidate(c, d, idx);
            hit = d < r2;
            if (hit) key = ((unsigned long long)__float_as_uint(d) << 32) | idx;
          }
          int rank = 0;
          for (uint32_t base2 = 0; base2 < total; base2 += 32) {
            const uint32_t c2 = base2 + lane;
            unsigned long long key2 = ~0ull;
            if (c2 < total) {
              float d2; uint32_t idx2;
              candidate(c2, d2, idx2);
              if (d2 < r2) key2 = ((unsigned long long)__float_as_uint(d2) << 32) | idx2;
            }
#pragma unroll 8
            for (int t = 0; t < 32; ++t) {
              const unsigned long long other = __shfl_sync(0xffffffffu, key2, t);
              rank += (other < key) ? 1 : 0;
            }
          }
          if (hit && rank < ld) orow[rank] = (long long)(key & 0xffffffffull);
        }
      }
      for (long long k = (long long)n_hits + lane; k < ld; k += 32) orow[k] = pad;
    }
  }
  if (lane == 0 && n_hits > 0) atomicMax(&sh_max, n_hits);
  __syncthreads();
  if (threadIdx.x == 0 && sh_max > 0) atomicMax(max_count, sh_max);
}

struct G2Workspace {
  int* q_off;
  int* s_off;
  BBox* bbox;
  CellGrid* grids;
  uint32_t* cell_cnt;  // batch*kCellsPerCloud + 1, scanned in place
  int* cell_of;
  int* rank_of;
  float4* sorted;
  uint32_t* scan_ws;
  int* scalars;
  size_t bytes;
};

static G2Workspace carve_g2(void* ws, size_t ws_bytes, int64_t ns, int batch, bool* ok) {
  Carver c(ws, ws_bytes);
  G2Workspace w;
  const size_t ncell = (size_t)batch * kCellsPerCloud + 1;
  w.q_off = c.take<int>(batch + 1);
  w.s_off = c.take<int>(batch + 1);
  w.bbox = c.take<BBox>(batch);
  w.grids = c.take<CellGrid>(batch);
  w.cell_cnt = c.take<uint32_t>(ncell);
  w.cell_of = c.take<int>(ns);
  w.rank_of = c.take<int>(ns);
  w.sorted = c.take<float4>(ns);
  w.scan_ws = c.take<uint32_t>(scan_workspace_elems((int64_t)ncell));
  w.scalars = c.take<int>(8);
  w.bytes = c.off;
  *ok = c.ok;
  return w;
}

// =============================================================================================
// G1  grid subsampling
// =============================================================================================

struct VoxelGeom {
  float ox, oy, oz;
  float voxel;
  unsigned long long nx, ny;
};

__global__ void voxel_geometry_kernel(const BBox* __restrict__ bbox, const int* __restrict__ off, int batch, float voxel,
                                      float inv_voxel, VoxelGeom* __restrict__ geom) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  VoxelGeom g;
  g.voxel = voxel;
  if (off[b + 1] == off[b]) { g.ox = g.oy = g.oz = 0.f; g.nx = g.ny = 1ull; geom[b] = g; return; }
  const float mnx = ord2f(bbox[b].mn[0]), mny = ord2f(bbox[b].mn[1]), mnz = ord2f(bbox[b].mn[2]);
  const float mxx = ord2f(bbox[b].mx[0]), mxy = ord2f(bbox[b].mx[1]);
  // grid_subsampling_cpu.cpp:11  floor(minCorner * float(1./voxel)) * voxel
  g.ox = __fmul_rn(floorf(__fmul_rn(mnx, inv_voxel)), voxel);
  g.oy = __fmul_rn(floorf(__fmul_rn(mny, inv_voxel)), voxel);
  g.oz = __fmul_rn(floorf(__fmul_rn(mnz, inv_voxel)), voxel);
  // :13-20  size_t(floor((max - origin) / voxel) + 1), the +1 in double
  g.nx = (unsigned long long)(long long)(floor((double)__fdiv_rn(__fsub_rn(mxx, g.ox), voxel)) + 1.0);
  g.ny = (unsigned long long)(long long)(floor((double)__fdiv_rn(__fsub_rn(mxy, g.oy), voxel)) + 1.0);
  geom[b] = g;
}

__device__ __forceinline__ unsigned long long voxel_key(const VoxelGeom& g, float x, float y, float z) {
  // :32-35; negative float -> size_t behaves as (uint64)(int64) on x86-64
  const unsigned long long ix = (unsigned long long)(long long)floorf(__fdiv_rn(__fsub_rn(x, g.ox), g.voxel));
  const unsigned long long iy = (unsigned long long)(long long)floorf(__fdiv_rn(__fsub_rn(y, g.oy), g.voxel));
  const unsigned long long iz = (unsigned long long)(long long)floorf(__fdiv_rn(__fsub_rn(z, g.oz), g.voxel));
  return ix + g.nx * iy + g.nx * g.ny * iz;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

constexpr unsigned long long kEmptyKey = ~0ull;

// Hash table layout: cloud b owns slots [2*off[b] + b, 2*off[b+1] + b]  (2*len probe slots + one
// dedicated slot for the key that equals the empty marker).
__global__ void __launch_bounds__(256) voxel_insert_kernel(const float* __restrict__ pts, const int* __restrict__ off,
                                                           int batch, const VoxelGeom* __restrict__ geom,
                                                           unsigned long long* __restrict__ tab_key,
                                                           int* __restrict__ tab_first, uint32_t* __restrict__ tab_cnt,
                                                           int* __restrict__ slot_of) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= off[batch]) return;
  const int b = find_segment(off, batch, i);
  const VoxelGeom g = geom[b];
  const unsigned long long key = voxel_key(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  const long long tbase = 2ll * off[b] + b;
  const unsigned int cap = 2u * (unsigned int)(off[b + 1] - off[b]);
  long long slot;
  if (key == kEmptyKey) {
    slot = tbase + cap;
  } else {
    unsigned int h = (unsigned int)(mix64(key) % cap);
    for (;;) {
      unsigned long long* p = &tab_key[tbase + h];
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(p);
      if (cur == kEmptyKey) cur = atomicCAS(p, kEmptyKey, key);
      if (cur == kEmptyKey || cur == key) break;
      h = h + 1 == cap ? 0u : h + 1;
    }
    slot = tbase + h;
  }
  slot_of[i] = (int)slot;
  atomicMin(&tab_first[slot], i);
  atomicAdd(&tab_cnt[slot], 1u);
}

__global__ void __launch_bounds__(256) voxel_flag_kernel(const int* __restrict__ off, int batch,
                                                         const int* __restrict__ tab_first, const int* __restrict__ slot_of,
                                                         uint32_t* __restrict__ flag, int n_cap) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_cap) return;
  flag[i] = (i < off[batch] && tab_first[slot_of[i]] == i) ? 1u : 0u;
}

// first-occurrence points create their voxel record; vid = global voxel id in first-occurrence order
__global__ void __launch_bounds__(256) voxel_init_kernel(const float* __restrict__ pts, const int* __restrict__ off, int batch,
                                                         const VoxelGeom* __restrict__ geom, const uint32_t* __restrict__ fscan,
                                                         const int* __restrict__ tab_first, const uint32_t* __restrict__ tab_cnt,
                                                         const int* __restrict__ slot_of, int* __restrict__ tab_vid,
                                                         unsigned long long* __restrict__ vkey, uint32_t* __restrict__ vcnt,
                                                         int64_t* __restrict__ out_lengths, int64_t* __restrict__ out_total,
                                                         int n_cap) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = off[batch];
  if (i <= batch) {
    if (i < batch) out_lengths[i] = (int64_t)fscan[off[i + 1]] - (int64_t)fscan[off[i]];
    else *out_total = (int64_t)fscan[n];
  }
  if (i == 0) vcnt[fscan[n]] = 0u;  // terminator for the exclusive scan of vcnt
  if (i >= n) return;
  const int slot = slot_of[i];
  if (tab_first[slot] != i) return;
  const int b = find_segment(off, batch, i);
  const uint32_t v = fscan[i];
  tab_vid[slot] = (int)v;
  vkey[v] = voxel_key(geom[b], pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  vcnt[v] = tab_cnt[slot];
}

__global__ void __launch_bounds__(256) voxel_scatter_kernel(const int* __restrict__ off, int batch,
                                                            const int* __restrict__ slot_of, const int* __restrict__ tab_vid,
                                                            const uint32_t* __restrict__ voff, uint32_t* __restrict__ vfill,
                                                            int* __restrict__ plist) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= off[batch]) return;
  const int v = tab_vid[slot_of[i]];
  const uint32_t p = voff[v] + atomicAdd(&vfill[v], 1u);
  plist[p] = i;
}

// one thread per voxel: order its points by input index, sum sequentially (SampledData::update),
// scale by float(1.0 / count)  (grid_subsampling_cpu.cpp:46)
__global__ void __launch_bounds__(128) voxel_barycenter_kernel(const float* __restrict__ pts, const int* __restrict__ off,
                                                               int batch, const uint32_t* __restrict__ fscan,
                                                               const uint32_t* __restrict__ voff, int* __restrict__ plist,
                                                               float* __restrict__ bary) {
  pdl_wait();
  pdl_trigger();
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nv = fscan[off[batch]];
  if (v >= nv) return;
  const uint32_t s = voff[v], e = voff[v + 1];
  for (uint32_t a = s + 1; a < e; ++a) {  // insertion sort (voxels hold a handful of points)
    const int x = plist[a];
    uint32_t c = a;
    while (c > s && plist[c - 1] > x) { plist[c] = plist[c - 1]; --c; }
    plist[c] = x;
  }
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (uint32_t a = s; a < e; ++a) {
    const int i = plist[a];
    sx = __fadd_rn(sx, pts[3 * i]); sy = __fadd_rn(sy, pts[3 * i + 1]); sz = __fadd_rn(sz, pts[3 * i + 2]);
  }
  const float sc = (float)(1.0 / (double)(int)(e - s));
  bary[3 * v] = __fmul_rn(sx, sc); bary[3 * v + 1] = __fmul_rn(sy, sc); bary[3 * v + 2] = __fmul_rn(sz, sc);
}

// ---- libstdc++ unordered_map iteration-order replay ------------------------------------------
// Elements t = 0..D-1 are the distinct voxel keys in first-occurrence order.  The table grows
// through the bucket ladder below (measured, oracle/probe_ladder.cpp); each rehash re-inserts the
// current list in order, each insertion puts the node at the front of its bucket's run, or at
// the front of the whole list when the bucket is empty.  After a phase with bucket count nb the
// list therefore is: buckets by DESCENDING first-touch position, inside a bucket by DESCENDING
// position, where "position" indexes the phase's insertion sequence
//     [previous list ... , new elements in time order].
__constant__ unsigned int c_ladder[23] = {13u,      29u,      59u,      127u,     257u,      541u,      1109u,    2357u,
                                          5087u,    10273u,   20753u,   42043u,   85229u,    172933u,   351061u,  712697u,
                                          1447153u, 2938679u, 5967347u, 12117689u, 24607243u, 49969847u, 101473717u};
constexpr int kLadderLen = 23;
constexpr int kReplayThreads = 1024;
constexpr int kReplayCluster = 8;  // CTAs per cloud for the global-memory phases
constexpr int kReplaySmemElems = 5087;  // phases up to this bucket count run out of shared memory
constexpr unsigned int kNoTouch = 0xffffffffu;

// generic-address load that is never served from a stale L1 line (works for shared and global)
__device__ __forceinline__ unsigned int ld_cg(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }

// exclusive SUFFIX sum over w[0..n) in place, one CTA.  sh: >= 33 uints.
__device__ void block_suffix_scan(unsigned int* w, unsigned int n, unsigned int* sh) {
  const unsigned int T = blockDim.x;
  const unsigned int chunk = (n + T - 1) / T;
  const unsigned int lo = min(n, threadIdx.x * chunk), hi = min(n, lo + chunk);
  unsigned int s = 0;
  for (unsigned int i = lo; i < hi; ++i) s += ld_cg(&w[i]);
  // inclusive suffix scan across threads: reverse the lane order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = T >> 5;
  unsigned int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int y = __shfl_down_sync(0xffffffffu, x, o);
    if (lane + o < 32) x += y;
  }
  if (lane == 0) sh[warp] = x;  // warp total
  __syncthreads();
  if (warp == 0) {
    unsigned int t = lane < nwarp ? sh[lane] : 0u;
    unsigned int u = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int y = __shfl_down_sync(0xffffffffu, u, o);
      if (lane + o < 32) u += y;
    }
    sh[lane] = u - t;  // sum of warps strictly after this one
  }
  __syncthreads();
  unsigned int run = sh[warp] + (x - s);  // sum of everything after this thread's chunk
  __syncthreads();
  for (unsigned int i = hi; i > lo; --i) {
    const unsigned int v = ld_cg(&w[i - 1]);
    w[i - 1] = run;
    run += v;
  }
}

// The same exclusive SUFFIX sum by ALL CTAs of the cluster (global-memory phases: n reaches tens of thousands, and one
// CTA walking 20+ dependent L2 round trips per thread was a quarter of the phase).  Every thread takes one contiguous
// chunk; per-CTA totals cross the cluster through distributed shared memory.  Integer sums: any order gives the same
// result.  Ends with the CTA totals consumed; the caller's cluster barrier orders the writes to w.
__device__ void cluster_suffix_scan(unsigned int* w, unsigned int n, unsigned int* sh, unsigned int* sh_tot, unsigned int crank,
                                    unsigned int ncta) {
  const unsigned int T = blockDim.x;
  const unsigned int GT = ncta * T;
  const unsigned int chunk = (n + GT - 1) / GT;
  const unsigned int g = crank * T + threadIdx.x;
  const unsigned int lo = min(n, g * chunk), hi = min(n, lo + chunk);
  unsigned int s = 0;
  for (unsigned int i = lo; i < hi; ++i) s += ld_cg(&w[i]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = T >> 5;
  unsigned int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int y = __shfl_down_sync(0xffffffffu, x, o);
    if (lane + o < 32) x += y;
  }
  if (lane == 0) sh[warp] = x;  // warp total
  __syncthreads();
  if (warp == 0) {
    unsigned int t = lane < nwarp ? sh[lane] : 0u;
    unsigned int u = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int y = __shfl_down_sync(0xffffffffu, u, o);
      if (lane + o < 32) u += y;
    }
    sh[lane] = u - t;  // sum of warps strictly after this one
    if (lane == 0) *sh_tot = u;  // CTA total
  }
  // cluster barrier: totals visible to the peers (it is a CTA barrier as well)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  unsigned int after = 0;
  {
    const unsigned int local = (unsigned int)__cvta_generic_to_shared(sh_tot);
    for (unsigned int r = crank + 1; r < ncta; ++r) {
      unsigned int remote, v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
      after += v;
    }
  }
  unsigned int run = after + sh[warp] + (x - s);  // sum of everything after this thread's chunk
  for (unsigned int i = hi; i > lo; --i) {
    const unsigned int v = ld_cg(&w[i - 1]);
    w[i - 1] = run;
    run += v;
  }
}

// One phase of the replay over arrays that live either in the shared memory of the cluster's CTA 0 (CL_ACTIVE = 1:
// only that CTA works, barriers are __syncthreads) or in global memory (all CL CTAs of the cluster work, barriers are
// cluster barriers).  `bkc` caches each position's bucket so the 64-bit modulo runs once per phase.
template <bool CLUSTER>
__device__ __forceinline__ void replay_barrier() {
  if (CLUSTER) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

template <bool CLUSTER>
__device__ __forceinline__ void replay_phase(const unsigned long long* __restrict__ key, unsigned int nb, unsigned int n,
                                             unsigned int n_prev, const unsigned int* lin, unsigned int* lout, unsigned int* ft,
                                             unsigned int* cn, unsigned int* fill, unsigned int* w, unsigned int* tmp,
                                             unsigned int* bkc, unsigned int* sh_scan, unsigned int gtid, unsigned int GT,
                                             bool scan_cta, unsigned int* sh_tot = nullptr) {
  for (unsigned int k = gtid; k < nb; k += GT) { ft[k] = kNoTouch; cn[k] = 0u; fill[k] = 0u; }
  replay_barrier<CLUSTER>();
  for (unsigned int s = gtid; s < n; s += GT) {
    const unsigned int e = s < n_prev ? ld_cg(&lin[s]) : s;
    const unsigned int bk = (unsigned int)(key[e] % nb);
    bkc[s] = bk;
    atomicMin(&ft[bk], s);
    atomicAdd(&cn[bk], 1u);
  }
  replay_barrier<CLUSTER>();
  for (unsigned int s = gtid; s < n; s += GT) {
    const unsigned int bk = ld_cg(&bkc[s]);
    w[s] = (ld_cg(&ft[bk]) == s) ? ld_cg(&cn[bk]) : 0u;
  }
  replay_barrier<CLUSTER>();
  if (CLUSTER && sh_tot != nullptr) cluster_suffix_scan(w, n, sh_scan, sh_tot, gtid / blockDim.x, GT / blockDim.x);
  else if (scan_cta) block_suffix_scan(w, n, sh_scan);
  replay_barrier<CLUSTER>();
  for (unsigned int s = gtid; s < n; s += GT) {
    const unsigned int bk = ld_cg(&bkc[s]);
    const unsigned int base = ld_cg(&w[ld_cg(&ft[bk])]);
    const unsigned int slot = atomicAdd(&fill[bk], 1u);
    tmp[base + slot] = s;
  }
  replay_barrier<CLUSTER>();
  for (unsigned int s = gtid; s < n; s += GT) {
    const unsigned int bk = ld_cg(&bkc[s]);
    if (ld_cg(&ft[bk]) != s) continue;
    const unsigned int base = ld_cg(&w[s]), c = ld_cg(&cn[bk]);
    for (unsigned int a = 1; a < c; ++a) {  // descending insertion sort of the bucket's positions
      const unsigned int x = ld_cg(&tmp[base + a]);
      unsigned int j = a;
      while (j > 0 && ld_cg(&tmp[base + j - 1]) < x) { tmp[base + j] = ld_cg(&tmp[base + j - 1]); --j; }
      tmp[base + j] = x;
    }
    for (unsigned int a = 0; a < c; ++a) {
      const unsigned int s2 = ld_cg(&tmp[base + a]);
      lout[base + a] = s2 < n_prev ? ld_cg(&lin[s2]) : s2;
    }
  }
  replay_barrier<CLUSTER>();
}

// CL CTAs (one thread-block cluster) per cloud.  Phases whose tables fit in shared memory run in the cluster's CTA 0
// alone; the long tail (bucket counts 10273 ... ) runs over global memory with all CL x 1024 threads, which is what
// hides the L2 latency of the dependent gathers.  CL = 1 is the single-CTA form for small clouds.
template <int CL>
__global__ void __launch_bounds__(kReplayThreads, 1) hash_order_replay_kernel(
    const int* __restrict__ off, int batch, const uint32_t* __restrict__ fscan, const unsigned long long* __restrict__ vkey,
    const float* __restrict__ bary, unsigned int* g_list0, unsigned int* g_list1, unsigned int* g_w, unsigned int* g_tmp,
    unsigned int* g_bkc, unsigned int* g_ft, unsigned int* g_cn, unsigned int* g_fill, float* __restrict__ out_points,
    int cluster_scan) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ unsigned int smem[];
  __shared__ unsigned int sh_scan[33];
  __shared__ unsigned int sh_tot;  // this CTA's total in the cluster-wide scan (read by the peers through DSMEM)
  const int b = blockIdx.x / CL;
  const unsigned int crank = blockIdx.x % CL;  // cluster dims (CL,1,1): rank inside the cluster
  const unsigned int vbase = fscan[off[b]];
  const unsigned int D = fscan[off[b + 1]] - vbase;
  if (D == 0) return;
  const bool has_global = D > (unsigned int)kReplaySmemElems;
  if (crank != 0 && !has_global) return;  // nobody in this cluster will touch a cluster barrier
  const unsigned long long* key = vkey + vbase;
  // global scratch of this cloud (element arrays start at the cloud's first point, tables at 3x)
  const size_t ebase = (size_t)off[b];
  const size_t tbase = 3 * (size_t)off[b] + 64 * (size_t)b;
  const unsigned int tid = threadIdx.x, T = blockDim.x;
  unsigned int* lin = nullptr;
  unsigned int n_prev = 0;
  int p = 0;

  // ---- shared-memory phases (CTA 0)
  {
    unsigned int* l0 = smem; unsigned int* l1 = smem + kReplaySmemElems;
    unsigned int* w = smem + 2 * kReplaySmemElems; unsigned int* tmp = smem + 3 * kReplaySmemElems;
    unsigned int* ft = smem + 4 * kReplaySmemElems; unsigned int* cn = smem + 5 * kReplaySmemElems;
    unsigned int* fill = smem + 6 * kReplaySmemElems; unsigned int* bkc = smem + 7 * kReplaySmemElems;
    for (; p < kLadderLen && c_ladder[p] <= (unsigned int)kReplaySmemElems; ++p) {
      const unsigned int nb = c_ladder[p];
      const unsigned int n = min(D, nb);
      unsigned int* lout = (lin == l0) ? l1 : l0;  // must not alias the input list
      if (crank == 0) replay_phase<false>(key, nb, n, n_prev, lin, lout, ft, cn, fill, w, tmp, bkc, sh_scan, tid, T, true);
      lin = lout;
      n_prev = n;
      if (n == D) break;
    }
  }
  if (has_global) {
    // hand the list over to global memory, then continue with every CTA of the cluster
    unsigned int* l0 = g_list0 + ebase; unsigned int* l1 = g_list1 + ebase;
    if (crank == 0) {
      for (unsigned int s = tid; s < n_prev; s += T) l0[s] = lin[s];
    }
    lin = l0;
    replay_barrier<(CL > 1)>();
    const unsigned int gtid = crank * T + tid, GT = CL * T;
    for (; p < kLadderLen; ++p) {
      const unsigned int nb = c_ladder[p];
      const unsigned int n = min(D, nb);
      unsigned int* lout = (lin == l0) ? l1 : l0;
      replay_phase<(CL > 1)>(key, nb, n, n_prev, lin, lout, g_ft + tbase, g_cn + tbase, g_fill + tbase, g_w + ebase,
                             g_tmp + ebase, g_bkc + ebase, sh_scan, gtid, GT, crank == 0, cluster_scan ? &sh_tot : nullptr);
      lin = lout;
      n_prev = n;
      if (n == D) break;
    }
    // emission: position -> voxel
    for (unsigned int pos = gtid; pos < D; pos += GT) {
      const unsigned int e = ld_cg(&lin[pos]);
      const size_t src = 3 * (size_t)(vbase + e), dst = 3 * (size_t)(vbase + pos);
      out_points[dst] = bary[src]; out_points[dst + 1] = bary[src + 1]; out_points[dst + 2] = bary[src + 2];
    }
  } else {
    for (unsigned int pos = tid; pos < D; pos += T) {
      const unsigned int e = lin[pos];
      const size_t src = 3 * (size_t)(vbase + e), dst = 3 * (size_t)(vbase + pos);
      out_points[dst] = bary[src]; out_points[dst + 1] = bary[src + 1]; out_points[dst + 2] = bary[src + 2];
    }
  }
}

struct G1Workspace {
  int* off;
  BBox* bbox;
  VoxelGeom* geom;
  unsigned long long* tab_key;  // 2n + batch
  int* tab_first;
  uint32_t* tab_cnt;
  int* tab_vid;
  int* slot_of;      // n
  uint32_t* fscan;   // n + 1
  unsigned long long* vkey;  // n
  uint32_t* vcnt;    // n + 1 (scanned in place -> voff)
  uint32_t* vfill;   // n
  int* plist;        // n
  float* bary;       // 3n
  unsigned int *list0, *list1, *w, *tmp, *bkc;  // n each
  unsigned int *ft, *cn, *fill;           // 3n + 64*batch each
  uint32_t* scan_ws;
  int* scalars;
  size_t tab_slots;
  size_t tbl_elems;
  size_t zero_bytes;  // span tab_cnt .. end of vfill (cleared by one memset)
  size_t bytes;
};

static G1Workspace carve_g1(void* ws, size_t ws_bytes, int64_t n, int batch, bool* ok) {
  Carver c(ws, ws_bytes);
  G1Workspace w;
  w.tab_slots = 2 * (size_t)n + (size_t)batch + 1;
  w.tbl_elems = 3 * (size_t)n + 64 * (size_t)batch + 64;
  w.off = c.take<int>(batch + 1);
  w.bbox = c.take<BBox>(batch);
  w.geom = c.take<VoxelGeom>(batch);
  w.tab_key = c.take<unsigned long long>(w.tab_slots);
  w.tab_first = c.take<int>(w.tab_slots);
  // tab_cnt | vcnt | vfill are carved back to back: one memset clears all three (zero_bytes)
  w.tab_cnt = c.take<uint32_t>(w.tab_slots);
  w.vcnt = c.take<uint32_t>(n + 2);
  w.vfill = c.take<uint32_t>(n + 1);
  w.zero_bytes = (size_t)(reinterpret_cast<char*>(w.vfill + (n + 1)) - reinterpret_cast<char*>(w.tab_cnt));
  w.tab_vid = c.take<int>(w.tab_slots);
  w.slot_of = c.take<int>(n + 1);
  w.fscan = c.take<uint32_t>(n + 2);
  w.vkey = c.take<unsigned long long>(n + 1);
  w.plist = c.take<int>(n + 1);
  w.bary = c.take<float>(3 * (size_t)n + 3);
  w.list0 = c.take<unsigned int>(n + 1);
  w.list1 = c.take<unsigned int>(n + 1);
  w.w = c.take<unsigned int>(n + 1);
  w.tmp = c.take<unsigned int>(n + 1);
  w.bkc = c.take<unsigned int>(n + 1);
  w.ft = c.take<unsigned int>(w.tbl_elems);
  w.cn = c.take<unsigned int>(w.tbl_elems);
  w.fill = c.take<unsigned int>(w.tbl_elems);
  w.scan_ws = c.take<uint32_t>(scan_workspace_elems(n + 2));
  w.scalars = c.take<int>(8);
  w.bytes = c.off;
  *ok = c.ok;
  return w;
}

}  // namespace gr

// =================================================================================================
// C ABI
// =================================================================================================
using namespace gr;

extern "C" size_t gr_radius_neighbors_workspace_size(int64_t nq, int64_t ns, int batch) {
  (void)nq;
  bool ok;
  return carve_g2(nullptr, 0, ns, batch, &ok).bytes;
}

/* Same search; `reuse_grid` != 0 promises that `ws` still holds the cell grid a previous call built for the SAME
 * (s_points, s_lengths, radius): the support cloud is not binned again (of the 13 searches of one pyramid only 5 use
 * a new support cloud / radius combination). */
// count_zeroed: *out_max_count is already 0 in stream order (gr_radius_pyramid clears all its counters with one memset)
static int radius_neighbors_impl(const float* q_points, const float* s_points, const int64_t* q_lengths,
                                 const int64_t* s_lengths, int batch, int64_t nq, int64_t ns, float radius,
                                 int64_t* out_idx, int64_t ld, int32_t* out_max_count, void* ws, size_t ws_bytes,
                                 int reuse_grid, bool count_zeroed, void* stream) {
  if (batch <= 0 || nq < 0 || ns < 0 || !q_lengths || !s_lengths || !out_max_count || (out_idx && ld <= 0) ||
      nq >= (1ll << 31) || ns >= (1ll << 31) || (int64_t)batch * kCellsPerCloud >= (1ll << 31))
    return GR_ERR_BAD_ARG;
  if ((nq > 0 && !q_points) || (ns > 0 && !s_points)) return GR_ERR_BAD_ARG;
  bool ok;
  G2Workspace w = carve_g2(ws, ws_bytes, ns, batch, &ok);
  if (!ws || !ok) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t ncell = (size_t)batch * kCellsPerCloud + 1;

  if (!count_zeroed) GR_CHECK_CUDA(cudaMemsetAsync(out_max_count, 0, sizeof(int32_t), st));
  const long long* q_len_inline = batch <= kInlineBatch ? reinterpret_cast<const long long*>(q_lengths) : nullptr;
  if (reuse_grid) {
    // only the query offsets change: the search kernel derives them from the lengths itself (small batches), else one
    // prep launch (the bounding boxes that kernel resets are not needed once the grid exists)
    if (!q_len_inline) {
      GR_CHECK_CUDA(launch_pdl(prep_offsets_kernel, dim3(1), dim3(256), (size_t)(0), st, q_lengths, w.q_off, nullptr, nullptr, batch, w.bbox, w.scalars, 8));
      GR_CHECK_LAUNCH("prep_offsets_kernel");
    }
    if (nq > 0) {
      const float r2 = radius * radius;
      GR_CHECK_CUDA(launch_pdl(radius_search_kernel, dim3(ceil_div(nq, kSearchWarps)), dim3(kSearchWarps * 32), (size_t)(0), st, q_points, w.sorted, w.cell_cnt, w.grids, w.q_off, w.s_off, batch, r2, reinterpret_cast<long long*>(out_idx),
          (long long)ld, out_max_count, q_len_inline));
      GR_CHECK_LAUNCH("radius_search_kernel");
    }
    return GR_OK;
  }
  GR_CHECK_CUDA(cudaMemsetAsync(w.cell_cnt, 0, ncell * sizeof(uint32_t), st));
  GR_CHECK_CUDA(launch_pdl(prep_offsets_kernel, dim3(1), dim3(256), (size_t)(0), st, q_lengths, w.q_off, s_lengths, w.s_off, batch, w.bbox, w.scalars, 8));
  GR_CHECK_LAUNCH("prep_offsets_kernel");
  if (ns > 0) {
    GR_CHECK_CUDA(launch_pdl(bbox_kernel, dim3(ceil_div(ns, 256)), dim3(256), (size_t)(0), st, s_points, w.s_off, batch, w.bbox));
    GR_CHECK_LAUNCH("bbox_kernel");
  }
  GR_CHECK_CUDA(launch_pdl(cell_geometry_kernel, dim3(ceil_div(batch, 128)), dim3(128), (size_t)(0), st, w.bbox, w.s_off, batch, radius, w.grids));
  GR_CHECK_LAUNCH("cell_geometry_kernel");
  if (ns > 0) {
    GR_CHECK_CUDA(launch_pdl(cell_count_kernel, dim3(ceil_div(ns, 256)), dim3(256), (size_t)(0), st, s_points, w.s_off, batch, w.grids, w.cell_cnt, w.cell_of,
                                                         w.rank_of));
    GR_CHECK_LAUNCH("cell_count_kernel");
  }
  int rc = exclusive_scan_u32(w.cell_cnt, w.cell_cnt, (int64_t)ncell, w.scan_ws, st);
  if (rc != GR_OK) return rc;
  if (ns > 0) {
    GR_CHECK_CUDA(launch_pdl(cell_scatter_kernel, dim3(ceil_div(ns, 256)), dim3(256), (size_t)(0), st, s_points, w.s_off, batch, w.cell_cnt, w.cell_of, w.rank_of,
                                                           w.sorted));
    GR_CHECK_LAUNCH("cell_scatter_kernel");
  }
  if (nq > 0) {
    const float r2 = radius * radius;  // radius_neighbors_cpu.cpp:12 (host float multiply, one rounding)
    GR_CHECK_CUDA(launch_pdl(radius_search_kernel, dim3(ceil_div(nq, kSearchWarps)), dim3(kSearchWarps * 32), (size_t)(0), st, q_points, w.sorted, w.cell_cnt, w.grids, w.q_off, w.s_off, batch, r2, reinterpret_cast<long long*>(out_idx),
        (long long)ld, out_max_count, (const long long*)nullptr));
    GR_CHECK_LAUNCH("radius_search_kernel");
  }
  return GR_OK;
}

extern "C" int gr_radius_neighbors_cached(const float* q_points, const float* s_points, const int64_t* q_lengths,
                                          const int64_t* s_lengths, int batch, int64_t nq, int64_t ns, float radius,
                                          int64_t* out_idx, int64_t ld, int32_t* out_max_count, void* ws, size_t ws_bytes,
                                          int reuse_grid, void* stream) {
  return radius_neighbors_impl(q_points, s_points, q_lengths, s_lengths, batch, nq, ns, radius, out_idx, ld, out_max_count, ws, ws_bytes,
                               reuse_grid, false, stream);
}

extern "C" int gr_radius_neighbors(const float* q_points, const float* s_points, const int64_t* q_lengths,
                                   const int64_t* s_lengths, int batch, int64_t nq, int64_t ns, float radius,
                                   int64_t* out_idx, int64_t ld, int32_t* out_max_count, void* ws, size_t ws_bytes,
                                   void* stream) {
  return gr_radius_neighbors_cached(q_points, s_points, q_lengths, s_lengths, batch, nq, ns, radius, out_idx, ld, out_max_count,
                                    ws, ws_bytes, 0, stream);
}

extern "C" size_t gr_grid_subsample_workspace_size(int64_t n_points, int batch) {
  bool ok;
  return carve_g1(nullptr, 0, n_points, batch, &ok).bytes;
}

extern "C" int gr_grid_subsample(const float* points, const int64_t* lengths, int batch, int64_t n_points,
                                 float voxel_size, float* out_points, int64_t* out_lengths, int64_t* out_total, void* ws,
                                 size_t ws_bytes, void* stream) {
  if (batch <= 0 || n_points < 0 || !lengths || !out_lengths || !out_total || !(voxel_size > 0.f) ||
      n_points >= (1ll << 29))
    return GR_ERR_BAD_ARG;
  if (n_points > 0 && (!points || !out_points)) return GR_ERR_BAD_ARG;
  bool ok;
  G1Workspace w = carve_g1(ws, ws_bytes, n_points, batch, &ok);
  if (!ws || !ok) return GR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = (int)n_points;
  const int blocks = ceil_div(n + 1, 256);

  GR_CHECK_CUDA(cudaMemsetAsync(w.tab_key, 0xff, w.tab_slots * sizeof(unsigned long long), st));
  GR_CHECK_CUDA(cudaMemsetAsync(w.tab_first, 0x7f, w.tab_slots * sizeof(int), st));
  // tab_cnt, vfill and vcnt in one memset (adjacent in the workspace).  vcnt: the scan below walks all n + 1 capacity slots;
  // only the first (number of voxels) + 1 are written by voxel_init_kernel -- the tail is cleared so that no kernel reads
  // uninitialised memory (initcheck-clean)
  GR_CHECK_CUDA(cudaMemsetAsync(w.tab_cnt, 0, w.zero_bytes, st));
  GR_CHECK_CUDA(launch_pdl(prep_offsets_kernel, dim3(1), dim3(256), (size_t)(0), st, lengths, w.off, nullptr, nullptr, batch, w.bbox, w.scalars, 8));
  GR_CHECK_LAUNCH("prep_offsets_kernel");
  if (n > 0) {
    GR_CHECK_CUDA(launch_pdl(bbox_kernel, dim3(ceil_div(n, 256)), dim3(256), (size_t)(0), st, points, w.off, batch, w.bbox));
    GR_CHECK_LAUNCH("bbox_kernel");
  }
  const float inv_voxel = (float)(1.0 / (double)voxel_size);  // "1. / voxel_size" narrowed by cloud.h:83
  GR_CHECK_CUDA(launch_pdl(voxel_geometry_kernel, dim3(ceil_div(batch, 128)), dim3(128), (size_t)(0), st, w.bbox, w.off, batch, voxel_size, inv_voxel, w.geom));
  GR_CHECK_LAUNCH("voxel_geometry_kernel");
  if (n > 0) {
    GR_CHECK_CUDA(launch_pdl(voxel_insert_kernel, dim3(ceil_div(n, 256)), dim3(256), (size_t)(0), st, points, w.off, batch, w.geom, w.tab_key, w.tab_first, w.tab_cnt,
                                                          w.slot_of));
    GR_CHECK_LAUNCH("voxel_insert_kernel");
  }
  GR_CHECK_CUDA(launch_pdl(voxel_flag_kernel, dim3(blocks), dim3(256), (size_t)(0), st, w.off, batch, w.tab_first, w.slot_of, w.fscan, n));
  GR_CHECK_LAUNCH("voxel_flag_kernel");
  int rc = exclusive_scan_u32(w.fscan, w.fscan, (int64_t)n + 1, w.scan_ws, st);
  if (rc != GR_OK) return rc;
  GR_CHECK_CUDA(launch_pdl(voxel_init_kernel, dim3(ceil_div((int64_t)max(n, batch + 1), 256)), dim3(256), (size_t)(0), st, points, w.off, batch, w.geom, w.fscan, w.tab_first, w.tab_cnt, w.slot_of, w.tab_vid, w.vkey, w.vcnt, out_lengths,
      out_total, n));
  GR_CHECK_LAUNCH("voxel_init_kernel");
  if (n == 0) return GR_OK;
  rc = exclusive_scan_u32(w.vcnt, w.vcnt, (int64_t)n + 1, w.scan_ws, st);
  if (rc != GR_OK) return rc;
  GR_CHECK_CUDA(launch_pdl(voxel_scatter_kernel, dim3(ceil_div(n, 256)), dim3(256), (size_t)(0), st, w.off, batch, w.slot_of, w.tab_vid, w.vcnt, w.vfill, w.plist));
  GR_CHECK_LAUNCH("voxel_scatter_kernel");
  GR_CHECK_CUDA(launch_pdl(voxel_barycenter_kernel, dim3(ceil_div(n, 128)), dim3(128), (size_t)(0), st, points, w.off, batch, w.fscan, w.vcnt, w.plist, w.bary));
  GR_CHECK_LAUNCH("voxel_barycenter_kernel");
  {
    // clouds that may exceed the shared-memory phases get a cluster of kReplayCluster CTAs each
    const size_t smem = 8 * (size_t)kReplaySmemElems * sizeof(unsigned int);
    GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(hash_order_replay_kernel<1>), (int)smem));
    GR_CHECK_CUDA(ensure_smem_attr(reinterpret_cast<const void*>(hash_order_replay_kernel<kReplayCluster>), (int)smem));
    static int cl_knob = -1;
    if (cl_knob < 0) { const char* e = getenv("GAUSSREG_REPLAY_CLUSTER"); cl_knob = e ? atoi(e) : kReplayCluster; }
    static int scan_knob = -1;  // 1: the suffix scans of the global-memory phases run on every CTA of the cluster
    if (scan_knob < 0) { const char* e = getenv("GAUSSREG_REPLAY_SCAN"); scan_knob = e ? atoi(e) : 1; }
    if (n > kReplaySmemElems && cl_knob > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)batch * kReplayCluster);
      cfg.blockDim = dim3(kReplayThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kReplayCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      GR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, hash_order_replay_kernel<kReplayCluster>, (const int*)w.off, batch,
                                       (const uint32_t*)w.fscan, (const unsigned long long*)w.vkey, (const float*)w.bary, w.list0,
                                       w.list1, w.w, w.tmp, w.bkc, w.ft, w.cn, w.fill, out_points, scan_knob));
    } else {
      GR_CHECK_CUDA(launch_pdl(hash_order_replay_kernel<1>, dim3(batch), dim3(kReplayThreads), (size_t)(smem), st, w.off, batch, w.fscan, w.vkey, w.bary, w.list0, w.list1,
                                                                        w.w, w.tmp, w.bkc, w.ft, w.cn, w.fill, out_points, 0));
    }
  }
  GR_CHECK_LAUNCH("hash_order_replay_kernel");
  return GR_OK;
}

/* G1 chain: every subsampling stage of a pyramid in one call (see include/gaussreg_b200.h). */
extern "C" int gr_grid_subsample_chain(const float* points, const int64_t* lengths, int batch, int64_t n_points,
                                       const float* voxel_sizes, int n_sub, float* out_points, int64_t* out_lengths,
                                       int64_t* out_totals, void* ws, size_t ws_bytes, void** stage_events, void* stream) {
  if (n_sub < 0 || n_sub > 15 || batch <= 0 || n_points < 0 || !lengths || (n_sub > 0 && (!voxel_sizes || !out_lengths || !out_totals)))
    return GR_ERR_BAD_ARG;
  // events per (thread, device), created once: a later cudaStreamWaitEvent refers to the record that precedes it, so the
  // handles can be re-recorded by the next chain while earlier waits are still pending
  struct Pool { int dev; cudaEvent_t ev[16]; };
  static thread_local std::vector<Pool> pools;
  Pool* pool = nullptr;
  if (stage_events) {
    int dev = 0;
    GR_CHECK_CUDA(cudaGetDevice(&dev));
    for (Pool& p : pools) if (p.dev == dev) pool = &p;
    if (!pool) {
      Pool p;
      p.dev = dev;
      for (int i = 0; i < 16; ++i) GR_CHECK_CUDA(cudaEventCreateWithFlags(&p.ev[i], cudaEventDisableTiming));
      pools.push_back(p);
      pool = &pools.back();
    }
    stage_events[0] = nullptr;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* in_pts = points;
  const int64_t* in_len = lengths;
  for (int i = 0; i < n_sub; ++i) {
    float* o_pts = out_points + (size_t)i * (size_t)n_points * 3;
    int64_t* o_len = out_lengths + (size_t)i * batch;
    const int rc = gr_grid_subsample(in_pts, in_len, batch, n_points, voxel_sizes[i], o_pts, o_len, out_totals + i, ws, ws_bytes, stream);
    if (rc != GR_OK) return rc;
    if (stage_events) {
      GR_CHECK_CUDA(cudaEventRecord(pool->ev[i + 1], st));
      stage_events[i + 1] = pool->ev[i + 1];
    }
    in_pts = o_pts;
    in_len = o_len;
  }
  return GR_OK;
}

/* G3: the pyramid's searches in one call (see include/gaussreg_b200.h). */
extern "C" int gr_radius_pyramid(const float* const* stage_points, const int64_t* const* stage_lengths, int n_stages, int batch,
                                 int64_t capacity, void* const* stage_grid_ws, size_t grid_ws_bytes, void* const* stage_ready_events,
                                 const gr_pyramid_search* searches, int n_searches, uint32_t built_mask, void* stream) {
  if (!stage_points || !stage_lengths || !stage_grid_ws || !searches || n_stages <= 0 || n_stages > 16 || n_searches < 0 || batch <= 0 ||
      capacity < 0)
    return GR_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool built[16] = {};
  for (int i = 0; i < 16; ++i) built[i] = (built_mask >> i) & 1u;  // grids an earlier call already left in stage_grid_ws
  // the searches' counters: one memset when they are consecutive int32s (the usual (n_searches,) tensor), else one each
  bool counters_contiguous = n_searches > 0 && searches[0].out_max_count != nullptr;
  for (int j = 0; j < n_searches && counters_contiguous; ++j) counters_contiguous = searches[j].out_max_count == searches[0].out_max_count + j;
  if (counters_contiguous) GR_CHECK_CUDA(cudaMemsetAsync(searches[0].out_max_count, 0, (size_t)n_searches * sizeof(int32_t), st));
  int waited = 0;
  for (int j = 0; j < n_searches; ++j) {
    const gr_pyramid_search& q = searches[j];
    if (q.query_stage < 0 || q.query_stage >= n_stages || q.support_stage < 0 || q.support_stage >= n_stages) return GR_ERR_BAD_ARG;
    const int need = q.query_stage > q.support_stage ? q.query_stage : q.support_stage;
    while (waited < need) {
      ++waited;
      if (stage_ready_events && stage_ready_events[waited])
        GR_CHECK_CUDA(cudaStreamWaitEvent(st, static_cast<cudaEvent_t>(stage_ready_events[waited]), 0));
    }
    const int rc = radius_neighbors_impl(stage_points[q.query_stage], stage_points[q.support_stage], stage_lengths[q.query_stage],
                                         stage_lengths[q.support_stage], batch, capacity, capacity, q.radius, q.out_idx, q.limit,
                                         q.out_max_count, stage_grid_ws[q.support_stage], grid_ws_bytes,
                                         built[q.support_stage] ? 1 : 0, counters_contiguous, stream);
    if (rc != GR_OK) return rc;
    built[q.support_stage] = true;
  }
  return GR_OK;
}
