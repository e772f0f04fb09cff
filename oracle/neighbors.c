/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the neighbour pyramid.
 *
 * Plain-C restatement of the reference's two native ops.  Nothing under gaussreg_b200/ may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Pinning: the reference has no golden vectors, so this restatement is
 * pinned against the reference's own sources compiled by oracle/Makefile into
 * oracle/_ref/libgaussreg_ref.so (tests/test_oracle_neighbors.py), and against fixtures in
 * tests/golden/ generated from that library (tests/golden/make_neighbor_golden.py).
 *
 * G1  grid_subsampling      <- geotransformer/extensions/cpu/grid_subsampling/grid_subsampling_cpu.cpp:3-75
 * G2  radius_neighbors      <- geotransformer/extensions/cpu/radius_neighbors/radius_neighbors_cpu.cpp:3-91
 *                              + geotransformer/extensions/extra/nanoflann/nanoflann.hpp:249-252,432-440,1280-1288
 *
 * Build: gcc -O3 -ffp-contract=off (no FMA contraction: the reference is built for baseline
 * x86-64 where float a*b+c is two roundings).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * libstdc++ std::unordered_map<size_t, T> emulation (GCC 13 <bits/hashtable.h>):
 *   - identity hash, bucket = key % bucket_count
 *   - all nodes live on ONE singly linked list; bucket[b] points at the node *before* the
 *     first node of bucket b (or at the list head sentinel)
 *   - a node inserted into a non-empty bucket goes to the FRONT of that bucket's run;
 *     a node inserted into an empty bucket goes to the FRONT of the whole list
 *   - _Prime_rehash_policy, max_load_factor 1.0, growth factor 2: rehash to the next listed
 *     prime >= max(n+1, 2*nb) when n+1 > nb; first insert goes 1 -> 13.  The rehash walks the
 *     list in order and re-inserts every node with the same two rules.
 * The resulting iteration order is what grid_subsampling_cpu.cpp:45-47 emits.
 * The prime ladder below was measured with oracle/probe_ladder.cpp (g++ 13.3) and is
 * re-checked by tests/test_oracle_neighbors.py::test_bucket_ladder.
 * ------------------------------------------------------------------------------------------ */
static const uint64_t k_ladder[] = {
    13ull,      29ull,      59ull,       127ull,      257ull,      541ull,
    1109ull,    2357ull,    5087ull,     10273ull,    20753ull,    42043ull,
    85229ull,   172933ull,  351061ull,   712697ull,   1447153ull,  2938679ull,
    5967347ull, 12117689ull, 24607243ull, 49969847ull, 101473717ull};
#define N_LADDER ((int)(sizeof(k_ladder) / sizeof(k_ladder[0])))

int oracle_ladder(uint64_t* out, int cap) {
  int n = N_LADDER < cap ? N_LADDER : cap;
  for (int i = 0; i < n; ++i) out[i] = k_ladder[i];
  return N_LADDER;
}

#define NIL (-1)
#define HEAD (-2) /* "before begin" sentinel */

typedef struct {
  int64_t n;        /* element count */
  uint64_t nb;      /* bucket count */
  int ladder_pos;   /* index of nb in k_ladder, -1 while nb == 1 */
  int64_t head;     /* first node */
  int64_t* bucket;  /* node-before-first of each bucket, NIL when empty */
  int64_t* next;    /* per node */
  uint64_t* key;    /* per node */
} htab_t;

static inline int64_t next_of(const htab_t* h, int64_t prev) { return prev == HEAD ? h->head : h->next[prev]; }
static inline void set_next_of(htab_t* h, int64_t prev, int64_t v) {
  if (prev == HEAD) h->head = v; else h->next[prev] = v;
}

static void htab_link(htab_t* h, int64_t* bucket, uint64_t nb, int64_t node, uint64_t* bbegin_bkt, int rehashing) {
  uint64_t b = h->key[node] % nb;
  if (bucket[b] != NIL) {
    h->next[node] = next_of(h, bucket[b]);
    set_next_of(h, bucket[b], node);
  } else {
    h->next[node] = h->head;
    h->head = node;
    if (h->next[node] != NIL) {
      /* the bucket that used to start at the list head now starts after `node` */
      uint64_t ob = rehashing ? *bbegin_bkt : (h->key[h->next[node]] % nb);
      bucket[ob] = node;
    }
    bucket[b] = HEAD;
    if (rehashing) *bbegin_bkt = b;
  }
}

static int htab_rehash(htab_t* h) {
  if (h->ladder_pos + 1 >= N_LADDER) return -1;
  h->ladder_pos += 1;
  uint64_t nb = k_ladder[h->ladder_pos];
  int64_t* nbk = (int64_t*)malloc(sizeof(int64_t) * nb);
  if (!nbk) return -1;
  for (uint64_t i = 0; i < nb; ++i) nbk[i] = NIL;
  int64_t p = h->head;
  h->head = NIL;
  uint64_t bbegin = 0;
  while (p != NIL) {
    int64_t nx = h->next[p];
    htab_link(h, nbk, nb, p, &bbegin, 1);
    p = nx;
  }
  free(h->bucket);
  h->bucket = nbk;
  h->nb = nb;
  return 0;
}

/* returns node index of key, inserting if absent (*inserted set) */
static int64_t htab_find_or_insert(htab_t* h, uint64_t key, int* inserted) {
  *inserted = 0;
  if (h->n > 0) {
    uint64_t b = key % h->nb;
    int64_t prev = h->bucket[b];
    if (prev != NIL) {
      int64_t p = next_of(h, prev);
      for (;;) {
        if (h->key[p] == key) return p;
        int64_t nx = h->next[p];
        if (nx == NIL || h->key[nx] % h->nb != b) break;
        p = nx;
      }
    }
  }
  /* _M_need_rehash(n_bkt, n_elt, 1): with load factor 1 the table holds at most nb nodes */
  if ((uint64_t)(h->n + 1) > (h->ladder_pos < 0 ? 0 : h->nb)) {
    if (htab_rehash(h) != 0) return -1;
  }
  int64_t node = h->n++;
  h->key[node] = key;
  htab_link(h, h->bucket, h->nb, node, NULL, 0);
  *inserted = 1;
  return node;
}

/* G1, one cloud.  grid_subsampling_cpu.cpp:3-48 */
static int64_t grid_subsample_single(const float* pts, int64_t n, float voxel, float* out) {
  if (n <= 0) return 0; /* the reference reads points[0] (cloud.cpp:5,23): empty clouds are UB there */
  float mnx = pts[0], mny = pts[1], mnz = pts[2];
  float mxx = mnx, mxy = mny, mxz = mnz;
  for (int64_t i = 0; i < n; ++i) { /* cloud.cpp:4-37, strict comparisons */
    float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (x < mnx) mnx = x;
    if (y < mny) mny = y;
    if (z < mnz) mnz = z;
    if (x > mxx) mxx = x;
    if (y > mxy) mxy = y;
    if (z > mxz) mxz = z;
  }
  /* :11  floor(minCorner * (1. / voxel_size)) * voxel_size ; "1./voxel" is a double division
   * narrowed to float by operator*(PointXYZ, const float) (cloud.h:83) */
  const float inv = (float)(1.0 / (double)voxel);
  const float ox = floorf(mnx * inv) * voxel;
  const float oy = floorf(mny * inv) * voxel;
  const float oz = floorf(mnz * inv) * voxel;
  /* :13-20  size_t(floor(float) + 1) evaluated in double */
  const uint64_t nx = (uint64_t)(int64_t)(floor((double)((mxx - ox) / voxel)) + 1.0);
  const uint64_t ny = (uint64_t)(int64_t)(floor((double)((mxy - oy) / voxel)) + 1.0);

  htab_t h;
  h.n = 0; h.nb = 1; h.ladder_pos = -1; h.head = NIL;
  h.bucket = NULL;
  h.next = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  h.key = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
  float* sum = (float*)calloc((size_t)n * 3, sizeof(float));
  int* cnt = (int*)calloc((size_t)n, sizeof(int));
  if (!h.next || !h.key || !sum || !cnt) return -1;

  for (int64_t i = 0; i < n; ++i) { /* :28-42 */
    float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    /* negative -> size_t is (uint64)(int64) on x86-64 (cvttsd2si) */
    uint64_t ix = (uint64_t)(int64_t)floor((double)((x - ox) / voxel));
    uint64_t iy = (uint64_t)(int64_t)floor((double)((y - oy) / voxel));
    uint64_t iz = (uint64_t)(int64_t)floor((double)((z - oz) / voxel));
    uint64_t key = ix + nx * iy + nx * ny * iz; /* wraps mod 2^64 like size_t */
    int ins;
    int64_t node = htab_find_or_insert(&h, key, &ins);
    if (node < 0) return -1;
    cnt[node] += 1; /* grid_subsampling_cpu.h:17-20: sequential float += in input order */
    sum[3 * node] += x;
    sum[3 * node + 1] += y;
    sum[3 * node + 2] += z;
  }
  int64_t m = 0;
  for (int64_t p = h.head; p != NIL; p = h.next[p]) { /* :45-47 */
    const float s = (float)(1.0 / (double)cnt[p]);
    out[3 * m] = sum[3 * p] * s;
    out[3 * m + 1] = sum[3 * p + 1] * s;
    out[3 * m + 2] = sum[3 * p + 2] * s;
    ++m;
  }
  free(h.bucket); free(h.next); free(h.key); free(sum); free(cnt);
  return m;
}

/* G1, stacked batch.  grid_subsampling_cpu.cpp:50-75.  Returns total output points or -1. */
int64_t oracle_grid_subsample(const float* points, const int64_t* lengths, int batch, float voxel,
                              float* out_points, int64_t* out_lengths) {
  int64_t start = 0, total = 0;
  for (int b = 0; b < batch; ++b) {
    int64_t m = grid_subsample_single(points + 3 * start, lengths[b], voxel, out_points + 3 * total);
    if (m < 0) return -1;
    out_lengths[b] = m;
    total += m;
    start += lengths[b];
  }
  return total;
}

/* ------------------------------------------------------------------------------------------
 * G2.  Every support point j of the query's batch element with
 *        d = ((qx-sx)^2 + (qy-sy)^2) + (qz-sz)^2  <  radius*radius      (float32, strict)
 * sorted ascending by d.  The reference's std::sort is unstable on equal d; the oracle (and the
 * CUDA path) break ties by ascending index, tests canonicalise the reference's rows the same way.
 * Rows are padded with Ns_total; output width = max row count (radius_neighbors_cpu.cpp:68-90).
 *
 * Candidates come from a uniform cell grid (cell edge > radius) instead of the KD-tree: the
 * candidate set is a superset of the ball, the acceptance test is the formula above, so the
 * result is that of an exhaustive scan (oracle_radius_neighbors_brute, kept for the tests).
 * ------------------------------------------------------------------------------------------ */
typedef struct { float d; int64_t j; } cand_t;

static int cand_cmp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->j > y->j) - (x->j < y->j);
}

static inline float sqdist(const float* q, const float* s) {
  float dx = q[0] - s[0], dy = q[1] - s[1], dz = q[2] - s[2];
  float r = dx * dx; /* 0 + dx*dx */
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

typedef struct {
  int64_t* rows;    /* concatenated sorted neighbour lists */
  int64_t* offs;    /* nq + 1 */
  int64_t cap;
} rows_t;

static int rows_push(rows_t* r, int64_t pos, int64_t v) {
  if (pos >= r->cap) {
    int64_t nc = r->cap * 2 + 1024;
    int64_t* p = (int64_t*)realloc(r->rows, sizeof(int64_t) * (size_t)nc);
    if (!p) return -1;
    r->rows = p; r->cap = nc;
  }
  r->rows[pos] = v;
  return 0;
}

static int64_t emit_padded(const rows_t* r, int64_t nq, int64_t ns_total, int64_t* out_idx, int64_t out_cap_width) {
  int64_t w = 0;
  for (int64_t i = 0; i < nq; ++i) {
    int64_t c = r->offs[i + 1] - r->offs[i];
    if (c > w) w = c;
  }
  if (out_idx == NULL) return w;
  if (w > out_cap_width) return -2;
  for (int64_t i = 0; i < nq; ++i) {
    int64_t c = r->offs[i + 1] - r->offs[i];
    for (int64_t k = 0; k < w; ++k) out_idx[i * w + k] = k < c ? r->rows[r->offs[i] + k] : ns_total;
  }
  return w;
}

/* Returns the output width (global max count).  out_idx must hold nq*out_cap_width int64;
 * pass out_idx == NULL to only query the width. */
int64_t oracle_radius_neighbors(const float* q, const float* s, const int64_t* q_len, const int64_t* s_len,
                                int batch, float radius, int64_t* out_idx, int64_t out_cap_width) {
  int64_t nq = 0, ns = 0;
  for (int b = 0; b < batch; ++b) { nq += q_len[b]; ns += s_len[b]; }
  const float r2 = radius * radius; /* radius_neighbors_cpu.cpp:12 */
  rows_t R; R.rows = NULL; R.cap = 0;
  R.offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nq + 1));
  if (!R.offs) return -1;
  R.offs[0] = 0;
  int64_t pos = 0, q0 = 0, s0 = 0;
  cand_t* cand = NULL; int64_t cand_cap = 0;
  for (int b = 0; b < batch; ++b) {
    const int64_t nsb = s_len[b], nqb = q_len[b];
    const float* sp = s + 3 * s0;
    /* cell grid over the support cloud */
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int64_t j = 0; j < nsb; ++j)
      for (int a = 0; a < 3; ++a) {
        double v = sp[3 * j + a];
        if (j == 0 || v < mn[a]) mn[a] = v;
        if (j == 0 || v > mx[a]) mx[a] = v;
      }
    double cell = (double)radius * 1.0001 + 1e-12;
    int64_t dim[3];
    for (int a = 0; a < 3; ++a) {
      double ext = (mx[a] - mn[a]) / cell;
      if (ext > 1024.0) { /* cap the table; coarser cells stay a superset */
        cell = (mx[a] - mn[a]) / 1024.0;
      }
    }
    for (int a = 0; a < 3; ++a) dim[a] = (int64_t)floor((mx[a] - mn[a]) / cell) + 1;
    int64_t ncell = dim[0] * dim[1] * dim[2];
    int64_t* cstart = (int64_t*)calloc((size_t)(ncell + 1), sizeof(int64_t));
    int64_t* cidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nsb > 0 ? nsb : 1));
    int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nsb > 0 ? nsb : 1));
    if (!cstart || !cidx || !order) return -1;
    for (int64_t j = 0; j < nsb; ++j) {
      int64_t c[3];
      for (int a = 0; a < 3; ++a) {
        c[a] = (int64_t)floor(((double)sp[3 * j + a] - mn[a]) / cell);
        if (c[a] < 0) c[a] = 0;
        if (c[a] >= dim[a]) c[a] = dim[a] - 1;
      }
      cidx[j] = c[0] + dim[0] * (c[1] + dim[1] * c[2]);
      cstart[cidx[j] + 1] += 1;
    }
    for (int64_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
    {
      int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncell > 0 ? ncell : 1));
      if (!fill) return -1;
      memcpy(fill, cstart, sizeof(int64_t) * (size_t)ncell);
      for (int64_t j = 0; j < nsb; ++j) order[fill[cidx[j]]++] = j;
      free(fill);
    }
    for (int64_t i = 0; i < nqb; ++i) {
      const float* qp = q + 3 * (q0 + i);
      int64_t nc = 0;
      if (nsb > 0) {
        int64_t lo[3], hi[3];
        int empty = 0;
        for (int a = 0; a < 3; ++a) {
          double f = ((double)qp[a] - mn[a]) / cell;
          lo[a] = (int64_t)floor(f) - 1;
          hi[a] = (int64_t)floor(f) + 1;
          if (lo[a] < 0) lo[a] = 0;
          if (hi[a] >= dim[a]) hi[a] = dim[a] - 1;
          if (lo[a] > hi[a]) empty = 1;
        }
        if (!empty)
          for (int64_t cz = lo[2]; cz <= hi[2]; ++cz)
            for (int64_t cy = lo[1]; cy <= hi[1]; ++cy)
              for (int64_t cx = lo[0]; cx <= hi[0]; ++cx) {
                int64_t c = cx + dim[0] * (cy + dim[1] * cz);
                for (int64_t t = cstart[c]; t < cstart[c + 1]; ++t) {
                  int64_t j = order[t];
                  float d = sqdist(qp, sp + 3 * j);
                  if (d < r2) { /* nanoflann.hpp:249-252 strict */
                    if (nc >= cand_cap) {
                      cand_cap = cand_cap * 2 + 256;
                      cand = (cand_t*)realloc(cand, sizeof(cand_t) * (size_t)cand_cap);
                      if (!cand) return -1;
                    }
                    cand[nc].d = d; cand[nc].j = j; ++nc;
                  }
                }
              }
      }
      qsort(cand, (size_t)nc, sizeof(cand_t), cand_cmp);
      for (int64_t k = 0; k < nc; ++k)
        if (rows_push(&R, pos + k, cand[k].j + s0) != 0) return -1; /* :83 batch offset */
      pos += nc;
      R.offs[q0 + i + 1] = pos;
    }
    free(cstart); free(cidx); free(order);
    q0 += nqb; s0 += nsb;
  }
  int64_t w = emit_padded(&R, nq, ns, out_idx, out_cap_width);
  free(R.rows); free(R.offs); free(cand);
  return w;
}

/* Exhaustive O(Nq*Ns) version of the same specification (small inputs only). */
int64_t oracle_radius_neighbors_brute(const float* q, const float* s, const int64_t* q_len, const int64_t* s_len,
                                      int batch, float radius, int64_t* out_idx, int64_t out_cap_width) {
  int64_t nq = 0, ns = 0;
  for (int b = 0; b < batch; ++b) { nq += q_len[b]; ns += s_len[b]; }
  const float r2 = radius * radius;
  rows_t R; R.rows = NULL; R.cap = 0;
  R.offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nq + 1));
  if (!R.offs) return -1;
  R.offs[0] = 0;
  int64_t pos = 0, q0 = 0, s0 = 0;
  for (int b = 0; b < batch; ++b) {
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)(s_len[b] > 0 ? s_len[b] : 1));
    if (!cand) return -1;
    for (int64_t i = 0; i < q_len[b]; ++i) {
      int64_t nc = 0;
      for (int64_t j = 0; j < s_len[b]; ++j) {
        float d = sqdist(q + 3 * (q0 + i), s + 3 * (s0 + j));
        if (d < r2) { cand[nc].d = d; cand[nc].j = j; ++nc; }
      }
      qsort(cand, (size_t)nc, sizeof(cand_t), cand_cmp);
      for (int64_t k = 0; k < nc; ++k)
        if (rows_push(&R, pos + k, cand[k].j + s0) != 0) return -1;
      pos += nc;
      R.offs[q0 + i + 1] = pos;
    }
    free(cand);
    q0 += q_len[b]; s0 += s_len[b];
  }
  int64_t w = emit_padded(&R, nq, ns, out_idx, out_cap_width);
  free(R.rows); free(R.offs);
  return w;
}
