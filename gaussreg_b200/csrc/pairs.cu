// Config 3 / 4 plumbing: ONE neighbour pyramid for P scene pairs, then per-pair views.
//
// The reference's collate stacks a batch as [ref_1 .. ref_P, src_1 .. src_P] (geotransformer/utils/data.py:139-189),
// but its model only accepts batch_size 1 (experiments/.../model.py:77-89).  Building the pyramid once for P pairs
// amortises ~170 latency-bound launches and two host syncs over the batch; the kernels below then re-order every
// per-stage array to pair-major order [ref_1, src_1, ref_2, src_2, ..] and re-base the neighbour indices to the
// pair's own stacked cloud, so that pair i's tensors are plain row slices -- exactly the tensors the single-pair
// collate produces (clouds never see each other's points: the searches are per cloud).
#include "common.cuh"

namespace gr {

struct PairMap {
  const long long* ref_off;  // [P+1] prefix sums of the ref cloud lengths of one stage
  const long long* src_off;  // [P+1] prefix sums of the src cloud lengths
  int P;
};

// pair-major row r -> (pair i, row of the batch-stacked array)
__device__ __forceinline__ void locate(const PairMap& m, long long r, int* pair, long long* stacked_row) {
  int lo = 0, hi = m.P;  // invariant: pm(lo) <= r < pm(hi), pm(i) = ref_off[i] + src_off[i]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (m.ref_off[mid] + m.src_off[mid] <= r) lo = mid; else hi = mid;
  }
  const long long l = r - (m.ref_off[lo] + m.src_off[lo]);
  const long long nr = m.ref_off[lo + 1] - m.ref_off[lo];
  *pair = lo;
  *stacked_row = l < nr ? m.ref_off[lo] + l : m.ref_off[m.P] + m.src_off[lo] + (l - nr);
}

__global__ void __launch_bounds__(256) pair_major_rows_kernel(const float* __restrict__ in, int C, long long n_rows, PairMap q,
                                                              float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * C) return;
  const long long r = t / C;
  int pair;
  long long src;
  locate(q, r, &pair, &src);
  out[t] = in[src * C + (t % C)];
}

// one warp per output row
__global__ void __launch_bounds__(256) pair_major_table_kernel(const long long* __restrict__ in, long long ld, int W, long long n_rows,
                                                               PairMap q, PairMap s, long long* __restrict__ out,
                                                               int* __restrict__ widths) {
  pdl_wait();
  pdl_trigger();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int lane = threadIdx.x & 31;
  int i;
  long long src;
  locate(q, r, &i, &src);
  const long long s_ref_total = s.ref_off[s.P], s_total = s_ref_total + s.src_off[s.P];
  const long long s_nr = s.ref_off[i + 1] - s.ref_off[i], s_ns = s.src_off[i + 1] - s.src_off[i];
  int cnt = 0;
  for (int c = lane; c < W; c += 32) {
    const long long g = in[src * ld + c];
    long long v;
    if (g >= s_total || g < 0) {
      v = s_nr + s_ns;  // the pair's own sentinel
    } else {
      v = g < s_ref_total ? g - s.ref_off[i] : g - s_ref_total - s.src_off[i] + s_nr;
      ++cnt;
    }
    out[r * ld + c] = v;
  }
  cnt = warp_sum(cnt);
  if (lane == 0 && cnt > 0) atomicMax(&widths[i], cnt);
}

}  // namespace gr

using namespace gr;

/* out (n_rows, C) = rows of the batch-stacked `in` ([ref_1..ref_P, src_1..src_P]) in pair-major order
 * ([ref_1, src_1, ref_2, src_2, ..]).  ref_off / src_off: device int64 [P+1] prefix sums of the cloud lengths. */
extern "C" int gr_pair_major_rows(const float* in, int C, int64_t n_rows, const int64_t* ref_off, const int64_t* src_off, int P,
                                  float* out, void* stream) {
  if (C <= 0 || n_rows < 0 || P <= 0 || !ref_off || !src_off) return GR_ERR_BAD_ARG;
  if (n_rows == 0) return GR_OK;
  if (!in || !out) return GR_ERR_BAD_ARG;
  PairMap q{reinterpret_cast<const long long*>(ref_off), reinterpret_cast<const long long*>(src_off), P};
  GR_CHECK_CUDA(launch_pdl(pair_major_rows_kernel, dim3(ceil_div(n_rows * C, 256)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), in, C, (long long)n_rows, q, out));
  GR_CHECK_LAUNCH("pair_major_rows_kernel");
  return GR_OK;
}

/* Neighbour table (n_rows, ld) of a batch-stacked pyramid -> pair-major rows with indices re-based to each pair's own
 * stacked support cloud (sentinel = that pair's support count).  q_*: prefix sums of the QUERY stage, s_*: of the
 * SUPPORT stage.  widths[i] (device int32, zeroed here) = the largest number of real neighbours in any row of pair i,
 * i.e. the width the single-pair search would have reported. */
extern "C" int gr_pair_major_table(const int64_t* in, int64_t ld, int W, int64_t n_rows, const int64_t* q_ref_off,
                                   const int64_t* q_src_off, const int64_t* s_ref_off, const int64_t* s_src_off, int P,
                                   int64_t* out, int32_t* widths, void* stream) {
  if (ld <= 0 || W < 0 || W > ld || n_rows < 0 || P <= 0 || !q_ref_off || !q_src_off || !s_ref_off || !s_src_off || !widths)
    return GR_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GR_CHECK_CUDA(cudaMemsetAsync(widths, 0, (size_t)P * sizeof(int32_t), st));
  if (n_rows == 0 || W == 0) return GR_OK;
  if (!in || !out) return GR_ERR_BAD_ARG;
  PairMap q{reinterpret_cast<const long long*>(q_ref_off), reinterpret_cast<const long long*>(q_src_off), P};
  PairMap s{reinterpret_cast<const long long*>(s_ref_off), reinterpret_cast<const long long*>(s_src_off), P};
  GR_CHECK_CUDA(launch_pdl(pair_major_table_kernel, dim3(ceil_div(n_rows, 8)), dim3(256), (size_t)(0), st, reinterpret_cast<const long long*>(in), (long long)ld, W,
                                                               (long long)n_rows, q, s, reinterpret_cast<long long*>(out), widths));
  GR_CHECK_LAUNCH("pair_major_table_kernel");
  return GR_OK;
}
