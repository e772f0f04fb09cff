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

/* The whole subsampling chain of a pyramid (utils/data.py:24-31: stage i = grid_subsampling(stage i-1, voxel_sizes[i-1]))
 * in ONE call: n_sub dependent gr_grid_subsample calls on `stream`, every stage sized by the capacity n_points.
 *   out_points  (n_sub, n_points, 3) f32;  out_lengths (n_sub, batch) i64;  out_totals (n_sub) i64  -- all DEVICE
 *   stage_events (n_sub + 1 entries, may be NULL): [i] receives a cudaEvent_t recorded on `stream` behind stage i
 *   (i = 1..n_sub; [0] = NULL), owned by the library and valid until this thread's next chain call -- the
 *   `stage_ready_events` of gr_radius_pyramid. */
int gr_grid_subsample_chain(const float* points, const int64_t* lengths, int batch, int64_t n_points, const float* voxel_sizes,
                            int n_sub, float* out_points, int64_t* out_lengths, int64_t* out_totals, void* ws, size_t ws_bytes,
                            void** stage_events, void* stream);

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
/* Same, with reuse_grid != 0 promising that `ws` still holds the cell grid built by an earlier call for the same
 * (s_points, s_lengths, radius): the support cloud is not binned again. */
int gr_radius_neighbors_cached(const float* q_points, const float* s_points, const int64_t* q_lengths,
                               const int64_t* s_lengths, int batch, int64_t nq, int64_t ns, float radius,
                               int64_t* out_idx, int64_t ld, int32_t* out_max_count, void* ws, size_t ws_bytes,
                               int reuse_grid, void* stream);

/* All radius searches of a neighbour pyramid (utils/data.py:27-67: per stage `neighbors`, `subsampling`, `upsampling`) in
 * ONE call: the host spends ~50 us here instead of ~0.6 ms on thirteen Python-issued calls, which is what delayed
 * everything queued behind the pyramid.  Every stage's cloud lives in a capacity-sized buffer (`capacity` rows, true
 * sizes in the device-side stage_lengths); searches[j] is gr_radius_neighbors_cached(query stage, support stage) into
 * its own (capacity, limit) table, one cell grid per support stage (stage_grid_ws[s], each grid_ws_bytes >=
 * gr_radius_neighbors_workspace_size(capacity, capacity, batch)) shared by the searches of that stage in call order.
 * stage_ready_events[s] (cudaEvent_t, may be NULL) is waited on `stream` before the first search that touches stage s:
 * the grid-subsampling chain may run on another stream.  Bit s of built_mask: stage_grid_ws[s] already holds the grid of
 * stage s (an earlier call of this function searched that stage). */
typedef struct {
  int query_stage, support_stage;
  float radius;
  int64_t limit;
  int64_t* out_idx;
  int32_t* out_max_count;
} gr_pyramid_search;
int gr_radius_pyramid(const float* const* stage_points, const int64_t* const* stage_lengths, int n_stages, int batch,
                      int64_t capacity, void* const* stage_grid_ws, size_t grid_ws_bytes, void* const* stage_ready_events,
                      const gr_pyramid_search* searches, int n_searches, uint32_t built_mask, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense fp32 product with fused epilogue (used by K1 contraction, K2 Linear, T1-T3 projections,
 * attention products, M1/M2 similarity).  Replaces the ATen/cuBLAS calls behind nn.Linear
 * (kpconv/modules.py:73,98), torch.matmul (kpconv.py:105,109), torch.einsum (model.py:189,
 * rpe_transformer.py:56-57, vanilla_transformer.py:55).
 *   C[b] = act(alpha * A[b] . op(B[b]) / row_div[:,None] + bias[None,:] + residual[b])
 *   A (M,K) ld lda; B (N,K) if trans_b else (K,N); batch strides in elements; act 0 none 1 relu 2 leaky(0.1).
 * --------------------------------------------------------------------------------------------- */
int gr_gemm(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB, int trans_b,
            float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch, float alpha, const float* bias,
            const float* row_div, const float* residual, int64_t ldr, int64_t strideR, int act, void* stream);
/* 0 = fp32 FFMA kernels only; 1 (default) = tcgen05 kind::tf32 with the 3xTF32 split for large trans_b products
 * (env GAUSSREG_GEMM=simt|tc selects the initial mode). */
/* Same product for a static weight W (N,K) whose gr_pack_weight_tf32x3 image is supplied: the weight tiles reach
 * shared memory by bulk async copies (no conversion work in the kernel). */
int gr_linear_packed(const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_packed, float* C, int64_t ldc,
                     int M, int N, int K, float alpha, const float* bias, const float* row_div, const float* residual,
                     int64_t ldr, int act, void* stream);
int gr_linear_packed16(const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_packed, const void* W_packed16,
                       float inv_scale16, float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias,
                       const float* row_div, const float* residual, int64_t ldr, int act, void* stream);
void gr_set_gemm_mode(int mode);
int gr_get_gemm_mode(void);
/* 1 if the calling thread's last gr_gemm ran on the tensor cores, 0 if on the FFMA kernel (bench bookkeeping). */
int gr_last_gemm_path(void);

/* ---------------------------------------------------------------------------------------------
 * K1  KPConv gather + kernel-point correlation (geotransformer/modules/kpconv/kpconv.py:79-122).
 *   A[m, k*C + c] = sum_h max(0, 1 - |(s_h - q_m) - kp_k| / sigma) * s_feats[idx[m,h], c]
 *   row_div[m]    = max(1, #{h : sum_c s_feats[idx[m,h], c] > 0})
 * The contraction with weights (15*C, C_out) and the /row_div + bias epilogue are one gr_gemm call.
 * --------------------------------------------------------------------------------------------- */
size_t gr_kpconv_aggregate_workspace_size(int64_t n_support);
int gr_kpconv_aggregate(const float* s_feats, int C, const float* q_points, const float* s_points,
                        const int64_t* neighbor_idx, int H, int64_t ld_idx, int M, int Ns, const float* kernel_points,
                        int n_kernel_points, float sigma, float* A, float* row_div, void* ws, size_t ws_bytes,
                        void* stream);

/* K2  GroupNorm over all rows of the stacked pair (kpconv/modules.py:33-50) + optional add + activation. */
size_t gr_group_norm_workspace_size(int64_t n_rows, int groups);
int gr_group_norm(const float* x, int64_t n_rows, int C, int groups, const float* gamma, const float* beta, float eps,
                  const float* add, int act, float* y, void* ws, size_t ws_bytes, void* stream);
/* T2/T3  y = LayerNorm(a + b) (rpe_transformer.py:101-103, output_layer.py:14-21); b may be NULL. */
int gr_layer_norm_add(const float* a, const float* b, int64_t rows, int C, const float* gamma, const float* beta,
                      float eps, float* y, void* stream);

/* K3  neighbourhood max-pool (kpconv/functional.py:54-67), nearest-upsample + skip concat
 * (functional.py:6-22, backbone.py:195-208), zero-padded row gather (modules/ops/index_select.py). */
int gr_maxpool(const float* x, int Ns, int C, const int64_t* idx, int H, int64_t ld_idx, int M, float* out, void* stream);
int gr_upsample_concat(const float* coarse, int Nc, int C1, const int64_t* idx, int64_t ld_idx, const float* skip,
                       int C2, int M, float* out, void* stream);
int gr_gather_rows(const float* x, int n, int C, const int64_t* idx, int64_t rows, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2/K4  KPConv blocks and the whole KPConvFPN as single entry points
 * (geotransformer/modules/kpconv/modules.py:53-225, experiments/.../backbone.py:164-212).
 *
 * The host structs below hold DEVICE pointers to the module parameters (state_dict tensors) plus derived
 * operand images the caller prepares once per parameter version:
 *   weight_packed          gr_pack_weight_tf32x3 image of `weight`            (may be NULL: converted on the fly)
 *   weights_kmajor         (out, 15*in) K-major copy of KPConv.weights        (may be NULL: NN product on `weights`)
 *   weights_kmajor_packed  gr_pack_weight_tf32x3 image of weights_kmajor      (may be NULL)
 * GroupNorm statistics of every Linear / KPConv output are produced by the product's own epilogue (per-tile
 * partial sums folded in a fixed order), so a block costs no separate statistics pass over its activations.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  const float* weight;        /* nn.Linear weight (out, in) */
  const float* weight_packed;
  const float* bias;          /* (out) or NULL */
  const float* gn_weight;     /* GroupNorm affine (out); NULL = no norm (LastUnaryBlock) */
  const float* gn_bias;
  int in_channels, out_channels; /* in_channels == 0: the block is nn.Identity */
  int leaky_relu;             /* LeakyReLU(0.1) after the norm (UnaryBlock has_relu) */
  /* Decoder blocks only (their input is cat[nearest_upsample(coarse), skip], backbone.py:195-208): with split_k = C_coarse
   * and the packed images of weight[:, :split_k] / weight[:, split_k:], the product is evaluated as
   * upsample(coarse W_lo^T) + skip W_hi^T -- a row gather commutes with the right-multiplication, so the coarse half
   * runs on the 4x fewer coarse rows and no (M, C_coarse + C_skip) tensor is written.  split_k = 0: plain product. */
  int split_k;
  const float* weight_packed_lo;
  const float* weight_packed_hi;
  /* Optional fp16-split images (gr_pack_weight_f16x3) of weight / weight[:, :split_k] / weight[:, split_k:] and the inverses
   * of their packing scales: with them the products run on kind::f16 (K = 16 per tensor-core dispatch instead of 8). */
  const void* weight_packed16;
  const void* weight_packed16_lo;
  const void* weight_packed16_hi;
  float inv_scale16, inv_scale16_lo, inv_scale16_hi;
} gr_unary_weights;

typedef struct {
  const float* weights;       /* (15, in, out) */
  const float* weights_kmajor;
  const float* weights_kmajor_packed;
  const float* bias;          /* (out) or NULL */
  const float* kernel_points; /* (15, 3) */
  float sigma;
  int in_channels, out_channels;
  const void* weights_kmajor_packed16; /* gr_pack_weight_f16x3 image of weights_kmajor (may be NULL) */
  float inv_scale16;
} gr_kpconv_weights;

typedef struct {
  int kind;                   /* 0 = ConvBlock (KPConv + GN + LeakyReLU), 1 = ResidualBlock */
  int strided;                /* ResidualBlock: shortcut = maxpool over the neighbour table */
  gr_unary_weights unary1;    /* ResidualBlock only */
  gr_kpconv_weights conv;
  const float* gn_conv_weight; /* GroupNorm after the KPConv */
  const float* gn_conv_bias;
  gr_unary_weights unary2;
  gr_unary_weights shortcut;
} gr_block_weights;

#define GR_FPN_STAGES 5
#define GR_FPN_BLOCKS 14
typedef struct {
  gr_block_weights blocks[GR_FPN_BLOCKS]; /* encoder1_1 ... encoder5_3 in forward order (backbone.py:164-193) */
  gr_unary_weights decoder4, decoder3, decoder2;
  int group_norm;                          /* number of groups */
  float eps;
} gr_fpn_weights;

typedef struct {
  const float* points[GR_FPN_STAGES];
  int n_points[GR_FPN_STAGES];
  const int64_t* neighbors[GR_FPN_STAGES];   int neighbors_w[GR_FPN_STAGES];   int64_t neighbors_ld[GR_FPN_STAGES];
  const int64_t* subsampling[GR_FPN_STAGES]; int subsampling_w[GR_FPN_STAGES]; int64_t subsampling_ld[GR_FPN_STAGES];
  const int64_t* upsampling[GR_FPN_STAGES];  int upsampling_w[GR_FPN_STAGES];  int64_t upsampling_ld[GR_FPN_STAGES];
} gr_pyramid;

/* y = [LeakyReLU]( GroupNorm(x W^T + b) [+ add] ), or x W^T + b when w->gn_weight is NULL.  `act_after_add`
 * (0 none / 2 leaky) applies when the block itself has no activation (modules.py:207-225). */
size_t gr_unary_block_workspace_size(int64_t rows, int out_channels, int groups);
int gr_unary_block(const gr_unary_weights* h_w, const float* x, int64_t rows, int groups, float eps, const float* add,
                   int act_after_add, float* y, void* ws, size_t ws_bytes, void* stream);
/* y = LeakyReLU(GroupNorm(KPConv(...))) when gn_weight is given, the raw KPConv output otherwise. */
size_t gr_kpconv_block_workspace_size(int M, int Ns, int in_channels, int out_channels, int groups);
int gr_kpconv_block(const gr_kpconv_weights* h_w, const float* gn_weight, const float* gn_bias, int groups, float eps,
                    const float* s_feats, const float* q_points, const float* s_points, const int64_t* neighbor_idx, int H,
                    int64_t ld_idx, int M, int Ns, float* y, void* ws, size_t ws_bytes, void* stream);
/* KPConvFPN.forward: feats (n_points[0], in) -> out_l2 (n1, C2), out_l3 (n2, C3), out_l4 (n3, C4), out_f5 (n4, C5)
 * = the list backbone.py:210-212 returns.  Every kernel is issued from this one call. */
size_t gr_kpconv_fpn_workspace_size(const gr_fpn_weights* h_w, const gr_pyramid* h_pyr);
int gr_kpconv_fpn(const gr_fpn_weights* h_w, const gr_pyramid* h_pyr, const float* feats, float* out_l2, float* out_l3,
                  float* out_l4, float* out_f5, void* ws, size_t ws_bytes, void* stream);
/* gr_kpconv_fpn entering at block `start_block` (0 or 2): with 2, `feats` is the output of encoder1_2 -- the two stage-0
 * blocks (gr_kpconv_block) need only the input cloud and neighbors[0] and may have been queued while the rest of the
 * pyramid was still being built (backbone.py:166-167 vs :168-212). */
int gr_kpconv_fpn_from(const gr_fpn_weights* w, const gr_pyramid* pyr, int start_block, const float* feats, float* out_l2,
                       float* out_l3, float* out_l4, float* out_f5, void* ws, size_t ws_bytes, void* stream);

/* N1  farthest-point subsampling to `point_limit` (demo.py:44-47; restates exact FPS, the result the third-party
 * fpsample.bucket_fps_kdline_sampling accelerates).  points (n,3) -> out_idx (k) i64 in selection order. */
size_t gr_farthest_point_sample_workspace_size(int64_t n_points);
int gr_farthest_point_sample(const float* points, int64_t n_points, int k, int64_t start_idx, int64_t* out_idx, void* ws,
                             size_t ws_bytes, void* stream);

/* N4  Gaussian merge (gs_fusion.py:231-262): the second 3DGS cloud under the estimated similarity transform -- positions,
 * log-scales, rotation quaternions (quaternion_to_matrix / matrix_to_quaternion, :70-159) and SH bands 1-3 (sh_rotation,
 * :53-68).  cloud / out: (n,59) f32 device rows with pitches ld_in / ld_out; h_rotation: HOST 3x3 unit rotation (row-major);
 * log_scale = log(scale); h_translation: HOST float[3]; h_sh: HOST double[9 + 25 + 49] band matrices. */
int gr_gaussian_transform(const float* cloud, int64_t ld_in, int64_t n, const float* h_rotation, float scale, float log_scale,
                          const float* h_translation, const double* h_sh, float* out, int64_t ld_out, void* stream);

/* N3  similarity-transform RANSAC over the LGR correspondences (model.py:209-215, utils/open3d.py:169-198; restates the
 * published algorithm of open3d==0.11.2 registration_ransac_based_on_correspondence with
 * TransformationEstimationPointToPoint(with_scaling=True): third party, randomised -> statistical parity only).
 * ref_corr / src_corr (capacity,3); d_num_corr: DEVICE int32 count of valid rows (NULL = capacity); fallback: device (4,4)
 * returned when no hypothesis has an inlier (NULL = identity); T_out (4,4) = [c R | t]; info[2] (may be NULL) = {inliers of
 * the best hypothesis, its index}.  Deterministic for a given seed (counter-based sampler). */
size_t gr_similarity_ransac_workspace_size(int num_hypotheses);
int gr_similarity_ransac(const float* ref_corr, const float* src_corr, const int32_t* d_num_corr, int capacity,
                         int num_hypotheses, int sample_size, float distance_threshold, uint64_t seed, int refit,
                         const float* fallback, float* T_out, int32_t* info, void* ws, size_t ws_bytes, void* stream);

/* P1  point-to-node partition (modules/ops/pointcloud_partition.py:61-111). */
size_t gr_point_to_node_workspace_size(int64_t n_points, int64_t n_nodes);
int gr_point_to_node_partition(const float* points, int N, const float* nodes, int M, int point_limit,
                               int32_t* point_to_node, uint8_t* node_masks, int64_t* knn_indices, uint8_t* knn_masks,
                               void* ws, size_t ws_bytes, void* stream);

/* T1  geometric structure embedding pieces (modules/geotransformer/geotransformer.py:26-72,
 * modules/transformer/positional_embedding.py:19-35). */
int gr_embedding_indices(const float* points, int N, float sigma_d, float sigma_a, int angle_k, float* d_idx, float* a_idx,
                         int32_t* knn, void* stream);
int gr_sinusoid_rows(const float* x, int64_t rows, const float* div_term, int n_div, float* E, void* stream);
int gr_embedding_combine(const float* D, const float* A, int64_t rows, int C, int k, float* out, void* stream);
/* T1 fused on the tensor cores (geotransformer.py:57-72 in one kernel): sinusoid operand tiles are generated in
 * shared memory, proj_d / proj_a run as tcgen05 3xTF32 MMAs into four TMEM accumulators, bias + max_k + sum in the
 * epilogue.  Weights must first be packed with gr_pack_weight_tf32x3 ((N,K) -> 2*roundup(N,256)*roundup(K,32) floats). */
int gr_pack_weight_tf32x3(const float* W, int N, int K, float* out, void* stream);
int gr_structure_embedding_fused(const float* d_idx, const float* a_idx, int64_t rows, int angle_k, const float* div_term,
                                 int hidden_dim, const float* wd_packed, const float* wa_packed, const float* bias_d,
                                 const float* bias_a, float* out, void* stream);

/* T1 with fp16-split operands: TF32 and fp16 carry the same 11 significant bits, so the two-part split is as accurate,
 * and kind::f16 issues at twice the kind::tf32 rate.  Valid here because the generated operand is sin / cos in [-1, 1]
 * and the static weights are pre-scaled by a power of two (divided out in the epilogue).
 * gr_pack_weight_f16x2: W (N <= 256, K <= 256) fp32 -> 256 KB of fp16 hi/lo tiles; gr_structure_embedding_fused_f16:
 * same contract as gr_structure_embedding_fused, inv_scale_* = 1 / the packing scales.
 * Replaces geotransformer/modules/geotransformer/geotransformer.py:57-72. */
size_t gr_packed_weight_f16_bytes(int N, int K);
int gr_pack_weight_f16x3(const float* W, int N, int K, float scale, void* out, void* stream);
int gr_gemm_f16_overflow_ptr(int** dev_ptr);
int gr_pack_weight_f16x2(const float* W, int N, int K, float scale, void* out, void* stream);
int gr_structure_embedding_fused_f16(const float* d_idx, const float* a_idx, int64_t rows, int angle_k, const float* div_term,
                                     int hidden_dim, const void* wd_packed, const void* wa_packed, float inv_scale_d,
                                     float inv_scale_a, const float* bias_d, const float* bias_a, float* out, void* stream);

/* T1 by tabulation.  proj(sinusoid(x)) is a band-limited function of ONE scalar per channel (highest angular frequency 1),
 * so geotransformer.py:57-72 = f_d(d) + max_k f_a(a_k) with f_d, f_a interpolated from exact fp64 node tables (cubic
 * Hermite, step 1/8 on the angle index range [0, 180/sigma_a]; quintic Hermite, step 1/2 on the distance index range
 * [0, 1024]: the first 129 nodes in shared memory, the rest read from L2; an index outside its table is evaluated
 * directly from the weights).  Error vs the exact function 7e-8
 * relative (the fp32 reference itself: 3e-7).  gr_structure_embedding_table_floats: table size (0 on bad arguments);
 * gr_structure_embedding_build_table: once per weight set; gr_structure_embedding_tabulated: same contract as
 * gr_structure_embedding_fused with raw (hidden_dim, hidden_dim) weights; hidden_dim % 64 == 0.
 * Replaces geotransformer/modules/geotransformer/geotransformer.py:57-72. */
int64_t gr_structure_embedding_table_floats(int hidden_dim, float sigma_a);
int gr_structure_embedding_build_table(const float* div_term, int hidden_dim, const float* W_d, const float* b_d,
                                       const float* W_a, const float* b_a, float sigma_a, float* table, void* stream);
int gr_structure_embedding_tabulated(const float* d_idx, const float* a_idx, int64_t rows, int angle_k, const float* table,
                                     float sigma_a, const float* div_term, int hidden_dim, const float* W_d, const float* b_d,
                                     const float* W_a, const float* b_a, float* out, void* stream);

/* gr_embedding_indices + gr_structure_embedding_tabulated in one call (d_idx, a_idx, knn: the intermediate buffers). */
int gr_structure_embedding_points(const float* points, int N, float sigma_d, float sigma_a, int angle_k, const float* table,
                                  const float* div_term, int hidden_dim, const float* W_d, const float* b_d, const float* W_a,
                                  const float* b_a, float* d_idx, float* a_idx, int32_t* knn, float* out, void* stream);

/* T2  RPE attention probabilities with the p-term reassociated (rpe_transformer.py:50-66), row softmax
 * (vanilla_transformer.py:66), F.normalize (model.py:143-144). */
int gr_rpe_attention_probs(const float* q, const float* k, const float* U, const float* qb, const float* emb, int N, int C,
                           int num_heads, float* P, void* stream);
/* same, q and k being column slices of wider matrices (row pitches ldq / ldk, e.g. a fused q|k|v projection) */
int gr_rpe_attention_probs_ld(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* U, const float* qb,
                              const float* emb, int N, int C, int num_heads, float* P, void* stream);
int gr_softmax_rows(float* x, int64_t rows, int cols, void* stream);
int gr_l2_normalize_rows(const float* x, int64_t rows, int C, float eps, float* y, void* stream);

/* T4  the whole RPE-conditional transformer (conditional_transformer.py:97-117: 'self' = RPE self-attention layer on
 * both clouds, 'cross' = sequential cross-attention layers) in one call: host-side orchestration of the kernels
 * above, so that the ~130 superpoint-sized launches are issued from C++ rather than from the interpreter.
 * All pointers are device pointers to the nn.Linear / nn.LayerNorm parameters of one layer. */
typedef struct {
  const float *wq, *bq, *wk, *bk, *wv, *bv;  /* attention.attention.proj_{q,k,v} (C,C),(C) */
  const float *wp, *bp;                      /* attention.attention.proj_p (self layers only, else NULL) */
  const float *wo, *bo;                      /* attention.linear */
  const float *ln1_g, *ln1_b;                /* attention.norm */
  const float *w1, *b1, *w2, *b2;            /* output.expand (2C,C), output.squeeze (C,2C) */
  const float *ln2_g, *ln2_b;                /* output.norm */
  /* optional: proj_q|proj_k|proj_v stacked row-wise, (3C,C) and (3C).  When set, q/k/v come out of one product
   * (self layers) or of a q product and a fused k|v product (cross layers); results are identical, there are
   * just fewer superpoint-sized launches.  NULL -> the separate matrices above are used. */
  const float *wqkv, *bqkv;
  /* optional: output.expand / output.squeeze weights transposed to K-major, (C,2C) and (2C,C), for the fused
   * AttentionOutput kernel (GAUSSREG_TF_MLP=1); NULL -> the two GEMMs + LayerNorm path. */
  const float *w1t, *w2t;
  int is_self;
} gr_layer_weights;
size_t gr_conditional_transformer_workspace_size(int N0, int N1, int C, int num_heads);
int gr_conditional_transformer(const gr_layer_weights* h_layers, int n_layers, float* feats0, float* feats1,
                               const float* emb0, const float* emb1, int N0, int N1, int C, int num_heads, void* ws,
                               size_t ws_bytes, void* stream);

/* M1  superpoint matching (modules/geotransformer/superpoint_matching.py:13-50). */
size_t gr_superpoint_matching_workspace_size(int Nr, int Ns, int k);
int gr_superpoint_matching(float* xy, int Nr, int Ns, const uint8_t* ref_masks, const uint8_t* src_masks, int k,
                           int dual_normalization, int64_t* ref_idx, int64_t* src_idx, float* scores, int32_t* count,
                           void* ws, size_t ws_bytes, void* stream);

/* S1  log-domain Sinkhorn with learnable dustbin (modules/sinkhorn/learnable_sinkhorn.py:5-66). */
int gr_sinkhorn(const float* scores, const uint8_t* row_masks, const uint8_t* col_masks, const float* alpha, int P, int K,
                int num_iterations, float inf, float* out, void* stream);

/* L1  local-to-global registration (modules/geotransformer/local_global_registration.py:11-235) and
 * L2  weighted Procrustes (modules/registration/procrustes.py:6-82), SVD on the device. */
size_t gr_lgr_workspace_size(int P, int K, int topk);
int gr_local_global_registration(const float* matching_scores, int P, int K, int ld, const float* ref_knn_points,
                                 const float* src_knn_points, const uint8_t* ref_knn_masks, const uint8_t* src_knn_masks,
                                 int topk, float acceptance_radius, int mutual, float confidence_threshold,
                                 int correspondence_threshold, int num_refinement_steps, float* ref_corr_points,
                                 float* src_corr_points, float* corr_scores, int32_t* num_corr, float* transform, void* ws,
                                 size_t ws_bytes, void* stream);
int gr_weighted_procrustes(const float* src_points, const float* ref_points, const float* weights, int B, int n, float eps,
                           float* transforms, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N1  Gaussian-splat cloud -> network input, the step in front of the path
 *     (experiments/geotransformer.gaussian_splatting.indoor/demo.py:30-75 _read_ply_by_opacity and :81-124
 *     load_data; geotransformer/utils/graphics_utils.py:34-89 eval_sh).
 * cloud: (n, ld) f32 rows in 3DGS property order without normals (gs_fusion.py:172-184):
 *     xyz 0..2 | f_dc 3..5 | f_rest 6..50 | opacity 51 | scale 52..54 | rot 55..58.
 * Arguments documented as HOST are read before the call returns; everything else is device memory.
 * --------------------------------------------------------------------------------------------- */
/* out_values[j] (device) = ranks[j]-th smallest (0-based) entry of column cols[j]; cols/ranks HOST, n_queries <= 16.
 * Exact order statistics for np.percentile (demo.py:40-42); the interpolation itself is host glue. */
size_t gr_column_order_stats_workspace_size(int n_queries);
int gr_column_order_stats(const float* cloud, int64_t n, int ld, const int32_t* cols, const int64_t* ranks, int n_queries,
                          float* out_values, void* ws, size_t ws_bytes, void* stream);
/* keep row i iff sigmoid(opacity) > opacity_min and lo[a] < xyz[a] < hi[a] (demo.py:34,40-43); lo/hi HOST double[3].
 * out_index (capacity n) holds the kept rows in ascending order, out_count their number. */
size_t gr_gaussian_select_workspace_size(int64_t n);
int gr_gaussian_select(const float* cloud, int64_t n, int ld, int opacity_col, float opacity_min, const double* lo,
                       const double* hi, int64_t* out_index, int64_t* out_count, void* ws, size_t ws_bytes, void* stream);
/* out_points (m,3) = cloud[index, 0:3] (index NULL = all rows); out_stats float[9] = {float32 column sums in input
 * order (numpy's axis-0 reduction order, demo.py:62), min xyz, max xyz}.  ws: >= 64 bytes. */
int gr_gather_points_stats(const float* cloud, int ld, const int64_t* index, int64_t m, float* out_points,
                           float* out_stats, void* ws, size_t ws_bytes, void* stream);
/* out_feats (m,4) = [sigmoid(opacity), 255*clip(SH_deg3(view dir)+0.5,0,1) RGB] (demo.py:63-72); view_point HOST double[3]. */
int gr_gaussian_features(const float* cloud, int ld, const int64_t* index, int64_t m, const double* view_point,
                         float* out_feats, void* stream);
/* points <- (points - center) [* scale] in float32 (demo.py:85-110); center3 HOST float[3]. */
int gr_points_normalize(float* points, int64_t m, const float* center3, float scale, int apply_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Config 3 / 4: one neighbour pyramid for P pairs (utils/data.py:139-189 stacks [ref_1..ref_P, src_1..src_P]; the
 * reference model is batch-1, model.py:77-89), then pair-major re-ordering so that every pair is a row slice.
 * ref_off / src_off: device int64 [P+1] prefix sums of the per-cloud lengths of one pyramid stage.
 * --------------------------------------------------------------------------------------------- */
int gr_pair_major_rows(const float* in, int C, int64_t n_rows, const int64_t* ref_off, const int64_t* src_off, int P,
                       float* out, void* stream);
int gr_pair_major_table(const int64_t* in, int64_t ld, int W, int64_t n_rows, const int64_t* q_ref_off,
                        const int64_t* q_src_off, const int64_t* s_ref_off, const int64_t* s_src_off, int P, int64_t* out,
                        int32_t* widths, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAUSSREG_B200_H_ */
