"""GeoTransformer coarse-registration model, eval forward
(reference: experiments/geotransformer.gaussian_splatting.indoor/model.py:19-227).

Differences from the reference, by design:
  * forward-only (the `if self.training` branches of model.py:74,111,165 are out of scope);
  * `estimated_transform` is the deterministic LocalGlobalRegistration result by default (the parity target).  The
    reference then overwrites it with Open3D's similarity RANSAC (model.py:209-215: third-party, randomised, CPU);
    `model.ransac = True` (or `create_model(cfg, ransac=True)`) applies the device restatement of that step
    (csrc/ransac.cu: seeded, statistical parity only) and returns it as `estimated_transform`, keeping the LGR
    result under `lgr_transform`;
  * the random init below consumes the torch / numpy RNG streams exactly like the reference, so
    `torch.manual_seed(s); np.random.seed(s); create_model(cfg)` yields the reference's weights
    (checked against tests/golden/weights_checksum.npz).
"""
import threading

import torch
import torch.nn as nn

from . import ops
from .modules import (KPConvFPN, GeometricTransformer, SuperPointMatching, LocalGlobalRegistration,
                      LearnableLogOptimalTransport)


class GeoTransformer(nn.Module):
    def __init__(self, cfg, ransac=False, ransac_seed=0):
        super().__init__()
        # N3: similarity RANSAC after LGR (model.py:209-215: ransac_n 5, 10 000 iterations, threshold 0.05 in eval)
        self.ransac, self.ransac_seed = bool(ransac), int(ransac_seed)
        self.ransac_n, self.ransac_iterations, self.ransac_distance = 5, 10000, 0.05
        self.num_points_in_patch = cfg.model.num_points_in_patch
        self.matching_radius = cfg.model.ground_truth_matching_radius
        b = cfg.backbone
        self.backbone = KPConvFPN(b.input_dim, b.output_dim, b.init_dim, b.kernel_size, b.init_radius, b.init_sigma, b.group_norm)
        g = cfg.geotransformer
        self.transformer = GeometricTransformer(g.input_dim, g.output_dim, g.hidden_dim, g.num_heads, g.blocks, g.sigma_d,
                                                g.sigma_a, g.angle_k, reduction_a=g.reduction_a)
        cm = cfg.coarse_matching
        self.coarse_matching = SuperPointMatching(cm.num_correspondences, cm.dual_normalization)
        f = cfg.fine_matching
        self.fine_matching = LocalGlobalRegistration(
            f.topk, f.acceptance_radius, mutual=f.mutual, confidence_threshold=f.confidence_threshold,
            use_dustbin=f.use_dustbin, use_global_score=f.use_global_score,
            correspondence_threshold=f.correspondence_threshold, correspondence_limit=f.correspondence_limit,
            num_refinement_steps=f.num_refinement_steps)
        self.optimal_transport = LearnableLogOptimalTransport(cfg.model.num_sinkhorn_iterations)
        self._count_slots = {}

    @torch.no_grad()
    def forward(self, data_dict):
        if self.training:
            raise RuntimeError("gaussreg_b200 implements the eval forward only: call model.eval()")
        out = {}
        feats = data_dict["features"]
        lengths = data_dict["lengths"]
        # stage lengths are needed on the host to slice ref / src (model.py:77-79 does three .item() syncs)
        host = data_dict.get("lengths_host")  # extension: stage lengths already known on the host (batched pyramid)
        if host is not None:
            ref_length_c, ref_length_f, ref_length = int(host[-1][0]), int(host[1][0]), int(host[0][0])
        else:
            lens = torch.stack([lengths[-1], lengths[1], lengths[0]]).cpu()
            ref_length_c, ref_length_f, ref_length = int(lens[0, 0]), int(lens[1, 0]), int(lens[2, 0])
        points_c, points_f, points = data_dict["points"][-1], data_dict["points"][1], data_dict["points"][0]
        ref_points_c, src_points_c = points_c[:ref_length_c], points_c[ref_length_c:]

        # 3a. geometric structure embedding (geotransformer.py:57-72) depends on the superpoint coordinates only: it is
        # queued BEFORE the backbone (and before any other host work of this function: the GPU is idle since the stage-size
        # read), so that the GPU is busy while the host reads the neighbour-table widths (data.LazyTables) and prepares the
        # backbone call
        ref_emb = self.transformer.embedding(ref_points_c)
        src_emb = self.transformer.embedding(src_points_c)

        ref_points_f, src_points_f = points_f[:ref_length_f], points_f[ref_length_f:]
        out["ref_points_c"], out["src_points_c"] = ref_points_c, src_points_c
        out["ref_points_f"], out["src_points_f"] = ref_points_f, src_points_f
        out["ref_points"], out["src_points"] = points[:ref_length], points[ref_length:]

        # 2. KPConv FPN (model.py:129-132)
        feats_list = self.backbone(feats, data_dict)
        feats_c, feats_f = feats_list[-1], feats_list[0]

        # 1. point-to-node partition (model.py:99-109).  Its results are first needed by the superpoint matching, so it is
        # issued AFTER the backbone: between the stage sizes and the backbone call the host is the bottleneck (the GPU
        # finishes the embedding before the host has trimmed the tables and prepared the call), and every wrapper call
        # issued there delays the backbone by its host time
        K = self.num_points_in_patch
        _, ref_node_masks, ref_node_knn_indices, ref_node_knn_masks = ops.point_to_node_partition(ref_points_f, ref_points_c, K)
        _, src_node_masks, src_node_knn_indices, src_node_knn_masks = ops.point_to_node_partition(src_points_f, src_points_c, K)

        # 3b. geometric transformer (model.py:135-147)
        ref_feats_c, src_feats_c = self.transformer(ref_points_c, src_points_c, feats_c[:ref_length_c], feats_c[ref_length_c:],
                                                    embeddings=(ref_emb, src_emb))
        del ref_emb, src_emb
        both = ops.stacked_rows(ref_feats_c, src_feats_c)
        if both is not None:
            both = ops.l2_normalize_rows(both)
            ref_feats_c_norm, src_feats_c_norm = both[:ref_feats_c.shape[0]], both[ref_feats_c.shape[0]:]
        else:
            ref_feats_c_norm = ops.l2_normalize_rows(ref_feats_c)
            src_feats_c_norm = ops.l2_normalize_rows(src_feats_c)
        out["ref_feats_c"], out["src_feats_c"] = ref_feats_c_norm, src_feats_c_norm
        ref_feats_f, src_feats_f = feats_f[:ref_length_f], feats_f[ref_length_f:]
        out["ref_feats_f"], out["src_feats_f"] = ref_feats_f, src_feats_f

        # 6. superpoint correspondences (model.py:156-163).  The reference returns min(k, #valid pairs) of them -- a
        # data-dependent length that costs a device->host round trip right where the GPU would run dry (0.2 ms of
        # bubbles: the host can only queue the patch gathers, Sinkhorn and LGR after it).  Almost always the length is k:
        # the tail is queued for k pairs at once, the count follows through pinned memory, and only if it turns out
        # smaller is the tail redone on the shorter arrays.
        ref_ci, src_ci, node_corr_scores, count = self.coarse_matching(ref_feats_c_norm, src_feats_c_norm, ref_node_masks,
                                                                       src_node_masks, lazy=True)
        count_host, count_ready = self._count_slot(feats.device)
        count_host.copy_(count, non_blocking=True)
        count_ready.record()
        self._tail(out, ref_ci, src_ci, node_corr_scores, ref_node_knn_indices, src_node_knn_indices, ref_points_f, src_points_f,
                   ref_feats_f, src_feats_f, feats_f.shape[1], feats.device)
        count_ready.synchronize()  # completed long ago: the host has been queueing the tail meanwhile
        c = int(count_host[0])
        if c < ref_ci.shape[0]:
            self._tail(out, ref_ci[:c], src_ci[:c], node_corr_scores[:c], ref_node_knn_indices, src_node_knn_indices, ref_points_f,
                       src_points_f, ref_feats_f, src_feats_f, feats_f.shape[1], feats.device)
        return out

    def _count_slot(self, dev):
        """A pinned int32 + event per (device, stream, thread): the asynchronous read-back of the correspondence count."""
        key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, threading.get_ident())
        slot = self._count_slots.get(key)
        if slot is None:
            slot = (torch.empty((1,), dtype=torch.int32, pin_memory=True), torch.cuda.Event())
            self._count_slots[key] = slot
        return slot

    def _tail(self, out, ref_ci, src_ci, node_corr_scores, ref_node_knn_indices, src_node_knn_indices, ref_points_f, src_points_f,
              ref_feats_f, src_feats_f, C, dev):
        K = self.num_points_in_patch
        out["ref_node_corr_indices"], out["src_node_corr_indices"] = ref_ci, src_ci
        out["node_corr_scores"] = node_corr_scores

        # 7.2 patch gathers (model.py:171-186)
        P = ref_ci.shape[0]
        ref_knn_idx = ops_gather_index(ref_node_knn_indices, ref_ci)
        src_knn_idx = ops_gather_index(src_node_knn_indices, src_ci)
        ref_knn_masks = ref_knn_idx != ref_points_f.shape[0]
        src_knn_masks = src_knn_idx != src_points_f.shape[0]
        ref_knn_points = ops.gather_rows(ref_points_f, ref_knn_idx)
        src_knn_points = ops.gather_rows(src_points_f, src_knn_idx)
        ref_knn_feats = ops.gather_rows(ref_feats_f, ref_knn_idx)
        src_knn_feats = ops.gather_rows(src_feats_f, src_knn_idx)
        out["ref_node_corr_knn_points"], out["src_node_corr_knn_points"] = ref_knn_points, src_knn_points
        out["ref_node_corr_knn_masks"], out["src_node_corr_knn_masks"] = ref_knn_masks, src_knn_masks

        # 8. optimal transport (model.py:189-193)
        scores = torch.empty((P, K, K), dtype=torch.float32, device=dev)
        ops.gemm_batched(ref_knn_feats.data_ptr(), C, K * C, src_knn_feats.data_ptr(), C, K * C, True, scores.data_ptr(), K,
                         K * K, K, K, C, P, alpha=1.0 / C ** 0.5)
        matching_scores = self.optimal_transport(scores, ref_knn_masks, src_knn_masks)
        out["matching_scores"] = matching_scores

        # 9. local-to-global registration (model.py:196-207); the dustbin row/col is skipped inside the kernel
        if self.ransac:
            # the correspondence buffers stay padded on the device: the RANSAC kernels read the count there (no sync)
            ref_pad, src_pad, sc_pad, num, T = self.fine_matching.forward_device(ref_knn_points, src_knn_points, ref_knn_masks,
                                                                                 src_knn_masks, matching_scores)
            T_sim, info = ops.similarity_ransac(ref_pad, src_pad, num, self.ransac_iterations, self.ransac_n, self.ransac_distance,
                                                seed=self.ransac_seed, fallback=T)
            c = int(num.item())  # the reference returns (C,3) tensors: data-dependent shape
            out["ref_corr_points"], out["src_corr_points"], out["corr_scores"] = ref_pad[:c], src_pad[:c], sc_pad[:c]
            out["lgr_transform"], out["estimated_transform"], out["ransac_info"] = T, T_sim, info
            return
        ref_corr, src_corr, corr_scores, T = self.fine_matching(ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks,
                                                                matching_scores, node_corr_scores)
        out["ref_corr_points"], out["src_corr_points"], out["corr_scores"] = ref_corr, src_corr, corr_scores
        out["estimated_transform"] = out["lgr_transform"] = T


def ops_gather_index(table, rows):
    """table[rows] for an int64 (M,K) table: a row gather (model.py:171-172); knn masks follow from the sentinel."""
    return table.index_select(0, rows)


def create_model(config, ransac=False, ransac_seed=0):
    return GeoTransformer(config, ransac=ransac, ransac_seed=ransac_seed)
