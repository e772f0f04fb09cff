"""Superpoint matching, optimal transport and local-to-global registration
(reference: geotransformer/modules/geotransformer/{superpoint_matching,local_global_registration}.py,
geotransformer/modules/sinkhorn/learnable_sinkhorn.py, geotransformer/modules/registration/procrustes.py)."""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops


class SuperPointMatching(nn.Module):
    """superpoint_matching.py:7-50."""

    def __init__(self, num_correspondences, dual_normalization=True):
        super().__init__()
        self.num_correspondences = num_correspondences
        self.dual_normalization = dual_normalization

    @torch.no_grad()
    def forward(self, ref_feats, src_feats, ref_masks=None, src_masks=None, lazy=False):
        """lazy=True (extension): no host sync -- returns the full-length (k) arrays plus the device scalar `count`; the caller
        must check count == k before trusting entries beyond it (GeoTransformer.forward does, after queueing its tail)."""
        dev = ref_feats.device
        if ref_masks is None:
            ref_masks = torch.ones(ref_feats.shape[0], dtype=torch.bool, device=dev)
        if src_masks is None:
            src_masks = torch.ones(src_feats.shape[0], dtype=torch.bool, device=dev)
        ref_idx, src_idx, scores, count = ops.superpoint_matching(
            ref_feats, src_feats, ref_masks, src_masks, self.num_correspondences, self.dual_normalization)
        if lazy:
            return ref_idx, src_idx, scores, count
        # the reference returns min(k, #valid pairs) entries (data-dependent length)
        c = int(count.item())
        if c < self.num_correspondences:
            ref_idx, src_idx, scores = ref_idx[:c], src_idx[:c], scores[:c]
        return ref_idx, src_idx, scores


class LearnableLogOptimalTransport(nn.Module):
    """learnable_sinkhorn.py:5-70."""

    def __init__(self, num_iterations, inf=1e12):
        super().__init__()
        self.num_iterations = num_iterations
        self.register_parameter("alpha", torch.nn.Parameter(torch.tensor(1.0)))
        self.inf = inf

    @torch.no_grad()
    def forward(self, scores, row_masks=None, col_masks=None):
        B, M, N = scores.shape
        if row_masks is None:
            row_masks = torch.ones((B, M), dtype=torch.bool, device=scores.device)
        if col_masks is None:
            col_masks = torch.ones((B, N), dtype=torch.bool, device=scores.device)
        return ops.sinkhorn(scores, row_masks, col_masks, self.alpha, self.num_iterations, self.inf)

    def __repr__(self):
        return self.__class__.__name__ + "(num_iterations={})".format(self.num_iterations)


def weighted_procrustes(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-5, return_transform=False):
    """procrustes.py:6-82."""
    if weights is None:
        weights = torch.ones_like(src_points[..., 0])
    if weight_thresh > 0.0:
        weights = torch.where(weights < weight_thresh, torch.zeros_like(weights), weights)
    T = ops.weighted_procrustes(src_points, ref_points, weights, eps)
    if return_transform:
        return T
    return T[..., :3, :3], T[..., :3, 3]


class WeightedProcrustes(nn.Module):
    """procrustes.py:85-100."""

    def __init__(self, weight_thresh=0.0, eps=1e-5, return_transform=False):
        super().__init__()
        self.weight_thresh, self.eps, self.return_transform = weight_thresh, eps, return_transform

    @torch.no_grad()
    def forward(self, src_points, tgt_points, weights=None):
        return weighted_procrustes(src_points, tgt_points, weights, self.weight_thresh, self.eps, self.return_transform)


class LocalGlobalRegistration(nn.Module):
    """local_global_registration.py:11-235 for the configuration the model uses (mutual matching,
    no dustbin, no global score, no correspondence limit)."""

    def __init__(self, k: int, acceptance_radius: float, mutual: bool = True, confidence_threshold: float = 0.05,
                 use_dustbin: bool = False, use_global_score: bool = False, correspondence_threshold: int = 3,
                 correspondence_limit: Optional[int] = None, num_refinement_steps: int = 5):
        super().__init__()
        if use_dustbin or use_global_score or correspondence_limit is not None or not mutual:
            raise NotImplementedError("only the GaussReg configuration (config.py:116-125) is implemented")
        self.k, self.acceptance_radius, self.mutual = k, acceptance_radius, mutual
        self.confidence_threshold = confidence_threshold
        self.use_dustbin, self.use_global_score = use_dustbin, use_global_score
        self.correspondence_threshold, self.correspondence_limit = correspondence_threshold, correspondence_limit
        self.num_refinement_steps = num_refinement_steps
        self.procrustes = WeightedProcrustes(return_transform=True)

    @torch.no_grad()
    def forward_device(self, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat):
        """No host sync: padded correspondence buffers + device count + transform."""
        return ops.local_global_registration(score_mat, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, self.k,
                                             self.acceptance_radius, self.mutual, self.confidence_threshold,
                                             self.correspondence_threshold, self.num_refinement_steps)

    @torch.no_grad()
    def forward(self, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat, global_scores=None):
        ref_c, src_c, sc, num, T = self.forward_device(ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat)
        c = int(num.item())  # the reference returns (C,3) tensors: data-dependent shape
        return ref_c[:c], src_c[:c], sc[:c], T
