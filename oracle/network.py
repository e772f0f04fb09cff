"""TEST INFRASTRUCTURE ONLY -- torch fp32 (CPU) restatement of the reference's network half of the
coarse-registration forward.

Functional style: every function takes plain tensors plus the reference's ``state_dict`` (same 316
keys), so it can be checked (a) against golden vectors produced by the real reference modules
(tests/golden/make_network_golden.py, tests/test_oracle_network.py) and (b) against the CUDA path
on any seeded input (tests/test_*_gpu.py).  Floating-point work, so a torch reference is kept as
the oracle (tolerances are written in the tests).

Nothing under gaussreg_b200/ imports this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# O1 helpers
# ------------------------------------------------------------------------------------------------


def pairwise_distance(x, y, normalized=False):
    """geotransformer/modules/ops/pairwise_distance.py:4-31 (channel-last form)."""
    xy = torch.matmul(x, y.transpose(-1, -2))
    if normalized:
        sq = 2.0 - 2.0 * xy
    else:
        x2 = torch.sum(x ** 2, dim=-1).unsqueeze(-1)
        y2 = torch.sum(y ** 2, dim=-1).unsqueeze(-2)
        sq = x2 - 2 * xy + y2
    return sq.clamp(min=0.0)


def apply_transform(points, transform):
    """geotransformer/modules/ops/transformation.py:7-60."""
    if transform.ndim == 2:
        R, t = transform[:3, :3], transform[:3, 3]
        shape = points.shape
        return (points.reshape(-1, 3) @ R.transpose(-1, -2) + t).reshape(shape)
    R, t = transform[:, :3, :3], transform[:, None, :3, 3]
    return points @ R.transpose(-1, -2) + t


def gather_rows(data, index):
    """index_select(data, index, dim=0) for an index of any rank (modules/ops/index_select.py:4-31)."""
    return data.index_select(0, index.reshape(-1)).view(*index.shape, *data.shape[1:])


# ------------------------------------------------------------------------------------------------
# P1 point-to-node partition
# ------------------------------------------------------------------------------------------------


def point_to_node_partition(points, nodes, point_limit):
    """geotransformer/modules/ops/pointcloud_partition.py:61-111."""
    sq = pairwise_distance(nodes, points)  # (M, N)
    point_to_node = sq.min(dim=0)[1]
    node_masks = torch.zeros(nodes.shape[0], dtype=torch.bool)
    node_masks.index_fill_(0, point_to_node, True)
    match = torch.zeros_like(sq, dtype=torch.bool)
    match[point_to_node, torch.arange(points.shape[0])] = True
    sq = sq.masked_fill(~match, 1e12)
    knn_idx = sq.topk(k=point_limit, dim=1, largest=False)[1]
    knn_node = gather_rows(point_to_node, knn_idx)
    node_idx = torch.arange(nodes.shape[0]).unsqueeze(1).expand(-1, point_limit)
    knn_masks = torch.eq(knn_node, node_idx)
    knn_idx = knn_idx.masked_fill(~knn_masks, points.shape[0])
    return point_to_node, node_masks, knn_idx, knn_masks


# ------------------------------------------------------------------------------------------------
# K1-K4 KPConv backbone
# ------------------------------------------------------------------------------------------------


def kpconv(sd, prefix, s_feats, q_points, s_points, neighbor_indices, sigma):
    """geotransformer/modules/kpconv/kpconv.py:79-122."""
    W, bias, kp = sd[prefix + ".weights"], sd.get(prefix + ".bias"), sd[prefix + ".kernel_points"]
    s_points = torch.cat([s_points, torch.zeros_like(s_points[:1]) + 1e6], 0)
    neighbors = gather_rows(s_points, neighbor_indices) - q_points.unsqueeze(1)  # (M,H,3)
    diff = neighbors.unsqueeze(2) - kp  # (M,H,K,3)
    sq = torch.sum(diff ** 2, dim=3)
    w = torch.clamp(1 - torch.sqrt(sq) / sigma, min=0.0).transpose(1, 2)  # (M,K,H)
    s_feats = torch.cat((s_feats, torch.zeros_like(s_feats[:1])), 0)
    nf = gather_rows(s_feats, neighbor_indices)  # (M,H,C)
    wf = torch.matmul(w, nf).permute(1, 0, 2)  # (K,M,C)
    out = torch.matmul(wf, W).sum(dim=0)  # (M,C_out)
    num = torch.sum(torch.gt(torch.sum(nf, dim=-1), 0.0), dim=-1)
    num = torch.max(num, torch.ones_like(num))
    out = out / num.unsqueeze(1)
    if bias is not None:
        out = out + bias
    return out


def group_norm(sd, prefix, x, groups=32):
    """geotransformer/modules/kpconv/modules.py:33-50: statistics over all rows of the stacked pair."""
    y = F.group_norm(x.transpose(0, 1).unsqueeze(0), groups, sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"])
    return y.squeeze(0).transpose(0, 1)


def unary_block(sd, prefix, x, has_relu=True, groups=32):
    """modules.py:53-83."""
    x = F.linear(x, sd[prefix + ".mlp.weight"], sd[prefix + ".mlp.bias"])
    x = group_norm(sd, prefix + ".norm", x, groups)
    return F.leaky_relu(x, 0.1) if has_relu else x


def maxpool(x, neighbor_indices):
    """geotransformer/modules/kpconv/functional.py:54-67."""
    x = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return gather_rows(x, neighbor_indices).max(1)[0]


def nearest_upsample(x, upsample_indices):
    """functional.py:6-22."""
    x = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return x.index_select(0, upsample_indices[:, 0])


def conv_block(sd, prefix, s_feats, q_points, s_points, idx, sigma, groups=32):
    """modules.py:104-146."""
    x = kpconv(sd, prefix + ".KPConv", s_feats, q_points, s_points, idx, sigma)
    return F.leaky_relu(group_norm(sd, prefix + ".norm", x, groups), 0.1)


def residual_block(sd, prefix, s_feats, q_points, s_points, idx, sigma, strided=False, groups=32):
    """modules.py:149-225."""
    x = unary_block(sd, prefix + ".unary1", s_feats, True, groups) if (prefix + ".unary1.mlp.weight") in sd else s_feats
    x = kpconv(sd, prefix + ".KPConv", x, q_points, s_points, idx, sigma)
    x = F.leaky_relu(group_norm(sd, prefix + ".norm_conv", x, groups), 0.1)
    x = unary_block(sd, prefix + ".unary2", x, False, groups)
    sc = maxpool(s_feats, idx) if strided else s_feats
    if (prefix + ".unary_shortcut.mlp.weight") in sd:
        sc = unary_block(sd, prefix + ".unary_shortcut", sc, False, groups)
    return F.leaky_relu(x + sc, 0.1)


def kpconv_fpn(sd, feats, data, init_sigma, groups=32, prefix="backbone", taps=None):
    """experiments/geotransformer.gaussian_splatting.indoor/backbone.py:164-212.
    Returns [f_s2, f_s3, f_s4, f_s5]."""
    P, NB, SUB, UP = data["points"], data["neighbors"], data["subsampling"], data["upsampling"]
    s = init_sigma
    p = prefix + "."

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    f1 = tap("encoder1_1", conv_block(sd, p + "encoder1_1", feats, P[0], P[0], NB[0], s, groups))
    f1 = tap("encoder1_2", residual_block(sd, p + "encoder1_2", f1, P[0], P[0], NB[0], s, False, groups))
    f2 = tap("encoder2_1", residual_block(sd, p + "encoder2_1", f1, P[1], P[0], SUB[0], s, True, groups))
    f2 = tap("encoder2_2", residual_block(sd, p + "encoder2_2", f2, P[1], P[1], NB[1], s * 2, False, groups))
    f2 = tap("encoder2_3", residual_block(sd, p + "encoder2_3", f2, P[1], P[1], NB[1], s * 2, False, groups))
    f3 = tap("encoder3_1", residual_block(sd, p + "encoder3_1", f2, P[2], P[1], SUB[1], s * 2, True, groups))
    f3 = tap("encoder3_2", residual_block(sd, p + "encoder3_2", f3, P[2], P[2], NB[2], s * 4, False, groups))
    f3 = tap("encoder3_3", residual_block(sd, p + "encoder3_3", f3, P[2], P[2], NB[2], s * 4, False, groups))
    f4 = tap("encoder4_1", residual_block(sd, p + "encoder4_1", f3, P[3], P[2], SUB[2], s * 4, True, groups))
    f4 = tap("encoder4_2", residual_block(sd, p + "encoder4_2", f4, P[3], P[3], NB[3], s * 8, False, groups))
    f4 = tap("encoder4_3", residual_block(sd, p + "encoder4_3", f4, P[3], P[3], NB[3], s * 8, False, groups))
    f5 = tap("encoder5_1", residual_block(sd, p + "encoder5_1", f4, P[4], P[3], SUB[3], s * 8, True, groups))
    f5 = tap("encoder5_2", residual_block(sd, p + "encoder5_2", f5, P[4], P[4], NB[4], s * 16, False, groups))
    f5 = tap("encoder5_3", residual_block(sd, p + "encoder5_3", f5, P[4], P[4], NB[4], s * 16, False, groups))
    l4 = torch.cat([nearest_upsample(f5, UP[3]), f4], dim=1)
    l4 = tap("decoder4", unary_block(sd, p + "decoder4", l4, True, groups))
    l3 = torch.cat([nearest_upsample(l4, UP[2]), f3], dim=1)
    l3 = tap("decoder3", unary_block(sd, p + "decoder3", l3, True, groups))
    l2 = torch.cat([nearest_upsample(l3, UP[1]), f2], dim=1)
    l2 = tap("decoder2", F.linear(l2, sd[p + "decoder2.mlp.weight"], sd[p + "decoder2.mlp.bias"]))
    return [l2, l3, l4, f5]


# ------------------------------------------------------------------------------------------------
# T1-T4 geometric transformer
# ------------------------------------------------------------------------------------------------


def sinusoidal_embedding(idx, div_term):
    """geotransformer/modules/transformer/positional_embedding.py:19-35 (interleaved sin/cos)."""
    om = idx.reshape(-1, 1, 1) * div_term.view(1, -1, 1)
    emb = torch.cat([torch.sin(om), torch.cos(om)], dim=2)
    return emb.view(*idx.shape, div_term.numel() * 2)


def embedding_indices(points, sigma_d, sigma_a, angle_k):
    """geotransformer/modules/geotransformer/geotransformer.py:26-55.  points (N,3)."""
    dist = torch.sqrt(pairwise_distance(points, points))
    d_idx = dist / sigma_d
    k = angle_k
    knn = dist.topk(k=k + 1, dim=1, largest=False)[1][:, 1:]  # (N,k)
    knn_points = points[knn]  # (N,k,3)
    ref_v = knn_points - points.unsqueeze(1)  # (N,k,3)
    anc_v = points.unsqueeze(0) - points.unsqueeze(1)  # (N,N,3)
    ref_v = ref_v.unsqueeze(1).expand(-1, points.shape[0], -1, -1)  # (N,N,k,3)
    anc_v = anc_v.unsqueeze(2).expand(-1, -1, k, -1)
    sin_v = torch.linalg.norm(torch.cross(ref_v, anc_v, dim=-1), dim=-1)
    cos_v = torch.sum(ref_v * anc_v, dim=-1)
    a_idx = torch.atan2(sin_v, cos_v) * (180.0 / (sigma_a * np.pi))
    return d_idx, a_idx, knn


def structure_embedding(sd, points, sigma_d, sigma_a, angle_k, prefix="transformer.embedding"):
    """geotransformer.py:57-72 (reduction 'max').  points (N,3) -> (N,N,C)."""
    d_idx, a_idx, _ = embedding_indices(points, sigma_d, sigma_a, angle_k)
    div = sd[prefix + ".embedding.div_term"]
    d_emb = F.linear(sinusoidal_embedding(d_idx, div), sd[prefix + ".proj_d.weight"], sd[prefix + ".proj_d.bias"])
    a_emb = F.linear(sinusoidal_embedding(a_idx, div), sd[prefix + ".proj_a.weight"], sd[prefix + ".proj_a.bias"])
    return d_emb + a_emb.max(dim=2)[0]


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _heads(x, h):
    n, c = x.shape
    return x.view(n, h, c // h).permute(1, 0, 2)  # (h, n, c/h)


def attention_output(sd, prefix, x):
    """geotransformer/modules/transformer/output_layer.py:6-21."""
    h = _lin(sd, prefix + ".squeeze", F.relu(_lin(sd, prefix + ".expand", x)))
    return F.layer_norm(x + h, (x.shape[-1],), sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"])


def rpe_layer(sd, prefix, x, emb, num_heads):
    """rpe_transformer.py:18-131 with memory == input (self attention).  x (N,C), emb (N,N,C)."""
    a = prefix + ".attention.attention"
    q = _heads(_lin(sd, a + ".proj_q", x), num_heads)
    k = _heads(_lin(sd, a + ".proj_k", x), num_heads)
    v = _heads(_lin(sd, a + ".proj_v", x), num_heads)
    n = x.shape[0]
    p = _lin(sd, a + ".proj_p", emb).view(n, n, num_heads, -1).permute(2, 0, 1, 3)  # (h,n,m,c)
    s_p = torch.einsum("hnc,hnmc->hnm", q, p)
    s_e = torch.einsum("hnc,hmc->hnm", q, k)
    s = F.softmax((s_e + s_p) / (q.shape[-1] ** 0.5), dim=-1)
    hid = torch.matmul(s, v).permute(1, 0, 2).reshape(n, -1)
    hid = _lin(sd, prefix + ".attention.linear", hid)
    y = F.layer_norm(hid + x, (x.shape[-1],), sd[prefix + ".attention.norm.weight"], sd[prefix + ".attention.norm.bias"])
    return attention_output(sd, prefix + ".output", y)


def cross_layer(sd, prefix, x, mem, num_heads):
    """vanilla_transformer.py:15-129."""
    a = prefix + ".attention.attention"
    q = _heads(_lin(sd, a + ".proj_q", x), num_heads)
    k = _heads(_lin(sd, a + ".proj_k", mem), num_heads)
    v = _heads(_lin(sd, a + ".proj_v", mem), num_heads)
    s = F.softmax(torch.einsum("hnc,hmc->hnm", q, k) / (q.shape[-1] ** 0.5), dim=-1)
    hid = torch.matmul(s, v).permute(1, 0, 2).reshape(x.shape[0], -1)
    hid = _lin(sd, prefix + ".attention.linear", hid)
    y = F.layer_norm(hid + x, (x.shape[-1],), sd[prefix + ".attention.norm.weight"], sd[prefix + ".attention.norm.bias"])
    return attention_output(sd, prefix + ".output", y)


def geometric_transformer(sd, ref_points, src_points, ref_feats, src_feats, cfg, prefix="transformer", taps=None):
    """geotransformer.py:114-155 + conditional_transformer.py:97-117 (sequential cross blocks)."""
    g = cfg["geotransformer"]
    e0 = structure_embedding(sd, ref_points, g["sigma_d"], g["sigma_a"], g["angle_k"], prefix + ".embedding")
    e1 = structure_embedding(sd, src_points, g["sigma_d"], g["sigma_a"], g["angle_k"], prefix + ".embedding")
    if taps is not None:
        taps["ref_embeddings"], taps["src_embeddings"] = e0, e1
    f0, f1 = _lin(sd, prefix + ".in_proj", ref_feats), _lin(sd, prefix + ".in_proj", src_feats)
    for i, block in enumerate(g["blocks"]):
        lp = f"{prefix}.transformer.layers.{i}"
        if block == "self":
            f0 = rpe_layer(sd, lp, f0, e0, g["num_heads"])
            f1 = rpe_layer(sd, lp, f1, e1, g["num_heads"])
        else:
            f0 = cross_layer(sd, lp, f0, f1, g["num_heads"])
            f1 = cross_layer(sd, lp, f1, f0, g["num_heads"])
        if taps is not None:
            taps[f"layer{i}_ref"], taps[f"layer{i}_src"] = f0, f1
    return _lin(sd, prefix + ".out_proj", f0), _lin(sd, prefix + ".out_proj", f1)


# ------------------------------------------------------------------------------------------------
# M1 superpoint matching, S1 Sinkhorn
# ------------------------------------------------------------------------------------------------


def superpoint_matching(ref_feats, src_feats, ref_masks, src_masks, num_correspondences, dual_normalization=True):
    """geotransformer/modules/geotransformer/superpoint_matching.py:13-50."""
    ref_idx = torch.nonzero(ref_masks, as_tuple=True)[0]
    src_idx = torch.nonzero(src_masks, as_tuple=True)[0]
    rf, sf = ref_feats[ref_idx], src_feats[src_idx]
    scores = torch.exp(-pairwise_distance(rf, sf, normalized=True))
    if dual_normalization:
        scores = (scores / scores.sum(dim=1, keepdim=True)) * (scores / scores.sum(dim=0, keepdim=True))
    k = min(num_correspondences, scores.numel())
    corr_scores, corr = scores.view(-1).topk(k=k, largest=True)
    return ref_idx[corr // scores.shape[1]], src_idx[corr % scores.shape[1]], corr_scores


def log_optimal_transport(scores, row_masks, col_masks, alpha, num_iterations, inf=1e12):
    """geotransformer/modules/sinkhorn/learnable_sinkhorn.py:5-66."""
    B, M, N = scores.shape
    prm = torch.zeros(B, M + 1, dtype=torch.bool)
    prm[:, :M] = ~row_masks
    pcm = torch.zeros(B, N + 1, dtype=torch.bool)
    pcm[:, :N] = ~col_masks
    psm = torch.logical_or(prm.unsqueeze(2), pcm.unsqueeze(1))
    pc = alpha.expand(B, M, 1)
    pr = alpha.expand(B, 1, N + 1)
    ps = torch.cat([torch.cat([scores, pc], dim=-1), pr], dim=1)
    ps = ps.masked_fill(psm, -inf)
    nvr = row_masks.float().sum(1)
    nvc = col_masks.float().sum(1)
    norm = -torch.log(nvr + nvc)
    log_mu = torch.empty(B, M + 1)
    log_mu[:, :M] = norm.unsqueeze(1)
    log_mu[:, M] = torch.log(nvc) + norm
    log_mu[prm] = -inf
    log_nu = torch.empty(B, N + 1)
    log_nu[:, :N] = norm.unsqueeze(1)
    log_nu[:, N] = torch.log(nvr) + norm
    log_nu[pcm] = -inf
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(num_iterations):
        u = log_mu - torch.logsumexp(ps + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(ps + u.unsqueeze(2), dim=1)
    out = ps + u.unsqueeze(2) + v.unsqueeze(1)
    return out - norm.unsqueeze(1).unsqueeze(2)


# ------------------------------------------------------------------------------------------------
# L1 / L2 local-to-global registration
# ------------------------------------------------------------------------------------------------


def weighted_procrustes(src, ref, weights, eps=1e-5):
    """geotransformer/modules/registration/procrustes.py:6-82 -> (B,4,4) or (4,4)."""
    squeeze = src.ndim == 2
    if squeeze:
        src, ref, weights = src.unsqueeze(0), ref.unsqueeze(0), weights.unsqueeze(0)
    B = src.shape[0]
    w = torch.where(weights < 0.0, torch.zeros_like(weights), weights)
    w = (w / (torch.sum(w, dim=1, keepdim=True) + eps)).unsqueeze(2)
    sc = torch.sum(src * w, dim=1, keepdim=True)
    rc = torch.sum(ref * w, dim=1, keepdim=True)
    H = (src - sc).permute(0, 2, 1) @ (w * (ref - rc))
    U, _, V = torch.svd(H)
    Ut = U.transpose(1, 2)
    eye = torch.eye(3).unsqueeze(0).repeat(B, 1, 1)
    eye[:, -1, -1] = torch.sign(torch.det(V @ Ut))
    R = V @ eye @ Ut
    t = (rc.permute(0, 2, 1) - R @ sc.permute(0, 2, 1)).squeeze(2)
    T = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    return T.squeeze(0) if squeeze else T


def correspondence_matrix(score_mat, ref_masks, src_masks, k, conf, mutual=True):
    """local_global_registration.py:49-83 (score_mat already exp'ed, no dustbin)."""
    mask = torch.logical_and(ref_masks.unsqueeze(2), src_masks.unsqueeze(1))
    B, R, S = score_mat.shape
    rs, ri = score_mat.topk(k=k, dim=2)
    ref_sm = torch.zeros_like(score_mat).scatter_(2, ri, rs)
    ss, si = score_mat.topk(k=k, dim=1)
    src_sm = torch.zeros_like(score_mat).scatter_(1, si, ss)
    rc, sc = ref_sm > conf, src_sm > conf
    corr = torch.logical_and(rc, sc) if mutual else torch.logical_or(rc, sc)
    return torch.logical_and(corr, mask)


def local_global_registration(ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat, cfg, taps=None,
                              force_best=None):
    """local_global_registration.py:137-235 with the config of config.py:116-125
    (no dustbin, no global score, no correspondence limit)."""
    f = cfg["fine_matching"]
    radius, steps, thr = f["acceptance_radius"], f["num_refinement_steps"], f["correspondence_threshold"]
    score_mat = torch.exp(score_mat)
    corr_mat = correspondence_matrix(score_mat, ref_knn_masks, src_knn_masks, f["topk"], f["confidence_threshold"], f["mutual"])
    score_mat = score_mat * corr_mat.float()
    b_idx, r_idx, s_idx = torch.nonzero(corr_mat, as_tuple=True)
    ref_c = ref_knn_points[b_idx, r_idx]
    src_c = src_knn_points[b_idx, s_idx]
    sc = score_mat[b_idx, r_idx, s_idx]

    def rescore(T):
        res = torch.linalg.norm(ref_c - apply_transform(src_c, T), dim=1)
        return sc * (res < radius).float()

    counts = torch.bincount(b_idx, minlength=score_mat.shape[0])
    keep = torch.nonzero(counts >= thr, as_tuple=True)[0]
    if taps is not None:
        taps["patch_counts"] = counts
    if keep.numel() > 0:
        starts = torch.cumsum(counts, 0) - counts
        mc = int(counts[keep].max())
        bs = torch.zeros(keep.numel(), mc, 3)
        br = torch.zeros(keep.numel(), mc, 3)
        bw = torch.zeros(keep.numel(), mc)
        for j, pch in enumerate(keep.tolist()):
            a, c = int(starts[pch]), int(counts[pch])
            bs[j, :c], br[j, :c], bw[j, :c] = src_c[a:a + c], ref_c[a:a + c], sc[a:a + c]
        Ts = weighted_procrustes(bs, br, bw)
        aligned = apply_transform(src_c.unsqueeze(0), Ts)
        inl = torch.linalg.norm(ref_c.unsqueeze(0) - aligned, dim=2) < radius
        best = inl.sum(dim=1).argmax() if force_best is None else torch.as_tensor(force_best)  # test hook: alternative tie-break
        if taps is not None:
            taps["local_transforms"], taps["inlier_counts"], taps["best_index"] = Ts, inl.sum(dim=1), best
        cur = sc * inl[best].float()
    else:
        T = weighted_procrustes(src_c, ref_c, sc)
        cur = rescore(T)
    T = weighted_procrustes(src_c, ref_c, cur)
    for _ in range(steps - 1):
        T = weighted_procrustes(src_c, ref_c, rescore(T))
    return ref_c, src_c, sc, T


# ------------------------------------------------------------------------------------------------
# full eval forward (model.py:69-222 without the Open3D RANSAC of :209-215)
# ------------------------------------------------------------------------------------------------

DEFAULT_CFG = {
    "backbone": {"num_stages": 5, "init_voxel_size": 0.025, "kernel_size": 15, "init_radius": 0.0625,
                 "init_sigma": 0.05, "group_norm": 32, "input_dim": 4, "init_dim": 64, "output_dim": 256},
    "model": {"num_points_in_patch": 128, "num_sinkhorn_iterations": 100},
    "coarse_matching": {"num_correspondences": 256, "dual_normalization": True},
    "geotransformer": {"input_dim": 2048, "hidden_dim": 256, "output_dim": 256, "num_heads": 4,
                       "blocks": ["self", "cross", "self", "cross", "self", "cross"], "sigma_d": 0.2,
                       "sigma_a": 15, "angle_k": 3, "reduction_a": "max"},
    "fine_matching": {"topk": 3, "acceptance_radius": 0.1, "mutual": True, "confidence_threshold": 0.05,
                      "use_dustbin": False, "use_global_score": False, "correspondence_threshold": 3,
                      "correspondence_limit": None, "num_refinement_steps": 5},
}


def forward(sd, data, cfg=DEFAULT_CFG, taps=None):
    """data: dict with 'features', 'points', 'lengths', 'neighbors', 'subsampling', 'upsampling'
    (lists of CPU tensors as produced by registration_collate_fn_stack_mode)."""
    out = {}
    rl_c, rl_f = int(data["lengths"][-1][0]), int(data["lengths"][1][0])
    pc, pf = data["points"][-1], data["points"][1]
    ref_c, src_c, ref_f, src_f = pc[:rl_c], pc[rl_c:], pf[:rl_f], pf[rl_f:]
    K = cfg["model"]["num_points_in_patch"]
    _, ref_nm, ref_knn, ref_km = point_to_node_partition(ref_f, ref_c, K)
    _, src_nm, src_knn, src_km = point_to_node_partition(src_f, src_c, K)
    ref_knn_pts = gather_rows(torch.cat([ref_f, torch.zeros_like(ref_f[:1])], 0), ref_knn)
    src_knn_pts = gather_rows(torch.cat([src_f, torch.zeros_like(src_f[:1])], 0), src_knn)

    feats_list = kpconv_fpn(sd, data["features"], data, cfg["backbone"]["init_sigma"], cfg["backbone"]["group_norm"], taps=taps)
    feats_c, feats_f = feats_list[-1], feats_list[0]
    rfc, sfc = geometric_transformer(sd, ref_c, src_c, feats_c[:rl_c], feats_c[rl_c:], cfg, taps=taps)
    rfc_n, sfc_n = F.normalize(rfc, p=2, dim=1), F.normalize(sfc, p=2, dim=1)
    out["ref_feats_c"], out["src_feats_c"] = rfc_n, sfc_n
    ref_ff, src_ff = feats_f[:rl_f], feats_f[rl_f:]

    cm = cfg["coarse_matching"]
    ref_ci, src_ci, node_scores = superpoint_matching(rfc_n, sfc_n, ref_nm, src_nm, cm["num_correspondences"], cm["dual_normalization"])
    out["ref_node_corr_indices"], out["src_node_corr_indices"], out["node_corr_scores"] = ref_ci, src_ci, node_scores

    r_knn, s_knn = ref_knn[ref_ci], src_knn[src_ci]
    r_km, s_km = ref_km[ref_ci], src_km[src_ci]
    r_kp, s_kp = ref_knn_pts[ref_ci], src_knn_pts[src_ci]
    r_kf = gather_rows(torch.cat([ref_ff, torch.zeros_like(ref_ff[:1])], 0), r_knn)
    s_kf = gather_rows(torch.cat([src_ff, torch.zeros_like(src_ff[:1])], 0), s_knn)
    ms = torch.einsum("bnd,bmd->bnm", r_kf, s_kf) / feats_f.shape[1] ** 0.5
    ms = log_optimal_transport(ms, r_km, s_km, sd["optimal_transport.alpha"], cfg["model"]["num_sinkhorn_iterations"])
    out["matching_scores"] = ms
    rcp, scp, cs, T = local_global_registration(r_kp, s_kp, r_km, s_km, ms[:, :-1, :-1], cfg, taps=taps)
    out.update(ref_corr_points=rcp, src_corr_points=scp, corr_scores=cs, estimated_transform=T,
               ref_node_knn_indices=ref_knn, src_node_knn_indices=src_knn, ref_node_masks=ref_nm, src_node_masks=src_nm,
               ref_node_knn_masks=ref_km, src_node_knn_masks=src_km)
    return out
