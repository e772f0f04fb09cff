"""GPU parity tests for the network half of the hot path (SURVEY.md section 8 rows P1, K1-K4, T1-T4, M1,
M2, S1, L1, L2): every CUDA kernel, called through the C ABI behind the reference's module interface,
against the CPU oracle (oracle/network.py, itself pinned to reference goldens) on the same seeded
inputs -- teacher-forced per stage, then end to end against the reference's golden transform.

Tolerances (fp32 path): rel-L2 <= 1e-5 for single kernels, <= 2e-4 after the 14-block backbone /
6-layer transformer; index outputs exact (ties aside); transform: 1e-4 Frobenius (north_star)."""
import os

import numpy as np
import pytest
import torch

from gaussreg_b200 import ops
from gaussreg_b200 import modules as gm
from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
from gaussreg_b200.data import registration_collate_fn_stack_mode
from gaussreg_b200.synthetic import make_pair_inputs
from oracle import network as onet
from tests.helpers import GOLDEN_DIR, golden_spec, oracle_data, rel_l2, seeded_model, to_cuda

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,trans_b", [(1, 1, 1, True), (37, 53, 19, True), (300, 256, 60, False), (1000, 64, 480, False),
                                           (5000, 128, 64, True), (460, 256, 2048, True), (129, 257, 515, False),
                                           (20000, 256, 256, True)])
def test_gemm_epilogues(M, N, K, trans_b):
    a = _rand(M, K, seed=1)
    b = _rand(N, K, seed=2) if trans_b else _rand(K, N, seed=2)
    bias, div, res = _rand(N, seed=3), torch.rand(M, generator=torch.Generator().manual_seed(4)) * 5 + 1, _rand(M, N, seed=5)
    ref = (a.double() @ (b.double().t() if trans_b else b.double()))
    got = ops.gemm(a.cuda(), b.cuda(), trans_b).cpu()
    assert rel_l2(got, ref) < 5e-6
    full = torch.nn.functional.leaky_relu(0.5 * ref / div.double()[:, None] + bias.double() + res.double(), 0.1)
    got = ops.gemm(a.cuda(), b.cuda(), trans_b, bias=bias.cuda(), alpha=0.5, row_div=div.cuda(), residual=res.cuda(),
                   act="leaky_relu").cpu()
    assert rel_l2(got, full) < 5e-6


def test_gemm_batched_head_slices():
    H, N, M, dh = 4, 77, 91, 64
    C = H * dh
    q, k = _rand(N, C, seed=1).cuda(), _rand(M, C, seed=2).cuda()
    P = torch.empty((H, N, M), device="cuda")
    ops.gemm_batched(q.data_ptr(), C, dh, k.data_ptr(), C, dh, True, P.data_ptr(), M, N * M, N, M, dh, H, alpha=0.125)
    want = torch.einsum("nhc,mhc->hnm", q.view(N, H, dh).double(), k.view(M, H, dh).double()) * 0.125
    assert rel_l2(P.cpu(), want.cpu()) < 2e-6


# ---------------------------------------------------------------------------------------------- K1-K3
@pytest.mark.parametrize("C,Co,H", [(4, 64, 20), (32, 32, 20), (64, 64, 30), (128, 128, 43), (256, 256, 47), (512, 512, 12),
                                    (16, 24, 89), (8, 8, 5), (32, 32, 89), (64, 64, 49), (32, 64, 7)])
def test_kpconv_vs_oracle(C, Co, H):
    g = torch.Generator().manual_seed(C + H)
    Ns, M = 700, 500
    s_pts = torch.rand(Ns, 3, generator=g)
    q_pts = s_pts[:M] + 0.01 * torch.randn(M, 3, generator=g)
    idx = torch.randint(0, Ns + 1, (M, H), generator=g)  # includes the shadow index Ns
    idx[::7, H // 2:] = Ns
    feats = torch.randn(Ns, C, generator=g)
    feats[::5] = -feats[::5].abs()  # rows with non-positive sums: exercise the neighbour-count rule
    W = torch.randn(15, C, Co, generator=g) / (15 * C) ** 0.5
    bias = torch.randn(Co, generator=g)
    kp = (torch.rand(15, 3, generator=g) - 0.5) * 0.3
    kp[0] = 0
    sigma = 0.25
    sd = {"k.weights": W, "k.bias": bias, "k.kernel_points": kp}
    want = onet.kpconv(sd, "k", feats, q_pts, s_pts, idx, sigma)
    got = ops.kpconv(feats.cuda(), q_pts.cuda(), s_pts.cuda(), idx.cuda(), W.cuda(), bias.cuda(), kp.cuda(), sigma).cpu()
    assert rel_l2(got, want) < 1e-5


@pytest.mark.parametrize("N,C", [(1000, 32), (5003, 64), (777, 2048), (300, 1024), (3001, 256), (15000, 256)])
def test_group_norm_and_fusions(N, C):
    x, add = _rand(N, C, seed=1) * 3 + 0.5, _rand(N, C, seed=2)
    gamma, beta = _rand(C, seed=3), _rand(C, seed=4)
    sd = {"n.norm.weight": gamma, "n.norm.bias": beta}
    want = onet.group_norm(sd, "n", x, 32)
    got = ops.group_norm(x.cuda(), 32, gamma.cuda(), beta.cuda()).cpu()
    assert rel_l2(got, want) < 1e-5
    want2 = torch.nn.functional.leaky_relu(want + add, 0.1)
    got2 = ops.group_norm(x.cuda(), 32, gamma.cuda(), beta.cuda(), add=add.cuda(), act="leaky_relu").cpu()
    assert rel_l2(got2, want2) < 1e-5


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    for p in module.parameters():
        p.data = torch.randn(p.shape, generator=g) * (0.5 if p.dim() == 1 else 1.0 / max(p.shape[-1], 1) ** 0.5) + (1.0 if p.dim() == 1 else 0.0)
    return module


@pytest.mark.parametrize("N,Cin,Cout,relu", [(5003, 64, 128, True), (3001, 1024, 256, False), (20000, 128, 32, True), (777, 256, 512, True),
                                             (4100, 2048, 1024, True), (60000, 64, 32, True), (967, 3072, 1024, False)])
def test_unary_block_fused_groupnorm_statistics_vs_oracle(N, Cin, Cout, relu):
    """UnaryBlock through gr_unary_block: the GroupNorm statistics come from the tensor-core product's epilogue
    (direct tile epilogue, split-K reduction, or the separate pass when the FFMA kernel runs) -- all vs the oracle."""
    blk = _randomize(gm.UnaryBlock(Cin, Cout, 32, has_relu=relu), N).eval()
    sd = {"u." + k: v.clone() for k, v in blk.state_dict().items()}
    x, add = _rand(N, Cin, seed=1) * 2 + 0.3, _rand(N, Cout, seed=2)
    want = onet.unary_block(sd, "u", x, relu)
    blk = blk.cuda()
    got = blk(x.cuda()).cpu()
    assert rel_l2(got, want) < 1e-5
    got2 = blk(x.cuda()).cpu()
    assert torch.equal(got, got2)  # fixed reduction order: bit-stable run to run
    if not relu:
        want3 = torch.nn.functional.leaky_relu(want + add, 0.1)
        got3 = blk(x.cuda(), add=add.cuda(), act_after_add="leaky_relu").cpu()
        assert rel_l2(got3, want3) < 1e-5


@pytest.mark.parametrize("Cin,Cout,strided", [(64, 128, False), (128, 128, True), (256, 256, False)])
def test_residual_and_conv_block_vs_oracle(Cin, Cout, strided):
    g = torch.Generator().manual_seed(Cin + Cout)
    Ns, M, H = 6000, (2500 if strided else 6000), 26
    s_pts = torch.rand(Ns, 3, generator=g)
    q_pts = s_pts[:M].clone() if not strided else s_pts[:M] + 0.005 * torch.randn(M, 3, generator=g)
    idx = torch.randint(0, Ns + 1, (M, H), generator=g)
    idx[::9, H // 3:] = Ns
    feats = torch.randn(Ns, Cin, generator=g)
    blk = _randomize(gm.ResidualBlock(Cin, Cout, 15, 0.3, 0.25, 32, strided=strided), 7).eval()
    sd = {"b." + k: v.clone() for k, v in blk.state_dict().items()}
    want = onet.residual_block(sd, "b", feats, q_pts, s_pts, idx, 0.25, strided)
    got = blk.cuda()(feats.cuda(), q_pts.cuda(), s_pts.cuda(), idx.cuda()).cpu()
    assert rel_l2(got, want) < 2e-5
    cb = _randomize(gm.ConvBlock(Cin, Cout, 15, 0.3, 0.25, 32), 8).eval()
    sd = {"c." + k: v.clone() for k, v in cb.state_dict().items()}
    want = onet.conv_block(sd, "c", feats, q_pts, s_pts, idx, 0.25)
    got = cb.cuda()(feats.cuda(), q_pts.cuda(), s_pts.cuda(), idx.cuda()).cpu()
    assert rel_l2(got, want) < 1e-5


def test_layer_norm_maxpool_upsample_gather():
    x, y = _rand(333, 256, seed=1), _rand(333, 256, seed=2)
    gamma, beta = _rand(256, seed=3), _rand(256, seed=4)
    want = torch.nn.functional.layer_norm(x + y, (256,), gamma, beta)
    assert rel_l2(ops.layer_norm_add(x.cuda(), y.cuda(), gamma.cuda(), beta.cuda()).cpu(), want) < 1e-5
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(400, 96, generator=g)
    idx = torch.randint(0, 401, (250, 17), generator=g)
    assert torch.equal(ops.maxpool(feats.cuda(), idx.cuda()).cpu(), onet.maxpool(feats, idx))
    skip = torch.randn(250, 40, generator=g)
    want = torch.cat([onet.nearest_upsample(feats, idx), skip], 1)
    assert torch.equal(ops.upsample_concat(feats.cuda(), idx.cuda(), skip.cuda()).cpu(), want)
    pad = torch.cat([feats, torch.zeros(1, 96)], 0)
    assert torch.equal(ops.gather_rows(feats.cuda(), idx.cuda()).cpu(), onet.gather_rows(pad, idx))


# ---------------------------------------------------------------------------------------------- P1
def test_point_to_node_partition_vs_oracle():
    """P1: exact agreement is expected; the only admissible differences are TIES of the reference's own arithmetic
    (pairwise_distance.py:19-30 evaluates x^2 - 2xy + y^2 in fp32, whose rounding noise of ~|x|^2 * 2^-23 can reorder
    two candidates that are equidistant to that precision).  Every mismatch is checked to be such a tie, and the exact
    rates are written to gpurun_out/p1_parity.txt."""
    data = oracle_data(dict(seed=3, n_points=6000))
    n_c, n_f = int(data["lengths"][-1][0]), int(data["lengths"][1][0])
    nodes, pts = data["points"][-1][:n_c], data["points"][1][:n_f]
    p2n, nm, ki, km = onet.point_to_node_partition(pts, nodes, 128)
    g_p2n, g_nm, g_ki, g_km = [t.cpu() for t in ops.point_to_node_partition(pts.cuda(), nodes.cuda(), 128)]
    d2 = torch.cdist(pts.double(), nodes.double()) ** 2                       # (n_f, n_c) exact to fp64
    tie_eps = 4.0 * float((pts.double() ** 2).sum(1).max()) * 2.0 ** -23      # rounding noise of the fp32 expansion
    bad = torch.nonzero(g_p2n != p2n).flatten()
    for i in bad.tolist():
        assert abs(float(d2[i, g_p2n[i]] - d2[i, p2n[i]])) <= tie_eps, (i, float(d2[i, g_p2n[i]]), float(d2[i, p2n[i]]))
    assert torch.equal(g_nm, nm)
    rows_bad = torch.nonzero(~(g_ki == ki).all(1)).flatten()
    for r in rows_bad.tolist():
        # same multiset of points up to re-assigned tie points; order differs only between equidistant points
        a, b = g_ki[r], ki[r]
        diff = torch.nonzero(a != b).flatten()
        for c in diff.tolist():
            ia, ib = int(a[c]), int(b[c])
            da = float(d2[ia, r]) if ia < n_f else float("inf")
            db = float(d2[ib, r]) if ib < n_f else float("inf")
            tied_assignment = (ia in bad.tolist()) or (ib in bad.tolist())
            assert tied_assignment or abs(da - db) <= tie_eps, (r, c, ia, ib, da, db)
    assert torch.equal(g_km, g_ki != n_f)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/p1_parity.txt", "w") as f:
        f.write(f"point_to_node: {n_f - bad.numel()}/{n_f} identical assignments ({bad.numel()} ties of the reference's fp32 distance); "
                f"knn rows: {n_c - rows_bad.numel()}/{n_c} identical\n")
    assert bad.numel() <= max(2, n_f // 2000) and rows_bad.numel() <= max(4, n_c // 50), (bad.numel(), rows_bad.numel())
    # small limit: truncation keeps the nearest ones
    _, _, ki8, km8 = onet.point_to_node_partition(pts, nodes, 8)
    _, _, g_ki8, g_km8 = [t.cpu() for t in ops.point_to_node_partition(pts.cuda(), nodes.cuda(), 8)]
    assert (g_ki8 == ki8).all(1).float().mean().item() > 0.99


# ---------------------------------------------------------------------------------------------- T1-T4
def _transformer_pair(seed=0):
    m = seeded_model(0)
    return m, m.state_dict()


def test_structure_embedding_vs_oracle():
    m, sd = _transformer_pair()
    g = torch.Generator().manual_seed(1)
    pts = (torch.rand(311, 3, generator=g) - 0.5) * torch.tensor([4.0, 3.0, 2.5])
    want = onet.structure_embedding(sd, pts, 0.2, 15, 3)
    emb = m.transformer.embedding.cuda()
    got = emb(pts.cuda()).cpu()  # default: Hermite tables of proj(sinusoid(.)) (csrc/embedding_tab.cu)
    assert rel_l2(got, want) < 2e-6 and float((got - want).abs().max()) < 2e-5
    os.environ["GAUSSREG_T1"] = "tc"
    try:
        got = emb(pts.cuda()).cpu()  # fused tensor-core kernel
    finally:
        del os.environ["GAUSSREG_T1"]
    assert rel_l2(got, want) < 1e-5
    emb.use_fused = False
    emb.CHUNK_ROWS = 20000  # unfused path, chunked
    got = emb(pts.cuda()).cpu()
    emb.use_fused = True
    assert rel_l2(got, want) < 1e-5
    d_idx, a_idx, knn = onet.embedding_indices(pts, 0.2, 15, 3)
    gd, ga, gk = [t.cpu() for t in ops.embedding_indices(pts.cuda(), 0.2, 15, 3)]
    assert torch.equal(gk.long(), knn)
    assert rel_l2(gd, d_idx) < 1e-6 and float((ga - a_idx).abs().max()) < 1e-4


def test_structure_embedding_table_edges():
    """Table path: indices beyond the distance table (direct evaluation), angle_k 1..3, the exact fp64 function."""
    m, sd = _transformer_pair()
    emb = m.transformer.embedding.cuda()
    g = torch.Generator().manual_seed(5)
    # 40 m cloud: dist / sigma_d reaches ~300, beyond the shared-memory nodes ([0, 64]) -> global-memory tier of the table;
    # 400 m cloud: beyond the table ([0, 1024]) -> direct evaluation.  The oracle's own fp32 phase x * w carries an error
    # of x * 2^-24, which is what the tolerance follows
    for extent, tol in ((40.0, 5e-6), (400.0, 5e-5)):
        pts = (torch.rand(97, 3, generator=g) - 0.5) * extent
        want = onet.structure_embedding(sd, pts, 0.2, 15, 3)
        got = emb(pts.cuda()).cpu()
        assert rel_l2(got, want) < tol, extent
    # angle_k = 1, 2 through the C-ABI wrapper, against the exact function in fp64
    pts = (torch.rand(150, 3, generator=g) - 0.5) * torch.tensor([4.0, 3.0, 2.5])
    Wd, bd = emb.proj_d.weight.detach().double().cpu(), emb.proj_d.bias.detach().double().cpu()
    Wa, ba = emb.proj_a.weight.detach().double().cpu(), emb.proj_a.bias.detach().double().cpu()
    div = emb.embedding.div_term.detach().double().cpu()

    def exact(x, W, b):
        om = x.double().reshape(-1, 1) * div.reshape(1, -1)
        E = torch.stack([torch.sin(om), torch.cos(om)], dim=2).reshape(x.numel(), -1)
        return E @ W.T + b

    for k in (1, 2, 3):
        d_idx, a_idx, _ = ops.embedding_indices(pts.cuda(), 0.2, 15, k)
        got = ops.structure_embedding_tabulated(d_idx, a_idx, emb.embedding.div_term, emb.proj_d.weight, emb.proj_d.bias,
                                                emb.proj_a.weight, emb.proj_a.bias, 15).cpu()
        fa = exact(a_idx.cpu(), Wa, ba).reshape(150, 150, k, -1).max(dim=2)[0]
        ex = (exact(d_idx.cpu(), Wd, bd).reshape(150, 150, -1) + fa).float()
        assert rel_l2(got, ex) < 5e-7, k
    # other angle bandwidths: sigma_a = 7.5 doubles the angle table (still in shared memory), sigma_a = 3 does not fit one
    # SM's shared memory -> GR_ERR_CAPACITY -> the module falls back to the tensor-core kernel
    for sigma_a in (7.5, 3.0):
        emb.sigma_a, emb.factor_a = sigma_a, 180.0 / (sigma_a * np.pi)
        want = onet.structure_embedding(sd, pts, 0.2, sigma_a, 3)
        got = emb(pts.cuda()).cpu()
        assert rel_l2(got, want) < 1e-5, sigma_a
    emb.sigma_a, emb.factor_a = 15, 180.0 / (15 * np.pi)
    # NaN coordinates propagate as they do through torch.max
    bad = pts.clone()
    bad[3, 1] = float("nan")
    got = emb(bad.cuda()).cpu()
    assert torch.isnan(got[3]).all() and torch.isnan(got[:, 3]).all()
    # a changed weight rebuilds the table
    with torch.no_grad():
        emb.proj_a.bias.add_(1.0)
    got2 = emb(pts.cuda()).cpu()
    with torch.no_grad():
        emb.proj_a.bias.sub_(1.0)
    got1 = emb(pts.cuda()).cpu()
    assert float((got2 - got1 - 1.0).abs().max()) < 1e-5


def test_transformer_layers_vs_oracle():
    m, sd = _transformer_pair()
    g = torch.Generator().manual_seed(2)
    n0, n1 = 203, 157
    p0, p1 = torch.rand(n0, 3, generator=g) * 3, torch.rand(n1, 3, generator=g) * 3
    f0, f1 = torch.randn(n0, 256, generator=g), torch.randn(n1, 256, generator=g)
    e0 = onet.structure_embedding(sd, p0, 0.2, 15, 3)
    want_self = onet.rpe_layer(sd, "transformer.transformer.layers.0", f0, e0, 4)
    want_cross = onet.cross_layer(sd, "transformer.transformer.layers.1", f0, f1, 4)
    tr = m.transformer.cuda()
    got_self, _ = tr.transformer.layers[0](f0.cuda(), f0.cuda(), e0.cuda())
    got_cross, _ = tr.transformer.layers[1](f0.cuda(), f1.cuda())
    assert rel_l2(got_self.cpu(), want_self) < 1e-5
    assert rel_l2(got_cross.cpu(), want_cross) < 1e-5
    # whole transformer (2048-d inputs)
    x0, x1 = torch.randn(n0, 2048, generator=g), torch.randn(n1, 2048, generator=g)
    cfg = {"geotransformer": dict(make_cfg().geotransformer)}
    w0, w1 = onet.geometric_transformer(sd, p0, p1, x0, x1, cfg)
    g0, g1 = tr(p0.cuda()[None], p1.cuda()[None], x0.cuda()[None], x1.cuda()[None])
    assert g0.shape == (1, n0, 256)
    assert rel_l2(g0[0].cpu(), w0) < 5e-5 and rel_l2(g1[0].cpu(), w1) < 5e-5


# ---------------------------------------------------------------------------------------------- M1 / S1 / L1 / L2
def test_superpoint_matching_vs_oracle():
    g = torch.Generator().manual_seed(4)
    rf = torch.nn.functional.normalize(torch.randn(431, 256, generator=g), dim=1)
    sf = torch.nn.functional.normalize(torch.randn(397, 256, generator=g) + 0.5 * rf[:397], dim=1)
    rm, sm = torch.ones(431, dtype=torch.bool), torch.ones(397, dtype=torch.bool)
    rm[[5, 77]] = False
    sm[[0, 396]] = False
    wr, ws, wsc = onet.superpoint_matching(rf, sf, rm, sm, 256)
    gr, gs, gsc = gm.SuperPointMatching(256)(rf.cuda(), sf.cuda(), rm.cuda(), sm.cuda())
    assert set(zip(gr.tolist(), gs.tolist())) == set(zip(wr.tolist(), ws.tolist()))
    assert rel_l2(gsc.cpu(), wsc) < 1e-5
    assert torch.equal(gr.cpu(), wr) and torch.equal(gs.cpu(), ws)
    # fewer valid pairs than k
    wr, ws, wsc = onet.superpoint_matching(rf[:9], sf[:7], rm[:9], sm[:7], 256)
    gr, gs, gsc = gm.SuperPointMatching(256)(rf[:9].cuda(), sf[:7].cuda(), rm[:9].cuda(), sm[:7].cuda())
    assert gr.shape == wr.shape and set(zip(gr.tolist(), gs.tolist())) == set(zip(wr.tolist(), ws.tolist()))


def _patch_problem(P=64, K=128, seed=6):
    g = torch.Generator().manual_seed(seed)
    ref_pts = torch.rand(P, K, 3, generator=g)
    ang = 0.3
    R = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32)
    perm = torch.stack([torch.randperm(K, generator=g) for _ in range(P)])
    src_pts = torch.gather((ref_pts - 0.1) @ R, 1, perm[..., None].expand(-1, -1, 3)) + 0.002 * torch.randn(P, K, 3, generator=g)
    rm = torch.rand(P, K, generator=g) > 0.15
    sm = torch.rand(P, K, generator=g) > 0.15
    rf = torch.randn(P, K, 32, generator=g)
    sf = torch.gather(rf, 1, perm[..., None].expand(-1, -1, 32)) + 0.3 * torch.randn(P, K, 32, generator=g)
    scores = torch.einsum("bnd,bmd->bnm", rf, sf) / 32 ** 0.5
    return ref_pts, src_pts, rm, sm, scores


def test_sinkhorn_vs_oracle():
    _, _, rm, sm, scores = _patch_problem()
    alpha = torch.tensor(1.0)
    want = onet.log_optimal_transport(scores, rm, sm, alpha, 100)
    got = ops.sinkhorn(scores.cuda(), rm.cuda(), sm.cuda(), alpha.cuda(), 100).cpu()
    valid = want > -1e11
    assert torch.equal(valid, got > -1e11)
    err_log, err_exp = float((got[valid] - want[valid]).abs().max()), rel_l2(got[valid].exp(), want[valid].exp())
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sinkhorn_parity.txt", "w") as f:
        f.write(f"max |log P - log P_oracle| = {err_log:.3e}, rel-L2 of P = {err_exp:.3e}\n")
    # rel-L2 <= 1e-5 on the transport plan; the log-domain values agree to 1e-4 absolute
    assert err_exp < 1e-5 and err_log < 1e-4, (err_log, err_exp)


def test_sinkhorn_large_magnitude_scores_stay_finite():
    """Scores of magnitude ~100 (trained weights can produce them): the log-sum-exp shift taken from the previous
    iterate must not underflow a whole line (guarded: the line is redone with its true maximum)."""
    g = torch.Generator().manual_seed(3)
    P, K = 8, 128
    scores = torch.randn(P, K, K, generator=g) * 60.0
    scores[0] = 100.0 * torch.eye(K) - 50.0          # one dominant entry per line, everything else far below
    scores[1, :, :] = -100.0
    scores[1, torch.arange(K), torch.arange(K).flip(0)] = 100.0
    rm = torch.ones(P, K, dtype=torch.bool)
    sm = torch.ones(P, K, dtype=torch.bool)
    rm[2, 100:] = False
    sm[3, :5] = False
    alpha = torch.tensor(1.0)
    want = onet.log_optimal_transport(scores, rm, sm, alpha, 100)
    got = ops.sinkhorn(scores.cuda(), rm.cuda(), sm.cuda(), alpha.cuda(), 100).cpu()
    valid = want > -1e11
    assert torch.equal(valid, got > -1e11)
    assert bool(torch.isfinite(got[valid]).all())
    assert float((got[valid] - want[valid]).abs().max()) < 5e-3  # |values| up to ~300: 1e-5 relative
    assert rel_l2(got[valid].exp(), want[valid].exp()) < 1e-4


def test_lgr_and_procrustes_vs_oracle():
    ref_pts, src_pts, rm, sm, scores = _patch_problem()
    ms = onet.log_optimal_transport(scores, rm, sm, torch.tensor(1.0), 100)
    cfg = {"fine_matching": dict(make_cfg().fine_matching)}
    taps = {}
    w_ref, w_src, w_sc, w_T = onet.local_global_registration(ref_pts, src_pts, rm, sm, ms[:, :-1, :-1], cfg, taps=taps)
    lgr = gm.LocalGlobalRegistration(3, 0.1, True, 0.05, False, False, 3, None, 5)
    g_ref, g_src, g_sc, g_T = [t.cpu() for t in lgr(ref_pts.cuda(), src_pts.cuda(), rm.cuda(), sm.cuda(), ms.cuda(), None)]
    assert g_sc.shape == w_sc.shape and w_sc.shape[0] > 500
    assert torch.equal(g_ref, w_ref) and torch.equal(g_src, w_src)
    assert rel_l2(g_sc, w_sc) < 1e-5
    assert float((g_T - w_T).norm()) < 1e-4, (g_T, w_T)
    # batched Procrustes incl. planar (rank-2), collinear-ish and zero-weight problems
    g = torch.Generator().manual_seed(8)
    src = torch.randn(40, 50, 3, generator=g)
    src[:10, :, 2] = 0.0  # planar
    Rz = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    ref = src @ Rz.t() + torch.tensor([0.5, -0.25, 1.0]) + 0.01 * torch.randn(40, 50, 3, generator=g)
    ref[:10, :, 2] = 1.0
    w = torch.rand(40, 50, generator=g)
    w[20:25, 10:] = 0.0
    want = onet.weighted_procrustes(src, ref, w)
    got = ops.weighted_procrustes(src.cuda(), ref.cuda(), w.cuda()).cpu()
    assert float((got - want).abs().max()) < 2e-5


# ---------------------------------------------------------------------------------------------- end to end
TAP_BLOCKS = ["encoder1_1", "encoder1_2", "encoder2_1", "encoder2_3", "encoder3_3", "encoder4_3", "encoder5_3",
              "decoder4", "decoder3", "decoder2"]


def _run_gpu_model(spec):
    model = seeded_model(0).cuda()
    d = make_pair_inputs(**spec)
    dd = {k: d[k] for k in ("ref_points", "src_points", "ref_feats", "src_feats")}
    cfg = make_cfg()
    data = registration_collate_fn_stack_mode([dd], cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                              cfg.backbone.init_radius, NEIGHBOR_LIMITS)
    out = model(data)  # the product path: the backbone is ONE native call (csrc/backbone.cu)
    # taps per block come from the per-module path, which must agree with the native call bit for bit
    taps = {}
    hooks = [getattr(model.backbone, n).register_forward_hook(lambda m, i, o, n=n: taps.__setitem__(n, o)) for n in TAP_BLOCKS]
    by_module = model.backbone.forward_modules(data["features"], data)
    for h in hooks:
        h.remove()
    native = model.backbone(data["features"], data)
    torch.cuda.synchronize()
    # The encoder is the same kernel sequence in both paths: bit-identical.  The decoders differ by ONE reassociation in the
    # native call (cat[up(c), s] W^T evaluated as up(c W_lo^T) + s W_hi^T, gr_unary_weights.split_k): fp32 rounding only.
    n_equal = 0
    for a, b in zip(native, by_module):
        if torch.equal(a, b):
            n_equal += 1
        else:
            rel = float((a - b).norm() / b.norm())
            assert rel < 3e-6, rel
    assert n_equal >= 1
    return model, data, out, taps


@pytest.mark.parametrize("case", ["room5k", "textured3k"])
def test_full_forward_vs_oracle_and_reference_golden(case):
    gold = np.load(os.path.join(GOLDEN_DIR, f"network_golden_{case}.npz"))
    spec = golden_spec(gold)
    model, data, out, taps = _run_gpu_model(spec)
    assert np.array_equal(torch.stack(data["lengths"]).cpu().numpy(), gold["lengths"])
    odata = oracle_data(spec)
    otaps = {}
    with torch.no_grad():
        want = onet.forward(seeded_model(0).state_dict(), odata, taps=otaps)
    report = {}
    for n in TAP_BLOCKS:
        report[n] = rel_l2(taps[n].cpu(), otaps[n])
        assert report[n] < 2e-4, (n, report)
    report["ref_feats_c"] = rel_l2(out["ref_feats_c"].cpu(), want["ref_feats_c"])
    report["src_feats_c"] = rel_l2(out["src_feats_c"].cpu(), want["src_feats_c"])
    assert report["ref_feats_c"] < 2e-4 and report["src_feats_c"] < 2e-4, report
    got_pairs = set(zip(out["ref_node_corr_indices"].tolist(), out["src_node_corr_indices"].tolist()))
    want_pairs = set(zip(gold["ref_node_corr_indices"].tolist(), gold["src_node_corr_indices"].tolist()))
    report["pair_agreement"] = len(got_pairs & want_pairs) / max(len(want_pairs), 1)
    report["num_corr"] = (int(out["corr_scores"].shape[0]), int(gold["corr_scores"].shape[0]))
    T = out["estimated_transform"].cpu().numpy()
    report["T_err_vs_reference"] = float(np.linalg.norm(T - gold["estimated_transform"]))
    report["T_err_vs_oracle"] = float(np.linalg.norm(T - want["estimated_transform"].numpy()))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/e2e_parity_{case}.txt", "w") as f:
        f.write(repr(report) + "\n")
    assert report["pair_agreement"] >= 0.98, report
    # (a) Our LGR kernel is the reference's function of ITS OWN input: feed the CPU oracle the Sinkhorn output and the
    #     patch points / masks the GPU path produced, and compare transforms (tolerance 1e-4, north_star).
    cfgd = {"fine_matching": dict(make_cfg().fine_matching)}
    ms_gpu = out["matching_scores"].cpu()
    _, _, sc_o, T_tf = onet.local_global_registration(out["ref_node_corr_knn_points"].cpu(), out["src_node_corr_knn_points"].cpu(),
                                                       out["ref_node_corr_knn_masks"].cpu(), out["src_node_corr_knn_masks"].cpu(),
                                                       ms_gpu[:, :-1, :-1], cfgd)
    report["T_err_vs_oracle_LGR_on_gpu_inputs"] = float(np.linalg.norm(T - T_tf.numpy()))
    assert sc_o.shape[0] == out["corr_scores"].shape[0]
    with open(f"gpurun_out/e2e_parity_{case}.txt", "w") as f:
        f.write(repr(report) + "\n")
    # north_star tolerance, no escape hatch: 1e-4 Frobenius against the REFERENCE's golden transform, identical superpoint
    # pairs, identical correspondence count
    assert report["T_err_vs_reference"] < 1e-4 and report["pair_agreement"] == 1.0, report
    assert report["num_corr"][0] == report["num_corr"][1], report
    # Diagnostic (b): the CPU restatement of LGR fed the GPU path's own Sinkhorn output.  With the seeded RANDOM weights
    # some cases are ill-conditioned for the reference algorithm itself (rank-deficient per-patch hypotheses decide an
    # argmax): if this diagnostic misses, it must be because the reference algorithm moves by more than the tolerance
    # under fp32-rounding-sized (1e-6 relative) input noise -- shown below and reported as a visible xfail, never
    # silently accepted.  The golden comparison above has already passed at this point.
    if report["T_err_vs_oracle_LGR_on_gpu_inputs"] >= 1e-4:
        g = torch.Generator().manual_seed(0)
        moves = []
        for _ in range(6):
            noisy = ms_gpu[:, :-1, :-1] * (1.0 + 1e-6 * torch.randn(ms_gpu[:, :-1, :-1].shape, generator=g))
            rp, sp = out["ref_node_corr_knn_points"].cpu(), out["src_node_corr_knn_points"].cpu()
            rp = rp * (1.0 + 1e-6 * torch.randn(rp.shape, generator=g))
            sp = sp * (1.0 + 1e-6 * torch.randn(sp.shape, generator=g))
            _, _, _, T_n = onet.local_global_registration(rp, sp, out["ref_node_corr_knn_masks"].cpu(),
                                                          out["src_node_corr_knn_masks"].cpu(), noisy, cfgd)
            moves.append(float(np.linalg.norm(T_n.numpy() - T_tf.numpy())))
        report["reference_LGR_move_under_1e-6_noise"] = moves
        with open(f"gpurun_out/e2e_parity_{case}.txt", "w") as f:
            f.write(repr(report) + "\n")
        assert max(moves) > 1e-4, report  # otherwise the miss is ours
        pytest.xfail(f"golden transform matched to {report['T_err_vs_reference']:.1e}; the reference LGR restatement is unstable on "
                     f"this case (moves {max(moves):.2e} under 1e-6 input noise), diagnostic (b) = {report['T_err_vs_oracle_LGR_on_gpu_inputs']:.2e}")


# ---------------------------------------------------------------------------------------------- tcgen05 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (4096, 256, 256), (1000, 64, 480), (777, 128, 1920), (20000, 32, 480),
                                   (460, 256, 2048), (300, 512, 128), (5000, 2048, 512), (129, 200, 36), (60000, 64, 60)])
def test_gemm_tensor_core_3xtf32_vs_fp32(M, N, K):
    """tcgen05 kind::tf32 with the hi/lo split must agree with the fp32 FFMA kernel to fp32 rounding level."""
    from gaussreg_b200 import _lib
    L = _lib.lib()
    a, b = _rand(M, K, seed=11).cuda(), _rand(N, K, seed=12).cuda()
    bias, div, res = _rand(N, seed=13).cuda(), (torch.rand(M, generator=torch.Generator().manual_seed(4)) * 5 + 1).cuda(), _rand(M, N, seed=15).cuda()
    ref = a.double() @ b.double().t()
    try:
        L.gr_set_gemm_mode(0)
        simt = ops.gemm(a, b, True, bias=bias, alpha=0.5, row_div=div, residual=res, act="leaky_relu")
        L.gr_set_gemm_mode(1)
        tc = ops.gemm(a, b, True, bias=bias, alpha=0.5, row_div=div, residual=res, act="leaky_relu")
        plain = ops.gemm(a, b, True)
    finally:
        L.gr_set_gemm_mode(1)
    torch.cuda.synchronize()
    # measured: 1e-6 .. 4e-6 (the tensor core's fp32 accumulation truncates; K is sliced at 256 and the slices
    # are summed in fp32), vs ~2e-7 for the FFMA kernel
    assert rel_l2(plain.cpu(), ref.cpu()) < 5e-6
    assert rel_l2(tc.cpu(), simt.cpu()) < 5e-6


@pytest.mark.parametrize("M,N,K", [(4096, 256, 256), (5000, 384, 480), (3000, 64, 960), (2500, 2048, 512), (2048, 96, 36), (60000, 32, 480)])
def test_linear_with_packed_weights(M, N, K):
    """Static weights pre-packed into the tensor-core operand format (bulk-copied B tiles) == plain product."""
    x, w, b = _rand(M, K, seed=21).cuda(), _rand(N, K, seed=22).cuda(), _rand(N, seed=23).cuda()
    div = (torch.rand(M, generator=torch.Generator().manual_seed(4)) * 3 + 1).cuda()
    ref = (x.double() @ w.double().t()) / div.double()[:, None] + b.double()
    got = ops.linear(x, w, bias=b, row_div=div)
    assert rel_l2(got.cpu(), ref.cpu()) < 5e-6
