#!/usr/bin/env python
"""Times gr_kpconv_aggregate (+ the contraction) on the real layer shapes of one 30k+30k pair (CUDA events, L2 flushed)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gaussreg_b200 import ops
from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
from gaussreg_b200.data import precompute_data_stack_mode
from gaussreg_b200.synthetic import make_pair_inputs

dev = torch.device("cuda", 0)
cfg = make_cfg()
p = make_pair_inputs(0, int(os.environ.get("N", "30000")))
pts = torch.from_numpy(np.concatenate([p["ref_points"], p["src_points"]])).to(dev)
lens = torch.tensor([p["ref_points"].shape[0], p["src_points"].shape[0]], dtype=torch.int64, device=dev)
d = precompute_data_stack_mode(pts, lens, cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius, NEIGHBOR_LIMITS)
P, NB, SUB = d["points"], d["neighbors"], d["subsampling"]
layers = [("1_2", 32, P[0], P[0], NB[0]), ("2_1", 32, P[1], P[0], SUB[0]), ("2_2", 64, P[1], P[1], NB[1]), ("3_1", 64, P[2], P[1], SUB[1]),
          ("3_2", 128, P[2], P[2], NB[2]), ("4_1", 128, P[3], P[2], SUB[2]), ("4_2", 256, P[3], P[3], NB[3]), ("5_1", 256, P[4], P[3], SUB[3]),
          ("5_2", 512, P[4], P[4], NB[4])]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device="cpu").manual_seed(0)
kp = ((torch.rand(15, 3, generator=g) - 0.5) * 0.1).to(dev)
tot_a = tot_g = 0.0
for name, C, q, s, idx in layers:
    feats = torch.randn(s.shape[0], C, generator=g).to(dev)
    W = (torch.randn(15, C, C, generator=g) / (15 * C) ** 0.5).to(dev)
    Wk = ops._kmajor_weights(W)
    sigma = 0.05 * (q.shape[0] and 1)
    ta, tg = [], []
    for it in range(6):
        flush.fill_(it)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        A, rd = ops.kpconv_aggregate(feats, q, s, idx, kp, 0.05)
        e1.record()
        out = ops.linear(A, Wk, row_div=rd)
        e2.record()
        torch.cuda.synchronize()
        ta.append(e0.elapsed_time(e1)); tg.append(e1.elapsed_time(e2))
    a, b = sorted(ta[1:])[2], sorted(tg[1:])[2]
    tot_a += a; tot_g += b
    M, H = idx.shape
    print(f"{name}: C={C} M={M} H={H} Ns={s.shape[0]}  aggregate {a*1e3:7.1f} us   gemm {b*1e3:7.1f} us   A={M*15*C*4/1e6:.0f} MB")
print(f"total aggregate {tot_a:.3f} ms, gemm {tot_g:.3f} ms  (one layer per shape class; the model has 13)")
# raw memory rates for context
x = torch.empty(1 << 28, dtype=torch.float32, device=dev); y = torch.empty_like(x)
for nm, fn, nbytes in (("memset 1GB", lambda: x.zero_(), 4 * x.numel()), ("copy 1GB", lambda: y.copy_(x), 8 * x.numel())):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{nm}: {nbytes / e0.elapsed_time(e1) / 1e6:.0f} GB/s")
