"""How long does the HOST need to issue one pair?  A tiny cloud makes the GPU work negligible (same launch count)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussreg_b200 import _lib
from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
from gaussreg_b200.data import precompute_data_stack_mode
from gaussreg_b200.model import create_model
from gaussreg_b200.synthetic import make_pair_inputs

cfg = make_cfg()
torch.manual_seed(0); np.random.seed(0)
model = create_model(cfg).eval().cuda()
for n in (1500, 30000):
    d = make_pair_inputs(0, n)
    pts = torch.from_numpy(np.concatenate([d["ref_points"], d["src_points"]])).cuda()
    feats = torch.from_numpy(np.concatenate([d["ref_feats"], d["src_feats"]])).cuda()
    lens = torch.tensor([n, n], dtype=torch.int64, device="cuda")
    def step():
        data = precompute_data_stack_mode(pts, lens, cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius, NEIGHBOR_LIMITS)
        data["features"] = feats
        return model(data)["estimated_transform"]
    for _ in range(3): step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    tp = tm = 0.0
    for _ in range(10):
        a = time.perf_counter()
        data = precompute_data_stack_mode(pts, lens, cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius, NEIGHBOR_LIMITS)
        data["features"] = feats
        b = time.perf_counter()
        T = model(data)["estimated_transform"]
        c = time.perf_counter()
        tp += b - a; tm += c - b
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n={n}: wall/step {1e2*(t2-t0):.2f} ms, host returns after {1e2*(t1-t0):.2f} ms/step (pyramid {1e2*tp:.2f} + model {1e2*tm:.2f}), launches/step {(_lib.launch_count()-l0)/10:.0f}")
