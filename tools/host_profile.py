"""cProfile of the HOST side of one pair (tiny cloud -> the GPU is idle, the wall time is the Python issue time)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
from gaussreg_b200.data import precompute_data_stack_mode
from gaussreg_b200.model import create_model
from gaussreg_b200.synthetic import make_pair_inputs

cfg = make_cfg()
torch.manual_seed(0); np.random.seed(0)
model = create_model(cfg).eval().cuda()
n = 1500
d = make_pair_inputs(0, n)
pts = torch.from_numpy(np.concatenate([d["ref_points"], d["src_points"]])).cuda()
feats = torch.from_numpy(np.concatenate([d["ref_feats"], d["src_feats"]])).cuda()
lens = torch.tensor([n, n], dtype=torch.int64, device="cuda")
def step():
    data = precompute_data_stack_mode(pts, lens, cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius, NEIGHBOR_LIMITS)
    data["features"] = feats
    return model(data)["estimated_transform"]
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
