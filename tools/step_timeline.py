"""Host vs device timeline of one 30k+30k step: for every C-ABI call the host time spent inside the call, the device time
between its first and last kernel (CUDA events) and the host clock at which it was issued; plus the step's wall time."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussreg_b200 import _lib
from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
from gaussreg_b200.data import precompute_data_stack_mode
from gaussreg_b200.model import create_model
from gaussreg_b200.synthetic import make_pair_inputs

class Prof:
    def __init__(self, lib): self._lib, self.rec, self.on = lib, [], False
    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("gr_") or "workspace_size" in name or name in ("gr_last_error", "gr_launch_count", "gr_get_gemm_mode", "gr_last_gemm_path"):
            return fn
        def w(*a):
            if not self.on: return fn(*a)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); s.record(); r = fn(*a); e.record(); t1 = time.perf_counter()
            self.rec.append((name, t0, t1, s, e)); return r
        return w

cfg = make_cfg(); torch.manual_seed(0); np.random.seed(0)
model = create_model(cfg).eval().cuda()
lib = Prof(_lib.lib()); _lib._lib = lib
n = int(os.environ.get("N", "30000"))
d = make_pair_inputs(0, n)
pts = torch.from_numpy(np.concatenate([d["ref_points"], d["src_points"]])).cuda()
feats = torch.from_numpy(np.concatenate([d["ref_feats"], d["src_feats"]])).cuda()
lens = torch.tensor([n, n], dtype=torch.int64, device="cuda")
def step():
    data = precompute_data_stack_mode(pts, lens, cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius, NEIGHBOR_LIMITS,
                                      early=model.backbone.forward_early if os.environ.get('GAUSSREG_EARLY', '0') == '1' else None, features=feats)
    data["features"] = feats
    return model(data)["estimated_transform"]
for _ in range(5): step()
torch.cuda.synchronize()
# plain wall / device time of 10 steps
t0 = time.perf_counter(); s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(10): step()
e0.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"10 steps: wall {1e2*(t1-t0):.3f} ms/step, device {s0.elapsed_time(e0)/10:.3f} ms/step")
lib.on = True
tA = time.perf_counter(); sA = torch.cuda.Event(enable_timing=True); sA.record()
step(); torch.cuda.synchronize(); tB = time.perf_counter()
lib.on = False
agg = {}
print(f"profiled step wall {1e3*(tB-tA):.3f} ms; calls {len(lib.rec)}")
for name, a, b, s, e in lib.rec:
    g = agg.setdefault(name, [0, 0.0, 0.0]); g[0] += 1; g[1] += 1e3 * (b - a); g[2] += s.elapsed_time(e)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{k:36s} n={v[0]:3d} host {v[1]:7.3f} ms  device {v[2]:7.3f} ms")
print("timeline (ms since step start): name host_start host_end dev_start dev_end")
for name, a, b, s, e in lib.rec:
    print(f"  {name:34s} {1e3*(a-tA):7.3f} {1e3*(b-tA):7.3f} {sA.elapsed_time(s):7.3f} {sA.elapsed_time(e):7.3f}")
