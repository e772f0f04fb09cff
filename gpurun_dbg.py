import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
from gaussreg_b200.data import precompute_data_stack_mode
from gaussreg_b200.model import create_model
from gaussreg_b200.synthetic import make_pair_inputs
cfg = make_cfg(); torch.manual_seed(0); np.random.seed(0)
model = create_model(cfg).eval().cuda()
p = make_pair_inputs(0, 30000)
pts = torch.from_numpy(np.concatenate([p['ref_points'], p['src_points']])).cuda(); feats = torch.from_numpy(np.concatenate([p['ref_feats'], p['src_feats']])).cuda()
lens = torch.tensor([30000,30000], device='cuda')
def step():
    data = precompute_data_stack_mode(pts, lens, 5, 0.025, 0.0625, NEIGHBOR_LIMITS); data['features']=feats
    return data, model(data)
for _ in range(3): step()
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5): d,o = step()
t_issue=(time.perf_counter()-t)/5
torch.cuda.synchronize(); t_all=(time.perf_counter()-t)/5
print('wall per step %.2f ms (issue-side %.2f ms)'%(t_all*1e3, t_issue*1e3))
# phases
def timeit(fn, n=5):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): r=fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
print('pyramid %.2f ms'%timeit(lambda: precompute_data_stack_mode(pts, lens, 5, 0.025, 0.0625, NEIGHBOR_LIMITS)))
data,_=step()
print('backbone %.2f ms'%timeit(lambda: model.backbone(feats, data)))
import cProfile, pstats
pr=cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
