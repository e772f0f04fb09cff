"""Config-3 throughput against the number of host worker threads (one stream each) and the host-side issue time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussreg_b200 import parallel
from gaussreg_b200.config import make_cfg
from gaussreg_b200.model import create_model
from gaussreg_b200.synthetic import make_pair_inputs

cfg = make_cfg(); torch.manual_seed(0); np.random.seed(0)
model = create_model(cfg).eval().cuda()
pool = [make_pair_inputs(1000 + i, 30000) for i in range(16)]
jobs = [pool[i % 16] for i in range(int(os.environ.get("PAIRS", "96")))]
parallel.register_pairs(model, jobs[:8], streams=1); torch.cuda.synchronize()
for w in [int(x) for x in os.environ.get("WORKERS", "1,2,3,4,6,8").split(",")]:
    kw = dict(streams=1) if w == 1 else dict(workers=w)
    parallel.register_pairs(model, jobs[:8], **kw); torch.cuda.synchronize()
    t0 = time.perf_counter()
    T = parallel.register_pairs(model, jobs, **kw).cpu()
    sec = time.perf_counter() - t0
    print(f"workers={w}: {len(jobs)/sec:7.1f} pairs/s  ({1e3*sec/len(jobs):.2f} ms per pair)", flush=True)
