"""The RPE-conditional transformer alone at the bench's superpoint counts: CUDA-event time per call (warm), and a
convenient target for `ncu --cache-control none` launch lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussreg_b200 import ops
from gaussreg_b200.config import make_cfg
from gaussreg_b200.model import create_model

cfg = make_cfg(); torch.manual_seed(0); np.random.seed(0)
model = create_model(cfg).eval().cuda()
tf = model.transformer
N0, N1 = int(os.environ.get("N0", "479")), int(os.environ.get("N1", "488"))
g = torch.Generator().manual_seed(1)
p0, p1 = torch.rand(N0, 3, generator=g).cuda() * 4, torch.rand(N1, 3, generator=g).cuda() * 4
f0, f1 = torch.randn(N0, 2048, generator=g).cuda(), torch.randn(N1, 2048, generator=g).cuda()
e0, e1 = tf.embedding(p0), tf.embedding(p1)
reps = int(os.environ.get("REPS", "20"))
for _ in range(3):
    out = tf(p0, p1, f0, f1, embeddings=(e0, e1))
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    out = tf(p0, p1, f0, f1, embeddings=(e0, e1))
e.record(); torch.cuda.synchronize()
print(f"transformer (in_proj + 6 layers + out_proj), N0={N0} N1={N1}: {s.elapsed_time(e)/reps:.3f} ms per call")
