"""Times gr_linear_packed (tcgen05 3xTF32, TMA-fed) on the backbone's product shapes: CUDA events over back-to-back
repetitions on rotating operand buffers larger than L2 (so that A really streams from HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussreg_b200 import ops, _lib

shapes = [(60000, 32, 480), (41907, 64, 960), (15432, 128, 1920), (3733, 256, 3840), (967, 512, 7680),
          (41907, 256, 64), (41907, 64, 256), (60000, 128, 64), (15432, 512, 128), (15432, 128, 512),
          (15432, 512, 1536), (3733, 1024, 3072), (41907, 256, 768), (967, 2048, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
tot = 0.0
for M, N, K in shapes:
    nbuf = max(2, int(300e6 // (M * K * 4)) + 1)
    A = [torch.randn(M, K, generator=g).to(dev) for _ in range(min(nbuf, 6))]
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    out = torch.empty(M, N, device=dev)
    for i in range(3):
        ops.linear(A[i % len(A)], W, out=out)
    torch.cuda.synchronize()
    reps = 12
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        ops.linear(A[i % len(A)], W, out=out)
    e.record(); torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / reps
    tot += us
    gf = 2.0 * M * N * K / 1e9
    gb = 4.0 * (M * K + M * N) / 1e9
    print(f"{M:6d} x {N:5d} x {K:5d}: {us:8.1f} us  {gf/us*1e3:7.1f} TFLOP/s fp32-eq  {gb/us*1e6/1e3:6.2f} TB/s (A+C)  path={_lib.lib().gr_last_gemm_path()}")
print(f"sum {tot:.1f} us")
