import sys, torch, os
sys.path.insert(0, '.')
from gaussreg_b200 import ops
for N in (256, 128, 64):
  for K in (32, 64, 128, 256, 512, 1024):
    M=148*128*8
    x=torch.randn(M,K,device='cuda'); w=torch.randn(N,K,device='cuda'); b=torch.randn(N,device='cuda')
    fn=lambda: ops.linear(x,w,bias=b)
    for _ in range(2): y=fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True); s.record()
    for _ in range(5): y=fn()
    e.record(); torch.cuda.synchronize(); ms=s.elapsed_time(e)/5
    ctas = (M//128)*max(1,N//256 if N>=256 else 1)
    per_cta_us = ms*1e3/ (ctas/148) / (2 if N==64 else 1)
    print(f"N={N:4d} K={K:5d} {ms:7.3f} ms  per-CTA-slot {ms*1e3/(ctas/148):7.2f} us  ({2*M*N*K/ms/1e9:6.1f} TF/s)")
