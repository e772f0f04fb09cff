import sys, torch, time
sys.path.insert(0, '.')
from gaussreg_b200 import ops, _lib
L = _lib.lib()
shapes = [(262144,256,256),(786432,256,256),(60000,64,60),(60000,32,480),(60000,128,32),(60000,128,64),(42000,32,480),(42000,64,960),(42000,256,64),(15400,128,1920),(15400,512,128),(3700,256,3840),(3700,1024,256),(967,512,7680),(967,2048,512),(967,256,2048),(3700,1024,3072),(15400,512,1536),(42000,256,768)]
for (M,N,K) in shapes:
    a=torch.randn(M,K,device='cuda'); b=torch.randn(N,K,device='cuda'); out=torch.empty(M,N,device='cuda')
    res=[]
    for mode in (0,1):
        L.gr_set_gemm_mode(mode)
        for _ in range(2): ops.gemm(a,b,True,out=out)
        torch.cuda.synchronize()
        s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): ops.gemm(a,b,True,out=out)
        e.record(); torch.cuda.synchronize()
        ms=s.elapsed_time(e)/5
        res.append((ms, 2*M*N*K/ms/1e9))
    print(f"{M:7d} {N:5d} {K:5d}  simt {res[0][0]:8.3f} ms {res[0][1]:7.1f} TF/s | tc {res[1][0]:8.3f} ms {res[1][1]:7.1f} TF/s")
