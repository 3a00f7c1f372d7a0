import sys, torch
sys.path.insert(0,'/root/repo/flashattention.jl_b200')
import fa_sm100a as fa
bf=torch.bfloat16
for (X,Y,d,B,W) in ((128,128,64,8,7),(64,64,64,8,7),(256,256,64,4,13)):
    q,k,v=(fa.jl_empty((X,Y,d,B),bf).normal_() for _ in range(3))
    O,l,m=fa.circulant_fa(q,k,v,W)
    g=torch.randn_like(O)
    for name,fn in (("fwd",lambda: fa.circulant_fa(q,k,v,W)),("bwd",lambda: fa.circulant_fa_backward(q,k,v,O,g,l,m,W))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        t=e0.elapsed_time(e1)/5
        fl=4.0*X*Y*W*W*d*B*(2.5 if name=="bwd" else 1)
        by=(4 if name=="fwd" else 8)*X*Y*d*B*2
        print(X,Y,d,B,W,name,"ms",round(t,4),"GFLOP/s",round(fl/t/1e6,1),"alg GB/s",round(by/t/1e6,1),flush=True)
