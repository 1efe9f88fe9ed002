import sys, torch
sys.path.insert(0, "/root/repo")
import rmcl_b200
from rmcl_b200 import ops
dev="cuda"
g=torch.Generator().manual_seed(0)
for (B,C,K) in [(128,128,65536),(256,256,65536),(128,64,65536)]:
    q=torch.randn(B,C,generator=g).bfloat16().to(dev); k=torch.randn(B,C,generator=g).bfloat16().to(dev)
    queues=[torch.nn.functional.normalize(torch.randn(C,K,generator=g),dim=0).bfloat16().to(dev) for _ in range(10)]
    ops.profile_enable(True)
    acc={"prep":0,"partial":0,"finalize":0}
    n=40
    for i in range(n+5):
        ops.infonce_fwd_bwd(q,k,queues[i%10],0.07,normalize_k=True,path="tcgen05",want=("loss","dq","k_hat"))
        torch.cuda.synchronize()
        st=ops.profile_infonce_ms()
        if i>=5:
            for kk in acc: acc[kk]+=st[kk]*1000/n
    ops.profile_enable(False)
    print(B,C,K,{kk:round(v,1) for kk,v in acc.items()}, flush=True)
