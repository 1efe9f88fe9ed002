import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev="cuda"
g = torch.Generator().manual_seed(0)
for (B,C,K) in [(130,64,136),(64,256,520),(256,128,1000),(1024,128,4096)]:
    q = torch.randn(B, C, generator=g).to(dev); k = torch.randn(B, C, generator=g).to(dev)
    queue = torch.randn(C, K, generator=g).bfloat16().to(dev)
    r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05")
    torch.cuda.synchronize()
    print(B, C, K, "loss", r["loss"].item(), flush=True)
pk=[torch.randn(1000,device=dev), torch.randn(33,7,device=dev)]; pq=[torch.randn(1000,device=dev), torch.randn(33,7,device=dev)]
ops.ema_multi_(ops.EmaPlan(pk,pq),0.999)
qd=torch.randn(64,256,device=dev); pd=torch.zeros(1,dtype=torch.int64,device=dev)
ops.enqueue_(qd, torch.randn(32,64,device=dev), pd)
for mode in ("ref_linf","sign_linf","l2"):
    ops.pgd_step_(torch.zeros(5,1001,device=dev), torch.randn(5,1001,device=dev), 0.05, 0.03, mode)
    ops.pgd_step_(torch.zeros(3,4096,device=dev), torch.randn(3,4096,device=dev), 0.05, 0.03, mode)
torch.cuda.synchronize(); print("done")
