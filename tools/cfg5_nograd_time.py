import sys, torch
sys.path.insert(0, "/root/repo")
import rmcl_b200
from rmcl_b200 import ops
dev="cuda"; g=torch.Generator().manual_seed(0)
B,C,K=512,768,262144
q=torch.randn(B,C,generator=g).bfloat16().to(dev); k=torch.randn(B,C,generator=g).bfloat16().to(dev)
queue=torch.nn.functional.normalize(torch.randn(C,K,generator=g),dim=0).bfloat16().to(dev)
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1000
print("cfg5 two-pass: with grad %.1f us, statistics only %.1f us" % (t(lambda: ops.infonce_fwd_bwd(q,k,queue,0.07,normalize_k=True)), t(lambda: ops.infonce_fwd_bwd(q,k,queue,0.07,normalize_k=True,need_grad=False,want=("argmax","k_hat","lse")))))
