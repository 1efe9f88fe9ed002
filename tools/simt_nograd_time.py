import sys, torch
sys.path.insert(0, "/root/repo")
import rmcl_b200
from rmcl_b200 import ops
dev="cuda"; g=torch.Generator().manual_seed(0)
B,C,K=256,256,65536
q=torch.randn(B,C,generator=g).to(dev); k=torch.randn(B,C,generator=g).to(dev)
queue=torch.nn.functional.normalize(torch.randn(C,K,generator=g),dim=0).to(dev)
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1000
a=ops.infonce_fwd_bwd(q,k,queue,0.07,normalize_k=True,path="simt")
b=ops.infonce_fwd_bwd(q,k,queue,0.07,normalize_k=True,need_grad=False,path="simt",want=("argmax","k_hat","loss","lse"))
print("fp32 SIMT with grad %.1f us, without grad %.1f us; argmax equal %s lse equal %s" % (
  t(lambda: ops.infonce_fwd_bwd(q,k,queue,0.07,normalize_k=True,path="simt")),
  t(lambda: ops.infonce_fwd_bwd(q,k,queue,0.07,normalize_k=True,need_grad=False,path="simt",want=("argmax","k_hat"))),
  torch.equal(a["argmax"],b["argmax"]), torch.equal(a["lse"],b["lse"])))
