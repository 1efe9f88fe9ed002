import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev="cuda"
g = torch.Generator().manual_seed(0)
B,C,K=256,256,65536
q = torch.randn(B, C, generator=g).bfloat16().to(dev); k = torch.randn(B, C, generator=g).bfloat16().to(dev)
queue = torch.randn(C, K, generator=g).bfloat16().to(dev)
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
for it in range(4):
    flush.zero_(); torch.cuda.synchronize()
    r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss","dq","k_hat"))
    torch.cuda.synchronize()
print("loss", r["loss"].item())
