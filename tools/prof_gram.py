"""Target of the ncu capture of the Gram-path Barlow-Twins kernels: B128 and B1024 at D8192, two calls each."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
g = torch.Generator().manual_seed(0)
for B in (128, 1024):
    k = torch.randn(B, 8192, generator=g).cuda()
    q = 0.7 * k + 0.7 * torch.randn(B, 8192, generator=g).cuda()
    for _ in range(2):
        r = ops.barlow_fwd_bwd(q, k, 1.0 / B, 0.0051, path="gram")
    torch.cuda.synchronize()
    print(B, float(r["loss"]))
