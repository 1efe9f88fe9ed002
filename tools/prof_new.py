"""One call each of the fused Barlow-Twins loss (B128 D8192) and the cfg5 two-pass InfoNCE (B512 C768 K262144) —
the target of the ncu captures of barlow_tc_kernel / infonce_s_kernel / infonce_pv_kernel."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
kb = torch.randn(128, 8192, generator=g).to(dev)
qb = 0.7 * kb + 0.7 * torch.randn(128, 8192, generator=g).to(dev)
B, C, K = 512, 768, 262144
q = torch.randn(B, C, generator=g).bfloat16().to(dev)
k = torch.randn(B, C, generator=g).bfloat16().to(dev)
queue = torch.nn.functional.normalize(torch.randn(C, K, generator=g), dim=0).bfloat16().to(dev)
torch.cuda.synchronize()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    r1 = ops.barlow_fwd_bwd(qb, kb, 1.0 / 128, 0.0051)
    r2 = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05")
torch.cuda.synchronize()
print("ok", float(r1["loss"]), float(r2["loss"]))
