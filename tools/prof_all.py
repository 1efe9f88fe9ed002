"""Short driver for ncu --set full: one call of every kernel at its bench shape (see profiles/README.md)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
import bench
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
shapes = bench.load_shapes()
pk = [torch.randn(s, device=dev, generator=g) for s in shapes]
pq = [torch.randn(s, device=dev, generator=g) for s in shapes]
plan = ops.EmaPlan(pk, pq)
B, C, K = 256, 256, 65536
q = torch.randn(B, C, device=dev, generator=g).bfloat16(); k = torch.randn(B, C, device=dev, generator=g).bfloat16()
queue = torch.randn(C, K, device=dev, generator=g).bfloat16()
ptr = torch.zeros(1, dtype=torch.int64, device=dev)
grad = torch.randn(128, 3, 384, 384, device=dev, generator=g); delta = torch.zeros_like(grad)
for it in range(3):
    ops.ema_multi_(plan, 0.999)
    r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss", "dq", "k_hat"))
    ops.enqueue_(queue, r["k_hat"], ptr)
    ops.pgd_step_(delta, grad, 0.05, 8 / 255, "ref_linf")
    torch.cuda.synchronize()
print("loss", r["loss"].item())
