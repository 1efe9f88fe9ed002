"""Short driver for ncu (round 2): a few warm-up rounds, then one call of every kernel of the path at its bench shape, in a
fixed order (see profiles/README.md for the command lines and tools/make_traffic.py for the extraction)."""
import os, sys, torch
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
q4 = torch.randn(128, 128, device=dev, generator=g); k4 = torch.randn(128, 128, device=dev, generator=g)
queue4 = torch.randn(128, K, device=dev, generator=g)                      # fp32 queue, cfg4 shape: split-operand path
stats = None
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for it in range(rounds):
    ops.ema_multi_(plan, 0.999, check_storage=False)
    r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, want=("loss", "dq", "k_hat"))
    ops.enqueue_(queue, r["k_hat"], ptr)
    ops.pgd_step_(delta, grad, 0.05, 8 / 255, "ref_linf")
    ops.pgd_step_(delta, grad, 0.5, 1.0, "l2")
    ops.pgd_step_(delta, grad, 2 / 255, 8 / 255, "sign_linf")
    r4 = ops.infonce_fwd_bwd(q4, k4, queue4, 0.07, normalize_k=True, want=("loss", "dq", "k_hat"))
    stats = ops.QueueStats(queue4)
    torch.cuda.synchronize()
print("loss", r["loss"].item(), r4["loss"].item())
