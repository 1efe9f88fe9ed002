"""Barlow-Twins loss under data parallelism (batch 128 per rank, projector 8192): the reference's dataflow — per-rank
c = q.T @ k / bs, all_reduce(c) of the 268 MB matrix (objectives.py:480-482), loss, autograd backward — against the
all-gather of the two [B, D] projections + the fused Gram-path kernels (ops.barlow_twins_loss with dist.Gather).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/barlow_dist_time.py
"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
from rmcl_b200.dist import Gather

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
B, D, lam = 128, 8192, 0.0051
bs = world * B
g = torch.Generator(device=dev).manual_seed(10 + rank)
k = torch.randn(B, D, device=dev, generator=g)
q0 = 0.7 * k + 0.7 * torch.randn(B, D, device=dev, generator=g)


def reference():
    q = q0.detach().requires_grad_(True)
    c = q.T @ k
    c.div_(bs)
    dist.all_reduce(c)
    on_diag = torch.diagonal(c).add_(-1).pow_(2).sum()
    n = c.shape[0]
    off_diag = c.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten().pow_(2).sum()
    (on_diag + lam * off_diag).backward()
    return (on_diag + lam * off_diag).detach(), q.grad


gather = Gather()


def fused():
    q = q0.detach().requires_grad_(True)
    on, offs = ops.barlow_twins_loss(q, k, 1.0 / bs, lam, gather)
    (on + offs).backward()
    return (on + offs).detach(), q.grad


def timed(fn, n):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1000], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


lr, gr = reference()
lf, gf = fused()
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
us_ref, us_fused = timed(reference, 5), timed(fused, 20)
if rank == 0:
    print(json.dumps({"n_gpus": world, "batch_per_gpu": B, "D": D, "reference_dataflow_us": round(us_ref, 1), "fused_us": round(us_fused, 1),
                      "speedup": round(us_ref / us_fused, 1), "loss_rel": rel(lf, lr), "dq_rel": rel(gf, gr),
                      "exchange": "all-reduce of %.0f MB vs all-gather of %.1f MB" % (D * D * 4 / 1e6, 2 * world * B * D * 4 / 1e6)}), flush=True)
dist.destroy_process_group()
