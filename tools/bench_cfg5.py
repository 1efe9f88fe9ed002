"""BASELINE configs[4] — InfoNCE scaling stress: dim 768, queue 262144 (bf16, replicated), global batch 4096 =
512 per GPU at 8 B200.  One step per rank = fused InfoNCE fwd+bwd (two-pass tcgen05) -> NCCL all-gather of the
normalised keys -> ring-buffer enqueue of the gathered keys (identical on every rank).  Weak scaling.

    python tools/bench_cfg5.py                                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cfg5.py
"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200  # noqa: F401
from rmcl_b200 import ops

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B, C, K, tau = 512, 768, 262144, 0.07
steps, warmup = 50, 5
g = torch.Generator(device=dev).manual_seed(100 + rank)
gq = torch.Generator(device=dev).manual_seed(7)
queue = torch.nn.functional.normalize(torch.randn(C, K, device=dev, generator=gq), dim=0).bfloat16()
ptr = torch.zeros(1, dtype=torch.int64, device=dev)
q = torch.randn(B, C, device=dev, generator=g).bfloat16()
k = torch.randn(B, C, device=dev, generator=g).bfloat16()
gathered = torch.empty(world * B, C, dtype=torch.float32, device=dev)
p2p = None
if world > 1 and "--nccl" not in sys.argv:       # fused peer-memory exchange (rmcl_gather_enqueue_p2p) unless --nccl
    try:
        from rmcl_b200.dist import P2PKeyExchange
        p2p = P2PKeyExchange(B, C, dev)
    except Exception as e:      # noqa: BLE001
        sys.stderr.write(f"peer-memory exchange unavailable ({e}); using NCCL all-gather\n")


def step():
    r = ops.infonce_fwd_bwd(q, k, queue, tau, normalize_k=True, path="tcgen05", want=("loss", "dq", "k_hat"))
    keys = r["k_hat"]
    if p2p is not None:
        p2p.enqueue_(queue, keys, ptr)
        return r
    if world > 1:
        dist.all_gather_into_tensor(gathered, keys)
        keys = gathered
    ops.enqueue_(queue, keys, ptr)
    return r


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(warmup):
    r = step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    r = step()
e1.record()
barrier()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    qs = [torch.empty_like(queue[:, :4096]) for _ in range(world)]      # replicated queue stays identical
    dist.all_gather(qs, queue[:, :4096].contiguous())
    same = all(torch.equal(qs[0], x) for x in qs)
else:
    same = True
if rank == 0:
    per = ms / steps
    print(json.dumps({"workload": "cfg5: InfoNCE fwd+bwd B512/GPU C768 K262144 bf16 (two-pass tcgen05) + key all-gather + enqueue",
                      "n_gpus": world, "exchange": "p2p" if p2p is not None else ("nccl" if world > 1 else None), "global_batch": world * B, "steps": steps, "ms_per_step": per,
                      "value": world * steps / (ms * 1e-3), "unit": "rank-steps/s", "scaling": "weak",
                      "infonce_tflops_per_gpu": 4.0 * B * C * (K + 1) / (per * 1e-3) / 1e12,
                      "queues_identical_across_ranks": same, "ptr": int(ptr.item()), "loss": float(r["loss"])}), flush=True)
if world > 1:
    dist.destroy_process_group()
