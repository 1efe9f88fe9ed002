"""GPU, world_size 2 over NCCL (skipped with fewer than two devices): the data-parallel exchange of the
path — key all-gather in rank order, then the identical on-device enqueue on every replica — checked
against the oracle, plus the silent skip on a short gathered batch (objectives.py:242-243)."""
import os
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    import rmcl_b200
    from rmcl_b200 import objectives
    B, C, K = 16, 64, 256
    dev = torch.device("cuda", rank)
    g = torch.Generator().manual_seed(100 + rank)
    k_local = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1).to(dev)
    queue0 = torch.randn(C, K, generator=torch.Generator().manual_seed(0))

    class Mod:  # the attributes dequeue_and_enqueue reads (objectives.py:238-248)
        pass
    m = Mod()
    m.proj_queue, m.proj_queue_ptr = queue0.to(dev), torch.tensor([K - world * B], dtype=torch.int64, device=dev)
    m.per_step_bs = world * B
    objectives.dequeue_and_enqueue(m, k_local)          # gather + enqueue: pointer wraps to 0
    m.per_step_bs = world * B + 1                        # short/odd batch: silently skipped
    objectives.dequeue_and_enqueue(m, k_local)
    torch.cuda.synchronize()
    gathered = rmcl_b200.concat_all_gather(k_local)
    torch.save({"queue": m.proj_queue.cpu(), "ptr": m.proj_queue_ptr.item(), "gathered": gathered.cpu(),
                "k_local": k_local.cpu()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_gather_and_replicated_enqueue_world2():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rmcl_oracle as O
    world, B, C, K = 2, 16, 64, 256
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, os.path.join(d, "pg"), d), nprocs=world, join=True)
        r = [torch.load(os.path.join(d, f"r{i}.pt")) for i in range(world)]
    assert torch.equal(r[0]["gathered"], r[1]["gathered"])
    assert torch.equal(r[0]["gathered"], O.concat_all_gather([r[0]["k_local"], r[1]["k_local"]]))   # rank order
    queue0 = torch.randn(C, K, generator=torch.Generator().manual_seed(0))
    want_q, want_p = O.dequeue_and_enqueue(queue0, K - world * B, r[0]["gathered"], K, per_step_bs=world * B)
    for x in r:
        assert x["ptr"] == want_p == 0
        assert torch.equal(x["queue"], want_q)          # replicas bit-identical, no broadcast needed


def _barlow_worker(rank, world, init_file, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    import rmcl_b200
    from rmcl_b200 import ops
    from rmcl_b200.dist import Gather
    B, D, lam = 128, 1024, 0.0051
    dev = torch.device("cuda", rank)
    g = torch.Generator().manual_seed(300 + rank)
    k = torch.randn(B, D, generator=g)
    q0 = 0.7 * k + 0.7 * torch.randn(B, D, generator=g)
    q = q0.clone().to(dev).requires_grad_(True)
    on, offs = ops.barlow_twins_loss(q, k.to(dev), 1.0 / (world * B), lam, Gather())
    (on + offs).backward()
    torch.cuda.synchronize()
    torch.save({"on": on.detach().cpu(), "offs": offs.detach().cpu(), "dq": q.grad.cpu(), "q": q0, "k": k},
               os.path.join(out_dir, f"b{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_barlow_twins_gathered_world2():
    """Barlow-Twins loss over two ranks: NCCL all-gather of the projections + the fused kernel on the gathered
    batch (256 rows) against the reference dataflow (per-rank products, all-reduced matrix, per-rank backward)."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rmcl_oracle as O
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_barlow_worker, args=(world, os.path.join(d, "pg"), d), nprocs=world, join=True)
        r = [torch.load(os.path.join(d, f"b{i}.pt")) for i in range(world)]
    ref = O.barlow_twins([x["q"].bfloat16().double() for x in r], [x["k"].bfloat16().double() for x in r], 256, 0.0051)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    for i in range(world):
        assert rel(r[i]["on"], ref["on_diag"]) < 1e-4 and rel(r[i]["offs"], 0.0051 * ref["off_diag"]) < 1e-4
        assert rel(r[i]["dq"], ref["dq"][i]) < 1e-2
    assert torch.equal(r[0]["on"], r[1]["on"])          # every rank evaluates the identical gathered problem


def _p2p_worker(rank, world, init_file, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    import rmcl_b200
    from rmcl_b200 import ops
    from rmcl_b200.dist import P2PKeyExchange
    B, C, K, steps = 32, 96, 512, 10                     # 10 steps of 64 keys: the ring wraps, both staging slots are reused
    ex = P2PKeyExchange(B, C, dev)
    queue = torch.randn(C, K, generator=torch.Generator().manual_seed(0)).to(dev)
    ref_queue = queue.clone()
    ptr = torch.tensor([K - 2 * world * B], dtype=torch.int64, device=dev)
    ref_ptr = ptr.clone()
    shadow = ops.QueueShadow()
    shadow.get(queue)
    g = torch.Generator().manual_seed(500 + rank)
    for s in range(steps):
        keys = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1).to(dev)
        ex.enqueue_(queue, keys, ptr, shadow=shadow)                               # ONE kernel: push, signal, wait, enqueue
        ops.enqueue_(ref_queue, rmcl_b200.concat_all_gather(keys), ref_ptr)        # NCCL all-gather + enqueue kernel
        if rank == 0 and s % 3 == 0:
            torch.cuda._sleep(20_000_000)                                          # skew the ranks: the flags must hold them together
    torch.cuda.synchronize()
    ok = torch.equal(queue, ref_queue) and int(ptr.item()) == int(ref_ptr.item())
    ok_shadow = torch.equal(shadow.get(queue), queue.bfloat16())
    # the facade's opt-in (pl_module.rmcl_p2p_exchange): same result as its NCCL branch, short batches skipped on both
    from rmcl_b200 import objectives

    class Mod:
        pass
    a, b = Mod(), Mod()
    for m_, p2p in ((a, True), (b, False)):
        m_.proj_queue, m_.proj_queue_ptr = queue.clone(), ptr.clone()
        m_.per_step_bs, m_.rmcl_p2p_exchange, m_.rmcl_bf16_queue = world * B, p2p, False
    for s in range(3):
        keys = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1).to(dev)
        for m_ in (a, b):
            m_.per_step_bs = world * B + (1 if s == 1 else 0)          # step 1: gathered batch != per_step_bs -> skipped
            objectives.dequeue_and_enqueue(m_, keys)
    torch.cuda.synchronize()
    ok = ok and torch.equal(a.proj_queue, b.proj_queue) and int(a.proj_queue_ptr.item()) == int(b.proj_queue_ptr.item())
    ok = ok and int(a.proj_queue_ptr.item()) == (int(ptr.item()) + 2 * world * B) % K
    torch.save({"queue": queue.cpu(), "ptr": int(ptr.item()), "ok": ok, "ok_shadow": ok_shadow}, os.path.join(out_dir, f"p{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_p2p_gather_enqueue_matches_nccl_path_world2():
    """rmcl_gather_enqueue_p2p (one kernel per rank over NVLink peer memory) against the NCCL all-gather + enqueue pair:
    queue, bf16 shadow and pointer bit-identical on every rank, over 10 skewed steps with ring wrap-around."""
    import torch.multiprocessing as mp
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_p2p_worker, args=(world, os.path.join(d, "pg"), d), nprocs=world, join=True)
        r = [torch.load(os.path.join(d, f"p{i}.pt")) for i in range(world)]
    for x in r:
        assert x["ok"] and x["ok_shadow"]
    assert torch.equal(r[0]["queue"], r[1]["queue"]) and r[0]["ptr"] == r[1]["ptr"]
