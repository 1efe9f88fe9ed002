"""CPU, world_size 2 over gloo: the N>1 host logic — rank-ordered key gather, identical enqueue on
every replica, the short-batch skip — checked against the oracle."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    import rmcl_b200
    import rmcl_oracle as O
    from rmcl_b200.dist import concat_all_gather, gathered_batch_matches
    B, C, K = 4, 8, 32
    g = torch.Generator().manual_seed(100 + rank)
    k_local = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1)
    gathered = concat_all_gather(k_local)
    assert gathered.shape == (world * B, C)
    assert torch.equal(gathered[rank * B:(rank + 1) * B], k_local)
    # every rank applies the same enqueue to its replica (oracle stands in for the CUDA kernel on CPU)
    queue = torch.randn(C, K, generator=torch.Generator().manual_seed(0))
    q1, p1 = O.dequeue_and_enqueue(queue, 8, gathered, K, per_step_bs=world * B)
    assert p1 == 16
    # short batch -> skipped (objectives.py:242-243)
    assert not gathered_batch_matches(world * B + 1, gathered.shape[0])
    assert gathered_batch_matches(None, gathered.shape[0]) and gathered_batch_matches(world * B, gathered.shape[0])
    torch.save({"gathered": gathered, "queue": q1, "ptr": p1}, os.path.join(out_dir, f"r{rank}.pt"))
    # a second dtype / higher rank tensor
    t = torch.arange(6, dtype=torch.bfloat16).view(1, 2, 3) + rank
    gt = concat_all_gather(t)
    assert gt.shape == (world, 2, 3) and gt[1, 0, 0].item() == 1.0
    dist.destroy_process_group()


def test_key_gather_and_replicated_enqueue_world2():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "pg")
        mp.spawn(_worker, args=(world, init_file, d), nprocs=world, join=True)
        r0, r1 = torch.load(os.path.join(d, "r0.pt")), torch.load(os.path.join(d, "r1.pt"))
    assert torch.equal(r0["gathered"], r1["gathered"])
    assert torch.equal(r0["queue"], r1["queue"]) and r0["ptr"] == r1["ptr"]
    # rank order: rank 0's keys first
    g0 = torch.nn.functional.normalize(torch.randn(4, 8, generator=torch.Generator().manual_seed(100)), dim=1)
    assert torch.equal(r0["gathered"][:4], g0)


def test_single_process_gather_is_identity():
    from rmcl_b200.dist import concat_all_gather
    t = torch.randn(3, 5)
    assert concat_all_gather(t) is t


def _barlow_worker(rank, world, init_file, out_dir):
    """The Barlow-Twins exchange on the host side: the reference all-reduces c = q.T k / bs (objectives.py:482);
    the CUDA path all-gathers q and k and evaluates the matrix from the gathered batch.  Here both are done with
    gloo and the oracle stands in for the kernel: same c, same loss, and the per-rank gradient rows line up."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    import rmcl_oracle as O
    from rmcl_b200.dist import Gather
    B, D, lam = 6, 24, 0.0051
    g = torch.Generator().manual_seed(200 + rank)
    k = torch.randn(B, D, generator=g, dtype=torch.float64)
    q = 0.7 * k + 0.7 * torch.randn(B, D, generator=g, dtype=torch.float64)
    gather = Gather()
    assert gather.rank == rank and gather.world == world
    qa, ka = gather(q), gather(k)
    assert torch.equal(qa[rank * B:(rank + 1) * B], q)
    mine = O.barlow_twins([qa], [ka], world * B, lam)                    # what the kernel computes (b0 = rank * B)
    c = q.T @ k / (world * B)                                            # what the reference computes
    dist.all_reduce(c)
    assert torch.allclose(mine["c"], c, rtol=1e-12, atol=1e-14)
    torch.save({"loss": mine["loss"], "dq_local": mine["dq"][0][rank * B:(rank + 1) * B], "q": q, "k": k},
               os.path.join(out_dir, f"b{rank}.pt"))
    dist.destroy_process_group()


def test_barlow_gather_replaces_allreduce_world2():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rmcl_oracle as O
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_barlow_worker, args=(world, os.path.join(d, "pg"), d), nprocs=world, join=True)
        r = [torch.load(os.path.join(d, f"b{i}.pt")) for i in range(world)]
    assert torch.equal(r[0]["loss"], r[1]["loss"])
    ref = O.barlow_twins([x["q"] for x in r], [x["k"] for x in r], 12, 0.0051)   # the reference's per-rank backward
    for i in range(world):
        assert torch.allclose(r[i]["dq_local"], ref["dq"][i], rtol=1e-12, atol=1e-14)
