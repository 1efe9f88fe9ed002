"""BASELINE configs[3] / SURVEY 8(d) cfg4 — the full RMCL pretraining step around the kernels, data parallel:

    ViLT-B/32-shaped backbone (random init, torch, bf16 autocast; 12 x 768, patch 32, 384^2 images -> 144 patches + cls,
    40 text tokens = 185 tokens), 128 samples per GPU, queue 128 x 65536, tau 0.07, m 0.999, 3 PGD steps (eps 8/255,
    fp32 as in the reference, pgd_attack_vilt.py:141) ->
        rmcl_b200.compute_moco_contrastive(pl_module, batch)      # EMA, key fwd, clean fwd, PGD (3 fwd+bwd), attacked fwd,
                                                                  # fused InfoNCE, key all-gather, enqueue
        loss.backward(); flat NCCL all-reduce of the gradients; fused AdamW

The backbone GEMMs are torch's (north_star keeps them there); this script measures what the step costs end to end and
what share of it the hand-written kernels are, at 1/2/4/8 GPUs (weak scaling).

    python tools/full_step.py [--steps 10] [--batch 128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/full_step.py
"""
import argparse
import json
import os
import sys
import types

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200  # noqa: E402
from rmcl_b200 import ops  # noqa: E402


class Block(nn.Module):
    def __init__(self, dim=768, heads=12, mlp=3072):
        super().__init__()
        self.heads = heads
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)
        self.fc1, self.fc2 = nn.Linear(dim, mlp), nn.Linear(mlp, dim)

    def forward(self, x, mask=None):          # the synthetic batch has no padding: the mask is all ones
        B, N, C = x.shape
        qkv = self.qkv(self.norm1(x)).reshape(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        x = x + self.proj(a.transpose(1, 2).reshape(B, N, C))
        x = x + self.fc2(F.gelu(self.fc1(self.norm2(x))))
        return x, None


class VitB32(nn.Module):
    """vilt/modules/vision_transformer.py 'vit_base_patch32_384' shapes: patch 32, 12 blocks, 768 wide."""

    def __init__(self, dim=768, depth=12, img=384, patch=32):
        super().__init__()
        self.patch_embed = nn.Conv2d(3, dim, patch, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, (img // patch) ** 2 + 1, dim) * 0.02)
        self.blocks = nn.ModuleList([Block(dim) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)

    def visual_embed(self, img, max_image_len=200, mask_it=False):
        x = self.patch_embed(img).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1).to(x.dtype), x], dim=1) + self.pos_embed.to(x.dtype)
        return x, torch.ones(x.shape[0], x.shape[1], dtype=torch.long, device=x.device), None, None


class BertEmbeddings(nn.Module):
    def __init__(self, vocab=30522, dim=768, max_len=40):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, dim)
        self.position_embeddings = nn.Embedding(max_len, dim)
        self.token_type_embeddings = nn.Embedding(2, dim)
        self.LayerNorm = nn.LayerNorm(dim)

    def forward(self, ids):
        pos = torch.arange(ids.shape[1], device=ids.device)[None, :]
        return self.LayerNorm(self.word_embeddings(ids) + self.position_embeddings(pos) + self.token_type_embeddings(torch.zeros_like(ids)))


class Pooler(nn.Module):
    def __init__(self, dim=768):
        super().__init__()
        self.dense = nn.Linear(dim, dim)

    def forward(self, x):
        return torch.tanh(self.dense(x[:, 0]))


class MOCOHead(nn.Module):       # vilt/modules/heads.py:129-143
    def __init__(self, i=768, h=768, o=128):
        super().__init__()
        self.projector = nn.Sequential(nn.Linear(i, h), nn.LayerNorm(h), nn.ReLU(), nn.Linear(h, o, bias=False))

    def forward(self, x):
        return self.projector(x)


class RmclModule(nn.Module):
    """The attributes compute_moco_contrastive reads (vilt_module.py:69-107)."""

    def __init__(self, per_step_bs, K=65536, C=128, n_pgd=3, bf16_queue=True):
        super().__init__()
        from copy import deepcopy
        self.text_embeddings, self.token_type_embeddings = BertEmbeddings(), nn.Embedding(2, 768)
        self.transformer, self.pooler, self.moco_head = VitB32(), Pooler(), MOCOHead(o=C)
        self.k_text_embeddings, self.k_token_type_embeddings = deepcopy(self.text_embeddings), deepcopy(self.token_type_embeddings)
        self.k_transformer, self.k_moco_head = deepcopy(self.transformer), deepcopy(self.moco_head)
        for l in (self.k_text_embeddings, self.k_token_type_embeddings, self.k_transformer, self.k_moco_head):
            for p in l.parameters():
                p.requires_grad = False
        self.momentum, self.temperature, self.num_negative, self.per_step_bs = 0.999, 0.07, K, per_step_bs
        self.text_view, self.image_view, self.augmentation = False, True, False
        self.cosine = nn.CosineSimilarity(dim=1, eps=1e-6)
        self.register_buffer("proj_queue", torch.randn(C, K))
        self.register_buffer("proj_queue_ptr", torch.zeros(1, dtype=torch.long))
        self.max_image_len = 200
        self.pgd_attacker = rmcl_b200.PGDAttack_moco(dict(adv_steps_img=n_pgd, adv_lr_img=0.05, adv_max_norm_img=8 / 255,
                                                          max_image_len=200))
        self.rmcl_bf16_queue = bf16_queue       # main-step InfoNCE on the bf16 shadow (tcgen05); PGD inner loss stays fp32
        self.train_moco_loss = self.val_moco_loss = lambda x: x
        self.logged = {}

    def log(self, name, value, **kw):
        self.logged[name] = value

    def infer(self, batch, mask_text=False, mask_image=False, **kw):
        return rmcl_b200.PGDAttack.infer(self, batch, mask_text, mask_image, **kw)

    def infer_k(self, batch, mask_text=False, mask_image=False):
        view = types.SimpleNamespace(text_embeddings=self.k_text_embeddings, token_type_embeddings=self.k_token_type_embeddings,
                                     transformer=self.k_transformer, pooler=self.pooler, max_image_len=200)
        return rmcl_b200.PGDAttack.infer(view, batch, mask_text, mask_image)


def run(steps=10, warmup=3, B=128, pgd_steps=3, host_batch=False, clock_sampler=None, min_timed_s=0.0):
    """Runs the full step; returns the result dict on rank 0 (None elsewhere).  ``host_batch``: the batch lives in pinned
    host memory and is copied to the device inside every timed step, and the loss is read back to the host (the
    end-to-end arm of bench.py --config cfg4)."""
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True      # the fp32 PGD forwards/backwards (autocast off, as in the reference) on TF32
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)                               # identical initial weights and queue on every rank
    mod = RmclModule(per_step_bs=world * B, n_pgd=pgd_steps).to(dev).train()
    params = [p for p in mod.parameters() if p.requires_grad]
    n_query = sum(p.numel() for p in params)
    n_key = sum(p.numel() for n, p in mod.named_parameters() if n.startswith("k_"))
    opt = torch.optim.AdamW(params, lr=1e-5, fused=True)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    batch = {"image": [torch.randn(B, 3, 384, 384, device=dev, generator=g)], "text": ["x"] * B,
             "text_ids": torch.randint(1000, 30000, (B, 40), device=dev, generator=g),
             "text_labels": torch.full((B, 40), -100, device=dev), "text_masks": torch.ones(B, 40, dtype=torch.long, device=dev)}
    host = {k: (v[0] if isinstance(v, list) else v).cpu().pin_memory() for k, v in batch.items() if k != "text"} if host_batch else None
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values()) if host_batch else 0
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    flat = torch.empty(n_query, device=dev) if world > 1 else None

    # time spent inside the hand-written kernels, by CUDA events around every rmcl_b200.ops entry point
    kernel_events = []
    originals = {}

    def instrument(name):
        fn = originals[name] = getattr(ops, name)

        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            kernel_events.append((name, e0, e1))
            return r
        setattr(ops, name, wrapped)

    for name in ("ema_multi_", "infonce_fwd_bwd", "infonce_loss", "enqueue_", "pgd_step_"):
        instrument(name)
    orig_stats_init = ops.QueueStats.__init__

    def stats_init(self, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_stats_init(self, *a, **k)
        e1.record()
        kernel_events.append(("queue_stats", e0, e1))
    ops.QueueStats.__init__ = stats_init

    def step():
        opt.zero_grad(set_to_none=True)
        if host_batch:
            for k, v in host.items():
                if k == "image":
                    batch["image"][0].copy_(v, non_blocking=True)
                else:
                    batch[k].copy_(v, non_blocking=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ret = rmcl_b200.compute_moco_contrastive(mod, {k: (list(v) if isinstance(v, list) else v) for k, v in batch.items()})
            loss = ret["moco_loss"]
        loss.backward()
        if world > 1:       # DDP's job, done explicitly: one flat all-reduce (no overlap with the backward)
            torch._foreach_mul_([p.grad for p in params], 1.0 / world)
            off = 0
            for p in params:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                off += p.numel()
            dist.all_reduce(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        opt.step()
        if host_batch:
            loss_host.copy_(loss.detach().float(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            last = step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, last

    for _ in range(warmup):
        loss = step()
    barrier()
    kernel_events.clear()
    if clock_sampler is not None:
        clock_sampler.start()
    ms, loss = timed(steps)
    blocks = [ms]
    while sum(blocks) * 1e-3 < min_timed_s and len(blocks) < 50:
        more, loss = timed(steps)
        blocks.append(more)
    clocks = clock_sampler.stop() if clock_sampler is not None else None
    ms = sorted(blocks)[len(blocks) // 2]
    shares = {}
    for name, a, b in kernel_events:
        shares[name] = shares.get(name, 0.0) + a.elapsed_time(b) / (steps * len(blocks))
    for name, fn in originals.items():
        setattr(ops, name, fn)
    ops.QueueStats.__init__ = orig_stats_init
    # replicated state must agree on every rank after the run (queue + pointer: the exchange's parity evidence)
    same = True
    if world > 1:
        ref_q, ref_p = mod.proj_queue.clone(), mod.proj_queue_ptr.clone()
        dist.broadcast(ref_q, 0)
        dist.broadcast(ref_p, 0)
        flag = torch.tensor([int(torch.equal(ref_q, mod.proj_queue) and torch.equal(ref_p, mod.proj_queue_ptr))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    if rank != 0:
        return None
    per = ms / steps
    return {
        "workload": "cfg4: full RMCL step, ViLT-B/32-shaped torch backbone (bf16 autocast; PGD forwards fp32/TF32), "
                    f"B{B}/GPU, C128 K65536, {pgd_steps} PGD steps, key all-gather, flat gradient all-reduce, fused AdamW",
        "n_gpus": world, "global_batch": world * B, "steps": steps, "ms_per_step": per, "blocks_ms": blocks,
        "value": world * steps / (ms * 1e-3), "unit": "rank-steps/s", "samples_per_s": world * B * steps / (ms * 1e-3),
        "scaling": "weak", "query_params": n_query, "key_params": n_key,
        "ms_per_step_in_rmcl_kernels": shares, "rmcl_kernel_share": sum(shares.values()) / per,
        "loss": float(loss.detach()), "queue_ptr": int(mod.proj_queue_ptr.item()), "queue_identical_across_ranks": same,
        "pgd_success_rate": float(mod.logged.get("moco_attack/PGD_success_rate", float("nan"))),
        "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 if host_batch else 0, "clocks": clocks}


def bench_main(args, metric, ClockSampler):
    """bench.py --config cfg4: the bench line (same keys as the cfg2 line) for the full step."""
    local, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    steps, warm = min(args.steps, 20), max(3, min(args.warmup, 5))            # a step is ~0.2 s
    dev_res = run(steps=steps, warmup=warm, clock_sampler=ClockSampler(local))
    e2e_res = run(steps=steps, warmup=2, host_batch=True)
    if dev_res is not None:
        line = {"metric": metric, "value": dev_res["value"], "unit": "steps/s", "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": dev_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": dev_res["workload"], "per_gpu_batch": 128, "global_batch": dev_res["global_batch"],
                           "parallelism": f"dp{world}", "unit_of_value": "128-sample rank-steps per second, summed over ranks",
                           "l2": "inputs larger than L2 (226 MB of images, 450 MB of parameters per step)"},
                "e2e": {"value": e2e_res["value"], "unit": "steps/s", "h2d_bytes_per_step": e2e_res["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": e2e_res["d2h_bytes_per_step"], "ms_per_step": e2e_res["ms_per_step"],
                        "api": "rmcl_b200.compute_moco_contrastive (drop-in for objectives.py:217) with the batch in pinned host memory"},
                "kernels_ms_per_step": dev_res["ms_per_step_in_rmcl_kernels"], "rmcl_kernel_share": dev_res["rmcl_kernel_share"],
                "parity_check": {"ok": dev_res["queue_identical_across_ranks"], "queue_identical_across_ranks": dev_res["queue_identical_across_ranks"],
                                 "queue_ptr": dev_res["queue_ptr"]},
                "samples_per_s": dev_res["samples_per_s"], "gpu_launches": None, "clocks": dev_res["clocks"], "loss": dev_res["loss"],
                "roofline": None}
        print(json.dumps(line), flush=True)
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--pgd-steps", type=int, default=3)
    args = ap.parse_args()
    res = run(args.steps, args.warmup, args.batch, args.pgd_steps)
    if res is not None:
        print(json.dumps(res), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
