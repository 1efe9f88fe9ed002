"""SURVEY 8(d) "GPU baseline to beat": the reference's own expressions run eagerly by torch on the same B200, next to the
kernels that replace them, at the BASELINE shapes.  The expressions are restated here from the reference lines they come
from (this tool does not import oracle/): they are what `vilt/modules/objectives.py` and `attack/pgd_attack_vilt.py` execute
on a GPU today.

    EMA       objectives.py:219-224 (x4 at 257-260): per-tensor k.data = k.data*m + q.data*(1-m), 161 tensors
    InfoNCE   objectives.py:326-334+351: normalize, queue.clone(), two einsums, cat, /T, CrossEntropyLoss, backward
    enqueue   objectives.py:244-248: int(ptr) sync, strided copy, pointer write
    PGD       pgd_attack_vilt.py:162-173: inf-norm, clamp, step, add, clamp
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200  # noqa: E402
from rmcl_b200 import ops  # noqa: E402

dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000      # us


out = {}
g = torch.Generator(device=dev).manual_seed(0)

# ---- EMA over the real ViLT-B/32 key-encoder shape list
shapes = []
for line in open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vilt_b32_key_encoder_shapes.txt")):
    line = line.strip()
    if line and not line.startswith("#"):
        shapes.append(tuple(int(d) for d in line.split("x")))
pk = [torch.randn(s, device=dev, generator=g) for s in shapes]
pq = [torch.randn(s, device=dev, generator=g) for s in shapes]
m = 0.999


def ema_eager():
    for k, q in zip(pk, pq):
        k.data = k.data * m + q.data * (1.0 - m)


plan = ops.EmaPlan(pk, pq)
out["ema_161_tensors"] = {"eager_us": timed(ema_eager, 10), "kernel_us": timed(lambda: ops.ema_multi_(plan, m), 20)}
del pk, pq, plan

# ---- InfoNCE fwd + bwd
for tag, B, C, K, qdt in (("infonce_cfg2_bf16", 256, 256, 65536, torch.bfloat16), ("infonce_cfg2_fp32", 256, 256, 65536, torch.float32),
                          ("infonce_cfg4_bf16", 128, 128, 65536, torch.bfloat16), ("infonce_cfg5_bf16", 512, 768, 262144, torch.bfloat16)):
    q_raw = torch.randn(B, C, device=dev, generator=g)
    k = F.normalize(torch.randn(B, C, device=dev, generator=g), dim=1)
    queue = F.normalize(torch.randn(C, K, device=dev, generator=g), dim=0).to(qdt)

    def eager():
        q = q_raw.detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(qdt == torch.bfloat16)):
            qn = F.normalize(q, dim=1)
            qu = queue.clone().detach()
            l_pos = torch.einsum("nc,nc->n", [qn, k]).unsqueeze(-1)
            l_neg = torch.einsum("nc,ck->nk", [qn, qu])
            logits = torch.cat([l_pos, l_neg], dim=1) / 0.07
        labels = torch.zeros(logits.shape[0], dtype=torch.long, device=dev)
        loss = F.cross_entropy(logits.float(), labels)
        loss.backward()
        return loss

    def fused():
        return ops.infonce_fwd_bwd(q_raw, k, queue, 0.07, want=("loss", "dq", "argmax"))

    n = 5 if K > 100000 else 20
    out[tag] = {"eager_us": timed(eager, n), "kernel_us": timed(fused, n)}
    del q_raw, k, queue

# ---- enqueue
B, C, K = 256, 256, 65536
queue = torch.randn(C, K, device=dev, generator=g)
keys = F.normalize(torch.randn(B, C, device=dev, generator=g), dim=1)
ptr = torch.zeros(1, dtype=torch.long, device=dev)


def enq_eager():
    p = int(ptr)
    queue[:, p:p + B] = keys.T
    ptr[0] = (p + B) % K


ptr2 = torch.zeros(1, dtype=torch.long, device=dev)
out["enqueue_cfg2"] = {"eager_us": timed(enq_eager, 50), "kernel_us": timed(lambda: ops.enqueue_(queue, keys, ptr2), 50)}

# ---- PGD update, B128 pixels
grad = torch.randn(128, 3, 384, 384, device=dev, generator=g)
delta = torch.zeros_like(grad)
lr, eps = 0.05, 8 / 255


def pgd_eager():
    global delta
    gcl = grad.clone().detach().float()
    denorm = torch.norm(gcl.view(gcl.size(0), -1), dim=1, p=float("inf")).view(-1, 1, 1, 1)
    denorm = torch.clamp(denorm, min=1e-8)
    step = (lr * gcl / denorm).to(delta)
    delta = (delta + step).detach()
    delta = torch.clamp(delta, -eps, eps).detach()


d2 = torch.zeros_like(grad)
out["pgd_pixels_b128"] = {"eager_us": timed(pgd_eager, 10), "kernel_us": timed(lambda: ops.pgd_step_(d2, grad, lr, eps, "ref_linf"), 20)}

for v in out.values():
    v["speedup"] = v["eager_us"] / v["kernel_us"]
print(json.dumps({k: {a: round(b, 1) for a, b in v.items()} for k, v in out.items()}), flush=True)
