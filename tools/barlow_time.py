"""Barlow-Twins cross-correlation loss fwd+bwd at the reference's size (batch 128, projector 8192,
vilt_module.py:115, config.py:181): the fused tcgen05 chain (prep + barlow_tc_kernel + finalize) against the
reference's own expression chain (objectives.py:480-486 + autograd) run eagerly by torch on the same GPU."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev = "cuda"
lam = 0.0051
shapes = [(128, 8192), (256, 8192), (1024, 8192), (128, 2048)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]


def eager(q, k, bs):
    q = q.detach().requires_grad_(True)
    c = q.T @ k
    c.div_(bs)
    on_diag = torch.diagonal(c).add_(-1).pow_(2).sum()
    n = c.shape[0]
    off_diag = c.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten().pow_(2).sum()
    loss = on_diag + lam * off_diag
    loss.backward()
    return loss.detach(), q.grad


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000


g = torch.Generator().manual_seed(0)
for (B, D) in shapes:
    k = torch.randn(B, D, generator=g).to(dev)
    q = (0.7 * k + 0.7 * torch.randn(B, D, generator=g).to(dev))
    for _ in range(3):
        r = ops.barlow_fwd_bwd(q, k, 1.0 / B, lam, path="gram")
        le, ge = eager(q, k, B)
    us = timed(lambda: ops.barlow_fwd_bwd(q, k, 1.0 / B, lam, path="gram"), 20)
    us_d = float("nan")
    if B <= 256:
        ops.barlow_fwd_bwd(q, k, 1.0 / B, lam, path="direct")
        us_d = timed(lambda: ops.barlow_fwd_bwd(q, k, 1.0 / B, lam, path="direct"), 20)
    # the same call replayed as a CUDA graph: without the host-side launch overhead of three/four short kernels
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(10):
                ops.barlow_fwd_bwd(q, k, 1.0 / B, lam, path="gram")
    gr.replay()
    us_g = timed(gr.replay, 5) / 10
    us_e = timed(lambda: eager(q, k, B), 5)
    torch.backends.cuda.matmul.allow_tf32 = True
    us_e_tf32 = timed(lambda: eager(q, k, B), 5)
    torch.backends.cuda.matmul.allow_tf32 = False
    rel = lambda x, y: ((x.double() - y.double()).norm() / y.double().norm()).item()
    print(f"B={B} D={D}: gram {us:.1f} us (graph replay {us_g:.1f} us) | direct {us_d:.1f} us = {4.0*B*D*D/us_d/1e6:.0f} TF/s | eager torch fp32 "
          f"{us_e:.1f} us, tf32 {us_e_tf32:.1f} us | speed-up of gram {us_e/us:.1f}x | loss rel {rel(r['loss'], le):.2e} "
          f"dq rel {rel(r['dq'], ge):.2e}", flush=True)
