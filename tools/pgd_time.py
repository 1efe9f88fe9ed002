"""cfg3 PGD update kernel alone (B128; pixel 3x384x384 and embedding 185x768 perturbations), every mode, CUDA-event
timed over consecutive launches; RMCL_B200_LIB selects an experiment build for same-box A/B runs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
out = []
for tag, shape, mode, lr, eps in (("pixel ref_linf", (128, 3, 384, 384), "ref_linf", 0.05, 8 / 255),
                                  ("pixel l2", (128, 3, 384, 384), "l2", 0.5, 1.0),
                                  ("embed ref_linf", (128, 185, 768), "ref_linf", 0.05, 8 / 255),
                                  ("embed l2", (128, 185, 768), "l2", 0.5, 1.0),
                                  ("pixel sign", (128, 3, 384, 384), "sign_linf", 2 / 255, 8 / 255),
                                  ("embed sign", (128, 185, 768), "sign_linf", 2 / 255, 8 / 255)):
    grad = torch.randn(shape, device=dev, generator=g)
    delta = torch.zeros(shape, device=dev)
    for _ in range(3):
        ops.pgd_step_(delta, grad, lr, eps, mode)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.pgd_step_(delta, grad, lr, eps, mode)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 5 * 1000)
    t = min(ts)
    out.append(f"{tag} {t:.1f} us = {12.0 * grad.numel() / t / 1e3 / 6537:.3f}")
    del grad, delta
print(os.environ.get("RMCL_B200_LIB", "product"), "|", " | ".join(out), flush=True)
