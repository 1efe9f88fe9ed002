"""The tcgen05 flash pass alone (RMCL_INFONCE_DEBUG_PARTIAL_ONLY) over consecutive launches, CUDA-graph replay over queue
copies larger than L2; RMCL_B200_LIB selects an experiment build (same-box A/B)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
out = []
for B, C, K in ((256, 256, 65536), (128, 128, 65536)):
    torch.manual_seed(0)
    q = torch.randn(B, C, device="cuda").bfloat16(); k = torch.randn(B, C, device="cuda").bfloat16()
    n_copies = max(2, int(math.ceil(160e6 / (C * K * 2))) + 1)
    queues = [torch.randn(C, K, device="cuda").bfloat16() for _ in range(n_copies)]
    ops.infonce_fwd_bwd(q, k, queues[0], 0.07, normalize_k=True, path="tcgen05", want=("loss", "dq", "k_hat"))
    n = 8 * n_copies
    def batch():
        for j in range(n):
            ops.infonce_fwd_bwd(q, k, queues[j % n_copies], 0.07, normalize_k=True, path="tcgen05", want=(), _partial_only=True)
    batch(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        batch()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1000)
    t = min(ts)
    out.append(f"B{B} C{C}: {t:.2f} us = {4.0 * B * C * (K + 1) / t / 1e6:.0f} TF/s")
    del queues
print(os.environ.get("RMCL_B200_LIB", "product"), "| flash pass alone:", " | ".join(out), flush=True)
