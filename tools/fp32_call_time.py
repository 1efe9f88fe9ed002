"""Whole fp32-accurate InfoNCE call (fp32 queue -> bf16 hi/lo planes -> split-operand two-pass tcgen05 kernels) by CUDA-graph
replay over queue copies larger than L2, at cfg2 and the cfg4 shape; RMCL_B200_LIB selects an experiment build."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
out = []
for B, C, K in ((256, 256, 65536), (128, 128, 65536)):
    torch.manual_seed(0)
    q = torch.randn(B, C, device="cuda"); k = torch.randn(B, C, device="cuda")
    n_copies = max(2, int(math.ceil(160e6 / (C * K * 4))) + 1)
    queues = [torch.randn(C, K, device="cuda") for _ in range(n_copies)]
    n = 6 * n_copies
    def batch():
        for j in range(n):
            ops.infonce_fwd_bwd(q, k, queues[j % n_copies], 0.07, normalize_k=True, want=("loss", "dq", "k_hat"))
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
    out.append(f"B{B} C{C}: {min(ts):.2f} us")
    del queues
print(os.environ.get("RMCL_B200_LIB", "product"), "| fp32-accurate InfoNCE call:", " | ".join(out), flush=True)
