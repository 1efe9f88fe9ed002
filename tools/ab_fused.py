"""Same-process A/B is not possible (the toggle is read once), so this runs one mode per invocation:
RMCL_B200_INFONCE_FUSED=0|1 python tools/ab_fused.py — whole InfoNCE call by CUDA-graph replay over queue copies larger
than L2 and by a plain Python loop (torch-extension binding), at the cfg2 and cfg4 shapes, with and without gradient."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
torch.manual_seed(0)
res = []
for B, C, K, ncopy in ((256, 256, 65536, 6), (128, 128, 65536, 12)):
    q = torch.randn(B, C, device="cuda").bfloat16(); k = torch.randn(B, C, device="cuda").bfloat16()
    queues = [torch.nn.functional.normalize(torch.randn(C, K, device="cuda"), dim=0).bfloat16() for _ in range(ncopy)]
    for need_grad in (True, False):
        want = ("loss", "dq", "k_hat") if need_grad else ("argmax", "k_hat")
        def batch():
            for j in range(8 * ncopy):
                ops.infonce_fwd_bwd(q, k, queues[j % ncopy], 0.07, normalize_k=True, need_grad=need_grad, want=want)
        batch(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            batch()
        g.replay(); torch.cuda.synchronize()
        ts, tl = [], []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / (8 * ncopy) * 1000)
            e0.record(); batch(); e1.record(); torch.cuda.synchronize()
            tl.append(e0.elapsed_time(e1) / (8 * ncopy) * 1000)
        res.append(f"B{B} C{C} grad={int(need_grad)}: graph {min(ts):.2f} us, python loop {min(tl):.2f} us")
        del g
print("fused=" + os.environ.get("RMCL_B200_INFONCE_FUSED", "1"), ops.infonce_launch_names(256, 256, 65536, torch.bfloat16), "|", " | ".join(res), flush=True)
