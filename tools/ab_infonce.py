"""Same-box A/B of library builds (RMCL_B200_LIB): whole fused-InfoNCE call at cfg2 by CUDA-graph replay,
cycling over queue copies larger than L2, plus the in-step event-bracketed stage times."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
B, C, K = 256, 256, 65536
torch.manual_seed(0)
q = torch.randn(B, C, device="cuda").bfloat16(); k = torch.randn(B, C, device="cuda").bfloat16()
queues = [torch.randn(C, K, device="cuda").bfloat16() for _ in range(6)]
want = ("loss", "dq", "k_hat")
def batch():
    for j in range(48):
        ops.infonce_fwd_bwd(q, k, queues[j % 6], 0.07, normalize_k=True, path="tcgen05", want=want)
batch(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    batch()
g.replay(); torch.cuda.synchronize()
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 48 * 1000)
print(os.environ.get("RMCL_B200_LIB", "product"), "whole InfoNCE call, graph replay: min %.2f us  median %.2f us" % (min(ts), sorted(ts)[3]))
