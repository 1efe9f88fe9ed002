"""tcgen05 InfoNCE at cfg2: call with gradient vs the statistics-only call (need_grad=False), graph replay over queue copies > L2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
B, C, K = 256, 256, 65536
torch.manual_seed(0)
q = torch.randn(B, C, device="cuda").bfloat16(); k = torch.randn(B, C, device="cuda").bfloat16()
queues = [torch.randn(C, K, device="cuda").bfloat16() for _ in range(6)]
res = {}
for name, kw in (("grad", dict(want=("loss", "dq", "k_hat"))), ("nograd", dict(need_grad=False, want=("argmax", "k_hat", "lse")))):
    def batch():
        for j in range(48):
            ops.infonce_fwd_bwd(q, k, queues[j % 6], 0.07, normalize_k=True, path="tcgen05", **kw)
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
    res[name] = min(ts)
a = ops.infonce_fwd_bwd(q, k, queues[0], 0.07, normalize_k=True, path="tcgen05")
b = ops.infonce_fwd_bwd(q, k, queues[0], 0.07, normalize_k=True, path="tcgen05", need_grad=False, want=("argmax", "lse", "loss"))
print(os.environ.get("RMCL_B200_LIB", "product"), "whole call: with grad %.2f us, statistics only %.2f us; lse equal %s argmax equal %s" % (
    res["grad"], res["nograd"], torch.equal(a["lse"], b["lse"]), torch.equal(a["argmax"], b["argmax"])))
