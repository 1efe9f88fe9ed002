import sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops, _lib
dev="cuda"
g = torch.Generator().manual_seed(0)
L = _lib.lib()
B,C,K=256,256,65536
q = torch.randn(B, C, generator=g).bfloat16().to(dev); k = torch.randn(B, C, generator=g).bfloat16().to(dev)
queue = torch.randn(C, K, generator=g).bfloat16().to(dev)
for it in range(3):
    r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss","dq","k_hat"))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 8)()
    L.rmcl_debug_read(buf, 8)
    print("finalize cycles: pdl_wait_done", buf[0], "stats_done", buf[1], "after_barrier", buf[2], "po_done", buf[3], "after_barrier2", buf[4], "grad_done", buf[5], "end", buf[6])
# whole-op timing
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5): ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss","dq","k_hat"))
torch.cuda.synchronize()
e0.record()
for _ in range(50): ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss","dq","k_hat"))
e1.record(); torch.cuda.synchronize()
print("InfoNCE op (L2-warm queue, back-to-back):", e0.elapsed_time(e1)/50*1000, "us")
