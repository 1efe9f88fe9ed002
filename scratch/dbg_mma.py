import sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops, _lib
dev="cuda"
g = torch.Generator().manual_seed(0)
L = _lib.lib()
for (B,C,K) in [(256,256,65536),(256,128,65536),(256,64,65536)]:
    q = torch.randn(B, C, generator=g).bfloat16().to(dev); k = torch.randn(B, C, generator=g).bfloat16().to(dev)
    queue = torch.randn(C, K, generator=g).bfloat16().to(dev)
    for it in range(2):
        r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss","dq","k_hat"))
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 8)()
        L.rmcl_debug_read(buf, 8)
        print(f"C={C}: G1(0)={buf[0]} G1(1)={buf[1]} G2(0)={buf[2]} G2(1)={buf[3]} cycles (issue->mbarrier observed)")
