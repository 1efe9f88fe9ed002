import sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops, _lib
dev="cuda"
g = torch.Generator().manual_seed(0)
B,C,K=256,256,65536
q = torch.randn(B, C, generator=g).bfloat16().to(dev); k = torch.randn(B, C, generator=g).bfloat16().to(dev)
queue = torch.randn(C, K, generator=g).bfloat16().to(dev)
L = _lib.lib()
for it in range(2):
    r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss","dq","k_hat"))
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
L.rmcl_debug_read(buf, 512)
d = list(buf)
print("tile: softmax[wait_start, s_full_seen, decided, p_arrive] | mma[iter_start, after_G1(i+2)_issue, p_full_seen]")
for i in range(14):
    e = d[16+i*8:16+i*8+8]
    print(i, e[0:4], "|", e[4:7])
