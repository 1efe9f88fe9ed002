import sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops, _lib
dev="cuda"
g = torch.Generator().manual_seed(0)
B,C,K=256,256,65536
q = torch.randn(B, C, generator=g).bfloat16().to(dev); k = torch.randn(B, C, generator=g).bfloat16().to(dev)
queue = torch.randn(C, K, generator=g).bfloat16().to(dev)
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
L = _lib.lib()
for it in range(3):
    flush.zero_(); torch.cuda.synchronize()
    r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=("loss","dq","k_hat"))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 256)()
    L.rmcl_debug_read(buf, 256)
    d = list(buf)
    print(f"iter {it}: q_loaded@{d[0]} loop_end@{d[1]} epilogue_end@{d[2]} all_end@{d[3]}")
    print("  sm  (wait_s, softmax, end@):", [(d[8+4*i], d[9+4*i], d[10+4*i]) for i in range(14)])
    print("  sm detail (ldtm, max+pairsync, exp, odone_wait):", [(d[160+4*i], d[161+4*i], d[162+4*i], d[163+4*i]) for i in range(14)])
    print("  mma (wait_p, @, wait_k, @):", [(d[100+4*i], d[101+4*i], d[102+4*i], d[103+4*i]) for i in range(14)])
