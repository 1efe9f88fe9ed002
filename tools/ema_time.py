import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev="cuda"
shapes = bench.load_shapes()
g = torch.Generator(device=dev).manual_seed(0)
pk = [torch.randn(s, device=dev, generator=g) for s in shapes]
pq = [torch.randn(s, device=dev, generator=g) for s in shapes]
n = sum(p.numel() for p in pk)
for chunk in (8192, 16384, 32768, 65536):
    plan = ops.EmaPlan(pk, pq, chunk_elems=chunk)
    for _ in range(3): ops.ema_multi_(plan, 0.999)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.ema_multi_(plan, 0.999)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1)/20*1000
    print(f"chunk={chunk}: {us:.1f} us  {12*n/us/1e3:.0f} GB/s ({12*n/us/1e3/6537:.3f})", flush=True)
