"""GPU time of one whole fused-InfoNCE call (prep + tcgen05 partial + finalize) by CUDA-graph replay,
L2-warm (10 calls per graph) and after an L2 flush (single call): free of the host launch overhead
that dominates a Python loop of 40 us calls."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev="cuda"
g = torch.Generator().manual_seed(0)
flush = torch.empty(512*1024*1024, dtype=torch.uint8, device=dev)
for (B,C,K,norm) in [(256,256,65536,False),(256,256,65536,True),(128,128,65536,True),(512,128,65536,True),(256,64,65536,True)]:
    q = torch.randn(B, C, generator=g).bfloat16().to(dev); k = torch.randn(B, C, generator=g).bfloat16().to(dev)
    queue = torch.randn(C, K, generator=g)
    if norm: queue = torch.nn.functional.normalize(queue, dim=0)
    queue = queue.bfloat16().to(dev)
    want=("loss","dq","k_hat")
    for _ in range(3): ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=want)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(10):
                r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=want)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gr.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1)/50*1000
    # cold: flush L2 between single replays of a 1-op graph
    with torch.cuda.stream(st):
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, stream=st):
            r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path="tcgen05", want=want)
    torch.cuda.synchronize()
    ts=[]
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        e0.record(); g1.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)*1000)
    print(f"B={B} C={C} K={K} norm={norm}: graph x10 L2-warm {us:.1f} us/op ({4.0*B*C*K/us/1e6:.0f} TF/s); single-op graph after L2 flush {min(ts):.1f} us ({4.0*B*C*K/min(ts)/1e6:.0f} TF/s, queue {C*K*2/min(ts)/1e3:.0f} GB/s)")
