"""BASELINE cfg5 per-GPU InfoNCE (B512 C768 K262144 bf16): two-pass tcgen05 path vs the SIMT path — agreement
at full size and CUDA-event time of the whole call (prep + S pass + PV pass + finalize)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev = "cuda"
shapes = [(512, 768, 262144), (512, 512, 262144), (256, 768, 65536)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
g = torch.Generator().manual_seed(0)
for (B, C, K) in shapes:
    q = torch.randn(B, C, generator=g).bfloat16().to(dev)
    k = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1).bfloat16().to(dev)
    queue = torch.nn.functional.normalize(torch.randn(C, K, generator=g), dim=0).bfloat16().to(dev)
    out = {}
    for path in ("tcgen05", "simt"):
        n = 10 if path == "tcgen05" else 2
        for _ in range(2):
            r = ops.infonce_fwd_bwd(q, k, queue, 0.07, path=path)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            r = ops.infonce_fwd_bwd(q, k, queue, 0.07, path=path)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1000
        out[path] = r
        print(f"B={B} C={C} K={K} {path}: {us:.1f} us/call = {4.0*B*C*K/us/1e6:.0f} TF/s", flush=True)
    a, b = out["simt"], out["tcgen05"]
    rel = lambda x, y: ((x.double() - y.double()).norm() / y.double().norm()).item()
    print(f"   tc vs simt: lse {rel(b['lse'], a['lse']):.2e} loss {rel(b['loss'], a['loss']):.2e} dq {rel(b['dq'], a['dq']):.2e} "
          f"argmax equal {(a['argmax'] == b['argmax']).float().mean().item():.4f}", flush=True)
