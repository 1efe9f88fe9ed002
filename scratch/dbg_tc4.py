import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev = "cuda"
def run(B, C, K):
    g = torch.Generator().manual_seed(B + C + K)
    q = torch.randn(B, C, generator=g); k = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1)
    queue = torch.randn(C, K, generator=g).bfloat16()
    a = ops.infonce_fwd_bwd(q.to(dev), k.to(dev), queue.to(dev), 0.07, path="simt")
    out = []
    for _ in range(3):
        b = ops.infonce_fwd_bwd(q.to(dev), k.to(dev), queue.to(dev), 0.07, path="tcgen05")
        torch.cuda.synchronize()
        dq = b["dq"]
        out.append((torch.isnan(dq).any(1).sum().item(), round(((dq - a["dq"]).abs().max() / a["dq"].abs().max()).item(), 5)))
    print(f"dbg={os.environ.get('RMCL_TC_DEBUG')} B={B} C={C} K={K}: {out}", flush=True)
for cfg in [(256, 256, 65536), (128, 256, 262144), (2048, 256, 16384), (1024, 128, 65536), (1024, 64, 131072), (256, 256, 53248)]:
    run(*cfg)
