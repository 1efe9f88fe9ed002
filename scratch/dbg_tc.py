import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def run(B, C, K, norm):
    g = torch.Generator().manual_seed(B + C + K)
    q = torch.randn(B, C, generator=g); k = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1)
    queue = torch.randn(C, K, generator=g)
    if norm: queue = torch.nn.functional.normalize(queue, dim=0)
    queue = queue.bfloat16()
    a = ops.infonce_fwd_bwd(q.to(dev), k.to(dev), queue.to(dev), 0.07, path="simt")
    b = ops.infonce_fwd_bwd(q.to(dev), k.to(dev), queue.to(dev), 0.07, path="tcgen05")
    torch.cuda.synchronize()
    dq = b["dq"]
    nan_rows = torch.isnan(dq).any(1).nonzero().flatten()
    err = ((dq - a["dq"]).abs().max() / a["dq"].abs().max()).item()
    print(f"B={B} C={C} K={K} norm={norm}: nan_rows={nan_rows.numel()} first={nan_rows[:8].tolist()} relerr_vs_simt={err:.3e} "
          f"lse_err={((b['lse']-a['lse']).abs().max()/a['lse'].abs().max()).item():.2e}", flush=True)
for K in (8192, 16384, 32768, 49152, 65536):
    run(256, 256, K, False)
run(2048, 256, 4096, False)
run(1024, 128, 8192, False)
run(1024, 64, 8192, False)
run(256, 256, 65536, True)
for _ in range(3):
    run(256, 256, 65536, False)
