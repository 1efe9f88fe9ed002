import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import ops
dev = "cuda"
def plan(B, C, K, sms=148):
    rows, tile = 128, (64 if C == 256 else 128)
    rb = (B + rows - 1) // rows; b_pad = rb * rows
    tiles = (K + tile - 1) // tile
    splits = max(1, min(sms // rb, tiles))
    tps = (tiles + splits - 1) // splits
    cps = tps * tile
    S = (K + cps - 1) // cps
    off = 0; offs = {}
    def take(name, n):
        nonlocal off
        offs[name] = off; off = (off + n + 255) // 256 * 256
    take("qhat", B*C*4); take("khat", B*C*4); take("inv", B*4); take("pos2", B*4); take("qb", b_pad*C*2)
    take("m", S*B*4); take("l", S*B*4); take("av", S*B*4); take("ai", S*B*4); take("o", S*B*C*4)
    return S, cps, offs
def run(B, C, K, norm=False):
    g = torch.Generator().manual_seed(B + C + K)
    q = torch.randn(B, C, generator=g); k = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1)
    queue = torch.randn(C, K, generator=g).bfloat16()
    ops._ws_cache.clear()
    b = ops.infonce_fwd_bwd(q.to(dev), k.to(dev), queue.to(dev), 0.07, path="tcgen05")
    torch.cuda.synchronize()
    S, cps, offs = plan(B, C, K)
    (ws,) = ops._ws_cache.values()
    a0 = (-ws.data_ptr()) % 256
    raw = ws[a0:]
    m = raw[offs["m"]:offs["m"] + S*B*4].view(torch.float32).view(S, B)
    l = raw[offs["l"]:offs["l"] + S*B*4].view(torch.float32).view(S, B)
    o = raw[offs["o"]:offs["o"] + S*B*C*4].view(torch.float32).view(S, B, C)
    bad = ~torch.isfinite(o)
    bs = bad.any(2)
    print(f"B={B} C={C} K={K} S={S} cps={cps}: dq nan rows={torch.isnan(b['dq']).any(1).sum().item()} bad (split,row) pairs={bs.sum().item()}")
    if bs.any():
        idx = bs.nonzero()
        print("  bad splits:", sorted(set(idx[:, 0].tolist()))[:20], " rows sample:", idx[:10, 1].tolist())
        s0, r0 = idx[0].tolist()
        print("  m,l at first bad:", m[s0, r0].item(), l[s0, r0].item(), " bad cols:", bad[s0, r0].nonzero().flatten()[:10].tolist(), bad[s0, r0].sum().item())
        print("  values:", o[s0, r0, :8].tolist())
    print("  m finite:", torch.isfinite(m).all().item(), " l finite:", torch.isfinite(l).all().item(), " o absmax finite part:", o[~bad].abs().max().item())
for K in (57344, 61440, 65536, 65536 + 896, 131072):
    run(256, 256, K)
run(128, 256, 65536)
