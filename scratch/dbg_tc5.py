import sys, os, math, torch
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
    return S, cps, offs, tile
def run(B, C, K):
    g = torch.Generator().manual_seed(B + C + K)
    q = torch.randn(B, C, generator=g); k = torch.nn.functional.normalize(torch.randn(B, C, generator=g), dim=1)
    queue = torch.randn(C, K, generator=g).bfloat16()
    ops._ws_cache.clear()
    b = ops.infonce_fwd_bwd(q.to(dev), k.to(dev), queue.to(dev), 0.07, path="tcgen05")
    torch.cuda.synchronize()
    S, cps, offs, tile = plan(B, C, K)
    (ws,) = ops._ws_cache.values()
    raw = ws[(-ws.data_ptr()) % 256:]
    m = raw[offs["m"]:offs["m"] + S*B*4].view(torch.float32).view(S, B).double()
    l = raw[offs["l"]:offs["l"] + S*B*4].view(torch.float32).view(S, B).double()
    o = raw[offs["o"]:offs["o"] + S*B*C*4].view(torch.float32).view(S, B, C).double()
    qh = raw[offs["qb"]:offs["qb"] + B*C*2].view(torch.bfloat16).view(-1, C)[:B].double()
    qd = queue.to(dev).double()
    scale2 = math.log2(math.e) / 0.07
    print(f"B={B} C={C} K={K} S={S} tiles/split={cps//tile}")
    nbad = 0
    for s in range(S):
        cols = qd[:, s*cps:min((s+1)*cps, K)]
        sl = (qh @ cols) * scale2                       # [B, n]
        p = torch.exp2(sl - m[s][:, None])
        o_ref = p @ cols.T
        l_ref = p.sum(1)
        err = (o[s] - o_ref).abs() / (o_ref.abs().max(1, keepdim=True).values + 1e-30)
        bad = (err > 0.05) | ~torch.isfinite(o[s])
        lerr = ((l[s] - l_ref).abs() / l_ref).max().item()
        if bad.any() or lerr > 1e-3:
            nbad += 1
            rows_bad = bad.any(1).nonzero().flatten()
            # growth history of the first bad row
            r0 = rows_bad[0].item() if rows_bad.numel() else 0
            tmax = sl[r0].view(-1, tile).max(1).values
            print(f" split {s}: bad rows {rows_bad.numel()} [{rows_bad[:4].tolist()}..{rows_bad[-1].item() if rows_bad.numel() else ''}] bad elems {bad.sum().item()} lerr={lerr:.2e}")
            print("   row", r0, "tile maxima:", [round(x, 1) for x in tmax.tolist()][:30], " m_used:", round(m[s][r0].item(), 1))
            bc = bad[r0].nonzero().flatten()
            print("   bad cols:", bc[:16].tolist(), " ratio o/o_ref at bad:", (o[s][r0][bc[:6]] / o_ref[r0][bc[:6]]).tolist())
            if nbad >= 4: break
    print(" total bad splits so far:", nbad)
run(1024, 64, 131072)
run(256, 256, 65536)
