"""In-kernel timeline of the tcgen05 InfoNCE kernel (CTA (0,0)) at cfg2, inside a step-like sequence
(EMA-sized L2 flush before each call).  Prints cycles since kernel entry; layout: csrc/infonce_tc.cu."""
import sys

import torch

sys.path.insert(0, ".")
import rmcl_b200  # noqa: E402
from rmcl_b200 import ops  # noqa: E402

B, C, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (256, 256, 65536)))
torch.manual_seed(0)
q = torch.randn(B, C, device="cuda").bfloat16()
k = torch.nn.functional.normalize(torch.randn(B, C, device="cuda"), dim=1).bfloat16()
queue = torch.randn(C, K, device="cuda").bfloat16()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def call():
    ops.infonce_fwd_bwd(q, k, queue, 0.07, path="tcgen05", want=("loss", "dq", "argmax"))


for _ in range(3):
    flush.zero_()
    call()
rows = []
for rep in range(5):
    flush.zero_()
    t = ops.tc_timeline(call).tolist()
    rows.append(t)
# the same call with the stage events of rmcl_profile_enable (what bench.py reports per kernel)
ops.profile_enable(True)
HEAD0 = 8 + 8 * 40
for rep in range(5):
    flush.zero_()
    tt = ops.tc_timeline(call).tolist()
    st = [tt[HEAD0 + 2 * c] for c in range(148)]
    en = [tt[HEAD0 + 2 * c + 1] for c in range(148)]
    print("profiled rep", rep, ops.profile_infonce_ms(), "kernel span by %globaltimer:", max(en) - min(st), "ns;",
          " ".join(f"{n}={tt[i] - tt[0]}" for i, n in enumerate(["entry", "setup", "prep+barrier", "q_in_tmem", "o_all_done", "stats", "stored", "barrier2"])))
ops.profile_enable(False)
t = rows[-1]
t0 = t[0]
names = ["entry", "setup", "prep+barrier", "q_in_tmem", "o_all_done", "stats", "stored", "barrier2"]
for rep, tt in enumerate(rows):
    print("rep", rep, " ".join(f"{n}={tt[i] - tt[0]}" for i, n in enumerate(names)))
for rep, tt in enumerate(rows):
    if tt[312]:
        print("rep", rep, "prep rows written at", tt[320] - tt[0], "| finalize row of team 0:",
              " ".join(f"{n}={tt[312 + i] - tt[0]}" for i, n in enumerate(["start", "stats", "synced", "partials_summed", "synced2", "dq_written", "done"])))
HEAD = 8 + 8 * 40
n_cta = 148 if (B, C, K) == (256, 256, 65536) else 0
for rep, tt in enumerate(rows):
    if n_cta:
        st = [tt[HEAD + 2 * c] for c in range(n_cta)]
        en = [tt[HEAD + 2 * c + 1] for c in range(n_cta)]
        t00 = min(st)
        dur = sorted(e - s for s, e in zip(st, en))
        print(f"rep {rep} CTA start spread {max(st) - t00} ns, first start -> last end {max(en) - t00} ns, "
              f"CTA duration min/median/max {dur[0]}/{dur[len(dur) // 2]}/{dur[-1]} ns")
print("tile  S_issue S_landed P_stored O_issue Pbuf_free S_free_seen S_in_regs decided")
for i in range(40):
    v = t[8 + 8 * i: 16 + 8 * i]
    if not any(v):
        break
    print(f"{i:4d} " + " ".join(f"{(x - t0) if x else 0:8d}" for x in v))
