"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM,
TMA tile loads = UTMALDG, bulk async copies = UBLKCP / UBLKPF, tcgen05.commit = UTCBAR) in the built library.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "robust-multimodal-contrastive-learning_b200", "librmcl_b200.so")
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTCBAR", "UTCCP", "SYNCS", "ELECT", "MUFU.EX2", "HMMA", "FFMA", "REDG", "ATOMG"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|rmcl::|void ", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    m = re.match(r"[\w:]+(<[^()]*>)?", name)
    return m.group(0) if m else name


counts, order, cur, i = [], [], None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        order.append(short(names[i] if i < len(names) else m.group(1)))
        counts.append(collections.Counter())
        cur = counts[-1]
        i += 1
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for k in MN:
            if op.startswith(k):
                cur[k] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)}: instruction counts per kernel (sm_100a); columns with no hits anywhere are omitted")
used = [k for k in MN if any(c[k] for c in counts)]
print("kernel".ljust(78) + "".join(k.rjust(10) for k in ["total"] + used))
tot = collections.Counter()
for name, c in zip(order, counts):
    if not any(c[k] for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UBLKPF")) and "--all" not in sys.argv:
        continue
    print(name[:77].ljust(78) + str(c["_total"]).rjust(10) + "".join(str(c[k]).rjust(10) for k in used))
    tot.update(c)
print("TOTAL (listed kernels)".ljust(78) + str(tot["_total"]).rjust(10) + "".join(str(tot[k]).rjust(10) for k in used))
print(f"\n{len(order)} kernels in the library; those without tensor-core / TMA / bulk-copy instructions (EMA, PGD, enqueue, prep, finalize, SIMT InfoNCE, "
      "queue statistics) are listed with --all")
