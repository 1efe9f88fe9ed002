"""profiles/traffic.json from an `ncu --set full` raw CSV page of tools/prof_r2.py: DRAM bytes (read + write) per launch of
every kernel bench.py reports, stamped with the hash of the sources the capture was taken from (bench.py flags a stale
table), plus a human-readable counter summary.

    ncu -i gpurun_out/r2_prof.ncu-rep --page raw --csv > profiles/r2_ncu_full.raw.csv
    python tools/make_traffic.py profiles/r2_ncu_full.raw.csv"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

WANT = {"dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "gpu__time_duration.sum": "us",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "launch__registers_per_thread": "regs"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        rec = {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for name, key in WANT.items():
            if name in col and r[col[name]] not in ("", "n/a"):
                rec[key] = float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1)
        out.append(rec)
    # the LAST launch of each kernel variant is the warm one
    last = {}
    for rec in out:
        last[rec["kernel"] + rec["grid"]] = rec
    def pick(sub, nth=-1, grid=None):
        c = [r for r in out if sub in r["kernel"] and (grid is None or grid in r["grid"])]
        return c[nth] if c else None
    table = {}
    def put(key, rec):
        if rec is not None and "rd" in rec:
            table[key] = int(rec["rd"] + rec.get("wr", 0))
    put("ema", pick("ema_multi_kernel"))
    fused = pick("infonce_tc_kernel")
    put("infonce_call", fused)
    put("infonce_partial", fused)
    put("enqueue", pick("enqueue_kernel"))
    pgd = [r for r in out if "pgd_ticket_kernel" in r["kernel"]]
    if len(pgd) >= 3:
        put("pgd_pixel_ref_linf", pgd[-3]); put("pgd_pixel_l2", pgd[-2]); put("pgd_pixel_sign_linf", pgd[-1])
    s_pass, pv = pick("infonce_s_kernel"), pick("infonce_pv_kernel")
    if s_pass and pv:
        table["infonce_cfg4_b128_c128_fp32_call"] = int(s_pass["rd"] + s_pass["wr"] + pv["rd"] + pv["wr"])
    table["_meta"] = {"source_hash": bench.source_hash(), "capture": os.path.relpath(path, ROOT),
                      "what": "dram__bytes_read.sum + dram__bytes_write.sum of the last launch of each kernel in tools/prof_r2.py (ncu --set full)"}
    json.dump(table, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(table, indent=1))
    print("\n| kernel | grid | us | dram MB (rd+wr) | dram % | tensor pipe % | warps active % | regs |\n|---|---|---|---|---|---|---|---|")
    for rec in last.values():
        short = rec["kernel"].split("(")[0].replace("void ", "").replace("rmcl::", "").replace("(anonymous namespace)::", "")
        print(f"| `{short[:70]}` | {rec['grid']} | {rec.get('us', 0):.1f} | {rec.get('rd', 0) / 1e6:.2f} + {rec.get('wr', 0) / 1e6:.2f} | "
              f"{rec.get('dram_pct', 0):.1f} | {rec.get('tensor_pct', 0):.1f} | {rec.get('warps_pct', 0):.1f} | {int(rec.get('regs', 0))} |")


if __name__ == "__main__":
    main(sys.argv[1])
