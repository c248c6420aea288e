#!/usr/bin/env python3
"""Compact per-launch table from `ncu -i X.ncu-rep --page raw --csv` (one line per profiled launch).
usage: ncu -i rep --page raw --csv > raw.csv; python tools/ncu_table.py raw.csv [--md]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel", lambda s: s.split("(")[0].replace("void ", "")[:34]),
        ("Grid Size", "grid", str),
        ("gpu__time_duration.sum", "ms", None),
        ("dram__bytes_read.sum", "rd", None), ("dram__bytes_write.sum", "wr", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", None),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", None),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%", None),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaH%", None),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", None),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", None),
        ("launch__registers_per_thread", "regs", None),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar", None),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_lsb", None),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_ssb", None),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", None),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio", None),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", None),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankc", None)]
units = rows[1]
def fmt(name, v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if name in ("rd", "wr"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        return "%.3fG" % (x * mult / 1e9)
    if name == "ms":
        mult = {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
        return "%.3f" % (x * mult)
    return "%.1f" % x if abs(x) < 1e6 else "%.3g" % x
md = "--md" in sys.argv
names = [c[1] for c in cols if c[0] in idx]
print(("| " + " | ".join(names) + " |") if md else "  ".join("%-8s" % n for n in names))
if md:
    print("|" + "---|" * len(names))
for r in rows[2:]:
    out = []
    for key, name, f in cols:
        if key not in idx:
            continue
        v = r[idx[key]]
        out.append(f(v) if f else fmt(name, v, units[idx[key]]))
    print(("| " + " | ".join(out) + " |") if md else "  ".join("%-8s" % o for o in out))
