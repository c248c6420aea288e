#!/usr/bin/env python3
"""Build the tracked per-round summaries under profiles/ from the scratch captures in gpurun_out/.
usage: python tools/make_profiles.py r01"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list of the bench command (cfg3) ----
src = os.path.join(G, f"{rnd}_launches_cfg3.csv")
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [r[ki].split("(")[0].replace("void ", "") for r in data]
t = [float(r[vi].replace(",", "")) for r in data]
open(os.path.join(P, f"{rnd}_launches_cfg3.csv"), "w").write(open(src).read())
is_first = lambda n: n.startswith("ntt_pass_kernel<1, 1")          # first kernel of every LDE (DIF pass with inverse roots)
starts = [i for i, n in enumerate(names) if is_first(n) and (i == 0 or not is_first(names[i - 1]))]
s, e = starts[1], starts[2]          # warm-up step, TIMED step, then the per-phase re-runs
agg = collections.OrderedDict()
for n, x in zip(names[s:e], t[s:e]):
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += x
tot = sum(a[1] for a in agg.values())
plain = json.loads(open(os.path.join(G, f"{rnd}_plain_cfg3.log")).read().strip().splitlines()[-1])
out = [f"# {rnd} - launch list of `python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras` (1xB200)", "",
       f"Source: `profiles/{rnd}_launches_cfg3.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are",
       f"serialised and cold-cache: compare shares, not absolutes). Window = the timed step (launches {s}..{e - 1} of {len(data)}).", "",
       "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, (n, x) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {n} | {x / 1e6:.2f} | {x / tot * 100:.1f}% |")
out += ["", f"Step total under ncu: {tot / 1e6:.1f} ms. The same command without the profiler (CUDA events on the launching stream):", "",
        "```", json.dumps({k: plain[k] for k in ("value", "phases_s", "gpu_launches", "clocks")}), "```",
        "", "Shares agree: merkle_leaf_kernel %.1f%% of the ncu window vs %.1f%% of the plain step (merkle_leaf phase / value)." %
        (agg["merkle_leaf_kernel"][1] / tot * 100, plain["phases_s"].get("merkle_leaf", plain["phases_s"]["merkle"]) / plain["value"] * 100)]
open(os.path.join(P, f"{rnd}_launches_cfg3.md"), "w").write("\n".join(out) + "\n")

# ---- ncu --set full tables ----
traffic = {}
def table(rep, title, out_name, key=None):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_raw.csv"
    open(tmp, "w").write(raw)
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_table.py"), tmp, "--md"], capture_output=True, text=True).stdout
    open(os.path.join(P, out_name), "w").write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` ({rep}, kept in gpurun_out/ scratch); "
                                               "columns: ms = gpu__time_duration, rd/wr = dram__bytes_read/write.sum, dram% = gpu__dram_throughput pct of peak, "
                                               "issue% = smsp__issue_active, alu% / fmaH% = alu pipe instructions / fmaheavy pipe cycles (pct of peak), st_* = "
                                               "warp stall cycles per issued instruction.\n\n" + md)
    if key:
        r = list(csv.reader(raw.splitlines()))
        h, u = r[0], r[1]
        ix = {x: i for i, x in enumerate(h)}
        def gb(row, col):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[ix[col]]]
            return float(row[ix[col]].replace(",", "")) * mult
        lde = 0.0
        for row in r[2:]:
            nm = row[ix["Kernel Name"]]
            b = gb(row, "dram__bytes_read.sum") + gb(row, "dram__bytes_write.sum")
            if "merkle_leaf" in nm:
                traffic.setdefault(key, {})["merkle_leaf_kernel"] = b
            if "ntt_" in nm:
                lde += b
        if lde:
            traffic.setdefault(key, {})["lde"] = lde

table(f"{rnd}_prof_cfg3.ncu-rep", f"{rnd} - ncu --set full, cfg3 (2^23 x 256, blowup 2): the five LDE kernels and the leaf hash", f"{rnd}_ncu_cfg3.md", "cfg3")
table(f"{rnd}_prof_ntt_slab.ncu-rep", f"{rnd} - ncu --set full, one 32-column slab of cfg3 (pass structure 8/8/7)", f"{rnd}_ncu_ntt_slab.md")
table(f"{rnd}_prof_leaf_cfg2.ncu-rep", f"{rnd} - ncu --set full, merkle_leaf_kernel at cfg2 (2^21 rows x 64 cols)", f"{rnd}_ncu_leaf_cfg2.md", "cfg2")
if traffic:
    tp = os.path.join(P, "traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old.update(traffic)
    json.dump(old, open(tp, "w"), indent=1)
print(open(os.path.join(P, f"{rnd}_launches_cfg3.md")).read())
print(json.dumps(traffic))
