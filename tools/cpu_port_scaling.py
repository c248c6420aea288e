#!/usr/bin/env python
"""Validates the scaling law bench.py uses to extrapolate the CPU port (oracle/gl_oracle.c) from a row sample to the full
row count: LDE ~ rows * log2(rows), hashing / FRI ~ rows.  Runs the port at 2^k rows for several k up to the FULL row count
of cfg3 (2^23) at a reduced column count (the whole 256-column buffer needs 80 GiB of host memory; columns are independent
in the LDE and the hash is linear in ceil(C/8) + 1 permutations per row), and compares each measured time with the value
extrapolated from the previous size.  Output: JSON on stdout (kept under profiles/).
    python tools/cpu_port_scaling.py [--cols 32] [--bits 17 19 21 23]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--cols", type=int, default=32)
ap.add_argument("--bits", type=int, nargs="+", default=[17, 19, 21, 23])
a = ap.parse_args()
threads = os.cpu_count() or 1
out = {"cols": a.cols, "blowup": 2, "threads": threads, "runs": []}
prev = None
for b in a.bits:
    d = bench.cpu_commit_sample(b, a.cols, 1, b, threads, 0x5EED0003)     # sample_bits == n_bits: no extrapolation inside
    run = {"rows_log2": b, "lde_s": d["lde_s"], "merkle_s": d["merkle_s"], "fri_s": d["fri_s"], "total_s": d["sample_s"]}
    if prev is not None:
        pb, pd = prev
        ratio = float(1 << (b - pb))
        log_ratio = (b + b + 1) / float(pb + pb + 1)
        pred = pd["lde_s"] * ratio * log_ratio + pd["merkle_s"] * ratio + pd["fri_s"] * ratio
        run["predicted_from_2^%d_s" % pb] = pred
        run["measured_over_predicted"] = d["sample_s"] / pred
    out["runs"].append(run)
    prev = (b, d)
    print(json.dumps(run), file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
