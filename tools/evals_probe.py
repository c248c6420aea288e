"""Stand-alone driver of the evaluation kernels for ncu captures:  python tools/evals_probe.py [n_bits] [cols] [n_open]"""
import sys, time, pathlib
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import pil2_stark_js_b200 as m

n_bits, cols, n_open = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 21), (2, 256), (3, 2)))
ctx = m.default_context(0)
rng = np.random.default_rng(1)
buf = rng.integers(0, 0xFFFFFFFF00000001, size=cols << n_bits, dtype=np.uint64)
xi = rng.integers(0, 0xFFFFFFFF00000001, size=3, dtype=np.uint64)
levs = ctx.compute_levs(xi, list(range(n_open)), n_bits)
dbuf = ctx.upload(buf)
ev = [(c, 1, o) for o in range(n_open) for c in range(cols)]
for rep in range(3):
    t0 = time.perf_counter()
    out = ctx.compute_evals(dbuf, cols, n_bits, n_bits, ev, levs, n_open)
    dt = time.perf_counter() - t0
    print(f"evals {len(ev)} over 2^{n_bits} x {cols}: {dt * 1e3:.2f} ms  ({8 * (cols << n_bits) / dt / 1e9:.0f} GB/s)", flush=True)
