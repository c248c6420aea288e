"""Stand-alone driver of the FRI-polynomial kernels for ncu captures:  python tools/fripol_probe.py [ext_bits] [cols] [n_open]"""
import sys, time, pathlib
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import pil2_stark_js_b200 as m

ext_bits, cols, n_open = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 22), (2, 256), (3, 2)))
ctx = m.default_context(0)
rng = np.random.default_rng(1)
buf = rng.integers(0, 0xFFFFFFFF00000001, size=cols << ext_bits, dtype=np.uint64)
xi, vf1, vf2 = (rng.integers(0, 0xFFFFFFFF00000001, size=3, dtype=np.uint64) for _ in range(3))
openings = list(range(n_open))
dbuf = ctx.upload(buf)
xdiv = ctx.x_div_x_sub_xi(xi, openings, ext_bits - 1, ext_bits, download=False)
terms = [(dbuf, cols, c, 1, o) for o in openings for c in range(cols)]
evals = rng.integers(0, 0xFFFFFFFF00000001, size=(len(terms), 3), dtype=np.uint64)
for rep in range(3):
    t0 = time.perf_counter()
    f = ctx.fri_pol(terms, evals, openings, xdiv, vf1, vf2, ext_bits, download=False)
    ctx.sync()
    dt = time.perf_counter() - t0
    f.free()
    print(f"fri_pol {len(terms)} terms over 2^{ext_bits} x {cols}: {dt * 1e3:.2f} ms  ({8 * (cols << ext_bits) / dt / 1e9:.0f} GB/s)", flush=True)
