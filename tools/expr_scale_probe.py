"""How do the compiled and the interpreted expression kernels scale with the program length?  The golden quotient program of the sm_all
AIR (169 records, 40 live temporaries) repeated k times, over 2^bits rows, both paths, device-resident buffers.
usage: python tools/expr_scale_probe.py [bits] [k ...]   (k = 24 needs > 15 min of NVRTC: keep k <= 12)"""
import json, os, pathlib, sys, time, types
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from test_oracle_expressions import make_domain_ctx, prover_side_program
import pil2_stark_js_b200 as m
from pil2_stark_js_b200 import prover_helpers as H
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 18
sizes = [int(a) for a in sys.argv[2:]] or [1, 4, 16]
gpu = m.default_context(0)
qcode = json.loads((ROOT / "tests" / "golden" / "sm_all_q_code.json").read_text())
rng = np.random.default_rng(3)
d = make_domain_ctx(qcode, rng, bits - 1, bits)
info = d["pilInfo"]
E = 1 << bits
d["xDivXSubXi_ext"] = rng.integers(0, H.P, size=3 * 2 * E, dtype=np.uint64)
d["f_ext"] = np.zeros(3 * E, dtype=np.uint64)
d["evals"] = [[int(x) for x in rng.integers(0, H.P, size=3, dtype=np.uint64)] for _ in range(5)]
ctx = types.SimpleNamespace(**d); ctx.gpu = gpu
names = [k for k in d if isinstance(d[k], np.ndarray) and (k.endswith("_ext") or k == "Zi_ext")]
for n in sizes:
    base, code = prover_side_program(qcode), []
    for rep in range(n):
        for c in base:                                # fresh temporary ids per copy
            c2 = json.loads(json.dumps(c))
            for r in [c2["dest"]] + c2["src"]:
                if r["type"] == "tmp":
                    r["id"] += 100000 * rep
            code.append(c2)
    cc = H.compile_code(ctx, code, "ext")
    held = {}
    for name, rw in cc.buffers:
        held[name] = gpu.upload(np.ascontiguousarray(H._host_buffer(ctx, name), dtype=np.uint64).reshape(-1))
    ctx.dev_buffers = {k: (v, v.words) for k, v in held.items()}
    res = {}
    for mode in ("jit", "interp"):
        os.environ["PIL2GPU_EXPR"] = mode
        t0 = time.perf_counter(); H.calculateExps(ctx, code, "ext"); gpu.sync(); first = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(3): H.calculateExps(ctx, code, "ext")
        gpu.sync(); res[mode] = ((time.perf_counter() - t0) / 3, first)
    print(f"{len(code)} records ({cc.ops.size // 16} live), {cc.n_slots} slots, 2^{bits} rows: compiled {res['jit'][0] * 1e3:.2f} ms (first call incl. NVRTC {res['jit'][1]:.1f} s), "
          f"interpreted {res['interp'][0] * 1e3:.2f} ms", flush=True)
    for b in held.values(): b.free()
