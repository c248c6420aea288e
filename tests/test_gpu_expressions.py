"""f4 on the GPU: the expression interpreter (pil2gpu_calculate_exps_dev through pil2_stark_js_b200.prover_helpers, the mirror of
src/prover/prover_helpers.js:23-110) against the oracle's calculate_exps, bit-exact -- on the quotient program the reference generated
for its golden sm_all proof (tests/golden/sm_all_q_code.json) and on random programs that exercise every operand kind, both
domains, negative row offsets, mixed dimensions and stores into committed polynomials."""
import json
import pathlib
import types

import numpy as np
import pytest

from oracle import expressions as X
from oracle import gl_spec as S
from test_oracle_expressions import prover_side_program, make_domain_ctx

pytestmark = pytest.mark.gpu
P = S.P
ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def gpu():
    import pil2_stark_js_b200 as m
    return m.default_context(0)


@pytest.fixture(params=["jit", "interp"], autouse=True)
def expr_mode(request, monkeypatch):
    """Every test runs through the run-time compiled kernel (NVRTC; "jit" makes a missing compiler an error, so the GPU box proves the
    compiled path is the one that ran) and through the interpreter."""
    if request.param == "jit":                              # no libnvrtc on this machine: the compiled path cannot be exercised (same image everywhere, but be explicit)
        from pil2_stark_js_b200 import _lib
        probe = np.zeros(16, dtype=np.uint32)
        probe[0], probe[1], probe[4], probe[7] = 3, 1, 0 | (1 << 8), 3 | (1 << 8)          # t0 = x
        if _lib.load().pil2gpu_expr_jit_check(probe.ctypes.data, 1, None, 0, 4, 1, None, 0) == -5:
            pytest.skip("NVRTC not available")
    monkeypatch.setenv("PIL2GPU_EXPR", request.param)
    return request.param


@pytest.fixture(scope="module")
def qcode():
    return json.loads((ROOT / "tests" / "golden" / "sm_all_q_code.json").read_text())


def ns(d, gpu):
    c = types.SimpleNamespace(**{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()})
    c.gpu = gpu
    return c


@pytest.mark.parametrize("n_bits,ext_bits", [(4, 5), (6, 8), (10, 11)])
def test_golden_quotient_program_over_the_extended_domain(gpu, qcode, n_bits, ext_bits):
    from pil2_stark_js_b200 import prover_helpers as H
    rng = np.random.default_rng(n_bits)
    d = make_domain_ctx(qcode, rng, n_bits, ext_bits)
    code = prover_side_program(qcode)
    want = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    X.calculate_exps(want, code, "ext")
    ctx = ns(d, gpu)
    cc = H.callCalculateExps(4, {"code": code}, "ext", ctx)
    assert np.array_equal(ctx.q_ext, want["q_ext"])
    assert cc.n_slots <= H.MAX_SLOTS and cc.written == {"q_ext"}
    assert int(ctx.q_ext.max()) < P


def test_golden_program_with_device_resident_buffers(gpu, qcode):
    """ctx.dev_buffers: the stage buffers stay in HBM (as after extendAndMerkelize with device_resident = True); only q_ext moves."""
    from pil2_stark_js_b200 import prover_helpers as H
    rng = np.random.default_rng(9)
    d = make_domain_ctx(qcode, rng, 8, 9)
    code = prover_side_program(qcode)
    want = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    X.calculate_exps(want, code, "ext")
    ctx = ns(d, gpu)
    held = {name: gpu.upload(getattr(ctx, name)) for name in ("cm1_ext", "cm2_ext", "cm3_ext", "const_ext")}
    ctx.dev_buffers = {k: (v, v.words) for k, v in held.items()}
    for name in held:                      # poison the host copies: they must not be read
        getattr(ctx, name)[:] = 0
    H.calculateExps(ctx, code, "ext")
    assert np.array_equal(ctx.q_ext, want["q_ext"])
    for b in held.values():
        b.free()


def random_program(rng, info, dom, n_records=60):
    """Records over every operand kind the reference's getRef knows (prover_helpers.js:152-219)."""
    code, tmps = [], []
    cm1 = [i for i, p in enumerate(info["cmPolsMap"]) if p["dim"] == 1]
    cm3 = [i for i, p in enumerate(info["cmPolsMap"]) if p["dim"] == 3 and p["stage"] < 4]
    primes = [0, 0, 1, -1, 2]

    def src():
        k = int(rng.integers(0, 9))
        if k == 0 and tmps:
            t = tmps[int(rng.integers(0, len(tmps)))]
            return {"type": "tmp", "id": t[0], "dim": t[1]}
        if k == 1:
            return {"type": "const", "id": int(rng.integers(0, info["nConstants"])), "prime": primes[int(rng.integers(0, 5))]}
        if k == 2:
            return {"type": "cm", "id": cm1[int(rng.integers(0, len(cm1)))], "prime": primes[int(rng.integers(0, 5))]}
        if k == 3:
            return {"type": "cm", "id": cm3[int(rng.integers(0, len(cm3)))], "prime": primes[int(rng.integers(0, 5))]}
        if k == 4:
            return {"type": "number", "value": str(int(rng.integers(0, P, dtype=np.uint64)))}
        if k == 5:
            return {"type": "public", "id": int(rng.integers(0, 3))}
        if k == 6:
            st = int(rng.integers(2, 5))
            return {"type": "challenge", "stage": st, "stageId": int(rng.integers(0, {2: 2, 3: 3, 4: 1}[st]))}
        if k == 7:
            return {"type": "x"}
        if dom == "ext":
            return {"type": ["xDivXSubXi", "Zi"][int(rng.integers(0, 2))], "id": int(rng.integers(0, 2)), "boundaryId": 0}
        return {"type": "eval", "id": int(rng.integers(0, 5))}

    def dim_of(r):
        if r["type"] == "tmp":
            return r["dim"]
        if r["type"] == "cm":
            return info["cmPolsMap"][r["id"]]["dim"]
        return 3 if r["type"] in ("challenge", "xDivXSubXi", "eval") else 1

    for k in range(n_records):
        op = ["add", "sub", "mul", "mul", "copy", "muladd"][int(rng.integers(0, 6))]
        srcs = [src() for _ in range({"copy": 1, "muladd": 3}.get(op, 2))]
        dim = max(dim_of(s) for s in srcs)
        dest = {"type": "tmp", "id": 100 + k, "dim": dim}
        code.append({"op": op, "dest": dest, "src": srcs})
        tmps.append((100 + k, dim))
    # stores: an F3 result into the last stage-3 polynomial (never read above: cm3 excludes nothing, so use the Q stage) and q / f
    last3 = next(t for t in reversed(tmps) if t[1] == 3)
    qpol = next(i for i, p in enumerate(info["cmPolsMap"]) if p["stage"] == 4)
    code.append({"op": "copy", "dest": {"type": "cm", "id": qpol, "prime": 0}, "src": [{"type": "tmp", "id": last3[0], "dim": 3}]})
    if dom == "ext":
        code.append({"op": "add", "dest": {"type": "q", "id": 0, "dim": 3}, "src": [{"type": "tmp", "id": last3[0], "dim": 3}, {"type": "number", "value": "5"}]})
        code.append({"op": "mul", "dest": {"type": "f", "id": 0, "dim": 3}, "src": [{"type": "tmp", "id": tmps[0][0], "dim": tmps[0][1]}, {"type": "x"}]})
    return code


@pytest.mark.parametrize("dom", ["n", "ext"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_programs_vs_oracle(gpu, qcode, dom, seed):
    from pil2_stark_js_b200 import prover_helpers as H
    rng = np.random.default_rng(100 + seed)
    n_bits, ext_bits = 5, 7
    d = make_domain_ctx(qcode, rng, n_bits, ext_bits)
    info = d["pilInfo"]
    N, E = 1 << n_bits, 1 << ext_bits
    d["const_n"] = rng.integers(0, P, size=info["nConstants"] * N, dtype=np.uint64)
    for st, w in info["mapSectionsN"].items():
        d[st + "_n"] = rng.integers(0, P, size=w * N, dtype=np.uint64)
    d["x_n"] = np.array([pow(S.root_of_unity(n_bits), i, P) for i in range(N)], dtype=np.uint64)
    d["xDivXSubXi_ext"] = rng.integers(0, P, size=3 * 2 * E, dtype=np.uint64)
    d["f_ext"] = np.zeros(3 * E, dtype=np.uint64)
    d["evals"] = [[int(x) for x in rng.integers(0, P, size=3, dtype=np.uint64)] for _ in range(5)]
    code = random_program(rng, info, dom)
    want = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    X.calculate_exps(want, code, dom)
    ctx = ns(d, gpu)
    cc = H.calculateExps(ctx, code, dom)
    for name in cc.written:
        assert np.array_equal(getattr(ctx, name), want[name]), name
    assert "cm4_" + dom in cc.written


def test_compiled_kernels_persist_in_the_cache_directory(qcode, tmp_path):
    """PIL2GPU_JIT_CACHE: the first process compiles the golden program and leaves a cubin in the directory, a second process loads it
    (no NVRTC call: much faster first evaluation) and computes the same q_ext."""
    import os, subprocess, sys, time
    code = (
        "import json, sys, time, types, pathlib; import numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_oracle_expressions import prover_side_program, make_domain_ctx\n"
        "import pil2_stark_js_b200 as m\nfrom pil2_stark_js_b200 import prover_helpers as H\n"
        "qcode = json.loads(pathlib.Path(%r).read_text())\n"
        "d = make_domain_ctx(qcode, np.random.default_rng(4), 6, 7)\n"
        "ctx = types.SimpleNamespace(**d); ctx.gpu = m.default_context(0)\n"
        "ctx.gpu.sync(); t0 = time.perf_counter(); H.calculateExps(ctx, prover_side_program(qcode), 'ext'); dt = time.perf_counter() - t0\n"
        "np.save(sys.argv[1], ctx.q_ext); print(dt)\n"
    ) % (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden" / "sm_all_q_code.json"))
    env = dict(os.environ, PIL2GPU_EXPR="jit", PIL2GPU_JIT_CACHE=str(tmp_path))
    out = []
    for k in range(2):
        r = subprocess.run([sys.executable, "-c", code, str(tmp_path / ("q%d.npy" % k))], env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0 and "NVRTC" in r.stderr and "not found" in r.stderr:
            pytest.skip("NVRTC not available")
        assert r.returncode == 0, r.stderr[-800:]
        out.append(float(r.stdout.strip().splitlines()[-1]))
        assert len(list(tmp_path.glob("pil2gpu_expr_*_sm100a.cubin"))) == 1
    assert np.array_equal(np.load(tmp_path / "q0.npy"), np.load(tmp_path / "q1.npy"))
    assert out[1] < 0.5 * out[0], out                                # the second process did not compile


def test_host_buffer_entry_point(gpu, qcode):
    """pil2gpu_calculate_exps (what the N-API addon's calculateExps and js/prover_helpers.js call): host arrays in, the written buffers
    updated in place, same values as the oracle; buffers not flagged `written` come back untouched."""
    import ctypes
    from pil2_stark_js_b200 import prover_helpers as H, _lib
    rng = np.random.default_rng(21)
    d = make_domain_ctx(qcode, rng, 6, 7)
    code = prover_side_program(qcode)
    want = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    X.calculate_exps(want, code, "ext")
    ctx = ns(d, gpu)
    cc = H.compile_code(ctx, code, "ext")
    host = [np.ascontiguousarray(H._host_buffer(ctx, name), dtype=np.uint64).reshape(-1).copy() for name, _ in cc.buffers]
    before = [h.copy() for h in host]
    arr = (_lib.ExprHostBuffer * len(host))()
    for i, ((name, rw), h) in enumerate(zip(cc.buffers, host)):
        arr[i] = _lib.ExprHostBuffer(h.ctypes.data, rw, 1, 1 if name in cc.written else 0)
    rc = gpu._L.pil2gpu_calculate_exps(gpu.handle, cc.ops.ctypes.data, cc.ops.size // 16, cc.consts.ctypes.data, cc.consts.size // 3, arr, len(host), 7, 1)
    assert rc == 0, gpu._L.pil2gpu_last_error().decode()
    for (name, _), h, b in zip(cc.buffers, host, before):
        if name in cc.written:
            assert np.array_equal(h, want[name].reshape(-1)), name
        else:
            assert np.array_equal(h, b), name


def test_program_errors(gpu, qcode):
    from pil2_stark_js_b200 import prover_helpers as H, Pil2GpuError
    rng = np.random.default_rng(5)
    ctx = ns(make_domain_ctx(qcode, rng, 3, 4), gpu)
    with pytest.raises(ValueError):
        H.calculateExps(ctx, [{"op": "div", "dest": {"type": "tmp", "id": 0, "dim": 1}, "src": [{"type": "x"}, {"type": "x"}]}], "ext")
    with pytest.raises(ValueError):
        H.calculateExps(ctx, [{"op": "copy", "dest": {"type": "q", "id": 0, "dim": 3}, "src": [{"type": "x"}]}], "n")      # "Accessing q in domain n"
    with pytest.raises(ValueError):
        H.calculateExps(ctx, [{"op": "copy", "dest": {"type": "tmp", "id": 1, "dim": 1}, "src": [{"type": "tmp", "id": 0, "dim": 1}]}], "ext")
    wide = [{"op": "copy", "dest": {"type": "tmp", "id": k, "dim": 1}, "src": [{"type": "x"}]} for k in range(70)]
    wide += [{"op": "add", "dest": {"type": "tmp", "id": 100 + k, "dim": 1},
              "src": [{"type": "tmp", "id": k, "dim": 1}, {"type": "tmp", "id": 99 + k, "dim": 1} if k else {"type": "x"}]} for k in range(70)]
    wide.append({"op": "copy", "dest": {"type": "q", "id": 0, "dim": 3}, "src": [{"type": "tmp", "id": 169, "dim": 1}]})
    with pytest.raises(ValueError):
        H.calculateExps(ctx, wide, "ext")                              # 70 temporaries alive at once
    # records whose temporary nobody reads are dropped before slots are assigned: the same 140 records without the final store are an empty program
    assert H.compile_code(ctx, wide[:-1], "ext").ops.size == 0
    # the C ABI validates the records itself
    bad = np.zeros(16, dtype=np.uint32)
    bad[0], bad[1] = 9, 2
    assert gpu._L.pil2gpu_calculate_exps_dev(gpu.handle, bad.ctypes.data, 1, None, 0, None, 0, 4, 1) == -1
    assert "opcode" in gpu._L.pil2gpu_last_error().decode()
