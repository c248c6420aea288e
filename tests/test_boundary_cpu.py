"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/pil2gpu.h declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import pathlib
import re
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "pil2gpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pil2gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pil2_stark_js_b200 import _lib
    L = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/pil2gpu.h but not exported"
        assert s in _lib._SIGS, f"{s} has no ctypes signature in _lib.py"
    assert set(_lib._SIGS) == set(syms)


def test_layout_helpers_match_reference_formula():
    # pure functions of the ABI (no device needed): _getNNodes (merklehash_p.js:28-42)
    from pil2_stark_js_b200 import _lib
    from oracle import gl_spec as S
    L = _lib.load()
    for h in [1, 2, 3, 4, 5, 7, 8, 33, 256, 1000, 1 << 18, (1 << 24)]:
        assert L.pil2gpu_merkle_nnodes(h) == S.merkle_n_nodes(4 * h)
    for k in range(1, 25):
        assert L.pil2gpu_merkle_nnodes(1 << k) == 8 * (1 << k) - 4
        assert L.pil2gpu_merkle_depth(1 << k) == k
    assert L.pil2gpu_merkle_depth(33) == 6 and L.pil2gpu_merkle_depth(1) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on CPU-only hosts")
    import pil2_stark_js_b200 as m
    with pytest.raises(m.Pil2GpuError) as e:
        m.Context(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    import numpy as np
    with pytest.raises(m.Pil2GpuError):
        m.fft(np.zeros(8, dtype=np.uint64), 1, 3, np.zeros(8, dtype=np.uint64))


def test_product_does_not_import_oracle():
    pkg = ROOT / "pil2_stark_js_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_addon_type_checks():
    """include/pil2gpu.h must compile as C (the boundary is a C ABI), and the N-API addon must type-check against it
    (Node.js is absent from this image: tests/stubs/node_api.h declares the N-API functions the addon uses)."""
    import subprocess
    hdr = ROOT / "include" / "pil2gpu.h"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", str(hdr)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", str(ROOT / "tests" / "stubs"),
                        str(ROOT / "napi" / "pil2gpu_addon.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_js_shims_bind_only_exported_addon_functions():
    """Every addon.<name> used by js/*.js is registered by the addon's Init table."""
    addon_src = (ROOT / "napi" / "pil2gpu_addon.cc").read_text()
    registered = set(re.findall(r'\{"([A-Za-z0-9]+)", nullptr,', addon_src))
    used = set()
    for f in (ROOT / "js").glob("*.js"):
        used |= set(re.findall(r"\baddon\.([A-Za-z0-9]+)\(", f.read_text()))
    assert used and used <= registered, f"js uses unregistered addon functions: {sorted(used - registered)}"


def _build_addon_driver(tmp_path):
    import subprocess
    exe = tmp_path / "napi_mock"
    lib_dir = ROOT / "pil2_stark_js_b200"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "tests" / "stubs"), "-o", str(exe),
                        str(ROOT / "tests" / "stubs" / "napi_mock.cc"), str(ROOT / "napi" / "pil2gpu_addon.cc"), "-L", str(lib_dir), "-l:libpil2gpu.so",
                        f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_addon_argument_checks_execute_under_a_mock_napi(tmp_path):
    """The N-API addon never runs under Node here, so its checks are EXECUTED against a small in-process model of the N-API calls it
    makes (tests/stubs/napi_mock.cc): wrong-sized buffers and page lists are RangeErrors, wrong types and null contexts TypeErrors,
    all raised before anything reaches libpil2gpu (the context handed in is a fake pointer that must never be dereferenced)."""
    import subprocess
    from pil2_stark_js_b200 import _lib
    _lib.load()
    exe = _build_addon_driver(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL CHECKS PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
    assert r.stdout.count("ok  :") >= 25


def test_expression_jit_source_compiles_offline():
    """The run-time compiled form of the constraint-expression evaluator (csrc/expr_jit.cuh): the CUDA source generated for the golden
    quotient program of the sm_all AIR (169 records) is accepted by NVRTC for sm_100a -- no device needed -- and mirrors the records:
    one statement per record, temporaries as registers, the F3 x F3 products as gl3_mul."""
    import ctypes, json, sys, types
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    from test_oracle_expressions import prover_side_program, make_domain_ctx
    from pil2_stark_js_b200 import prover_helpers as H, _lib
    L = _lib.load()
    qcode = json.loads((ROOT / "tests" / "golden" / "sm_all_q_code.json").read_text())
    d = make_domain_ctx(qcode, np.random.default_rng(1), 4, 5)
    cc = H.compile_code(types.SimpleNamespace(**d), prover_side_program(qcode), "ext")
    rw = np.array([r for _, r in cc.buffers], dtype=np.uint64)
    buf = ctypes.create_string_buffer(1 << 20)
    rc = L.pil2gpu_expr_jit_check(cc.ops.ctypes.data, len(cc.ops) // 16, rw.ctypes.data, len(rw), 5, 1, buf, 1 << 20)
    src = buf.value.decode()
    body = src[src.index('extern "C"'):]
    recs = cc.ops.reshape(-1, 16)
    n_f3_products = sum(1 for r in recs if r[0] in (2, 4) and ((r[7] >> 8) & 255) == 3 and ((r[10] >> 8) & 255) == 3)
    assert len(recs) == 146 and cc.n_slots == 17              # 23 of the program's 169 records compute temporaries nobody reads: dropped
    assert body.count("\n    { const gl3 a = ") == len(recs) and body.count("f3_mul(a, b)") == n_f3_products and "xval" in body
    L.pil2gpu_last_error.restype = ctypes.c_char_p
    if rc == -5:                                                        # PIL2GPU_E_UNSUPPORTED: no libnvrtc on this machine
        import pytest
        pytest.skip(L.pil2gpu_last_error().decode())
    assert rc == 0, L.pil2gpu_last_error().decode()
    bad = np.zeros(16, dtype=np.uint32)
    bad[0] = 9
    assert L.pil2gpu_expr_jit_check(bad.ctypes.data, 1, None, 0, 4, 1, None, 0) == -1


def test_js_shims_only_call_exported_addon_functions():
    """Node is absent, so the JS shims never run here: at least every `addon.<name>` they call must be an export of the N-API addon, and
    their brackets must balance (a crude stand-in for `node --check`)."""
    import re
    exported = set(re.findall(r'\{"(\w+)", nullptr, \w+, nullptr', (ROOT / "napi" / "pil2gpu_addon.cc").read_text()))
    assert "calculateExps" in exported and len(exported) >= 30
    for f in sorted((ROOT / "js").glob("*.js")):
        src = f.read_text()
        for name in set(re.findall(r"\baddon\.(\w+)\s*\(", src)):
            assert name in exported, f"{f.name} calls addon.{name}, which the addon does not export"
        s = re.sub(r"//[^\n]*", "", src)
        s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
        for q in ('"', "'", "`"):
            s = re.sub(q + r"(?:\\.|[^" + q + r"\\])*" + q, q + q, s)
        stack, pairs = [], {")": "(", "]": "[", "}": "{"}
        for ch in s:
            if ch in "([{":
                stack.append(ch)
            elif ch in ")]}":
                assert stack and stack.pop() == pairs[ch], f"{f.name}: unbalanced {ch}"
        assert not stack, f"{f.name}: unclosed {stack[-1]}"
