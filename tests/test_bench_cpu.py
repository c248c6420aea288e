"""bench.py pieces that run without a GPU: the reference arm (CPU port of the path, one JSON line in the contract's shape), the
synthetic-input generator and the FRI step schedule."""
import json
import pathlib
import subprocess
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "s" and d["higher_is_better"] is False and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "tiny"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_synthetic_generator_and_fri_schedule():
    import bench
    assert bench.fri_steps(24) == [24, 20, 16, 12, 8, 4] and bench.fri_steps(13) == [13, 9, 5, 4] and bench.fri_steps(4) == [4]
    a = bench.splitmix_field(0x5EED0003, 0, 1000)
    assert a.dtype == np.uint64 and (a < np.uint64(bench.P)).all()
    # splitmix64 of index 0 with that seed, by hand
    z = (0x5EED0003 + 0x9E3779B97F4A7C15) & (2**64 - 1)
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    z ^= z >> 31
    assert int(a[0]) == (z - bench.P if z >= bench.P else z)
    assert np.array_equal(bench.splitmix_field(7, 10, 5), bench.splitmix_field(7, 0, 15)[10:])


def test_ours_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", "tiny"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def _oracle_step_buffers(n_bits, cols, blow, seed, n_queries):
    """What one bench step leaves behind, built by the oracle (torch CPU tensors stand in for the device buffers)."""
    import torch
    import bench
    from oracle import gl_oracle as C
    ext_bits = n_bits + blow
    E = 1 << ext_bits
    src = bench.splitmix_field(seed, 0, cols << n_bits)
    dst = C.lde(src, cols, n_bits, ext_bits)
    nodes = C.merkelize(dst, cols, E)
    steps = bench.fri_steps(ext_bits)
    chal = [np.ascontiguousarray(bench.splitmix_field(seed + 2 + s, 0, 3)) for s in range(len(steps))]
    pol = bench.splitmix_field(seed + 1, 0, 3 << steps[0]).reshape(-1, 3)
    fri_pol, fri_rows, fri_nodes = [pol], [], []
    gs = 1 << (steps[0] - steps[1])
    rows0 = np.ascontiguousarray(pol.reshape(gs, 1 << steps[1], 3).transpose(1, 0, 2)).reshape(-1)
    fri_rows.append(rows0); fri_nodes.append(C.merkelize(rows0, 3 * gs, 1 << steps[1]))
    for s in range(1, len(steps)):
        nxt = steps[s + 1] if s + 1 < len(steps) else None
        p, rows = C.fri_fold(fri_pol[-1], steps[s - 1], steps[s], nxt, steps[0], [int(x) for x in chal[s]])
        fri_pol.append(p)
        if nxt is not None:
            fri_rows.append(rows); fri_nodes.append(C.merkelize(rows, 3 << (steps[s] - nxt), 1 << nxt))
    rng = np.random.default_rng(7)
    queries = rng.integers(0, E, size=n_queries, dtype=np.uint64)
    q_rows = np.empty(n_queries * cols, dtype=np.uint64)
    q_sib = np.empty(n_queries * ext_bits * 4, dtype=np.uint64)
    for k, q in enumerate(queries):
        row, sib = C.group_proof(dst, nodes, cols, E, int(q))
        q_rows[k * cols:(k + 1) * cols] = row
        q_sib[k * ext_bits * 4:(k + 1) * ext_bits * 4] = sib.reshape(-1)
    fq_rows, fq_sib = [], []
    qs = queries.copy()
    for s in range(len(steps) - 1):
        w, h = 3 << (steps[s] - steps[s + 1]), 1 << steps[s + 1]
        qs = qs % np.uint64(h)
        r_, s_ = np.empty(n_queries * w, dtype=np.uint64), np.empty(n_queries * max(1, steps[s + 1]) * 4, dtype=np.uint64)
        for k, q in enumerate(qs):
            row, sib = C.group_proof(fri_rows[s], fri_nodes[s], w, h, int(q))
            r_[k * w:(k + 1) * w] = row
            s_[k * steps[s + 1] * 4:(k + 1) * steps[s + 1] * 4] = sib.reshape(-1)
        fq_rows.append(r_); fq_sib.append(s_)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).reshape(-1).view(np.int64).copy())
    return dict(dst=t(dst), nodes=t(nodes), root=[int(x) for x in nodes[-4:]], queries=queries, q_rows=q_rows, q_sib=q_sib, steps=steps, chal=chal,
                fri_pol=[t(p) for p in fri_pol], fri_nodes=[t(n) for n in fri_nodes], fq_rows=fq_rows, fq_sib=fq_sib)


def test_parity_spot_check_accepts_oracle_buffers_and_flags_corruption():
    """The check bench.py runs on the real cfg3 buffers, driven here with buffers the oracle produced (must pass) and with
    single corrupted words in the extended buffer, a leaf digest, a sibling and a FRI layer (each must be reported)."""
    import torch
    import bench
    n_bits, cols, blow, seed, nq = 9, 24, 1, 0x5EED0003, 8
    b = _oracle_step_buffers(n_bits, cols, blow, seed, nq)
    run = lambda **kw: bench.parity_spot_check(torch, n_bits, cols, blow, seed, kw.get("dst", b["dst"]), kw.get("nodes", b["nodes"]), b["root"],
                                               b["queries"], kw.get("q_rows", b["q_rows"]), kw.get("q_sib", b["q_sib"]), b["steps"], b["chal"],
                                               kw.get("fri_pol", b["fri_pol"]), b["fri_nodes"], kw.get("fq_rows", b["fq_rows"]), b["fq_sib"],
                                               n_cols_checked=cols, n_leaves_checked=1 << (n_bits + blow + 2))
    ok = run()
    assert ok["status"] == "ok" and ok["columns"] == cols and ok["main_paths"] == nq and ok["fri_fold_links"] == nq * (len(b["steps"]) - 1)
    d2 = b["dst"].clone(); d2[5 * cols + 3] ^= 1
    assert run(dst=d2)["status"] == "MISMATCH"
    n2 = b["nodes"].clone(); n2[4 * 17] ^= 1
    assert any("leaf digest" in f for f in run(nodes=n2)["failures"])
    s2 = b["q_sib"].copy(); s2[7] ^= np.uint64(1)
    assert any("main tree" in f for f in run(q_sib=s2)["failures"])
    fp = [p.clone() for p in b["fri_pol"]]; fp[-1][0] ^= 1
    bad_final = run(fri_pol=fp)
    hits_q0 = any(int(q) % (1 << b["steps"][-1]) == 0 for q in b["queries"])
    assert (bad_final["status"] == "MISMATCH") == hits_q0
    fr = [r.copy() for r in b["fq_rows"]]; fr[1][4] ^= np.uint64(1)
    assert any("FRI layer" in f or "fold link" in f for f in run(fq_rows=fr)["failures"])
