"""f4 (constraint-expression evaluation) in the oracle: pinned by the reference's golden proof.

The program is the quotient constraint of the sm_all AIR exactly as the reference generated it for the verifier of its golden proof
(tests/golden/sm_all_q_code.json <- test/compressor/verifier.circom VerifyEvaluations0 :277-499, extracted by
tests/golden/make_sm_all_q_code.py).  stark_verify.js:138-151 accepts a proof iff executeCode(qVerifier.code) on the proof's evals
equals Q(xi) = sum_i xi^(N i) evals[Q_i]; the golden proof must pass that check through the oracle's execute_code."""
import json
import pathlib

import numpy as np
import pytest

from oracle import expressions as X
from oracle import gl_spec as S
from test_oracle_f_rows import golden_challenges

P = S.P
ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def qcode():
    return json.loads((ROOT / "tests" / "golden" / "sm_all_q_code.json").read_text())


def golden_all_challenges(golden):
    """challenges[stage-1][id] of the golden proof: stage 2 (2), stage 3 (3), stage Q = 4 (1), xi = stage 5 (1)."""
    class T(S.Transcript):
        getField = S.Transcript.get_field
    t = T()
    r = golden["roots"]
    t.put(r["const"]); t.put(golden["publics"]); t.put(r["stage1"])
    st2 = [t.getField(), t.getField()]
    t.put(r["stage2"])
    st3 = [t.getField() for _ in range(3)]
    t.put(r["stage3"])
    q = [t.getField()]
    t.put(r["stageQ"])
    xi = [t.getField()]
    return [[], st2, st3, q, xi]


def test_golden_proof_satisfies_its_quotient_program(golden, qcode):
    ch = golden_all_challenges(golden)
    xi = ch[4][0]
    assert xi == golden_challenges(golden)[0]
    n_bits = 10
    xn = [1, 0, 0]
    for _ in range(1 << n_bits):
        xn = S.f3_mul(xn, xi)
    zh = S.f3_inv(X.f_sub(xn, 1))                                              # Z = 1 / (xi^N - 1): verifier.circom:293-294
    ctx = {"evals": golden["evals"], "challenges": ch, "publics": golden["publics"], "Z": zh, "starkInfo": qcode["starkInfo"]}
    res = X.execute_code(qcode["qVerifier"]["code"], ctx)
    q = X.f_add(golden["evals"][41], S.f3_mul(xn, golden["evals"][42]))       # stark_verify.js:140-147
    assert res == q
    # a corrupted evaluation must break the relation
    bad = dict(ctx, evals=[list(e) for e in golden["evals"]])
    bad["evals"][12][0] = (bad["evals"][12][0] + 1) % P
    assert X.execute_code(qcode["qVerifier"]["code"], bad) != q


def prover_side_program(qcode):
    """The same program as the prover runs it over the extended domain (prover_helpers.js getRef :152-219): every eval reference
    becomes the polynomial it evaluates (evMap), and the last record stores to q instead of a tmp."""
    ev = qcode["starkInfo"]["evMap"]
    code = []
    for c in qcode["qVerifier"]["code"]:
        src = []
        for s in c["src"]:
            if s["type"] == "eval":
                e = ev[s["id"]]
                src.append({"type": e["type"], "id": e["id"], "prime": e["prime"]})
            else:
                src.append(dict(s))
        code.append({"op": c["op"], "dest": dict(c["dest"]), "src": src})
    code[-1]["dest"] = {"type": "q", "id": 0, "dim": 3}
    return code


def make_domain_ctx(qcode, rng, n_bits=4, ext_bits=5):
    info = dict(qcode["starkInfo"])
    E = 1 << ext_bits
    ctx = {"pilInfo": info, "nBits": n_bits, "nBitsExt": ext_bits, "publics": [int(x) for x in rng.integers(0, P, size=3, dtype=np.uint64)],
           "challenges": [[], *[[[int(x) for x in rng.integers(0, P, size=3, dtype=np.uint64)] for _ in range(k)] for k in (2, 3, 1, 1)]],
           "const_ext": rng.integers(0, P, size=info["nConstants"] * E, dtype=np.uint64),
           "x_ext": np.array(X.build_x_ext(ext_bits), dtype=np.uint64), "Zi_ext": np.array(X.build_zi_every_row(n_bits, ext_bits), dtype=np.uint64),
           "q_ext": np.zeros(3 * E, dtype=np.uint64)}
    for st, w in info["mapSectionsN"].items():
        ctx[st + "_ext"] = rng.integers(0, P, size=w * E, dtype=np.uint64)
    return ctx


def test_prover_side_program_matches_the_verifier_at_one_row(qcode):
    """calculate_exps over the extended domain and execute_code at a point are the same program on different operand sources:
    feeding execute_code the values row i of the buffers holds (as `evals`) gives q_ext[i]."""
    rng = np.random.default_rng(3)
    ctx = make_domain_ctx(qcode, rng)
    code = prover_side_program(qcode)
    X.calculate_exps(ctx, code, "ext")
    info = ctx["pilInfo"]
    E, eb = 1 << ctx["nBitsExt"], ctx["nBitsExt"] - ctx["nBits"]
    for i in (0, 7, E - 1):
        evals = []
        for e in info["evMap"]:
            row = (i + (e["prime"] << eb)) % E
            if e["type"] == "const":
                evals.append(int(ctx["const_ext"][e["id"] + row * info["nConstants"]]))
            else:
                p = info["cmPolsMap"][e["id"]]
                b, w = ctx["cm%d_ext" % p["stage"]], info["mapSectionsN"]["cm%d" % p["stage"]]
                pos = p["stagePos"] + row * w
                evals.append(int(b[pos]) if p["dim"] == 1 else [int(b[pos]), int(b[pos + 1]), int(b[pos + 2])])
        # the verifier reads `x` as the evaluation challenge (stark_verify.js:257-260): give it x_ext[i] there
        vch = [list(c) for c in ctx["challenges"]]
        vch[4] = [int(ctx["x_ext"][i])]
        v = X.execute_code(qcode["qVerifier"]["code"], {"evals": evals, "challenges": vch, "publics": ctx["publics"],
                                                          "Z": int(ctx["Zi_ext"][i]), "starkInfo": info})
        v = v if not isinstance(v, int) else [v, 0, 0]
        assert [int(x) for x in ctx["q_ext"][3 * i:3 * i + 3]] == v


def test_mixed_dimension_arithmetic_matches_f3g():
    """f3g.js:47-104 on mixed operands (base-field scalar with an extension element)."""
    a, b, k = [5, 6, 7], [P - 1, 2, 3], 11
    assert X.f_add(k, a) == [16, 6, 7] and X.f_add(a, k) == [16, 6, 7] and X.f_add(a, b) == [4, 8, 10]
    assert X.f_sub(k, a) == [6, P - 6, P - 7] and X.f_sub(a, k) == [P - 6, 6, 7] and X.f_sub(3, 5) == P - 2
    assert X.f_mul(k, a) == [55, 66, 77] and X.f_mul(a, k) == [55, 66, 77] and X.f_mul(a, b) == S.f3_mul(a, b)
    assert X.f_mul([1, 2, 3], [4, 5, P - 1]) == [17, 23, 18]                   # test/f3g.test.js:33-38
