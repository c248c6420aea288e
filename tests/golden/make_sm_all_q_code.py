#!/usr/bin/env python3
"""Extract the quotient-constraint program of the reference's sm_all AIR from the verifier the reference generated for its golden
proof (test/compressor/verifier.circom, template VerifyEvaluations0 :277-499) into the program format of the reference's own
expression evaluators: a list of {op, dest, src} records (src/prover/prover_helpers.js:87-110, src/stark/stark_verify.js:222-241).

Run in the build container only (reads /root/reference).  The circom template is that program printed line by line (one signal per
record), so the translation is mechanical:
    CMul()(a, b)                                  -> mul
    [a[0] + b[0], a[1] + b[1], a[2] + b[2]]       -> add      (likewise sub)
    [k + c[0], c[1], c[2]] / [e[0] - k, e[1], e[2]] / [k - e[0], -e[1], -e[2]] / [e[0] - publics[i], ...]
                                                  -> add / sub with a base-field operand (number / public)
    [t[0] * k, t[1] * k, t[2] * k]                -> mul by a number
    signal tmp_k[3] <== evals[i];                 -> copy
References: evals[i] -> {type: eval}; challengesStage2/3[i] -> {type: challenge, stage 2/3}; challengeQ -> stage 4; challengeXi ->
{type: x} (stark_verify.js:257-260); publics[i]; Zh -> {type: Zi, boundaryId 0 = everyRow}.  Output: tests/golden/sm_all_q_code.json
with the verifier-side program ("qVerifier") and what is needed to run the same program on the prover side over the extended
domain (evMap / cmPolsMap / mapSectionsN as implied by MapValues0 :708- and CalculateFRIPolValue0 :541-672)."""
import json, pathlib, re

REF = pathlib.Path("/root/reference/test/compressor/verifier.circom")
OUT = pathlib.Path(__file__).resolve().parent / "sm_all_q_code.json"


def ref_of(tok):
    tok = tok.strip()
    m = re.fullmatch(r"tmp_(\d+)", tok)
    if m:
        return {"type": "tmp", "id": int(m.group(1)), "dim": 3}
    m = re.fullmatch(r"evals\[(\d+)\]", tok)
    if m:
        return {"type": "eval", "id": int(m.group(1)), "dim": 3}
    m = re.fullmatch(r"challengesStage(\d)\[(\d+)\]", tok)
    if m:
        return {"type": "challenge", "stage": int(m.group(1)), "stageId": int(m.group(2)), "dim": 3}
    if tok == "challengeQ":
        return {"type": "challenge", "stage": 4, "stageId": 0, "dim": 3}
    if tok == "challengeXi":
        return {"type": "x", "dim": 3}
    if tok == "Zh":
        return {"type": "Zi", "boundaryId": 0, "dim": 3}
    m = re.fullmatch(r"publics\[(\d+)\]", tok)
    if m:
        return {"type": "public", "id": int(m.group(1)), "dim": 1}
    if re.fullmatch(r"\d+", tok):
        return {"type": "number", "value": tok, "dim": 1}
    raise ValueError("unknown operand: " + tok)


def strip0(tok):
    """'name[0]' -> 'name' for the first component of a triple"""
    tok = tok.strip()
    if re.fullmatch(r"publics\[\d+\]", tok):
        return tok                                   # a scalar with its own index
    return tok[:-3] if tok.endswith("[0]") else tok


def translate(lines):
    code = []
    for ln in lines:
        ln = ln.strip()
        m = re.fullmatch(r"signal tmp_(\d+)\[3\] <== (.*);", ln)
        if not m:
            continue
        dest = {"type": "tmp", "id": int(m.group(1)), "dim": 3}
        rhs = m.group(2).strip()
        mm = re.fullmatch(r"CMul\(\)\((.*), (.*)\)", rhs)
        if mm:
            code.append({"op": "mul", "dest": dest, "src": [ref_of(mm.group(1)), ref_of(mm.group(2))]})
            continue
        if not rhs.startswith("["):
            code.append({"op": "copy", "dest": dest, "src": [ref_of(rhs)]})
            continue
        comps = [c.strip() for c in rhs[1:-1].split(",")]
        assert len(comps) == 3, ln
        c0 = comps[0]
        mm = re.fullmatch(r"(.+?) ([+\-*]) (.+)", c0)
        assert mm, ln
        a, op, b = strip0(mm.group(1)), mm.group(2), strip0(mm.group(3))
        ra, rb = ref_of(a), ref_of(b)
        # consistency of components 1, 2 with the F3g semantics of a mixed operation (f3g.js:47-104)
        def comp(name, k):
            return f"{name}[{k}]"
        for k in (1, 2):
            if op == "*":
                want = f"{comp(a, k)} * {b}"
            elif ra["dim"] == 3 and rb["dim"] == 3:
                want = f"{comp(a, k)} {op} {comp(b, k)}"
            elif ra["dim"] == 1:
                want = comp(b, k) if op == "+" else "-" + comp(b, k)
            else:
                want = comp(a, k)
            assert comps[k].replace("  ", " ") == want, (ln, comps[k], want)
        code.append({"op": {"+": "add", "-": "sub", "*": "mul"}[op], "dest": dest, "src": [ra, rb]})
    return code


def main():
    text = REF.read_text().splitlines()
    start = next(i for i, l in enumerate(text) if "template parallel VerifyEvaluations0()" in l)
    end = next(i for i in range(start, len(text)) if "signal xAcc[2][3]" in text[i])
    code = translate(text[start:end])
    assert code[-1]["dest"]["id"] == 168 and code[-1]["src"][1]["type"] == "Zi"
    # tree layout of the proof (MapValues0 :708-): stage -> (number of polynomials, dim)
    cm_pols = []
    for stage, n, dim in ((1, 15, 1), (2, 2, 3), (3, 7, 3), (4, 2, 3)):
        for k in range(n):
            cm_pols.append({"stage": stage, "stagePos": k * dim, "dim": dim, "stageId": k})
    first = {1: 0, 2: 15, 3: 17, 4: 24}
    ev = []       # evals index -> evMap entry, read off CalculateFRIPolValue0 :541-672
    def add(kind, col, prime, stage=None, dim=1):
        if kind == "const":
            ev.append({"type": "const", "id": col, "prime": prime})
        else:
            ev.append({"type": "cm", "id": first[stage] + col // dim, "prime": prime})
    for c in range(6):
        add("const", c, 0)
    for c in (6, 7, 8):
        add("const", c, 0); add("const", c, 1)
    for c in (0, 1):
        add("cm", c, 0, 1); add("cm", c, 1, 1)
    for c in (2, 3, 4, 7, 8, 9, 10, 11, 12):
        add("cm", c, 0, 1)
    add("cm", 13, 1, 1)
    add("cm", 14, 0, 1); add("cm", 14, 1, 1)
    add("cm", 0, 0, 2, 3); add("cm", 0, 1, 2, 3); add("cm", 3, 0, 2, 3)
    for k in (0, 1, 2):
        add("cm", 3 * k, 0, 3, 3); add("cm", 3 * k, 1, 3, 3)
    for k in (3, 4, 5, 6):
        add("cm", 3 * k, 0, 3, 3)
    add("cm", 0, 0, 4, 3); add("cm", 3, 0, 4, 3)
    assert len(ev) == 43
    out = {
        "source": "pil2-stark-js test/compressor/verifier.circom VerifyEvaluations0 :277-499 (+ MapValues0 :708-, CalculateFRIPolValue0 :541-672)",
        "starkInfo": {"nStages": 3, "qDim": 3, "qDeg": 2, "nConstants": 9, "openingPoints": [0, 1],
                      "mapSectionsN": {"cm1": 15, "cm2": 6, "cm3": 21, "cm4": 6},
                      "boundaries": [{"name": "everyRow"}], "cmPolsMap": cm_pols, "evMap": ev,
                      "starkStruct": {"nBits": 10, "nBitsExt": 11, "nQueries": 8, "steps": [{"nBits": 11}, {"nBits": 7}, {"nBits": 3}]}},
        "qVerifier": {"code": code},
    }
    OUT.write_text(json.dumps(out, separators=(",", ":")))
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(code), "records")


if __name__ == "__main__":
    main()
