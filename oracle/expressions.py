"""CPU restatement of the reference's constraint-expression evaluation.  TEST INFRASTRUCTURE ONLY.

Follows
  src/prover/prover_helpers.js:23-110   calculateExps / compileCode: one straight-line program of {op, dest, src} records run at
                                        every row of a domain ("n" or "ext"); the reference turns it into a JS function body
  src/prover/prover_helpers.js:112-265  setRef / getRef / evalMap: how a reference becomes a buffer access (row offset for
                                        `prime`, stage buffer + stagePos for `cm`, Zi_ext / x_ext / xDivXSubXi_ext tables)
  src/stark/stark_verify.js:222-300     executeCode: the same program format evaluated once, at the challenge point, on the
                                        proof's `evals` (verifier side)
  src/helpers/f3g.js:47-104             add / sub / mul on mixed base-field (int) and cubic-extension ([a0,a1,a2]) operands
Parity: PINNED -- tests/test_oracle_expressions.py runs the quotient program the reference generated for its golden sm_all proof
(test/compressor/verifier.circom VerifyEvaluations0 :277-499, translated by tests/golden/make_sm_all_q_code.py) through
execute_code on the proof's evals and obtains Q(xi) = evals[41] + xi^N evals[42], the relation stark_verify.js:138-151 checks."""
from .gl_spec import P, f3_mul, inv, root_of_unity, SHIFT


# ---- F3g arithmetic on mixed operands (f3g.js:47-104): int = base field, list of 3 = extension ---------------------------
def f_add(a, b):
    if isinstance(a, int):
        if isinstance(b, int):
            return (a + b) % P
        return [(a + b[0]) % P, b[1], b[2]]
    if isinstance(b, int):
        return [(a[0] + b) % P, a[1], a[2]]
    return [(a[0] + b[0]) % P, (a[1] + b[1]) % P, (a[2] + b[2]) % P]


def f_sub(a, b):
    if isinstance(a, int):
        if isinstance(b, int):
            return (a - b) % P
        return [(a - b[0]) % P, (-b[1]) % P, (-b[2]) % P]
    if isinstance(b, int):
        return [(a[0] - b) % P, a[1], a[2]]
    return [(a[0] - b[0]) % P, (a[1] - b[1]) % P, (a[2] - b[2]) % P]


def f_mul(a, b):
    if isinstance(a, int):
        if isinstance(b, int):
            return (a * b) % P
        return [(a * b[0]) % P, (a * b[1]) % P, (a * b[2]) % P]
    if isinstance(b, int):
        return [(a[0] * b) % P, (a[1] * b) % P, (a[2] * b) % P]
    return f3_mul(a, b)


def _apply(op, src):
    if op == "add":
        return f_add(src[0], src[1])
    if op == "sub":
        return f_sub(src[0], src[1])
    if op == "mul":
        return f_mul(src[0], src[1])
    if op == "muladd":                                       # stark_verify.js:235
        return f_add(f_mul(src[0], src[1]), src[2])
    if op == "copy":
        return src[0]
    raise ValueError("Invalid op:" + str(op))


# ---- verifier side: executeCode (stark_verify.js:222-300) ---------------------------------------------------------------
def execute_code(code, ctx):
    """ctx: dict with evals (list of F3), challenges[stage-1][id], publics, Z (and Z_fr / Z_lr / Z_frame<i> when used), starkInfo
    {nStages, boundaries}; optional consts, subproofValues, xDivXSubXi.  Returns the value of the last destination."""
    tmp = {}

    def get(r):
        t = r["type"]
        if t == "tmp":
            return tmp[r["id"]]
        if t == "const":
            return ctx["consts"][r["id"]]
        if t == "eval":
            return ctx["evals"][r["id"]]
        if t == "number":
            return int(r["value"]) % P
        if t == "public":
            return int(ctx["publics"][r["id"]])
        if t == "challenge":
            return ctx["challenges"][r["stage"] - 1][r["stageId"]]
        if t == "subproofValue":
            return ctx["subproofValues"][r["id"]]
        if t == "xDivXSubXi":
            return ctx["xDivXSubXi"][r["id"]]
        if t == "x":
            return ctx["challenges"][ctx["starkInfo"]["nStages"] + 1][0]
        if t == "Zi":
            b = ctx["starkInfo"]["boundaries"][r["boundaryId"]]
            if b["name"] == "everyRow":
                return ctx["Z"]
            if b["name"] == "firstRow":
                return ctx["Z_fr"]
            if b["name"] == "lastRow":
                return ctx["Z_lr"]
            raise ValueError("Invalid boundary: " + str(b["name"]))
        raise ValueError("Invalid reference type get: " + str(t))

    for c in code:
        res = _apply(c["op"], [get(s) for s in c["src"]])
        if c["dest"]["type"] != "tmp":
            raise ValueError("Invalid reference type set: " + str(c["dest"]["type"]))
        tmp[c["dest"]["id"]] = res
    return get(code[-1]["dest"])


# ---- prover side: calculateExps over a domain (prover_helpers.js:33-76, getRef :152-219, setRef :112-150) -------------------
def calculate_exps(ctx, code, dom):
    """Runs `code` at every row of the domain (dom = "n" or "ext").  ctx: dict with pilInfo {nConstants, mapSectionsN, cmPolsMap,
    openingPoints, boundaries}, nBits, nBitsExt, buffers const_n / const_ext / cm<stage>_n / cm<stage>_ext (flat row-major lists or
    numpy arrays, modified in place when the program stores to a cm / q / f destination), x_n / x_ext, Zi_ext, xDivXSubXi_ext,
    q_ext, f_ext, challenges, publics, evals, subproofValues.  Returns nothing (like the reference with ret == false)."""
    n_bits, ext_bits = ctx["nBits"], ctx["nBitsExt"]
    N = 1 << (n_bits if dom == "n" else ext_bits)
    extend_bits = ext_bits - n_bits
    info = ctx["pilInfo"]

    def row_of(i, prime):
        if not prime:
            return i
        if dom == "n":
            nxt = prime + N if prime < 0 else prime
        else:
            nxt = ((prime + (1 << n_bits)) << extend_bits) if prime < 0 else (prime << extend_bits)   # :160-166 (N there is the base size)
        return (i + nxt) % N

    # NOTE prover_helpers.js:158-166 computes `next` for dom == "ext" from ctx.extN: (prime + extN) << extendBits for negative
    # primes; modulo extN that equals prime << extendBits, which is what the line above reduces to as well.
    def pol_ref(pol_id):
        p = info["cmPolsMap"][pol_id]
        st = "cm%d" % p["stage"]
        return st + "_" + dom, p["stagePos"], info["mapSectionsN"][st], p["dim"]

    def get(r, i, tmp):
        t = r["type"]
        if t == "tmp":
            return tmp[r["id"]]
        if t == "const":
            return int(ctx["const_" + dom][r["id"] + row_of(i, r.get("prime", 0)) * info["nConstants"]])
        if t == "cm":
            name, off, size, dim = pol_ref(r["id"])
            pos = off + row_of(i, r.get("prime", 0)) * size
            b = ctx[name]
            return int(b[pos]) if dim == 1 else [int(b[pos]), int(b[pos + 1]), int(b[pos + 2])]
        if t == "number":
            return int(r["value"]) % P
        if t == "public":
            return int(ctx["publics"][r["id"]])
        if t == "challenge":
            return ctx["challenges"][r["stage"] - 1][r["stageId"]]
        if t == "subproofValue":
            return ctx["subproofValues"][r["id"]]
        if t == "eval":
            return ctx["evals"][r["id"]]
        if t == "xDivXSubXi":
            n_open = len(info["openingPoints"])
            b = ctx["xDivXSubXi_ext"]
            pos = 3 * (r["id"] + n_open * i)
            return [int(b[pos]), int(b[pos + 1]), int(b[pos + 2])]
        if t == "x":
            return int(ctx["x_" + dom][i])
        if t == "Zi":
            return int(ctx["Zi_ext"][zi_index(info, r["boundaryId"]) * (1 << ext_bits) + i])
        raise ValueError("Invalid reference type get: " + str(t))

    def put(r, i, val, tmp):
        t = r["type"]
        if t == "tmp":
            tmp[r["id"]] = val
        elif t in ("q", "f"):
            if dom != "ext":
                raise ValueError("Accessing %s in domain n" % t)
            b = ctx[t + "_ext"]
            if t == "f" or r.get("dim", 3) == 3:
                v = val if not isinstance(val, int) else [val, 0, 0]
                b[3 * i], b[3 * i + 1], b[3 * i + 2] = v
            else:
                b[i] = val
        elif t == "cm":
            name, off, size, dim = pol_ref(r["id"])
            pos = off + row_of(i, r.get("prime", 0)) * size
            b = ctx[name]
            if dim == 1:
                b[pos] = val
            else:
                v = val if not isinstance(val, int) else [val, 0, 0]
                b[pos], b[pos + 1], b[pos + 2] = v
        else:
            raise ValueError("Invalid reference type set: " + str(t))

    for i in range(N):
        tmp = {}
        for c in code:
            put(c["dest"], i, _apply(c["op"], [get(s, i, tmp) for s in c["src"]]), tmp)


def zi_index(info, boundary_id):
    """Index of the Zi_ext table of a boundary (prover_helpers.js:201-215)."""
    b = info["boundaries"][boundary_id]
    for k, o in enumerate(info["boundaries"]):
        if b["name"] == "everyFrame":
            if o["name"] == "everyFrame" and o.get("offsetMin") == b.get("offsetMin") and o.get("offsetMax") == b.get("offsetMax"):
                return k
        elif o["name"] == b["name"]:
            return k
    raise ValueError("Something went wrong")


def build_x_ext(ext_bits):
    """ctx.x_ext (stark_gen_helpers.js:136-143): x_i = shift * w_ext^i."""
    w, x, out = root_of_unity(ext_bits), SHIFT, []
    for _ in range(1 << ext_bits):
        out.append(x)
        x = (x * w) % P
    return out


def build_zi_every_row(n_bits, ext_bits):
    """Zi_ext of the everyRow boundary (buildZhInv, stark_gen_helpers.js): 1 / (x^N - 1) on the extended coset; x^N takes only
    2^extendBits distinct values there."""
    eb = ext_bits - n_bits
    sn = pow(SHIFT, 1 << n_bits, P)
    w = root_of_unity(eb) if eb else 1
    vals, acc = [], sn
    for _ in range(1 << eb):
        vals.append(inv((acc - 1) % P))
        acc = (acc * w) % P
    return [vals[i % (1 << eb)] for i in range(1 << ext_bits)]
