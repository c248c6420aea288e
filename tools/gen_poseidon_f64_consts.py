#!/usr/bin/env python3
"""Constants of the FP64-resident partial rounds of the CUDA Poseidon-GL permutation (csrc/poseidon.cuh, poseidon_partial_f64).

During the 22 partial rounds lanes 1..11 are never touched by an S-box, so they stay on the FP64 pipe as two planes of exact
integers (value = L + 2^32 H mod p) and only lane 0 crosses to the integer pipes once per round.  The round constants of lanes
1..11 are deferred: with t(r) = s(r) - d(r) (d(r) supported on lanes 1..11, d(4) = 0),

    s(r+1) = M sigma(s(r)) + C(r+1)   ==>   t(r+1) = M sigma(t(r)) + v0(r+1) e0,    v(r+1) = M d(r) + C(r+1),  d(r+1) = v(r+1) with lane 0 cleared

so a round adds ONE scalar (to lane 0, folded into the double -> integer conversion constant) and the whole vector v(26) comes
back after the last partial MDS.  Everything is in Montgomery form (c * 2^64 mod p), like the rest of the CUDA permutation.

The conversion constants carry the magic number 2^52 + 2^51: adding them to a plane value y (|y| < 2^50) gives a double in
[2^52, 2^53) whose mantissa is y + 2^51 + c; the bias K = 2^51 + 2^83 of the pair of planes is taken out of c beforehand.

Writes  pil2_stark_js_b200/csrc/poseidon_rc_f64p.inc: one table per candidate block length NR (rounds 4..3+NR on the FP64 pipe, the
remaining partial rounds in limb form): NR-1 lane-0 pairs, then 12 pairs for the MDS output of the block's last round.  Self-check: a pure-Python model of the deferred form reproduces the textbook permutation (and the KAT of test/poseidon.test.js:13-20);
tests/test_oracle_spec.py checks the same model against the oracle.
"""
import pathlib

import re

ROOT = pathlib.Path(__file__).resolve().parents[1]
CSRC = ROOT / "pil2_stark_js_b200/csrc"

P = 0xFFFFFFFF00000001
R = (1 << 64) % P
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]            # glwasm.js:430; + 8 on M[0][0] (:431)
M = [[CIRC[(j - i) % 12] + (8 if i == j == 0 else 0) for j in range(12)] for i in range(12)]


def load_rc():
    """The 360 plain round constants, recovered from the Montgomery-form tables the CUDA code already includes
    (poseidon_rc.inc: round 0 whole; poseidon_rc_limbs.inc: rounds 1..29 as 22/22/20-bit limbs)."""
    strip = lambda t: re.sub(r"/\*.*?\*/", "", t, flags=re.S)
    r0 = [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]+)ULL", strip((CSRC / "poseidon_rc.inc").read_text()))]
    lim = [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]+)u", strip((CSRC / "poseidon_rc_limbs.inc").read_text()))]
    assert len(r0) == 12 and len(lim) == 30 * 36
    mont = r0 + [lim[3 * k] | (lim[3 * k + 1] << 22) | (lim[3 * k + 2] << 44) for k in range(29 * 12)]
    rinv = pow(R, P - 2, P)
    return [v * rinv % P for v in mont]


RC = load_rc()


def plain_perm(state):
    """Textbook form (glwasm.js:359-390): 30 x (add constants, S-box layer, MDS)."""
    s = list(state)
    for r in range(30):
        s = [(a + RC[12 * r + i]) % P for i, a in enumerate(s)]
        s = [pow(x, 7, P) for x in s] if (r < 4 or r >= 26) else [pow(s[0], 7, P)] + s[1:]
        s = mds(s)
    return s
MAGIC = (1 << 52) + (1 << 51)
NRS = [6, 8, 10, 12, 14, 16, 18, 22]        # candidate lengths of the FP64 block (even: the planes are re-normalised every second round)
K = ((1 << 51) + (1 << 83)) % P


def mds(v):
    return [sum(M[i][j] * v[j] for j in range(12)) % P for i in range(12)]


def deferred_tables(scale, nr=22):
    """(v0[r] for r = 5..3+nr, v(4+nr)) for an FP64 block covering rounds 4..3+nr, constants scaled by `scale` (1: plain, R: Montgomery)."""
    C = [[RC[12 * r + i] * scale % P for i in range(12)] for r in range(30)]
    d = [0] * 12
    lane0 = []
    for r in range(4, 4 + nr):
        v = [(a + b) % P for a, b in zip(mds(d), C[r + 1])]
        if r < 3 + nr:
            lane0.append(v[0])
            d = [0] + v[1:]
        else:
            final = v
    return lane0, final


def model_perm(state, nr=22):
    """Plain-domain model of the restructured permutation: rounds 4..3+nr in the deferred-constant form, the others as usual."""
    C = [[RC[12 * r + i] for i in range(12)] for r in range(30)] + [[0] * 12]
    lane0, final = deferred_tables(1, nr)
    s = [(a + b) % P for a, b in zip(state, C[0])]
    for r in range(4):
        s = [(a + b) % P for a, b in zip(mds([pow(x, 7, P) for x in s]), C[r + 1])]
    for r in range(4, 4 + nr):
        s = mds([pow(s[0], 7, P)] + s[1:])
        if r < 3 + nr:
            s[0] = (s[0] + lane0[r - 4]) % P
        else:
            s = [(a + b) % P for a, b in zip(s, final)]
    for r in range(4 + nr, 30):
        s = [pow(x, 7, P) for x in s] if r >= 26 else [pow(s[0], 7, P)] + s[1:]
        s = [(a + b) % P for a, b in zip(mds(s), C[r + 1])]
    return s


def halves(c):
    c = (c - K) % P
    return float(MAGIC + (c & 0xFFFFFFFF)), float(MAGIC + (c >> 32))


def main():
    import random
    rnd = random.Random(5)
    text = ("/* FP64-resident partial rounds 4..3+NR: 2^52 + 2^51 + 32-bit halves of (constant - K) in Montgomery form; per NR: NR-1 lane-0 "
            "pairs (after the MDS of rounds 4..2+NR), then 12 pairs (after round 3+NR) -- generated by tools/gen_poseidon_f64_consts.py */\n")
    for nr in NRS:
        for _ in range(2):
            st = [rnd.randrange(P) for _ in range(12)]
            assert model_perm(st, nr) == plain_perm(st), "deferred-constant model != textbook permutation"
        assert model_perm(list(range(12)), nr)[0] == 0xd64e1e3efc5b8e9e
        lane0, final = deferred_tables(R, nr)
        vals = []
        for c in lane0 + final:
            lo, hi = halves(c)
            assert int(lo) - MAGIC < (1 << 32) and int(hi) - MAGIC < (1 << 32)
            vals += [lo, hi]
        rows = ["    " + ", ".join("%.1f" % v for v in vals[i:i + 4]) + "," for i in range(0, len(vals), 4)]
        text += "__constant__ double POSEIDON_RC_F64P_%d[%d] = {\n" % (nr, len(vals)) + "\n".join(rows) + "\n};\n"
    out = ROOT / "pil2_stark_js_b200/csrc/poseidon_rc_f64p.inc"
    out.write_text(text)
    print("ok: tables for NR in", NRS, "->", out)


if __name__ == "__main__":
    main()
