#!/usr/bin/env python3
"""Constants of the tensor-core form of the 22 partial rounds of the CUDA Poseidon-GL permutation (csrc/poseidon_tc.cuh).

In a partial round only lane 0 meets an S-box, so with N = the MDS matrix with column 0 cleared and v = column 0 of the MDS matrix

    y(n+1) = N y(n) + v s_n + C(5+n),     s_n = y(n)_0 ^ 7          (y(n) = S-box inputs of round 4+n, n = 0..21)

and everything except the 22 scalars s_n is LINEAR in y(0) and in the earlier s_k:

    t_n   = y(n)_0 = <U_n, y(0)> + sum_{k<n} mu_{n-k} s_k + kappa_n        U_n = row 0 of N^n, mu_d = (N^(d-1) v)_0
    y(22) = N^22 y(0) + sum_k N^(21-k) v s_k + kappaF

The linear parts are products of CONSTANT matrices over F_p with per-permutation vectors -- dense contractions, batched over the 32
permutations of a warp on the tensor cores: the vector is taken as its bytes exactly as it lies in shared memory (K index = byte
position), the constant as the 8 byte limbs of  coefficient * 2^(8 b) mod p, and  mma.sync.m16n8k32.u8.u8.s32  sums the byte products
exactly (K <= 288: sums < 2^25); the 8 limb sums of an output recombine as  sum_b' D_b' 2^(8 b')  mod p.  The rounds are cut into
blocks of 8: the contributions of the s_k of earlier blocks come out of the block's GEMM, those of the block itself are 28 lazily
accumulated multiply-adds on the integer pipes.  All additive constants are in Montgomery form (c 2^64 mod p) like the rest of the
CUDA permutation; the multiplicative ones are plain (the maps are linear, the scaling passes through).

Per-permutation row in shared memory (288 bytes):  [0,88) y_1..y_11 | 88: the byte 1 (carries the additive constant) | 89..95: 0 |
[96 + 8 n, +8): slot n = the GEMM part of t_n, later s_n (n = 0..21) | [272, 288): 0.

Writes pil2_stark_js_b200/csrc/poseidon_tc_consts.inc: the A-operand fragments (the constants), 16 bytes per (k-step, m-tile, lane), for
the three block GEMMs (8 rows = the 8 rounds of the block; 3 / 5 / 7 k-steps) and the final GEMM (2 x 8 rows = the 12 output lanes,
9 k-steps), and mu_1..mu_7.  Self-check: a pure-Python model that mirrors the kernel's fragment indexing (an emulated m16n8k32 MMA over
the 32 lanes of a warp) reproduces the textbook partial rounds; tests/test_oracle_spec.py checks the model against the oracle.
"""
import importlib.util
import pathlib

ROOT = pathlib.Path(__file__).resolve().parents[1]
_spec = importlib.util.spec_from_file_location("gen_f64", ROOT / "tools" / "gen_poseidon_f64_consts.py")
_g = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_g)
P, R, M, RC, mds, plain_perm = _g.P, _g.R, _g.M, _g.RC, _g.mds, _g.plain_perm

ROW_BYTES = 288
BLOCK = 8                       # rounds per block
NBLOCKS = 3                     # rounds 0..7, 8..15, 16..21
KSTEPS_BLOCK = [3, 5, 7]        # K = 96 + 64 b bytes
KSTEPS_FINAL = 9


def matvec(A, x):
    return [sum(a * b for a, b in zip(row, x)) % P for row in A]


def matmul(A, B):
    return [[sum(A[i][k] * B[k][j] for k in range(12)) % P for j in range(12)] for i in range(12)]


N = [[0 if j == 0 else M[i][j] for j in range(12)] for i in range(12)]
V = [M[i][0] for i in range(12)]
I12 = [[int(i == j) for j in range(12)] for i in range(12)]
NPOW = [I12]
for _ in range(22):
    NPOW.append(matmul(NPOW[-1], N))
C = [[RC[12 * r + i] * R % P for i in range(12)] for r in range(30)]          # Montgomery-form additive constants

U = [NPOW[n][0] for n in range(23)]                                           # U[n] = row 0 of N^n
MU = [0] + [matvec(NPOW[d - 1], V)[0] for d in range(1, 23)]                  # mu_d, d = 1..22


def kappa_vec(n):
    """sum_{k<n} N^(n-1-k) C(5+k)."""
    acc = [0] * 12
    for k in range(n):
        acc = [(a + b) % P for a, b in zip(acc, matvec(NPOW[n - 1 - k], C[5 + k]))]
    return acc


KAPPA = [kappa_vec(n)[0] for n in range(23)]
KAPPA_F = kappa_vec(22)
G = [matvec(NPOW[21 - k], V) for k in range(22)]                              # G[k] = N^(21-k) v
F = NPOW[22]


def weights(a, c, g):
    """Field-element weight of every byte position of the row for an output with coefficients a[1..11] on y, constant c, g[k] on slot k."""
    W = [0] * ROW_BYTES
    for j in range(1, 12):
        for b in range(8):
            W[8 * (j - 1) + b] = a[j] * (1 << (8 * b)) % P
    W[88] = c % P
    for k, gk in enumerate(g):
        for b in range(8):
            W[96 + 8 * k + b] = gk * (1 << (8 * b)) % P
    return W


def block_rows(b):
    """Weights of the 8 output rows of block b: row j = the GEMM part of t_(8b+j)."""
    rows = []
    for j in range(BLOCK):
        n = BLOCK * b + j
        if n > 21 or n == 0:            # t_0 = y_0 never goes through the GEMM; rounds 22, 23 do not exist
            rows.append([0] * ROW_BYTES)
            continue
        g = [MU[n - k] if k < BLOCK * b else 0 for k in range(22)]
        rows.append(weights(U[n], KAPPA[n], g))
    return rows


def final_rows():
    rows = []
    for i in range(16):
        if i >= 12:
            rows.append([0] * ROW_BYTES)
            continue
        rows.append(weights(F[i], KAPPA_F[i], [G[k][i] for k in range(22)]))
    return rows


def frag_index(s, tile, lane):
    """u64 word index of the first of the two words (a0 | a1 << 32, a2 | a3 << 32) of a lane: one 16-byte shared-memory load per MMA."""
    return ((s * 4 + tile) * 32 + lane) * 2


def fragments(rows8, ksteps):
    """A-operand fragments of an 8-row group.  The M index of the MMA is (limb b', output row j): m-tile i holds limbs 2i (tile rows 0..7 =
    output rows 0..7) and 2i+1 (tile rows 8..15), so that lane 4g+t ends up with all 8 limb sums of output row g.  The K index is the
    byte position in the permutation's row: logical k = 4t+i <-> byte 32s + 8t + i, logical 16+4t+i <-> byte 32s + 8t + 4 + i, which
    makes the B fragment (the data) of lane 4g+t one 8-byte load at byte 32s + 8t of the row of permutation g of the n-tile."""
    out = [0] * (ksteps * 256)
    for s in range(ksteps):
        for tile in range(4):
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                a = [0, 0, 0, 0]
                for i in range(4):
                    a[0] |= ((rows8[g][32 * s + 8 * t + i] >> (8 * (2 * tile))) & 0xFF) << (8 * i)
                    a[1] |= ((rows8[g][32 * s + 8 * t + i] >> (8 * (2 * tile + 1))) & 0xFF) << (8 * i)
                    a[2] |= ((rows8[g][32 * s + 8 * t + 4 + i] >> (8 * (2 * tile))) & 0xFF) << (8 * i)
                    a[3] |= ((rows8[g][32 * s + 8 * t + 4 + i] >> (8 * (2 * tile + 1))) & 0xFF) << (8 * i)
                out[frag_index(s, tile, lane)] = a[0] | (a[1] << 32)
                out[frag_index(s, tile, lane) + 1] = a[2] | (a[3] << 32)
    return out


def tables():
    blocks = [fragments(block_rows(b), KSTEPS_BLOCK[b]) for b in range(NBLOCKS)]
    fr = final_rows()
    final = fragments(fr[:8], KSTEPS_FINAL) + fragments(fr[8:], KSTEPS_FINAL)
    return blocks, final


# ---------------------------------------------------------------------------------------------------------------------
# Model of the kernel: a warp of 32 permutations, shared-memory rows as bytearrays, the MMA emulated per lane
# ---------------------------------------------------------------------------------------------------------------------
def mma_m16n8k32(a, b, d):
    """Emulates mma.sync.m16n8k32.u8.u8.s32 over 32 lanes: a[lane] = 4 words, b[lane] = 2 words, d[lane] = 4 accumulators (in place)."""
    A = [[0] * 32 for _ in range(16)]
    B = [[0] * 8 for _ in range(32)]
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for i in range(16):
            row = g if (i < 4 or 8 <= i < 12) else g + 8
            col = 4 * t + (i & 3) + (16 if i >= 8 else 0)
            A[row][col] = (a[lane][i >> 2] >> (8 * (i & 3))) & 0xFF
        for i in range(8):
            B[4 * t + (i & 3) + (16 if i >= 4 else 0)][g] = (b[lane][i >> 2] >> (8 * (i & 3))) & 0xFF
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for i in range(4):
            row = g if i < 2 else g + 8
            col = 2 * t + (i & 1)
            d[lane][i] += sum(A[row][k] * B[k][col] for k in range(32))


def gemm_group(rows_mem, table, ksteps, ntiles=(0, 1, 2, 3)):
    """One 8-row group for the given n-tiles (8 permutations each) of the warp: returns {(perm, row j): value mod p}."""
    out = {}
    for q in ntiles:
        d = [[[0] * 4 for _ in range(32)] for _ in range(4)]               # [m-tile][lane][4]
        for s in range(ksteps):
            b = []
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                w = int.from_bytes(rows_mem[8 * q + g][32 * s + 8 * t: 32 * s + 8 * t + 8], "little")
                b.append([w & 0xFFFFFFFF, w >> 32])
            for tile in range(4):
                a = []
                for lane in range(32):
                    w0, w1 = table[frag_index(s, tile, lane)], table[frag_index(s, tile, lane) + 1]
                    a.append([w0 & 0xFFFFFFFF, w0 >> 32, w1 & 0xFFFFFFFF, w1 >> 32])
                mma_m16n8k32(a, b, d[tile])
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for e in range(2):
                limbs = []
                for tile in range(4):
                    limbs += [d[tile][lane][e], d[tile][lane][2 + e]]
                assert all(v < (1 << 25) for v in limbs)
                out[(8 * q + 2 * t + e, g)] = sum(v << (8 * bp) for bp, v in enumerate(limbs)) % P
    return out


def model_partial_warp(ys, blocks=None, final=None):
    """ys: 32 states (S-box inputs of round 4, any representatives, Montgomery form) -> S-box inputs of round 26, via the kernel's data flow."""
    if blocks is None:
        blocks, final = tables()
    mem = [bytearray(ROW_BYTES) for _ in range(32)]
    for p, y in enumerate(ys):
        for j in range(1, 12):
            mem[p][8 * (j - 1): 8 * j] = int(y[j]).to_bytes(8, "little")
        mem[p][88] = 1
    rinv = pow(R, P - 2, P)
    sbox = lambda t: pow(t * rinv % P, 7, P) * R % P                      # Montgomery-form S-box
    for b in range(NBLOCKS):
        lin = gemm_group(mem, blocks[b], KSTEPS_BLOCK[b])
        for p in range(32):
            s_blk = []
            for j in range(BLOCK):
                n = BLOCK * b + j
                if n > 21:
                    break
                acc = ys[p][0] if n == 0 else lin[(p, j)]
                for i, s in enumerate(s_blk):
                    acc += MU[j - i] * s
                s = sbox(acc % P)
                s_blk.append(s)
                mem[p][96 + 8 * n: 104 + 8 * n] = s.to_bytes(8, "little")
    outs = [[0] * 12 for _ in range(32)]
    for half in (0, 1):                                                      # two n-tiles per pass: both row groups, then their rows are overwritten
        for rg in (0, 1):
            res = gemm_group(mem, final[rg * KSTEPS_FINAL * 256:(rg + 1) * KSTEPS_FINAL * 256], KSTEPS_FINAL, ntiles=(2 * half, 2 * half + 1))
            for (perm, j), v in res.items():
                if 8 * rg + j < 12:
                    outs[perm][8 * rg + j] = v
    return outs


def textbook_partial(y):
    """Montgomery-form states: S-box inputs of round 4 -> S-box inputs of round 26."""
    rinv = pow(R, P - 2, P)
    s = [v * rinv % P for v in y]
    for r in range(4, 26):
        s = mds([pow(s[0], 7, P)] + s[1:])
        s = [(a + RC[12 * (r + 1) + i]) % P for i, a in enumerate(s)]
    return [v * R % P for v in s]


def main():
    import random
    rnd = random.Random(7)
    blocks, final = tables()
    ys = [[rnd.randrange(1 << 64) for _ in range(12)] for _ in range(32)]    # any 64-bit representatives
    ys[3] = [(1 << 64) - 1] * 12
    ys[4] = [0] * 12
    got = model_partial_warp(ys, blocks, final)
    for p in range(32):
        assert got[p] == textbook_partial(ys[p]), "tensor-core model != textbook partial rounds (perm %d)" % p
    text = ("/* Tensor-core form of the 22 partial rounds: A-operand fragments (a0 | a1 << 32, a2 | a3 << 32) per (k-step, m-tile, lane) of the block GEMMs\n"
            "   (3 / 5 / 7 k-steps) and of the final GEMM (2 row groups x 9 k-steps), then mu_1..mu_7 -- generated by tools/gen_poseidon_tc_consts.py */\n")
    words = [w for blk in blocks for w in blk] + final
    assert len(words) == (3 + 5 + 7 + 18) * 256
    text += "#define POSEIDON_TC_WORDS %d\n" % len(words)
    text += "__device__ __align__(16) const unsigned long long POSEIDON_TC_FRAGS[POSEIDON_TC_WORDS] = {\n"
    text += "\n".join("    " + ", ".join("0x%016xULL" % w for w in words[i:i + 6]) + "," for i in range(0, len(words), 6)) + "\n};\n"
    text += "__constant__ unsigned long long POSEIDON_TC_MU[8] = {0ULL, " + ", ".join("0x%xULL" % MU[d] for d in range(1, 8)) + "};\n"
    out = ROOT / "pil2_stark_js_b200/csrc/poseidon_tc_consts.inc"
    out.write_text(text)
    print("ok:", len(words), "fragment words ->", out, " mu:", [hex(MU[d]) for d in range(1, 8)])


if __name__ == "__main__":
    main()
