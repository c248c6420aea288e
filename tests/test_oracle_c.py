"""Pins the C oracle (oracle/gl_oracle.c) against the reference vectors and the Python spec (CPU only)."""
import random
import numpy as np
import pytest
from oracle import gl_spec as S
from oracle import gl_oracle as C
from oracle import sm_all

P = S.P


def rand_field(rnd, n):
    return [rnd.randrange(P) for _ in range(n)]


def test_poseidon_kats_c():
    assert list(C.poseidon_perm([0] * 12))[:4] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca]
    assert list(C.poseidon_perm(list(range(12))))[:4] == [0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2]
    assert list(C.poseidon_perm([P - 1] * 12))[:4] == [0xbe0085cfc57a8357, 0xd95af71847d05c09, 0xcf55a13d33c1c953, 0x95803a74f4530e82]
    rnd = random.Random(7)
    for _ in range(20):
        st = rand_field(rnd, 12)
        assert [int(v) for v in C.poseidon_perm(st)] == S.poseidon_perm(st)


@pytest.mark.parametrize("bits,npols", [(0, 3), (1, 2), (3, 1), (5, 2), (8, 5)])
def test_ntt_vs_spec(bits, npols):
    rnd = random.Random(bits * 31 + npols)
    src = rand_field(rnd, npols << bits)
    assert [int(v) for v in C.ntt(src, npols, bits)] == S.fft_p(src, npols, bits)
    assert [int(v) for v in C.ntt(src, npols, bits, inverse=True)] == S.fft_p(src, npols, bits, inverse=True)


@pytest.mark.parametrize("bits,ext,npols", [(3, 4, 1), (5, 6, 3), (6, 8, 2), (4, 4, 2)])
def test_lde_vs_spec(bits, ext, npols):
    rnd = random.Random(bits + 100 * ext)
    src = rand_field(rnd, npols << bits)
    assert [int(v) for v in C.lde(src, npols, bits, ext)] == S.interpolate(src, npols, bits, ext)


def test_ntt_reference_pattern_roundtrip():
    # test/fft_p.test.js:120-190 pattern (v = row index), round trip ifft(fft(x)) == x, threads vs serial
    n_bits, npols = 14, 5
    src = np.repeat(np.arange(1 << n_bits, dtype=np.uint64), npols)
    f = C.ntt(src, npols, n_bits)
    assert np.array_equal(C.ntt(src, npols, n_bits, threads=1), f)
    assert np.array_equal(C.ntt(f, npols, n_bits, inverse=True), src)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("w", [0, 1, 2, 3, 4, 5, 8, 9, 15, 16, 24, 25, 32, 33, 50])
def test_linear_hash_widths(w, split):
    # width sweep of test/glwasm.test.js:198-230
    rnd = random.Random(w)
    v = rand_field(rnd, w)
    assert [int(x) for x in C.linear_hash(v, split)] == S.linear_hash(v, split)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n,npols", [(256, 3), (256, 9), (33, 6), (1, 9), (2, 1), (7, 40)])
def test_merkle_vs_spec(n, npols, split):
    buff = [i + j * 1000 for i in range(n) for j in range(npols)]
    nodes = C.merkelize(buff, npols, n, split)
    tree = S.merkelize(buff, npols, n, split)
    assert [int(x) for x in nodes] == tree["nodes"]
    for idx in {0, n - 1, n // 2}:
        row, sib = C.group_proof(buff, nodes, npols, n, idx)
        v, mp = S.get_group_proof(tree, idx)
        assert [int(x) for x in row] == v and [[int(x) for x in s] for s in sib] == mp
    with pytest.raises(IndexError):
        C.group_proof(buff, nodes, npols, n, n)


def test_golden_roots_c(golden):
    buff, w = sm_all.committed_trace((1, 2))
    ext = C.lde(buff, w, 10, 11)
    nodes = C.merkelize(ext, w, 2048)
    assert [int(x) for x in nodes[-4:]] == golden["roots"]["stage1"]
    cbuff, cw = sm_all.constant_trace()
    nodes_c = C.merkelize(C.lde(cbuff, cw, 10, 11), cw, 2048)
    assert [int(x) for x in nodes_c[-4:]] == golden["roots"]["const"]


def test_merkle_2_18_x10_selfconsistent():
    # test/merklehash_p.test.js:79 shape (2^18 rows x 10), pattern i + 1000 j
    n, npols = 1 << 18, 10
    i = np.arange(n, dtype=np.uint64)[:, None]
    j = np.arange(npols, dtype=np.uint64)[None, :]
    buff = (i + 1000 * j).reshape(-1)
    nodes = C.merkelize(buff, npols, n)
    assert nodes.size == 8 * n - 4
    for idx in (3, n - 1):
        row, sib = C.group_proof(buff, nodes, npols, n, idx)
        assert S.verify_group_proof([int(x) for x in nodes[-4:]], [[int(x) for x in s] for s in sib], idx,
                                    [int(x) for x in row])


def test_fri_fold_vs_spec():
    rnd = random.Random(11)
    steps = [9, 5, 2]
    pol = [[rnd.randrange(P) for _ in range(3)] for _ in range(512)]
    ch = [rnd.randrange(P) for _ in range(3)]
    r1 = S.fri_fold(steps, 1, pol, ch)
    pol2, rows = C.fri_fold(np.array(pol, dtype=np.uint64), 9, 5, 2, 9, ch)
    assert pol2.tolist() == r1["pol"]
    assert rows.tolist() == S.transposed_buffer(r1["pol"], 2)
    ch2 = [rnd.randrange(P) for _ in range(3)]
    r2 = S.fri_fold(steps, 2, r1["pol"], ch2)
    pol3, rows3 = C.fri_fold(pol2, 5, 2, None, 9, ch2)
    assert rows3 is None and pol3.tolist() == r2["pol"]


@pytest.mark.parametrize("n_bits,ext_bits,steps", [(6, 7, [7, 4, 2]), (8, 10, [10, 6, 3]), (7, 8, [8, 3])])
def test_fri_chain_keeps_low_degree(n_bits, ext_bits, steps):
    """The property FRI.verify relies on (fri.js:158-171), independent of how the fold is coded: folding the evaluations of a
    polynomial of degree < N on the coset 7<w_ext> through the whole chain leaves a final polynomial of degree < 2^(last -
    (ext - n)); a polynomial of full degree does not."""
    rng = np.random.default_rng(n_bits)
    base = rng.integers(0, S.P, size=(1 << n_bits) * 3, dtype=np.uint64)
    low = C.lde(base, 3, n_bits, ext_bits).reshape(-1, 3)                 # an F3 polynomial of degree < N, coordinate-wise
    full = rng.integers(0, S.P, size=(1 << ext_bits, 3), dtype=np.uint64)
    for pol, expect_low in ((low, True), (full, False)):
        cur = pol
        for s in range(1, len(steps)):
            ch = [int(x) for x in rng.integers(0, S.P, size=3, dtype=np.uint64)]
            nxt = steps[s + 1] if s + 1 < len(steps) else None
            cur, rows = C.fri_fold(cur, steps[s - 1], steps[s], nxt, steps[0], ch)
            if nxt is not None:                                           # rows = getTransposedBuffer of the folded layer (fri.js:187-202)
                w = 1 << nxt
                assert np.array_equal(rows.reshape(w, -1, 3), cur.reshape(-1, w, 3).transpose(1, 0, 2))
        coeffs = S.intt([[int(v) for v in e] for e in cur])
        max_deg = 1 << (steps[-1] - (ext_bits - n_bits))
        high_is_zero = all(c == [0, 0, 0] for c in coeffs[max_deg:])
        assert high_is_zero == expect_low
