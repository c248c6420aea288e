"""Pins the pure-Python spec oracle against the reference's own vectors (CPU only)."""
import random
import pytest
from oracle import gl_spec as S
from oracle import sm_all

P = S.P


def test_poseidon_kats():
    # test/poseidon.test.js:13-38 (also test/glwasm.test.js:143-187)
    assert S.poseidon([0] * 8) == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca]
    assert S.poseidon(list(range(8)), [8, 9, 10, 11]) == [
        0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2]
    assert S.poseidon([P - 1] * 8, [P - 1] * 4) == [
        0xbe0085cfc57a8357, 0xd95af71847d05c09, 0xcf55a13d33c1c953, 0x95803a74f4530e82]


def test_f3_kats():
    # test/f3g.test.js:13-38
    assert S.f3_mul([1, 2, 3], [4, 5, P - 1]) == [17, 23, 18]
    a = [random.randrange(P) for _ in range(3)]
    assert S.f3_mul(a, S.f3_inv(a)) == [1, 0, 0]
    assert pow(S.W32, 1 << 31, P) == P - 1
    assert S.SHIFT_INV == 2635249152773512046


def test_ntt_roundtrip_and_definition():
    # test/fft.test.js:16-34 (round trip) + the DFT definition
    rnd = random.Random(1)
    a = [rnd.randrange(P) for _ in range(16)]
    A = S.ntt(a)
    w = S.root_of_unity(4)
    for k in range(16):
        assert A[k] == sum(a[j] * pow(w, j * k, P) for j in range(16)) % P
    assert S.intt(A) == a
    a3 = [[rnd.randrange(P) for _ in range(3)] for _ in range(64)]
    assert S.intt(S.ntt(a3)) == a3


def test_extend_pol_is_coset_evaluation():
    rnd = random.Random(2)
    a = [rnd.randrange(P) for _ in range(8)]
    coefs = S.intt(a)
    ext = S.extend_pol(a, 1)
    w = S.root_of_unity(4)
    for j in range(16):
        x = (S.SHIFT * pow(w, j, P)) % P
        assert ext[j] == sum(c * pow(x, i, P) for i, c in enumerate(coefs)) % P


def test_merkle_layout_sizes():
    # power-of-two heights: 8h-4 words; h=1 quirk: 8 words, root = zeros
    for k in range(1, 12):
        assert S.merkle_n_nodes(4 << k) == 8 * (1 << k) - 4
    assert S.merkle_n_nodes(4) == 8
    t = S.merkelize(list(range(1, 10)), 9, 1)
    assert S.merkle_root(t) == [0, 0, 0, 0]


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n,npols", [(256, 3), (256, 9), (33, 6), (5, 17), (2, 1)])
def test_merkle_selfconsistency(n, npols, split):
    # test/merklehash_p.test.js shapes incl. the non-power-of-two (33,6) case
    buff = [i + j * 1000 for i in range(n) for j in range(npols)]
    tree = S.merkelize(buff, npols, n, split)
    for idx in {0, 3 % n, n - 1, n // 2}:
        v, mp = S.get_group_proof(tree, idx)
        assert S.verify_group_proof(S.merkle_root(tree), mp, idx, v, split)
    with pytest.raises(IndexError):
        S.get_group_proof(tree, n)


def test_split_hash_shapes():
    # widths of test/glwasm.test.js:198-230; structural checks of linearhash_gpu.js:31-67
    for w in [0, 1, 2, 3, 4]:
        v = list(range(1, w + 1))
        assert S.linear_hash(v, True) == v + [0] * (4 - w)
    v = list(range(1, 9))           # one batch of 8 -> single sponge digest, passthrough of 4 words
    assert S.linear_hash(v, True) == S.linear_hash(v, False)
    v = list(range(1, 51))          # batch = 13 -> 4 batches -> 16 words -> sponge
    d = []
    for b in range(0, 50, 13):
        d += S.linear_hash(v[b:b + 13], False)
    assert S.linear_hash(v, True) == S.linear_hash(d, False)


# ------------------------------------------------------------------------------------------------
# Golden proof: test/compressor/verifier.proof.zkin.json
# ------------------------------------------------------------------------------------------------
def _transcript(golden):
    """Challenge order: test/compressor/verifier.circom:90-235 (== src/prover/prover.js:42-122)."""
    t = S.Transcript()
    r = golden["roots"]
    t.put(r["const"]); t.put(golden["publics"]); t.put(r["stage1"])
    ch = {"stage2": [t.get_field(), t.get_field()]}
    t.put(r["stage2"])
    ch["stage3"] = [t.get_field() for _ in range(3)]
    t.put(r["stage3"])
    ch["Q"] = t.get_field()
    t.put(r["stageQ"])
    ch["xi"] = t.get_field()
    for e in golden["evals"]:
        t.put(e)
    ch["fri"] = [t.get_field(), t.get_field()]
    steps = [t.get_field()]
    t.put(r["fri1"]); steps.append(t.get_field())
    t.put(r["fri2"]); steps.append(t.get_field())
    for e in golden["final_pol"]:
        t.put(e)
    steps.append(t.get_field())
    ch["fri_steps"] = steps
    t2 = S.Transcript()
    t2.put(steps[3])
    ch["queries"] = t2.get_permutations(8, 11)
    return ch


def test_golden_transcript_queries(golden):
    assert _transcript(golden)["queries"] == [891, 1628, 1228, 1991, 1856, 415, 833, 296]


def test_golden_root1_and_rootC(golden):
    buff, w = sm_all.committed_trace((1, 2))
    assert buff[(sm_all.N - 1) * w + 0] == golden["publics"][2] == 74469561660084004
    ext = S.interpolate(buff, w, 10, 11)
    tree1 = S.merkelize(ext, w, 2048)
    assert S.merkle_root(tree1) == golden["roots"]["stage1"]
    cbuff, cw = sm_all.constant_trace()
    cext = S.interpolate(cbuff, cw, 10, 11)
    treec = S.merkelize(cext, cw, 2048)
    assert S.merkle_root(treec) == golden["roots"]["const"]
    # the opened rows and paths of the proof are exactly what get_group_proof returns
    q = _transcript(golden)["queries"]
    for k, idx in enumerate(q):
        v, mp = S.get_group_proof(tree1, idx)
        assert v == golden["layer0"]["stage1"]["rows"][k]
        assert mp == golden["layer0"]["stage1"]["siblings"][k]
        v, mp = S.get_group_proof(treec, idx)
        assert v == golden["layer0"]["const"]["rows"][k]
        assert mp == golden["layer0"]["const"]["siblings"][k]


def test_golden_all_merkle_paths(golden):
    q = _transcript(golden)["queries"]
    for name in ["const", "stage1", "stage2", "stage3", "stageQ"]:
        for k, idx in enumerate(q):
            assert S.verify_group_proof(golden["roots"][name], golden["layer0"][name]["siblings"][k], idx,
                                        golden["layer0"][name]["rows"][k])
    for k, idx in enumerate(q):
        assert S.verify_group_proof(golden["roots"]["fri1"], golden["fri1"]["siblings"][k], idx % 128,
                                    golden["fri1"]["rows"][k])
        assert S.verify_group_proof(golden["roots"]["fri2"], golden["fri2"]["siblings"][k], idx % 8,
                                    golden["fri2"]["rows"][k])


def test_golden_fri_fold_links(golden):
    # fri.js:107-174 on the fixture: s1 group --challenge[1]--> element of s2 group --challenge[2]--> finalPol
    ch = _transcript(golden)
    q = ch["queries"]
    steps = [11, 7, 3]
    for k, q0 in enumerate(q):
        g1 = [golden["fri1"]["rows"][k][3 * i:3 * i + 3] for i in range(16)]
        q1 = q0 % 128
        shift1 = S.SHIFT                                            # polBits = 11, no squaring yet
        ev = S.fri_verify_fold(g1, steps[0], shift1, ch["fri_steps"][1], q1)
        g2 = [golden["fri2"]["rows"][k][3 * i:3 * i + 3] for i in range(16)]
        q2 = q1 % 8
        assert g2[q1 // 8] == ev
        shift2 = pow(S.SHIFT, 1 << 4, P)                            # squared (11-7) times
        ev2 = S.fri_verify_fold(g2, steps[1], shift2, ch["fri_steps"][2], q2)
        assert golden["final_pol"][q2] == ev2


def test_fri_fold_prover_matches_verifier():
    # prover-side fold (fri.js:22-81) against the verifier-side link on a synthetic chain
    rnd = random.Random(5)
    steps = [8, 5, 2]
    pol = [[rnd.randrange(P) for _ in range(3)] for _ in range(256)]
    ch = [[rnd.randrange(P) for _ in range(3)] for _ in range(3)]
    r0 = S.fri_fold(steps, 0, pol, ch[0])
    assert r0["pol"] == pol
    r1 = S.fri_fold(steps, 1, r0["pol"], ch[1])
    r2 = S.fri_fold(steps, 2, r1["pol"], ch[2])
    assert r2["tree"] is None and len(r2["proof"]) == 4
    for q0 in [0, 7, 100, 255]:
        q1 = q0 % 32
        v, mp = S.get_group_proof(r0["tree"], q1)
        assert S.verify_group_proof(r0["proof"]["root"], mp, q1, v)
        grp = [v[3 * i:3 * i + 3] for i in range(8)]
        assert grp == [pol[q1 + 32 * j] for j in range(8)]
        ev = S.fri_verify_fold(grp, 8, S.SHIFT, ch[1], q1)
        assert ev == r1["pol"][q1]
        q2 = q1 % 4
        v2, mp2 = S.get_group_proof(r1["tree"], q2)
        grp2 = [v2[3 * i:3 * i + 3] for i in range(8)]
        assert grp2[q1 // 4] == ev
        ev2 = S.fri_verify_fold(grp2, 5, pow(S.SHIFT, 8, P), ch[2], q2)
        assert ev2 == r2["pol"][q2]


def test_fp64_partial_round_constants_model_vs_oracle():
    """csrc/poseidon.cuh runs the partial rounds with deferred round constants (tools/gen_poseidon_f64_consts.py): the generator's
    pure-Python model of that form, for every block length it emits tables for, is the oracle permutation; the emitted table is the
    one the generator would write now (i.e. the committed .inc is not stale)."""
    import importlib.util
    import pathlib
    import random
    root = pathlib.Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("gen_f64", root / "tools" / "gen_poseidon_f64_consts.py")
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    rnd = random.Random(11)
    for nr in g.NRS:
        st = [rnd.randrange(S.P) for _ in range(12)]
        assert g.model_perm(st, nr) == S.poseidon_perm(st)
    assert g.plain_perm(list(range(12))) == S.poseidon_perm(list(range(12)))
    # the conversion constants: magic + halves of (c - K), all exact doubles below 2^53
    lane0, final = g.deferred_tables(g.R, 22)
    for c in lane0 + final:
        lo, hi = g.halves(c)
        assert lo < 2.0 ** 53 and hi < 2.0 ** 53
        assert ((int(lo) - g.MAGIC) + ((int(hi) - g.MAGIC) << 32) + g.K) % S.P == c
    inc = (root / "pil2_stark_js_b200/csrc/poseidon_rc_f64p.inc").read_text()
    assert "%.1f" % g.halves(lane0[0])[0] in inc and "POSEIDON_RC_F64P_22[66]" in inc


def test_tensor_core_partial_rounds_model_vs_oracle():
    """csrc/poseidon_tc.cuh (opt-in leaf kernel) runs the 22 partial rounds as constant-matrix GEMMs over byte limbs plus 71 lazy
    multiply-adds (tools/gen_poseidon_tc_consts.py).  The generator's lane-exact model of that data flow (emulated m16n8k32 MMA, the
    kernel's fragment and row layout) equals the oracle's rounds 4..25 for a warp of states incl. all-ones / zero words, and the
    committed table is the one the generator writes now."""
    import importlib.util
    import pathlib
    import random
    root = pathlib.Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("gen_tc", root / "tools" / "gen_poseidon_tc_consts.py")
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    R = (1 << 64) % S.P
    rinv = pow(R, S.P - 2, S.P)
    rnd = random.Random(13)
    ys = [[rnd.randrange(1 << 64) for _ in range(12)] for _ in range(32)]
    ys[0] = [(1 << 64) - 1] * 12
    ys[1] = [0] * 12
    ys[2] = [S.P - 1] * 12
    got = g.model_partial_warp(ys)
    for y, out in zip(ys, got):
        s = [v * rinv % S.P for v in y]                       # Montgomery form -> plain S-box inputs of round 4
        for r in range(4, 26):
            s = [pow(s[0], 7, S.P)] + s[1:]
            s = [sum(S.MDS[i][j] * s[j] for j in range(12)) % S.P for i in range(12)]
            s = [(a + S.RC[12 * (r + 1) + i]) % S.P for i, a in enumerate(s)]
        assert out == [v * R % S.P for v in s]
    # mu_d of the in-block multiply-adds: small integers below 2^52 (the kernel's lazy accumulator relies on it)
    assert g.MU[1] == 25 and all(g.MU[d] < (1 << 52) for d in range(1, 8)) and all(g.MU[d] < (1 << 32) for d in range(1, 5))
    blocks, final = g.tables()
    inc = (root / "pil2_stark_js_b200/csrc/poseidon_tc_consts.inc").read_text()
    assert "0x%016xULL" % blocks[1][100] in inc and "0x%016xULL" % final[-1] in inc and "POSEIDON_TC_WORDS 8448" in inc


def test_lazy_f3_product_accumulator_model():
    """csrc/gl.cuh gl3_mul accumulates the 12 partial products of an F3 product lazily in three 64-bit columns with carry counts and
    reduces each coordinate once (gl_acc_mac / gl_acc_reduce: low 128 bits through 2^64 = 2^32 - 1 and 2^96 = -1, what lies above 2^128
    comes off as top * 2^32 since 2^128 = -2^32 mod p).  The same word-level steps in Python give the oracle's F3 product, for random
    and for all-ones (non-canonical, maximal carries) operands."""
    import random
    M32, M64 = (1 << 32) - 1, (1 << 64) - 1

    def mac(A, x, y):                                     # A = [a0, c0, a1, c1, a2, c2]: 64-bit columns + carry counts
        x0, x1, y0, y1 = x & M32, x >> 32, y & M32, y >> 32
        for col, p in ((0, x0 * y0), (2, x0 * y1), (2, x1 * y0), (4, x1 * y1)):
            s = A[col] + p
            A[col], A[col + 1] = s & M64, A[col + 1] + (s >> 64)

    def reduce(A):
        a0, c0, a1, c1, a2, c2 = A
        l1 = (a0 >> 32) + (a1 & M32)
        h0 = (a2 & M32) + (a1 >> 32) + (l1 >> 32)
        h1 = (a2 >> 32) + (h0 >> 32)
        top = c2 + (h1 >> 32)
        h0 = (h0 & M32) + c0
        h1 = (h1 & M32) + c1 + (h0 >> 32)
        top += h1 >> 32
        lo = ((l1 & M32) << 32) | (a0 & M32)
        hi = ((h1 & M32) << 32) | (h0 & M32)
        assert top < 1 << 31
        # exactness of the decomposition, then the reduction identities
        total = a0 + (c0 << 64) + ((a1 + (c1 << 64)) << 32) + ((a2 + (c2 << 64)) << 64)
        assert total == lo + (hi << 64) + (top << 128)
        r = (lo + (hi & M32) * ((1 << 32) - 1) - (hi >> 32)) % S.P           # gl_reduce128: 2^64 = 2^32 - 1, 2^96 = -1
        return (r - (top << 32)) % S.P                                        # 2^128 = -2^32

    def f3_mul_lazy(a, b):
        r0, r1, r2 = [0] * 6, [0] * 6, [0] * 6
        for acc, terms in ((r0, ((0, 0), (1, 2), (2, 1))), (r1, ((0, 1), (1, 0), (1, 2), (2, 1), (2, 2))), (r2, ((0, 2), (1, 1), (2, 0), (2, 2)))):
            for i, j in terms:
                mac(acc, a[i], b[j])
        return [reduce(r0), reduce(r1), reduce(r2)]

    rnd = random.Random(17)
    cases = [([M64] * 3, [M64] * 3), ([S.P - 1] * 3, [S.P - 1] * 3), ([0, 1, 0], [0, 0, 1])]
    cases += [([rnd.randrange(1 << 64) for _ in range(3)], [rnd.randrange(1 << 64) for _ in range(3)]) for _ in range(200)]
    for a, b in cases:
        assert f3_mul_lazy(a, b) == S.f3_mul([x % S.P for x in a], [x % S.P for x in b])


def test_three_level_post_order_walk_model():
    """csrc/merkle.cuh merkle_levels_kernel: thread i walks the depth-LV subtree over input nodes [2^LV i, 2^LV (i+1)) in post-order with a
    rolled loop (leaf pair, leaf pair, their parent, ...).  The same state machine in Python -- pending level, parked left siblings,
    'a node exists iff its first input node does', missing nodes count as the zero pad -- reproduces every level of the oracle's tree
    for ragged sizes and LV = 1, 2, 3."""
    def level_offsets(p_in, n_in, lv):
        off, n = [p_in], n_in
        for _ in range(lv):
            off.append(off[-1] + 4 * (n + (n & 1)))
            n = (n + 1) // 2
        return off

    def walk(nodes, p_in, n_in, lv, i):
        first = i << lv
        if first >= n_in:
            return
        off = level_offsets(p_in, n_in, lv)
        left, cur, k, pending = {}, [0, 0, 0, 0], 0, 0
        for _ in range((1 << lv) - 1):
            if pending == 0:
                lvl, idx = 1, (first >> 1) + k
                n0 = 2 * idx
                a = list(nodes[p_in + 4 * n0:p_in + 4 * n0 + 4])
                b = list(nodes[p_in + 4 * (n0 + 1):p_in + 4 * (n0 + 1) + 4])       # n0 + 1 == n_in: the stored zero pad
                k += 1
            else:
                lvl, idx = pending + 1, (first >> (pending + 1)) + ((k - 1) >> pending)
                a, b = left[pending], cur
            if (idx << lvl) < n_in:
                cur = [int(x) for x in S.poseidon_perm([int(x) for x in a] + [int(x) for x in b] + [0, 0, 0, 0])[:4]]
                nodes[off[lvl] + 4 * idx:off[lvl] + 4 * idx + 4] = cur
            else:
                cur = [0, 0, 0, 0]
            j = (k - 1) >> (lvl - 1)
            if (j & 1) and lvl < lv:
                pending = lvl
            else:
                if lvl < lv:
                    left[lvl] = cur
                pending = 0

    import random
    rng = random.Random(3)
    for height in (8, 11, 13, 16, 21, 37):
        leaves = [rng.randrange(S.P) for _ in range(4 * height)]
        want = S.merkelize([int(x) for x in leaves], 4, height)["nodes"]   # width 4: the leaf digest of a row is the row itself
        for lv in (1, 2, 3):
            nodes = [0] * len(want)
            nodes[:4 * height] = [int(x) for x in leaves]
            p_in, n = 0, height
            while n > 1:
                remaining, m = 0, n
                while m > 1:
                    m, remaining = (m + 1) // 2, remaining + 1
                step = min(lv, remaining)                                  # the launcher never walks past the root either
                for i in range((n + (1 << step) - 1) >> step):
                    walk(nodes, p_in, n, step, i)
                for _ in range(step):
                    p_in += 4 * (n + (n & 1))
                    n = (n + 1) // 2
            assert nodes == [int(x) for x in want], (height, lv)
