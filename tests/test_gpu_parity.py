"""GPU parity tests: the CUDA path (through the C ABI / Python mirror) against the CPU oracle, the reference's
known-answer vectors and the committed golden proof.  Bit-exact (integer field arithmetic): the tolerance is zero."""
import random
import numpy as np
import pytest

from oracle import gl_spec as S
from oracle import gl_oracle as C
from oracle import sm_all

pytestmark = pytest.mark.gpu
P = S.P


@pytest.fixture(scope="module")
def ctx():
    import pil2_stark_js_b200 as m
    return m.default_context(0)


def rnd_field(seed, n):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a)


# ---------------------------------------------------------------- extendAndMerkelize (host buffers, slab pipeline)
@pytest.mark.parametrize("bits,ext,npols,split", [(12, 13, 64, False), (12, 14, 80, False), (13, 14, 128, False), (12, 13, 96, True),
                                                   (12, 13, 48, False), (5, 6, 64, False), (12, 12, 64, False), (12, 13, 256, False), (12, 13, 320, False)])
def test_extend_and_merkelize_vs_oracle(ctx, bits, ext, npols, split):
    """stark_gen_helpers.js:388-412 = interpolate + merkelize.  Shapes 1-3 and 7-9 take the column-slab pipeline
    (npols % 16 == 0, >= 4 slabs, standard hash; the last two with mixed 32/64-column slabs); the others fall back to the
    whole-buffer path."""
    src = rnd_field(1000 + bits * 7 + npols, npols << bits)
    dst, nodes, root = ctx.extend_and_merkelize(src, npols, bits, ext, split)
    want_dst = C.lde(src, npols, bits, ext)
    assert np.array_equal(dst, want_dst)
    want_nodes = C.merkelize(want_dst, npols, 1 << ext, split=split)
    assert np.array_equal(nodes, want_nodes)
    assert np.array_equal(root, want_nodes[-4:])
    # root only (no downloads)
    _, _, root2 = ctx.extend_and_merkelize(src, npols, bits, ext, split, want_dst=False, want_nodes=False)
    assert np.array_equal(root2, root)


# ---------------------------------------------------------------- Poseidon / linear hash
def test_poseidon_kats(ctx):
    # test/poseidon.test.js:13-38
    assert [int(x) for x in ctx.poseidon([0] * 12)[:4]] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca]
    assert [int(x) for x in ctx.poseidon(list(range(12)))[:4]] == [0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2]
    assert [int(x) for x in ctx.poseidon([P - 1] * 12)[:4]] == [0xbe0085cfc57a8357, 0xd95af71847d05c09, 0xcf55a13d33c1c953, 0x95803a74f4530e82]


def test_poseidon_random_and_edge_states(ctx):
    states = [rnd_field(s, 12) for s in range(8)]
    states += [np.array([P - 1, 0, 1, 2**32 - 1, 2**32, 2**63, P - 2, 2**32 + 1, 0xFFFFFFFF00000000, 5, 6, 7], dtype=np.uint64)]
    for st in states:
        assert np.array_equal(ctx.poseidon(st), C.poseidon_perm(st))


@pytest.mark.parametrize("split", [False, True])
def test_linear_hash_width_sweep(ctx, split):
    # widths of test/glwasm.test.js:198-230
    for w in [0, 1, 2, 3, 4, 5, 8, 9, 15, 16, 24, 25, 32, 33, 50, 256]:
        v = rnd_field(100 + w, w)
        assert np.array_equal(ctx.linear_hash(v, split), C.linear_hash(v, split)), (w, split)


# ---------------------------------------------------------------- NTT
@pytest.mark.parametrize("bits,npols", [(0, 1), (1, 3), (2, 2), (3, 1), (5, 2), (8, 17), (9, 16), (10, 5), (12, 33), (13, 3)])
def test_ntt_vs_oracle(ctx, bits, npols):
    src = rnd_field(bits * 97 + npols, npols << bits)
    dst = np.empty_like(src)
    ctx.ntt(src, npols, bits, dst)
    assert np.array_equal(dst, C.ntt(src, npols, bits))
    ctx.ntt(src, npols, bits, dst, inverse=True)
    assert np.array_equal(dst, C.ntt(src, npols, bits, inverse=True))


def test_fft_p_reference_shapes(ctx):
    # test/fft_p.test.js:47 (5 bits x 2 cols), :120/:156 (18 bits x 5 cols), inputs v = row index
    from pil2_stark_js_b200 import fft_p
    for bits, npols in [(5, 2), (18, 5)]:
        src = np.repeat(np.arange(1 << bits, dtype=np.uint64), npols)
        dst = np.empty_like(src)
        fft_p.fft(src, npols, bits, dst)
        assert np.array_equal(dst, C.ntt(src, npols, bits))
        back = np.empty_like(src)
        fft_p.ifft(dst, npols, bits, back)
        assert np.array_equal(back, src)
        fft_p.ifft(src, npols, bits, dst)
        assert np.array_equal(dst, C.ntt(src, npols, bits, inverse=True))


# ---------------------------------------------------------------- LDE (interpolate)
@pytest.mark.parametrize("bits,ext,npols", [(0, 0, 2), (0, 1, 1), (1, 2, 3), (3, 4, 1), (3, 3, 2), (5, 8, 3), (9, 10, 16), (10, 11, 15),
                                            (10, 12, 7), (11, 12, 20), (13, 14, 5), (12, 15, 2)])
def test_lde_vs_oracle(ctx, bits, ext, npols):
    src = rnd_field(bits * 31 + ext * 7 + npols, npols << bits)
    dst = np.full(npols << ext, 0xDEADBEEF, dtype=np.uint64)   # prior contents must not matter
    ctx.lde(src, npols, bits, dst, ext)
    assert np.array_equal(dst, C.lde(src, npols, bits, ext))


def test_interpolate_reference_shapes(ctx):
    # test/fft_p.test.js:82 (3 bits x 1, extendPol) and :193 (18 bits x 5, ext 1)
    from pil2_stark_js_b200 import fft_p
    src = np.arange(8, dtype=np.uint64)
    dst = np.empty(16, dtype=np.uint64)
    fft_p.interpolate(src, 1, 3, dst, 4)
    assert [int(x) for x in dst] == S.extend_pol(list(range(8)), 1)
    bits, npols = 18, 5
    src = np.repeat(np.arange(1 << bits, dtype=np.uint64), npols)
    dst = np.empty(npols << (bits + 1), dtype=np.uint64)
    fft_p.interpolate(src, npols, bits, dst, bits + 1)
    assert np.array_equal(dst, C.lde(src, npols, bits, bits + 1))


# ---------------------------------------------------------------- Merkle
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n,npols", [(256, 3), (256, 9), (33, 6), (1, 9), (2, 1), (7, 40), (1025, 17), (5000, 8)])
def test_merkelize_vs_oracle(ctx, n, npols, split):
    # shapes of test/merklehash_p.test.js (pattern i + 1000 j) incl. the non-power-of-two (33,6)
    i = np.arange(n, dtype=np.uint64)[:, None]
    j = np.arange(npols, dtype=np.uint64)[None, :]
    buff = np.ascontiguousarray((i + 1000 * j).reshape(-1))
    nodes = ctx.merkelize(buff, npols, n, split)
    assert np.array_equal(nodes, C.merkelize(buff, npols, n, split))


def test_leaf_hash_tensor_core_form_vs_oracle(tmp_path):
    """The opt-in leaf kernel with the partial rounds on the tensor cores (PIL2GPU_LEAF_TC=1, csrc/poseidon_tc.cuh): same nodes as the
    oracle for ragged heights (clamped tail rows), widths with a partial last chunk, and all-(p-1) / zero rows.  Runs in a child
    process because the switch is read once per process."""
    import os, subprocess, sys, pathlib
    root = pathlib.Path(__file__).resolve().parents[1]
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import pil2_stark_js_b200 as m\n"
        "ctx = m.default_context(0)\n"
        "rng = np.random.default_rng(5)\n"
        "for k, (n, w) in enumerate([(4099, 40), (8192, 9), (5000, 256), (4096, 5)]):\n"
        "    buff = rng.integers(0, 0xFFFFFFFF00000001, size=n * w, dtype=np.uint64)\n"
        "    buff[:w] = 0xFFFFFFFF00000000; buff[w:2 * w] = 0\n"
        "    np.save(%r + '/in%%d.npy' %% k, buff); np.save(%r + '/out%%d.npy' %% k, ctx.merkelize(buff, w, n, False))\n"
    ) % (str(root), str(tmp_path), str(tmp_path))
    env = dict(os.environ, PIL2GPU_LEAF_TC="1")
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=600)
    for k, (n, w) in enumerate([(4099, 40), (8192, 9), (5000, 256), (4096, 5)]):
        buff = np.load(tmp_path / ("in%d.npy" % k))
        assert np.array_equal(np.load(tmp_path / ("out%d.npy" % k)), C.merkelize(buff, w, n, False))


@pytest.mark.parametrize("split", [False, True])
def test_merklehash_p_interface(ctx, split, tmp_path):
    # test/merklehash_p.test.js: merkelize -> getGroupProof -> verifyGroupProof, (2^18, 10); file save/restore :101-132
    from pil2_stark_js_b200 import buildMerkleHash, OutOfRange
    MH = buildMerkleHash(split)
    n, npols = 1 << 18, 10
    i = np.arange(n, dtype=np.uint64)[:, None]
    j = np.arange(npols, dtype=np.uint64)[None, :]
    pols = np.ascontiguousarray((i + 1000 * j).reshape(-1))
    tree = MH.merkelize(pols, npols, n)
    assert tree["elements"] is pols and tree["nodes"].size == 8 * n - 4
    assert np.array_equal(tree["nodes"], C.merkelize(pols, npols, n, split))
    idx = 3
    groupElements, mp = MH.getGroupProof(tree, idx)
    root = MH.root(tree)
    assert MH.verifyGroupProof(root, mp, idx, groupElements)
    assert S.verify_group_proof(root, mp, idx, groupElements, split)           # and by the independent CPU verifier
    bad = list(groupElements); bad[0] = (bad[0] + 1) % P
    assert not MH.verifyGroupProof(root, mp, idx, bad)
    with pytest.raises(OutOfRange):
        MH.getGroupProof(tree, n)
    f = tmp_path / "tree.bin"
    MH.writeToFile(tree, str(f))
    t2 = MH.readFromFile(str(f))
    assert t2["width"] == npols and t2["height"] == n
    assert np.array_equal(t2["elements"], pols) and np.array_equal(t2["nodes"], tree["nodes"])
    raw = np.fromfile(str(f), dtype="<u8")
    assert raw[0] == npols and raw[1] == n and raw.size == 2 + pols.size + tree["nodes"].size


# ---------------------------------------------------------------- golden proof (test/compressor/verifier.proof.zkin.json)
def _golden_queries(golden, T):
    t = T()
    r = golden["roots"]
    t.put(r["const"]); t.put(golden["publics"]); t.put(r["stage1"])
    t.getField(); t.getField()
    t.put(r["stage2"])
    for _ in range(3):
        t.getField()
    t.put(r["stage3"]); t.getField()
    t.put(r["stageQ"]); t.getField()
    for e in golden["evals"]:
        t.put(e)
    t.getField(); t.getField()
    steps = [t.getField()]
    t.put(r["fri1"]); steps.append(t.getField())
    t.put(r["fri2"]); steps.append(t.getField())
    for e in golden["final_pol"]:
        t.put(e)
    steps.append(t.getField())
    t2 = T()
    t2.put(steps[3])
    return steps, t2.getPermutations(8, 11)


def test_golden_transcript_on_gpu(ctx, golden):
    from pil2_stark_js_b200 import Transcript
    steps, q = _golden_queries(golden, Transcript)
    assert q == [891, 1628, 1228, 1991, 1856, 415, 833, 296]


def test_golden_commit_roots_and_paths(ctx, golden):
    # sm_all traces -> device-resident commit (LDE x2 + Merkle) -> root1 / rootC and every opened row + sibling path
    q = [891, 1628, 1228, 1991, 1856, 415, 833, 296]
    for trace_fn, name in [(sm_all.committed_trace, "stage1"), (sm_all.constant_trace, "const")]:
        buff, w = trace_fn()
        src = np.array(buff, dtype=np.uint64)
        tree, root = ctx.commit(src, w, 10, 11)
        assert [int(x) for x in root] == golden["roots"][name]
        rows, sib = tree.group_proofs(q)
        assert rows.tolist() == golden["layer0"][name]["rows"]
        assert sib.tolist() == golden["layer0"][name]["siblings"]
        elems, nodes = tree.download()
        assert np.array_equal(elems, C.lde(src, w, 10, 11))
        assert np.array_equal(nodes, C.merkelize(elems, w, 2048))
        tree.free()


def test_golden_merkle_paths_verify_on_gpu(ctx, golden):
    from pil2_stark_js_b200 import buildMerkleHash
    MH = buildMerkleHash(False)
    q = [891, 1628, 1228, 1991, 1856, 415, 833, 296]
    for name in ["const", "stage1", "stage2", "stage3", "stageQ"]:
        for k in (0, 5):
            assert MH.verifyGroupProof(golden["roots"][name], golden["layer0"][name]["siblings"][k], q[k], golden["layer0"][name]["rows"][k])
    for k in (1, 7):
        assert MH.verifyGroupProof(golden["roots"]["fri1"], golden["fri1"]["siblings"][k], q[k] % 128, golden["fri1"]["rows"][k])
        assert MH.verifyGroupProof(golden["roots"]["fri2"], golden["fri2"]["siblings"][k], q[k] % 8, golden["fri2"]["rows"][k])


def test_golden_fri_fold_links(ctx, golden):
    # Embed the opened groups of the fixture in full-size layers and fold them on the GPU with the transcript challenges:
    # s1 group --challenge[1]--> element of the s2 group --challenge[2]--> finalPol[q mod 8]  (fri.js:47-61)
    from pil2_stark_js_b200 import Transcript
    steps, q = _golden_queries(golden, Transcript)
    lay1 = rnd_field(1, 3 << 11).reshape(-1, 3).copy()
    lay2 = rnd_field(2, 3 << 7).reshape(-1, 3).copy()
    for k, q0 in enumerate(q):
        q1, q2 = q0 % 128, q0 % 8
        for i in range(16):
            lay1[i * 128 + q1] = golden["fri1"]["rows"][k][3 * i:3 * i + 3]
            lay2[i * 8 + q2] = golden["fri2"]["rows"][k][3 * i:3 * i + 3]
    p2, rows, nodes = ctx.fri_fold(lay1, 11, 7, 3, 11, steps[1])
    p3, _, _ = ctx.fri_fold(lay2, 7, 3, None, 11, steps[2])
    for k, q0 in enumerate(q):
        q1, q2 = q0 % 128, q0 % 8
        assert [int(x) for x in p2[q1]] == golden["fri2"]["rows"][k][3 * (q1 // 8):3 * (q1 // 8) + 3]
        assert [int(x) for x in p3[q2]] == golden["final_pol"][q2]


# ---------------------------------------------------------------- FRI vs oracle
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("steps", [[9, 5, 2], [12, 8, 4, 2], [10, 5, 0], [8, 7, 3], [6, 6, 2], [13, 7, 1], [14, 9, 4]])
def test_fri_chain_vs_oracle(ctx, steps, split):
    pol = rnd_field(sum(steps), 3 << steps[0]).reshape(-1, 3)
    rng = random.Random(7)
    # step 0: identity fold + commit of the first layer
    ch = [rng.randrange(P) for _ in range(3)]
    p, rows, nodes = ctx.fri_fold(pol, steps[0], steps[0], steps[1], steps[0], ch, split)
    assert np.array_equal(p, pol)
    exp_rows = np.array(S.transposed_buffer(pol.tolist(), steps[1]), dtype=np.uint64)
    assert np.array_equal(rows, exp_rows)
    assert np.array_equal(nodes, C.merkelize(exp_rows, 3 << (steps[0] - steps[1]), 1 << steps[1], split))
    cur = pol
    for s in range(1, len(steps)):
        ch = [rng.randrange(P) for _ in range(3)]
        nxt = steps[s + 1] if s + 1 < len(steps) else None
        p, rows, nodes = ctx.fri_fold(cur, steps[s - 1], steps[s], nxt, steps[0], ch, split)
        ep, erows = C.fri_fold(cur, steps[s - 1], steps[s], nxt, steps[0], ch)
        assert np.array_equal(p, ep)
        if nxt is not None:
            assert np.array_equal(rows, erows)
            assert np.array_equal(nodes, C.merkelize(erows, 3 << (steps[s] - nxt), 1 << nxt, split))
        else:
            assert rows is None and nodes is None
        cur = p


@pytest.mark.parametrize("steps", [[14, 6, 2], [13, 0], [15, 8, 1], [12, 5]])
def test_fri_wide_folds_vs_oracle(ctx, steps):
    """fri.js:38-41 puts no bound on the fold width; folds wider than one kernel handles (2^6) are split into chunks with the
    challenge squared once per halving -- same field elements as the reference's ifft + Horner (the oracle)."""
    pol = rnd_field(sum(steps) + 3, 3 << steps[0]).reshape(-1, 3)
    rng = random.Random(11)
    cur = pol
    for s in range(1, len(steps)):
        ch = [rng.randrange(P) for _ in range(3)]
        nxt = steps[s + 1] if s + 1 < len(steps) else None
        p, rows, nodes = ctx.fri_fold(cur, steps[s - 1], steps[s], nxt, steps[0], ch)
        ep, erows = C.fri_fold(cur, steps[s - 1], steps[s], nxt, steps[0], ch)
        assert np.array_equal(p, ep), (steps, s)
        if nxt is not None:
            assert np.array_equal(rows, erows)
            assert np.array_equal(nodes, C.merkelize(erows, 3 << (steps[s] - nxt), 1 << nxt))
        cur = p


def test_fri_class_prove_then_verify(ctx):
    # the reference's integration style (test/stark/helpers.js): prove, then the verifier must accept
    from pil2_stark_js_b200 import FRI, buildMerkleHash, Transcript
    MH = buildMerkleHash(False)
    struct = {"nBits": 7, "nBitsExt": 8, "nQueries": 4, "steps": [{"nBits": 8}, {"nBits": 5}, {"nBits": 2}]}
    fri = FRI(struct, MH)
    # a genuine low-degree polynomial (deg < 2^7) evaluated on the coset 7*<w_8>, in F3
    coef = rnd_field(3, 3 << 7).reshape(-1, 3)
    cols = np.ascontiguousarray(coef.reshape(-1))            # 128 rows x 3 "columns" = the three F3 components
    ext = np.empty(3 << 8, dtype=np.uint64)
    tmp = np.empty_like(cols)
    ctx.ntt(cols, 3, 7, tmp)                                 # evaluations on <w_7>, then LDE to the coset
    ctx.lde(tmp, 3, 7, ext, 8)
    pol = ext.reshape(-1, 3)
    tr = Transcript()
    proof, trees, challenges = [], [], []
    cur = pol
    for step in range(3):
        ch = tr.getField(); challenges.append(ch)
        r = fri.fold(step, cur, ch)
        cur = r["pol"]
        proof.append(r["proof"] if isinstance(r["proof"], dict) else r["proof"])
        trees.append(r["tree"])
        tr.put(r["proof"]["root"] if isinstance(r["proof"], dict) else r["proof"])
    final = proof[2]
    chq = tr.getField()
    t2 = Transcript(); t2.put(chq)
    queries = t2.getPermutations(4, 8)
    # layer proofs: proof[s] for s>=1 opens tree s-1's successor; mirror fri.proofQueries' layout (fri.js:83-105)
    layer = [{"root": proof[0]["root"]}, {"root": proof[1]["root"]}]
    qs = list(queries)
    opened = []
    for s in (1, 2):
        qs = [x % (1 << struct["steps"][s]["nBits"]) for x in qs]
        opened.append([MH.getGroupProof(trees[s - 1], x) for x in qs])
    vproof = [{"polQueries": [[[int(v) for v in pol[x]], None] for x in queries]},
              {"root": proof[0]["root"], "polQueries": opened[0]},
              {"root": proof[1]["root"], "polQueries": opened[1]}, final]

    def check0(query, idx):        # layer-0 consistency is the STARK's job; FRI starts from the committed first layer
        return None

    # verify the two folding links + final degree with the reference's verifier logic, starting at step 1
    vfri = FRI({"nBits": 7, "nBitsExt": 8, "nQueries": 4, "steps": struct["steps"]}, MH)
    polBits = 8
    for i, x in enumerate(queries):
        x1 = x % 32
        v, mp = opened[0][i]
        assert MH.verifyGroupProof(proof[0]["root"], mp, x1, v)
        grp = [v[3 * k:3 * k + 3] for k in range(8)]
        assert grp == [[int(c) for c in pol[x1 + 32 * k]] for k in range(8)]
        ev = S.fri_verify_fold(grp, 8, 7, challenges[1], x1)
        x2 = x1 % 4
        v2, mp2 = opened[1][i]
        assert MH.verifyGroupProof(proof[1]["root"], mp2, x2, v2)
        assert v2[3 * (x1 // 4):3 * (x1 // 4) + 3] == ev
        ev2 = S.fri_verify_fold([v2[3 * k:3 * k + 3] for k in range(8)], 5, pow(7, 8, P), challenges[2], x2)
        assert final[x2] == ev2
    # final polynomial has degree < 2^(2 - 1): coefficients above maxDeg vanish (fri.js:158-171)
    c = S.intt([list(e) for e in final])
    assert all(ci == [0, 0, 0] for ci in c[3:])


# ---------------------------------------------------------------- multi-GPU (runs when the box has >= 2 GPUs)
def _mgpu_worker(rank, world, port, n_bits, blow, cols, q, mode="peer"):
    import os
    os.environ["PIL2GPU_EXCHANGE"] = mode
    import torch
    import torch.distributed as dist
    from pil2_stark_js_b200.sharded import GpuEngine, ShardedCommit
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(5)
        full = rng.integers(0, P, size=(1 << n_bits, cols), dtype=np.uint64)
        sc = ShardedCommit(GpuEngine(torch, rank), dist, rank, world)
        cg = sc.shard_cols(cols)
        slab = torch.from_numpy(np.ascontiguousarray(full[:, rank * cg:(rank + 1) * cg]).reshape(-1).view(np.int64)).cuda()
        buf = sc.buffers(cols, n_bits, n_bits + blow)
        root = sc.commit(slab, cols, n_bits, n_bits + blow, buf)
        root = sc.commit(slab, cols, n_bits, n_bits + blow, buf)          # twice: the receive buffers are reused
        E = 1 << (n_bits + blow)
        qs = [0, 1, E - 1, E // world, E // world - 1, 1234 % E]
        rows_q, sib_q = buf["tree"].open(torch.tensor(qs, dtype=torch.int64, device="cuda"))
        lw, lh = 48, 1 << (n_bits - 2)
        layer_np = np.random.default_rng(9).integers(0, P, size=lw * lh, dtype=np.uint64)
        layer = torch.from_numpy(layer_np.view(np.int64)).cuda()
        eng = sc.e
        lt, lroot = sc.commit_rows(layer, lw, lh, eng.empty(eng.nnodes(lh // world)), eng.empty(4 * world), eng.empty(max(8, eng.nnodes(world))))
        lrows, lsib = lt.open(torch.tensor([0, lh - 1, lh // 2 + 1], dtype=torch.int64, device="cuda"))
        # the rows next to the commit over the row-sharded buffer: evaluations at xi and the FRI polynomial
        from pil2_stark_js_b200.sharded import sharded_evals, sharded_fri_pol
        frng = np.random.default_rng(11)
        xi, vf1, vf2 = (frng.integers(0, P, size=3, dtype=np.uint64) for _ in range(3))
        ev_map = [("t", c, 1, o) for o in (0, 1) for c in range(0, cols, 3)] + [("t", 29, 3, 1), ("t", 33, 3, 0)]
        trees = {"t": buf["tree"]}
        sev = sharded_evals(eng, dist, rank, world, trees, [(n_, c, d, o) for n_, c, d, o in ev_map], xi, [0, 1], n_bits, n_bits + blow)
        sf = sharded_fri_pol(eng, dist, rank, world, trees, [(n_, c, d, (0, 1)[o]) for n_, c, d, o in ev_map], sev, xi, [0, 1], vf1, vf2, n_bits,
                             n_bits + blow)
        torch.cuda.synchronize()
        u = lambda t: t.cpu().numpy().view(np.uint64).copy()
        q.put((rank, u(root), u(buf["nodes"]), u(buf["top"]), sc.exchange_kind(buf),
               (qs, u(rows_q), u(sib_q), layer_np, u(lroot), u(lrows), u(lsib)), (xi, vf1, vf2, ev_map, sev, u(sf))))
        dist.barrier()
        sc.release(buf)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_sharded_commit_two_gpus(mode):
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from pil2_stark_js_b200.sharded import assemble_nodes
    world, n_bits, blow, cols = 2, 12, 1, 64
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx_mp = mp.get_context("spawn")
    q = ctx_mp.Queue()
    procs = [ctx_mp.Process(target=_mgpu_worker, args=(r, world, port, n_bits, blow, cols, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    full = rng.integers(0, P, size=(1 << n_bits, cols), dtype=np.uint64)
    ext = C.lde(full.reshape(-1), cols, n_bits, n_bits + blow)
    nodes = C.merkelize(ext, cols, 1 << (n_bits + blow))
    for _, root, _, _, kind, extra, frows in res:
        xi, vf1, vf2, ev_map, sev, sf = frows                        # sharded evaluations + FRI polynomial == single-process oracle
        levs = [C.lev(xi, o, n_bits) for o in (0, 1)]
        want_ev = C.evals({"t": (ext, cols)}, ev_map, levs, n_bits, blow)
        assert np.array_equal(sev, want_ev)
        want_f = C.fri_polynomial({"t": (ext, cols)}, [(n_, c, d, (0, 1)[o]) for n_, c, d, o in ev_map], want_ev, [0, 1],
                                  C.x_div_x_sub_xi(xi, [0, 1], n_bits, n_bits + blow), vf1, vf2, n_bits + blow)
        assert np.array_equal(sf.reshape(-1, 3), want_f)
        assert np.array_equal(root, nodes[-4:])
        assert ("peer stores" in kind) == (mode == "peer"), f"exchange used: {kind}"
        qs, rows_q, sib_q, layer, lroot, lrows, lsib = extra          # sharded proofQueries + a sharded layer tree
        for k, qi in enumerate(qs):
            r, sb = C.group_proof(ext, nodes, cols, 1 << (n_bits + blow), qi)
            assert np.array_equal(rows_q[k], r) and np.array_equal(sib_q[k].reshape(-1), np.asarray(sb, dtype=np.uint64).reshape(-1))
        lw, lh = 48, 1 << (n_bits - 2)
        lnodes = C.merkelize(layer, lw, lh)
        assert np.array_equal(lroot, lnodes[-4:])
        for k, qi in enumerate([0, lh - 1, lh // 2 + 1]):
            r, sb = C.group_proof(layer, lnodes, lw, lh, qi)
            assert np.array_equal(lrows[k], r) and np.array_equal(lsib[k].reshape(-1), np.asarray(sb, dtype=np.uint64).reshape(-1))
    stitched = assemble_nodes([r[2] for r in res], res[0][3], (1 << (n_bits + blow)) // world, world, C.merkle_nnodes)
    assert np.array_equal(stitched, nodes)


def test_group_proofs_dev_and_rows_only_fold(ctx):
    """pil2gpu_tree_group_proofs_dev (device indices, "not mine" slots zero-filled) and pil2gpu_fri_fold_dev with
    nodes_out == NULL (rows only): the single-GPU pieces of the sharded query / FRI-layer path."""
    import ctypes
    from pil2_stark_js_b200._lib import vp, check
    L = ctx._L
    w, h = 24, 512
    elems = rnd_field(4, w * h)
    tree = ctx.tree_from_host(elems, w, h)
    nodes = C.merkelize(elems, w, h)
    idx = np.array([3, 0xFFFFFFFFFFFFFFFF, 511, 700], dtype=np.uint64)
    d_idx, d_rows, d_sib = ctx.upload(idx), ctx.alloc(4 * w), ctx.alloc(4 * 9 * 4)
    check(L.pil2gpu_tree_group_proofs_dev(ctx.handle, tree._h, d_idx.ptr, 4, d_rows.ptr, d_sib.ptr))
    rows, sib = d_rows.download().reshape(4, w), d_sib.download().reshape(4, 9, 4)
    for k, i in enumerate(idx):
        if i >= h:
            assert not rows[k].any() and not sib[k].any()
        else:
            r, sb = C.group_proof(elems, nodes, w, h, int(i))
            assert np.array_equal(rows[k], r) and np.array_equal(sib[k].reshape(-1), np.asarray(sb, dtype=np.uint64).reshape(-1))
    # rows-only fold: step 0 of a chain 2^10 -> 2^6
    pol = rnd_field(8, 3 << 10)
    ch = np.array([3, 5, 7], dtype=np.uint64)
    d_pol, d_rows2 = ctx.upload(pol), ctx.alloc(3 << 10)
    check(L.pil2gpu_fri_fold_dev(ctx.handle, d_pol.ptr, 10, 10, 6, 10, vp(ch.ctypes.data), 0, d_pol.ptr, d_rows2.ptr, None))
    want = np.ascontiguousarray(pol.reshape(16, 64, 3).transpose(1, 0, 2)).reshape(-1)
    assert np.array_equal(d_rows2.download(), want)
    assert np.array_equal(d_pol.download(), pol)
    for b in (d_idx, d_rows, d_sib, d_pol, d_rows2):
        b.free()


def test_paged_bigbuffer_entry_points(ctx):
    """pilcom BigBuffer = a list of BigUint64Array pages: the _paged twins take ragged page lists (an empty page included) and
    must give the single-buffer result; page lists that do not add up are rejected."""
    import ctypes
    import pil2_stark_js_b200 as m
    from pil2_stark_js_b200._lib import check
    L = ctx._L
    n_bits, ext_bits, cols = 9, 10, 12

    def pages(arr, cuts):
        parts = np.split(arr, cuts)
        ptrs = (ctypes.c_void_p * len(parts))(*[p.ctypes.data if p.size else 0 for p in parts])
        words = np.array([p.size for p in parts], dtype=np.uint64)
        return parts, ptrs, words

    src = rnd_field(61, cols << n_bits)
    dst = np.zeros(cols << ext_bits, dtype=np.uint64)
    sp, sptr, sw = pages(src, [7, 7, 1000, 4097])                                 # ragged, with an empty page
    dp, dptr, dw = pages(dst, [5000, 5001])
    check(L.pil2gpu_lde_paged(ctx.handle, sptr, sw.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(sp), dptr,
                              dw.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(dp), cols, n_bits, ext_bits))
    want = C.lde(src, cols, n_bits, ext_bits)
    assert np.array_equal(np.concatenate(dp), want)
    nodes = np.empty(ctx.merkle_nnodes(1 << ext_bits), dtype=np.uint64)
    ep, eptr, ew = pages(want, [1, 12 * 3 + 5, 9000])                             # page boundaries inside rows
    check(L.pil2gpu_merkelize_paged(ctx.handle, eptr, ew.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(ep), cols, 1 << ext_bits, 0,
                                    nodes.ctypes.data))
    assert np.array_equal(nodes, C.merkelize(want, cols, 1 << ext_bits))
    short = np.array([p.size for p in ep[:-1]], dtype=np.uint64)                  # one page missing
    with pytest.raises(m.Pil2GpuError):
        check(L.pil2gpu_merkelize_paged(ctx.handle, eptr, short.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(ep) - 1, cols, 1 << ext_bits, 0,
                                        nodes.ctypes.data))
    long_w = sw.copy()
    long_w[0] += 1                                                                # pages hold one word too many
    with pytest.raises(m.Pil2GpuError):
        check(L.pil2gpu_lde_paged(ctx.handle, sptr, long_w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(sp), dptr,
                                  dw.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(dp), cols, n_bits, ext_bits))


def test_c_abi_error_behaviour(ctx):
    """Every failure is a negative return code with a message, never a crash: null pointers, overlapping buffers, sizes beyond
    the field's 2-adicity, unsupported blow-ups, empty trees -- and the context keeps working afterwards."""
    import ctypes
    from pil2_stark_js_b200._lib import vp
    L = ctx._L
    a, b = ctx.alloc(64), ctx.alloc(128)
    msg = lambda: L.pil2gpu_last_error().decode()
    assert L.pil2gpu_ntt_dev(ctx.handle, None, b.ptr, 2, 3, 0) == -1 and "null" in msg()
    assert L.pil2gpu_ntt_dev(ctx.handle, a.ptr, a.ptr, 2, 3, 0) == -1 and "overlap" in msg()
    assert L.pil2gpu_ntt_dev(ctx.handle, a.ptr, b.ptr, 2, 33, 0) == -1 and "2-adicity" in msg()
    assert L.pil2gpu_ntt_dev(ctx.handle, a.ptr, b.ptr, 0, 3, 0) == -1
    assert L.pil2gpu_lde_dev(ctx.handle, a.ptr, b.ptr, 2, 4, 3) == -1 and "nBitsExt" in msg()
    assert L.pil2gpu_lde_dev(ctx.handle, a.ptr, b.ptr, 1, 2, 11) == -5 and "not supported" in msg()
    assert L.pil2gpu_merkelize_dev(ctx.handle, a.ptr, 4, 0, 0, b.ptr) == -1
    assert L.pil2gpu_merkelize_dev(ctx.handle, a.ptr, 4, 4, 0, None) == -1
    assert L.pil2gpu_fri_fold_dev(ctx.handle, a.ptr, 3, 4, -1, 4, vp(np.zeros(3, dtype=np.uint64).ctypes.data), 0, b.ptr, None, None) == -1
    assert L.pil2gpu_fri_fold_range_dev(ctx.handle, a.ptr, 0, 10, 3, -1, 10, vp(np.zeros(3, dtype=np.uint64).ctypes.data), 0, 0, b.ptr, None) == -5
    assert L.pil2gpu_compute_q_dev(ctx.handle, a.ptr, 0, 1, 2, 3, b.ptr) == -1
    t = vp()
    assert L.pil2gpu_tree_alloc(ctx.handle, 4, 0, ctypes.byref(t)) == -1
    assert L.pil2gpu_create(99, None, ctypes.byref(t)) == -1 and "out of range" in msg()
    assert L.pil2gpu_merkle_nnodes(0) == 0 and L.pil2gpu_merkle_depth(0) == 0
    # still healthy
    x = rnd_field(1, 16)
    y = np.empty(16, dtype=np.uint64)
    ctx.ntt(x, 2, 3, y)
    assert np.array_equal(y, C.ntt(x, 2, 3))
    a.free(); b.free()


# ---------------------------------------------------------------- full-size properties (sizes the oracle cannot sweep)
def _fadd(a, b):
    """elementwise (a + b) mod p on canonical uint64 arrays"""
    s = a.astype(object) + b.astype(object)
    return np.array([int(x) % P for x in s], dtype=np.uint64) if a.size < 1 << 12 else _fadd_np(a, b)


def _fadd_np(a, b):
    with np.errstate(over="ignore"):
        s = a + b
        wrapped = s < a                                   # 2^64 = 2^32 - 1 (mod p)
        s = np.where(wrapped, s + np.uint64(0xFFFFFFFF), s)
        return np.where(s >= np.uint64(P), s - np.uint64(P), s)


def test_ntt_roundtrip_and_linearity_large(ctx):
    """2^20 rows x 8 cols (three passes 7/7/6): INTT(NTT(x)) == x and NTT(a + b) == NTT(a) + NTT(b)."""
    bits, npols = 20, 8
    a, b = rnd_field(51, npols << bits), rnd_field(52, npols << bits)
    fa, fb, fs, back = (np.empty_like(a) for _ in range(4))
    ctx.ntt(a, npols, bits, fa)
    ctx.ntt(b, npols, bits, fb)
    ctx.ntt(_fadd_np(a, b), npols, bits, fs)
    assert np.array_equal(fs, _fadd_np(fa, fb))
    ctx.ntt(fa, npols, bits, back, inverse=True)
    assert np.array_equal(back, a)
    assert int(fa.max()) < P and int(back.max()) < P      # canonical outputs


def test_lde_decimation_and_linearity_large(ctx):
    """2^19 rows x 16 cols: every 2nd / 4th row of the blowup-2 / blowup-4 extension is the blowup-1 extension (the same
    coset 7<w_N>), and the extension is linear.  Exercises the fused middle kernel with B = 1, 2, 4 on three-pass plans."""
    bits, npols = 19, 16
    a, b = rnd_field(61, npols << bits), rnd_field(62, npols << bits)
    e1 = np.empty(npols << bits, dtype=np.uint64)
    e2 = np.empty(npols << (bits + 1), dtype=np.uint64)
    e4 = np.empty(npols << (bits + 2), dtype=np.uint64)
    ctx.lde(a, npols, bits, e1, bits)
    ctx.lde(a, npols, bits, e2, bits + 1)
    ctx.lde(a, npols, bits, e4, bits + 2)
    assert np.array_equal(e2.reshape(-1, npols)[0::2].reshape(-1), e1)
    assert np.array_equal(e4.reshape(-1, npols)[0::4].reshape(-1), e1)
    assert np.array_equal(e4.reshape(-1, npols)[0::2].reshape(-1), e2)
    eb, es = np.empty_like(e2), np.empty_like(e2)
    ctx.lde(b, npols, bits, eb, bits + 1)
    ctx.lde(_fadd_np(a, b), npols, bits, es, bits + 1)
    assert np.array_equal(es, _fadd_np(e2, eb))


def test_cfg2_commit_root_vs_oracle(ctx):
    """BASELINE config 2 at full size (2^20 rows x 64 cols, blowup 2): device-resident commit root, the host-buffer
    pipelined commit and the C oracle agree; random openings verify against the root."""
    bits, npols = 20, 64
    src = rnd_field(71, npols << bits)
    tree, root = ctx.commit(src, npols, bits, bits + 1)
    ext = C.lde(src, npols, bits, bits + 1)
    nodes = C.merkelize(ext, npols, 1 << (bits + 1))
    assert np.array_equal(root, nodes[-4:])
    _, nodes_pipe, root_pipe = ctx.extend_and_merkelize(src, npols, bits, bits + 1, want_dst=False)
    assert np.array_equal(nodes_pipe, nodes) and np.array_equal(root_pipe, root)
    import pil2_stark_js_b200 as m
    MH = m.buildMerkleHash(False, ctx)
    idxs = [0, 1, (1 << (bits + 1)) - 1, 123456, 1 << bits]
    rows, sibs = tree.group_proofs(idxs)
    for q, idx in enumerate(idxs):
        assert np.array_equal(rows[q], ext[idx * npols:(idx + 1) * npols])
        assert MH.verifyGroupProof([int(x) for x in root], [[int(x) for x in s] for s in sibs[q]], idx, [int(x) for x in rows[q]])
    tree.free()
