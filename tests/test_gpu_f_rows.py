"""GPU parity for the rows built after the core commit path: the LDE with fused row scatter, the quotient
commit (8 f1), the evaluations at xi and the FRI denominators (8 f3) -- all through the C ABI, bit-exact against the oracle."""
import ctypes
import types

import numpy as np
import pytest

from oracle import gl_oracle as C
from oracle import gl_spec as S

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


# ---------------------------------------------------------------- LDE with the row scatter fused into its last pass
@pytest.mark.parametrize("n_bits,blow,cols,world", [(6, 1, 32, 2), (10, 1, 64, 4), (12, 2, 32, 2), (13, 1, 128, 8), (4, 1, 16, 2),
                                                    (11, 1, 24, 1)])
def test_lde_scatter_virtual_ranks(ctx, n_bits, blow, cols, world):
    """All `world` ranks played by one GPU, one after the other: every rank's slab is extended with
    pil2gpu_lde_scatter_dev into the `world` receive buffers; each receive buffer must then hold that rank's rows as
    column tiles (what the all-to-all used to deliver) and hash to the oracle's subtree."""
    from pil2_stark_js_b200._lib import vp, check
    L = ctx._L
    ext = n_bits + blow
    E, cg = 1 << ext, cols // world
    rows_local = E // world
    full = rnd_field(77 + n_bits + cols, cols << n_bits).reshape(1 << n_bits, cols)
    want = C.lde(full.reshape(-1), cols, n_bits, ext).reshape(E, cols)
    recv = [ctx.alloc(cg * E) for _ in range(world)]
    peers = (ctypes.c_void_p * world)(*[r.ptr.value for r in recv])
    dst = ctx.alloc(cg * E)
    for g in range(world):
        slab = ctx.upload(np.ascontiguousarray(full[:, g * cg:(g + 1) * cg]))
        check(L.pil2gpu_lde_scatter_dev(ctx.handle, slab.ptr, dst.ptr, cg, n_bits, ext, peers, world, g, 0, 0))
        ctx.sync()
        slab.free()
    for h in range(world):
        tiles = recv[h].download().reshape(world, rows_local, cg)
        got = np.ascontiguousarray(tiles.transpose(1, 0, 2)).reshape(rows_local, cols)
        assert np.array_equal(got, want[h * rows_local:(h + 1) * rows_local]), f"rank {h} rows differ"
        if cg % 8 == 0 or world == 1:
            nodes = ctx.alloc(ctx.merkle_nnodes(rows_local))
            check(L.pil2gpu_merkelize_tiled_dev(ctx.handle, recv[h].ptr, world, cg, rows_local * cg, rows_local, 0, nodes.ptr))
            exp = C.merkelize(np.ascontiguousarray(want[h * rows_local:(h + 1) * rows_local]).reshape(-1), cols, rows_local)
            assert np.array_equal(nodes.download(), exp)
            nodes.free()
    for r in recv:
        r.free()
    dst.free()


def test_lde_scatter_argument_checks(ctx):
    from pil2_stark_js_b200 import Pil2GpuError
    from pil2_stark_js_b200._lib import check
    L = ctx._L
    a, b = ctx.alloc(64), ctx.alloc(128)
    peers3 = (ctypes.c_void_p * 3)(b.ptr.value, b.ptr.value, b.ptr.value)
    with pytest.raises(Pil2GpuError):      # 3 ranks: not a power of two
        check(L.pil2gpu_lde_scatter_dev(ctx.handle, a.ptr, b.ptr, 8, 3, 4, peers3, 3, 0, 0, 0))
    peers2 = (ctypes.c_void_p * 2)(b.ptr.value, None)
    with pytest.raises(Pil2GpuError):      # null peer
        check(L.pil2gpu_lde_scatter_dev(ctx.handle, a.ptr, b.ptr, 8, 3, 4, peers2, 2, 0, 0, 0))
    with pytest.raises(Pil2GpuError):      # slab columns outside the tile
        check(L.pil2gpu_lde_scatter_dev(ctx.handle, a.ptr, b.ptr, 8, 3, 4, (ctypes.c_void_p * 1)(b.ptr.value), 1, 0, 8, 4))
    a.free(); b.free()


# ---------------------------------------------------------------- quotient commit (computeQStark)
@pytest.mark.parametrize("n_bits,ext_bits,q_dim,q_deg,split", [(4, 5, 3, 2, False), (6, 8, 3, 3, False), (10, 11, 3, 2, False),
                                                               (10, 12, 3, 4, True), (12, 13, 1, 2, False), (9, 10, 2, 1, False),
                                                               (13, 14, 3, 2, False), (5, 5, 3, 1, False)])
def test_compute_q_vs_oracle(ctx, n_bits, ext_bits, q_dim, q_deg, split):
    q = rnd_field(31 * n_bits + q_deg, q_dim << ext_bits)
    ext, nodes, root = ctx.compute_q(q, q_dim, q_deg, n_bits, ext_bits, split)
    want = C.compute_q(q, q_dim, q_deg, n_bits, ext_bits)
    assert np.array_equal(ext, want)
    want_nodes = C.merkelize(want, q_dim * q_deg, 1 << ext_bits, split=split)
    assert np.array_equal(nodes, want_nodes)
    assert np.array_equal(root, want_nodes[-4:])


def test_compute_q_rejects_oversized_degree(ctx):
    from pil2_stark_js_b200 import Pil2GpuError
    q = rnd_field(1, 3 << 6)
    with pytest.raises(Pil2GpuError):
        ctx.compute_q(q, 3, 4, 5, 6)          # qDeg 4 > blowup 2


def test_compute_q_recombines_large(ctx):
    """2^18 x 3, blowup 4, qDeg 3 (beyond what the oracle sweeps quickly): stark_verify.js:140-147 identity at sampled
    rows -- sum_p x^(N p) cmQ[p] == q_ext, for a quotient of degree < qDeg * N built by LDE of random coefficients."""
    n_bits, ext_bits, q_dim, q_deg = 18, 20, 3, 3
    n, ne = 1 << n_bits, 1 << ext_bits
    # q_ext = evaluations on 7<w_E> of a random polynomial of degree < q_deg * n: NTT_E of coefficients c_j * 7^j
    coef = np.zeros((ne, q_dim), dtype=np.uint64)
    coef[:q_deg * n] = rnd_field(5, q_deg * n * q_dim).reshape(-1, q_dim)
    pw = np.empty(ne, dtype=object)
    acc = 1
    for j in range(ne):
        pw[j] = acc
        acc = acc * 7 % P
    scaled = np.array([[int(coef[j, k]) * pw[j] % P for k in range(q_dim)] for j in range(q_deg * n)], dtype=np.uint64)
    coef[:q_deg * n] = scaled
    q_ext = np.empty_like(coef).reshape(-1)
    ctx.ntt(np.ascontiguousarray(coef).reshape(-1), q_dim, ext_bits, q_ext)
    ext, _, _ = ctx.compute_q(q_ext, q_dim, q_deg, n_bits, ext_bits, want_nodes=False)
    ext = ext.reshape(ne, q_dim * q_deg)
    q_ext = q_ext.reshape(ne, q_dim)
    w = S.root_of_unity(ext_bits)
    for j in [0, 1, 12345, ne // 2 + 3, ne - 1]:
        x = 7 * pow(w, j, P) % P
        xn = pow(x, n, P)
        for k in range(q_dim):
            tot, xa = 0, 1
            for p in range(q_deg):
                tot = (tot + xa * int(ext[j, p * q_dim + k])) % P
                xa = xa * xn % P
            assert tot == int(q_ext[j, k])


# ---------------------------------------------------------------- evaluations at xi, x / (x - xi)
@pytest.mark.parametrize("n_bits,opening", [(0, 0), (1, 1), (4, 0), (7, 1), (10, -1), (13, 2)])
def test_compute_lev_vs_oracle(ctx, n_bits, opening):
    xi = rnd_field(3 + n_bits, 3)
    lev = ctx.compute_levs(xi, [opening], n_bits)
    got = lev.download().reshape(-1, 3)
    assert np.array_equal(got, C.lev(xi, opening, n_bits))
    lev.free()


@pytest.mark.parametrize("mode", ["mma", "mma1", "mma2", "scalar"])
@pytest.mark.parametrize("n_bits,extend_bits,size,n_evals,openings", [
    (6, 1, 9, 5, [0, 1, -1]), (10, 1, 37, 40, [0, 1, -1]), (12, 2, 15, 300, [0, 1, -1]), (13, 1, 128, 100, [0, 1]), (3, 0, 4, 2, [0, 1, -1]),
    (11, 1, 64, 64, [0]), (10, 0, 33, 50, [0, 1, -1, 2]), (15, 1, 8, 30, [0, 1]), (12, 1, 2, 4, [0, 1, 2, 3, 4]),
    (10, 1, 37, 40, [0, 1]), (12, 0, 301, 64, [0, 1]), (11, 2, 300, 50, [0]), (13, 1, 256, 90, [0, -1])])
def test_compute_evals_vs_oracle(ctx, monkeypatch, mode, n_bits, extend_bits, size, n_evals, openings):
    """Evaluation sums through the tensor-core byte-limb GEMM (n_bits >= 10, <= 4 openings) and through the per-evaluation
    kernel (PIL2GPU_EVALS=scalar; also what small inputs and > 4 openings use): both bit-exact against the oracle."""
    if mode != "mma":                                        # mma1 / mma2: force the first / second tensor-core formulation whatever the width (mma2: <= 2 openings)
        monkeypatch.setenv("PIL2GPU_EVALS", mode)
    rng = np.random.default_rng(n_bits + size)
    xi = rnd_field(11, 3)
    ext_bits = n_bits + extend_bits
    buf = rnd_field(13 + size, size << ext_bits)
    if size >= 3:
        buf[:3] = [P - 1, P - 1, 0xFFFFFFFF]                 # extreme limbs in the first row
    ev = []
    for _ in range(n_evals):
        dim = 3 if (size >= 3 and rng.integers(0, 3) == 0) else 1
        ev.append((int(rng.integers(0, size - dim + 1)), dim, int(rng.integers(0, len(openings)))))
    levs = ctx.compute_levs(xi, openings, n_bits)
    dbuf = ctx.upload(buf)
    got = ctx.compute_evals(dbuf, size, n_bits, ext_bits, ev, levs, len(openings))
    levs_o = [C.lev(xi, o, n_bits) for o in openings]
    want = C.evals({"b": (buf, size)}, [("b", o, d, l) for o, d, l in ev], levs_o, n_bits, extend_bits)
    assert np.array_equal(got, want)
    dbuf.free(); levs.free()


def test_compute_evals_mma_saturated_limbs(ctx):
    """All-0xFF bytes in both factors over a full 2^15-row chunk: the s32 limb accumulators reach 2^15 * 255^2 = 2130706432 < 2^31."""
    n_bits, size = 15, 32
    n = 1 << n_bits
    buf = np.full(size * n, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)          # non-canonical on purpose: interpreted mod p
    lev = np.full(3 * n, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    dlev, dbuf = ctx.upload(lev), ctx.upload(buf)
    got = ctx.compute_evals(dbuf, size, n_bits, n_bits, [(0, 1, 0), (7, 3, 0)], dlev, 1)
    v = (2**64 - 1) % P
    s = n * v * v % P
    assert [int(x) for x in got[0]] == [s, s, s]
    r = [s, s, s]
    x1 = [r[2], (r[0] + r[2]) % P, r[1]]
    x2t = [r[2], (r[0] + r[2]) % P, r[1]]
    x2 = [x2t[2], (x2t[0] + x2t[2]) % P, x2t[1]]
    assert [int(x) for x in got[1]] == [(r[i] + x1[i] + x2[i]) % P for i in range(3)]
    dlev.free(); dbuf.free()


def test_compute_evals_range_checks(ctx):
    from pil2_stark_js_b200 import Pil2GpuError
    xi = rnd_field(1, 3)
    levs = ctx.compute_levs(xi, [0], 4)
    dbuf = ctx.upload(rnd_field(2, 5 << 5))
    with pytest.raises(Pil2GpuError):
        ctx.compute_evals(dbuf, 5, 4, 5, [(4, 3, 0)], levs, 1)       # columns 4..6 of a 5-word row
    with pytest.raises(Pil2GpuError):
        ctx.compute_evals(dbuf, 5, 4, 5, [(0, 1, 1)], levs, 1)       # opening index 1 of 1
    with pytest.raises(Pil2GpuError):
        ctx.compute_evals(dbuf, 5, 4, 5, [(0, 2, 0)], levs, 1)       # dim 2
    dbuf.free(); levs.free()


@pytest.mark.parametrize("n_bits,ext_bits,openings", [(3, 5, [0, 1, -2]), (9, 10, [0, 1]), (12, 14, [0, 1, -1, 5]), (0, 0, [0]), (2, 2, [0, 1])])
def test_x_div_x_sub_xi_vs_oracle(ctx, n_bits, ext_bits, openings):
    xi = rnd_field(19 + ext_bits, 3)
    got = ctx.x_div_x_sub_xi(xi, openings, n_bits, ext_bits)
    assert np.array_equal(got, C.x_div_x_sub_xi(xi, openings, n_bits, ext_bits))


# ---------------------------------------------------------------- the prover-side callers (stark_gen_helpers mirror)
def test_stark_gen_helpers_stage_flow(ctx):
    """extendAndMerkelize -> computeQStark -> computeEvalsStark -> xDivXSubXi with the reference's ctx field names, checked
    row by row against the oracle's restatement of the same functions."""
    import pil2_stark_js_b200 as m
    from pil2_stark_js_b200 import stark_gen_helpers as H
    n_bits, ext_bits = 8, 9
    N, extN = 1 << n_bits, 1 << ext_bits
    pil = {"nStages": 1, "qDim": 3, "qDeg": 2, "mapSectionsN": {"cm1": 10, "cm2": 6}, "nConstants": 4,
           "openingPoints": [0, 1],
           "cmPolsMap": [{"stage": 1, "stagePos": 0, "dim": 1}, {"stage": 1, "stagePos": 3, "dim": 3}, {"stage": 2, "stagePos": 0, "dim": 3},
                         {"stage": 2, "stagePos": 3, "dim": 3}],
           "evMap": [{"type": "cm", "id": 0, "prime": 0}, {"type": "cm", "id": 0, "prime": 1}, {"type": "cm", "id": 1, "prime": 0},
                     {"type": "const", "id": 2, "prime": 1}, {"type": "cm", "id": 2, "prime": 0}, {"type": "cm", "id": 3, "prime": 0}],
           "starkStruct": {"nBits": n_bits, "nBitsExt": ext_bits, "nQueries": 8, "steps": [{"nBits": ext_bits}, {"nBits": 5}]}}
    c = types.SimpleNamespace(pilInfo=pil, nBits=n_bits, nBitsExt=ext_bits, N=N, extN=extN, extendBits=1, trees={}, gpu=ctx,
                              MH=m.buildMerkleHash(False, ctx))
    c.cm1_n = rnd_field(1, 10 * N)
    c.cm1_ext = np.zeros(10 * extN, dtype=np.uint64)
    c.const_ext = C.lde(rnd_field(2, 4 * N), 4, n_bits, ext_bits)
    c.q_ext = rnd_field(3, 3 * extN)
    c.cm2_ext = None
    root1 = H.extendAndMerkelize(1, c)
    want1 = C.lde(c.cm1_n, 10, n_bits, ext_bits)
    assert np.array_equal(c.cm1_ext, want1)
    assert root1[0] == [int(x) for x in C.merkelize(want1, 10, extN)[-4:]]
    root2 = H.computeQStark(c)
    want2 = C.compute_q(c.q_ext, 3, 2, n_bits, ext_bits)
    assert np.array_equal(c.cm2_ext, want2)
    assert root2[0] == [int(x) for x in C.merkelize(want2, 6, extN)[-4:]]
    assert c.MH.root(c.trees[2]) == root2[0]
    c.challenges = {2: [[int(x) for x in rnd_field(4, 3)]]}
    evals = H.computeEvalsStark(c)
    levs = [C.lev(np.array(c.challenges[2][0], dtype=np.uint64), o, n_bits) for o in (0, 1)]
    want_ev = C.evals({"cm1": (want1, 10), "cm2": (want2, 6), "const": (c.const_ext, 4)},
                      [("cm1", 0, 1, 0), ("cm1", 0, 1, 1), ("cm1", 3, 3, 0), ("const", 2, 1, 1), ("cm2", 0, 3, 0), ("cm2", 3, 3, 0)],
                      levs, n_bits, 1)
    assert evals == [[int(x) for x in r] for r in want_ev]
    xd = H.computeXDivXSubXi(c)
    assert np.array_equal(xd.reshape(-1, 2, 3), C.x_div_x_sub_xi(np.array(c.challenges[2][0], dtype=np.uint64), [0, 1], n_bits, ext_bits))
    q = H.getPermutationsStark(c, [1, 2, 3])
    t = S.Transcript()
    t.put([1, 2, 3])
    assert q == t.get_permutations(8, ext_bits)
    # FRI polynomial from the honest evaluations, then the FRI chain on it: the final polynomial passes the degree check of
    # fri.js:158-171 only because f_ext really is of degree < N -- commit, evaluations, xDivXSubXi and friExp all have to agree
    c.challenges[4] = [[int(x) for x in rnd_field(5, 3)], [int(x) for x in rnd_field(6, 3)]]
    f = H.computeFRIPol(c)
    ev_map_o = [("cm1", 0, 1, 0), ("cm1", 0, 1, 1), ("cm1", 3, 3, 0), ("const", 2, 1, 1), ("cm2", 0, 3, 0), ("cm2", 3, 3, 0)]
    want_f = C.fri_polynomial({"cm1": (want1, 10), "cm2": (want2, 6), "const": (c.const_ext, 4)}, ev_map_o, np.array(evals, dtype=np.uint64), [0, 1],
                              xd, np.array(c.challenges[4][0], dtype=np.uint64), np.array(c.challenges[4][1], dtype=np.uint64), ext_bits)
    assert np.array_equal(f, want_f)
    # the same stages with everything kept in HBM (ctx.device_resident): identical roots, evaluations, f_ext and openings
    d = types.SimpleNamespace(pilInfo=pil, nBits=n_bits, nBitsExt=ext_bits, N=N, extN=extN, extendBits=1, trees={}, gpu=ctx, MH=c.MH,
                              device_resident=True, cm1_n=c.cm1_n, q_ext=c.q_ext, challenges=c.challenges, dev_buffers=None)
    assert H.extendAndMerkelize(1, d) == root1 and H.computeQStark(d) == root2
    const_dev = ctx.upload(c.const_ext)
    d.dev_buffers["const_ext"] = const_dev
    assert isinstance(d.trees[1], m.DeviceTree) and set(d.dev_buffers) == {"cm1_ext", "cm2_ext", "const_ext"}
    assert H.computeEvalsStark(d) == evals
    assert np.array_equal(H.computeFRIPol(d), f)
    for idx in (0, 5, extN - 1):
        assert c.MH.getGroupProof(d.trees[1], idx) == c.MH.getGroupProof(c.trees[1], idx)
        assert c.MH.getGroupProof(d.trees[2], idx) == c.MH.getGroupProof(c.trees[2], idx)
    const_dev.free(); d.trees[1].free(); d.trees[2].free()
    c.fri = m.FRI(pil["starkStruct"], c.MH)
    c.friPol, c.friProof, c.friTrees = {0: f}, {0: {}}, {}
    H.computeFRIFolding(0, c, [0, 0, 0])
    last = H.computeFRIFolding(1, c, [int(x) for x in rnd_field(7, 3)])
    coeffs = S.intt([[int(v) for v in e] for e in np.asarray(last).reshape(-1, 3)])
    max_deg = 1 << (5 - (ext_bits - n_bits))
    assert all(cf == [0, 0, 0] for cf in coeffs[max_deg:]) and any(cf != [0, 0, 0] for cf in coeffs[:max_deg])


@pytest.mark.parametrize("n_bits,blow,cols,world", [(10, 1, 128, 2), (12, 1, 256, 4), (9, 2, 64, 2), (8, 1, 48, 2)])
def test_lde_scatter_from_host_virtual_ranks(ctx, n_bits, blow, cols, world):
    """pil2gpu_lde_scatter: every virtual rank pulls its column slab straight out of the full-width HOST buffer (row pitch =
    all columns) in sub-slabs; the receive buffers must equal the oracle's rows, tile by tile."""
    from pil2_stark_js_b200._lib import check
    L = ctx._L
    ext = n_bits + blow
    E, cg = 1 << ext, cols // world
    rows_local = E // world
    full = rnd_field(123 + cols, cols << n_bits).reshape(1 << n_bits, cols)
    want = C.lde(full.reshape(-1), cols, n_bits, ext).reshape(E, cols)
    recv = [ctx.alloc(cg * E) for _ in range(world)]
    peers = (ctypes.c_void_p * world)(*[r.ptr.value for r in recv])
    for g in range(world):
        check(L.pil2gpu_lde_scatter(ctx.handle, ctypes.c_void_p(full.ctypes.data + 8 * g * cg), cols, cg, n_bits, ext, peers, world, g))
    ctx.sync()
    for h in range(world):
        tiles = recv[h].download().reshape(world, rows_local, cg)
        got = np.ascontiguousarray(tiles.transpose(1, 0, 2)).reshape(rows_local, cols)
        assert np.array_equal(got, want[h * rows_local:(h + 1) * rows_local]), f"rank {h} rows differ"
    for r in recv:
        r.free()


# ---------------------------------------------------------------- const tree files -> device-resident trees
def test_const_tree_file_to_device(ctx, tmp_path):
    """buildConstTree (stark_buildConstTree.js:17-33) -> writeToFile -> readFromFileToDevice / `.cnts` with tree_to_device:
    root and group proofs from the device handle equal the host tree's, and "Out of range" still throws."""
    import pil2_stark_js_b200 as m
    from pil2_stark_js_b200 import stark_consts_file as CF
    n_bits, ext_bits, n_consts = 9, 10, 9
    MH = m.buildMerkleHash(False, ctx)
    const_n = rnd_field(3, n_consts << n_bits)
    const_ext = np.empty(n_consts << ext_bits, dtype=np.uint64)
    m.interpolate(const_n, n_consts, n_bits, const_ext, ext_bits, ctx=ctx)
    tree = MH.merkelize(const_ext, n_consts, 1 << ext_bits)
    assert np.array_equal(tree["nodes"], C.merkelize(C.lde(const_n, n_consts, n_bits, ext_bits), n_consts, 1 << ext_bits))
    fn = tmp_path / "const.tree"
    MH.writeToFile(tree, fn)
    dev = MH.readFromFileToDevice(fn, chunk_words=1000)            # several chunks per array
    assert MH.root(dev) == MH.root(tree)
    for idx in (0, 1, 513, (1 << ext_bits) - 1):
        assert MH.getGroupProof(dev, idx) == MH.getGroupProof(tree, idx)
        assert MH.verifyGroupProof(MH.root(dev), MH.getGroupProof(dev, idx)[1], idx, MH.getGroupProof(dev, idx)[0])
    assert MH.getElement(dev, 7, 3) == MH.getElement(tree, 7, 3)
    with pytest.raises(m.OutOfRange):
        MH.getGroupProof(dev, 1 << ext_bits)
    w = S.root_of_unity(n_bits)
    consts = {"fixedPolsEvals": const_n, "constTree": tree, "x_n": np.array([pow(w, i, P) for i in range(1 << n_bits)], dtype=np.uint64),
              "x_ext": np.array([7 * pow(S.root_of_unity(ext_bits), i, P) % P for i in range(1 << ext_bits)], dtype=np.uint64)}
    fn2 = tmp_path / "setup.cnts"
    CF.writePilStarkConstsFile(consts, fn2)
    back = CF.readPilStarkConstsFile(fn2, ctx=ctx, tree_to_device=True)
    assert np.array_equal(back["fixedPolsEvals"], const_n) and np.array_equal(back["x_ext"], consts["x_ext"])
    assert MH.root(back["constTree"]) == MH.root(tree)
    assert MH.getGroupProof(back["constTree"], 77) == MH.getGroupProof(tree, 77)
    e, n = back["constTree"].download()
    assert np.array_equal(e, const_ext) and np.array_equal(n, tree["nodes"])
    dev.free(); back["constTree"].free()


# ---------------------------------------------------------------- FRI polynomial (friExp over the extended domain)
@pytest.mark.parametrize("n_bits,ext_bits,openings,sizes,n_terms", [
    (4, 6, [0, 1], (4, 5), 6), (8, 9, [0, 1, -1], (37, 6), 40), (10, 11, [0, 1], (256, 12), 300), (6, 8, [0], (3, 3), 4),
    (9, 10, [0, 1, -1, 2], (64, 9), 100), (11, 12, [1, -1, 0], (130, 16), 200)])
def test_fri_pol_vs_oracle(ctx, n_bits, ext_bits, openings, sizes, n_terms):
    rng = np.random.default_rng(n_bits * 31 + n_terms)
    ne = 1 << ext_bits
    bufs = {"a": (rnd_field(1 + n_terms, sizes[0] * ne), sizes[0]), "b": (rnd_field(2 + n_terms, sizes[1] * ne), sizes[1])}
    bufs["a"][0][:3] = [P - 1, 0, 0xFFFFFFFF]
    ev_map = []
    for _ in range(n_terms):
        name = "a" if rng.integers(0, 3) else "b"
        size = bufs[name][1]
        dim = 3 if (size >= 3 and rng.integers(0, 4) == 0) else 1
        ev_map.append((name, int(rng.integers(0, size - dim + 1)), dim, int(openings[rng.integers(0, len(openings))])))
    evals = rnd_field(5, 3 * n_terms).reshape(-1, 3)
    xi = rnd_field(6, 3)
    vf1, vf2 = rnd_field(7, 3), rnd_field(8, 3)
    xdiv_h = C.x_div_x_sub_xi(xi, openings, n_bits, ext_bits)
    want = C.fri_polynomial(bufs, ev_map, evals, openings, xdiv_h, vf1, vf2, ext_bits)
    dev = {k: ctx.upload(v[0]) for k, v in bufs.items()}
    xdiv = ctx.x_div_x_sub_xi(xi, openings, n_bits, ext_bits, download=False)
    assert np.array_equal(xdiv.download().reshape(xdiv_h.shape), xdiv_h)
    terms = [(dev[name], bufs[name][1], off, dim, prime) for name, off, dim, prime in ev_map]
    got = ctx.fri_pol(terms, evals, openings, xdiv, vf1, vf2, ext_bits)
    assert np.array_equal(got, want)
    for b in list(dev.values()) + [xdiv]:
        b.free()


def test_fri_pol_low_degree_large(ctx):
    """2^14 rows, blowup 4, 96 columns: honest evaluations (computed by the evaluation kernels) make f_ext a polynomial of degree
    < N -- INTT on the GPU, coefficients above N vanish after undoing the coset shift is not even needed (zero stays zero)."""
    n_bits, ext_bits, size = 14, 16, 96
    n, ne = 1 << n_bits, 1 << ext_bits
    openings = [0, 1]
    trace = rnd_field(40, size * n)
    ext = np.empty(size * ne, dtype=np.uint64)
    ctx.lde(trace, size, n_bits, ext, ext_bits)
    xi, vf1, vf2 = rnd_field(41, 3), rnd_field(42, 3), rnd_field(43, 3)
    ev_map = [(c, 1, o) for o in (0, 1) for c in range(0, size, 2)] + [(5, 3, 0)]
    dext = ctx.upload(ext)
    levs = ctx.compute_levs(xi, openings, n_bits)
    evals = ctx.compute_evals(dext, size, n_bits, ext_bits, [(c, d, openings.index(o)) for c, d, o in ev_map], levs, 2)
    xdiv = ctx.x_div_x_sub_xi(xi, openings, n_bits, ext_bits, download=False)
    f = ctx.fri_pol([(dext, size, c, d, o) for c, d, o in ev_map], evals, openings, xdiv, vf1, vf2, ext_bits)
    coef = np.empty(3 * ne, dtype=np.uint64)
    ctx.ntt(np.ascontiguousarray(f).reshape(-1), 3, ext_bits, coef, inverse=True)
    coef = coef.reshape(ne, 3)
    assert not coef[n:].any(), "f_ext is not of degree < N"
    assert coef[:n].any()
    evals[3, 1] ^= np.uint64(1)                                  # one wrong evaluation: no longer a polynomial of degree < N
    f2 = ctx.fri_pol([(dext, size, c, d, o) for c, d, o in ev_map], evals, openings, xdiv, vf1, vf2, ext_bits)
    ctx.ntt(np.ascontiguousarray(f2).reshape(-1), 3, ext_bits, coef.reshape(-1), inverse=True)
    assert coef.reshape(ne, 3)[n:].any()
    for b in (dext, levs, xdiv):
        b.free()


def test_sharded_rows_single_rank_two_tiles(ctx):
    """sharded_evals / sharded_fri_pol on one rank whose buffer is stored as two column tiles (the layout the row exchange
    delivers): the tile -> buffer mapping, without any collective."""
    import torch
    from pil2_stark_js_b200.sharded import GpuEngine, ShardedTree, sharded_evals, sharded_fri_pol
    n_bits, ext_bits, cols = 10, 11, 64
    ne, cg = 1 << ext_bits, 32
    ext = C.lde(rnd_field(3, cols << n_bits), cols, n_bits, ext_bits).reshape(ne, cols)
    tiles = np.concatenate([np.ascontiguousarray(ext[:, t * cg:(t + 1) * cg]).reshape(-1) for t in range(2)])
    eng = GpuEngine(torch, 0)
    dt = torch.from_numpy(tiles.view(np.int64)).cuda()
    tree = ShardedTree(eng, None, 0, 1, dt, 2, cg, ne, None, None, None)
    xi, vf1, vf2 = rnd_field(4, 3), rnd_field(5, 3), rnd_field(6, 3)
    ev_map = [("t", c, 1, o) for o in (0, 1) for c in range(0, cols, 5)] + [("t", 29, 3, 1), ("t", 33, 3, 0)]
    ev = sharded_evals(eng, None, 0, 1, {"t": tree}, ev_map, xi, [0, 1], n_bits, ext_bits)
    want_ev = C.evals({"t": (ext.reshape(-1), cols)}, ev_map, [C.lev(xi, o, n_bits) for o in (0, 1)], n_bits, 1)
    assert np.array_equal(ev, want_ev)
    f = sharded_fri_pol(eng, None, 0, 1, {"t": tree}, ev_map, ev, xi, [0, 1], vf1, vf2, n_bits, ext_bits)
    want_f = C.fri_polynomial({"t": (ext.reshape(-1), cols)}, ev_map, want_ev, [0, 1], C.x_div_x_sub_xi(xi, [0, 1], n_bits, ext_bits), vf1, vf2,
                              ext_bits)
    torch.cuda.synchronize()
    assert np.array_equal(f.cpu().numpy().view(np.uint64).reshape(-1, 3), want_f)
    with pytest.raises(ValueError):
        sharded_evals(eng, None, 0, 1, {"t": tree}, [("t", 31, 3, 0)], xi, [0, 1], n_bits, ext_bits)      # straddles the two tiles


# ---------------------------------------------------------------- pins against the golden proof (verifier.proof.zkin.json)
def test_golden_evals_on_gpu(ctx, golden):
    """f3 on the GPU against the reference's own vector: sm_all constant / stage-1 traces -> device-resident commit -> LEv
    vectors and evaluation sums at the transcript's xi (Transcript hashing on the GPU too) == evals[0..27] of the proof.
    Evaluation map: verifier.circom:541-672 (tests/test_oracle_f_rows.py:golden_ev_map)."""
    from oracle import sm_all
    from pil2_stark_js_b200 import Transcript
    from test_oracle_f_rows import golden_ev_map, golden_challenges
    xi, _, _, _, q = golden_challenges(golden, Transcript)
    assert q == [891, 1628, 1228, 1991, 1856, 415, 833, 296]
    ev_map = golden_ev_map()
    levs = ctx.compute_levs(xi, [0, 1], 10)
    for name, fn in (("const", sm_all.constant_trace), ("stage1", sm_all.committed_trace)):
        buff, w = fn()
        tree, root = ctx.commit(np.array(buff, dtype=np.uint64), w, 10, 11)
        assert [int(x) for x in root] == golden["roots"][name]
        idx = [i for i, e in enumerate(ev_map) if e[0] == name]
        got = ctx.compute_evals(tree.elements_ptr, w, 10, 11, [(ev_map[i][1], ev_map[i][2], ev_map[i][3]) for i in idx], levs, 2)
        assert got.tolist() == [golden["evals"][i] for i in idx], name
        tree.free()
    levs.free()


def test_golden_fri_polynomial_on_gpu(ctx, golden):
    """fri_pol + xDivXSubXi on the GPU against the reference's own vector: friExp at the 8 query rows, from the opened rows of
    all five trees, the proof's evals and the transcript challenges, equals the value the first FRI layer opens there
    (verifier.circom:501-704)."""
    from pil2_stark_js_b200 import Transcript
    from test_oracle_f_rows import golden_ev_map, golden_challenges, _golden_row_buffers
    xi, vf1, vf2, _, q = golden_challenges(golden, Transcript)
    ev_map = golden_ev_map()
    bufs = _golden_row_buffers(golden, q)
    dev = {k: ctx.upload(v[0]) for k, v in bufs.items()}
    xdiv = ctx.x_div_x_sub_xi(xi, [0, 1], 10, 11, download=False)
    terms = [(dev[name], bufs[name][1], off, dim, prime) for name, off, dim, prime in ev_map]
    f = ctx.fri_pol(terms, golden["evals"], [0, 1], xdiv, vf1, vf2, 11)
    for k, idx in enumerate(q):
        j = idx >> 7
        assert [int(x) for x in f[idx]] == golden["fri1"]["rows"][k][3 * j:3 * j + 3], (k, idx)
    # and through the host-buffer entry point a JS caller uses (computeFRIStark's arithmetic in one call)
    hterms = [(bufs[name][0], bufs[name][1], off, dim, prime) for name, off, dim, prime in ev_map]
    f2 = ctx.fri_pol_host(hterms, golden["evals"], [0, 1], xi, vf1, vf2, 10, 11)
    assert np.array_equal(f2, f)
    for b in list(dev.values()) + [xdiv]:
        b.free()
