"""N>1 path on CPU: the multi-GPU choreography of pil2_stark_js_b200/sharded.py (column slabs -> all-to-all -> row tiles
-> sub-roots -> top tree) run by 2 and 4 gloo ranks with a stand-in engine, checked against the single-process oracle.
The stand-in engine uses the oracle for the arithmetic (this is a test of the exchange/assembly logic, not of kernels)."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gl_oracle as C
from pil2_stark_js_b200.sharded import ShardedCommit, assemble_nodes


def _at(ptr, words):
    """numpy view of `words` u64 at a host address (the sharded helpers hand tile pointers to the engine)."""
    import ctypes
    return np.ctypeslib.as_array((ctypes.c_uint64 * int(words)).from_address(int(ptr)))


class OracleEngine:
    device = "cpu"

    def empty(self, words):
        return torch.zeros(int(words), dtype=torch.int64)

    # ---- rows next to the commit: stand-ins with the GpuEngine signatures, arithmetic by the C oracle ----
    def compute_levs(self, xi, openings, n_bits):
        lev = np.concatenate([C.lev(np.asarray(xi, dtype=np.uint64), o, n_bits, threads=1).reshape(-1) for o in openings])
        return torch.from_numpy(lev.view(np.int64).copy())

    def compute_evals(self, buf_ptr, size, n_bits, ext_bits, descs, lev, n_lev):
        buf = _at(buf_ptr, size << ext_bits).copy()
        levs = self._u(lev).reshape(n_lev, 1 << n_bits, 3)
        return C.evals({"b": (buf, size)}, [("b", o, d, l) for o, d, l in descs], [levs[i] for i in range(n_lev)], n_bits, ext_bits - n_bits)

    def x_div_x_sub_xi(self, xi, openings, n_bits, ext_bits):
        x = C.x_div_x_sub_xi(np.asarray(xi, dtype=np.uint64), openings, n_bits, ext_bits, threads=1).reshape(-1)
        return torch.from_numpy(x.view(np.int64).copy())

    def fri_pol(self, terms, evals, openings, xdiv, vf1, vf2, ext_bits):
        bufs, ev_map = {}, []
        for ptr, size, off, dim, prime in terms:
            bufs.setdefault((ptr, size), (_at(ptr, size << ext_bits).copy(), size))
            ev_map.append(((ptr, size), off, dim, prime))
        # xdiv here is the rank's ROW SLICE of the table: the oracle only indexes it by row, so the slice is a table of its own
        f = C.fri_polynomial(bufs, ev_map, evals, openings, self._u(xdiv), np.asarray(vf1, dtype=np.uint64), np.asarray(vf2, dtype=np.uint64),
                             ext_bits, threads=1)
        return torch.from_numpy(f.reshape(-1).view(np.int64).copy())

    def nnodes(self, height):
        return C.merkle_nnodes(height)

    @staticmethod
    def _u(t):
        return t.numpy().view(np.uint64)

    def lde(self, src, cols, n_bits, ext_bits, dst):
        self._u(dst)[:] = C.lde(self._u(src), cols, n_bits, ext_bits, threads=1)

    def merkelize_tiled(self, tiles, n_tiles, tile_cols, rows, nodes, split=False):
        t = self._u(tiles)[:n_tiles * rows * tile_cols].reshape(n_tiles, rows, tile_cols)
        full = np.ascontiguousarray(t.transpose(1, 0, 2)).reshape(-1)
        self._u(nodes)[:] = C.merkelize(full, n_tiles * tile_cols, rows, split, threads=1)

    def merkelize(self, elems, width, height, nodes):
        self._u(nodes)[:C.merkle_nnodes(height)] = C.merkelize(self._u(elems)[:width * height].copy(), width, height, threads=1)

    def group_proofs(self, tiles, n_tiles, tile_cols, rows, nodes, idxs, n_idx, rows_out, sib_out):
        width = n_tiles * tile_cols
        t = self._u(tiles)[:n_tiles * rows * tile_cols].reshape(n_tiles, rows, tile_cols)
        full = np.ascontiguousarray(t.transpose(1, 0, 2)).reshape(-1)
        depth = max(rows.bit_length() - 1, 0)
        ro, so = self._u(rows_out), self._u(sib_out)
        nd = self._u(nodes)[:C.merkle_nnodes(rows)].copy()
        for q in range(n_idx):
            i = int(self._u(idxs)[q])
            if i >= rows:
                ro[q * width:(q + 1) * width] = 0
                so[q * depth * 4:(q + 1) * depth * 4] = 0
            else:
                r, sb = C.group_proof(full, nd, width, rows, i)
                ro[q * width:(q + 1) * width] = r
                so[q * depth * 4:(q + 1) * depth * 4] = np.asarray(sb, dtype=np.uint64).reshape(-1)

    def compute_q(self, q_ext, q_dim, q_deg, n_bits, ext_bits, cmq):
        self._u(cmq)[:] = C.compute_q(self._u(q_ext).copy(), q_dim, q_deg, n_bits, ext_bits, threads=1)

    def fri_fold_range(self, src, in_layout, prev_bits, cur_bits, next_bits, step0_bits, challenge, row0, n_rows, pol_out, rows_out):
        """stand-in for pil2gpu_fri_fold_range_dev: the whole fold by the oracle, then only the requested rows are written"""
        nx = 1 << (prev_bits - cur_bits)
        a = self._u(src)[:3 << prev_bits].reshape(-1, 3)
        if in_layout == 1:                                           # rows of the previous layer -> polynomial order
            a = np.ascontiguousarray(a.reshape(1 << cur_bits, nx, 3).transpose(1, 0, 2)).reshape(-1, 3)
        nb = cur_bits if next_bits is None else next_bits
        if prev_bits == cur_bits:
            pol2 = a.copy()
        else:
            pol2, _ = C.fri_fold(a, prev_bits, cur_bits, None, step0_bits, [int(x) for x in challenge], threads=1)
        gs = 1 << (cur_bits - nb)
        rows = np.ascontiguousarray(pol2.reshape(gs, 1 << nb, 3).transpose(1, 0, 2)).reshape(-1)
        if n_rows == 0:
            row0, n_rows = 0, 1 << nb
        if rows_out is not None:
            self._u(rows_out)[row0 * 3 * gs:(row0 + n_rows) * 3 * gs] = rows[row0 * 3 * gs:(row0 + n_rows) * 3 * gs]
        if pol_out is not None:
            po = self._u(pol_out)[:3 << cur_bits].reshape(-1, 3)
            for j in range(gs):
                po[row0 + (j << nb):row0 + n_rows + (j << nb)] = pol2[row0 + (j << nb):row0 + n_rows + (j << nb)]

    def tree_from_digests(self, nodes, height):
        d = self._u(nodes)[:4 * height].copy()
        self._u(nodes)[:C.merkle_nnodes(height)] = C.merkelize(d, 4, height, threads=1)   # width 4 = passthrough leaves


class PeerOracleEngine(OracleEngine):
    """Stand-in for the fused peer-store exchange: `lde_scatter` delivers the rows to their owners itself (here through a
    gloo all-to-all), so the test covers the branch of ShardedCommit.commit that skips all_to_all_single and relies on the
    two one-word all-reduces for ordering."""

    def __init__(self, dist_mod):
        self.dist = dist_mod
        self.scatter_calls = 0

    def open_exchange(self, dist_mod, rank, world, recv_words):
        return {"recv": self.empty(recv_words), "rank": rank, "world": world, "flag": torch.zeros(1, dtype=torch.int32)}

    def close_exchange(self, ex):
        ex["closed"] = True

    def lde_scatter(self, src, cols, n_bits, ext_bits, dst, ex):
        self.lde(src, cols, n_bits, ext_bits, dst)
        self.dist.all_to_all_single(ex["recv"], dst)
        self.scatter_calls += 1


def _worker(rank, world, port, n_bits, blow, cols, split, q, peer=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        full = rng.integers(0, 0xFFFFFFFF00000001, size=(1 << n_bits, cols), dtype=np.uint64)
        eng = PeerOracleEngine(dist) if peer else OracleEngine()
        sc = ShardedCommit(eng, dist, rank, world)
        cg = sc.shard_cols(cols)
        slab = torch.from_numpy(np.ascontiguousarray(full[:, rank * cg:(rank + 1) * cg]).reshape(-1).view(np.int64))
        buf = sc.buffers(cols, n_bits, n_bits + blow)
        root = sc.commit(slab, cols, n_bits, n_bits + blow, buf, split)
        # proofQueries over the sharded tree: rows + siblings of the full-height tree, on every rank
        qs = [0, 1, (1 << (n_bits + blow)) - 1, (1 << (n_bits + blow)) // world, 37 % (1 << (n_bits + blow)), 5]
        rows_q, sib_q = buf["tree"].open(torch.tensor(qs, dtype=torch.int64))
        # a layer tree over rows every rank holds in full (FRI layers)
        lw, lh = 6, 1 << (n_bits - 1)
        layer = torch.from_numpy(rng.integers(0, 0xFFFFFFFF00000001, size=lw * lh, dtype=np.uint64).view(np.int64))
        lt, lroot = sc.commit_rows(layer, lw, lh, eng.empty(eng.nnodes(lh // world)), eng.empty(4 * world), eng.empty(max(8, eng.nnodes(world))))
        lrows, lsib = lt.open(torch.tensor([0, lh - 1, lh // 2], dtype=torch.int64))
        extra = (qs, rows_q.numpy().view(np.uint64).copy(), sib_q.numpy().view(np.uint64).copy(), layer.numpy().view(np.uint64).copy(),
                 lroot.numpy().view(np.uint64).copy(), lrows.numpy().view(np.uint64).copy(), lsib.numpy().view(np.uint64).copy())
        # evaluations at xi and the FRI polynomial over the row-sharded buffer
        from pil2_stark_js_b200.sharded import sharded_evals, sharded_fri_pol
        frng = np.random.default_rng(11)
        xi, vf1, vf2 = (frng.integers(0, 0xFFFFFFFF00000001, size=3, dtype=np.uint64) for _ in range(3))
        ev_map = [("t", c, 1, o) for o in (0, 1) for c in range(0, cols, 3)] + [("t", 9, 3, 1), ("t", 16, 3, 0)]
        sev = sharded_evals(eng, dist, rank, world, {"t": buf["tree"]}, ev_map, xi, [0, 1], n_bits, n_bits + blow)
        sf = sharded_fri_pol(eng, dist, rank, world, {"t": buf["tree"]}, ev_map, sev, xi, [0, 1], vf1, vf2, n_bits, n_bits + blow)
        extra = extra + ((xi, vf1, vf2, ev_map, sev, sf.numpy().view(np.uint64).copy()),)
        if peer:
            assert eng.scatter_calls == 1 and "peer stores" in sc.exchange_kind(buf)
            root = sc.commit(slab, cols, n_bits, n_bits + blow, buf, split)      # buffers are reusable
            sc.release(buf)
        else:
            assert "NCCL" in sc.exchange_kind(buf)
        q.put((rank, root.numpy().view(np.uint64).copy(), buf["nodes"].numpy().view(np.uint64).copy(),
               buf["top"].numpy().view(np.uint64).copy(), extra))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,split,peer", [(2, False, False), (4, False, False), (2, True, False), (2, False, True), (4, False, True)])
def test_sharded_commit_matches_single_process(world, split, peer):
    n_bits, blow, cols = 6, 1, 32
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bits, blow, cols, split, q, peer)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    full = rng.integers(0, 0xFFFFFFFF00000001, size=(1 << n_bits, cols), dtype=np.uint64)
    ext = C.lde(full.reshape(-1), cols, n_bits, n_bits + blow)
    nodes = C.merkelize(ext, cols, 1 << (n_bits + blow), split)
    E = 1 << (n_bits + blow)
    for _, root, _, _, extra in res:
        assert np.array_equal(root, nodes[-4:])
        qs, rows_q, sib_q, layer, lroot, lrows, lsib, frows = extra
        xi, vf1, vf2, ev_map, sev, sf = frows                        # sharded evaluations + FRI polynomial == single-process oracle
        want_ev = C.evals({"t": (ext, cols)}, ev_map, [C.lev(xi, o, n_bits) for o in (0, 1)], n_bits, blow)
        assert np.array_equal(sev, want_ev)
        want_f = C.fri_polynomial({"t": (ext, cols)}, ev_map, want_ev, [0, 1], C.x_div_x_sub_xi(xi, [0, 1], n_bits, n_bits + blow), vf1, vf2,
                                  n_bits + blow)
        assert np.array_equal(sf.reshape(-1, 3), want_f)
        for k, qi in enumerate(qs):
            r, sb = C.group_proof(ext, nodes, cols, E, qi)
            assert np.array_equal(rows_q[k], r) and np.array_equal(sib_q[k].reshape(-1), np.asarray(sb, dtype=np.uint64).reshape(-1)), f"query {qi}"
        if not split:
            lw, lh = 6, 1 << (n_bits - 1)
            lnodes = C.merkelize(layer, lw, lh)
            assert np.array_equal(lroot, lnodes[-4:])
            for k, qi in enumerate([0, lh - 1, lh // 2]):
                r, sb = C.group_proof(layer, lnodes, lw, lh, qi)
                assert np.array_equal(lrows[k], r) and np.array_equal(lsib[k].reshape(-1), np.asarray(sb, dtype=np.uint64).reshape(-1))
    rows_local = (1 << (n_bits + blow)) // world
    stitched = assemble_nodes([r[2] for r in res], res[0][3], rows_local, world, C.merkle_nnodes)
    assert np.array_equal(stitched, nodes)


def test_shard_validation():
    sc = ShardedCommit(OracleEngine(), None, 0, 3)
    with pytest.raises(ValueError):
        sc.shard_cols(32)
    sc = ShardedCommit(OracleEngine(), None, 0, 4)
    with pytest.raises(ValueError):
        sc.shard_cols(16)      # 4 columns per GPU: sponge chunks would straddle tiles
    assert sc.shard_cols(64) == 16


# ---------------------------------------------------------------- sharded FRI chain
def _fri_worker(rank, world, port, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pil2_stark_js_b200.sharded import ShardedFri, open_trees
        eng = OracleEngine()
        rng = np.random.default_rng(17)
        pol0 = rng.integers(0, 0xFFFFFFFF00000001, size=3 << steps[0], dtype=np.uint64)
        chal = [rng.integers(0, 0xFFFFFFFF00000001, size=3, dtype=np.uint64) for _ in steps]
        fri = ShardedFri(eng, dist, rank, world, steps, min_rows_per_rank=32)
        roots, final = fri.run(torch.from_numpy(pol0.view(np.int64).copy()), chal)
        queries = torch.tensor([0, 1, (1 << steps[0]) - 1, 12345 % (1 << steps[0]), 777], dtype=torch.int64)
        opened = open_trees(fri.query_pairs(queries))
        # quotient commit: transforms replicated, tree sharded
        from pil2_stark_js_b200.sharded import sharded_compute_q, ShardedCommit
        qe = rng.integers(0, 0xFFFFFFFF00000001, size=3 << 7, dtype=np.uint64)
        sc = ShardedCommit(eng, dist, rank, world)
        cmq = eng.empty(6 << 7)
        qt, qroot = sharded_compute_q(sc, torch.from_numpy(qe.view(np.int64).copy()), 3, 2, 6, 7, cmq, eng.empty(eng.nnodes(128 // world)), eng.empty(4 * world),
                                      eng.empty(max(8, eng.nnodes(world))))
        qrows, qsib = qt.open(torch.tensor([0, 127, 64], dtype=torch.int64))
        q.put((rank, fri.sharded, [r.numpy().view(np.uint64).copy() for r in roots], final.numpy().view(np.uint64).copy(),
               [(r.numpy().view(np.uint64).copy(), sb.numpy().view(np.uint64).copy()) for r, sb in opened],
               (qe, qroot.numpy().view(np.uint64).copy(), qrows.numpy().view(np.uint64).copy(), qsib.numpy().view(np.uint64).copy())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,steps", [(2, [11, 8, 5, 2]), (4, [12, 9, 7, 3]), (2, [9, 4])])
def test_sharded_fri_chain_matches_single_process(world, steps):
    """Every layer root, the final polynomial and the opened rows / sibling paths of every layer tree of the sharded chain equal
    the single-process oracle chain (fri.js:22-105); big layers are sharded, small ones redundant."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fri_worker, args=(r, world, port, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(17)
    pol0 = rng.integers(0, 0xFFFFFFFF00000001, size=3 << steps[0], dtype=np.uint64)
    chal = [rng.integers(0, 0xFFFFFFFF00000001, size=3, dtype=np.uint64) for _ in steps]
    # single-process chain
    cur = pol0.reshape(-1, 3)
    L = len(steps) - 1
    layer_rows, layer_nodes = [], []
    for s in range(L):
        if s >= 1:
            cur, _ = C.fri_fold(cur, steps[s - 1], steps[s], None, steps[0], [int(x) for x in chal[s]])
        gs = 1 << (steps[s] - steps[s + 1])
        rows = np.ascontiguousarray(cur.reshape(gs, 1 << steps[s + 1], 3).transpose(1, 0, 2)).reshape(-1)
        layer_rows.append(rows)
        layer_nodes.append(C.merkelize(rows, 3 * gs, 1 << steps[s + 1]))
    final, _ = C.fri_fold(cur, steps[L - 1], steps[L], None, steps[0], [int(x) for x in chal[L]])
    queries = [0, 1, (1 << steps[0]) - 1, 12345 % (1 << steps[0]), 777]
    assert any(res[0][1]) or len(steps) == 2                               # at least one layer really is sharded (the 2-step case is all replicas)
    for _, sharded, roots, fin, opened, qres in res:
        qe, qroot, qrows, qsib = qres                                      # sharded quotient commit == single-process oracle
        want_q = C.compute_q(qe, 3, 2, 6, 7)
        qnodes = C.merkelize(want_q, 6, 128)
        assert np.array_equal(qroot, qnodes[-4:])
        for k, qi in enumerate([0, 127, 64]):
            r, sb = C.group_proof(want_q, qnodes, 6, 128, qi)
            assert np.array_equal(qrows[k], r) and np.array_equal(qsib[k].reshape(-1), np.asarray(sb, dtype=np.uint64).reshape(-1))
        assert np.array_equal(fin.reshape(-1, 3), final)
        for s in range(L):
            assert np.array_equal(roots[s], layer_nodes[s][-4:]), f"root of layer {s}"
            w, h = 3 << (steps[s] - steps[s + 1]), 1 << steps[s + 1]
            rows_q, sib_q = opened[s]
            for k, qi in enumerate(queries):
                r, sb = C.group_proof(layer_rows[s], layer_nodes[s], w, h, qi % h)
                assert np.array_equal(rows_q[k], r), (s, qi)
                assert np.array_equal(sib_q[k].reshape(-1), np.asarray(sb, dtype=np.uint64).reshape(-1)), (s, qi)
