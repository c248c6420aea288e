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


class OracleEngine:
    def empty(self, words):
        return torch.zeros(int(words), dtype=torch.int64)

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
        qs, rows_q, sib_q, layer, lroot, lrows, lsib = extra
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
