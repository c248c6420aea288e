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
        if peer:
            assert eng.scatter_calls == 1 and "peer stores" in sc.exchange_kind(buf)
            root = sc.commit(slab, cols, n_bits, n_bits + blow, buf, split)      # buffers are reusable
            sc.release(buf)
        else:
            assert "NCCL" in sc.exchange_kind(buf)
        q.put((rank, root.numpy().view(np.uint64).copy(), buf["nodes"].numpy().view(np.uint64).copy(),
               buf["top"].numpy().view(np.uint64).copy()))
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
    for _, root, _, _ in res:
        assert np.array_equal(root, nodes[-4:])
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
