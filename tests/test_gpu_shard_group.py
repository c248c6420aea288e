"""The multi-GPU commit group of the C ABI (pil2gpu_shard_*: peer-mapped receive buffers + mailboxes, flag barriers, no collective
library) against the single-process oracle.  Two ways of forming a group are exercised:
  * one process, several contexts wired with pil2gpu_shard_connect_local -- on ONE GPU here (every rank is a context with its own
    stream on device 0), the way a single Node / Python thread would drive several GPUs;
  * one process per rank, handles exchanged through a queue, pil2gpu_shard_connect (CUDA IPC) -- the Node-worker shape; needs 2 GPUs.
Roots, every node, and the opened rows / sibling paths must equal extendAndMerkelize + getGroupProof of the whole trace."""
import ctypes

import numpy as np
import pytest

from oracle import gl_oracle as C

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


def _lib():
    from pil2_stark_js_b200 import _lib
    return _lib.load(), _lib.check


class _Rank:
    """One rank of a group living in this process: its own context (own stream) on `device`."""

    def __init__(self, device, rank, world, recv_words, stage_words):
        import torch
        self.torch = torch
        self.L, self.check = _lib()
        self.dev = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.h = ctypes.c_void_p()
        self.check(self.L.pil2gpu_create(device, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(self.h)))
        self.sh = ctypes.c_void_p()
        self.check(self.L.pil2gpu_shard_create(self.h, rank, world, recv_words, stage_words, ctypes.byref(self.sh)))
        self.rank, self.world = rank, world

    def dev_u64(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).reshape(-1)).to(self.dev)
        self.torch.cuda.synchronize()
        return t

    def empty(self, words):
        return self.torch.empty(int(words), dtype=self.torch.int64, device=self.dev)

    def close(self):
        self.check(self.L.pil2gpu_shard_destroy(self.sh))
        self.L.pil2gpu_destroy(self.h)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _u(t):
    return t.cpu().numpy().view(np.uint64).copy()


def _check_against_oracle(full, cols, n_bits, ext_bits, world, roots, nodes_per_rank, tops, queries, rows, sibs, split=False):
    from pil2_stark_js_b200.sharded import assemble_nodes
    E = 1 << ext_bits
    ext = C.lde(full.reshape(-1), cols, n_bits, ext_bits)
    nodes = C.merkelize(ext, cols, E, split)
    for r in range(world):
        assert np.array_equal(roots[r], nodes[-4:]), f"root of rank {r}"
    stitched = assemble_nodes(nodes_per_rank, tops[0], E // world, world, C.merkle_nnodes)
    assert np.array_equal(stitched, nodes)
    for r in range(world):
        for k, qi in enumerate(queries):
            want_row, want_sib = C.group_proof(ext, nodes, cols, E, int(qi))
            assert np.array_equal(rows[r][k], want_row)
            assert np.array_equal(sibs[r][k].reshape(-1), np.asarray(want_sib, dtype=np.uint64).reshape(-1))


@pytest.mark.parametrize("world,n_bits,blow,cols,split", [(2, 10, 1, 32, False), (4, 9, 2, 64, False), (2, 8, 1, 64, True), (1, 8, 1, 24, False)])
def test_shard_group_one_process(world, n_bits, blow, cols, split):
    """Every rank is a context on device 0; one host thread enqueues the ranks' calls one after the other -- nothing in the API blocks
    the host, the ranks meet in the flag barriers on the device.  (Several ranks on ONE device need their streams on different hardware
    queues -- tests/conftest.py raises CUDA_DEVICE_MAX_CONNECTIONS; with one context per device, the production shape, there is nothing
    to configure.)"""
    import torch
    L, check = _lib()
    ext_bits, cg = n_bits + blow, cols // world
    E = 1 << ext_bits
    rows_local = E // world
    rng = np.random.default_rng(100 + world)
    full = rng.integers(0, P, size=(1 << n_bits, cols), dtype=np.uint64)
    queries = np.array([0, E - 1, rows_local, rows_local - 1, E // 2 + 3, 5], dtype=np.uint64) % E
    nq, depth = len(queries), ext_bits
    ranks = [_Rank(0, r, world, cg << ext_bits, nq * (cols + 4 * depth)) for r in range(world)]
    try:
        group = (ctypes.c_void_p * world)(*[rk.sh.value for rk in ranks])
        check(L.pil2gpu_shard_connect_local(group, world))
        nn = int(L.pil2gpu_merkle_nnodes(rows_local))
        bufs = []
        for rk in ranks:
            src = rk.dev_u64(full[:, rk.rank * cg:(rk.rank + 1) * cg])
            bufs.append({"src": src, "work": rk.empty(cg << ext_bits), "nodes": rk.empty(nn), "root": rk.empty(4), "idx": rk.dev_u64(queries),
                         "rows": rk.empty(nq * cols), "sib": rk.empty(nq * depth * 4)})
        for it in range(2):                                  # twice: the second commit reuses receive buffers and mailboxes
            for rk, b in zip(ranks, bufs):
                check(L.pil2gpu_shard_commit_dev(rk.sh, _p(b["src"]), _p(b["work"]), cols, n_bits, ext_bits, int(split), _p(b["nodes"]), _p(b["root"])))
            for rk, b in zip(ranks, bufs):
                check(L.pil2gpu_shard_open_dev(rk.sh, _p(b["nodes"]), cols, ext_bits, _p(b["idx"]), nq, _p(b["rows"]), _p(b["sib"])))
        for rk in ranks:
            check(L.pil2gpu_shard_status(rk.sh))             # synchronises; reports a barrier that timed out
        torch.cuda.synchronize()
        tops = []
        for rk in ranks:
            words = max(8, int(L.pil2gpu_merkle_nnodes(world)))
            host = np.empty(words, dtype=np.uint64)
            check(L.pil2gpu_d2h(rk.h, ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(L.pil2gpu_shard_top_nodes_dev(rk.sh)), words * 8))
            check(L.pil2gpu_sync(rk.h))
            tops.append(host)
        _check_against_oracle(full, cols, n_bits, ext_bits, world, [_u(b["root"]) for b in bufs], [_u(b["nodes"]) for b in bufs], tops, queries,
                              [_u(b["rows"]).reshape(nq, cols) for b in bufs], [_u(b["sib"]).reshape(nq, depth, 4) for b in bufs], split)
    finally:
        for rk in ranks:
            rk.close()


def test_shard_group_argument_checks():
    L, check = _lib()
    from pil2_stark_js_b200._lib import Pil2GpuError
    rk = _Rank(0, 0, 2, 1 << 12, 64)
    try:
        sh = ctypes.c_void_p()
        for bad in [(4, 4), (0, 3), (0, 32), (0, 0)]:      # rank >= world, world not a power of two, too many ranks, empty group
            with pytest.raises(Pil2GpuError):
                check(L.pil2gpu_shard_create(rk.h, bad[0], bad[1], 64, 0, ctypes.byref(sh)))
        a = rk.empty(1 << 12)
        with pytest.raises(Pil2GpuError, match="not connected"):
            check(L.pil2gpu_shard_commit_dev(rk.sh, _p(a), _p(a), 16, 4, 5, 0, _p(a), None))
        with pytest.raises(Pil2GpuError, match="not connected"):
            check(L.pil2gpu_shard_barrier(rk.sh))
        with pytest.raises(Pil2GpuError):
            check(L.pil2gpu_shard_connect(rk.sh, None, 2))
        lone = _Rank(0, 0, 1, 1 << 12, 64)
        try:
            with pytest.raises(Pil2GpuError, match="multiple of the number of ranks"):
                check(L.pil2gpu_shard_commit_dev(lone.sh, _p(a), _p(a), 0, 4, 5, 0, _p(a), None))
            with pytest.raises(Pil2GpuError, match="too small"):
                check(L.pil2gpu_shard_commit_dev(lone.sh, _p(a), _p(a), 16, 10, 11, 0, _p(a), None))
            idx = lone.dev_u64(np.zeros(64, dtype=np.uint64))
            with pytest.raises(Pil2GpuError, match="staging too small"):
                check(L.pil2gpu_shard_open_dev(lone.sh, _p(a), 16, 5, _p(idx), 64, _p(a), _p(a)))
        finally:
            lone.close()
    finally:
        rk.close()


def test_shard_barrier_timeout_is_reported(monkeypatch):
    """A peer that never arrives must not hang the GPU: the barrier kernel gives up after PIL2GPU_SHARD_TIMEOUT_MS and
    pil2gpu_shard_status reports it."""
    monkeypatch.setenv("PIL2GPU_SHARD_TIMEOUT_MS", "200")
    L, check = _lib()
    from pil2_stark_js_b200._lib import Pil2GpuError
    ranks = [_Rank(0, r, 2, 1 << 10, 0) for r in range(2)]
    try:
        group = (ctypes.c_void_p * 2)(*[rk.sh.value for rk in ranks])
        check(L.pil2gpu_shard_connect_local(group, 2))
        check(L.pil2gpu_shard_barrier(ranks[0].sh))          # rank 1 never calls
        with pytest.raises(Pil2GpuError, match="timed out"):
            check(L.pil2gpu_shard_status(ranks[0].sh))
    finally:
        for rk in ranks:
            rk.close()


def _ipc_worker(rank, world, devices, n_bits, blow, cols, qin, qout):
    import torch
    torch.cuda.set_device(devices[rank])
    L, check = _lib()
    ext_bits, cg = n_bits + blow, cols // world
    E = 1 << ext_bits
    rng = np.random.default_rng(77)
    full = rng.integers(0, P, size=(1 << n_bits, cols), dtype=np.uint64)
    queries = np.array([1, E - 2, E // world, E // world - 1], dtype=np.uint64)
    nq, depth = len(queries), ext_bits
    rk = _Rank(devices[rank], rank, world, cg << ext_bits, nq * (cols + 4 * depth))
    try:
        handle = (ctypes.c_uint8 * 128)()
        check(L.pil2gpu_shard_handles(rk.sh, handle))
        qout.put((rank, bytes(handle)))
        allh = qin.get(timeout=120)                          # the parent gathers and redistributes: no collective library anywhere
        check(L.pil2gpu_shard_connect(rk.sh, (ctypes.c_uint8 * (128 * world)).from_buffer_copy(allh), world))
        nn = int(L.pil2gpu_merkle_nnodes(E // world))
        src = rk.dev_u64(full[:, rank * cg:(rank + 1) * cg])
        work, nodes, root, idx = rk.empty(cg << ext_bits), rk.empty(nn), rk.empty(4), rk.dev_u64(queries)
        rows, sib = rk.empty(nq * cols), rk.empty(nq * depth * 4)
        for _ in range(2):
            check(L.pil2gpu_shard_commit_dev(rk.sh, _p(src), _p(work), cols, n_bits, ext_bits, 0, _p(nodes), _p(root)))
            check(L.pil2gpu_shard_open_dev(rk.sh, _p(nodes), cols, ext_bits, _p(idx), nq, _p(rows), _p(sib)))
        check(L.pil2gpu_shard_status(rk.sh))
        words = max(8, int(L.pil2gpu_merkle_nnodes(world)))
        top = np.empty(words, dtype=np.uint64)
        check(L.pil2gpu_d2h(rk.h, ctypes.c_void_p(top.ctypes.data), ctypes.c_void_p(L.pil2gpu_shard_top_nodes_dev(rk.sh)), words * 8))
        check(L.pil2gpu_sync(rk.h))
        qout.put((rank, _u(root), _u(nodes), top, _u(rows).reshape(nq, cols), _u(sib).reshape(nq, depth, 4)))
        qin.get(timeout=120)                                 # keep the mappings alive until every rank has finished
    finally:
        rk.close()


@pytest.mark.parametrize("world", [2])
def test_shard_group_ipc_processes(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs (one process per GPU, CUDA IPC)")
    n_bits, blow, cols = 11, 1, 64
    ctx_mp = mp.get_context("spawn")
    qins, qout = [ctx_mp.Queue() for _ in range(world)], ctx_mp.Queue()
    procs = [ctx_mp.Process(target=_ipc_worker, args=(r, world, list(range(world)), n_bits, blow, cols, qins[r], qout)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        hs = dict(qout.get(timeout=300) for _ in range(world))
        for qi in qins:
            qi.put(b"".join(hs[r] for r in range(world)))
        res = sorted([qout.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    finally:
        for qi in qins:
            qi.put(b"done")
        for p in procs:
            p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    rng = np.random.default_rng(77)
    full = rng.integers(0, P, size=(1 << n_bits, cols), dtype=np.uint64)
    E = 1 << (n_bits + blow)
    queries = np.array([1, E - 2, E // world, E // world - 1], dtype=np.uint64)
    _check_against_oracle(full, cols, n_bits, n_bits + blow, world, [r[1] for r in res], [r[2] for r in res], [r[3] for r in res], queries,
                          [r[4] for r in res], [r[5] for r in res])
