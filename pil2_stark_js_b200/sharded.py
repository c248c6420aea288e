"""Multi-GPU commit (one process per GPU, torch.distributed for the plumbing) -- SURVEY 8(e).

    rank g owns columns [g*C/G, (g+1)*C/G) of the trace            (columns are independent NTTs: no communication)
    1. LDE of the column slab on the local GPU                       N x C/G  ->  E x C/G
    2. ONE exchange over NVLink: rank g sends rows [h*E/G, (h+1)*E/G) of its slab to rank h; rank h receives G tiles
       (one per source rank) of E/G rows x C/G columns -- its row range across all columns, kept as column tiles.
       Default: the exchange is FUSED into the LDE -- the last butterfly pass stores every finished row straight into the
       owner's receive buffer (peer memory mapped with CUDA IPC, pil2gpu_lde_scatter_dev), so the transfer overlaps the
       pass tile by tile and only a one-word all-reduce (the "all stores have landed" barrier) remains.  Fallback
       (PIL2GPU_EXCHANGE=nccl, or IPC mapping unavailable): local LDE followed by one NCCL all_to_all_single.
    3. leaf hashing + subtree of the E/G local rows straight from the tiles (pil2gpu_merkelize_tiled_dev, no repack)
    4. the G sub-roots (G x 32 bytes) reach every rank -- stored into the peers' mailboxes + a flag barrier (pil2gpu_shard_hash_dev), or
       an all-gather on the NCCL path; the top log2(G) levels are hashed redundantly on every rank
The root (and every node) equals the single-GPU tree: contiguous leaf ranges make each local root the level-log2(E/G)
node of the reference layout.  The reference has no counterpart (it is single-process, workerpool threads).

The choreography is written against a small `engine` interface so that the same code runs on CUDA tensors through the
C ABI (GpuEngine) and, in the CPU tests, on a stand-in engine under the gloo backend.
"""
import ctypes

import numpy as np


class GpuEngine:
    """Device work through libpil2gpu.so; tensors are torch int64 CUDA tensors holding u64 bit patterns."""

    def __init__(self, torch, device_index):
        from . import _lib
        self.torch, self.L, self.check = torch, _lib.load(), _lib.check
        h = ctypes.c_void_p()
        # The library must enqueue on the stream torch.distributed orders its collectives against.  torch reports the
        # default stream as handle 0, which pil2gpu_create reads as "make your own stream": pass cudaStreamLegacy (0x1).
        stream = torch.cuda.current_stream().cuda_stream or 1
        self.check(self.L.pil2gpu_create(device_index, ctypes.c_void_p(stream), ctypes.byref(h)))
        self.h = h
        self.device = torch.device("cuda", device_index)

    def empty(self, words):
        return self.torch.empty(int(words), dtype=self.torch.int64, device=self.device)

    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def nnodes(self, height):
        return int(self.L.pil2gpu_merkle_nnodes(height))

    def lde(self, src, cols, n_bits, ext_bits, dst):
        self.check(self.L.pil2gpu_lde_dev(self.h, self._p(src), self._p(dst), cols, n_bits, ext_bits))

    # ---- peer-memory exchange: the commit group of the C ABI (pil2gpu_shard_*: receive buffers + mailboxes mapped with CUDA IPC,
    # flag barriers on the stream); torch.distributed only carries the 2 x 64-byte handles of every rank, once ----
    def _tensor_view(self, ptr, words):
        class _Raw:       # torch view of a raw device allocation
            pass
        view = _Raw()
        view.__cuda_array_interface__ = {"shape": (int(words),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}
        return self.torch.as_tensor(view, device=self.device)

    def open_exchange(self, dist, rank, world, recv_words, stage_words=0):
        """Create this rank's end of the commit group and connect it to its peers.  Returns {"recv": tensor view of the receive
        buffer, "peers": host array of device pointers, "shard": handle, ...} or None when peer mapping is unavailable on any rank
        (the caller then uses the NCCL all-to-all)."""
        import os
        torch = self.torch
        if os.environ.get("PIL2GPU_EXCHANGE", "peer") == "nccl":
            return None
        ok, sh, err = 1, ctypes.c_void_p(), ""
        try:
            self.check(self.L.pil2gpu_shard_create(self.h, rank, world, int(recv_words), int(stage_words), ctypes.byref(sh)))
            handle = (ctypes.c_uint8 * 128)()
            self.check(self.L.pil2gpu_shard_handles(sh, handle))
        except Exception as ex:       # noqa: BLE001 -- any failure means "no peer mapping here"; all ranks must agree below
            ok, err = 0, str(ex)
            handle = (ctypes.c_uint8 * 128)()
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        if ok:
            try:
                allh = (ctypes.c_uint8 * (128 * world)).from_buffer_copy(b"".join(handles))
                self.check(self.L.pil2gpu_shard_connect(sh, allh, world))
            except Exception as ex:   # noqa: BLE001
                ok, err = 0, str(ex)
        flag = torch.tensor([ok], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if err:
                print(f"[pil2gpu] rank {rank}: peer exchange unavailable ({err}); using NCCL all_to_all", flush=True)
            if sh:
                self.L.pil2gpu_shard_destroy(sh)
            return None
        recv = self._tensor_view(self.L.pil2gpu_shard_recv_dev(sh), recv_words)
        peers = ctypes.cast(self.L.pil2gpu_shard_peer_recv(sh), ctypes.POINTER(ctypes.c_void_p))
        sub = self._tensor_view(self.L.pil2gpu_shard_sub_roots_dev(sh), 4 * world)
        top = self._tensor_view(self.L.pil2gpu_shard_top_nodes_dev(sh), max(8, self.nnodes(world)))
        return {"recv": recv, "peers": peers, "shard": sh, "rank": rank, "world": world, "sub": sub, "top": top}

    def close_exchange(self, ex):
        if ex:
            self.torch.cuda.synchronize()
            self.check(self.L.pil2gpu_shard_status(ex["shard"]))
            self.L.pil2gpu_shard_destroy(ex["shard"])

    def exchange_barrier(self, ex):
        """Flag barrier over the mailboxes, enqueued on the stream (no NCCL call, no host round trip)."""
        self.check(self.L.pil2gpu_shard_barrier(ex["shard"]))

    def shard_hash(self, ex, cols, ext_bits, split, nodes, root_out):
        """Leaf hashing + local subtree + sub-root exchange (peer stores + flag barrier) + top tree: the root on every rank."""
        self.check(self.L.pil2gpu_shard_hash_dev(ex["shard"], cols, ext_bits, int(split), self._p(nodes), self._p(root_out)))

    def lde_scatter(self, src, cols, n_bits, ext_bits, dst, ex):
        self.check(self.L.pil2gpu_lde_scatter_dev(self.h, self._p(src), self._p(dst), cols, n_bits, ext_bits, ex["peers"], ex["world"],
                                                  ex["rank"], 0, 0))

    def lde_scatter_host(self, src_host, cols, n_bits, ext_bits, ex):
        """Pinned host slab -> peers: upload in sub-slabs overlapped with the LDE + peer stores (asynchronous)."""
        self.check(self.L.pil2gpu_lde_scatter(self.h, ctypes.c_void_p(src_host.data_ptr()), cols, cols, n_bits, ext_bits, ex["peers"],
                                              ex["world"], ex["rank"]))

    def merkelize_tiled(self, tiles, n_tiles, tile_cols, rows, nodes, split=False):
        self.check(self.L.pil2gpu_merkelize_tiled_dev(self.h, self._p(tiles), n_tiles, tile_cols, rows * tile_cols, rows, int(split),
                                                      self._p(nodes)))

    def tree_from_digests(self, nodes, height):
        self.check(self.L.pil2gpu_merkle_tree_from_digests_dev(self.h, self._p(nodes), height))

    def merkelize(self, elems, width, height, nodes):
        self.check(self.L.pil2gpu_merkelize_dev(self.h, self._p(elems), width, height, 0, self._p(nodes)))

    def group_proofs(self, tiles, n_tiles, tile_cols, rows, nodes, idxs, n_idx, rows_out, sib_out):
        """getGroupProof (merklehash_p.js:142-168) for n_idx device-resident leaf indices of a (tiled) device tree; an index
        >= rows marks a slot owned by another rank and is zero-filled."""
        t = ctypes.c_void_p()
        self.check(self.L.pil2gpu_tree_wrap_tiled_dev(self.h, self._p(tiles), n_tiles, tile_cols, rows * tile_cols, self._p(nodes), rows,
                                                      ctypes.byref(t)))
        try:
            self.check(self.L.pil2gpu_tree_group_proofs_dev(self.h, t, self._p(idxs), n_idx, self._p(rows_out), self._p(sib_out)))
        finally:
            self.L.pil2gpu_tree_free(None, t)       # a wrapper: owns nothing on the device

    # ---- evaluations at xi / FRI polynomial on device tensors (the rows next to the commit) ----
    def compute_levs(self, xi, openings, n_bits):
        lev = self.empty(len(openings) * (3 << n_bits))
        xi = np.ascontiguousarray(xi, dtype=np.uint64)
        for i, o in enumerate(openings):
            self.check(self.L.pil2gpu_compute_lev_dev(self.h, ctypes.c_void_p(xi.ctypes.data), int(o), n_bits,
                                                      ctypes.c_void_p(lev.data_ptr() + 8 * i * (3 << n_bits))))
        return lev

    def compute_evals(self, buf_ptr, size, n_bits, ext_bits, descs, lev, n_lev):
        from . import _lib
        n = len(descs)
        out = np.empty((n, 3), dtype=np.uint64)
        if n:
            arr = (_lib.EvalDesc * n)(*[_lib.EvalDesc(int(o), int(d), int(l)) for o, d, l in descs])
            self.check(self.L.pil2gpu_compute_evals_dev(self.h, ctypes.c_void_p(buf_ptr), size, n_bits, ext_bits, arr, n, self._p(lev), n_lev,
                                                        ctypes.c_void_p(out.ctypes.data)))
        return out

    def x_div_x_sub_xi(self, xi, openings, n_bits, ext_bits):
        out = self.empty(3 * len(openings) << ext_bits)
        xi = np.ascontiguousarray(xi, dtype=np.uint64)
        op = (ctypes.c_int32 * len(openings))(*[int(o) for o in openings])
        self.check(self.L.pil2gpu_x_div_x_sub_xi_dev(self.h, ctypes.c_void_p(xi.ctypes.data), op, len(openings), n_bits, ext_bits, self._p(out)))
        return out

    def fri_pol(self, terms, evals, openings, xdiv, vf1, vf2, ext_bits):
        from . import _lib
        n = len(terms)
        arr = (_lib.FriTerm * n)(*[_lib.FriTerm(int(p), int(size), int(off), int(dim), int(prime)) for p, size, off, dim, prime in terms])
        ev = np.ascontiguousarray(np.asarray(evals, dtype=np.uint64).reshape(-1))
        op = (ctypes.c_int32 * len(openings))(*[int(o) for o in openings])
        a, b = (np.ascontiguousarray(v, dtype=np.uint64).reshape(-1) for v in (vf1, vf2))
        f = self.empty(3 << ext_bits)
        self.check(self.L.pil2gpu_fri_pol_dev(self.h, arr, n, ctypes.c_void_p(ev.ctypes.data), op, len(openings), self._p(xdiv),
                                              ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data), ext_bits, self._p(f)))
        return f

    def compute_q(self, q_ext, q_dim, q_deg, n_bits, ext_bits, cmq):
        self.check(self.L.pil2gpu_compute_q_dev(self.h, self._p(q_ext), q_dim, q_deg, n_bits, ext_bits, self._p(cmq)))

    def fri_fold_range(self, src, in_layout, prev_bits, cur_bits, next_bits, step0_bits, challenge, row0, n_rows, pol_out, rows_out):
        """One rank's share of a FRI fold (pil2gpu_fri_fold_range_dev): rows [row0, row0 + n_rows) of the next layer's
        transposed buffer at their absolute positions; src is the polynomial (in_layout 0) or the previous layer's rows (1)."""
        ch = np.ascontiguousarray(challenge, dtype=np.uint64)
        self.check(self.L.pil2gpu_fri_fold_range_dev(self.h, self._p(src), int(in_layout), prev_bits, cur_bits, -1 if next_bits is None else next_bits,
                                                     step0_bits, ctypes.c_void_p(ch.ctypes.data), int(row0), int(n_rows),
                                                     self._p(pol_out) if pol_out is not None else None,
                                                     self._p(rows_out) if rows_out is not None else None))

    def launches(self):
        return int(self.L.pil2gpu_launch_count(self.h))


def _tile_terms(tree, offset, dim):
    """(tile pointer, column inside the tile) of a polynomial of a ShardedTree; its columns must lie in one tile."""
    t = offset // tree.tile_cols
    local = offset - t * tree.tile_cols
    if local + dim > tree.tile_cols:
        raise ValueError("a polynomial's columns straddle two column tiles")
    return tree.tiles.data_ptr() + 8 * t * tree.rows_local * tree.tile_cols, local


def sharded_evals(engine, dist, rank, world, trees, ev_map, xi, openings, n_bits, ext_bits):
    """computeEvalsStark (stark_gen_helpers.js:210-273) over row-sharded extended buffers: every rank sums over the base rows it
    holds (its slice of the LEv vectors), the per-rank partial sums are gathered and added mod p.  trees: name -> ShardedTree;
    ev_map: list of (tree name, column, dim, opening index).  Returns the (n, 3) uint64 array on every rank."""
    import torch
    g_bits = world.bit_length() - 1
    nb, neb = n_bits - g_bits, ext_bits - g_bits
    if nb < 0:
        raise ValueError("more ranks than base rows")
    n_lev, Nl = len(openings), 1 << nb
    lev = engine.compute_levs(xi, openings, n_bits).view(n_lev, 1 << n_bits, 3)[:, rank * Nl:(rank + 1) * Nl, :].contiguous().view(-1)
    out = np.zeros((len(ev_map), 3), dtype=object)
    by_tile = {}
    for i, (name, off, dim, oi) in enumerate(ev_map):
        ptr, local = _tile_terms(trees[name], off, dim)
        by_tile.setdefault((ptr, trees[name].tile_cols), []).append((i, local, dim, oi))
    P = 0xFFFFFFFF00000001
    part = np.zeros((len(ev_map), 3), dtype=np.uint64)
    for (ptr, size), items in by_tile.items():
        vals = engine.compute_evals(ptr, size, nb, neb, [(l, d, o) for _, l, d, o in items], lev, n_lev)
        for (i, _, _, _), v in zip(items, vals):
            part[i] = v
    if world == 1:
        return part
    mine = torch.from_numpy(part.view(np.int64).reshape(-1)).to(engine.device)
    allp = engine.empty(world * mine.numel())
    dist.all_gather_into_tensor(allp, mine)
    allp = allp.cpu().numpy().view(np.uint64).reshape(world, len(ev_map), 3)
    for i in range(len(ev_map)):
        for c in range(3):
            out[i, c] = sum(int(allp[r, i, c]) for r in range(world)) % P
    return out.astype(np.uint64)


def sharded_fri_pol(engine, dist, rank, world, trees, ev_map, evals, xi, openings, vf1, vf2, n_bits, ext_bits):
    """computeFRIStark (stark_gen_helpers.js:275-334) over row-sharded extended buffers: f_ext is row-local, so every rank
    evaluates friExp on the rows it holds (each column tile enters as one buffer of pil2gpu_fri_pol_dev) and one all-gather
    assembles the 2^ext_bits x 3 polynomial for the FRI chain.  ev_map: list of (tree name, column, dim, prime)."""
    g_bits = world.bit_length() - 1
    neb = ext_bits - g_bits
    R, n_open = 1 << neb, len(openings)
    xdiv = engine.x_div_x_sub_xi(xi, openings, n_bits, ext_bits).view(1 << ext_bits, n_open * 3)[rank * R:(rank + 1) * R].contiguous().view(-1)
    terms = []
    for name, off, dim, prime in ev_map:
        ptr, local = _tile_terms(trees[name], off, dim)
        terms.append((ptr, trees[name].tile_cols, local, dim, prime))
    f_local = engine.fri_pol(terms, evals, openings, xdiv, vf1, vf2, neb)
    if world == 1:
        return f_local
    f = engine.empty(3 << ext_bits)
    dist.all_gather_into_tensor(f, f_local)
    return f


def sharded_compute_q(sc, q_ext, q_dim, q_deg, n_bits, ext_bits, cmq, nodes, sub, top):
    """computeQStark (stark_gen_helpers.js:168-208) over G ranks that all hold q_ext (the quotient is evaluated row-locally, so after
    an all-gather -- or when every rank evaluates the whole domain -- each rank has it): the 3-6 column transforms are too narrow
    to shard and run on every rank, the tree -- two thirds of the cost at production sizes -- is hashed sharded by rows
    (ShardedCommit.commit_rows).  cmq: qDim*qDeg << ext_bits words; nodes: nnodes(2^ext_bits / G); sub: 4 G; top: nnodes(G).
    Returns (ShardedTree, root tensor)."""
    sc.e.compute_q(q_ext, q_dim, q_deg, n_bits, ext_bits, cmq)
    return sc.commit_rows(cmq, q_dim * q_deg, 1 << ext_bits, nodes, sub, top)


class ShardedTree:
    """A Merkle tree whose leaves are spread over the ranks in contiguous ranges: rank h holds rows [h*R, (h+1)*R) (as
    n_tiles column tiles), the subtree over them (`nodes`, reference layout of a height-R tree) and -- like every rank --
    the gathered sub-roots (`sub`) with the levels above them (`top`, reference layout of a height-G tree)."""

    def __init__(self, engine, dist, rank, world, tiles, n_tiles, tile_cols, rows_local, nodes, sub, top):
        self.e, self.dist, self.rank, self.world = engine, dist, rank, world
        self.tiles, self.n_tiles, self.tile_cols, self.rows_local = tiles, n_tiles, tile_cols, rows_local
        self.nodes, self.sub, self.top = nodes, sub, top
        self.width = n_tiles * tile_cols
        self.replica = None            # (global rank, global world, dist) when this is a whole tree replicated on every rank

    def reduce_to_root(self):
        """Sub-root all-gather + the top log2(G) levels (hashed redundantly on every rank).  Returns the 4-word root tensor."""
        G, e = self.world, self.e
        if G == 1:
            return self.nodes[-4:]
        nn = e.nnodes(self.rows_local)
        self.dist.all_gather_into_tensor(self.sub, self.nodes[nn - 4:nn].contiguous())
        self.top[:4 * G].copy_(self.sub)
        e.tree_from_digests(self.top, G)
        nt = e.nnodes(G)
        return self.top[nt - 4:nt]

    def open_words(self, n_queries):
        """Words of the flat buffer open_local fills: Q rows of `width` words, then Q x local depth x 4 sibling words."""
        dl = max(self.rows_local.bit_length() - 1, 0)
        return n_queries * self.width + n_queries * dl * 4

    def open_local(self, queries, flat):
        """This rank's contribution to getGroupProof for global leaf indices `queries`: the rows and lower siblings of the
        leaves it owns, zeros for the others, written into `flat` (open_words(Q) words) -- ready for a sum all-reduce."""
        e, R = self.e, self.rows_local
        Q = int(queries.numel())
        dl = max(R.bit_length() - 1, 0)
        local = queries - self.rank * R
        idx = local.clone()
        idx[(local < 0) | (local >= R)] = -1                              # 0xFFFF...: "not mine"
        rows_out = flat[:Q * self.width]
        if dl:
            e.group_proofs(self.tiles, self.n_tiles, self.tile_cols, R, self.nodes, idx, Q, rows_out, flat[Q * self.width:Q * self.width + Q * dl * 4])
        else:
            e.group_proofs(self.tiles, self.n_tiles, self.tile_cols, R, self.nodes, idx, Q, rows_out, e.empty(4))

    def open_finish(self, queries, flat):
        """After the all-reduce of `flat`: (rows [Q, width], siblings [Q, depth, 4]); the top log2(G) siblings come from the
        replicated top tree."""
        e, G, R = self.e, self.world, self.rows_local
        Q = int(queries.numel())
        dl = max(R.bit_length() - 1, 0)
        rows = flat[:Q * self.width].view(Q, self.width)
        sib_local = flat[Q * self.width:Q * self.width + Q * dl * 4].view(Q, dl, 4)
        if G == 1:
            return rows, sib_local
        dt = G.bit_length() - 1
        owners = (queries // R).contiguous()
        sub_rows, sib_top = e.empty(Q * 4), e.empty(Q * dt * 4)
        e.group_proofs(self.sub, 1, 4, G, self.top, owners, Q, sub_rows, sib_top)
        import torch
        return rows, torch.cat([sib_local, sib_top.view(Q, dt, 4)], dim=1)

    def open(self, queries):
        """getGroupProof for global leaf indices `queries` (int64 tensor on the engine's device, identical on every rank):
        every rank gathers the rows it owns, one sum all-reduce combines them (the other ranks contribute zeros), and the
        top-level siblings come from the replicated top tree.  Returns (rows [Q, width], siblings [Q, depth, 4]) on every rank."""
        return open_trees([(self, queries)])[0]


def open_trees(pairs):
    """proofQueries (fri.js:83-105) over several sharded trees with ONE collective: every (tree, queries) pair gathers into its
    slice of a single flat buffer, one sum all-reduce combines all of them.  Returns [(rows, siblings)] in order."""
    if not pairs:
        return []
    t0 = pairs[0][0]
    sizes = [t.open_words(int(q.numel())) for t, q in pairs]
    flat = t0.e.empty(max(1, sum(sizes)))
    off = 0
    views = []
    group = None
    for (t, q), n in zip(pairs, sizes):
        v = flat[off:off + n]
        t.open_local(q, v)
        if t.replica is not None:
            if t.replica[0] != 0:
                v.zero_()              # a replicated tree: rank 0 alone contributes to the sum
            if t.replica[1] > 1:
                group = t.replica[2]
        elif t.world > 1:
            group = t.dist
        views.append(v)
        off += n
    if group is not None:
        group.all_reduce(flat)
    return [t.open_finish(q, v) for (t, q), v in zip(pairs, views)]


class ShardedCommit:
    """Column-sharded LDE -> all-to-all -> row-sharded hashing -> gathered tree top."""

    def __init__(self, engine, dist, rank, world):
        self.e, self.dist, self.rank, self.world = engine, dist, rank, world

    def shard_cols(self, cols):
        if cols % self.world:
            raise ValueError(f"nPols ({cols}) must be divisible by the number of GPUs ({self.world})")
        cg = cols // self.world
        if self.world > 1 and cg % 8:
            raise ValueError("columns per GPU must be a multiple of 8 (sponge chunks must not straddle tiles)")
        return cg

    def buffers(self, cols, n_bits, ext_bits):
        cg = self.shard_cols(cols)
        rows_local = (1 << ext_bits) // self.world
        e = self.e
        buf = {"dst": e.empty(cg << ext_bits), "nodes": e.empty(e.nnodes(rows_local)),
               "top": e.empty(max(8, e.nnodes(self.world))), "sub": e.empty(4 * self.world), "exchange": None}
        if self.world > 1 and hasattr(e, "open_exchange"):
            buf["exchange"] = e.open_exchange(self.dist, self.rank, self.world, cg << ext_bits)
        buf["recv"] = buf["exchange"]["recv"] if buf["exchange"] else e.empty(cg << ext_bits)
        buf["root"] = e.empty(4)
        return buf

    def release(self, buf):
        if buf.get("exchange"):
            self.e.close_exchange(buf["exchange"])
            buf["exchange"] = None

    def exchange_kind(self, buf):
        if not buf.get("exchange"):
            return "NCCL all_to_all_single"
        if buf["exchange"].get("shard") is not None:
            return "peer stores fused into the last LDE pass (CUDA IPC over NVLink), flag barriers + sub-roots over peer mailboxes (no NCCL in the commit)"
        return "peer stores fused into the last LDE pass (CUDA IPC over NVLink)"

    def commit(self, src_slab, cols, n_bits, ext_bits, buf, split=False):
        """src_slab: this rank's N x C/G column slab (row-major, on the device).  Returns the 4-word root tensor (on every rank)."""
        tiles = self.extend_exchange(src_slab, cols, n_bits, ext_bits, buf)
        return self.hash_and_root(tiles, cols, ext_bits, buf, split)

    def extend_exchange(self, src_slab, cols, n_bits, ext_bits, buf, host_src=False):
        """LDE of the local column slab + the column -> row exchange.  Returns the tensor holding this rank's rows as G column
        tiles.  host_src: src_slab is a pinned HOST tensor, uploaded in sub-slabs overlapped with the LDE (peer exchange only)."""
        G, e = self.world, self.e
        cg = self.shard_cols(cols)
        if (1 << ext_bits) % G:
            raise ValueError("extended height must be divisible by the number of GPUs")
        ex = buf.get("exchange")
        if host_src and not (G > 1 and ex):
            raise ValueError("host-source extension needs the peer exchange")
        if G > 1 and ex:
            # Barrier BEFORE the stores: whatever the peers still do with their receive buffers (hashing of the previous
            # commit, a download) is stream-ordered before their signal.
            self._barrier(ex)
            if host_src:
                e.lde_scatter_host(src_slab, cg, n_bits, ext_bits, ex)
            else:
                e.lde_scatter(src_slab, cg, n_bits, ext_bits, buf["dst"], ex)
            self._barrier(ex)                                               # every rank's stores have landed (stream-ordered)
            return buf["recv"]
        e.lde(src_slab, cg, n_bits, ext_bits, buf["dst"])
        if G > 1:
            self.dist.all_to_all_single(buf["recv"], buf["dst"])            # equal splits: chunk h = rows of rank h
            return buf["recv"]
        return buf["dst"]

    def _barrier(self, ex):
        if hasattr(self.e, "exchange_barrier"):
            self.e.exchange_barrier(ex)                                     # flag barrier on the stream (C ABI)
        else:
            self.dist.all_reduce(ex["flag"])                                # stand-in engines: a one-word all-reduce

    def hash_and_root(self, tiles, cols, ext_bits, buf, split=False):
        G, e = self.world, self.e
        cg = self.shard_cols(cols)
        rows_local = (1 << ext_bits) // G
        ex = buf.get("exchange")
        if G > 1 and ex and ex.get("shard") is not None and tiles is buf["recv"]:
            # the whole second half in the library: hashing, sub-roots stored into every mailbox, flag barrier, top tree
            e.shard_hash(ex, cols, ext_bits, split, buf["nodes"], buf["root"])
            buf["sub"], buf["top"] = ex["sub"], ex["top"]                   # the gathered sub-roots / top tree live in the mailbox
            buf["tree"] = ShardedTree(e, self.dist, self.rank, G, tiles, G, cg, rows_local, buf["nodes"], ex["sub"], ex["top"])
            return buf["root"]
        e.merkelize_tiled(tiles, G, cg, rows_local, buf["nodes"], split)
        buf["tree"] = ShardedTree(e, self.dist, self.rank, G, tiles, G, cg, rows_local, buf["nodes"], buf["sub"], buf["top"])
        return buf["tree"].reduce_to_root()

    def commit_rows(self, rows, width, height, nodes, sub, top):
        """Layer tree over rows every rank already holds in full (FRI layers, fri.js:63-71): rank h hashes leaves
        [h*height/G, (h+1)*height/G) and the sub-roots are gathered.  Returns (ShardedTree, root tensor)."""
        G, e = self.world, self.e
        cnt = height // G
        mine = rows[self.rank * cnt * width:(self.rank + 1) * cnt * width]
        e.merkelize(mine, width, cnt, nodes)
        t = ShardedTree(e, self.dist, self.rank, G, mine, 1, width, cnt, nodes, sub, top)
        return t, t.reduce_to_root()


class ShardedFri:
    """FRI chain (fri.js:22-81 per step) over G ranks that all hold the FRI polynomial.  Layer tree s (over the transposed rows of
    P_s, fri.js:187-202) is SHARDED while it is big enough to be worth a collective (>= min_rows_per_rank leaves per rank): rank h
    computes rows [h*R, (h+1)*R) of the layer straight from the previous layer (pil2gpu_fri_fold_range_dev: the outputs a row needs
    are exactly that row), hashes them, and the rows of the layer are all-gathered so that every rank can fold the next one -- the
    transposed rows of layer s hold the 2^k inputs of every output of fold s+1 contiguously, so the chain never returns to the
    polynomial order.  Smaller layers are computed redundantly on every rank (no collective).  Roots, trees and the final
    polynomial equal the single-process chain."""

    def __init__(self, engine, dist, rank, world, steps, min_rows_per_rank=1024):
        self.e, self.dist, self.rank, self.world, self.steps = engine, dist, rank, world, list(steps)
        self.nl = len(steps) - 1
        e = engine
        self.sharded = [world > 1 and ((1 << steps[s + 1]) // world) >= min_rows_per_rank and ((1 << steps[s + 1]) // world) % 32 == 0
                        for s in range(self.nl)]
        self.rows = [e.empty(3 << steps[s]) for s in range(self.nl)]
        self.width = [3 << (steps[s] - steps[s + 1]) for s in range(self.nl)]
        self.height = [1 << steps[s + 1] for s in range(self.nl)]
        self.nodes, self.sub, self.top, self.trees = [], [], [], [None] * self.nl
        for s in range(self.nl):
            if self.sharded[s]:
                self.nodes.append(e.empty(e.nnodes(self.height[s] // world)))
                self.sub.append(e.empty(4 * world))
                self.top.append(e.empty(max(8, e.nnodes(world))))
            else:
                self.nodes.append(e.empty(e.nnodes(self.height[s])))
                self.sub.append(None)
                self.top.append(None)
        self.final = e.empty(3 << steps[-1])
        self.roots = [None] * self.nl

    def run(self, pol0, challenges):
        """pol0: 3 * 2^steps[0] words (on every rank); challenges[s]: the F3 challenge of fold s (challenges[0] is unused: step
        0 is the identity fold, fri.js:48-49).  Fills rows / trees / roots / final."""
        e, G, st = self.e, self.world, self.steps
        for s in range(self.nl):
            # layer s holds P_s = fold_s(P_{s-1}) (P_0 = pol0, fold 0 is the identity).  Input of fold s: pol0 for s <= 1, the
            # (complete) rows of layer s-1 afterwards -- row g of layer s-1 is the input group of output g
            if s <= 1:
                src, layout, prev_bits = pol0, 0, st[0]
            else:
                src, layout, prev_bits = self.rows[s - 1], 1, st[s - 1]
            h, w = self.height[s], self.width[s]
            if self.sharded[s]:
                R = h // G
                e.fri_fold_range(src, layout, prev_bits, st[s], st[s + 1], st[0], challenges[s], self.rank * R, R, None, self.rows[s])
                mine = self.rows[s][self.rank * R * w:(self.rank + 1) * R * w]
                e.merkelize(mine, w, R, self.nodes[s])
                t = ShardedTree(e, self.dist, self.rank, G, mine, 1, w, R, self.nodes[s], self.sub[s], self.top[s])
                self.roots[s] = t.reduce_to_root()
                self.trees[s] = t
                if s >= 1:                                  # the next fold reads every row of this layer (layer 0's successor reads pol0)
                    self.dist.all_gather_into_tensor(self.rows[s], mine if str(getattr(e, "device", "")) != "cpu" else mine.clone())
            else:
                e.fri_fold_range(src, layout, prev_bits, st[s], st[s + 1], st[0], challenges[s], 0, 0, None, self.rows[s])
                e.merkelize(self.rows[s], w, h, self.nodes[s])
                nn = e.nnodes(h)
                self.roots[s] = self.nodes[s][nn - 4:nn]
                self.trees[s] = ShardedTree(e, self.dist, 0, 1, self.rows[s], 1, w, h, self.nodes[s], None, None)
                self.trees[s].replica = (self.rank, G, self.dist)
        # last fold: P_{L-1} -> P_L (the final polynomial), no tree (fri.js:72-79)
        L = len(st) - 1
        if L == 0:
            self.final.copy_(pol0)
        elif L == 1:
            e.fri_fold_range(pol0, 0, st[0], st[1], None, st[0], challenges[1], 0, 0, self.final, None)
        else:
            e.fri_fold_range(self.rows[L - 1], 1, st[L - 1], st[L], None, st[0], challenges[L], 0, 0, self.final, None)
        return self.roots, self.final

    def query_pairs(self, queries):
        """(tree, indices) of every layer tree for proofQueries (fri.js:96-104: layer s is opened at q mod 2^steps[s+1])."""
        return [(self.trees[s], queries % self.height[s]) for s in range(self.nl)]


def assemble_nodes(local_nodes_per_rank, top_nodes, rows_local, world, nnodes_fn):
    """Host-side helper (tests / downloads): stitch per-rank subtree node arrays and the top tree into the reference
    `nodes` layout of the full tree (power-of-two sizes)."""
    E = rows_local * world
    out = np.zeros(nnodes_fn(E), dtype=np.uint64)
    # levels of the local subtrees: level l has rows_local >> l nodes per rank, stored contiguously per rank
    off_local, off_global, n = 0, 0, rows_local
    while n >= 1:
        for r in range(world):
            out[off_global + r * n * 4: off_global + (r + 1) * n * 4] = local_nodes_per_rank[r][off_local:off_local + n * 4]
        if n == 1:
            break
        off_local += n * 4
        off_global += n * world * 4
        n >>= 1
    # levels above the sub-roots come from the top tree (its leaf level is the sub-root level just written)
    top_off, m = 0, world
    while m > 1:
        top_off += m * 4
        off_global += m * 4
        m >>= 1
        out[off_global:off_global + m * 4] = top_nodes[top_off:top_off + m * 4]
    return out


# ------------------------------------------------------------------------------------------------------------------
# bench.py --gpus N entry (launched under torchrun, one rank per GPU)
# ------------------------------------------------------------------------------------------------------------------
def bench_main(args, rank, world, local_rank, dist, bench):
    import json
    import torch
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = GpuEngine(torch, local_rank)
    L, check, vp = eng.L, eng.check, ctypes.c_void_p
    n_bits, cols, blow = bench.WORKLOADS[args.workload]
    ext_bits = n_bits + blow
    sc = ShardedCommit(eng, dist, rank, world)
    cg = sc.shard_cols(cols)
    seed = 0x5EED0000 + 3
    src = eng.empty(cg << n_bits)
    check(L.pil2gpu_synth2d_dev(eng.h, vp(src.data_ptr()), 1 << n_bits, cg, cols, rank * cg, seed))
    buf = sc.buffers(cols, n_bits, ext_bits)
    # FRI chain: every rank holds the FRI polynomial; big layers are sharded by rows (fold + leaf hashing of the
    # rank's own rows, rows all-gathered for the next fold), small ones are computed redundantly on every rank (ShardedFri)
    steps = bench.fri_steps(ext_bits)
    nl = len(steps) - 1                                   # layer trees
    h0 = 1 << steps[1] if nl else 1
    chal = [np.ascontiguousarray(bench.splitmix_field(seed + 2 + s, 0, 3)) for s in range(len(steps))]
    fri_pol0 = eng.empty(3 << steps[0])
    check(L.pil2gpu_synth_dev(eng.h, vp(fri_pol0.data_ptr()), 3 << steps[0], seed + 1, 0))
    sfri = ShardedFri(eng, dist, rank, world, steps)
    shard_l0 = bool(nl and sfri.sharded[0])
    npp = lambda a: vp(a.ctypes.data)
    rng = np.random.default_rng(7)
    queries_np = rng.integers(0, 1 << ext_bits, size=bench.N_QUERIES, dtype=np.int64)
    queries = torch.from_numpy(queries_np).to(eng.device)
    opened = {}

    def fri_chain():
        sfri.run(fri_pol0, chal)

    def open_queries():
        # proofQueries (fri.js:83-105): every tree is opened where its rows live, ONE sum all-reduce combines all of them
        res = open_trees([(buf["tree"], queries)] + sfri.query_pairs(queries))
        opened["main"], opened["fri"] = res[0], res[1:]

    root_host = torch.empty(4, dtype=torch.int64, pin_memory=True)

    def step():
        root = sc.commit(src, cols, n_bits, ext_bits, buf)
        fri_chain()
        open_queries()
        root_host.copy_(root, non_blocking=True)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([eng.launches() - l0], device="cuda")
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    # ---- per-phase device times of one more step (CUDA events on the launching stream, max over ranks; outside the headline
    # region) and the dominant kernel's own launch time on this rank's share of the rows ----
    def timed(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        return out, (a, b)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    tiles, ev_lde = timed(lambda: sc.extend_exchange(src, cols, n_bits, ext_bits, buf))
    rows_local = (1 << ext_bits) // world
    _, ev_hash = timed(lambda: eng.merkelize_tiled(tiles, world, cg, rows_local, buf["nodes"]))
    _, ev_tree = timed(lambda: eng.tree_from_digests(buf["nodes"], rows_local))      # the levels again: leaf kernel = hash - tree
    buf["tree"] = ShardedTree(eng, dist, rank, world, tiles, world, cg, rows_local, buf["nodes"], buf["sub"], buf["top"])
    _, ev_top = timed(lambda: buf["tree"].reduce_to_root())
    _, ev_fri = timed(fri_chain)
    _, ev_q = timed(open_queries)
    torch.cuda.synchronize()
    ph = torch.tensor([a.elapsed_time(b) / 1e3 for a, b in (ev_lde, ev_hash, ev_tree, ev_top, ev_fri, ev_q)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    t_lde, t_hash, t_tree, t_top, t_fri, t_q = (float(x) for x in ph.tolist())
    phases = {"lde+exchange": t_lde, "hash (leaves + local subtree)": t_hash, "local subtree levels": t_tree, "sub-root gather + top tree": t_top,
              "fri (big layers sharded, small ones replicated)": t_fri, "queries": t_q, "sum": t_lde + t_hash + t_top + t_fri + t_q,
              "how": "one extra step after the timed region, CUDA events per phase, max over ranks per phase"}
    mm, iw = ctypes.c_double(), ctypes.c_double()
    check(L.pil2gpu_bench_int_pipes(eng.h, ctypes.byref(mm), ctypes.byref(iw)))
    pk = torch.tensor([mm.value, iw.value], device="cuda", dtype=torch.float64)
    dist.all_reduce(pk, op=dist.ReduceOp.MIN)
    mm_v, iw_v = (float(x) for x in pk.tolist())

    # ---- the rows next to the commit over the row-sharded buffer: evaluations at xi and the FRI polynomial for every
    # column at two openings, timed as whole calls (max over ranks of the wall clock between synchronisations) ----
    extras = None
    if not getattr(args, "no_extras", False):
        import time
        xi = np.ascontiguousarray(bench.splitmix_field(seed + 10, 0, 3))
        vf1, vf2 = (np.ascontiguousarray(bench.splitmix_field(seed + 11 + i, 0, 3)) for i in range(2))
        ev_map = [("t", c, 1, o) for o in (0, 1) for c in range(cols)]
        trees = {"t": buf["tree"]}

        def wall(fn, reps=2):
            fn()                                           # warm-up (pool growth, first-launch costs)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = fn()
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / reps], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item()), out
        # These rows are reported next to the headline, never instead of it: a failure here (on every rank alike -- the inputs
        # are identical) is recorded in the line and must not cost the commit numbers measured above.
        try:
            t_ev, evs = wall(lambda: sharded_evals(eng, dist, rank, world, trees, ev_map, xi, [0, 1], n_bits, ext_bits))
            t_fp, _ = wall(lambda: sharded_fri_pol(eng, dist, rank, world, trees, ev_map, evs, xi, [0, 1], vf1, vf2, n_bits, ext_bits))
            qd, qg = 3, 1 << blow
            q_ext = eng.empty(qd << ext_bits)
            check(L.pil2gpu_synth_dev(eng.h, vp(q_ext.data_ptr()), qd << ext_bits, seed + 9, 0))
            cmq, qn = eng.empty((qd * qg) << ext_bits), eng.empty(eng.nnodes((1 << ext_bits) // world))
            qs, qt = eng.empty(4 * world), eng.empty(max(8, eng.nnodes(world)))
            t_q, _ = wall(lambda: sharded_compute_q(sc, q_ext, qd, qg, n_bits, ext_bits, cmq, qn, qs, qt))
            del q_ext, cmq, qn
            extras = {"q_commit": {"s": t_q, "shape": f"qDim {qd}, qDeg {qg}, 2^{ext_bits} rows: transforms on every rank, tree hashed sharded over {world} ranks"},
                      "evals": {"s": t_ev, "shape": f"{len(ev_map)} evaluations, base rows sharded over {world} ranks (LEv vectors + sums + gather)"},
                      "fri_pol": {"s": t_fp, "shape": f"{len(ev_map)} evMap terms, 2^{ext_bits} rows sharded over {world} ranks (xDivXSubXi + friExp + "
                                                      "all-gather)"}}
        except (ValueError, RuntimeError) as ex:
            extras = {"error": str(ex)[:300]}

    # ---- e2e: the same sharded commit with HOST buffers: every rank uploads its column slab from pinned memory and
    # downloads its share of the extended rows and of the nodes; rank 0 also moves the FRI polynomial and layers ----
    e2e = None
    if not args.no_e2e:
        import time
        pin = lambda t: torch.empty(t.numel(), dtype=torch.int64, pin_memory=True)
        src_host = pin(src)
        src_host.copy_(src)
        ext_dev = buf["recv"] if world > 1 else buf["dst"]
        out_host, nodes_host = pin(ext_dev), pin(buf["nodes"])
        # FRI: the polynomial goes up on every rank (every rank folds its share of the big layers); sharded layers come down from
        # the ranks that own them (rows + subtree nodes), replicated ones and the final polynomial from rank 0
        fri_host = {"pol0": pin(fri_pol0)}
        fri_host["pol0"].copy_(fri_pol0)
        fri_down = []
        for s_ in range(nl):
            if sfri.sharded[s_]:
                R_ = sfri.height[s_] // world
                fri_down += [sfri.rows[s_][rank * R_ * sfri.width[s_]:(rank + 1) * R_ * sfri.width[s_]], sfri.nodes[s_]]
            elif rank == 0:
                fri_down += [sfri.rows[s_], sfri.nodes[s_]]
        if rank == 0:
            fri_down.append(sfri.final)
        fri_host["down"] = [pin(t) for t in fri_down]
        torch.cuda.synchronize()

        copy_stream = torch.cuda.Stream()
        peer = buf.get("exchange") is not None

        import os
        trace = os.environ.get("PIL2GPU_TRACE") is not None
        marks = []

        def mark(name, stream=None):
            if trace:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream or torch.cuda.current_stream())
                marks.append((name, ev))

        def e2e_step():
            del marks[:]
            mark("start")
            if peer:
                # pinned slab -> sub-slab uploads overlapped with the LDE + peer stores; then the download of this rank's
                # extended rows (copy stream) overlaps their hashing
                tiles = sc.extend_exchange(src_host, cols, n_bits, ext_bits, buf, host_src=True)
                mark("lde+exchange")
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    out_host.copy_(tiles, non_blocking=True)
                    mark("rows down (copy stream)", copy_stream)
                root = sc.hash_and_root(tiles, cols, ext_bits, buf)
                mark("hash+root")
            else:
                src.copy_(src_host, non_blocking=True)
                root = sc.commit(src, cols, n_bits, ext_bits, buf)
                mark("commit")
                out_host.copy_(ext_dev, non_blocking=True)
                mark("rows down")
            # the FRI chain first: its kernels overlap the row download, whereas a device->host copy issued here would queue
            # behind those 32/G GiB on the copy engine and hold the chain back
            fri_pol0.copy_(fri_host["pol0"], non_blocking=True)
            fri_chain()
            mark("fri")
            nodes_host.copy_(buf["nodes"], non_blocking=True)
            mark("nodes down")
            for d, h in zip(fri_down, fri_host["down"]):
                h.copy_(d, non_blocking=True)
            root_host.copy_(root, non_blocking=True)
            mark("fri down")
            torch.cuda.synchronize()
            if trace:
                import sys as _sys
                print(f"[pil2gpu] rank {rank} e2e: " + ", ".join(f"{n} {marks[0][1].elapsed_time(ev):.1f} ms" for n, ev in marks[1:]), file=_sys.stderr, flush=True)

        e2e_step()
        n_e2e = max(1, min(args.steps, 3))
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = 8 * (cols << n_bits) + 8 * (3 << steps[0]) * world
        d2h = 8 * (cols << ext_bits) + 8 * world * buf["nodes"].numel() + 32 + 8 * (3 << steps[-1])
        for s_ in range(nl):              # rows + nodes of every layer, from wherever they live
            d2h += 8 * (3 << steps[s_]) + 8 * (world * eng.nnodes(sfri.height[s_] // world) if sfri.sharded[s_] else eng.nnodes(sfri.height[s_]))
        e2e = {"value": float(dt.item()), "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "call": ("per rank: pinned host slab -> pil2gpu_lde_scatter (sub-slab uploads overlapped with the LDE + peer stores) -> hashing "
                        "overlapped with the download of the extended rows; nodes -> pinned host" if peer else
                        "per rank: pinned host slab -> device, ShardedCommit.commit, extended rows + nodes -> pinned host") +
                       "; FRI polynomial up on every rank, sharded layers down from their owners, replicated layers from rank 0"}
    if rank == 0:
        clocks = sampler.stop()
        sec = float(ms.item()) / 1e3 / args.steps
        root = [int(x) & 0xFFFFFFFFFFFFFFFF for x in root_host.tolist()]
        a2a = 8 * (cg << ext_bits) * (world - 1) // world
        peaks = {}
        try:
            import os as _os
            peaks = json.load(open(_os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        t_leaf = max(t_hash - t_tree, 1e-9)
        roofline = bench.leaf_roofline(rows_local * ((cols + 7) // 8), 8 * cols * rows_local + 32 * rows_local, t_leaf, sec, mm_v, iw_v, hbm_peak,
                                       "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md, 6.65 TB/s)",
                                       None)
        roofline["kernel"] = f"merkle_leaf_kernel (per rank: {rows_local} of the 2^{ext_bits} rows, slowest rank)"
        roofline["note"] = (f"between the LDE and the hashing every GPU moves {a2a >> 20} MiB over NVLink ({sc.exchange_kind(buf)}); "
                            "ncu traffic is captured at N=1 only (profiles/)")
        line = {
            "metric": bench.METRIC, "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": bench.config_dict(args.workload, world),
            "rows_per_s": (1 << n_bits) / sec, "all_to_all_bytes_per_gpu": a2a, "exchange": sc.exchange_kind(buf),
            "fri": f"layers sharded by rows: {[bool(x) for x in sfri.sharded]} (the others replicated on every rank); {bench.N_QUERIES} queries "
                   "opened on the owning ranks, all trees combined with ONE all-reduce",
            "gpu_launches": int(launches.item()), "clocks": clocks,
            "root": root,
            "e2e": e2e, "next_rows": extras, "cpu_baseline": None,
            "phases_s": phases,
            "roofline": roofline,
            "fri_roots": [[int(x) & 0xFFFFFFFFFFFFFFFF for x in r.tolist()] for r in sfri.roots],
            "fri_final0": [int(x) & 0xFFFFFFFFFFFFFFFF for x in sfri.final[:3].tolist()],
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    sc.release(buf)
    dist.destroy_process_group()
