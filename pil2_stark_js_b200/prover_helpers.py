"""Mirror of the expression evaluators of src/prover/prover_helpers.js: callCalculateExps / calculateExps (:23-76) over the domains
"n" and "ext", with the reference's program format ({op, dest, src} records, operand references as in getRef / setRef :112-219).

The reference compiles a program into the body of a JavaScript function (compileCode :87-110) and calls it once per row; here
`compile_code` turns it into the fixed records of pil2gpu_calculate_exps_dev (include/pil2gpu.h) -- operands resolved to (device
buffer, column, row offset) / uniform-constant / temporary-slot triples, temporaries packed into slots by liveness -- and the GPU
interprets it with one thread per row.  `ctx` carries the reference's field names (pilInfo, nBits, nBitsExt, challenges, publics,
evals, subproofValues, const_n / const_ext / cm<stage>_n / cm<stage>_ext, Zi_ext, xDivXSubXi_ext, q_ext, f_ext) as numpy uint64
arrays, plus `gpu` (a Context); with `dev_buffers` (name -> (device pointer, words)) the named buffers are used in place on the
device and nothing is uploaded or downloaded for them (the way the committed stage buffers stay in HBM between stages)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check
from .context import default_context, DeviceBuffer

P = 0xFFFFFFFF00000001
OPCODES = {"add": 0, "sub": 1, "mul": 2, "copy": 3, "muladd": 4}
K_TMP, K_CONST, K_BUF, K_X = 0, 1, 2, 3
OP_WORDS = 16
MAX_SLOTS = 64


def _f3(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return [int(x) % P for x in v], 3
    return [int(v) % P, 0, 0], 1


def zi_index(info, boundary_id):
    """prover_helpers.js:201-215"""
    b = info["boundaries"][boundary_id]
    for k, o in enumerate(info["boundaries"]):
        if b["name"] == "everyFrame":
            if o["name"] == "everyFrame" and o.get("offsetMin") == b.get("offsetMin") and o.get("offsetMax") == b.get("offsetMax"):
                return k
        elif o["name"] == b["name"]:
            return k
    raise ValueError("Something went wrong")


class CompiledCode:
    """ops: uint32 [n_ops * 16]; consts: uint64 [n_consts * 3]; buffers: [(name, row_words)] in buffer-index order; written: names
    of the buffers the program stores to; n_slots: temporaries alive at once."""

    def __init__(self, ops, consts, buffers, written, n_slots, dom):
        self.ops, self.consts, self.buffers, self.written, self.n_slots, self.dom = ops, consts, buffers, written, n_slots, dom


def compile_code(ctx, code, dom):
    """compileCode (prover_helpers.js:87-110) for the GPU interpreter."""
    if dom not in ("n", "ext"):
        raise ValueError("Invalid dom")
    info = ctx.pilInfo
    n_bits, ext_bits = ctx.nBits, ctx.nBitsExt
    N = 1 << (n_bits if dom == "n" else ext_bits)
    extend_bits = ext_bits - n_bits
    consts, const_index = [], {}
    buffers, buf_index, written = [], {}, set()

    def const_of(value):
        v, dim = _f3(value)
        key = (tuple(v), dim)
        if key not in const_index:
            const_index[key] = len(consts)
            consts.append(v)
        return const_index[key], dim

    def buf_of(name, row_words):
        if name not in buf_index:
            buf_index[name] = len(buffers)
            buffers.append((name, int(row_words)))
        return buf_index[name]

    def row_offset(prime):
        if not prime:
            return 0
        nxt = prime if dom == "n" else prime << extend_bits            # getRef :158-166 / evalMap :227-233, reduced mod N
        return nxt % N

    def pol_ref(pol_id):
        p = info["cmPolsMap"][pol_id]
        st = "cm%d" % p["stage"]
        return st + "_" + dom, p["stagePos"], info["mapSectionsN"][st], p["dim"]

    # validation first (on the program as given: dead records are checked too)
    seen = set()
    for c in code:
        if c["op"] not in OPCODES:
            raise ValueError("Invalid op:" + str(c["op"]))
        nsrc = {"copy": 1, "muladd": 3}.get(c["op"], 2)
        if len(c["src"]) != nsrc:
            raise ValueError("op %s takes %d sources" % (c["op"], nsrc))
        for s in c["src"]:
            if s["type"] == "tmp" and s["id"] not in seen:
                raise ValueError("temporary %d read before it is written" % s["id"])
        if c["dest"]["type"] == "tmp":
            seen.add(c["dest"]["id"])
    # dead records: a temporary nobody reads (the code generator leaves some behind: 23 of the 169 records of the sm_all quotient
    # program) -- dropped, transitively, before slots are assigned
    code = list(code)
    while True:
        read = {s["id"] for c in code for s in c["src"] if s["type"] == "tmp"}
        live = [c for c in code if c["dest"]["type"] != "tmp" or c["dest"]["id"] in read]
        if len(live) == len(code):
            break
        code = live
    # liveness of the temporaries: last record that reads each id
    last_use = {}
    for k, c in enumerate(code):
        for s in c["src"]:
            if s["type"] == "tmp":
                last_use[s["id"]] = k
    slot_of, free, n_slots = {}, [], 0
    tmp_dim = {}

    def operand(r, k, is_dest=False):
        t = r["type"]
        if t == "tmp":
            if is_dest:
                if r["id"] in slot_of:                                  # re-assignment of a live temporary keeps its slot
                    slot = slot_of[r["id"]]
                else:
                    nonlocal n_slots
                    slot = free.pop() if free else n_slots
                    if slot == n_slots:
                        n_slots += 1
                    slot_of[r["id"]] = slot
                tmp_dim[r["id"]] = r.get("dim", 3)
                return [K_TMP | (tmp_dim[r["id"]] << 8), slot, 0]
            if r["id"] not in slot_of:
                raise ValueError("temporary %d read before it is written" % r["id"])
            return [K_TMP | (tmp_dim[r["id"]] << 8), slot_of[r["id"]], 0]
        if t == "const":
            b = buf_of("const_" + dom, info["nConstants"])
            return [K_BUF | (1 << 8) | (b << 16), r["id"], row_offset(r.get("prime", 0))]
        if t == "cm":
            name, off, size, dim = pol_ref(r["id"])
            b = buf_of(name, size)
            if is_dest:
                written.add(name)
            return [K_BUF | (dim << 8) | (b << 16), off, row_offset(r.get("prime", 0))]
        if t in ("q", "f"):
            if not is_dest or dom != "ext":
                raise ValueError("Accessing %s in domain %s" % (t, dom))
            dim = 3 if t == "f" else r.get("dim", 3)
            b = buf_of(t + "_ext", dim)
            written.add(t + "_ext")
            return [K_BUF | (dim << 8) | (b << 16), 0, 0]
        if is_dest:
            raise ValueError("Invalid reference type set: " + str(t))
        if t == "number":
            i, dim = const_of(int(r["value"]))
        elif t == "public":
            i, dim = const_of(int(ctx.publics[r["id"]]))
        elif t == "challenge":
            i, dim = const_of(ctx.challenges[r["stage"] - 1][r["stageId"]])
        elif t == "subproofValue":
            i, dim = const_of(ctx.subproofValues[r["id"]])
        elif t == "eval":
            i, dim = const_of(ctx.evals[r["id"]])
        elif t == "xDivXSubXi":
            b = buf_of("xDivXSubXi_ext", 3 * len(info["openingPoints"]))
            return [K_BUF | (3 << 8) | (b << 16), 3 * r["id"], 0]
        elif t == "x":
            return [K_X | (1 << 8), 0, 0]
        elif t == "Zi":
            # Zi_ext is [boundary][row] (prover_helpers.js:201-215): one single-column buffer per boundary table
            b = buf_of("Zi_ext:%d" % zi_index(info, r["boundaryId"]), 1)
            return [K_BUF | (1 << 8) | (b << 16), 0, 0]
        else:
            raise ValueError("Invalid reference type get: " + str(t))
        return [K_CONST | (dim << 8), i, 0]

    ops = np.zeros(len(code) * OP_WORDS, dtype=np.uint32)
    for k, c in enumerate(code):
        if c["op"] not in OPCODES:
            raise ValueError("Invalid op:" + str(c["op"]))
        nsrc = {"copy": 1, "muladd": 3}.get(c["op"], 2)
        if len(c["src"]) != nsrc:
            raise ValueError("op %s takes %d sources" % (c["op"], nsrc))
        rec = ops[k * OP_WORDS:(k + 1) * OP_WORDS]
        rec[0], rec[1] = OPCODES[c["op"]], nsrc
        srcs = [operand(s, k) for s in c["src"]]
        for s in c["src"]:                                              # slots whose last reader is this record are free again
            if s["type"] == "tmp" and last_use.get(s["id"]) == k and s["id"] in slot_of:
                free.append(slot_of.pop(s["id"]))
        rec[4:7] = operand(c["dest"], k, is_dest=True)
        for j, s in enumerate(srcs):
            rec[7 + 3 * j:10 + 3 * j] = s
        if c["dest"]["type"] == "tmp" and c["dest"]["id"] not in last_use:
            pass                                                        # a temporary nobody reads (the returned value): keeps its slot
    if n_slots > MAX_SLOTS:
        raise ValueError("the program keeps %d temporaries alive at once (at most %d)" % (n_slots, MAX_SLOTS))
    carr = np.array(consts, dtype=np.uint64).reshape(-1) if consts else np.zeros(0, dtype=np.uint64)
    return CompiledCode(ops, carr, buffers, written, n_slots, dom)


def _host_buffer(ctx, name):
    if name.startswith("Zi_ext:"):
        k = int(name.split(":")[1])
        E = 1 << ctx.nBitsExt
        return np.ascontiguousarray(np.asarray(ctx.Zi_ext, dtype=np.uint64)[k * E:(k + 1) * E])
    return getattr(ctx, name)


def calculateExps(ctx, code, dom, debug=False, ret=False, global_=False):
    """calculateExps (prover_helpers.js:33-76) with ret == false: runs `code["code"]` (or a bare list of records) at every row of
    the domain and updates the destination buffers of ctx.  Buffers named in ctx.dev_buffers stay on the device; the others are
    uploaded, and the ones the program writes are downloaded back into the ctx arrays."""
    if debug or ret:
        raise NotImplementedError("the debug / ret modes of calculateExps return per-row JavaScript values; use the CPU reference for them")
    records = code["code"] if isinstance(code, dict) else code
    g = getattr(ctx, "gpu", None) or default_context()
    cc = compile_code(ctx, records, dom)
    dev = getattr(ctx, "dev_buffers", None) or {}
    rows = 1 << (ctx.nBits if dom == "n" else ctx.nBitsExt)
    owned, arr = [], (_lib.ExprBuffer * max(1, len(cc.buffers)))()
    try:
        for i, (name, row_words) in enumerate(cc.buffers):
            if name in dev:
                ptr = dev[name][0] if isinstance(dev[name], (tuple, list)) else dev[name]
                ptr = ptr.ptr.value if isinstance(ptr, DeviceBuffer) else int(ptr)
            else:
                h = np.ascontiguousarray(_host_buffer(ctx, name), dtype=np.uint64).reshape(-1)
                if h.size != rows * row_words:
                    raise ValueError("%s holds %d words, expected %d" % (name, h.size, rows * row_words))
                b = g.upload(h)
                owned.append((name, b))
                ptr = b.ptr.value
            arr[i] = _lib.ExprBuffer(ptr, row_words)
        check(g._L.pil2gpu_calculate_exps_dev(g.handle, ctypes.c_void_p(cc.ops.ctypes.data), cc.ops.size // OP_WORDS, ctypes.c_void_p(cc.consts.ctypes.data) if cc.consts.size else None,
                                              cc.consts.size // 3, arr, len(cc.buffers), ctx.nBits if dom == "n" else ctx.nBitsExt, 1 if dom == "ext" else 0))
        for name, b in owned:
            if name in cc.written:
                np.asarray(getattr(ctx, name)).reshape(-1)[:] = b.download()
        g.sync()
    finally:
        for _, b in owned:
            b.free()
    return cc


def callCalculateExps(stage, code, dom, ctx, parallelExec=False, useThreads=False, debug=False, global_=False):
    """callCalculateExps (prover_helpers.js:23-31): the parallel / threaded variants of the reference are scheduling choices of
    its CPU implementation; on the GPU every row is its own thread."""
    return calculateExps(ctx, code, dom, debug, False, global_)
