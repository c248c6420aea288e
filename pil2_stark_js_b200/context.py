"""GPU context: one per (device, stream).  Mirrors nothing in the reference (which spins up a workerpool per call,
fft_p.js:121 / merklehash_p.js:53); it exists so twiddle tables are built once per process."""
import ctypes
import numpy as np

from . import _lib
from ._lib import vp, check
from .bigbuffer import BigBuffer, as_pages


def _as_u64(a, name="buffer"):
    if not isinstance(a, np.ndarray) or a.dtype != np.uint64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name} must be a C-contiguous numpy uint64 array (BigUint64Array layout)")
    return a


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class Context:
    def __init__(self, device=0, stream=None):
        self._L = _lib.load()
        h = vp()
        check(self._L.pil2gpu_create(int(device), vp(stream) if stream else None, ctypes.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.pil2gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("context is closed")
        return self._h

    def sync(self):
        check(self._L.pil2gpu_sync(self.handle))

    @property
    def launch_count(self):
        return int(self._L.pil2gpu_launch_count(self.handle))

    # ---- host-buffer entry points (numpy uint64 arrays or BigBuffers in / out) ----
    def ntt(self, src, n_pols, n_bits, dst, inverse=False):
        sp, sw, sn, ssz = as_pages(src, "buffSrc")
        dp, dw, dn, dsz = as_pages(dst, "buffDst")
        if ssz != n_pols << n_bits or dsz != n_pols << n_bits:
            raise ValueError("buffer size does not match nPols * 2^nBits")
        check(self._L.pil2gpu_ntt_paged(self.handle, sp, sw, sn, dp, dw, dn, n_pols, n_bits, int(inverse)))

    def lde(self, src, n_pols, n_bits, dst, n_bits_ext):
        sp, sw, sn, ssz = as_pages(src, "buffSrc")
        dp, dw, dn, dsz = as_pages(dst, "buffDst")
        if ssz != n_pols << n_bits or dsz != n_pols << n_bits_ext:
            raise ValueError("buffer size does not match nPols * 2^nBits / 2^nBitsExt")
        check(self._L.pil2gpu_lde_paged(self.handle, sp, sw, sn, dp, dw, dn, n_pols, n_bits, n_bits_ext))

    def poseidon(self, in12):
        a = np.ascontiguousarray(in12, dtype=np.uint64)
        if a.size != 12:
            raise ValueError("poseidon state must have 12 words")
        o = np.empty(12, dtype=np.uint64)
        check(self._L.pil2gpu_poseidon(self.handle, _ptr(a), _ptr(o)))
        return o

    def linear_hash(self, vals, split=False):
        a = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1)
        o = np.empty(4, dtype=np.uint64)
        check(self._L.pil2gpu_linear_hash(self.handle, _ptr(a) if a.size else None, a.size, int(split), _ptr(o)))
        return o

    def merkle_nnodes(self, height):
        return int(self._L.pil2gpu_merkle_nnodes(height))

    def merkelize(self, elems, width, height, split=False):
        ep, ew, en, esz = as_pages(elems, "buff")
        if esz != width * height:
            raise ValueError("buffer size does not match width * height")
        nodes = np.empty(self.merkle_nnodes(height), dtype=np.uint64)
        check(self._L.pil2gpu_merkelize_paged(self.handle, ep, ew, en, width, height, int(split), _ptr(nodes)))
        return nodes

    def fri_fold(self, pol, prev_bits, cur_bits, next_bits, step0_bits, challenge, split=False):
        """pol: (2^prev, 3) uint64.  next_bits=None: last step.  Returns (pol2 (2^cur,3), rows|None, nodes|None)."""
        pol = np.ascontiguousarray(pol, dtype=np.uint64).reshape(-1)
        if pol.size != 3 << prev_bits:
            raise ValueError("Invalid polynomial size")
        ch = np.ascontiguousarray(challenge, dtype=np.uint64).reshape(-1)
        pol2 = np.empty(3 << cur_bits, dtype=np.uint64)
        rows = nodes = None
        if next_bits is not None:
            rows = np.empty(3 << cur_bits, dtype=np.uint64)
            nodes = np.empty(self.merkle_nnodes(1 << next_bits), dtype=np.uint64)
        check(self._L.pil2gpu_fri_fold(self.handle, _ptr(pol), prev_bits, cur_bits, -1 if next_bits is None else next_bits,
                                       step0_bits, _ptr(ch), int(split), _ptr(pol2),
                                       _ptr(rows) if rows is not None else None, _ptr(nodes) if nodes is not None else None))
        return pol2.reshape(-1, 3), rows, nodes

    def extend_and_merkelize(self, src, n_pols, n_bits, n_bits_ext, split=False, want_dst=True, want_nodes=True, dst=None):
        """extendAndMerkelize (stark_gen_helpers.js:388-412) with host buffers (numpy arrays or BigBuffers): returns
        (dst|None, nodes|None, root[4]); `dst` may be a caller-allocated array / BigBuffer (cm<stage>_ext) that is filled in
        place.  Wide standard-hash traces go through the column-slab pipeline (H2D | LDE + hashing | D2H overlapped)."""
        sp, sw, sn, ssz = as_pages(src, "buffSrc")
        if ssz != n_pols << n_bits:
            raise ValueError("buffer size does not match nPols * 2^nBits")
        if dst is None and want_dst:
            dst = np.empty(n_pols << n_bits_ext, dtype=np.uint64)
        if dst is not None:
            dp, dw, dn, dsz = as_pages(dst, "buffDst")
            if dsz != n_pols << n_bits_ext:
                raise ValueError("buffer size does not match nPols * 2^nBitsExt")
        else:
            dp, dw, dn = None, None, 0
        nodes = np.empty(self.merkle_nnodes(1 << n_bits_ext), dtype=np.uint64) if want_nodes else None
        root = np.empty(4, dtype=np.uint64)
        check(self._L.pil2gpu_extend_and_merkelize_paged(self.handle, sp, sw, sn, n_pols, n_bits, n_bits_ext, int(split), dp, dw, dn,
                                                         _ptr(nodes) if want_nodes else None, _ptr(root)))
        return dst, nodes, root

    # ---- quotient polynomial / evaluations (stark_gen_helpers.js:168-323) ----
    def compute_q(self, q_ext, q_dim, q_deg, n_bits, n_bits_ext, split=False, want_ext=True, want_nodes=True, ext=None):
        """computeQStark (stark_gen_helpers.js:168-208) with host buffers (numpy arrays or BigBuffers): returns
        (cmQ_ext|None, nodes|None, root[4]); `ext` may be the caller's cm<Q>_ext buffer."""
        qp, qw, qn, qsz = as_pages(q_ext, "q_ext")
        if qsz != q_dim << n_bits_ext:
            raise ValueError("buffer size does not match qDim * 2^nBitsExt")
        if ext is None and want_ext:
            ext = np.empty((q_dim * q_deg) << n_bits_ext, dtype=np.uint64)
        if ext is not None:
            cp, cw, cn, csz = as_pages(ext, "cmQ_ext")
            if csz != (q_dim * q_deg) << n_bits_ext:
                raise ValueError("buffer size does not match qDim * qDeg * 2^nBitsExt")
        else:
            cp, cw, cn = None, None, 0
        nodes = np.empty(self.merkle_nnodes(1 << n_bits_ext), dtype=np.uint64) if want_nodes else None
        root = np.empty(4, dtype=np.uint64)
        check(self._L.pil2gpu_compute_q_paged(self.handle, qp, qw, qn, q_dim, q_deg, n_bits, n_bits_ext, int(split), cp, cw, cn,
                                              _ptr(nodes) if want_nodes else None, _ptr(root)))
        return ext, nodes, root

    def compute_q_tree(self, q_ext, q_dim, q_deg, n_bits, n_bits_ext, split=False):
        """computeQStark (stark_gen_helpers.js:168-208) with the result kept in HBM: q_ext (host) is uploaded, cmQ_ext and
        its tree never leave the device.  Returns (DeviceTree, root[4])."""
        _as_u64(q_ext, "q_ext")
        if q_ext.size != q_dim << n_bits_ext:
            raise ValueError("buffer size does not match qDim * 2^nBitsExt")
        q_dev = self.upload(q_ext)
        tree = self.tree_alloc(q_dim * q_deg, 1 << n_bits_ext)
        try:
            check(self._L.pil2gpu_compute_q_dev(self.handle, q_dev.ptr, q_dim, q_deg, n_bits, n_bits_ext, vp(tree.elements_ptr)))
            check(self._L.pil2gpu_merkelize_dev(self.handle, vp(tree.elements_ptr), q_dim * q_deg, 1 << n_bits_ext, int(split),
                                                vp(tree.nodes_ptr)))
            root = tree.root()
        except Exception:
            tree.free()
            raise
        finally:
            q_dev.free()
        return tree, root

    def alloc(self, words):
        return DeviceBuffer(self, words)

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
        b = DeviceBuffer(self, a.size)
        check(self._L.pil2gpu_h2d(self.handle, b.ptr, _ptr(a), a.size * 8))
        self.sync()
        return b

    def compute_levs(self, xi_challenge, openings, n_bits):
        """LEv vectors of computeEvalsStark (stark_gen_helpers.js:216-231), one per opening point, back to back on the
        device: DeviceBuffer of len(openings) x 2^n_bits x 3 words."""
        xi = np.ascontiguousarray(xi_challenge, dtype=np.uint64).reshape(-1)
        if xi.size != 3:
            raise ValueError("the challenge is an F3 element (3 words)")
        n = 1 << n_bits
        lev = DeviceBuffer(self, len(openings) * n * 3)
        for i, o in enumerate(openings):
            check(self._L.pil2gpu_compute_lev_dev(self.handle, _ptr(xi), int(o), n_bits, vp(lev.ptr.value + 8 * i * n * 3)))
        return lev

    def compute_evals(self, buf_ptr, size, n_bits, n_bits_ext, ev_descs, levs, n_lev):
        """Evaluation loop of computeEvalsStark (:250-264) over one device-resident extended buffer.  ev_descs: list of
        (offset, dim, opening index).  Returns an (n, 3) uint64 array."""
        n = len(ev_descs)
        out = np.empty((n, 3), dtype=np.uint64)
        if n == 0:
            return out
        desc = (_lib.EvalDesc * n)(*[_lib.EvalDesc(int(o), int(d), int(l)) for o, d, l in ev_descs])
        bp = buf_ptr.ptr if isinstance(buf_ptr, DeviceBuffer) else vp(buf_ptr)
        check(self._L.pil2gpu_compute_evals_dev(self.handle, bp, size, n_bits, n_bits_ext, desc, n, levs.ptr, n_lev, _ptr(out)))
        return out

    def compute_evals_host(self, xi_challenge, openings, n_bits, n_bits_ext, buf, size, ev_descs):
        """computeEvalsStark's arithmetic (:216-267) for one HOST extended buffer: LEv vectors + evaluation sums; only the
        2^n_bits rows the loop reads are uploaded.  ev_descs: list of (offset, dim, opening index).  Returns (n, 3) uint64."""
        _as_u64(buf, "buffer")
        if buf.size != size << n_bits_ext:
            raise ValueError("buffer size does not match size * 2^nBitsExt")
        xi = np.ascontiguousarray(xi_challenge, dtype=np.uint64).reshape(-1)
        n = len(ev_descs)
        out = np.empty((n, 3), dtype=np.uint64)
        if n == 0:
            return out
        op = (ctypes.c_int32 * len(openings))(*[int(o) for o in openings])
        desc = (_lib.EvalDesc * n)(*[_lib.EvalDesc(int(o), int(d), int(l)) for o, d, l in ev_descs])
        check(self._L.pil2gpu_compute_evals(self.handle, _ptr(xi), op, len(openings), n_bits, n_bits_ext, _ptr(buf), size, desc, n, _ptr(out)))
        return out

    def x_div_x_sub_xi_host(self, xi_challenge, openings, n_bits, n_bits_ext):
        xi = np.ascontiguousarray(xi_challenge, dtype=np.uint64).reshape(-1)
        op = (ctypes.c_int32 * len(openings))(*[int(o) for o in openings])
        out = np.empty(3 * len(openings) << n_bits_ext, dtype=np.uint64)
        check(self._L.pil2gpu_x_div_x_sub_xi(self.handle, _ptr(xi), op, len(openings), n_bits, n_bits_ext, _ptr(out)))
        return out.reshape(-1, len(openings), 3)

    def x_div_x_sub_xi(self, xi_challenge, openings, n_bits, n_bits_ext, download=True):
        """xDivXSubXi_ext of computeFRIStark (:289-323): (2^n_bits_ext, nOpenings, 3) array (or the DeviceBuffer)."""
        xi = np.ascontiguousarray(xi_challenge, dtype=np.uint64).reshape(-1)
        op = (ctypes.c_int32 * len(openings))(*[int(o) for o in openings])
        out = DeviceBuffer(self, 3 * len(openings) << n_bits_ext)
        check(self._L.pil2gpu_x_div_x_sub_xi_dev(self.handle, _ptr(xi), op, len(openings), n_bits, n_bits_ext, out.ptr))
        if not download:
            return out
        return out.download().reshape(-1, len(openings), 3)

    def fri_pol(self, terms, evals, openings, xdiv, vf1, vf2, n_bits_ext, download=True):
        """f_ext of computeFRIStark (stark_gen_helpers.js:325-334; friPolinomial.js:26-56).  terms: evMap in order as
        (DeviceBuffer | device pointer, row size, column offset, dim, prime); evals: (n, 3) array-like (ctx.evals); xdiv: the
        DeviceBuffer from x_div_x_sub_xi(download=False).  Returns the (2^n_bits_ext, 3) array (or the DeviceBuffer)."""
        n = len(terms)
        arr = (_lib.FriTerm * n)()
        for i, (buf, size, offset, dim, prime) in enumerate(terms):
            ptr = buf.ptr.value if isinstance(buf, DeviceBuffer) else int(buf)
            arr[i] = _lib.FriTerm(ptr, int(size), int(offset), int(dim), int(prime))
        ev = np.ascontiguousarray(np.asarray(evals, dtype=np.uint64).reshape(-1))
        if ev.size != 3 * n:
            raise ValueError("evals must hold one F3 value per evMap entry")
        op = (ctypes.c_int32 * len(openings))(*[int(o) for o in openings])
        a, b = (np.ascontiguousarray(v, dtype=np.uint64).reshape(-1) for v in (vf1, vf2))
        out = DeviceBuffer(self, 3 << n_bits_ext)
        check(self._L.pil2gpu_fri_pol_dev(self.handle, arr, n, _ptr(ev), op, len(openings), xdiv.ptr, _ptr(a), _ptr(b), n_bits_ext, out.ptr))
        if not download:
            return out
        res = out.download().reshape(-1, 3)
        out.free()
        return res

    def fri_pol_host(self, terms, evals, openings, xi_challenge, vf1, vf2, n_bits, n_bits_ext, want_xdiv=False):
        """All of computeFRIStark's arithmetic with HOST buffers: terms as (numpy buffer, row size, offset, dim, prime)."""
        n = len(terms)
        arr = (_lib.FriTerm * n)()
        keep = []
        for i, (buf, size, offset, dim, prime) in enumerate(terms):
            _as_u64(buf, "extended buffer")
            if buf.size != size << n_bits_ext:
                raise ValueError("buffer size does not match size * 2^nBitsExt")
            keep.append(buf)
            arr[i] = _lib.FriTerm(buf.ctypes.data, int(size), int(offset), int(dim), int(prime))
        ev = np.ascontiguousarray(np.asarray(evals, dtype=np.uint64).reshape(-1))
        op = (ctypes.c_int32 * len(openings))(*[int(o) for o in openings])
        xi, a, b = (np.ascontiguousarray(v, dtype=np.uint64).reshape(-1) for v in (xi_challenge, vf1, vf2))
        f = np.empty(3 << n_bits_ext, dtype=np.uint64)
        xd = np.empty(3 * len(openings) << n_bits_ext, dtype=np.uint64) if want_xdiv else None
        check(self._L.pil2gpu_fri_pol(self.handle, arr, n, _ptr(ev), op, len(openings), _ptr(xi), _ptr(a), _ptr(b), n_bits, n_bits_ext, _ptr(f),
                                      _ptr(xd) if want_xdiv else None))
        return (f.reshape(-1, 3), xd) if want_xdiv else f.reshape(-1, 3)

    # ---- device-resident commit ----
    def commit(self, src, n_pols, n_bits, n_bits_ext, split=False):
        """interpolate + merkelize with the LDE kept in HBM.  Returns (DeviceTree, root[4])."""
        _as_u64(src, "buffSrc")
        if src.size != n_pols << n_bits:
            raise ValueError("buffer size does not match nPols * 2^nBits")
        t = vp()
        root = np.empty(4, dtype=np.uint64)
        check(self._L.pil2gpu_commit(self.handle, _ptr(src), n_pols, n_bits, n_bits_ext, int(split), ctypes.byref(t), _ptr(root)))
        return DeviceTree(self, t), root

    def commit_dev(self, src_dev_ptr, n_pols, n_bits, n_bits_ext, split=False):
        t = vp()
        root = np.empty(4, dtype=np.uint64)
        check(self._L.pil2gpu_commit_dev(self.handle, vp(src_dev_ptr), n_pols, n_bits, n_bits_ext, int(split), ctypes.byref(t),
                                         _ptr(root)))
        return DeviceTree(self, t), root

    def tree_alloc(self, width, height):
        """Empty device tree to be filled from a file (DeviceTree.fill)."""
        t = vp()
        check(self._L.pil2gpu_tree_alloc(self.handle, width, height, ctypes.byref(t)))
        return DeviceTree(self, t)

    def tree_from_host(self, elems, width, height, split=False):
        _as_u64(elems, "buff")
        t = vp()
        check(self._L.pil2gpu_tree_from_host(self.handle, _ptr(elems), width, height, int(split), ctypes.byref(t)))
        return DeviceTree(self, t)


class DeviceBuffer:
    """A device allocation of u64 words owned through the C ABI (pil2gpu_dev_alloc / pil2gpu_dev_free)."""

    def __init__(self, ctx, words):
        self.ctx, self.words = ctx, int(words)
        p = vp()
        check(ctx._L.pil2gpu_dev_alloc(ctx.handle, self.words * 8, ctypes.byref(p)))
        self.ptr = p

    def download(self):
        a = np.empty(self.words, dtype=np.uint64)
        check(self.ctx._L.pil2gpu_d2h(self.ctx.handle, _ptr(a), self.ptr, self.words * 8))
        self.ctx.sync()
        return a

    def free(self):
        if self.ptr:
            self.ctx.sync()
            self.ctx._L.pil2gpu_dev_free(self.ctx.handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class DeviceTree:
    """Device-resident {elements, nodes, width, height} (the tree object of merklehash_p.js:46-51)."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        w, h = ctypes.c_uint64(), ctypes.c_uint64()
        check(ctx._L.pil2gpu_tree_width(handle, ctypes.byref(w), ctypes.byref(h)))
        self.width, self.height = int(w.value), int(h.value)

    def root(self):
        r = np.empty(4, dtype=np.uint64)
        check(self.ctx._L.pil2gpu_tree_root(self.ctx.handle, self._h, _ptr(r)))
        return r

    def group_proofs(self, idxs):
        idxs = [int(i) for i in idxs]
        if any(i < 0 for i in idxs):
            raise _lib.OutOfRange(_lib.E_RANGE, "Out of range")
        ia = np.asarray(idxs, dtype=np.uint64)
        depth = int(self.ctx._L.pil2gpu_merkle_depth(self.height))
        rows = np.empty((len(idxs), self.width), dtype=np.uint64)
        sib = np.empty((len(idxs), depth, 4), dtype=np.uint64)
        check(self.ctx._L.pil2gpu_tree_group_proofs(self.ctx.handle, self._h, _ptr(ia) if len(idxs) else None, len(idxs), _ptr(rows),
                                                    _ptr(sib)))
        return rows, sib

    def fill(self, which, offset_words, chunk):
        """Copy a host chunk into the elements (which = 0) or the nodes (which = 1) at a word offset."""
        a = np.ascontiguousarray(chunk, dtype=np.uint64).reshape(-1)
        check(self.ctx._L.pil2gpu_tree_fill(self.ctx.handle, self._h, int(which), int(offset_words), _ptr(a) if a.size else None, a.size))

    def download(self, elements=True, nodes=True):
        e = np.empty(self.width * self.height, dtype=np.uint64) if elements else None
        n = np.empty(self.ctx.merkle_nnodes(self.height), dtype=np.uint64) if nodes else None
        check(self.ctx._L.pil2gpu_tree_download(self.ctx.handle, self._h, _ptr(e) if e is not None else None,
                                                _ptr(n) if n is not None else None))
        return e, n

    @property
    def elements_ptr(self):
        return self.ctx._L.pil2gpu_tree_elements_dev(self._h)

    @property
    def nodes_ptr(self):
        return self.ctx._L.pil2gpu_tree_nodes_dev(self._h)

    def free(self):
        if self._h:
            self.ctx._L.pil2gpu_tree_free(self.ctx.handle, self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


_default = {}


def default_context(device=0):
    """Process-wide context on `device` (created on first use; raises without a CUDA device)."""
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
