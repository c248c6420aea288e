"""ctypes front-end of the C oracle (oracle/gl_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes, os, pathlib, subprocess
import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libgloracle.so"
_U64P = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    src = _DIR / "gl_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        # -march=native is avoided: the .so is built in one container and also used on the GPU box
        subprocess.check_call(["make", "-C", str(_DIR), "CFLAGS=-O3 -fPIC -Wall -Wextra -pthread"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        L.orc_ntt.argtypes = [_U64P, _U64P, ctypes.c_uint64, ctypes.c_uint, ctypes.c_int, ctypes.c_int]
        L.orc_lde.argtypes = [_U64P, _U64P, ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, ctypes.c_int]
        L.orc_poseidon_perm.argtypes = [_U64P, _U64P]
        L.orc_poseidon_perm.restype = None
        L.orc_linear_hash.argtypes = [_U64P, ctypes.c_uint64, ctypes.c_int, _U64P]
        L.orc_linear_hash.restype = None
        L.orc_merkle_nnodes.argtypes = [ctypes.c_uint64]
        L.orc_merkle_nnodes.restype = ctypes.c_uint64
        L.orc_merkelize.argtypes = [_U64P, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, _U64P, ctypes.c_int]
        L.orc_group_proof.argtypes = [_U64P, _U64P, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, _U64P, _U64P]
        L.orc_fri_fold.argtypes = [_U64P, ctypes.c_uint, ctypes.c_uint, ctypes.c_int, ctypes.c_uint, _U64P, _U64P,
                                   _U64P, ctypes.c_int]
        L.orc_compute_q.argtypes = [_U64P, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, _U64P, ctypes.c_int]
        L.orc_lev.argtypes = [_U64P, ctypes.c_int, ctypes.c_uint, _U64P, ctypes.c_int]
        L.orc_eval.argtypes = [_U64P, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, _U64P, ctypes.c_uint, ctypes.c_uint, _U64P]
        L.orc_eval.restype = None
        L.orc_xdivxsubxi.argtypes = [_U64P, ctypes.POINTER(ctypes.c_int), ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, _U64P, ctypes.c_int]
        L.orc_fri_pol.argtypes = [ctypes.POINTER(_U64P), _U64P, _U64P, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_uint64,
                                  _U64P, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_uint64, _U64P, _U64P, _U64P, ctypes.c_uint, _U64P,
                                  ctypes.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_U64P)


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def default_threads():
    return max(1, os.cpu_count() or 1)


def ntt(src, n_pols, n_bits, inverse=False, threads=None):
    src = _arr(src); dst = np.empty_like(src)
    assert src.size == n_pols << n_bits
    lib().orc_ntt(_p(src), _p(dst), n_pols, n_bits, int(inverse), threads or default_threads())
    return dst


def lde(src, n_pols, n_bits, n_bits_ext, threads=None):
    src = _arr(src); dst = np.empty(n_pols << n_bits_ext, dtype=np.uint64)
    assert src.size == n_pols << n_bits
    rc = lib().orc_lde(_p(src), _p(dst), n_pols, n_bits, n_bits_ext, threads or default_threads())
    assert rc == 0
    return dst


def poseidon_perm(state12):
    s = _arr(state12); o = np.empty(12, dtype=np.uint64)
    lib().orc_poseidon_perm(_p(s), _p(o))
    return o


def linear_hash(vals, split=False):
    v = _arr(vals); o = np.empty(4, dtype=np.uint64)
    lib().orc_linear_hash(_p(v) if v.size else None, v.size, int(split), _p(o))
    return o


def merkle_nnodes(height):
    return int(lib().orc_merkle_nnodes(height))


def merkelize(elems, width, height, split=False, threads=None):
    e = _arr(elems)
    assert e.size == width * height
    nodes = np.empty(merkle_nnodes(height), dtype=np.uint64)
    lib().orc_merkelize(_p(e), width, height, int(split), _p(nodes), threads or default_threads())
    return nodes


def group_proof(elems, nodes, width, height, idx):
    e = _arr(elems); nd = _arr(nodes)
    row = np.empty(width, dtype=np.uint64); sib = np.empty(4 * 64, dtype=np.uint64)
    d = lib().orc_group_proof(_p(e), _p(nd), width, height, idx, _p(row), _p(sib))
    if d < 0:
        raise IndexError("Out of range")
    return row, sib[:4 * d].reshape(d, 4).copy()


def fri_fold(pol, prev_bits, cur_bits, next_bits, step0_bits, challenge, threads=None):
    """pol: (2^prev, 3) uint64.  next_bits None => last step (no rows).  Returns (pol2, rows|None)."""
    pol = _arr(pol).reshape(-1)
    assert pol.size == 3 << prev_bits
    ch = _arr(challenge)
    pol2 = np.empty(3 << cur_bits, dtype=np.uint64)
    rows = np.empty(3 << cur_bits, dtype=np.uint64) if next_bits is not None else None
    lib().orc_fri_fold(_p(pol), prev_bits, cur_bits, 0 if next_bits is None else next_bits + 1, step0_bits, _p(ch),
                       _p(pol2), _p(rows) if rows is not None else None, threads or default_threads())
    return pol2.reshape(-1, 3), rows


def compute_q(q_ext, q_dim, q_deg, n_bits, n_bits_ext, threads=None):
    """computeQStark up to the merkelize (stark_gen_helpers.js:168-192)."""
    q_ext = _arr(q_ext)
    assert q_ext.size == q_dim << n_bits_ext
    out = np.empty((q_dim * q_deg) << n_bits_ext, dtype=np.uint64)
    rc = lib().orc_compute_q(_p(q_ext), q_dim, q_deg, n_bits, n_bits_ext, _p(out), threads or default_threads())
    assert rc == 0
    return out


def lev(xi_challenge, opening, n_bits, threads=None):
    """LEv of computeEvalsStark (stark_gen_helpers.js:216-231): (2^n_bits, 3) array."""
    xi = _arr(xi_challenge)
    out = np.empty(3 << n_bits, dtype=np.uint64)
    rc = lib().orc_lev(_p(xi), int(opening), n_bits, _p(out), threads or default_threads())
    assert rc == 0
    return out.reshape(-1, 3)


def evals(buffers, ev_map, levs, n_bits, extend_bits):
    """computeEvalsStark loop (stark_gen_helpers.js:234-267); same arguments as gl_spec.compute_evals with numpy buffers."""
    out = np.empty((len(ev_map), 3), dtype=np.uint64)
    bufs = {k: (_arr(v[0]), v[1]) for k, v in buffers.items()}
    levs = [_arr(l).reshape(-1) for l in levs]
    for i, (name, offset, dim, oi) in enumerate(ev_map):
        buf, size = bufs[name]
        r = np.empty(3, dtype=np.uint64)
        lib().orc_eval(_p(buf), size, offset, dim, _p(levs[oi]), n_bits, extend_bits, _p(r))
        out[i] = r
    return out


def x_div_x_sub_xi(xi_challenge, openings, n_bits, n_bits_ext, threads=None):
    """xDivXSubXi_ext of computeFRIStark (stark_gen_helpers.js:289-323): (2^n_bits_ext, nOpenings, 3) array."""
    xi = _arr(xi_challenge)
    op = (ctypes.c_int * len(openings))(*[int(o) for o in openings])
    out = np.empty(3 * len(openings) << n_bits_ext, dtype=np.uint64)
    rc = lib().orc_xdivxsubxi(_p(xi), op, len(openings), n_bits, n_bits_ext, _p(out), threads or default_threads())
    assert rc == 0
    return out.reshape(-1, len(openings), 3)


def fri_polynomial(buffers, ev_map, evals_, openings, xdiv, vf1, vf2, n_bits_ext, threads=None):
    """f_ext of computeFRIStark (friPolinomial.js:26-56 evaluated per row); same arguments as gl_spec.fri_polynomial with numpy
    buffers.  Returns a (2^n_bits_ext, 3) array."""
    from .gl_spec import js_object_key_order
    bufs = {k: (_arr(v[0]), int(v[1])) for k, v in buffers.items()}
    first_use = []
    for _, _, _, prime in ev_map:
        if prime not in first_use:
            first_use.append(prime)
    order = js_object_key_order(first_use)
    n = len(ev_map)
    ptrs = (_U64P * n)(*[_p(bufs[name][0]) for name, _, _, _ in ev_map])
    sizes = np.array([bufs[name][1] for name, _, _, _ in ev_map], dtype=np.uint64)
    offs = np.array([o for _, o, _, _ in ev_map], dtype=np.uint64)
    dims = (ctypes.c_int * n)(*[int(d) for _, _, d, _ in ev_map])
    group = (ctypes.c_int * n)(*[order.index(pr) for _, _, _, pr in ev_map])
    xidx = (ctypes.c_int * len(order))(*[list(openings).index(o) for o in order])
    ev = _arr(np.asarray(evals_, dtype=np.uint64).reshape(-1))
    xd, a, b = _arr(np.asarray(xdiv).reshape(-1)), _arr(vf1), _arr(vf2)
    out = np.empty(3 << n_bits_ext, dtype=np.uint64)
    rc = lib().orc_fri_pol(ptrs, _p(sizes), _p(offs), dims, group, n, _p(ev), xidx, len(order), len(openings), _p(xd), _p(a), _p(b), n_bits_ext,
                           _p(out), threads or default_threads())
    assert rc == 0
    return out.reshape(-1, 3)
