"""ctypes binding of libpil2gpu.so (include/pil2gpu.h).  No fallback: if the library cannot be loaded, or no
CUDA device is present, every operation raises."""
import ctypes
import pathlib

_PKG = pathlib.Path(__file__).resolve().parent
import os

# PIL2GPU_LIB: an alternative build of the same library (A/B measurements of kernel variants); default: the in-tree build
LIB_PATH = pathlib.Path(os.environ["PIL2GPU_LIB"]).resolve() if os.environ.get("PIL2GPU_LIB") else _PKG / "libpil2gpu.so"

OK, E_INVALID, E_RANGE, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5

u64p = ctypes.POINTER(ctypes.c_uint64)
u64pp = ctypes.POINTER(u64p)
vp = ctypes.c_void_p
c_u64, c_u32, c_i32, c_int, c_size = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32, ctypes.c_int, ctypes.c_size_t


class Pil2GpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pil2gpu error {code}: {msg}")
        self.code = code
        self.message = msg


class OutOfRange(Pil2GpuError, IndexError):
    """Mirrors `throw new Error("Out of range")` of merklehash_p.js:143."""


class EvalDesc(ctypes.Structure):
    """pil2gpu_eval_desc (include/pil2gpu.h)"""
    _fields_ = [("offset", ctypes.c_uint64), ("dim", ctypes.c_uint32), ("lev", ctypes.c_uint32)]


class ExprBuffer(ctypes.Structure):
    """pil2gpu_expr_buffer (include/pil2gpu.h)"""
    _fields_ = [("ptr_dev", ctypes.c_void_p), ("row_words", ctypes.c_uint64)]


class ExprHostBuffer(ctypes.Structure):
    """pil2gpu_expr_host_buffer (include/pil2gpu.h)"""
    _fields_ = [("ptr", ctypes.c_void_p), ("row_words", ctypes.c_uint64), ("read", ctypes.c_int32), ("written", ctypes.c_int32)]


class FriTerm(ctypes.Structure):
    """pil2gpu_fri_term (include/pil2gpu.h)"""
    _fields_ = [("buf_dev", ctypes.c_void_p), ("size", ctypes.c_uint64), ("offset", ctypes.c_uint64), ("dim", ctypes.c_uint32),
                ("prime", ctypes.c_int32)]


_SIGS = {
    # name: (restype, [argtypes])
    "pil2gpu_create": (c_int, [c_int, vp, ctypes.POINTER(vp)]),
    "pil2gpu_destroy": (None, [vp]),
    "pil2gpu_last_error": (ctypes.c_char_p, []),
    "pil2gpu_version": (ctypes.c_char_p, []),
    "pil2gpu_sync": (c_int, [vp]),
    "pil2gpu_release_workspace": (c_int, [vp]),
    "pil2gpu_launch_count": (c_u64, [vp]),
    "pil2gpu_dev_alloc": (c_int, [vp, c_size, ctypes.POINTER(vp)]),
    "pil2gpu_dev_free": (c_int, [vp, vp]),
    "pil2gpu_host_alloc": (c_int, [c_size, ctypes.POINTER(vp)]),
    "pil2gpu_host_free": (c_int, [vp]),
    "pil2gpu_h2d": (c_int, [vp, vp, vp, c_size]),
    "pil2gpu_d2h": (c_int, [vp, vp, vp, c_size]),
    "pil2gpu_shard_create": (c_int, [vp, c_u32, c_u32, c_u64, c_u64, ctypes.POINTER(vp)]),
    "pil2gpu_shard_destroy": (c_int, [vp]),
    "pil2gpu_shard_handles": (c_int, [vp, vp]),
    "pil2gpu_shard_connect": (c_int, [vp, vp, c_u32]),
    "pil2gpu_shard_connect_local": (c_int, [ctypes.POINTER(vp), c_u32]),
    "pil2gpu_shard_recv_dev": (vp, [vp]),
    "pil2gpu_shard_peer_recv": (vp, [vp]),
    "pil2gpu_shard_sub_roots_dev": (vp, [vp]),
    "pil2gpu_shard_top_nodes_dev": (vp, [vp]),
    "pil2gpu_shard_barrier": (c_int, [vp]),
    "pil2gpu_shard_status": (c_int, [vp]),
    "pil2gpu_shard_commit_dev": (c_int, [vp, vp, vp, c_u64, c_u32, c_u32, c_int, vp, vp]),
    "pil2gpu_shard_hash_dev": (c_int, [vp, c_u64, c_u32, c_int, vp, vp]),
    "pil2gpu_shard_open_dev": (c_int, [vp, vp, c_u64, c_u32, vp, c_u32, vp, vp]),
    "pil2gpu_ntt": (c_int, [vp, vp, vp, c_u64, c_u32, c_int]),
    "pil2gpu_ntt_dev": (c_int, [vp, vp, vp, c_u64, c_u32, c_int]),
    "pil2gpu_lde": (c_int, [vp, vp, vp, c_u64, c_u32, c_u32]),
    "pil2gpu_lde_dev": (c_int, [vp, vp, vp, c_u64, c_u32, c_u32]),
    "pil2gpu_lde_paged": (c_int, [vp, ctypes.POINTER(vp), u64p, c_u32, ctypes.POINTER(vp), u64p, c_u32, c_u64, c_u32, c_u32]),
    "pil2gpu_ntt_paged": (c_int, [vp, ctypes.POINTER(vp), u64p, c_u32, ctypes.POINTER(vp), u64p, c_u32, c_u64, c_u32, c_int]),
    "pil2gpu_extend_and_merkelize_paged": (c_int, [vp, ctypes.POINTER(vp), u64p, c_u32, c_u64, c_u32, c_u32, c_int, ctypes.POINTER(vp), u64p, c_u32,
                                                   vp, vp]),
    "pil2gpu_compute_q_paged": (c_int, [vp, ctypes.POINTER(vp), u64p, c_u32, c_u64, c_u64, c_u32, c_u32, c_int, ctypes.POINTER(vp), u64p, c_u32,
                                        vp, vp]),
    "pil2gpu_calculate_exps_dev": (c_int, [vp, vp, c_u32, vp, c_u32, vp, c_u32, c_u32, c_int]),
    "pil2gpu_calculate_exps": (c_int, [vp, vp, c_u32, vp, c_u32, vp, c_u32, c_u32, c_int]),
    "pil2gpu_expr_jit_check": (c_int, [vp, c_u32, vp, c_u32, c_u32, c_int, ctypes.c_char_p, c_u64]),
    "pil2gpu_fri_fold_range_dev": (c_int, [vp, vp, c_int, c_u32, c_u32, c_i32, c_u32, vp, c_u64, c_u64, vp, vp]),
    "pil2gpu_fri_fold_paged": (c_int, [vp, ctypes.POINTER(vp), u64p, c_u32, c_u32, c_u32, c_i32, c_u32, vp, c_int, ctypes.POINTER(vp), u64p, c_u32,
                                       ctypes.POINTER(vp), u64p, c_u32, vp]),
    "pil2gpu_ipc_export": (c_int, [vp, vp, vp]),
    "pil2gpu_ipc_open": (c_int, [vp, vp, ctypes.POINTER(vp)]),
    "pil2gpu_ipc_close": (c_int, [vp, vp]),
    "pil2gpu_lde_scatter_dev": (c_int, [vp, vp, vp, c_u64, c_u32, c_u32, ctypes.POINTER(vp), c_u32, c_u32, c_u64, c_u64]),
    "pil2gpu_lde_scatter": (c_int, [vp, vp, c_u64, c_u64, c_u32, c_u32, ctypes.POINTER(vp), c_u32, c_u32]),
    "pil2gpu_compute_q_dev": (c_int, [vp, vp, c_u64, c_u64, c_u32, c_u32, vp]),
    "pil2gpu_compute_q": (c_int, [vp, vp, c_u64, c_u64, c_u32, c_u32, c_int, vp, vp, vp]),
    "pil2gpu_compute_lev_dev": (c_int, [vp, vp, c_i32, c_u32, vp]),
    "pil2gpu_compute_evals_dev": (c_int, [vp, vp, c_u64, c_u32, c_u32, vp, c_u32, vp, c_u32, vp]),
    "pil2gpu_x_div_x_sub_xi_dev": (c_int, [vp, vp, vp, c_u32, c_u32, c_u32, vp]),
    "pil2gpu_compute_evals": (c_int, [vp, vp, vp, c_u32, c_u32, c_u32, vp, c_u64, vp, c_u32, vp]),
    "pil2gpu_x_div_x_sub_xi": (c_int, [vp, vp, vp, c_u32, c_u32, c_u32, vp]),
    "pil2gpu_fri_pol_dev": (c_int, [vp, vp, c_u32, vp, vp, c_u32, vp, vp, vp, c_u32, vp]),
    "pil2gpu_fri_pol": (c_int, [vp, vp, c_u32, vp, vp, c_u32, vp, vp, vp, c_u32, c_u32, vp, vp]),
    "pil2gpu_poseidon": (c_int, [vp, vp, vp]),
    "pil2gpu_linear_hash": (c_int, [vp, vp, c_u64, c_int, vp]),
    "pil2gpu_merkle_nnodes": (c_u64, [c_u64]),
    "pil2gpu_merkle_depth": (c_u32, [c_u64]),
    "pil2gpu_merkelize": (c_int, [vp, vp, c_u64, c_u64, c_int, vp]),
    "pil2gpu_merkelize_dev": (c_int, [vp, vp, c_u64, c_u64, c_int, vp]),
    "pil2gpu_merkelize_tiled_dev": (c_int, [vp, vp, c_u32, c_u64, c_u64, c_u64, c_int, vp]),
    "pil2gpu_merkle_tree_from_digests_dev": (c_int, [vp, vp, c_u64]),
    "pil2gpu_tree_wrap_tiled_dev": (c_int, [vp, vp, c_u32, c_u64, c_u64, vp, c_u64, ctypes.POINTER(vp)]),
    "pil2gpu_synth2d_dev": (c_int, [vp, vp, c_u64, c_u64, c_u64, c_u64, c_u64]),
    "pil2gpu_merkelize_paged": (c_int, [vp, ctypes.POINTER(vp), u64p, c_u32, c_u64, c_u64, c_int, vp]),
    "pil2gpu_commit": (c_int, [vp, vp, c_u64, c_u32, c_u32, c_int, ctypes.POINTER(vp), vp]),
    "pil2gpu_commit_dev": (c_int, [vp, vp, c_u64, c_u32, c_u32, c_int, ctypes.POINTER(vp), vp]),
    "pil2gpu_extend_and_merkelize": (c_int, [vp, vp, c_u64, c_u32, c_u32, c_int, vp, vp, vp]),
    "pil2gpu_synth_dev": (c_int, [vp, vp, c_u64, c_u64, c_u64]),
    "pil2gpu_bench_int_pipes": (c_int, [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "pil2gpu_tree_from_host": (c_int, [vp, vp, c_u64, c_u64, c_int, ctypes.POINTER(vp)]),
    "pil2gpu_tree_alloc": (c_int, [vp, c_u64, c_u64, ctypes.POINTER(vp)]),
    "pil2gpu_tree_fill": (c_int, [vp, vp, c_int, c_u64, vp, c_u64]),
    "pil2gpu_tree_width": (c_int, [vp, u64p, u64p]),
    "pil2gpu_tree_elements_dev": (vp, [vp]),
    "pil2gpu_tree_nodes_dev": (vp, [vp]),
    "pil2gpu_tree_root": (c_int, [vp, vp, vp]),
    "pil2gpu_tree_group_proofs": (c_int, [vp, vp, vp, c_u32, vp, vp]),
    "pil2gpu_tree_group_proofs_dev": (c_int, [vp, vp, vp, c_u32, vp, vp]),
    "pil2gpu_tree_download": (c_int, [vp, vp, vp, vp]),
    "pil2gpu_tree_free": (None, [vp, vp]),
    "pil2gpu_tree_wrap_dev": (c_int, [vp, vp, vp, c_u64, c_u64, ctypes.POINTER(vp)]),
    "pil2gpu_fri_fold": (c_int, [vp, vp, c_u32, c_u32, c_i32, c_u32, vp, c_int, vp, vp, vp]),
    "pil2gpu_fri_fold_dev": (c_int, [vp, vp, c_u32, c_u32, c_i32, c_u32, vp, c_int, vp, vp, vp]),
}

_lib = None


def load():
    """Load libpil2gpu.so (built in-tree by pil2_stark_js_b200/build.py; built on first use when missing and nvcc is present).
    Raises if it is missing and cannot be built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            try:                           # fresh checkout (the .so is git-ignored): build once if nvcc is here
                from . import build as _build
                _build.build()
            except Exception as ex:        # noqa: BLE001
                raise ImportError(
                    f"{LIB_PATH} is missing and could not be built ({ex}): run `python -m pil2_stark_js_b200.build` where nvcc "
                    "is installed.  There is no CPU fallback for the commit path.") from ex
        L = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        msg = load().pil2gpu_last_error().decode("utf-8", "replace")
        if rc == E_RANGE:
            raise OutOfRange(rc, msg)
        raise Pil2GpuError(rc, msg)
