"""Times pil2gpu_extend_and_merkelize variants on cfg3 with pinned host buffers (diagnostic)."""
import ctypes, time, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pil2_stark_js_b200 import _lib
L = _lib.load(); check = _lib.check; vp = ctypes.c_void_p
h = vp(); check(L.pil2gpu_create(0, None, ctypes.byref(h)))
n_bits, cols, blow = 23, 256, 1
if len(sys.argv) > 1: n_bits = int(sys.argv[1])
ext = n_bits + blow
def pinned(words):
    p = vp(); check(L.pil2gpu_host_alloc(int(words) * 8, ctypes.byref(p))); return p
sw, dw, nw = cols << n_bits, cols << ext, int(L.pil2gpu_merkle_nnodes(1 << ext))
t0 = time.perf_counter(); hs, hd, hn = pinned(sw), pinned(dw), pinned(nw); print("pinned alloc %.2f s" % (time.perf_counter() - t0))
d = vp(); check(L.pil2gpu_dev_alloc(h, sw * 8, ctypes.byref(d)))
check(L.pil2gpu_synth_dev(h, d, sw, 0x5EED0003, 0)); check(L.pil2gpu_d2h(h, hs, d, sw * 8)); check(L.pil2gpu_sync(h)); check(L.pil2gpu_dev_free(h, d))
root = np.zeros(4, dtype=np.uint64); rp = vp(root.ctypes.data)
def run(name, dst, nodes, reps=3):
    check(L.pil2gpu_extend_and_merkelize(h, hs, cols, n_bits, ext, 0, dst, nodes, rp))
    t0 = time.perf_counter()
    for _ in range(reps):
        check(L.pil2gpu_extend_and_merkelize(h, hs, cols, n_bits, ext, 0, dst, nodes, rp))
    print("%-32s %.3f s  root %x" % (name, (time.perf_counter() - t0) / reps, int(root[0])))
run("dst + nodes", hd, hn)
os.environ["PIL2GPU_TRACE"] = "1"; run("traced", hd, hn, reps=1); del os.environ["PIL2GPU_TRACE"]
run("nodes only (no dst download)", None, hn)
run("root only", None, None)
