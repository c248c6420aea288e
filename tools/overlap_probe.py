"""Do the LDE kernels and the leaf-hash kernel speed up when they share the GPU?  Two contexts (two streams) on one device: context A
runs an LDE, context B hashes an independent buffer; alone, then together.  (Diagnostic for overlapping the LDE of column slab k+1 with the
hashing of slab k inside one commit.)   usage: python tools/overlap_probe.py [n_bits] [cols]"""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pil2_stark_js_b200 import _lib
L = _lib.load(); check = _lib.check; vp = ctypes.c_void_p
n_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 22
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ext = n_bits + 1
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
ha, hb = vp(), vp()
check(L.pil2gpu_create(0, vp(sa.cuda_stream), ctypes.byref(ha))); check(L.pil2gpu_create(0, vp(sb.cuda_stream), ctypes.byref(hb)))
dev = lambda words: torch.empty(int(words), dtype=torch.int64, device="cuda")
src, dst = dev(cols << n_bits), dev(cols << ext)
ext_b, nodes = dev(cols << ext), dev(int(L.pil2gpu_merkle_nnodes(1 << ext)))
p = lambda t: vp(t.data_ptr())
check(L.pil2gpu_synth_dev(ha, p(src), cols << n_bits, 1, 0)); check(L.pil2gpu_synth_dev(ha, p(ext_b), cols << ext, 2, 0)); torch.cuda.synchronize()
def lde(): check(L.pil2gpu_lde_dev(ha, p(src), p(dst), cols, n_bits, ext))
def hsh(): check(L.pil2gpu_merkelize_dev(hb, p(ext_b), cols, 1 << ext, 0, p(nodes)))
def timed(fns, reps=3):
    for f in fns: f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for f in fns: f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
ta, tb = timed([lde]), timed([hsh])
tab, tba = timed([lde, hsh]), timed([hsh, lde])
print(f"2^{n_bits} x {cols}: LDE alone {ta*1e3:.1f} ms, hashing alone {tb*1e3:.1f} ms, sum {(ta+tb)*1e3:.1f} ms; together (LDE launched first) {tab*1e3:.1f} ms, "
      f"(hash first) {tba*1e3:.1f} ms")
# slab-wise: the LDE cut into 4 column slabs interleaved with 4 quarter-height hash launches, so that both kernels always have CTAs pending
q = cols // 4
slabs = [(dev(q << n_bits), dev(q << ext)) for _ in range(4)]
for s, _ in slabs: check(L.pil2gpu_synth_dev(ha, p(s), q << n_bits, 3, 0))
hq = (1 << ext) // 4
def mixed():
    for k, (s, d) in enumerate(slabs):
        check(L.pil2gpu_lde_dev(ha, p(s), p(d), q, n_bits, ext))
        check(L.pil2gpu_merkelize_dev(hb, vp(ext_b.data_ptr() + 8 * k * hq * cols), cols, hq, 0, p(nodes)))
print(f"4 LDE slabs on stream A interleaved with 4 quarter hashes on stream B: {timed([mixed])*1e3:.1f} ms")
