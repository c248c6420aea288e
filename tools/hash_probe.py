"""Leaf hashing A/B probe: merkelize of a device-resident 2^n_bits x cols buffer with the library named by PIL2GPU_LIB (default: the
in-tree build); checks a small case against the C oracle first.   usage: [PIL2GPU_LIB=...] python tools/hash_probe.py [n_bits] [cols]"""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pil2_stark_js_b200 import _lib
import pil2_stark_js_b200 as m
L = _lib.load(); check = _lib.check; vp = ctypes.c_void_p
n_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 22
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = m.default_context(0)
try:
    from oracle import gl_oracle as C
    rng = np.random.default_rng(5)
    for (n, w) in [(4099, 40), (8192, 9), (5000, 256)]:
        buff = rng.integers(0, 0xFFFFFFFF00000001, size=n * w, dtype=np.uint64)
        buff[:w] = 0xFFFFFFFF00000000; buff[w:2 * w] = 0
        ok = np.array_equal(ctx.merkelize(buff, w, n, False), C.merkelize(buff, w, n, False))
        print(f"parity {n} x {w}: {'ok' if ok else 'MISMATCH'}", flush=True)
except ImportError as ex:
    print("oracle not available:", ex)
s = torch.cuda.Stream()
h = vp(); check(L.pil2gpu_create(0, vp(s.cuda_stream), ctypes.byref(h)))
H = 1 << n_bits
buf = torch.empty(cols * H, dtype=torch.int64, device="cuda")
nodes = torch.empty(int(L.pil2gpu_merkle_nnodes(H)), dtype=torch.int64, device="cuda")
check(L.pil2gpu_synth_dev(h, vp(buf.data_ptr()), cols * H, 2, 0)); torch.cuda.synchronize()
def run(): check(L.pil2gpu_merkelize_dev(h, vp(buf.data_ptr()), cols, H, 0, vp(nodes.data_ptr())))
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    e0.record(s)
    for _ in range(3): run()
    e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
perms = H * ((cols + 7) // 8) + H - 1
print(f"{os.environ.get('PIL2GPU_LIB', 'default')}: merkelize 2^{n_bits} x {cols}: {ms:.2f} ms, {perms / ms / 1e6:.3f} Gperm/s, root {nodes[-4:].tolist()}")
