"""Mirror of pilcom's BigBuffer as the reference uses it (observed surface: `new BigBuffer(n)`, `.length`, `.slice(a, b)`,
`.set(arr, offset)`, `.getElement(i)`, `.setElement(i, v)` -- fft_p.js:28-29,61,89,102; merklehash_p.js:70,84,138,243,272;
stark_gen_helpers.js:104-114,186,254): one logical u64 buffer stored as a list of pages, because a single typed array cannot
hold the 16-64 GiB buffers of the production sizes.  Pages here are numpy uint64 arrays; with `pinned=True` they are allocated
through pil2gpu_host_alloc (page-locked: the library then copies them by DMA without staging), which is what the N-API addon's
`allocPinnedPage` gives the JS side.  Every *_paged entry point of the C ABI takes `pages()` of such a buffer."""
import ctypes

import numpy as np

from . import _lib

PAGE_LEN = 1 << 28          # elements per page (2 GiB), the order of magnitude pilcom uses; any page length works with the C ABI


class BigBuffer:
    def __init__(self, size, page_len=PAGE_LEN, pinned=False, page_lens=None, zero=True):
        self.length = int(size)
        self.pinned = bool(pinned)
        self._zero = bool(zero)          # zero=False: skip the zero fill of fresh pinned pages (a buffer that is fully overwritten)
        self._raw = []
        self.buffers = []
        lens = list(page_lens) if page_lens is not None else [min(page_len, self.length - i) for i in range(0, self.length, page_len)]
        if sum(lens) != self.length:
            raise ValueError("page lengths do not add up to the buffer size")
        for n in lens:
            self.buffers.append(self._new_page(int(n)))
        self._starts = np.concatenate([[0], np.cumsum([b.size for b in self.buffers])]).astype(np.int64)

    def _new_page(self, n):
        if not self.pinned or n == 0:
            return np.zeros(n, dtype=np.uint64) if self._zero else np.empty(n, dtype=np.uint64)
        L = _lib.load()
        p = ctypes.c_void_p()
        _lib.check(L.pil2gpu_host_alloc(n * 8, ctypes.byref(p)))
        self._raw.append(p)
        a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(n,))
        if self._zero:
            a[:] = 0
        return a

    def free(self):
        """Release pinned pages (ordinary pages are garbage-collected)."""
        L = _lib.load()
        self.buffers = []
        for p in self._raw:
            L.pil2gpu_host_free(p)
        self._raw = []

    def __del__(self):
        try:
            if self._raw:
                self.free()
        except Exception:
            pass

    @classmethod
    def from_array(cls, arr, page_len=PAGE_LEN, pinned=False, page_lens=None):
        arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1)
        b = cls(arr.size, page_len, pinned, page_lens)
        b.set(arr, 0)
        return b

    # ---- the reference's surface ----
    def _locate(self, idx):
        p = int(np.searchsorted(self._starts, idx, side="right")) - 1
        return p, idx - int(self._starts[p])

    def getElement(self, idx):
        if idx < 0 or idx >= self.length:
            raise IndexError("BigBuffer index out of range")
        p, o = self._locate(idx)
        return int(self.buffers[p][o])

    def setElement(self, idx, value):
        if idx < 0 or idx >= self.length:
            raise IndexError("BigBuffer index out of range")
        p, o = self._locate(idx)
        self.buffers[p][o] = int(value)

    def slice(self, a=0, b=None):
        """A COPY of elements [a, b) as one array (negative indices count from the end, like TypedArray.slice)."""
        b = self.length if b is None else b
        if a < 0:
            a += self.length
        if b < 0:
            b += self.length
        a, b = max(0, min(a, self.length)), max(0, min(b, self.length))
        out = np.empty(max(0, b - a), dtype=np.uint64)
        pos = a
        while pos < b:
            p, o = self._locate(pos)
            n = min(self.buffers[p].size - o, b - pos)
            out[pos - a:pos - a + n] = self.buffers[p][o:o + n]
            pos += n
        return out

    def set(self, arr, offset=0):
        arr = np.asarray(arr, dtype=np.uint64).reshape(-1)
        if offset < 0 or offset + arr.size > self.length:
            raise IndexError("BigBuffer.set out of range")
        pos, done = offset, 0
        while done < arr.size:
            p, o = self._locate(pos)
            n = min(self.buffers[p].size - o, arr.size - done)
            self.buffers[p][o:o + n] = arr[done:done + n]
            pos += n
            done += n

    def to_array(self):
        return self.slice(0, self.length)

    # ---- what the C ABI takes ----
    def pages(self):
        """(array of page pointers, array of page word counts, number of pages) for the *_paged entry points."""
        n = len(self.buffers)
        ptrs = (ctypes.c_void_p * max(1, n))(*[(b.ctypes.data if b.size else 0) for b in self.buffers])
        words = (ctypes.c_uint64 * max(1, n))(*[b.size for b in self.buffers])
        return ptrs, words, n


def as_pages(buff, name="buffer"):
    """Page description of a BigBuffer or of a plain C-contiguous numpy uint64 array (the one-page case)."""
    if isinstance(buff, BigBuffer):
        return buff.pages() + (buff.length,)
    if not isinstance(buff, np.ndarray) or buff.dtype != np.uint64 or not buff.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name} must be a BigBuffer or a C-contiguous numpy uint64 array (BigUint64Array layout)")
    ptrs = (ctypes.c_void_p * 1)(buff.ctypes.data if buff.size else 0)
    words = (ctypes.c_uint64 * 1)(buff.size)
    return ptrs, words, 1, buff.size
