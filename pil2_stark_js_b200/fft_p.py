"""Mirror of src/helpers/fft/fft_p.js:299-302 -- fft, ifft, interpolate over row-major BigBuffers.

Buffers are numpy uint64 arrays (the BigUint64Array / BigBuffer layout buff[row*nPols + col]).  As in the reference the
caller allocates buffDst, buffSrc is not modified, and the result fully overwrites buffDst.  The reference functions are
async; these are synchronous (the JS shim in js/fft_p.js wraps the same C calls in a Promise)."""
from .context import default_context


def fft(buffSrc, nPols, nBits, buffDst, ctx=None):
    """fft_p.js:178-180: buffDst[k*nPols+c] = sum_j buffSrc[j*nPols+c] * w_n^(j k)."""
    (ctx or default_context()).ntt(buffSrc, nPols, nBits, buffDst, inverse=False)


def ifft(buffSrc, nPols, nBits, buffDst, ctx=None):
    """fft_p.js:182-184."""
    (ctx or default_context()).ntt(buffSrc, nPols, nBits, buffDst, inverse=True)


def interpolate(buffSrc, nPols, nBits, buffDst, nBitsExt, ctx=None):
    """fft_p.js:187-297: low-degree extension onto the coset 7*<w_ext>: buffDst[j*nPols+c] = P_c(7 * w_ext^j)."""
    (ctx or default_context()).lde(buffSrc, nPols, nBits, buffDst, nBitsExt)
