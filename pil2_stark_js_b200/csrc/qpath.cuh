// Quotient-polynomial commit path (computeQStark, src/stark/stark_gen_helpers.js:168-208) for sm_100a.
//
// Reference semantics:  qq1 = ifft(q_ext)  (E x qDim, coefficients c_j * 7^j of Q);  chunk p < qDeg takes coefficients
// [pN, (p+1)N), scaled by shiftIn^p with shiftIn = 7^-N (:178-190), laid out as columns [p*qDim, (p+1)*qDim) of an E-row buffer
// whose rows >= N stay zero;  cmQ_ext = fft(qq2) over all E rows (:192);  merkelize (:198).
//
// Here: the INTT runs as DIF passes and leaves the coefficients in bit-reversed row order (no permutation pass, no 1/E
// scale); row q*B + bitrev_b(p) of that buffer is coefficient p*N + bitrev_n(q), i.e. exactly the bit-reversed input row q of
// chunk p's size-N transform.  q_split_kernel gathers those rows, applies shiftIn^p / E, and the forward transform of the
// zero-padded chunks is B coset NTTs of size N on w_E^r <w_N> (ntt_launch_coset_eval, unit shift) -- the zero rows are never
// materialised.
#pragma once
#include "gl.cuh"

// T[q * (qDim*qDeg) + p*qDim + k] = S[(q*B + bitrev_b(p)) * qDim + k] * fac[p]      (fac in Montgomery form)
__global__ void q_split_kernel(const u64* __restrict__ S, u64* __restrict__ T, const u64* __restrict__ fac, u64 N, int b_bits, u32 qDim,
                               u32 qDeg) {
    const u64 cols = (u64)qDim * qDeg;
    const u64 total = N * cols, stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const u64 q = i / cols;
        const u32 c = (u32)(i - q * cols);
        const u32 p = c / qDim, k = c - p * qDim;
        const u32 pr = b_bits == 0 ? 0u : (__brev(p) >> (32 - b_bits));
        T[i] = gl_mmul(S[((q << b_bits) + pr) * qDim + k], fac[p]);
    }
}
