// Poseidon-GL permutation (width 12 = rate 8 + capacity 4, x^7, 8 full + 22 partial rounds) for sm_100a.
//
// Functionally identical to the reference's two formulations, which agree with each other:
//   plain 30-round form      src/helpers/glwasm.js:359-390 (MDS :428-440, round constants :536-627)
//   optimised (sparse) form  src/helpers/hash/poseidon/poseidon.js:57-108
// One permutation per thread, the 12-word state lives in 24 registers, round constants come from
// __constant__ memory (warp-uniform index -> constant-cache broadcast).  The kernel is bound by the
// integer pipes (IMAD.WIDE.U32 / IADD3), not by HBM: see DESIGN.md "Poseidon".
#pragma once
#include "gl.cuh"

// 30 rounds x 12 lanes, plain form (generated: tools/gen_poseidon_rc.py)
__constant__ u64 POSEIDON_RC[360] = {
#include "poseidon_rc.inc"
};

// x -> x^7 : 4 multiplications (2 of them squarings)
GL_D u64 poseidon_sbox(u64 x) {
    u64 x2 = gl_sqr(x);
    u64 x3 = gl_mul(x2, x);
    u64 x4 = gl_sqr(x2);
    return gl_mul(x3, x4);
}

// Dense MDS layer: out_i = sum_j circ[(j - i) mod 12] * x_j  (+ 8*x_0 on lane 0), glwasm.js:428-440.
// The coefficients are < 2^6, so each state word is split into 32-bit halves and the two half-sums are
// accumulated exactly in 64 bits with one IMAD.WIDE.U32 per term (12*41*2^32 < 2^42); the halves are
// recombined with a single 96-bit reduction per lane.
GL_D void poseidon_mds(u64 x[12]) {
    const u32 circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 lo[12], hi[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        lo[j] = (u32)x[j];
        hi[j] = (u32)(x[j] >> 32);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 L = 0, H = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const u32 c = circ[(j - i + 12) % 12] + ((i == 0 && j == 0) ? 8u : 0u);
            L += (u64)lo[j] * c;
            H += (u64)hi[j] * c;
        }
        // value = L + H*2^32 = (L + Hh*EPS) + (Hl << 32), Hh = H >> 32 < 2^10
        u64 base = L + (u64)(u32)(H >> 32) * (u64)0xFFFFFFFFu;   // < 2^43, no wrap
        u64 top = (u64)(u32)H << 32;
        u64 r = base + top;
        if (r < top) r += GL_EPS;
        x[i] = r;
    }
}

// Plain-form permutation; state in lazy form on input, lazy form on output (callers canonicalise).
GL_D void poseidon_permute(u64 x[12]) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) x[i] = poseidon_sbox(gl_add(x[i], POSEIDON_RC[12 * r + i]));
        poseidon_mds(x);
    }
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
#pragma unroll
        for (int i = 1; i < 12; i++) x[i] = gl_add(x[i], POSEIDON_RC[12 * r + i]);
        x[0] = poseidon_sbox(gl_add(x[0], POSEIDON_RC[12 * r]));
        poseidon_mds(x);
    }
#pragma unroll 1
    for (int r = 26; r < 30; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) x[i] = poseidon_sbox(gl_add(x[i], POSEIDON_RC[12 * r + i]));
        poseidon_mds(x);
    }
}
