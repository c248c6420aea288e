// Candidate Poseidon-GL permutation for the probe: Montgomery-form state, MDS on 64-bit lanes (lo/hi halves).
#pragma once
#include "poseidon.cuh"

__constant__ u64 POSEIDON_RC_MONT[372] = {   // RC[r][i] * 2^64 mod p; rows 1..30 used post-MDS (row 30 = 0)
#include "poseidon_rc_mont.inc"
};

#define poseidon_sbox_mont poseidon_sbox

// y = circ-MDS * x on lanes of type T (wrap-around arithmetic; true results are non-negative and small)
template <typename T>
GL_D void poseidon_mds_lanes(T y[12], const T x[12]) {
    T xp[6], xm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        xp[i] = x[i] + x[i + 6];
        xm[i] = x[i] - x[i + 6];
    }
    T Q[6];
    Q[0] = 2 * xm[0] + 4 * xm[5] - 16 * xm[4] - xm[3] + xm[2] + xm[1];
    Q[1] = 2 * xm[1] - 4 * xm[0] - 16 * xm[5] - xm[4] + xm[3] + xm[2];
    Q[2] = 2 * xm[2] - 4 * xm[1] + 16 * xm[0] - xm[5] + xm[4] + xm[3];
    Q[3] = 2 * xm[3] - 4 * xm[2] + 16 * xm[1] + xm[0] + xm[5] + xm[4];
    Q[4] = 2 * xm[4] - 4 * xm[3] + 16 * xm[2] + xm[1] - xm[0] + xm[5];
    Q[5] = 2 * xm[5] - 4 * xm[4] + 16 * xm[3] + xm[2] - xm[1] - xm[0];
    T xpp[3], xpm[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        xpp[i] = xp[i] + xp[i + 3];
        xpm[i] = xp[i] - xp[i + 3];
    }
    const T s = xpp[0] + xpp[1] + xpp[2];
    T PP[3] = {16 * (s + xpp[2]), 16 * (s + xpp[0]), 16 * (s + xpp[1])};
    T PQ[3];
    PQ[0] = 8 * xpm[2] - xpm[0] - 2 * xpm[1];
    PQ[1] = (T)0 - xpm[1] - 8 * xpm[0] - 2 * xpm[2];
    PQ[2] = 2 * xpm[0] - xpm[2] - 8 * xpm[1];
    T Pv[6];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Pv[i] = PP[i] + PQ[i];
        Pv[i + 3] = PP[i] - PQ[i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        y[i] = Pv[i] + Q[i];
        y[i + 6] = Pv[i] - Q[i];
    }
    y[0] += 8 * x[0];
}

// L + H * 2^32 mod p for L, H < 2^43: 1 IMAD.WIDE + 4 ALU
GL_D u64 poseidon_join64(u64 L, u64 H) {
    const u32 H0 = (u32)H, H1 = (u32)(H >> 32);
    const u64 s = (u64)H1 * 0xFFFFFFFFu + L;     // H1 * 2^64 = H1 * EPS; no overflow
    const u32 s0 = (u32)s, s1 = (u32)(s >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 t1, c, m;\n\t"
        "add.cc.u32  t1, %3, %4;\n\t"
        "addc.u32    c, 0, 0;\n\t"
        "neg.s32     m, c;\n\t"
        "add.cc.u32  %0, %2, m;\n\t"
        "addc.u32    %1, t1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(s0), "r"(s1), "r"(H0));
    return ((u64)r1 << 32) | r0;
}

// MDS + next round's constants (Montgomery form), state lazy in / lazy out
GL_D void poseidon_mds_mont(u64 x[12], const u64* __restrict__ rc) {
    u64 lo[12], hi[12], yl[12], yh[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        lo[j] = (u64)(u32)x[j];
        hi[j] = x[j] >> 32;
    }
    poseidon_mds_lanes<u64>(yl, lo);
    poseidon_mds_lanes<u64>(yh, hi);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const u64 c = rc[i];
        x[i] = poseidon_join64(yl[i] + (u64)(u32)c, yh[i] + (c >> 32));
    }
}

// Permutation on a Montgomery-form state (x~ = x * 2^64 mod p, any 64-bit representative).
GL_D void poseidon_permute_lanes64(u64 x[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_addc(x[i], POSEIDON_RC_MONT[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        x[0] = poseidon_sbox_mont(x[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) x[i] = poseidon_sbox_mont(x[i]);
        }
        poseidon_mds_mont(x, POSEIDON_RC_MONT + 12 * (r + 1));
    }
}

