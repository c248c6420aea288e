// Candidate P5: shipped Montgomery permutation with every two-input add of the MDS forced onto the alu pipe
// (IADD3 with an opaque zero from constant memory as third operand; ptxas cannot turn that into IMAD.IADD).
#pragma once
#include "poseidon.cuh"
__constant__ u32 PZERO = 0;
GL_D void poseidon_mds_limb_z(u32 y[12], const u32 x[12], const u32* __restrict__ rc) {
    const u32 Z = PZERO;
    u32 xp[6], xm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        xp[i] = x[i] + x[i + 6] + Z;
        xm[i] = x[i] - x[i + 6] + Z;
    }
    // Q = negacyclic6([2,-4,16,1,-1,-1], xm):  Q_i = sum_e km_e * (+-) xm[(i-e) mod 6], sign flips on wrap
    u32 Q[6];
    Q[0] = 2 * xm[0] + 4 * xm[5] - 16 * xm[4] - xm[3] + xm[2] + xm[1];
    Q[1] = 2 * xm[1] - 4 * xm[0] - 16 * xm[5] - xm[4] + xm[3] + xm[2];
    Q[2] = 2 * xm[2] - 4 * xm[1] + 16 * xm[0] - xm[5] + xm[4] + xm[3];
    Q[3] = 2 * xm[3] - 4 * xm[2] + 16 * xm[1] + xm[0] + xm[5] + xm[4];
    Q[4] = 2 * xm[4] - 4 * xm[3] + 16 * xm[2] + xm[1] - xm[0] + xm[5];
    Q[5] = 2 * xm[5] - 4 * xm[4] + 16 * xm[3] + xm[2] - xm[1] - xm[0];
    u32 xpp[3], xpm[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        xpp[i] = xp[i] + xp[i + 3] + Z;
        xpm[i] = xp[i] - xp[i + 3] + Z;
    }
    // PP = cyclic3([16,32,16], xpp) = 16*(s + xpp[i-1]),  s = xpp0+xpp1+xpp2
    const u32 s = xpp[0] + xpp[1] + xpp[2];
    u32 PP[3] = {16 * (s + xpp[2]), 16 * (s + xpp[0]), 16 * (s + xpp[1])};
    // PQ = negacyclic3([-1,-8,2], xpm)
    u32 PQ[3];
    PQ[0] = 8 * xpm[2] - xpm[0] - 2 * xpm[1];
    PQ[1] = 0u - xpm[1] - 8 * xpm[0] - 2 * xpm[2];
    PQ[2] = 2 * xpm[0] - xpm[2] - 8 * xpm[1];
    u32 Pv[6];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Pv[i] = PP[i] + PQ[i] + Z;
        Pv[i + 3] = PP[i] - PQ[i] + Z;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        y[i] = Pv[i] + Q[i] + rc[3 * i];
        y[i + 6] = Pv[i] - Q[i] + rc[3 * (i + 6)];
    }
    y[0] += 8 * x[0];
}


GL_D void poseidon_mds_z(u64 x[12], const u32* __restrict__ rc_limbs) {
    u32 a[12], b[12], c[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const u32 lo = (u32)x[j], hi = (u32)(x[j] >> 32);
        a[j] = lo & 0x3FFFFFu;
        b[j] = __funnelshift_r(lo, hi, 22) & 0x3FFFFFu;
        c[j] = hi >> 12;
    }
    u32 ya[12], yb[12], yc[12];
    poseidon_mds_limb_z(ya, a, rc_limbs);
    poseidon_mds_limb_z(yb, b, rc_limbs + 1);
    poseidon_mds_limb_z(yc, c, rc_limbs + 2);
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
}
GL_D void poseidon_permute_mont_z(u64 x[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_addc(x[i], POSEIDON_RC0[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        x[0] = poseidon_sbox(x[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) x[i] = poseidon_sbox(x[i]);
        }
        poseidon_mds_z(x, POSEIDON_RC_LIMBS + r * 36);
    }
}
