// Candidate P3: Montgomery S-box + 3x22-bit-limb MDS with the adds split between the ALU pipe (IADD3/LEA) and the
// FMA pipe (IMAD with an opaque multiplier from constant memory, which ptxas cannot strength-reduce to IADD3).
#pragma once
#include "poseidon_mont.cuh"

__constant__ u32 PSTEER[8] = {1u, 0xFFFFFFFFu, 2u, 4u, 8u, 16u, 0xFFFFFFFEu, 0xFFFFFFF8u};
// fma-pipe helpers: d = a * K + b
GL_D u32 f_mad(u32 a, u32 k, u32 b) { u32 d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(k), "r"(b)); return d; }
#define F_ADD(a, b) f_mad((a), PSTEER[0], (b))            /* b + a      */
#define F_SUB(b, a) f_mad((a), PSTEER[1], (b))            /* b - a      */
#define F_MAD2(a, b) f_mad((a), PSTEER[2], (b))           /* b + 2a     */
#define F_MAD4(a, b) f_mad((a), PSTEER[3], (b))
#define F_MAD8(a, b) f_mad((a), PSTEER[4], (b))
#define F_MAD16(a, b) f_mad((a), PSTEER[5], (b))
#define F_MSUB2(a, b) f_mad((a), PSTEER[6], (b))          /* b - 2a     */
#define F_MSUB8(a, b) f_mad((a), PSTEER[7], (b))          /* b - 8a     */

// One limb plane.  STEER = 0: everything left to the compiler (ALU pipe in practice); STEER = 1: about half of the
// operations forced onto the FMA pipe.
template <int STEER>
GL_D void poseidon_mds_limb_s(u32 y[12], const u32 x[12], const u32* __restrict__ rc, int stride) {
    u32 xp[6], xm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        xp[i] = STEER ? F_ADD(x[i], x[i + 6]) : x[i] + x[i + 6];
        xm[i] = x[i] - x[i + 6];
    }
    u32 Q[6];
    if (STEER) {
        // Q_i = 2 xm_i -+ 4 xm_{i-1} -+ 16 xm_{i-2} -+ xm_{i-3} +- ... : half of each chain on the FMA pipe
        Q[0] = F_MAD4(xm[5], F_MAD2(xm[0], xm[2] + xm[1])) - (16 * xm[4] + xm[3]);
        Q[1] = F_MAD2(xm[1], xm[3] + xm[2]) - F_MAD16(xm[5], F_MAD4(xm[0], xm[4]));
        Q[2] = F_MAD16(xm[0], F_MAD2(xm[2], xm[4] + xm[3])) - (4 * xm[1] + xm[5]);
        Q[3] = F_MAD16(xm[1], F_MAD2(xm[3], xm[0] + xm[5])) + (xm[4] - 4 * xm[2]);
        Q[4] = F_MAD16(xm[2], F_MAD2(xm[4], xm[1] + xm[5])) - (4 * xm[3] + xm[0]);
        Q[5] = F_MAD16(xm[3], F_MAD2(xm[5], xm[2] - xm[1])) - (4 * xm[4] + xm[0]);
    } else {
        Q[0] = 2 * xm[0] + 4 * xm[5] - 16 * xm[4] - xm[3] + xm[2] + xm[1];
        Q[1] = 2 * xm[1] - 4 * xm[0] - 16 * xm[5] - xm[4] + xm[3] + xm[2];
        Q[2] = 2 * xm[2] - 4 * xm[1] + 16 * xm[0] - xm[5] + xm[4] + xm[3];
        Q[3] = 2 * xm[3] - 4 * xm[2] + 16 * xm[1] + xm[0] + xm[5] + xm[4];
        Q[4] = 2 * xm[4] - 4 * xm[3] + 16 * xm[2] + xm[1] - xm[0] + xm[5];
        Q[5] = 2 * xm[5] - 4 * xm[4] + 16 * xm[3] + xm[2] - xm[1] - xm[0];
    }
    u32 xpp[3], xpm[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        xpp[i] = xp[i] + xp[i + 3];
        xpm[i] = STEER ? F_SUB(xp[i], xp[i + 3]) : xp[i] - xp[i + 3];
    }
    const u32 s = xpp[0] + xpp[1] + xpp[2];
    u32 PP[3], PQ[3];
    if (STEER) {
        // PP_i = 16 (s + xpp_{i-1}) + rc  (the round constant rides on the FMA-pipe multiply-add)
        PP[0] = F_MAD16(s + xpp[2], 0u);
        PP[1] = F_MAD16(s + xpp[0], 0u);
        PP[2] = F_MAD16(s + xpp[1], 0u);
        PQ[0] = F_MSUB2(xpm[1], 8 * xpm[2] - xpm[0]);
        PQ[1] = F_MSUB8(xpm[0], 0u - xpm[1] - 2 * xpm[2]);
        PQ[2] = F_MSUB8(xpm[1], 2 * xpm[0] - xpm[2]);
    } else {
        PP[0] = 16 * (s + xpp[2]); PP[1] = 16 * (s + xpp[0]); PP[2] = 16 * (s + xpp[1]);
        PQ[0] = 8 * xpm[2] - xpm[0] - 2 * xpm[1];
        PQ[1] = 0u - xpm[1] - 8 * xpm[0] - 2 * xpm[2];
        PQ[2] = 2 * xpm[0] - xpm[2] - 8 * xpm[1];
    }
    u32 Pv[6];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Pv[i] = PP[i] + PQ[i];
        Pv[i + 3] = STEER ? F_SUB(PP[i], PQ[i]) : PP[i] - PQ[i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        // 3-input adds (IADD3) fold the round constant in for free on the ALU pipe
        y[i] = Pv[i] + Q[i] + rc[i * stride];
        y[i + 6] = STEER ? F_SUB(Pv[i] + rc[(i + 6) * stride], Q[i]) : Pv[i] - Q[i] + rc[(i + 6) * stride];
    }
    y[0] = STEER ? F_MAD8(x[0], y[0]) : y[0] + 8 * x[0];
}

__constant__ u32 POSEIDON_RC_MONT_LIMBS[31 * 36] = {
#include "poseidon_rc_mont_limbs.inc"
};

// limbs Y0 + Y1*2^22 + Y2*2^44 (each < 2^32) -> lazy u64:  2 IMAD.WIDE + 1 IMAD.WIDE + carry fix
GL_D u64 poseidon_join3(u32 Y0, u32 Y1, u32 Y2) {
    const u64 v = (u64)Y1 * (1u << 22) + Y0;             // < 2^55
    const u64 w = (u64)Y2 * (1u << 12);                  // Y2 * 2^44 = w * 2^32, w < 2^44
    const u32 w0 = (u32)w, w1 = (u32)(w >> 32);          // value = v + w0 * 2^32 + w1 * 2^64
    const u64 s = (u64)w1 * 0xFFFFFFFFu + v;             // w1 * 2^64 = w1 * EPS (w1 < 2^12): no overflow
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
        : "r"(s0), "r"(s1), "r"(w0));
    return ((u64)r1 << 32) | r0;
}

template <int STEER>
GL_D void poseidon_mds_steer(u64 x[12], const u32* __restrict__ rc_limbs) {
    u32 a[12], b[12], c[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const u32 lo = (u32)x[j], hi = (u32)(x[j] >> 32);
        a[j] = lo & 0x3FFFFFu;
        b[j] = __funnelshift_r(lo, hi, 22) & 0x3FFFFFu;
        c[j] = hi >> 12;
    }
    u32 ya[12], yb[12], yc[12];
    poseidon_mds_limb_s<STEER>(ya, a, rc_limbs, 3);
    poseidon_mds_limb_s<STEER>(yb, b, rc_limbs + 1, 3);
    poseidon_mds_limb_s<STEER>(yc, c, rc_limbs + 2, 3);
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = poseidon_join3(ya[i], yb[i], yc[i]);
}

template <int STEER>
GL_D void poseidon_permute_steer(u64 x[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_addc(x[i], POSEIDON_RC_MONT[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        x[0] = poseidon_sbox_mont(x[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) x[i] = poseidon_sbox_mont(x[i]);
        }
        poseidon_mds_steer<STEER>(x, POSEIDON_RC_MONT_LIMBS + 36 * r);
    }
}
