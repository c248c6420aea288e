// Candidate: limb-resident partial rounds.  The state crosses every round boundary as three limb planes (22/22/20 bits,
// the MDS works on them anyway).  A full round joins every lane back to a u64 for the S-box and splits it again (same
// instruction count as the shipped sbox -> split -> MDS -> join order); a PARTIAL round does that for lane 0 only and
// merely carry-normalises the other 11 lanes (10 ALU instead of join 11 + split 4, and no IMAD.WIDE).
// Normalisation keeps limbs non-negative by adding 2^12 to limb a; the resulting constant M * (0, 2^12, ..., 2^12) is
// folded into the round constants of rounds 5..26 (table POSEIDON_RC_LIMBS_LR, tools/gen_poseidon_rc.py).
#pragma once
#include "poseidon.cuh"

__constant__ u32 POSEIDON_RC_LIMBS_LR[22 * 36] = {
#include "poseidon_rc_limbs_lr.inc"
};

GL_D void poseidon_split(u64 x, u32& a, u32& b, u32& c) {
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    a = lo & 0x3FFFFFu;
    b = __funnelshift_r(lo, hi, 22) & 0x3FFFFFu;
    c = hi >> 12;
}

// (Y0, Y1, Y2) with Y0 < 2^31, Y1 < 2^32 - 2^10, Y2 < 2^31  ->  limbs of the same value + 2^12 (mod p):
// a in (0, 2^22 + 2^12], b < 2^23, c < 2^20.
GL_D void poseidon_renorm(u32 Y0, u32 Y1, u32 Y2, u32& a, u32& b, u32& c) {
    const u32 t1 = Y1 + (Y0 >> 22);
    const u32 t2 = Y2 + (t1 >> 22);
    const u32 top = t2 >> 20;                     // * 2^64 = top * 2^32 - top
    a = (Y0 & 0x3FFFFFu) - top + 4096u;
    b = (t1 & 0x3FFFFFu) + (top << 10);
    c = t2 & 0xFFFFFu;
}

GL_D void poseidon_permute_mont_lr(u64 x[12]) {
    u32 ya[12], yb[12], yc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) poseidon_split(gl_addc(x[i], POSEIDON_RC0[i]), ya[i], yb[i], yc[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        u32 a[12], b[12], c[12];
        const bool full = (r < 4 || r >= 26);
        poseidon_split(poseidon_sbox(poseidon_join(ya[0], yb[0], yc[0])), a[0], b[0], c[0]);
        if (full) {
#pragma unroll
            for (int i = 1; i < 12; i++) poseidon_split(poseidon_sbox(poseidon_join(ya[i], yb[i], yc[i])), a[i], b[i], c[i]);
        } else {
#pragma unroll
            for (int i = 1; i < 12; i++) poseidon_renorm(ya[i], yb[i], yc[i], a[i], b[i], c[i]);
        }
        const u32* __restrict__ rc = full ? (POSEIDON_RC_LIMBS + r * 36) : (POSEIDON_RC_LIMBS_LR + (r - 4) * 36);
        poseidon_mds_limb(ya, a, rc);
        poseidon_mds_limb(yb, b, rc + 1);
        poseidon_mds_limb(yc, c, rc + 2);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
}
