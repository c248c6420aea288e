// Poseidon-GL permutation (width 12 = rate 8 + capacity 4, x^7, 8 full + 22 partial rounds) for sm_100a.
//
// Functionally identical to the reference's two formulations, which agree with each other:
//   plain 30-round form      src/helpers/glwasm.js:359-390 (MDS :428-440, round constants :536-627)
//   optimised (sparse) form  src/helpers/hash/poseidon/poseidon.js:57-108
// One permutation per thread, the 12-word state lives in 24 registers, round constants come from
// __constant__ memory (warp-uniform index -> constant-cache broadcast).  The kernel is bound by the two
// half-rate integer pipes (alu: IADD3/LOP3/SHF, fmaheavy: IMAD/IMAD.WIDE), not by HBM: see DESIGN.md "Poseidon".
#pragma once
#include "gl.cuh"

// Round-0 constants (added before the first S-box layer), Montgomery form (generated: tools/gen_poseidon_rc.py)
__constant__ u64 POSEIDON_RC0[12] = {
#include "poseidon_rc.inc"
};

// The state is kept in Montgomery form (x * 2^64 mod p, any 64-bit representative): the S-box then costs 4 gl_mmul whose
// reductions are 9 ALU-pipe instructions each instead of 12, and the linear layer does not care about the scaling.
// x -> x^7 : 4 multiplications (2 of them squarings)
GL_D u64 poseidon_sbox(u64 x) {
    u64 x2 = gl_msqr(x);
    u64 x3 = gl_mmul(x2, x);
    u64 x4 = gl_msqr(x2);
    return gl_mmul(x3, x4);
}

// ---------------------------------------------------------------------------------------------------------------
// Dense MDS layer: out_i = sum_j circ[(j - i) mod 12] * x_j  (+ 8*x_0 on lane 0), glwasm.js:428-440.
//
// The 144 small multiplies are NOT done with multiplies.  The matrix is circulant with kernel
// k = [17,20,34,18,39,13,13,28,2,16,41,15] (y = k (*) x mod z^12 - 1), and its CRT split over
// z^12-1 = (z^6-1)(z^6+1) = (z^3-1)(z^3+1)(z^6+1) has power-of-two entries:
//     (k_lo + k_hi)/2 = [15,24,18,17,40,14]  ->  (.)/2 split again: [16,32,16] (cyclic 3) and [-1,-8,2] (negacyclic 3)
//     (k_lo - k_hi)/2 = [2,-4,16,1,-1,-1]                            (negacyclic 6)
// so one MDS on a vector of small integers is ~76 shift-adds, which ptxas spreads over the alu pipe (IADD3/LEA) and the
// fmaheavy pipe (IMAD.IADD / IMAD with a power-of-two immediate).  Each state word is cut into three limbs of 22/22/20
// bits (up to 2^23 after a limb-form re-normalisation, poseidon_renorm); limb sums stay below 264 * 2^23 + 2^22 < 2^32, so plain
// wrap-around u32 arithmetic is exact.  The round constant of
// the NEXT round rides on the last butterfly as the third operand of an IADD3.
// Alternatives measured on B200 (tools/probe/gl_probe.cu, Gperm/s): this 1.26; two 64-bit lanes with carries 1.20; adds
// forced onto the fmaheavy pipe 1.20; the same CRT on the FP64 pipe (exact, DFMA/DADD) 1.22; previous non-Montgomery 1.11.
// ---------------------------------------------------------------------------------------------------------------
GL_D void poseidon_mds_limb(u32 y[12], const u32 x[12], const u32* __restrict__ rc) {
    u32 xp[6], xm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        xp[i] = x[i] + x[i + 6];
        xm[i] = x[i] - x[i + 6];
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
        xpp[i] = xp[i] + xp[i + 3];
        xpm[i] = xp[i] - xp[i + 3];
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
        Pv[i] = PP[i] + PQ[i];
        Pv[i + 3] = PP[i] - PQ[i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        y[i] = Pv[i] + Q[i] + rc[3 * i];
        y[i + 6] = Pv[i] - Q[i] + rc[3 * (i + 6)];
    }
    y[0] += 8 * x[0];
}

// limbs -> field element: Y0 + Y1*2^22 + Y2*2^44 (any u32 limbs) mod p, with 2^64 = EPS.  3 IMAD.WIDE + 5 ALU.
GL_D u64 poseidon_join(u32 Y0, u32 Y1, u32 Y2) {
    const u64 v = (u64)Y1 * (1u << 22) + Y0;             // < 2^55
    const u64 w = (u64)Y2 * (1u << 12);                  // Y2 * 2^44 = w * 2^32, w < 2^44
    const u32 w0 = (u32)w, w1 = (u32)(w >> 32);          // value = v + w0 * 2^32 + w1 * 2^64
    const u64 s = (u64)w1 * 0xFFFFFFFFu + v;             // w1 * 2^64 = w1 * EPS (w1 < 2^12): no overflow
    const u32 s0 = (u32)s, s1 = (u32)(s >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 t1, c, m;\n\t"
        "add.cc.u32  t1, %3, %4;\n\t"      // + w0 * 2^32
        "addc.u32    c, 0, 0;\n\t"
        "neg.s32     m, c;\n\t"            // carry ? 0xFFFFFFFF : 0
        "add.cc.u32  %0, %2, m;\n\t"       // + carry * EPS (the wrapped value is < 2^56: no second wrap)
        "addc.u32    %1, t1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(s0), "r"(s1), "r"(w0));
    return ((u64)r1 << 32) | r0;
}

// u64 -> limbs (22/22/20 bits).  4 ALU.
GL_D void poseidon_split(u64 x, u32& a, u32& b, u32& c) {
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    a = lo & 0x3FFFFFu;
    b = __funnelshift_r(lo, hi, 22) & 0x3FFFFFu;
    c = hi >> 12;
}

// Carry re-normalisation of a lane that stays in limb form across a partial round: MDS outputs (Y0 < 2^31, Y1 < 2^32 - 2^10,
// Y2 < 2^31) -> limbs of the same value + 2^12 (mod p) with a in (0, 2^22 + 2^12], b < 2^23, c < 2^20, small enough for the
// next MDS to stay below 2^32 per limb (264 * 2^23 + 2^22 < 2^32).  The top carry folds with 2^64 = 2^32 - 1: + top * 2^10 on
// limb b, - top on limb a; the + 2^12 keeps limb a positive and its image under the MDS is already taken out of the round
// constants (POSEIDON_RC_LIMBS_PARTIAL).  8 ALU-pipe instructions, against join (3 IMAD.WIDE + 5) + split (4).
GL_D void poseidon_renorm(u32 Y0, u32 Y1, u32 Y2, u32& a, u32& b, u32& c) {
    const u32 t1 = Y1 + (Y0 >> 22);
    const u32 t2 = Y2 + (t1 >> 22);
    const u32 top = t2 >> 20;
    a = (Y0 & 0x3FFFFFu) - top + 4096u;
    b = (t1 & 0x3FFFFFu) + (top << 10);
    c = t2 & 0xFFFFFu;
}

// Montgomery-form constants of round r+1 pre-split into limbs, added after the MDS of round r (row 29 = 0), and the
// bias-compensated rows used after the MDS of the partial rounds 4..25 (see poseidon_renorm).
__constant__ u32 POSEIDON_RC_LIMBS[30 * 36] = {
#include "poseidon_rc_limbs.inc"
};
__constant__ u32 POSEIDON_RC_LIMBS_PARTIAL[22 * 36] = {
#include "poseidon_rc_limbs_partial.inc"
};

// Permutation of a MONTGOMERY-FORM state (x * 2^64 mod p, any 64-bit representative in and out).
// The state crosses every round boundary as three limb planes -- the form the MDS works on.  A full round joins every lane
// back to a u64 for the S-box and splits it again; a partial round does that for lane 0 only and re-normalises the other 11
// lanes in limb form (measured, tools/probe: 1.260 -> 1.288 Gperm/s).  One 30-iteration loop keeps a single copy of the
// S-box layer and of the MDS in the instruction cache.
GL_D void poseidon_permute_mont_limb(u64 x[12]) {
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
        const u32* __restrict__ rc = full ? (POSEIDON_RC_LIMBS + r * 36) : (POSEIDON_RC_LIMBS_PARTIAL + (r - 4) * 36);
        poseidon_mds_limb(ya, a, rc);
        poseidon_mds_limb(yb, b, rc + 1);
        poseidon_mds_limb(yc, c, rc + 2);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// FP64-resident partial rounds.
//
// In the 22 partial rounds only lane 0 sees an S-box; lanes 1..11 are mapped linearly from round to round.  They therefore
// leave the integer pipes altogether: each lane is kept as two planes of doubles holding exact integers, value = L + 2^32 H
// (mod p), and the MDS (the same CRT network as poseidon_mds_limb, 78 DADD/DFMA per plane -- the power-of-two coefficients
// are the multiplicands of the fused multiply-adds) runs on the FP64 pipe, which B200 issues at 16 lanes/clk/SMSP next to
// the alu and fmaheavy pipes and which is otherwise idle.  Signed intermediates are native, so no bias bookkeeping; the MDS
// grows a plane by at most 264x per round, so from |L|, |H| <= 2^32 two rounds stay below 2^49 (exact in a double) and the
// planes are re-normalised every second round (poseidon_renorm_f64, 8 FP64 instructions per lane).  Only lane 0 crosses
// pipes: double -> u64 before its S-box (2 DADD whose addend carries the magic number AND the round constant, 2 LOP3,
// poseidon_join_planes) and u64 -> double after it.  The round constants of lanes 1..11 are deferred through the linear map
// (tools/gen_poseidon_f64_consts.py): a round adds one scalar to lane 0, the accumulated vector comes back after round 25.
// ---------------------------------------------------------------------------------------------------------------
#include "poseidon_rc_f64p.inc"      // POSEIDON_RC_F64P_<NR>[2 * (NR + 11)] for the candidate block lengths NR
template <int NR> GL_D const double* poseidon_rc_f64p();
#define POSEIDON_F64P_TABLE(NR) template <> GL_D const double* poseidon_rc_f64p<NR>() { return POSEIDON_RC_F64P_##NR; }
POSEIDON_F64P_TABLE(6) POSEIDON_F64P_TABLE(8) POSEIDON_F64P_TABLE(10) POSEIDON_F64P_TABLE(12) POSEIDON_F64P_TABLE(14)
POSEIDON_F64P_TABLE(16) POSEIDON_F64P_TABLE(18) POSEIDON_F64P_TABLE(22)

GL_D void poseidon_mds_f64(double y[12], const double x[12]) {
    double xp[6], xm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        xp[i] = __dadd_rn(x[i], x[i + 6]);
        xm[i] = __dadd_rn(x[i], -x[i + 6]);
    }
    double Q[6];
    Q[0] = __fma_rn(2.0, xm[0], __fma_rn(4.0, xm[5], __fma_rn(-16.0, xm[4], __dadd_rn(__dadd_rn(xm[2], xm[1]), -xm[3]))));
    Q[1] = __fma_rn(2.0, xm[1], __fma_rn(-4.0, xm[0], __fma_rn(-16.0, xm[5], __dadd_rn(__dadd_rn(xm[3], xm[2]), -xm[4]))));
    Q[2] = __fma_rn(2.0, xm[2], __fma_rn(-4.0, xm[1], __fma_rn(16.0, xm[0], __dadd_rn(__dadd_rn(xm[4], xm[3]), -xm[5]))));
    Q[3] = __fma_rn(2.0, xm[3], __fma_rn(-4.0, xm[2], __fma_rn(16.0, xm[1], __dadd_rn(__dadd_rn(xm[0], xm[5]), xm[4]))));
    Q[4] = __fma_rn(2.0, xm[4], __fma_rn(-4.0, xm[3], __fma_rn(16.0, xm[2], __dadd_rn(__dadd_rn(xm[1], -xm[0]), xm[5]))));
    Q[5] = __fma_rn(2.0, xm[5], __fma_rn(-4.0, xm[4], __fma_rn(16.0, xm[3], __dadd_rn(__dadd_rn(xm[2], -xm[1]), -xm[0]))));
    double xpp[3], xpm[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        xpp[i] = __dadd_rn(xp[i], xp[i + 3]);
        xpm[i] = __dadd_rn(xp[i], -xp[i + 3]);
    }
    const double s = __dadd_rn(__dadd_rn(xpp[0], xpp[1]), xpp[2]);
    const double t[3] = {__dadd_rn(s, xpp[2]), __dadd_rn(s, xpp[0]), __dadd_rn(s, xpp[1])};
    double PQ[3];
    PQ[0] = __fma_rn(8.0, xpm[2], __fma_rn(-2.0, xpm[1], -xpm[0]));
    PQ[1] = __fma_rn(-8.0, xpm[0], __fma_rn(-2.0, xpm[2], -xpm[1]));
    PQ[2] = __fma_rn(2.0, xpm[0], __fma_rn(-8.0, xpm[1], -xpm[2]));
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double pa = __fma_rn(16.0, t[i], PQ[i]), pb = __fma_rn(16.0, t[i], -PQ[i]);
        y[i] = __dadd_rn(pa, Q[i]);
        y[i + 6] = __dadd_rn(pa, -Q[i]);
        y[i + 3] = __dadd_rn(pb, Q[i + 3]);
        y[i + 9] = __dadd_rn(pb, -Q[i + 3]);
    }
    y[0] = __fma_rn(8.0, x[0], y[0]);
}

// integer < 2^32 -> double (exact): 2^52 + v in the mantissa, minus 2^52.  1 MOV-class + 1 DADD.
GL_D double poseidon_u32_to_f64(u32 v) { return __dadd_rn(__hiloint2double(0x43300000, (int)v), -4503599627370496.0); }

// L + H * 2^32 mod p for L, H < 2^52: 1 IMAD.WIDE + 5 ALU (H = H1 2^32 + H0; H1 2^64 = H1 EPS < 2^52, no overflow; the wrap of
// the middle word folds once and cannot wrap again because the wrapped word is < 2^21).
GL_D u64 poseidon_join_planes(u64 L, u64 H) {
    const u32 H0 = (u32)H, H1 = (u32)(H >> 32);
    const u64 s = (u64)H1 * 0xFFFFFFFFu + L;
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

// Plane values (|yl|, |yh| < 2^50, signed) + constant -> u64 representative of yl + 2^32 yh + c.  cl / ch = 2^52 + 2^51 + the
// halves of c - (2^51 + 2^83): the sums lie in [2^52, 2^53), where the low 52 bits of the double ARE the integer.
GL_D u64 poseidon_f64_to_u64(double yl, double yh, double cl, double ch) {
    const double tl = __dadd_rn(yl, cl), th = __dadd_rn(yh, ch);
    const u64 L = ((u64)((u32)__double2hiint(tl) & 0xFFFFFu) << 32) | (u32)__double2loint(tl);
    const u64 H = ((u64)((u32)__double2hiint(th) & 0xFFFFFu) << 32) | (u32)__double2loint(th);
    return poseidon_join_planes(L, H);
}

// (L, H) with |L|, |H| < 2^50 -> same value mod p with |L| <= 2^31, |H| <= 2^31 + 2^19.  hh = round(H / 2^32) folds with
// 2^64 = 2^32 - 1 (H <- H - hh 2^32 + hh, L <- L - hh), then the carry lh = round(L / 2^32) of the low plane moves up.
GL_D void poseidon_renorm_f64(double& L, double& H) {
    const double C = 6755399441055744.0, I32 = 2.3283064365386963e-10;   // 1.5 * 2^52 (rounds to an integer), 2^-32
    const double hh = __dadd_rn(__fma_rn(H, I32, C), -C);
    H = __fma_rn(hh, -4294967295.0, H);
    L = __dadd_rn(L, -hh);
    const double lh = __dadd_rn(__fma_rn(L, I32, C), -C);
    L = __fma_rn(lh, -4294967296.0, L);
    H = __dadd_rn(H, lh);
}

// Rounds 4..3+NR: x = S-box inputs of round 4 (Montgomery form, any representatives) -> S-box inputs of round 4+NR.
// One round per loop iteration (the re-normalisation of every second round sits behind a warp-uniform branch): half the code of
// a two-round body, which keeps the whole hashing kernel inside the 32 KB L1.5 instruction cache.
#ifndef POSEIDON_F64P_UNROLL
#define POSEIDON_F64P_UNROLL 1
#endif
template <int NR>
GL_D void poseidon_partial_f64(u64 x[12]) {
    static_assert(NR % 2 == 0 && NR >= 2 && NR <= 22, "the FP64 block re-normalises every second round");
    const double* __restrict__ RCT = poseidon_rc_f64p<NR>();
    double L[12], H[12];
#pragma unroll
    for (int i = 1; i < 12; i++) {
        L[i] = poseidon_u32_to_f64((u32)x[i]);
        H[i] = poseidon_u32_to_f64((u32)(x[i] >> 32));
    }
    u64 x0 = x[0];
#pragma unroll 1
    for (int r = 0; r < NR; r += POSEIDON_F64P_UNROLL) {
#pragma unroll
        for (int h = 0; h < POSEIDON_F64P_UNROLL; h++) {
            const u64 a = poseidon_sbox(x0);
            L[0] = poseidon_u32_to_f64((u32)a);
            H[0] = poseidon_u32_to_f64((u32)(a >> 32));
            double yl[12], yh[12];
            poseidon_mds_f64(yl, L);
            poseidon_mds_f64(yh, H);
            const double* __restrict__ c = RCT + 2 * (r + h);
            x0 = poseidon_f64_to_u64(yl[0], yh[0], c[0], c[1]);
#pragma unroll
            for (int i = 1; i < 12; i++) { L[i] = yl[i]; H[i] = yh[i]; }
        }
        if (POSEIDON_F64P_UNROLL == 2 ? (r < NR - 2) : ((r & 1) && r < NR - 1)) {
#pragma unroll
            for (int i = 1; i < 12; i++) poseidon_renorm_f64(L[i], H[i]);
        }
    }
    x[0] = x0;
#pragma unroll
    for (int i = 1; i < 12; i++) x[i] = poseidon_f64_to_u64(L[i], H[i], RCT[2 * (NR - 1) + 2 * i], RCT[2 * (NR - 1) + 2 * i + 1]);
}

// The permutation with NR of the 22 partial rounds on the FP64 pipe: rounds 0..3 and 4+NR..29 run the limb-plane loop body of
// poseidon_permute_mont_limb (one copy of the code, executed for both stretches), rounds 4..3+NR poseidon_partial_f64.  NR
// balances the machine: an FP64 instruction takes two issue slots of the sub-partition (measured: tools/probe, "pipe mix"), so
// with every partial round on the FP64 pipe the kernel becomes issue-bound while the integer pipes idle, and with none the
// fmaheavy pipe is the limit.
template <int NR>
GL_D void poseidon_permute_mont_f64p(u64 x[12]) {
    u32 ya[12], yb[12], yc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) poseidon_split(gl_addc(x[i], POSEIDON_RC0[i]), ya[i], yb[i], yc[i]);
#pragma unroll 1
    for (int ph = 0; ph < 2; ph++) {
        const int r_end = ph ? 30 : 4;
#pragma unroll 1
        for (int r = ph ? 4 + NR : 0; r < r_end; r++) {
            u32 a[12], b[12], c[12];
            const bool full = (NR == 22) || (r < 4 || r >= 26);      // NR == 22: no limb-form partial rounds, the loop body is full rounds only
            poseidon_split(poseidon_sbox(poseidon_join(ya[0], yb[0], yc[0])), a[0], b[0], c[0]);
            if (full) {
#pragma unroll
                for (int i = 1; i < 12; i++) poseidon_split(poseidon_sbox(poseidon_join(ya[i], yb[i], yc[i])), a[i], b[i], c[i]);
            } else {
#pragma unroll
                for (int i = 1; i < 12; i++) poseidon_renorm(ya[i], yb[i], yc[i], a[i], b[i], c[i]);
            }
            const u32* __restrict__ rc = full ? (POSEIDON_RC_LIMBS + r * 36) : (POSEIDON_RC_LIMBS_PARTIAL + (r - 4) * 36);
            poseidon_mds_limb(ya, a, rc);
            poseidon_mds_limb(yb, b, rc + 1);
            poseidon_mds_limb(yc, c, rc + 2);
        }
        if (ph == 0) {
#pragma unroll
            for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
            poseidon_partial_f64<NR>(x);
#pragma unroll
            for (int i = 0; i < 12; i++) poseidon_split(x[i], ya[i], yb[i], yc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
}

// The permutation every hashing kernel calls.  POSEIDON_F64P = number of partial rounds on the FP64 pipe (build flag for A/B runs;
// 0 = the all-integer loop).  Measured on B200 (tools/probe/gl_probe.cu, profiles/r02_poseidon_f64.md): 22 -> 1.416 Gperm/s against
// 1.289 for the all-integer loop, every shorter FP64 block in between.
#ifndef POSEIDON_F64P
#define POSEIDON_F64P 22
#endif
GL_D void poseidon_permute_mont(u64 x[12]) {
#if POSEIDON_F64P > 0
    poseidon_permute_mont_f64p<POSEIDON_F64P>(x);
#else
    poseidon_permute_mont_limb(x);
#endif
}

// Permutation of a plain state; output canonical.  (Test hook / transcript: hashing kernels stay in Montgomery form.)
GL_D void poseidon_permute(u64 x[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_to_mont(x[i]);
    poseidon_permute_mont(x);
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_from_mont(x[i]);
}
