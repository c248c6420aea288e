// Poseidon-GL with the 22 partial rounds in tensor-core form (warp-collective: the 32 lanes of a warp permute 32 states together).
//
// Same function as poseidon_permute_mont (poseidon.cuh; reference: src/helpers/glwasm.js:359-390, src/helpers/hash/poseidon/poseidon.js:57-108).
// In a partial round only lane 0 meets an S-box.  With N = the MDS matrix with column 0 cleared and v = its column 0,
//     y(n+1) = N y(n) + v s_n + C(5+n),   s_n = y(n)_0 ^ 7                      (y(n) = S-box inputs of round 4+n, n = 0..21)
// so everything but the 22 scalars s_n is linear in y(0) and in the earlier s_k:
//     t_n   = y(n)_0 = <row 0 of N^n, y(0)> + sum_{k<n} mu_(n-k) s_k + const,       mu_d = (N^(d-1) v)_0
//     y(22) = N^22 y(0) + sum_k N^(21-k) v s_k + const
// Constant matrices over F_p times per-permutation vectors are dense contractions: they run on the tensor cores, batched over the warp.
// A vector is taken as its bytes exactly as it lies in shared memory (K index = byte position in the permutation's 288-byte row), a
// constant as the 8 byte limbs of coefficient * 2^(8b) mod p; mma.sync.m16n8k32.u8.u8.s32 sums the byte products exactly (K <= 288:
// sums < 2^25) and the 8 limb sums of an output recombine as sum_b' D_b' 2^(8b') mod p.  The 22 rounds are cut into blocks of 8: the s_k
// of earlier blocks enter t_n through the block's GEMM, those of the block itself through 28 lazily accumulated multiply-adds on the
// integer pipes (mu_1..mu_7 < 2^52).  Against the FP64-resident form (poseidon_partial_f64: 22 dense MDS applications, ~580 issue slots
// per round) the partial rounds cost the 22 S-boxes, 71 multiply-adds, 40 recombinations and ~530 IMMA per 32 permutations.
// Tables, row layout and a lane-exact Python model of this data flow: tools/gen_poseidon_tc_consts.py (its model is checked on the CPU by the
// test suite).
//
// Row of one permutation (PTC_ROW = 288 bytes, 16-byte aligned; 72 words = 8 mod 32 banks, so the A-fragment loads of 4 consecutive
// rows tile the 32 banks):  [0,88) y_1..y_11 | 88: byte 1 (carries the additive constant) | 89..95: 0 | [96 + 8n, +8): slot n = GEMM part
// of t_n, later s_n | [272,288): 0.  The final GEMM writes y(22) over [0,96).
#pragma once
#include "poseidon.cuh"

#include "poseidon_tc_consts.inc"

#define PTC_ROW 288
#define PTC_WARP_BYTES (32 * PTC_ROW)
#define PTC_TABLE_BYTES (POSEIDON_TC_WORDS * 8)
#define PTC_SLOT0 96
#ifndef PTC_HIST_SMEM
#define PTC_HIST_SMEM 1
#endif
#ifndef PTC_UNROLL_J
#define PTC_UNROLL_J 1      // 1: one rolled round body (47 KB kernel, 107 ms at 2^22 x 256); 8: straight-line blocks (72 KB, 121 ms: instruction cache)
#endif

constexpr int ptc_unroll_j = PTC_UNROLL_J;

GL_D void ptc_mma(u32 (&d)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// sum_b' D_b' 2^(8b') mod p for limb sums D < 2^25: two 50-bit planes, then L + 2^32 H.
GL_D u64 ptc_recombine(u32 d0, u32 d1, u32 d2, u32 d3, u32 d4, u32 d5, u32 d6, u32 d7) {
    const u64 L = (u64)d0 + ((u64)d1 << 8) + ((u64)d2 << 16) + ((u64)d3 << 24);
    const u64 H = (u64)d4 + ((u64)d5 << 8) + ((u64)d6 << 16) + ((u64)d7 << 24);
    return poseidon_join_planes(L, H);
}

// Lazy accumulator of lin + sum mu_d s: three 64-bit columns (weights 1, 2^32, 2^64) and the carry counts of the first two.  A product
// is one IMAD.WIDE with carry-out (ptxas fuses mad.lo.cc + madc.hi.cc) and the carries of two products share one IADD3.X.
struct PtcAcc {
    u32 a0l, a0h, c0, a1l, a1h, c1, a2l, a2h;
};
GL_D void ptc_acc_init(PtcAcc& A, u64 lin) {
    A.a0l = (u32)lin;
    A.a0h = (u32)(lin >> 32);
    A.c0 = A.a1l = A.a1h = A.c1 = A.a2l = A.a2h = 0;
}
GL_D void ptc_madc(u32& lo, u32& hi, u32& c, u32 a, u32 b) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(c) : "r"(a), "r"(b));
}
GL_D void ptc_mad(u32& lo, u32& hi, u32 a, u32 b) {      // no overflow possible
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
template <int D>
GL_D void ptc_mac(PtcAcc& A, u64 s) {
    const u64 mu = POSEIDON_TC_MU[D];
    const u32 s0 = (u32)s, s1 = (u32)(s >> 32), m0 = (u32)mu, m1 = (u32)(mu >> 32);
    ptc_madc(A.a0l, A.a0h, A.c0, s0, m0);
    ptc_madc(A.a1l, A.a1h, A.c1, s1, m0);
    if (D >= 5) {                          // mu_1..mu_4 < 2^32
        ptc_madc(A.a1l, A.a1h, A.c1, s0, m1);
        ptc_mad(A.a2l, A.a2h, s1, m1);     // 7 terms < 2^52
    }
}
// a0 + c0 2^64 + (a1 + c1 2^64) 2^32 + a2 2^64 mod p: lo = a0 + (a1l << 32), hi = a2 + a1h + c0 + carry (< 2^56); c1 2^96 = -c1 joins
// the "- h1" term of gl_reduce128 (h1 + c1 < 2^25 keeps its no-second-wrap argument).
GL_D u64 ptc_reduce(const PtcAcc& A) {
    u32 l1, h0, h1;
    asm("{\n\t"
        "add.cc.u32   %0, %3, %4;\n\t"       // a0h + a1l
        "addc.cc.u32  %1, %5, %6;\n\t"       // a2l + a1h + carry
        "addc.u32     %2, %7, 0;\n\t"
        "add.cc.u32   %1, %1, %8;\n\t"       // + c0
        "addc.u32     %2, %2, %9;\n\t"       // + c1 on the top word
        "}"
        : "=&r"(l1), "=&r"(h0), "=&r"(h1)
        : "r"(A.a0h), "r"(A.a1l), "r"(A.a2l), "r"(A.a1h), "r"(A.a2h), "r"(A.c0), "r"(A.c1));
    return gl_reduce128(((u64)h1 << 32) | h0, ((u64)l1 << 32) | A.a0l);
}

// The constants are the A operand: the M index is (limb b', output row j) with m-tile i = limbs 2i (tile rows 0..7 = output rows 0..7)
// and 2i+1 (tile rows 8..15), so lane 4g+t ends up holding all 8 limb sums of output row g, for the permutations 2t, 2t+1 of each
// n-tile.  The data are the B operand: lane 4g+t needs bytes 32s + 8t .. +7 of the row of permutation g of the n-tile -- one 8-byte
// load that lands in (b0, b1) as it is (the K index is permuted identically in the tables).
//
// Block GEMM: output rows j = 0..7 of block b (= the GEMM parts of t_8b..t_8b+7) for all 32 permutations, ks = 3 + 2b k-steps; the
// recombined values go to slots 8b..8b+7 of every row.
GL_D void ptc_block_gemm(unsigned char* wrows, const u64* __restrict__ tab, int ks, int b, int lane) {
    const int g = lane >> 2, t = lane & 3;
    u32 d[4][4][4];                         // [m-tile][n-tile][c]
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int c = 0; c < 4; c++) d[i][q][c] = 0;
    const unsigned char* b_base = wrows + g * PTC_ROW + 8 * t;
    const uint4* at = reinterpret_cast<const uint4*>(tab) + lane;
#pragma unroll 1
    for (int s = 0; s < ks; s++) {
        uint2 bq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) bq[q] = *reinterpret_cast<const uint2*>(b_base + 8 * q * PTC_ROW + 32 * s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 a = at[(s * 4 + i) * 32];
#pragma unroll
            for (int q = 0; q < 4; q++) ptc_mma(d[i][q], a.x, a.y, a.z, a.w, bq[q].x, bq[q].y);
        }
    }
    // c0, c1: limb 2i of output row g for permutations 8q + 2t, 8q + 2t + 1; c2, c3: limb 2i + 1
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const u64 v = ptc_recombine(d[0][q][e], d[0][q][2 + e], d[1][q][e], d[1][q][2 + e], d[2][q][e], d[2][q][2 + e], d[3][q][e], d[3][q][2 + e]);
            *reinterpret_cast<u64*>(wrows + (8 * q + 2 * t + e) * PTC_ROW + PTC_SLOT0 + 64 * b + 8 * g) = v;
        }
}

// Final GEMM for the n-tiles 2h, 2h+1 (permutations 16h..16h+15): both row groups (output lanes 0..7 and 8..15, of which 12 exist),
// 9 k-steps over the whole row; y(22) overwrites [0,96) of those permutations' own rows (no other n-tile reads them).
GL_D void ptc_final_gemm(unsigned char* wrows, const u64* __restrict__ tab, int h, int lane) {
    const int g = lane >> 2, t = lane & 3;
    u32 d[2][4][2][4];                      // [row group][m-tile][n-tile of the pair][c]
#pragma unroll
    for (int rg = 0; rg < 2; rg++)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int c = 0; c < 4; c++) d[rg][i][q][c] = 0;
    const unsigned char* b_base = wrows + (16 * h + g) * PTC_ROW + 8 * t;
    const uint4* at = reinterpret_cast<const uint4*>(tab) + lane;
#pragma unroll 1
    for (int s = 0; s < 9; s++) {
        const uint2 b0 = *reinterpret_cast<const uint2*>(b_base + 32 * s);
        const uint2 b1 = *reinterpret_cast<const uint2*>(b_base + 8 * PTC_ROW + 32 * s);
#pragma unroll
        for (int rg = 0; rg < 2; rg++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 a = at[((rg * 9 + s) * 4 + i) * 32];
                ptc_mma(d[rg][i][0], a.x, a.y, a.z, a.w, b0.x, b0.y);
                ptc_mma(d[rg][i][1], a.x, a.y, a.z, a.w, b1.x, b1.y);
            }
    }
    __syncwarp();        // every lane has read its B fragments before the rows are overwritten
#pragma unroll
    for (int rg = 0; rg < 2; rg++) {
        if (8 * rg + g < 12) {
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const u64 v = ptc_recombine(d[rg][0][q][e], d[rg][0][q][2 + e], d[rg][1][q][e], d[rg][1][q][2 + e], d[rg][2][q][e],
                                                d[rg][2][q][2 + e], d[rg][3][q][e], d[rg][3][q][2 + e]);
                    *reinterpret_cast<u64*>(wrows + (16 * h + 8 * q + 2 * t + e) * PTC_ROW + 8 * (8 * rg + g)) = v;
                }
        }
    }
}

#define PTC_OFF_BLOCK(b) ((b) == 0 ? 0 : (b) == 1 ? 768 : 2048)      // u64 words: 3, 5, 7 k-steps of 256 words
#define PTC_OFF_FINAL 3840

// Rounds 4..25 for the 32 states of a warp: x = S-box inputs of round 4 (Montgomery form, any representatives) -> S-box inputs of
// round 26.  wrows = the warp's 32 rows (bytes 272..287 of every row zeroed once by the caller), tab = the fragment tables in shared
// memory.  Must be called by all 32 lanes.
GL_D void poseidon_partial_tc(u64 x[12], unsigned char* wrows, const u64* __restrict__ tab, int lane) {
    unsigned char* my = wrows + lane * PTC_ROW;
    {
        ulonglong2* mp = reinterpret_cast<ulonglong2*>(my);
#pragma unroll
        for (int i = 0; i < 5; i++) mp[i] = make_ulonglong2(x[1 + 2 * i], x[2 + 2 * i]);
        mp[5] = make_ulonglong2(x[11], 1ULL);
    }
    u64 t = x[0];
    __syncwarp();
#pragma unroll 1
    for (int b = 0; b < 3; b++) {
        ptc_block_gemm(wrows, tab + PTC_OFF_BLOCK(b), 3 + 2 * b, b, lane);
        __syncwarp();
        u64* slot = reinterpret_cast<u64*>(my + PTC_SLOT0 + 64 * b);
        if (b > 0) t = slot[0];
        const int nj = (b == 2) ? 6 : 8;
#if PTC_HIST_SMEM
        // the s_k of the block are re-read from their slots (LSU pipe) instead of being kept in a register shift chain (14 moves per round)
#pragma unroll ptc_unroll_j
        for (int j = 0; j < 8; j++) {
            if (j >= nj) break;
            const u64 s = poseidon_sbox(t);
            const bool more = j + 1 < nj;
            PtcAcc A;
            if (more) {
                ptc_acc_init(A, slot[j + 1]);
                if (j >= 1) ptc_mac<2>(A, slot[j - 1]);
                if (j >= 2) ptc_mac<3>(A, slot[j - 2]);
                if (j >= 3) ptc_mac<4>(A, slot[j - 3]);
                if (j >= 4) ptc_mac<5>(A, slot[j - 4]);
                if (j >= 5) ptc_mac<6>(A, slot[j - 5]);
                if (j >= 6) ptc_mac<7>(A, slot[j - 6]);
            }
            slot[j] = s;
            if (more) {
                ptc_mac<1>(A, s);
                t = ptc_reduce(A);
            }
        }
#else
        u64 h[7];
#pragma unroll
        for (int i = 0; i < 7; i++) h[i] = 0;
#pragma unroll ptc_unroll_j
        for (int j = 0; j < 8; j++) {
            if (j >= nj) break;
            const u64 s = poseidon_sbox(t);
            const bool more = j + 1 < nj;
            PtcAcc A;
            if (more) {
                ptc_acc_init(A, slot[j + 1]);
                if (j >= 1) ptc_mac<2>(A, h[0]);
                if (j >= 2) ptc_mac<3>(A, h[1]);
                if (j >= 3) ptc_mac<4>(A, h[2]);
                if (j >= 4) ptc_mac<5>(A, h[3]);
                if (j >= 5) ptc_mac<6>(A, h[4]);
                if (j >= 6) ptc_mac<7>(A, h[5]);
            }
            slot[j] = s;
            if (more) {
                ptc_mac<1>(A, s);
                t = ptc_reduce(A);
            }
#pragma unroll
            for (int i = 6; i > 0; i--) h[i] = h[i - 1];
            h[0] = s;
        }
#endif
        __syncwarp();
    }
    ptc_final_gemm(wrows, tab + PTC_OFF_FINAL, 0, lane);
    ptc_final_gemm(wrows, tab + PTC_OFF_FINAL, 1, lane);
    __syncwarp();
    {
        const ulonglong2* mp = reinterpret_cast<const ulonglong2*>(my);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const ulonglong2 v = mp[i];
            x[2 * i] = v.x;
            x[2 * i + 1] = v.y;
        }
    }
}

// The permutation of a Montgomery-form state with the partial rounds in tensor-core form: rounds 0..3 and 26..29 run the limb-plane
// full-round body of poseidon.cuh (one copy of the code, executed for both stretches).  Warp-collective.
GL_D void poseidon_permute_mont_tc(u64 x[12], unsigned char* wrows, const u64* __restrict__ tab, int lane) {
    u32 ya[12], yb[12], yc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) poseidon_split(gl_addc(x[i], POSEIDON_RC0[i]), ya[i], yb[i], yc[i]);
#pragma unroll 1
    for (int ph = 0; ph < 2; ph++) {
        const int r_end = ph ? 30 : 4;
#pragma unroll 1
        for (int r = ph ? 26 : 0; r < r_end; r++) {
            u32 a[12], b[12], c[12];
#pragma unroll
            for (int i = 0; i < 12; i++) poseidon_split(poseidon_sbox(poseidon_join(ya[i], yb[i], yc[i])), a[i], b[i], c[i]);
            const u32* __restrict__ rc = POSEIDON_RC_LIMBS + r * 36;
            poseidon_mds_limb(ya, a, rc);
            poseidon_mds_limb(yb, b, rc + 1);
            poseidon_mds_limb(yc, c, rc + 2);
        }
        if (ph == 0) {
#pragma unroll
            for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
            poseidon_partial_tc(x, wrows, tab, lane);
#pragma unroll
            for (int i = 0; i < 12; i++) poseidon_split(x[i], ya[i], yb[i], yc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
}

// CTA-level setup: copy the fragment tables into shared memory and zero the K padding of every row.  smem = [tables | rows of warp 0 | ...].
GL_D void poseidon_tc_setup(unsigned char* smem, int nthreads, int tid) {
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(smem);
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(POSEIDON_TC_FRAGS);
    for (int i = tid; i < POSEIDON_TC_WORDS / 2; i += nthreads) dst[i] = src[i];
    unsigned char* my = smem + PTC_TABLE_BYTES + tid * PTC_ROW;
    *reinterpret_cast<ulonglong2*>(my + 272) = make_ulonglong2(0, 0);
    __syncthreads();
}
#define POSEIDON_TC_SMEM(nthreads) (PTC_TABLE_BYTES + (nthreads) * PTC_ROW)
