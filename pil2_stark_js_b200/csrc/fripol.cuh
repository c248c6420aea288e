// FRI polynomial f_ext (computeFRIStark, src/stark/stark_gen_helpers.js:325-334) for sm_100a.
//
// The reference evaluates the generated expression friExp row by row (friPolinomial.js:26-56):
//     F_o(x)  = Horner in vf2 over the evMap entries opened at o:  (...((p_1 - ev_1) vf2 + (p_2 - ev_2)) vf2 + ...)
//     f(x)    = Horner in vf1 over the openings:                   (...(F_a xDiv_a) vf1 + F_b xDiv_b) ...
// i.e.  f(x_k) = sum_g u_g xDiv_g(x_k) (S_g(k) - c_g),   S_g(k) = sum_i w_i p_i(x_k),   c_g = sum_i w_i ev_i,
// with w_i, u_g powers of the two challenges.  S is a dense contraction over the columns of the extended buffers --
// (E x C) . (C x 3 n_groups) over F_p -- and runs on the tensor cores as a byte-limb GEMM like the evaluation sums (evals.cuh):
// the A operand is the row as it lies in memory (u8 limbs of consecutive columns = the contraction index, no transposition),
// the B operand holds the byte limbs of  W'[(col, b)][(g, c)] = W[col][(g, c)] * 2^(8b) mod p,  pre-arranged on the host in
// fragment order.  K = 8 * columns <= 2^15 keeps the s32 sums exact (2^15 * 255^2 < 2^31); the 8 limb sums of an output are
// recombined with one quad reduction.  fripol_finish_kernel applies xDiv, the constants and the vf1 powers.
#pragma once
#include "evals.cuh"

#define FP_WARPS 4               // 32 rows per warp (two m16 tiles), 128 rows per CTA

// D(16x8) += A(16x32, u8) * B(32x8, u8): same instruction as evm_mma
// BF: fragment-ordered B, u32 pairs [(dstep * 2 + h) * NT + nt][lane][2]
// S[oc * rows + row] (+)= field value of sum_{col, b} V-byte * W'
template <int NT>
__global__ void __launch_bounds__(FP_WARPS * 32) fripol_mma_kernel(const u64* __restrict__ buf, u64 size, u64 rows, const uint2* __restrict__ BF,
                                                                    u32 dsteps, u64* __restrict__ S, int accumulate) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const u64 row0 = ((u64)blockIdx.x * FP_WARPS + warp) * 32;
    if (row0 >= rows) return;
    int acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
        for (int t = 0; t < NT; t++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[m][t][j] = 0;
    const u64 row_bytes = size * 8;
    const unsigned char* rp[4];          // rows gid, gid + 8 of the two m-tiles
#pragma unroll
    for (int q = 0; q < 4; q++) {
        u64 r = row0 + (q >> 1) * 16 + (q & 1) * 8 + gid;
        if (r >= rows) r = rows - 1;     // clamped rows are computed and discarded
        rp[q] = reinterpret_cast<const unsigned char*>(buf + r * size);
    }
    for (u32 j = 0; j < dsteps; j++) {
        const u64 off = (u64)j * 64 + 16 * tig;
        u32 w[4][4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            ulonglong2 v = make_ulonglong2(0, 0);
            if (off < row_bytes) v.x = *reinterpret_cast<const u64*>(rp[q] + off);
            if (off + 8 < row_bytes) v.y = *reinterpret_cast<const u64*>(rp[q] + off + 8);
            w[q][0] = (u32)v.x; w[q][1] = (u32)(v.x >> 32); w[q][2] = (u32)v.y; w[q][3] = (u32)(v.y >> 32);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 a0[4] = {w[0][2 * h], w[1][2 * h], w[0][2 * h + 1], w[1][2 * h + 1]};      // m-tile 0: rows gid, gid+8; k, k+16
            const u32 a1[4] = {w[2][2 * h], w[3][2 * h], w[2][2 * h + 1], w[3][2 * h + 1]};      // m-tile 1
            const uint2* __restrict__ bf = BF + ((size_t)(2 * j + h) * NT) * 32 + lane;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                const uint2 b = bf[t * 32];
                evm_mma(acc[0][t], a0, b.x, b.y);
                evm_mma(acc[1][t], a1, b.x, b.y);
            }
        }
    }
    // limb recombination: lane holds columns b' = 2 tig, 2 tig + 1 of rows gid (regs 0,1) and gid + 8 (regs 2,3)
    const u64 pw = (u64)1 << (16 * tig);     // 2^(16 tig) < p: canonical
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
        for (int t = 0; t < NT; t++)
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const u64 v = (u64)(u32)acc[m][t][2 * half] + ((u64)(u32)acc[m][t][2 * half + 1] << 8);
                u64 f = gl_mul(v, pw);
                f = gl_add(f, __shfl_xor_sync(0xFFFFFFFFu, f, 1));
                f = gl_add(f, __shfl_xor_sync(0xFFFFFFFFu, f, 2));
                const u64 r = row0 + m * 16 + half * 8 + gid;
                if (tig == 0 && r < rows) {
                    u64* o = S + (u64)t * rows + r;
                    *o = gl_canon(accumulate ? gl_add(*o, f) : f);
                }
            }
}

// ---- second formulation (r02): the rows are the B operand ------------------------------------------------------------------------
// N = rows (n-tile = 8 rows), K = (column, limb b) -- the B fragment of a lane is one buffer element as it lies in memory -- and the
// constants W'[(col, b)][oc] are the A operand with M = (limb pair, oc): m-tile i holds limbs 2i (tile rows 0..7 = the up to 8 outputs oc)
// and 2i+1, so a lane ends up with all 8 limb sums of (oc g; rows 2t, 2t+1 of every n-tile) and recombines them alone (no shuffles, no
// fragment register moves: the first formulation's A fragments interleave two rows and cost 8 moves per two MMAs).  A CTA is 8 warps x 32
// rows.  Every warp streams its rows through a private cp.async ring in steps of 8 columns (64 bytes per row, 2 KB per step, 2 steps in
// flight); the A fragments of a step (4 KB) are shared by the 8 warps through a CTA-wide ring filled one 16-byte copy per thread and step,
// made visible by a barrier every second step.  n_groups <= 2 (oc <= 8); more groups take fripol_mma_kernel.
#define FP2_THREADS 256
#define FP2_DPITCH 96                         // bytes per staged row of a step: 64 + 32, rows land 8 banks apart (conflict-free fragment loads)
#define FP2_DSTAGE (32 * FP2_DPITCH)
#define FP2_DRING 3
#define FP2_TPIECE 4096                       // A fragments of one step: 2 k-steps x 4 m-tiles x 32 lanes x 16 bytes
#define FP2_TRING 6
#define FP2_SMEM (FP2_TRING * FP2_TPIECE + 8 * FP2_DRING * FP2_DSTAGE)

// AF: [piece][k-step of the piece][m-tile][lane] uint4, zero padded to npieces + 3 pieces.  S[oc * rows + row] (+)= sum.
__global__ void __launch_bounds__(FP2_THREADS, 2) fripol_mma2_kernel(const u64* __restrict__ buf, u64 size, u64 rows, const uint4* __restrict__ AF,
                                                                     u32 npieces, int NT, u64* __restrict__ S, int accumulate) {
    extern __shared__ __align__(16) unsigned char fp2_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    unsigned char* tring = fp2_smem;
    unsigned char* dring = fp2_smem + FP2_TRING * FP2_TPIECE + warp * (FP2_DRING * FP2_DSTAGE);
    const u64 row0 = ((u64)blockIdx.x * 8 + warp) * 32;
    int acc[4][4][4];                         // [m-tile][n-tile][c]
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][j][c] = 0;
    // copies of this lane per step: chunks lane + 32 q (q < 4): row (lane >> 2) + 8 q, 16-byte part lane & 3 (columns 8 f + 2 part, + 1)
    const bool ok16 = ((size & 1) == 0) && ((reinterpret_cast<size_t>(buf) & 15) == 0);
    const int part = lane & 3;
    const u64* src[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        u64 r = row0 + (lane >> 2) + 8 * q;
        if (r >= rows) r = rows - 1;          // clamped rows are computed and discarded
        src[q] = buf + r * size + 2 * part;
    }
    const unsigned dsa = (unsigned)__cvta_generic_to_shared(dring) + (lane >> 2) * FP2_DPITCH + part * 16;
    const unsigned tsa = (unsigned)__cvta_generic_to_shared(tring) + threadIdx.x * 16;
    u32 nf = 0;                               // next step to issue
    auto issue = [&]() {                      // group nf: data of step nf, A fragments of step nf + 1 (step 0's ride along with group 0)
        if (nf < npieces) {
            const u64 c0 = (u64)8 * nf + 2 * part;
            const unsigned da = dsa + (nf % FP2_DRING) * FP2_DSTAGE;
            if (ok16) {
                const bool v = c0 < size;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da + q * 8 * FP2_DPITCH), "l"(v ? src[q] + 8 * nf : buf), "r"(v ? 16u : 0u) : "memory");
            } else {
                const bool v0 = c0 < size, v1 = c0 + 1 < size;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(da + q * 8 * FP2_DPITCH), "l"(v0 ? src[q] + 8 * nf : buf), "r"(v0 ? 8u : 0u) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(da + q * 8 * FP2_DPITCH + 8), "l"(v1 ? src[q] + 8 * nf + 1 : buf), "r"(v1 ? 8u : 0u) : "memory");
                }
            }
            if (nf == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tsa), "l"(AF + threadIdx.x) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tsa + ((nf + 1) % FP2_TRING) * FP2_TPIECE), "l"(AF + (size_t)(nf + 1) * 256 + threadIdx.x) : "memory");
        }
        nf++;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue();
    issue();
    for (u32 f = 0; f < npieces; f++) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // groups <= f have landed: data of step f, A fragments of steps <= f + 1
        if ((f & 1) == 0) __syncthreads(); else __syncwarp();     // even steps publish the A fragments of steps f, f + 1 to the whole CTA
        issue();                              // refills the slot of step f - 1: every lane is past that step (the sync above)
        const unsigned char* db = dring + (f % FP2_DRING) * FP2_DSTAGE + g * FP2_DPITCH + t * 8;
        const uint4* __restrict__ ab = reinterpret_cast<const uint4*>(tring + (f % FP2_TRING) * FP2_TPIECE) + lane;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            u64 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = *reinterpret_cast<const u64*>(db + 8 * j * FP2_DPITCH + 32 * h);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 a = ab[(h * 4 + i) * 32];
                const u32 af[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int j = 0; j < 4; j++) evm_mma(acc[i][j], af, (u32)v[j], (u32)(v[j] >> 32));
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (g < NT) {
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const u64 L = (u64)(u32)acc[0][j][e] + ((u64)(u32)acc[0][j][2 + e] << 8) + ((u64)(u32)acc[1][j][e] << 16) + ((u64)(u32)acc[1][j][2 + e] << 24);
                const u64 H = (u64)(u32)acc[2][j][e] + ((u64)(u32)acc[2][j][2 + e] << 8) + ((u64)(u32)acc[3][j][e] << 16) + ((u64)(u32)acc[3][j][2 + e] << 24);
                const u64 lo = L + (H << 32);
                const u64 f = gl_reduce128((H >> 32) + (u64)(lo < L), lo);
                const u64 r = row0 + 8 * j + 2 * t + e;
                if (r < rows) {
                    u64* o = S + (u64)g * rows + r;
                    *o = gl_canon(accumulate ? gl_add(*o, f) : f);
                }
            }
    }
}

struct FriPolFinish {
    u64 u[12][3];        // vf1 powers per group
    u64 c[12][3];        // sum_i w_i ev_i per group
    int xidx[12];        // column of the group's opening in xDivXSubXi_ext
    int n_groups, n_open;
};
// f[row] = sum_g u_g * xdiv[row][xidx_g] * (S[row][g] - c_g)
__global__ void fripol_finish_kernel(const u64* __restrict__ S, const u64* __restrict__ xdiv, const __grid_constant__ FriPolFinish P, u64 rows,
                                     u64* __restrict__ f) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    gl3 acc = {{0, 0, 0}};
    for (int g = 0; g < P.n_groups; g++) {
        const u64* s = S + (u64)(3 * g) * rows + r;          // S is planar: S[oc * rows + row], oc = 3 g + c (coalesced on both sides)
        const gl3 sg = {{gl_sub(s[0], P.c[g][0]), gl_sub(s[rows], P.c[g][1]), gl_sub(s[2 * rows], P.c[g][2])}};
        const u64* x = xdiv + 3 * (r * P.n_open + P.xidx[g]);
        const gl3 t = gl3_mul(gl3_mul(sg, gl3{{x[0], x[1], x[2]}}), gl3{{P.u[g][0], P.u[g][1], P.u[g][2]}});
        acc = gl3_add(acc, t);
    }
    acc = gl3_canon(acc);
    f[3 * r] = acc.c[0]; f[3 * r + 1] = acc.c[1]; f[3 * r + 2] = acc.c[2];
}

// ---- host-side F3 helpers (setup of the weights only) ----
struct h3 { u64 c[3]; };
static inline u64 glh_add(u64 a, u64 b) { return (u64)(((unsigned __int128)a + b) % GL_P); }
static inline u64 glh_sub(u64 a, u64 b) { return (u64)(((unsigned __int128)a + GL_P - b % GL_P) % GL_P); }
static inline h3 h3_add(h3 a, h3 b) { return h3{{glh_add(a.c[0], b.c[0]), glh_add(a.c[1], b.c[1]), glh_add(a.c[2], b.c[2])}}; }
static inline h3 h3_mul(h3 a, h3 b) {        // f3g.js:94-102 in schoolbook form: x^3 = x + 1, x^4 = x^2 + x
    const u64 a0b0 = glh_mul(a.c[0], b.c[0]), a0b1 = glh_mul(a.c[0], b.c[1]), a0b2 = glh_mul(a.c[0], b.c[2]);
    const u64 a1b0 = glh_mul(a.c[1], b.c[0]), a1b1 = glh_mul(a.c[1], b.c[1]), a1b2 = glh_mul(a.c[1], b.c[2]);
    const u64 a2b0 = glh_mul(a.c[2], b.c[0]), a2b1 = glh_mul(a.c[2], b.c[1]), a2b2 = glh_mul(a.c[2], b.c[2]);
    const u64 x3 = glh_add(a1b2, a2b1), x4 = a2b2;
    return h3{{glh_add(a0b0, x3), glh_add(glh_add(glh_add(a0b1, a1b0), x3), x4), glh_add(glh_add(glh_add(a0b2, a1b1), a2b0), x4)}};
}
static inline h3 h3_mulx(h3 r) { return h3{{r.c[2], glh_add(r.c[0], r.c[2]), r.c[1]}}; }      // x * (r0 + r1 x + r2 x^2)
