// FRI fold (+ transposed rows and leaf digests of the next layer's tree) for sm_100a.
//
// Replaces (semantics, not structure) src/stark/fri.js:22-81 and getTransposedBuffer :187-202:
//   step s > 0:  P'[g] = sum_i c_i * (challenge * sinv_g)^i,  c = INTT_nX(P[i*2^cur + g], i < nX),
//                sinv_g = (7^(2^(b0-prev)))^-1 * w_prev^-g,   nX = 2^(prev-cur)
//   rows for the next tree: row i (< 2^next) = [P'[i + 2^next*j] (3 words) for j < 2^(cur-next)]
// INTT-then-Horner is evaluated as log2(nX) exact radix-2 folds of the interpolant,
//   e'_j = (e_j + e_{j+n/2}) + beta * (e_j - e_{j+n/2}) * w_n^-j,   beta -> beta^2,   final scale 1/nX,
// which is the same field element (exact arithmetic) with ~1/3 fewer multiplications and half the registers.
// One thread per output point g; consecutive threads own consecutive g (coalesced 24-byte records).
#pragma once
#include "merkle.cuh"
#include "ntt.cuh"

#define FRI_MAX_FOLD_BITS 6     // nX <= 64 (the reference's configs use 2^4 and 2^5)
#define FRI_ROWS_PER_CTA 32

struct FriParams {
    int prev_bits, cur_bits, next_bits;   // next_bits < 0: last step (no rows / tree)
    int write_rows;                        // rows + leaf digests wanted
    int fuse_leaf_hash;                    // standard linear hash computed in-kernel
    int in_rows;                           // input layout: 0 = polynomial order pol[3*(j*2^cur + g)], 1 = the transposed rows of the
                                           // previous layer rows[3*(g*nX + j)] (its next_bits == this cur_bits): 16 contiguous inputs per output
    int write_pol;                         // store pol2[g] (0: rows only -- the sharded chain continues from the rows)
    u64 row0, n_rows;                      // rows [row0, row0 + n_rows) of the next layer (n_rows == 0: all) -- sharded chains
    u64 shift_inv;                         // (7^(2^(b0-prev)))^-1
    u64 nx_inv;                            // (2^(prev-cur))^-1
    u64 challenge[3];
};

GL_D gl3 fri_load3(const u64* __restrict__ p) { return gl3{{p[0], p[1], p[2]}}; }

// Evaluate the interpolant of e_0..e_{nX-1} (on <w_nX>) at beta; FOLD = log2(nX).
template <int FOLD>
GL_D gl3 fri_fold_point(const u64* __restrict__ pol, u64 g, int cur_bits, gl3 beta, const u64* __restrict__ tw_inv, int in_rows) {
    constexpr int NX = 1 << FOLD;
    // input j of output g: polynomial order base + j * 3 * 2^cur, rows order base + 3 j
    const u64* __restrict__ base = in_rows ? pol + 3 * g * NX : pol + 3 * g;
    const u64 sj = in_rows ? 3 : ((u64)3 << cur_bits);
    if constexpr (FOLD == 0) {
        return fri_load3(base);
    } else {
    gl3 e[NX / 2];
    // first fold straight from memory: pairs (j, j + NX/2)
#pragma unroll
    for (int j = 0; j < NX / 2; j++) {
        const gl3 a = fri_load3(base + (u64)j * sj);
        const gl3 b = fri_load3(base + (u64)(j + NX / 2) * sj);
        const u64 wj = tw_inv[(NX / 2) + j];                         // w_NX^-j (Montgomery form)
        e[j] = gl3_add(gl3_add(a, b), gl3_mul(beta, gl3_mscale(gl3_sub(a, b), wj)));
    }
#pragma unroll
    for (int lvl = 1; lvl < FOLD; lvl++) {
        beta = gl3_mul(beta, beta);
        const int n = NX >> lvl;   // current size
#pragma unroll
        for (int j = 0; j < NX / 2; j++) {
            if (j < n / 2) {
                const gl3 a = e[j], b = e[j + n / 2];
                const u64 wj = tw_inv[(n / 2) + j];                         // w_n^-j (Montgomery form)
                e[j] = gl3_add(gl3_add(a, b), gl3_mul(beta, gl3_mscale(gl3_sub(a, b), wj)));
            }
        }
    }
    return e[0];
    }
}

template <int FOLD>
__global__ void __launch_bounds__(FOLD >= 5 ? 256 : 512) fri_fold_kernel(const u64* pol, u64* pol2, u64* __restrict__ rows,   // pol2 == pol for the in-place identity
                                                        u64* __restrict__ nodes, FriParams P, NttTables tb) {                // step: no __restrict__ on those two
    extern __shared__ u64 fri_rows[];   // [rows_per_cta][3*gs] when fusing the leaf hash
    const int next_bits = P.next_bits < 0 ? P.cur_bits : P.next_bits;
    const u64 gs = 1ULL << (P.cur_bits - next_bits);   // group size (elements per row)
    const u64 n_rows = 1ULL << next_bits;
    const int rb = (int)(n_rows < FRI_ROWS_PER_CTA ? n_rows : FRI_ROWS_PER_CTA);
    const int ii = threadIdx.x % rb;
    const int jj = threadIdx.x / rb;
    const int jstep = blockDim.x / rb;
    const u64 i = P.row0 + (u64)blockIdx.x * rb + ii;
    const gl3 alpha = gl3{{P.challenge[0], P.challenge[1], P.challenge[2]}};
    for (u64 j = jj; j < gs; j += jstep) {
        const u64 g = i + (j << next_bits);
        // sinv_g = shift_inv * w_prev^-g
        const u32 E = P.prev_bits == 0 ? 0u : (0u - ((u32)g << (32 - P.prev_bits)));
        const u64 sinv = gl_mmul(P.shift_inv, ntt_root_pow(tb.bytepow, E));   // root in Montgomery form: plain product
        gl3 v = fri_fold_point<FOLD>(pol, g, P.cur_bits, gl3_scale(alpha, sinv), tb.tw_inv, P.in_rows);
        v = gl3_canon(gl3_scale(v, P.nx_inv));
        if (P.write_pol) { pol2[3 * g] = v.c[0]; pol2[3 * g + 1] = v.c[1]; pol2[3 * g + 2] = v.c[2]; }
        if (P.write_rows) {
            u64* r = rows + (i * gs + j) * 3;
            r[0] = v.c[0]; r[1] = v.c[1]; r[2] = v.c[2];
            if (P.fuse_leaf_hash) {
                u64* s = fri_rows + ((size_t)ii * gs + j) * 3;
                s[0] = v.c[0]; s[1] = v.c[1]; s[2] = v.c[2];
            }
        }
    }
    if (P.write_rows && P.fuse_leaf_hash) {
        __syncthreads();
        if (threadIdx.x < rb) {
            u64 d[4];
            merkle_sponge(fri_rows + (size_t)threadIdx.x * gs * 3, 3 * gs, d);
            const u64 row = P.row0 + (u64)blockIdx.x * rb + threadIdx.x;
#pragma unroll
            for (int k = 0; k < 4; k++) nodes[4 * row + k] = d[k];
        }
    }
}

// Launch one fold step.  Returns launches or -1 (unsupported fold width).
static int fri_launch_fold(const u64* pol, u64* pol2, u64* rows, u64* nodes, const FriParams& P, const NttTables& tb, cudaStream_t st) {
    const int fold = P.prev_bits - P.cur_bits;
    if (fold < 0 || fold > FRI_MAX_FOLD_BITS) return -1;
    const int next_bits = P.next_bits < 0 ? P.cur_bits : P.next_bits;
    const u64 gs = 1ULL << (P.cur_bits - next_bits);
    const u64 all_rows = 1ULL << next_bits;
    const u64 n_rows = P.n_rows ? P.n_rows : all_rows;
    const int rb = (int)(all_rows < FRI_ROWS_PER_CTA ? all_rows : FRI_ROWS_PER_CTA);   // the kernel derives rb from next_bits the same way
    if (n_rows % rb) return -1;
    u64 jb = gs;
    // registers: the fold keeps NX/2 F3 values live; keep CTAs at <= 256 threads for the wide folds
    const u64 max_threads = fold >= 5 ? 256 : 512;
    while ((u64)rb * jb > max_threads) jb >>= 1;
    if (jb == 0) jb = 1;
    const unsigned threads = (unsigned)(rb * jb);
    const unsigned blocks = (unsigned)(n_rows / rb);
    const size_t smem = (P.write_rows && P.fuse_leaf_hash) ? (size_t)rb * gs * 3 * sizeof(u64) : 0;
    if (smem > 200 * 1024) return -1;
#define FRI_CASE(F)                                                                                                          \
    case F:                                                                                                                  \
        if (smem > 48 * 1024 && cudaFuncSetAttribute(fri_fold_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1; \
        fri_fold_kernel<F><<<blocks, threads, smem, st>>>(pol, pol2, rows, nodes, P, tb);                                        \
        break;
    switch (fold) {
        FRI_CASE(0) FRI_CASE(1) FRI_CASE(2) FRI_CASE(3) FRI_CASE(4) FRI_CASE(5) FRI_CASE(6)
        default: return -1;
    }
#undef FRI_CASE
    return 1;
}
