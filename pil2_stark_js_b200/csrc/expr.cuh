// Constraint-expression evaluation over a domain for sm_100a.
//
// Replaces (semantics, not structure) the reference's per-row expression evaluators:
//   calculateExps / compileCode   src/prover/prover_helpers.js:33-110   a straight-line program of {op, dest, src} records (add / sub /
//                                 mul / copy on base-field or F3 operands) run at every row of the domain "n" or "ext"; the reference
//                                 turns the program into the body of a JavaScript function and calls it 2^nBitsExt times
//   getRef / setRef / evalMap     src/prover/prover_helpers.js:112-265  operand -> buffer access: row (i + next) mod N for `prime`,
//                                 stage buffer + stagePos for `cm`, the Zi_ext / x_ext / xDivXSubXi_ext tables, q_ext / f_ext as sinks
//   F.add / F.sub / F.mul         src/helpers/f3g.js:47-104             mixed base-field / extension operands
// Here the program is compiled on the host into fixed-size records (EXPR_OP_WORDS u32 each) and interpreted by one thread per row:
// every thread of a warp executes the same record, so there is no divergence; operands are (kind, dim, index) triples resolved
// against a table of device buffers, a table of uniform constants (numbers, publics, challenges, evals, subproof values) and a
// per-thread file of temporaries whose slots the host assigns by liveness (the reference's tmp ids are single-assignment: the
// 169-record quotient program of the sm_all AIR keeps 40 alive at once, not 169).  All values are canonical Goldilocks words; an F3 value is 3 words
// and a base-field value is its embedding (a, 0, 0), which makes the mixed-dimension rules of f3g.js fall out of the componentwise
// ones (only the product distinguishes a scaling from a full F3 multiplication).
#pragma once
#include "gl.cuh"
#include "ntt.cuh"

#define EXPR_OP_WORDS 16          // u32 words per record: [opcode, n_src, 0, 0, dest(3), src0(3), src1(3), src2(3)]
#define EXPR_MAX_SLOTS 64         // temporaries alive at once (3 words each, thread-local)
#define EXPR_MAX_BUFS 24
#define EXPR_THREADS 128

enum { EXPR_ADD = 0, EXPR_SUB = 1, EXPR_MUL = 2, EXPR_COPY = 3, EXPR_MULADD = 4 };
// operand word 0: kind | dim << 8 | buffer index << 16;  word 1: tmp slot / constant index / column offset;  word 2: row offset
enum { EXPR_K_TMP = 0, EXPR_K_CONST = 1, EXPR_K_BUF = 2, EXPR_K_X = 3 };

struct ExprBufs {
    u64* ptr[EXPR_MAX_BUFS];
    u64 row_words[EXPR_MAX_BUFS];
};

struct ExprVal {
    u64 c[3];
    int dim;
};

GL_D ExprVal expr_load(const u32* __restrict__ o, const u64* __restrict__ tmp, const u64* __restrict__ consts, const ExprBufs& bufs, u64 i, u64 mask,
                       int dom_bits, int x_shift, const NttTables& tb) {
    const u32 w0 = o[0];
    const int kind = (int)(w0 & 255), dim = (int)((w0 >> 8) & 255);
    ExprVal v;
    v.dim = dim;
    v.c[1] = v.c[2] = 0;
    if (kind == EXPR_K_TMP) {
        const u64* t = tmp + 3 * o[1];
        v.c[0] = t[0]; v.c[1] = t[1]; v.c[2] = t[2];
    } else if (kind == EXPR_K_CONST) {
        const u64* t = consts + 3 * (u64)o[1];
        v.c[0] = t[0]; v.c[1] = t[1]; v.c[2] = t[2];
    } else if (kind == EXPR_K_BUF) {
        const int b = (int)(w0 >> 16);
        const u64 row = (i + o[2]) & mask;
        const u64* p = bufs.ptr[b] + row * bufs.row_words[b] + o[1];
        v.c[0] = gl_canon(p[0]);
        if (dim == 3) { v.c[1] = gl_canon(p[1]); v.c[2] = gl_canon(p[2]); }
    } else {   // EXPR_K_X: x_i = w^i on the domain, times the coset shift 7 on the extended one (stark_gen_helpers.js:110-115,136-143)
        const u32 E = dom_bits == 0 ? 0u : ((u32)i << (32 - dom_bits));
        u64 x = gl_from_mont(ntt_root_pow(tb.bytepow, E));
        if (x_shift) x = gl_canon(gl_mul(x, GL_SHIFT));
        v.c[0] = x;
    }
    return v;
}

__global__ void __launch_bounds__(EXPR_THREADS) expr_kernel(const u32* __restrict__ ops, u32 n_ops, const u64* __restrict__ consts, const __grid_constant__ ExprBufs bufs,
                                                            int dom_bits, int x_shift, NttTables tb) {
    const u64 N = 1ULL << dom_bits, mask = N - 1;
    const u64 i = (u64)blockIdx.x * EXPR_THREADS + threadIdx.x;
    if (i >= N) return;
    u64 tmp[3 * EXPR_MAX_SLOTS];
    for (u32 k = 0; k < n_ops; k++) {
        const u32* __restrict__ o = ops + (size_t)k * EXPR_OP_WORDS;
        const int opc = (int)o[0];
        ExprVal a = expr_load(o + 7, tmp, consts, bufs, i, mask, dom_bits, x_shift, tb), r;
        if (opc == EXPR_COPY) {
            r = a;
        } else {
            const ExprVal b = expr_load(o + 10, tmp, consts, bufs, i, mask, dom_bits, x_shift, tb);
            r.dim = a.dim > b.dim ? a.dim : b.dim;
            if (opc == EXPR_ADD) {
#pragma unroll
                for (int c = 0; c < 3; c++) r.c[c] = gl_canon(gl_add(a.c[c], b.c[c]));
            } else if (opc == EXPR_SUB) {
#pragma unroll
                for (int c = 0; c < 3; c++) r.c[c] = gl_canon(gl_sub(a.c[c], b.c[c]));
            } else {   // MUL / MULADD
                if (a.dim == 3 && b.dim == 3) {
                    const gl3 p = gl3_canon(gl3_mul(gl3{{a.c[0], a.c[1], a.c[2]}}, gl3{{b.c[0], b.c[1], b.c[2]}}));
                    r.c[0] = p.c[0]; r.c[1] = p.c[1]; r.c[2] = p.c[2];
                } else if (a.dim == 1) {
#pragma unroll
                    for (int c = 0; c < 3; c++) r.c[c] = gl_canon(gl_mul(a.c[0], b.c[c]));
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) r.c[c] = gl_canon(gl_mul(a.c[c], b.c[0]));
                }
                if (opc == EXPR_MULADD) {
                    const ExprVal d = expr_load(o + 13, tmp, consts, bufs, i, mask, dom_bits, x_shift, tb);
                    if (d.dim > r.dim) r.dim = d.dim;
#pragma unroll
                    for (int c = 0; c < 3; c++) r.c[c] = gl_canon(gl_add(r.c[c], d.c[c]));
                }
            }
        }
        // destination
        const u32 d0 = o[4];
        const int dkind = (int)(d0 & 255), ddim = (int)((d0 >> 8) & 255);
        if (dkind == EXPR_K_TMP) {
            u64* t = tmp + 3 * o[5];
            t[0] = r.c[0]; t[1] = r.c[1]; t[2] = r.c[2];
        } else {   // EXPR_K_BUF: q_ext / f_ext / a committed polynomial
            const int b = (int)(d0 >> 16);
            const u64 row = (i + o[6]) & mask;
            u64* p = bufs.ptr[b] + row * bufs.row_words[b] + o[5];
            p[0] = r.c[0];
            if (ddim == 3) { p[1] = r.c[1]; p[2] = r.c[2]; }
        }
    }
}
