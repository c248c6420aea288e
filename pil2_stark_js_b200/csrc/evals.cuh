// Evaluations at the challenge point and the FRI quotient denominators for sm_100a.
//
// Replaces (semantics, not structure) the two serial BigInt loops that follow the commits in the reference prover:
//   computeEvalsStark   src/stark/stark_gen_helpers.js:210-273   LEv[i] = ifft(powers of xi*w^opening/7) (:216-231), then
//                                                                evals[e] = sum_k pol_e(7 w_N^k) * LEv[k]   (:234-267)
//   xDivXSubXi_ext      src/stark/stark_gen_helpers.js:289-323   x_k / (x_k - xi*w^opening), x_k = 7 w_E^k, via batchInverse
// All values are exact field elements, so any evaluation order gives the reference's words; outputs are canonical.
#pragma once
#include "gl.cuh"
#include "ntt.cuh"

// ---- powers of an F3 element: out[k] = x^k, k < n (row-major n x 3).  Each thread starts from x^(k0) built by
// square-and-multiply over a shared table of x^(2^j) and walks EV_POW_CHUNK consecutive powers.
#define EV_POW_CHUNK 16
__global__ void __launch_bounds__(256) f3_powers_kernel(const u64 x0, const u64 x1, const u64 x2, u64 n, u64* __restrict__ out) {
    __shared__ u64 sq[32][3];
    if (threadIdx.x == 0) {
        gl3 s = {{x0, x1, x2}};
        for (int j = 0; j < 32; j++) {
            sq[j][0] = s.c[0]; sq[j][1] = s.c[1]; sq[j][2] = s.c[2];
            s = gl3_mul(s, s);
        }
    }
    __syncthreads();
    const u64 k0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * EV_POW_CHUNK;
    if (k0 >= n) return;
    gl3 cur = {{1, 0, 0}};
    for (int j = 0; j < 32; j++)
        if ((k0 >> j) & 1) cur = gl3_mul(cur, gl3{{sq[j][0], sq[j][1], sq[j][2]}});
    const gl3 x = {{x0, x1, x2}};
    for (u64 k = k0; k < k0 + EV_POW_CHUNK && k < n; k++) {
        const gl3 c = gl3_canon(cur);
        out[3 * k] = c.c[0]; out[3 * k + 1] = c.c[1]; out[3 * k + 2] = c.c[2];
        cur = gl3_mul(cur, x);
    }
}

// ---- evaluations -------------------------------------------------------------------------------------------------
struct EvalDesc {      // one entry of pilInfo.evMap restricted to one buffer (stark_gen_helpers.js:236-249)
    u64 offset;        // column of the polynomial inside a row of the buffer
    u32 dim;           // 1 (base field) or 3 (three consecutive columns = one F3 value)
    u32 lev;           // which LEv vector (index into openingPoints)
};
#define EV_THREADS 256

// partial[(chunk * n_evals + e) * 3 ..] = sum over the chunk's rows k of buf[(k << eb) * size + offset] * lev[k]
// blockDim.x = EV_THREADS = ew * rw: thread (r, e) takes rows r, r + rw, ... of the chunk for evaluation e (+ ew, ...).
__global__ void __launch_bounds__(EV_THREADS) evals_partial_kernel(const u64* __restrict__ buf, u64 size, int eb, u64 n, const EvalDesc* __restrict__ desc,
                                                                   u32 n_evals, const u64* __restrict__ lev, u32 ew, u64* __restrict__ partial) {
    extern __shared__ u64 ev_sm[];   // [rw][ew][3]
    const u32 rw = EV_THREADS / ew;
    const u32 e_lane = threadIdx.x % ew, r = threadIdx.x / ew;
    const u64 rows_per_chunk = (n + gridDim.x - 1) / gridDim.x;
    const u64 kb = (u64)blockIdx.x * rows_per_chunk;
    const u64 ke = kb + rows_per_chunk < n ? kb + rows_per_chunk : n;
    for (u32 e0 = 0; e0 < n_evals; e0 += ew) {
        const u32 e = e0 + e_lane;
        gl3 acc = {{0, 0, 0}};
        if (e < n_evals) {
            const EvalDesc d = desc[e];
            const u64* __restrict__ lv = lev + (u64)d.lev * n * 3;
            for (u64 k = kb + r; k < ke; k += rw) {
                const u64* __restrict__ v = buf + (k << eb) * size + d.offset;
                const gl3 l = {{lv[3 * k], lv[3 * k + 1], lv[3 * k + 2]}};
                gl3 t;
                if (d.dim == 1) t = gl3_scale(l, v[0]);
                else t = gl3_mul(gl3{{v[0], v[1], v[2]}}, l);
                acc = gl3_add(acc, t);
            }
        }
        ev_sm[(r * ew + e_lane) * 3 + 0] = acc.c[0];
        ev_sm[(r * ew + e_lane) * 3 + 1] = acc.c[1];
        ev_sm[(r * ew + e_lane) * 3 + 2] = acc.c[2];
        __syncthreads();
        if (r == 0 && e < n_evals) {
            for (u32 rr = 1; rr < rw; rr++)
                acc = gl3_add(acc, gl3{{ev_sm[(rr * ew + e_lane) * 3], ev_sm[(rr * ew + e_lane) * 3 + 1], ev_sm[(rr * ew + e_lane) * 3 + 2]}});
            u64* o = partial + ((u64)blockIdx.x * n_evals + e) * 3;
            o[0] = acc.c[0]; o[1] = acc.c[1]; o[2] = acc.c[2];
        }
        __syncthreads();
    }
}
// out[e] = canonical sum over chunks of partial[chunk][e]
__global__ void evals_reduce_kernel(const u64* __restrict__ partial, u32 chunks, u32 n_evals, u64* __restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;   // (e, c)
    if (i >= n_evals * 3) return;
    u64 acc = 0;
    for (u32 ch = 0; ch < chunks; ch++) acc = gl_add(acc, partial[(u64)ch * n_evals * 3 + i]);
    out[i] = gl_canon(acc);
}

// ---- x / (x - xi) over the extended domain ----------------------------------------------------------------------------
// F3 inverse (f3g.js:136-172) split around the single base-field inversion so that a thread can share one inversion
// between several points (Montgomery's trick -- the same identity as the reference's batchInverse, f3g.js:370-385).
struct F3InvParts { u64 i1, i2, i3, t; };   // a^-1 = (i1, i2, i3) / t
GL_D F3InvParts gl3_inv_parts(const gl3& a) {
    const u64 aa = gl_mul(a.c[0], a.c[0]), ac = gl_mul(a.c[0], a.c[2]), ba = gl_mul(a.c[1], a.c[0]), bb = gl_mul(a.c[1], a.c[1]);
    const u64 bc = gl_mul(a.c[1], a.c[2]), cc = gl_mul(a.c[2], a.c[2]);
    const u64 aaa = gl_mul(aa, a.c[0]), aac = gl_mul(aa, a.c[2]), abc = gl_mul(ba, a.c[2]), abb = gl_mul(ba, a.c[1]), acc = gl_mul(ac, a.c[2]);
    const u64 bbb = gl_mul(bb, a.c[1]), bcc = gl_mul(bc, a.c[2]), ccc = gl_mul(cc, a.c[2]);
    u64 t = gl_neg(aaa);
    t = gl_sub(t, aac); t = gl_sub(t, aac); t = gl_add(t, abc); t = gl_add(t, abc); t = gl_add(t, abc); t = gl_add(t, abb);
    t = gl_sub(t, acc); t = gl_sub(t, bbb); t = gl_add(t, bcc); t = gl_sub(t, ccc);
    F3InvParts r;
    u64 i1 = gl_neg(aa);
    i1 = gl_sub(i1, ac); i1 = gl_sub(i1, ac); i1 = gl_add(i1, bc); i1 = gl_add(i1, bb); i1 = gl_sub(i1, cc);
    r.i1 = i1;
    r.i2 = gl_sub(ba, cc);
    r.i3 = gl_add(gl_sub(ac, bb), cc);
    r.t = t;
    return r;
}

#define XDIV_BATCH 8
#define XDIV_THREADS 128
// out[3 * (k * n_open + i) ..] = x_k / (x_k - xi_i),  x_k = 7 * w_E^k.  One thread: opening i, XDIV_BATCH points k.
__global__ void __launch_bounds__(XDIV_THREADS) xdiv_kernel(const u64* __restrict__ xi /* n_open x 3, canonical */, u32 n_open, int ext_bits,
                                                            NttTables tb, u64* __restrict__ out) {
    const u64 E = (u64)1 << ext_bits;
    const u32 i = blockIdx.y;
    const u64 kbase = (u64)blockIdx.x * (XDIV_THREADS * XDIV_BATCH) + threadIdx.x;
    const u64 z0 = xi[3 * i], z1 = xi[3 * i + 1], z2 = xi[3 * i + 2];
    u64 x[XDIV_BATCH], pre[XDIV_BATCH];
    F3InvParts parts[XDIV_BATCH];
    u64 run = 1;
#pragma unroll
    for (int j = 0; j < XDIV_BATCH; j++) {
        const u64 k = kbase + (u64)j * XDIV_THREADS;
        const u32 e = (k < E && ext_bits > 0) ? ((u32)k << (32 - ext_bits)) : 0u;
        // w_E^k * 2^64 (Montgomery form) -> x_k = 7 * w_E^k
        x[j] = gl_canon(gl_mmul(ext_bits == 0 ? GL_MONT_ONE : ntt_root_pow(tb.bytepow, e), GL_SHIFT));
        const gl3 den = {{gl_sub(x[j], z0), gl_neg(z1), gl_neg(z2)}};
        parts[j] = gl3_inv_parts(den);
        if (k >= E) parts[j].t = 1;            // padding lanes must not poison the shared inversion
        pre[j] = run;                          // product of t_0 .. t_{j-1}
        run = gl_mul(run, parts[j].t);
    }
    u64 inv = gl_inv(run);                     // one inversion for the batch
#pragma unroll
    for (int j = XDIV_BATCH - 1; j >= 0; j--) {
        const u64 tinv = gl_mul(inv, pre[j]);  // 1 / t_j
        inv = gl_mul(inv, parts[j].t);
        const u64 k = kbase + (u64)j * XDIV_THREADS;
        if (k < E) {
            const u64 s = gl_mul(tinv, x[j]);
            u64* o = out + 3 * (k * n_open + i);
            o[0] = gl_canon(gl_mul(parts[j].i1, s));
            o[1] = gl_canon(gl_mul(parts[j].i2, s));
            o[2] = gl_canon(gl_mul(parts[j].i3, s));
        }
    }
}
