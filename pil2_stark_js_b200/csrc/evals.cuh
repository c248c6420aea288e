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

// 160-bit accumulator for sums of 128-bit products: the modular reduction is paid once per thread, not once per product
// (a thread adds at most 2^32 products of < 2^128).
struct Acc160 {
    u64 lo, hi;
    u32 top;
};
GL_D void acc_mad(Acc160& a, u64 x, u64 y) {          // a += x * y
    const u64 pl = x * y, ph = __umul64hi(x, y);
    a.lo += pl;
    const u64 c0 = a.lo < pl;
    a.hi += ph;
    const u64 c1 = a.hi < ph;
    a.hi += c0;
    a.top += (u32)(c1 + (a.hi < c0));
}
GL_D void acc_add(Acc160& a, u64 x) {                 // a += x
    a.lo += x;
    const u64 c0 = a.lo < x;
    a.hi += c0;
    a.top += (u32)(a.hi < c0);
}
// top * 2^128 + hi * 2^64 + lo  mod p  (2^128 = -2^32 mod p; top * 2^32 <= p - 1 is canonical)
GL_D u64 acc_reduce(const Acc160& a) { return gl_sub(gl_reduce128(a.hi, a.lo), (u64)a.top << 32); }

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
            Acc160 s0 = {0, 0, 0}, s1 = {0, 0, 0}, s2 = {0, 0, 0};
            if (d.dim == 1) {
#pragma unroll 4
                for (u64 k = kb + r; k < ke; k += rw) {
                    const u64 v = buf[(k << eb) * size + d.offset];
                    acc_mad(s0, v, lv[3 * k]);
                    acc_mad(s1, v, lv[3 * k + 1]);
                    acc_mad(s2, v, lv[3 * k + 2]);
                }
            } else {
                for (u64 k = kb + r; k < ke; k += rw) {
                    const u64* __restrict__ v = buf + (k << eb) * size + d.offset;
                    const gl3 t = gl3_mul(gl3{{v[0], v[1], v[2]}}, gl3{{lv[3 * k], lv[3 * k + 1], lv[3 * k + 2]}});
                    acc_add(s0, t.c[0]);
                    acc_add(s1, t.c[1]);
                    acc_add(s2, t.c[2]);
                }
            }
            acc = gl3{{acc_reduce(s0), acc_reduce(s1), acc_reduce(s2)}};
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

// ---- evaluations as a byte-limb GEMM on the tensor cores ---------------------------------------------------------------
// evals[col][o][c] = sum_k V[k][col] * LEv_o[k][c] is a dense contraction over the 2^nBits rows: (size x N) . (N x 3*nOpen).
// Both factors are cut into their 8 byte limbs -- which is how they already sit in memory -- and the limb products are summed
// exactly by IMMA (mma.sync m16n8k32, u8 x u8 -> s32):
//     D[m][n] = sum_k A[m][k] * B[k][n],   m = (o*3 + c)*8 + b'  (byte b' of LEv_o[k][c]),   n = byte b of V[k][col]
// One n8 tile is one trace column, one half m16 tile is one (opening, coordinate) pair.  A K-chunk is at most 2^15 rows so
// that the s32 accumulators cannot overflow (2^15 * 255^2 < 2^31); at the end of a chunk every 8 x 8 block of limb sums is
// recombined, sum_{b,b'} D * 2^(8(b+b')) mod p, with one warp reduction, and written as a field element.  What is left is
// HBM-bound: the base rows of the buffer are read once, whatever the number of evaluations.
// LEv arrives pre-transposed by lev_bytes_kernel as LT[kstep][m][32 bytes of k], so an A fragment is one 32-bit load; the V
// tile (32 rows x 32 columns) is staged in shared memory with cp.async and byte-transposed 4 x 4 with PRMT.
#define EVM_THREADS 256
#define EVM_COLS 32              // columns per CTA (4 per warp)
#define EVM_KSTEP 32
#define EVM_PITCH 264            // bytes per staged row (256 + 8: rows 4 apart land 8 banks apart)
#define EVM_MAX_CHUNK 32768
#define EVM_STAGES 4             // cp.async ring: 3 K-steps (24 KiB per CTA) in flight hide the DRAM latency

// LT[(kstep * M + m) * 32 + (k % 32)] = byte (m % 8) of lev[(m / 24) * n * 3 + k * 3 + (m / 8) % 3]; rows m >= 24 * n_lev are zero.
__global__ void lev_bytes_kernel(const u64* __restrict__ lev, u64 n, u32 n_lev, u32 M, unsigned char* __restrict__ LT) {
    // one thread per 4 output bytes (k % 32 = 4q .. 4q+3 of one m): consecutive threads write consecutive words
    const u64 total = ((n + EVM_KSTEP - 1) / EVM_KSTEP) * M * (EVM_KSTEP / 4), stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const u32 q = (u32)(i % (EVM_KSTEP / 4));
        const u64 rest = i / (EVM_KSTEP / 4);
        const u32 m = (u32)(rest % M);
        const u64 ks = rest / M;
        u32 w = 0;
        if (m < 24 * n_lev) {
            const u32 o = m / 24, c = (m / 8) % 3, b = m % 8;
            const u64* __restrict__ src = lev + (u64)o * n * 3 + c;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u64 k = ks * EVM_KSTEP + 4 * q + j;
                if (k < n) w |= (u32)((src[k * 3] >> (8 * b)) & 0xFF) << (8 * j);
            }
        }
        reinterpret_cast<u32*>(LT)[i] = w;
    }
}

GL_D void evm_mma(int (&d)[4], const u32 (&a)[4], u32 b0, u32 b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
GL_D void evm_cp8(void* smem_dst, const void* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}

// grid = (column groups, K-chunks).  partial[(chunk * size + col) * (3 * n_lev) + oc] = sum over the chunk's rows (field element).
template <int MT>
__global__ void __launch_bounds__(EVM_THREADS) evals_mma_kernel(const u64* __restrict__ buf, u64 size, int eb, u64 n, u64 rows_per_chunk,
                                                                const unsigned char* __restrict__ LT, u32 n_lev, u64* __restrict__ partial) {
    __shared__ __align__(16) unsigned char Vs[EVM_STAGES][EVM_KSTEP * EVM_PITCH];
    __shared__ u64 pow8[16];                                  // 2^(8d) mod p, d < 15
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const u64 col0 = (u64)blockIdx.x * EVM_COLS;
    const u64 kb = (u64)blockIdx.y * rows_per_chunk;
    const u64 ke = kb + rows_per_chunk < n ? kb + rows_per_chunk : n;
    const u32 M = MT * 16;
    if (threadIdx.x < 16) {
        u64 v = 1;
        for (int i = 0; i < (int)threadIdx.x; i++) v = gl_canon(gl_mul(v, 256));
        pow8[threadIdx.x] = v;
    }
    int acc[MT][4][4];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int t = 0; t < 4; t++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][t][j] = 0;

    const u64 nsteps = (ke > kb) ? (ke - kb + EVM_KSTEP - 1) / EVM_KSTEP : 0;
    // stage loader: 32 rows x 32 columns of 8 bytes = 1024 copies, 4 per thread; always commits a group (possibly empty) so that
    // the wait_group arithmetic below is uniform
    auto stage = [&](u64 step) {
        if (step < nsteps) {
            const int sbuf = (int)(step % EVM_STAGES);
            const u64 k0 = kb + step * EVM_KSTEP;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int idx = threadIdx.x + q * EVM_THREADS;
                const int r = idx >> 5, c = idx & 31;
                unsigned char* dst = &Vs[sbuf][r * EVM_PITCH + c * 8];
                const u64 k = k0 + r;
                if (k < ke && col0 + c < size) evm_cp8(dst, buf + (k << eb) * size + col0 + c);
                else *reinterpret_cast<u64*>(dst) = 0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // A fragments of one K-step: LT[kstep][m][k]
    auto load_a = [&](u32 (&a)[MT][4], u64 step) {
        const unsigned char* __restrict__ lt = LT + ((kb / EVM_KSTEP + step) * M) * EVM_KSTEP;
#pragma unroll
        for (int i = 0; i < MT; i++) {
            const unsigned char* p0 = lt + (i * 16 + gid) * EVM_KSTEP + tig * 4;
            a[i][0] = *reinterpret_cast<const u32*>(p0);
            a[i][1] = *reinterpret_cast<const u32*>(p0 + 8 * EVM_KSTEP);
            a[i][2] = *reinterpret_cast<const u32*>(p0 + 16);
            a[i][3] = *reinterpret_cast<const u32*>(p0 + 8 * EVM_KSTEP + 16);
        }
    };
#pragma unroll
    for (int st = 0; st < EVM_STAGES - 1; st++) stage(st);
    const u32 sel = (u32)(gid & 3) | ((4u + (gid & 3)) << 4);          // PRMT: byte q of x, byte q of y
    u32 a[MT][4], an[MT][4];
    if (nsteps) load_a(a, 0);
    for (u64 s = 0; s < nsteps; s++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(EVM_STAGES - 2) : "memory");     // the group of step s has landed
        __syncthreads();                                     // ... for every thread; and everyone is done with step s - 1
        stage(s + EVM_STAGES - 1);                           // refills the buffer step s - 1 used
        if (s + 1 < nsteps) load_a(an, s + 1);
        const int cur = (int)(s % EVM_STAGES);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            // B fragment of column (warp*4 + t): byte gid of rows tig*4 .. +3 (b0) and +16 (b1)
            const unsigned char* vb = &Vs[cur][(tig * 4) * EVM_PITCH + (warp * 4 + t) * 8 + (gid & 4)];
            u32 w[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                w[j] = *reinterpret_cast<const u32*>(vb + j * EVM_PITCH);
                w[4 + j] = *reinterpret_cast<const u32*>(vb + (16 + j) * EVM_PITCH);
            }
            const u32 b0 = __byte_perm(__byte_perm(w[0], w[1], sel), __byte_perm(w[2], w[3], sel), 0x5410);
            const u32 b1 = __byte_perm(__byte_perm(w[4], w[5], sel), __byte_perm(w[6], w[7], sel), 0x5410);
#pragma unroll
            for (int i = 0; i < MT; i++) evm_mma(acc[i][t], a[i], b0, b1);
        }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) a[i][j] = an[i][j];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                         // pow8 visible even when the chunk is empty
    // recombination: (i, half) -> oc = 2i + half; lane holds limb sums (b' = gid; b = 2 tig, 2 tig + 1)
    const u32 n_oc = 3 * n_lev;
#pragma unroll
    for (int i = 0; i < MT; i++) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const u32 oc = 2 * i + half;
            if (oc >= n_oc) continue;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const u64 v0 = (u64)(u32)acc[i][t][2 * half], v1 = (u64)(u32)acc[i][t][2 * half + 1];
                u64 f = gl_add(gl_mul(v0, pow8[gid + 2 * tig]), gl_mul(v1, pow8[gid + 2 * tig + 1]));
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) f = gl_add(f, __shfl_xor_sync(0xFFFFFFFFu, f, off));
                const u64 col = col0 + warp * 4 + t;
                if (lane == 0 && col < size) partial[((u64)blockIdx.y * size + col) * n_oc + oc] = gl_canon(f);
            }
        }
    }
}

// ---- second formulation (r02): the limb weight folded into the LEv operand ---------------------------------------------------
// With A[m = (oc, b')][k = (row, b)] = byte b' of LEv[row][oc] * 2^(8b) mod p and B[k = (row, b)][n = col] = byte b of V[row][col] the
// contraction index runs over (row, limb) and the B fragment of a lane IS the 8 bytes of V[row r0 + t][col g] as they lie in memory:
// one 64-bit global load, no shared-memory staging and no byte transposition of the big operand.  The M index is (limb pair, oc) -- m-tile
// i holds limbs 2i (tile rows 0..7 = the up to 8 (opening, coordinate) pairs) and 2i+1 (rows 8..15) -- so lane 4g+t ends up with all 8
// limb sums of (oc g; columns 2t, 2t+1 of the n-tile) and recombines them alone.  The A fragments cost 7 shift-multiplies by 2^8 and 32
// byte permutes per (oc, row); warp w of the CTA builds them for k-step w of a 32-row super-step into shared memory (double buffered,
// one barrier per super-step) and all 8 warps -- 32 columns each -- use them.  s32 accumulators hold 4096 rows (4096 * 8 * 255^2 < 2^31),
// then fold into per-lane field accumulators.  n_lev <= 2 (up to 8 rows per m-tile half); more openings take evals_mma_kernel.
#define EV2_THREADS 256
#define EV2_COLS 256             // columns per CTA: 8 warps x 4 n-tiles x 8
#define EV2_FLUSH 128            // super-steps between folds of the s32 accumulators
#define EV2_RING 8               // k-steps of V (4 rows x 32 columns) in each warp's private cp.async ring
#define EV2_RPITCH 288           // bytes per staged row: 256 + 32, rows land 8 banks apart -> conflict-free 8-byte fragment loads
#define EV2_STAGE (4 * EV2_RPITCH)
#define EV2_SMEM (2 * 8 * 4 * 32 * 16 + 8 * EV2_RING * EV2_STAGE)

GL_D u64 ev2_mul256(u64 x) { return gl_reduce96((u32)(x >> 56), x << 8); }      // x * 2^8 mod p, any u64 in / out
GL_D void ev2_cp(void* smem_dst, const void* gmem_src, bool ok16, bool valid) {   // 16-byte (or 2 x 8-byte) copy, zero fill when !valid
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned n = valid ? 16u : 0u;
    if (ok16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem_src), "r"(n) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem_src), "r"(n >> 1) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa + 8), "l"(reinterpret_cast<const char*>(gmem_src) + 8), "r"(n >> 1) : "memory");
    }
}

__global__ void __launch_bounds__(EV2_THREADS, 2) evals_mma2_kernel(const u64* __restrict__ buf, u64 size, int eb, u64 n, u64 rows_per_chunk,
                                                                    const u64* __restrict__ lev, u32 n_lev, u64* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char ev2_smem[];
    uint4 (*Abuf)[8 * 4 * 32] = reinterpret_cast<uint4 (*)[8 * 4 * 32]>(ev2_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    unsigned char* ring = ev2_smem + 2 * 8 * 4 * 32 * 16 + warp * (EV2_RING * EV2_STAGE);
    const u32 n_oc = 3 * n_lev;
    const u64 kb = (u64)blockIdx.y * rows_per_chunk;
    const u64 ke = kb + rows_per_chunk < n ? kb + rows_per_chunk : n;           // both multiples of 32
    const u64 nsuper = (ke > kb) ? (ke - kb) / 32 : 0;
    const u64 cbase = (u64)blockIdx.x * EV2_COLS + 32 * warp;
    const u64 rstride = size << eb;                                              // words between consecutive base rows
    const u64* __restrict__ lp = (g < (int)n_oc) ? lev + (u64)(g / 3) * n * 3 + (g % 3) : nullptr;

    // A fragments of k-step `warp` of super-step S -> Abuf[S & 1]
    auto prep = [&](u64 S) {
        const u64 r = kb + 32 * S + 4 * warp + t;
        u64 x = (lp && r < ke) ? lp[r * 3] : 0;
        u32 lo[8], hi[8];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            lo[b] = (u32)x;
            hi[b] = (u32)(x >> 32);
            x = ev2_mul256(x);
        }
        uint4* dst = &Abuf[S & 1][(warp * 4) * 32 + lane];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u32* w = (i < 2) ? lo : hi;
            const u32 sel = (i & 1) ? 0x7362u : 0x5140u;                         // (w0[b], w1[b], w0[b+1], w1[b+1]), b = 0 or 2
            const u32 p01 = __byte_perm(w[0], w[1], sel), p23 = __byte_perm(w[2], w[3], sel);
            const u32 p45 = __byte_perm(w[4], w[5], sel), p67 = __byte_perm(w[6], w[7], sel);
            dst[i * 32] = make_uint4(__byte_perm(p01, p23, 0x5410), __byte_perm(p01, p23, 0x7632), __byte_perm(p45, p67, 0x5410),
                                     __byte_perm(p45, p67, 0x7632));
        }
    };

    int acc[4][4][4];                                                            // [m-tile][n-tile][c]
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][j][c] = 0;
    u64 facc[4][2];
#pragma unroll
    for (int j = 0; j < 4; j++) facc[j][0] = facc[j][1] = 0;
    auto flush = [&]() {
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                // limbs 2i: acc[i][j][e], 2i + 1: acc[i][j][2 + e]; each < 2^31
                const u64 L = (u64)(u32)acc[0][j][e] + ((u64)(u32)acc[0][j][2 + e] << 8) + ((u64)(u32)acc[1][j][e] << 16) + ((u64)(u32)acc[1][j][2 + e] << 24);
                const u64 H = (u64)(u32)acc[2][j][e] + ((u64)(u32)acc[2][j][2 + e] << 8) + ((u64)(u32)acc[3][j][e] << 16) + ((u64)(u32)acc[3][j][2 + e] << 24);
                const u64 lo = L + (H << 32);                                   // L + 2^32 H = (H >> 32 + carry) 2^64 + lo
                const u64 v = gl_reduce128((H >> 32) + (u64)(lo < L), lo);
                facc[j][e] = gl_add(facc[j][e], v);
                acc[0][j][e] = acc[0][j][2 + e] = acc[1][j][e] = acc[1][j][2 + e] = 0;
                acc[2][j][e] = acc[2][j][2 + e] = acc[3][j][e] = acc[3][j][2 + e] = 0;
            }
    };

    // V: every warp streams its own 32 columns through a private ring of EV2_RING k-steps with cp.async (the data never passes through
    // registers, 7 KB per warp in flight).  Lane l copies the 16-byte chunks l and l + 32 of a k-step: chunk c = row c / 16, columns
    // 2 (c % 16), +1.  Chunks past the row end are zero filled; odd row lengths (unaligned rows) copy 8 bytes at a time.
    const bool ok16 = ((size & 1) == 0) && ((reinterpret_cast<size_t>(buf) & 15) == 0);
    const u64 nk = nsuper * 8;
    const int crow = lane >> 4, ccol = 2 * (lane & 15);
    const bool v0 = cbase + ccol < size, v1 = cbase + ccol + 1 < size;
    const u64* __restrict__ nsrc = buf + ((kb + crow) << eb) * size + cbase + ccol;   // next k-step to copy, chunk `lane`; chunk lane + 32 is 2 rows down
    const u64 kstride = 4 * rstride, r2 = 2 * rstride;
    u32 kleft = (u32)nk, nslot = 0, cslot = 0;
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring) + crow * EV2_RPITCH + ccol * 8;
    auto issue = [&]() {                                     // next k-step (if any) -> next ring slot; always commits a group
        if (kleft) {
            const unsigned da = ring_sa + nslot * EV2_STAGE;
            if (ok16) {                                                          // even size: both columns of a chunk are in or out together
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(v0 ? nsrc : buf), "r"(v0 ? 16u : 0u) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da + 2 * EV2_RPITCH), "l"(v0 ? nsrc + r2 : buf), "r"(v0 ? 16u : 0u) : "memory");
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const u64* sh = nsrc + h * r2;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(da + 2 * h * EV2_RPITCH), "l"(v0 ? sh : buf), "r"(v0 ? 8u : 0u) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(da + 2 * h * EV2_RPITCH + 8), "l"(v1 ? sh + 1 : buf), "r"(v1 ? 8u : 0u) : "memory");
                }
            }
            nsrc += kstride;
            kleft--;
        }
        nslot = (nslot + 1) & (EV2_RING - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nsuper) prep(0);
#pragma unroll
    for (int k = 0; k < EV2_RING - 1; k++) issue();
    const unsigned char* vbase = ring + t * EV2_RPITCH + g * 8;
    for (u64 S0 = 0; S0 < nsuper; S0 += EV2_FLUSH) {
        const u64 S1 = S0 + EV2_FLUSH < nsuper ? S0 + EV2_FLUSH : nsuper;
#pragma unroll 1
        for (u64 S = S0; S < S1; S++) {
            __syncthreads();                                 // fragments of super-step S are written; everyone is done reading those of S - 1
            if (S + 1 < nsuper) prep(S + 1);
            const uint4* __restrict__ ab = &Abuf[S & 1][lane];
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                asm volatile("cp.async.wait_group %0;" ::"n"(EV2_RING - 2) : "memory");      // this lane's copies of the k-step have landed
                __syncwarp();                                                                // ... and every other lane's
                const unsigned char* vb = vbase + cslot * EV2_STAGE;
                cslot = (cslot + 1) & (EV2_RING - 1);
                u64 vv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) vv[j] = *reinterpret_cast<const u64*>(vb + 64 * j);
                __syncwarp();                                // the slot read one k-step ago is free in every lane: refill it
                issue();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint4 a = ab[(ks * 4 + i) * 32];
                    const u32 af[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) evm_mma(acc[i][j], af, (u32)vv[j], (u32)(vv[j] >> 32));
                }
            }
        }
        flush();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (g < (int)n_oc) {
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const u64 col = cbase + 8 * j + 2 * t + e;
                if (col < size) partial[((u64)blockIdx.y * size + col) * n_oc + g] = gl_canon(facc[j][e]);
            }
    }
}

// out[e] from the dense per-chunk results: dim 1 -> (o, c) of the column; dim 3 -> R(col) + x R(col+1) + x^2 R(col+2) in F3, with
// x (r0, r1, r2) = (r2, r0 + r2, r1).
__global__ void evals_gather_kernel(const u64* __restrict__ partial, u32 chunks, u64 size, u32 n_lev, const EvalDesc* __restrict__ desc, u32 n_evals,
                                    u64* __restrict__ out) {
    // one warp per evaluation: the lanes stride over the chunks, then a shuffle reduction
    const u32 e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n_evals) return;
    const EvalDesc d = desc[e];
    const u32 n_oc = 3 * n_lev;
    gl3 r[3];
    for (u32 j = 0; j < d.dim; j++) {
        gl3 a = {{0, 0, 0}};
        for (u32 ch = lane; ch < chunks; ch += 32) {
            const u64* p = partial + ((u64)ch * size + d.offset + j) * n_oc + 3 * d.lev;
            a = gl3_add(a, gl3{{p[0], p[1], p[2]}});
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int c = 0; c < 3; c++) a.c[c] = gl_add(a.c[c], __shfl_xor_sync(0xFFFFFFFFu, a.c[c], off));
        r[j] = a;
    }
    if (lane != 0) return;
    gl3 v = r[0];
    if (d.dim == 3) {
        const gl3 x1 = {{r[1].c[2], gl_add(r[1].c[0], r[1].c[2]), r[1].c[1]}};                    // x * r1
        const gl3 t2 = {{r[2].c[2], gl_add(r[2].c[0], r[2].c[2]), r[2].c[1]}};                    // x * r2
        const gl3 x2 = {{t2.c[2], gl_add(t2.c[0], t2.c[2]), t2.c[1]}};                            // x^2 * r2
        v = gl3_add(gl3_add(v, x1), x2);
    }
    v = gl3_canon(v);
    out[3 * e] = v.c[0]; out[3 * e + 1] = v.c[1]; out[3 * e + 2] = v.c[2];
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
