// Goldilocks NTT / INTT / low-degree extension over row-major buffers buff[row*nCols + col] for sm_100a.
//
// Replaces (semantics, not structure) the reference's blocked worker-thread FFT:
//   fft / ifft      src/helpers/fft/fft_p.js:114-184   (== F.fft / F.ifft per column, src/helpers/fft/fft.js:118-174)
//   interpolate     src/helpers/fft/fft_p.js:187-297   (dst[j*C+c] = P_c(7 * w_ext^j), == extendPol polutils.js:18-30)
//
// Design (see DESIGN.md "NTT"):
//  * A transform of 2^n rows is split into passes of t <= NTT_TMAX bits.  One CTA owns a tile of 2^t rows x
//    NTT_W columns in shared memory; rows of a tile are 2^lo apart in the transform index, the NTT_W columns
//    are adjacent in memory (NTT_W*8 = 128 contiguous bytes per row segment -> full-sector, coalesced access).
//  * Inside a tile the butterflies run radix-8 in registers (3 layers per shared-memory round trip).
//  * DIF passes go natural -> bit-reversed, DIT passes bit-reversed -> natural, so the LDE needs no permutation pass:
//        INTT (DIF, inverse roots)  ->  coefficients in bit-reversed order
//        -> B coset NTTs of size N (DIT, forward roots) written interleaved (row B*m + r) == NTT_E of the
//        zero-padded polynomial, in natural order.
//    The last INTT pass, the 1/N scale and the first NTT pass of every coset are one kernel; every pass after the first
//    runs in place in dst, so the LDE needs no scratch buffer beyond dst itself.
//  * No element is ever multiplied by a "glue" factor.  Both the Cooley-Tukey factor between passes
//    (w_m^(base_lo * bitrev(k))) and the coset scaling ((7 w_E^r)^i on coefficient i) are geometric in the index, so they
//    fold into the butterfly twiddles as one constant per layer:   layer s of a tile uses  G[s] * w_{2^(s+1)}^j  with
//        G[s] = w_m^(+-base_lo * 2^(t-1-s))  *  (7 w_E^r)^(2^(n-1-lo-s))          (second factor: coset passes only).
//    Each CTA builds that 2^t-entry table once in shared memory (one multiply per entry, shared by the NTT_W columns).
//  * Arithmetic: twiddles are stored in Montgomery form (w * 2^64), so a butterfly is one gl_mmul whose result is
//    canonical, followed by the 3/5-instruction gl_addc / gl_subc (gl.cuh).  DIT data stays lazy (any u64) between layers
//    and passes; DIF data stays canonical (the sum uses the complement trick in ntt_cadd).
#pragma once
#include "gl.cuh"

#define NTT_W 16        // columns per tile (128 contiguous bytes per row segment)
#define NTT_TMAX 9      // max log2(rows) per tile: 2^9 * 16 * 8 B = 64 KiB (+ twiddles) -> 3 CTAs / SM
#define NTT_THREADS 256

struct NttTables {
    const u64* bytepow;   // [4][256]: W32^(b << (8k)) * 2^64
    const u64* tw_fwd;    // [2^NTT_TMAX]: entry (1 << s) + j = w_{2^(s+1)}^j * 2^64, j < 2^s  (entry 0 unused)
    const u64* tw_inv;    // same with inverse roots
    const u64* pow7;      // [32]: 7^(2^k) * 2^64
};
#define NTT_TABLE_WORDS (1024 + 2 * (1 << NTT_TMAX) + 32)

// W32^E * 2^64 (canonical) for a 32-bit exponent E: any root of unity of order <= 2^32 to any power.
GL_D u64 ntt_root_pow(const u64* __restrict__ bytepow, u32 E) {
    u64 r = bytepow[3 * 256 + (E >> 24)];
    const u32 b2 = (E >> 16) & 255, b1 = (E >> 8) & 255, b0 = E & 255;
    if (b2) r = gl_mmul(r, bytepow[2 * 256 + b2]);
    if (b1) r = gl_mmul(r, bytepow[1 * 256 + b1]);
    if (b0) r = gl_mmul(r, bytepow[b0]);
    return r;
}

__device__ __forceinline__ u32 ntt_bitrev(u32 x, int bits) { return bits == 0 ? 0u : (__brev(x) >> (32 - bits)); }

// ---- setup kernel ---------------------------------------------------------------------------------------
__global__ void ntt_setup_tables(u64* bytepow, u64* tw_fwd, u64* tw_inv, u64* pow7) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1024) {
        const int k = i >> 8, b = i & 255;
        bytepow[i] = gl_canon(gl_to_mont(gl_pow(GL_W32, (u64)b << (8 * k))));
    }
    if (i >= 1 && i < (1 << NTT_TMAX)) {
        const int s = 31 - __clz(i), j = i - (1 << s);
        const u64 w = gl_pow(GL_W32, 1ULL << (32 - (s + 1)));   // w_{2^(s+1)}
        const u64 f = gl_canon(gl_pow(w, (u64)j));
        tw_fwd[i] = gl_canon(gl_to_mont(f));
        tw_inv[i] = gl_canon(gl_to_mont(gl_inv(f)));
    }
    if (i == 0) tw_fwd[0] = tw_inv[0] = GL_MONT_ONE;
    if (i < 32) {
        u64 v = GL_SHIFT;
        for (int k = 0; k < i; k++) v = gl_mul(v, v);
        pow7[i] = gl_canon(gl_to_mont(v));
    }
}

// ---- per-tile twiddle table -------------------------------------------------------------------------------
// TW[(1 << s) + j] = G[s] * w_{2^(s+1)}^(+-j) * 2^64 for the tile's layers s < t (global layers lo + s of a 2^n transform).
//   base_lo   low index bits of the tile (the Cooley-Tukey glue between passes), 0 when lo == 0
//   coset_r   < 0: plain transform;  >= 0: forward NTT on the coset 7 * w_E^r (E = 2^ext_bits)
// Must be called by all threads; ends with a barrier.
template <bool INVERSE>
__device__ __forceinline__ void ntt_build_tw(u64* __restrict__ TW, u64* __restrict__ G, int t, int lo, u32 base_lo, int coset_r, int n,
                                             int ext_bits, const NttTables& tb) {
    const u64* __restrict__ base = INVERSE ? tb.tw_inv : tb.tw_fwd;
    const bool plain = (base_lo == 0) && (coset_r < 0);
    if (plain) {
        for (int i = threadIdx.x; i < (1 << t); i += NTT_THREADS) TW[i] = base[i];
        __syncthreads();
        return;
    }
    if ((int)threadIdx.x < t) {
        const int s = threadIdx.x, sg = lo + s;
        u32 E = base_lo << (31 - sg);                    // w_m^(base_lo * 2^(t-1-s)) as a power of W32
        if (INVERSE) E = 0u - E;
        u64 g;
        if (coset_r >= 0) {
            const int k = n - 1 - sg;                    // (7 w_E^r)^(2^k)
            if (coset_r > 0) E += ((u32)coset_r << (32 - ext_bits)) << k;
            g = gl_mmul(ntt_root_pow(tb.bytepow, E), tb.pow7[k]);
        } else {
            g = ntt_root_pow(tb.bytepow, E);
        }
        G[s] = g;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (1 << t); i += NTT_THREADS) {
        const int s = (i == 0) ? 0 : 31 - __clz(i);
        TW[i] = gl_mmul(G[s], base[i]);
    }
    __syncthreads();
}


// ---- tile load ---------------------------------------------------------------------------------------------------
// 16-byte cp.async (LDGSTS) copies: every thread puts all of its row segments in flight at once and the data never
// passes through registers, so one DRAM latency covers the whole tile while the CTA builds its twiddle table.
GL_D void ntt_cp_async16(u64* smem_dst, const u64* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
GL_D void ntt_cp_async_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// ---- butterflies -----------------------------------------------------------------------------------------------
// canonical a, b -> canonical (a + b) mod p:  a - (p - b), + p on borrow.  7 ALU.
GL_D u64 ntt_cadd(u64 a, u64 b) { return gl_subc(a, GL_P - b); }

// tile: [2^t][NTT_W] u64 in shared memory.  TW: per-tile table (see ntt_build_tw).
// One "step" handles R consecutive layers in registers; `low` is the lowest k-bit of the step.
template <int R, bool DIF>
__device__ __forceinline__ void ntt_tile_step(u64* __restrict__ tile, const u64* __restrict__ TW, int t, int low) {
    const int groups = 1 << (t - R);
    const int c = threadIdx.x % NTT_W;
    for (int g = threadIdx.x / NTT_W; g < groups; g += NTT_THREADS / NTT_W) {
        const int glow = g & ((1 << low) - 1);
        const int kbase = ((g >> low) << (low + R)) | glow;
        u64* __restrict__ p = tile + kbase * NTT_W + c;
        u64 v[1 << R];
#pragma unroll
        for (int i = 0; i < (1 << R); i++) v[i] = p[(i << low) * NTT_W];
        if (DIF) {
#pragma unroll
            for (int lb = R - 1; lb >= 0; lb--) {
                const int b = low + lb;   // k-bit of this layer: pairs (k, k + 2^b), twiddle index k mod 2^b
#pragma unroll
                for (int i = 0; i < (1 << R); i++) {
                    if (i & (1 << lb)) continue;
                    const int x = ((i & ((1 << lb) - 1)) << low) | glow;
                    const u64 w = TW[(1 << b) + x];
                    const u64 a = v[i], bb = v[i | (1 << lb)];
                    v[i] = ntt_cadd(a, bb);
                    v[i | (1 << lb)] = gl_mmul(gl_subc(a, bb), w);
                }
            }
        } else {
#pragma unroll
            for (int lb = 0; lb < R; lb++) {
                const int b = low + lb;
#pragma unroll
                for (int i = 0; i < (1 << R); i++) {
                    if (i & (1 << lb)) continue;
                    const int x = ((i & ((1 << lb) - 1)) << low) | glow;
                    const u64 w = TW[(1 << b) + x];
                    const u64 a = v[i], tb = gl_mmul(v[i | (1 << lb)], w);
                    v[i] = gl_addc(a, tb);
                    v[i | (1 << lb)] = gl_subc(a, tb);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < (1 << R); i++) p[(i << low) * NTT_W] = v[i];
    }
}

// Full in-tile transform of 2^t points per column.  DIF: natural -> bit-reversed (layers from the top), canonical in/out;
// DIT: bit-reversed -> natural (layers from the bottom), any u64 in/out.  Caller syncs before; this syncs after every step.
template <bool DIF>
__device__ __forceinline__ void ntt_tile_transform(u64* tile, const u64* TW, int t) {
    if (DIF) {
        int top = t;   // bits [0, top) still to do
        while (top >= 3) { ntt_tile_step<3, true>(tile, TW, t, top - 3); top -= 3; __syncthreads(); }
        if (top == 2) { ntt_tile_step<2, true>(tile, TW, t, 0); __syncthreads(); }
        if (top == 1) { ntt_tile_step<1, true>(tile, TW, t, 0); __syncthreads(); }
    } else {
        int low = 0;
        const int rem = t % 3;
        if (rem == 1) { ntt_tile_step<1, false>(tile, TW, t, 0); low = 1; __syncthreads(); }
        if (rem == 2) { ntt_tile_step<2, false>(tile, TW, t, 0); low = 2; __syncthreads(); }
        while (low < t) { ntt_tile_step<3, false>(tile, TW, t, low); low += 3; __syncthreads(); }
    }
}

// Shared-memory layout: tile | [tile2] | TW (2^t) | G (16)
static inline size_t ntt_smem_bytes(int t, bool two_tiles) {
    const size_t words = ((size_t)NTT_W << t) * (two_tiles ? 2 : 1) + ((size_t)1 << t) + 16;
    return words * sizeof(u64);
}

struct NttPass {
    u64 C;             // columns of the buffer
    int n, lo, t;      // transform size, first index bit of this pass, bits of this pass
    u64 in_mul, out_mul;   // row = position * mul + z   (mul = 1: natural buffer, B: interleaved LDE layout)
    int bitrev_in;     // gather the input rows in bit-reversed order (first DIT pass of a natural -> natural transform)
    int canon_in;      // DIF: the input may hold non-canonical words (first pass reads the caller's buffer)
    int canon_out;     // DIT: canonicalise on store (last pass)
    int coset;         // forward LDE passes: blockIdx.z is the coset index r
    int ext_bits;
    u64 scale;         // Montgomery-form factor applied on store (plain INTT: 1/N), 0 = none
};

// ---- generic pass ------------------------------------------------------------------------------------------
// One pass over bits [lo, lo+t) of a 2^n-point transform of every column.
//   position(k) = (base_hi << (lo+t)) | (k << lo) | base_lo,   tile id = (base_hi << lo) | base_lo
// gridDim.x = 2^(n-t) tiles, gridDim.y = ceil(C / NTT_W) column chunks, gridDim.z = cosets.
template <bool DIF, bool INVERSE>
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(const u64* __restrict__ in, u64* __restrict__ out, NttPass P, NttTables tb) {
    extern __shared__ u64 ntt_smem[];
    const int t = P.t, lo = P.lo;
    u64* tile = ntt_smem;
    u64* TW = tile + ((size_t)NTT_W << t);
    u64* G = TW + ((size_t)1 << t);
    const u32 tile_id = blockIdx.x;
    const u32 base_lo = tile_id & ((1u << lo) - 1);
    const u32 base_hi = tile_id >> lo;
    const u64 c0 = (u64)blockIdx.y * NTT_W;
    const int cw = (int)((P.C - c0 < NTT_W) ? (P.C - c0) : NTT_W);
    const u64 z = blockIdx.z;
    const int rows = 1 << t;
    const int c = threadIdx.x % NTT_W;
    const u64 pos0 = ((u64)base_hi << (lo + t)) | base_lo;

    // loads first (they are in flight while the twiddle table is built)
    const bool async_ok = !P.bitrev_in && cw == NTT_W && (P.C % 2 == 0) && ((((size_t)in) & 15) == 0);
    if (async_ok) {
        const int cp = threadIdx.x % (NTT_W / 2);
        for (int k = threadIdx.x / (NTT_W / 2); k < rows; k += NTT_THREADS / (NTT_W / 2)) {
            const u64 pos = pos0 | ((u64)k << lo);
            ntt_cp_async16(tile + k * NTT_W + 2 * cp, in + (pos * P.in_mul + z) * P.C + c0 + 2 * cp);
        }
    } else {
        for (int k = threadIdx.x / NTT_W; k < rows; k += NTT_THREADS / NTT_W) {
            u64 v = 0;
            if (c < cw) {
                u64 pos = pos0 | ((u64)k << lo);
                if (P.bitrev_in) pos = (P.n == 0) ? 0 : (u64)(__brevll(pos) >> (64 - P.n));
                v = in[(pos * P.in_mul + z) * P.C + c0 + c];
            }
            tile[k * NTT_W + c] = v;
        }
    }
    ntt_build_tw<INVERSE>(TW, G, t, lo, base_lo, P.coset ? (int)z : -1, P.n, P.ext_bits, tb);   // ends with a barrier
    if (async_ok) { ntt_cp_async_wait(); __syncthreads(); }
    if (DIF && P.canon_in) {   // the caller's buffer may hold non-canonical words; DIF butterflies need canonical inputs
        for (int i = threadIdx.x; i < rows * NTT_W; i += NTT_THREADS) tile[i] = gl_canon(tile[i]);
        __syncthreads();
    }
    ntt_tile_transform<DIF>(tile, TW, t);
    for (int k = threadIdx.x / NTT_W; k < rows; k += NTT_THREADS / NTT_W) {
        if (c < cw) {
            u64 v = tile[k * NTT_W + c];
            if (P.scale) v = gl_mmul(v, P.scale);
            else if (!DIF && P.canon_out) v = gl_canon(v);
            const u64 pos = pos0 | ((u64)k << lo);
            out[(pos * P.out_mul + z) * P.C + c0 + c] = v;
        }
    }
}

// ---- fused LDE middle kernel ---------------------------------------------------------------------------------
// Tile = 2^t contiguous transform positions q (lo = 0).  Finishes the INTT (last t DIF layers, inverse roots), scales by
// 1/N, then for every coset r < B runs the first t DIT layers of the size-N forward NTT on 7 w_E^r <w_N> (coset folded into
// the twiddles) and stores to row (B*q + r).  Input row = q*in_mul (src: in_mul = 1; dst: in_mul = B).
__global__ void __launch_bounds__(NTT_THREADS) ntt_lde_fused_kernel(const u64* __restrict__ in, u64* __restrict__ out, u64 C, int n,
                                                                    int ext_bits, int t, u64 in_mul, u64 n_inv_mont, int canon_in,
                                                                    int canon_out, NttTables tb) {
    extern __shared__ u64 ntt_smem[];
    u64* tile = ntt_smem;
    u64* tile2 = tile + ((size_t)NTT_W << t);
    u64* TW = tile2 + ((size_t)NTT_W << t);
    u64* G = TW + ((size_t)1 << t);
    const int B = 1 << (ext_bits - n);
    const u64 q0 = (u64)blockIdx.x << t;
    const u64 c0 = (u64)blockIdx.y * NTT_W;
    const int cw = (int)((C - c0 < NTT_W) ? (C - c0) : NTT_W);
    const int rows = 1 << t;
    const int c = threadIdx.x % NTT_W;
    const bool async_ok = cw == NTT_W && (C % 2 == 0) && ((((size_t)in) & 15) == 0);
    if (async_ok) {
        const int cp = threadIdx.x % (NTT_W / 2);
        for (int k = threadIdx.x / (NTT_W / 2); k < rows; k += NTT_THREADS / (NTT_W / 2))
            ntt_cp_async16(tile + k * NTT_W + 2 * cp, in + ((q0 | (u64)k) * in_mul) * C + c0 + 2 * cp);
    } else {
        for (int k = threadIdx.x / NTT_W; k < rows; k += NTT_THREADS / NTT_W)
            tile[k * NTT_W + c] = (c < cw) ? in[((q0 | (u64)k) * in_mul) * C + c0 + c] : 0;
    }
    ntt_build_tw<true>(TW, G, t, 0, 0, -1, n, ext_bits, tb);
    if (async_ok) { ntt_cp_async_wait(); __syncthreads(); }
    if (canon_in) {
        for (int i = threadIdx.x; i < rows * NTT_W; i += NTT_THREADS) tile[i] = gl_canon(tile[i]);
        __syncthreads();
    }
    ntt_tile_transform<true>(tile, TW, t);          // coefficients, bit-reversed: position q holds a_{bitrev_n(q)}
    for (int i = threadIdx.x; i < rows * NTT_W; i += NTT_THREADS) tile[i] = gl_mmul(tile[i], n_inv_mont);   // * 1/N, once
    for (int r = 0; r < B; r++) {
        ntt_build_tw<false>(TW, G, t, 0, 0, r, n, ext_bits, tb);   // ends with a barrier (also orders the scale pass above)
        u64* work = tile;                                          // the last coset transforms the coefficients in place
        if (r + 1 < B) {
            work = tile2;
            for (int i = threadIdx.x; i < rows * NTT_W; i += NTT_THREADS) tile2[i] = tile[i];
            __syncthreads();
        }
        ntt_tile_transform<false>(work, TW, t);
        for (int k = threadIdx.x / NTT_W; k < rows; k += NTT_THREADS / NTT_W) {
            if (c < cw) {
                u64 v = work[k * NTT_W + c];
                if (canon_out) v = gl_canon(v);
                out[(((q0 | (u64)k) << (ext_bits - n)) + r) * C + c0 + c] = v;
            }
        }
        __syncthreads();
    }
}

// ---- host-side planning and launch -------------------------------------------------------------------------
struct NttPlan {
    int npass;
    int bits[8];   // bits per pass, from the top of the index (DIF order)
};
static inline NttPlan ntt_plan(int n, int tmax) {
    NttPlan p;
    p.npass = n == 0 ? 1 : (n + tmax - 1) / tmax;
    int rem = n;
    for (int i = 0; i < p.npass; i++) {
        int left = p.npass - i;
        int b = (rem + left - 1) / left;
        p.bits[i] = b;
        rem -= b;
    }
    return p;
}

template <typename K>
static inline cudaError_t ntt_set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// natural -> natural transform of every column; src != dst.  DIT passes: the first one gathers its rows from src in
// bit-reversed order, every later pass runs in place in dst.  Returns the number of kernel launches or -1.
static int ntt_launch_transform(const u64* src, u64* dst, u64 C, int n, bool inverse, const NttTables& tb, cudaStream_t st) {
    NttPlan p = ntt_plan(n, NTT_TMAX);
    const u64 n_inv = glh_to_mont(glh_inv((1ULL << n) % GL_P));
    int lo = 0;
    const unsigned ychunks = (unsigned)((C + NTT_W - 1) / NTT_W);
    int launches = 0;
    for (int i = p.npass - 1; i >= 0; i--) {
        const int t = p.bits[i];
        const bool first = (lo == 0), last = (i == 0);
        NttPass P;
        P.C = C; P.n = n; P.lo = lo; P.t = t; P.in_mul = 1; P.out_mul = 1;
        P.bitrev_in = first ? 1 : 0; P.canon_in = 0; P.canon_out = last ? 1 : 0; P.coset = 0; P.ext_bits = n;
        P.scale = (last && inverse) ? n_inv : 0;
        const u64* in = first ? src : dst;
        dim3 grid(1u << (n - t), ychunks, 1);
        const size_t smem = ntt_smem_bytes(t, false);
        if (inverse) {
            if (ntt_set_smem(ntt_pass_kernel<false, true>, smem) != cudaSuccess) return -1;
            ntt_pass_kernel<false, true><<<grid, NTT_THREADS, smem, st>>>(in, dst, P, tb);
        } else {
            if (ntt_set_smem(ntt_pass_kernel<false, false>, smem) != cudaSuccess) return -1;
            ntt_pass_kernel<false, false><<<grid, NTT_THREADS, smem, st>>>(in, dst, P, tb);
        }
        launches++;
        lo += t;
    }
    return launches;
}

// LDE src (2^n rows) -> dst (2^ext rows), all in dst after the first pass.  Returns the number of kernel launches or -1.
static int ntt_launch_lde(const u64* src, u64* dst, u64 C, int n, int ext_bits, const NttTables& tb, cudaStream_t st) {
    NttPlan p = ntt_plan(n, NTT_TMAX);
    const int B = 1 << (ext_bits - n);
    const u64 n_inv = glh_to_mont(glh_inv((1ULL << n) % GL_P));
    const unsigned ychunks = (unsigned)((C + NTT_W - 1) / NTT_W);
    int launches = 0;
    // INTT passes (DIF, inverse roots) except the last: src/dst rows q -> dst rows B*q
    int hi = n;
    for (int i = 0; i + 1 < p.npass; i++) {
        const int t = p.bits[i];
        const int lo = hi - t;
        NttPass P;
        P.C = C; P.n = n; P.lo = lo; P.t = t; P.in_mul = (i == 0) ? 1 : (u64)B; P.out_mul = (u64)B;
        P.bitrev_in = 0; P.canon_in = (i == 0) ? 1 : 0; P.canon_out = 0; P.coset = 0; P.ext_bits = ext_bits; P.scale = 0;
        dim3 grid(1u << (n - t), ychunks, 1);
        const size_t smem = ntt_smem_bytes(t, false);
        if (ntt_set_smem(ntt_pass_kernel<true, true>, smem) != cudaSuccess) return -1;
        ntt_pass_kernel<true, true><<<grid, NTT_THREADS, smem, st>>>(i == 0 ? src : dst, dst, P, tb);
        launches++;
        hi = lo;
    }
    // fused middle pass
    const int tf = p.bits[p.npass - 1];
    {
        dim3 grid(1u << (n - tf), ychunks, 1);
        const size_t smem = ntt_smem_bytes(tf, true);
        if (ntt_set_smem(ntt_lde_fused_kernel, smem) != cudaSuccess) return -1;
        const bool only = (p.npass == 1);
        ntt_lde_fused_kernel<<<grid, NTT_THREADS, smem, st>>>(only ? src : dst, dst, C, n, ext_bits, tf, only ? 1 : (u64)B, n_inv, only ? 1 : 0,
                                                              only ? 1 : 0, tb);
        launches++;
    }
    // remaining forward DIT passes, in place on the interleaved layout, one grid.z slice per coset
    int lo = tf;
    for (int i = p.npass - 2; i >= 0; i--) {
        const int t = p.bits[i];
        NttPass P;
        P.C = C; P.n = n; P.lo = lo; P.t = t; P.in_mul = (u64)B; P.out_mul = (u64)B;
        P.bitrev_in = 0; P.canon_in = 0; P.canon_out = (i == 0) ? 1 : 0; P.coset = 1; P.ext_bits = ext_bits; P.scale = 0;
        dim3 grid(1u << (n - t), ychunks, (unsigned)B);
        const size_t smem = ntt_smem_bytes(t, false);
        if (ntt_set_smem(ntt_pass_kernel<false, false>, smem) != cudaSuccess) return -1;
        ntt_pass_kernel<false, false><<<grid, NTT_THREADS, smem, st>>>(dst, dst, P, tb);
        launches++;
        lo += t;
    }
    return launches;
}
