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
//  * Passes are glued with the Cooley-Tukey twiddle w_m^(lo_index * bitrev(k)); DIF passes go natural ->
//    bit-reversed, DIT passes bit-reversed -> natural, so the LDE needs no permutation pass at all:
//        INTT (DIF, inverse roots)  ->  coefficients in bit-reversed order, scaled by (7 w_E^r)^i / N per coset r
//        -> B coset NTTs of size N (DIT, forward roots) written interleaved (row B*m + r) == NTT_E of the
//        zero-padded polynomial, in natural order.
//    The last INTT pass, the coset scaling and the first NTT pass are one kernel; every pass after the first
//    runs in place in dst, so the LDE needs no scratch buffer beyond dst itself.
//  * Twiddles: per-tile Cooley-Tukey factors come from four 256-entry tables of W32^(b << 8k) (3 multiplies each,
//    amortised over the NTT_W columns); in-tile factors from a 2^(TMAX-1) table staged in shared memory.
#pragma once
#include "gl.cuh"

#define NTT_W 16        // columns per tile (128 contiguous bytes per row segment)
#define NTT_TMAX 9      // max log2(rows) per tile: 2^9 * 16 * 8 B = 64 KiB (+ twiddles) -> 3 CTAs / SM
#define NTT_THREADS 256
#define NTT_TW_BITS 12  // in-tile twiddle table: w_{2^12}^j, j < 2^11

struct NttTables {
    const u64* bytepow;   // [4][256]: bytepow[k][b] = W32^(b << (8k))
    const u64* tw_fwd;    // [2^(NTT_TW_BITS-1)]: w_{2^TW}^j
    const u64* tw_inv;    // [2^(NTT_TW_BITS-1)]: w_{2^TW}^-j
};

// W32^E for a 32-bit exponent E (any root of unity of order <= 2^32 to any power).
GL_D u64 ntt_root_pow(const u64* __restrict__ bytepow, u32 E) {
    u64 r = bytepow[3 * 256 + (E >> 24)];
    u32 b2 = (E >> 16) & 255, b1 = (E >> 8) & 255, b0 = E & 255;
    if (b2) r = gl_mul(r, bytepow[2 * 256 + b2]);
    if (b1) r = gl_mul(r, bytepow[1 * 256 + b1]);
    if (b0) r = gl_mul(r, bytepow[b0]);
    return r;
}

__device__ __forceinline__ u32 ntt_bitrev(u32 x, int bits) { return bits == 0 ? 0u : (__brev(x) >> (32 - bits)); }

// ---- setup kernels --------------------------------------------------------------------------------------
__global__ void ntt_setup_tables(u64* bytepow, u64* tw_fwd, u64* tw_inv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1024) {
        int k = i >> 8, b = i & 255;
        bytepow[i] = gl_canon(gl_pow(GL_W32, (u64)b << (8 * k)));
    }
    if (i < (1 << (NTT_TW_BITS - 1))) {
        u64 w = gl_pow(GL_W32, 1ULL << (32 - NTT_TW_BITS));   // w_{2^TW}
        u64 f = gl_canon(gl_pow(w, (u64)i));
        tw_fwd[i] = f;
        tw_inv[i] = gl_canon(gl_inv(f));
    }
}

// Coset scale table for the LDE: scl[r][kappa] = delta_r^kappa with delta_r = (7 * w_E^r)^(2^(n-t)), kappa < 2^t,
// and base[r] = 7 * w_E^r (used per tile for gamma_r^(tile part) / N).
__global__ void ntt_setup_coset(u64* scl, u64* gam, const u64* bytepow, int n_bits, int ext_bits, int t) {
    int B = 1 << (ext_bits - n_bits);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int per = 1 << t;
    if (i >= B * per) return;
    int r = i / per, kappa = i % per;
    u64 g = gl_mul(GL_SHIFT, ntt_root_pow(bytepow, r == 0 ? 0u : ((u32)r << (32 - ext_bits))));   // 7 * w_E^r
    if (kappa == 0) gam[r] = gl_canon(g);
    u64 delta = gl_pow(g, 1ULL << (n_bits - t));
    scl[i] = gl_canon(gl_pow(delta, (u64)kappa));
}

// ---- in-tile butterflies ---------------------------------------------------------------------------------
// tile: [2^t][NTT_W] u64 in shared memory.  ltw: w^j (forward or inverse), j < 2^(t-1), already strided for t.
// One "step" handles R consecutive layers in registers; `low` is the lowest k-bit of the step.
template <int R, bool DIF>
__device__ __forceinline__ void ntt_tile_step(u64* __restrict__ tile, const u64* __restrict__ ltw, int t, int low) {
    const int groups = 1 << (t - R);
    const int items = groups * NTT_W;
    for (int item = threadIdx.x; item < items; item += NTT_THREADS) {
        const int c = item % NTT_W;
        const int g = item / NTT_W;
        const int glow = g & ((1 << low) - 1);
        const int kbase = ((g >> low) << (low + R)) | glow;
        u64 v[1 << R];
#pragma unroll
        for (int i = 0; i < (1 << R); i++) v[i] = tile[(kbase + (i << low)) * NTT_W + c];
        if (DIF) {
#pragma unroll
            for (int lb = R - 1; lb >= 0; lb--) {
                const int b = low + lb;   // k-bit of this layer: pairs (k, k + 2^b), twiddle w_{2^(b+1)}^(k mod 2^b)
#pragma unroll
                for (int i = 0; i < (1 << R); i++) {
                    if (i & (1 << lb)) continue;
                    const int x = ((i & ((1 << lb) - 1)) << low) | glow;
                    const u64 w = ltw[x << (t - b - 1)];
                    const u64 a = v[i], bb = v[i | (1 << lb)];
                    v[i] = gl_add(a, bb);
                    v[i | (1 << lb)] = gl_mul(gl_sub(a, bb), w);
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
                    const u64 w = ltw[x << (t - b - 1)];
                    const u64 a = v[i], tb = gl_mul(v[i | (1 << lb)], w);
                    v[i] = gl_add(a, tb);
                    v[i | (1 << lb)] = gl_sub(a, tb);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < (1 << R); i++) tile[(kbase + (i << low)) * NTT_W + c] = v[i];
    }
}

// Full in-tile transform of 2^t points per column.  DIF: natural -> bit-reversed (layers from the top);
// DIT: bit-reversed -> natural (layers from the bottom).  Caller syncs before; this syncs after every step.
template <bool DIF>
__device__ __forceinline__ void ntt_tile_transform(u64* tile, const u64* ltw, int t) {
    if (DIF) {
        int top = t;   // bits [0, top) still to do
        while (top >= 3) { ntt_tile_step<3, true>(tile, ltw, t, top - 3); top -= 3; __syncthreads(); }
        if (top == 2) { ntt_tile_step<2, true>(tile, ltw, t, 0); __syncthreads(); }
        if (top == 1) { ntt_tile_step<1, true>(tile, ltw, t, 0); __syncthreads(); }
    } else {
        int low = 0;
        const int rem = t % 3;
        if (rem == 1) { ntt_tile_step<1, false>(tile, ltw, t, 0); low = 1; __syncthreads(); }
        if (rem == 2) { ntt_tile_step<2, false>(tile, ltw, t, 0); low = 2; __syncthreads(); }
        while (low < t) { ntt_tile_step<3, false>(tile, ltw, t, low); low += 3; __syncthreads(); }
    }
}

// Shared-memory layout helpers
struct NttSmem {
    u64* tile;   // 2^t * NTT_W
    u64* tile2;  // second tile (LDE fused kernel only)
    u64* ltw;    // 2^(t-1) (>= 1)
    u64* ptw;    // 2^t per-row factors
};
__device__ __forceinline__ NttSmem ntt_smem_carve(u64* base, int t, bool two_tiles) {
    NttSmem s;
    s.tile = base;
    u64* p = base + ((size_t)NTT_W << t);
    s.tile2 = p;
    if (two_tiles) p += ((size_t)NTT_W << t);
    s.ltw = p;
    p += (t > 0) ? (1 << (t - 1)) : 1;
    s.ptw = p;
    return s;
}
static inline size_t ntt_smem_bytes(int t, bool two_tiles) {
    size_t words = ((size_t)NTT_W << t) * (two_tiles ? 2 : 1) + ((t > 0) ? (1u << (t - 1)) : 1) + ((size_t)1 << t);
    return words * sizeof(u64);
}

__device__ __forceinline__ void ntt_stage_local_twiddles(u64* ltw, const u64* __restrict__ table, int t) {
    const int n = (t > 0) ? (1 << (t - 1)) : 0;
    for (int j = threadIdx.x; j < n; j += NTT_THREADS) ltw[j] = table[(size_t)j << (NTT_TW_BITS - t)];
}

// ---- generic pass ------------------------------------------------------------------------------------------
// One pass over bits [lo, lo+t) of a 2^n-point transform of every column.
//   position(k) = (base_hi << (lo+t)) | (k << lo) | base_lo,   tile id = (base_hi << lo) | base_lo
//   input row   = (bitrev_in ? bitrev_n(position) : position) * in_mul + in_add
//                 (in_mul = 1 for a natural buffer, B for the interleaved LDE layout)
//   output row  = position * out_mul + out_add
// DIF: in-tile transform then multiply row k by w_m^(+-base_lo * bitrev_t(k)) (m = 2^(lo+t)), then by `scale`.
// DIT: the same factor is applied before the in-tile transform.
// gridDim.x = 2^(n-t) tiles, gridDim.y = ceil(C / NTT_W) column chunks, gridDim.z = cosets (out_add/in_add += z).
template <bool DIF, bool INVERSE>
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(const u64* __restrict__ in, u64* __restrict__ out, u64 C,
                                                               int n, int lo, int t, u64 in_mul, u64 in_add, u64 out_mul,
                                                               u64 out_add, int bitrev_in, u64 scale, NttTables tb) {
    extern __shared__ u64 ntt_smem[];
    NttSmem s = ntt_smem_carve(ntt_smem, t, false);
    const u32 tile_id = blockIdx.x;
    const u32 base_lo = tile_id & ((1u << lo) - 1);
    const u32 base_hi = tile_id >> lo;
    const u64 c0 = (u64)blockIdx.y * NTT_W;
    const int cw = (int)((C - c0 < NTT_W) ? (C - c0) : NTT_W);
    const u64 z = blockIdx.z;
    const int rows = 1 << t;

    ntt_stage_local_twiddles(s.ltw, INVERSE ? tb.tw_inv : tb.tw_fwd, t);
    // Cooley-Tukey factor per tile row: w_m^(base_lo * bitrev_t(k)), m = 2^(lo+t)
    const bool has_ptw = (lo > 0) && (base_lo != 0);
    if (lo > 0) {
        for (int k = threadIdx.x; k < rows; k += NTT_THREADS) {
            u32 e = base_lo * ntt_bitrev((u32)k, t);                 // < 2^(lo+t) <= 2^32
            u32 E = e << (32 - (lo + t));
            if (INVERSE) E = 0u - E;
            s.ptw[k] = ntt_root_pow(tb.bytepow, E);
        }
    }
    __syncthreads();
    // load
    for (int item = threadIdx.x; item < rows * NTT_W; item += NTT_THREADS) {
        const int c = item % NTT_W, k = item / NTT_W;
        u64 v = 0;
        if (c < cw) {
            u64 pos = ((u64)base_hi << (lo + t)) | ((u64)k << lo) | base_lo;
            if (bitrev_in) pos = (n == 0) ? 0 : (u64)(__brevll(pos) >> (64 - n));
            v = in[(pos * in_mul + in_add + z) * C + c0 + c];
            if (!DIF && has_ptw) v = gl_mul(v, s.ptw[k]);
        }
        s.tile[item] = v;
    }
    __syncthreads();
    ntt_tile_transform<DIF>(s.tile, s.ltw, t);
    // store
    for (int item = threadIdx.x; item < rows * NTT_W; item += NTT_THREADS) {
        const int c = item % NTT_W, k = item / NTT_W;
        if (c < cw) {
            u64 v = s.tile[item];
            if (DIF && has_ptw) v = gl_mul(v, s.ptw[k]);
            if (scale != 1) v = gl_mul(v, scale);
            const u64 pos = ((u64)base_hi << (lo + t)) | ((u64)k << lo) | base_lo;
            out[(pos * out_mul + out_add + z) * C + c0 + c] = gl_canon(v);
        }
    }
}

// ---- fused LDE middle kernel ---------------------------------------------------------------------------------
// Tile = 2^t contiguous transform positions q (lo = 0).  Finishes the INTT (last t DIF layers, inverse roots), then
// for every coset r < B: scales coefficient i = bitrev_n(q) by (7 w_E^r)^i / N, runs the first t DIT layers of the
// size-N forward NTT and stores to row (B*q + r).  Input row = q*in_mul (src: in_mul = 1; dst: in_mul = B).
__global__ void __launch_bounds__(NTT_THREADS) ntt_lde_fused_kernel(const u64* __restrict__ in, u64* __restrict__ out, u64 C, int n,
                                                                    int ext_bits, int t, u64 in_mul, u64 n_inv,
                                                                    const u64* __restrict__ scl, const u64* __restrict__ gam,
                                                                    NttTables tb) {
    extern __shared__ u64 ntt_smem[];
    NttSmem s = ntt_smem_carve(ntt_smem, t, true);
    const int B = 1 << (ext_bits - n);
    const u32 tile_id = blockIdx.x;
    const u64 c0 = (u64)blockIdx.y * NTT_W;
    const int cw = (int)((C - c0 < NTT_W) ? (C - c0) : NTT_W);
    const int rows = 1 << t;
    u64* ltw_inv = s.ltw;
    ntt_stage_local_twiddles(ltw_inv, tb.tw_inv, t);
    for (int item = threadIdx.x; item < rows * NTT_W; item += NTT_THREADS) {
        const int c = item % NTT_W, k = item / NTT_W;
        u64 v = 0;
        if (c < cw) v = in[((((u64)tile_id << t) | (u64)k) * in_mul) * C + c0 + c];
        s.tile[item] = v;
    }
    __syncthreads();
    ntt_tile_transform<true>(s.tile, ltw_inv, t);          // coefficients, bit-reversed: position q holds a_{bitrev_n(q)}
    ntt_stage_local_twiddles(s.ltw, tb.tw_fwd, t);          // (all threads passed the trailing barrier of the transform)
    const u32 tile_rev = ntt_bitrev(tile_id, n - t);        // low bits of the coefficient index
    for (int r = 0; r < B; r++) {
        // factor(k) = gamma_r^(tile_rev) / N * delta_r^(bitrev_t(k))
        const u64 base = gl_mul(gl_pow(gam[r], tile_rev), n_inv);
        for (int k = threadIdx.x; k < rows; k += NTT_THREADS) s.ptw[k] = gl_mul(base, scl[((size_t)r << t) + ntt_bitrev((u32)k, t)]);
        __syncthreads();
        for (int item = threadIdx.x; item < rows * NTT_W; item += NTT_THREADS) s.tile2[item] = gl_mul(s.tile[item], s.ptw[item / NTT_W]);
        __syncthreads();
        ntt_tile_transform<false>(s.tile2, s.ltw, t);
        for (int item = threadIdx.x; item < rows * NTT_W; item += NTT_THREADS) {
            const int c = item % NTT_W, k = item / NTT_W;
            if (c < cw) out[(((((u64)tile_id << t) | (u64)k) << (ext_bits - n)) + r) * C + c0 + c] = gl_canon(s.tile2[item]);
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
    const u64 n_inv = glh_inv((1ULL << n) % GL_P);
    int lo = 0;
    unsigned ychunks = (unsigned)((C + NTT_W - 1) / NTT_W);
    int launches = 0;
    for (int i = p.npass - 1; i >= 0; i--) {
        const int t = p.bits[i];
        const bool first = (lo == 0), last = (i == 0);
        const u64* in = first ? src : dst;
        dim3 grid(1u << (n - t), ychunks, 1);
        size_t smem = ntt_smem_bytes(t, false);
        const u64 scale = (last && inverse) ? n_inv : 1;
        if (inverse) {
            if (ntt_set_smem(ntt_pass_kernel<false, true>, smem) != cudaSuccess) return -1;
            ntt_pass_kernel<false, true><<<grid, NTT_THREADS, smem, st>>>(in, dst, C, n, lo, t, 1, 0, 1, 0, first ? 1 : 0, scale, tb);
        } else {
            if (ntt_set_smem(ntt_pass_kernel<false, false>, smem) != cudaSuccess) return -1;
            ntt_pass_kernel<false, false><<<grid, NTT_THREADS, smem, st>>>(in, dst, C, n, lo, t, 1, 0, 1, 0, first ? 1 : 0, scale, tb);
        }
        launches++;
        lo += t;
    }
    return launches;
}

// LDE src (2^n rows) -> dst (2^ext rows), all in dst after the first pass.  scl/gam: device scratch of
// (B << tmax) + B words prepared here.  Returns the number of kernel launches or -1.
static int ntt_launch_lde(const u64* src, u64* dst, u64 C, int n, int ext_bits, u64* scl, u64* gam, const NttTables& tb,
                          cudaStream_t st) {
    NttPlan p = ntt_plan(n, NTT_TMAX);
    const int B = 1 << (ext_bits - n);
    const u64 n_inv = glh_inv((1ULL << n) % GL_P);
    unsigned ychunks = (unsigned)((C + NTT_W - 1) / NTT_W);
    int launches = 0;
    // INTT passes (DIF, inverse roots) except the last: src/dst rows q -> dst rows B*q
    int hi = n;
    for (int i = 0; i + 1 < p.npass; i++) {
        const int t = p.bits[i];
        const int lo = hi - t;
        dim3 grid(1u << (n - t), ychunks, 1);
        size_t smem = ntt_smem_bytes(t, false);
        if (ntt_set_smem(ntt_pass_kernel<true, true>, smem) != cudaSuccess) return -1;
        if (i == 0)
            ntt_pass_kernel<true, true><<<grid, NTT_THREADS, smem, st>>>(src, dst, C, n, lo, t, 1, 0, (u64)B, 0, 0, 1, tb);
        else
            ntt_pass_kernel<true, true><<<grid, NTT_THREADS, smem, st>>>(dst, dst, C, n, lo, t, (u64)B, 0, (u64)B, 0, 0, 1, tb);
        launches++;
        hi = lo;
    }
    // fused middle pass
    const int tf = p.bits[p.npass - 1];
    {
        int total = B << tf;
        ntt_setup_coset<<<(total + 255) / 256, 256, 0, st>>>(scl, gam, tb.bytepow, n, ext_bits, tf);
        launches++;
        dim3 grid(1u << (n - tf), ychunks, 1);
        size_t smem = ntt_smem_bytes(tf, true);
        if (ntt_set_smem(ntt_lde_fused_kernel, smem) != cudaSuccess) return -1;
        const bool first = (p.npass == 1);
        ntt_lde_fused_kernel<<<grid, NTT_THREADS, smem, st>>>(first ? src : dst, dst, C, n, ext_bits, tf, first ? 1 : (u64)B, n_inv, scl,
                                                              gam, tb);
        launches++;
    }
    // remaining forward DIT passes, in place on the interleaved layout, one grid.z slice per coset
    int lo = tf;
    for (int i = p.npass - 2; i >= 0; i--) {
        const int t = p.bits[i];
        dim3 grid(1u << (n - t), ychunks, (unsigned)B);
        size_t smem = ntt_smem_bytes(t, false);
        if (ntt_set_smem(ntt_pass_kernel<false, false>, smem) != cudaSuccess) return -1;
        ntt_pass_kernel<false, false><<<grid, NTT_THREADS, smem, st>>>(dst, dst, C, n, lo, t, (u64)B, 0, (u64)B, 0, 0, 1, tb);
        launches++;
        lo += t;
    }
    return launches;
}
