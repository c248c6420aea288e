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
#include <cstdlib>
#include "gl.cuh"

#define NTT_W 16        // columns per tile (128 contiguous bytes per row segment)
#define NTT_TMAX 9      // max log2(rows) per tile: 2^9 * 16 * 8 B = 64 KiB (+ twiddles) -> 3 CTAs / SM
#define NTT_THREADS 256
#ifndef NTT_PASS_MIN_CTAS
#define NTT_PASS_MIN_CTAS 5
#endif
#ifndef NTT_FUSED_MIN_CTAS
#define NTT_FUSED_MIN_CTAS 5
#endif

struct NttTables {
    const u64* bytepow;   // [4][256]: W32^(b << (8k)) * 2^64
    const u64* tw_fwd;    // [2^NTT_TMAX]: entry (1 << s) + j = w_{2^(s+1)}^j * 2^64, j < 2^s  (entry 0 unused)
    const u64* tw_inv;    // same with inverse roots
    const u64* pow7;      // [32]: 7^(2^k) * 2^64
};
#define NTT_TABLE_WORDS (1024 + 2 * (1 << NTT_TMAX) + 32)

// Row scatter of the multi-GPU commit: the pass that finishes the LDE stores extended row R of this rank's
// column slab straight into the receive buffer of the rank that hashes row R (peer memory over NVLink / NVSwitch, mapped
// with CUDA IPC) instead of the local buffer -- the column -> row exchange rides on the stores of the last butterfly pass.
//   destination = peer[R >> rl_bits] + tile_off + (R & (2^rl_bits - 1)) * out_C + col_off + column
#define NTT_MAX_PEERS 16
struct NttScatter {
    u64* peer[NTT_MAX_PEERS];   // peer[h]: rank h's receive buffer (device pointer valid on this device)
    int rl_bits;                // log2(rows per rank)
    u64 tile_off;               // word offset of this rank's column tile inside a receive buffer
    u64 out_C;                  // columns of that tile (>= the columns of the slab being transformed)
    u64 col_off;                // first column of the slab inside the tile
};

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
                                             int ext_bits, const NttTables& tb, bool unit_shift = false) {
    const u64* __restrict__ base = INVERSE ? tb.tw_inv : tb.tw_fwd;
    const bool plain = (base_lo == 0) && (coset_r < 0 || (coset_r == 0 && unit_shift));
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
            g = ntt_root_pow(tb.bytepow, E);
            if (!unit_shift) g = gl_mmul(g, tb.pow7[k]);
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

// ---- tile layout --------------------------------------------------------------------------------------------
// A tile holds 2^t rows x NTT_W columns.  In shared memory it is stored as NTT_CP = NTT_W/2 column-pair regions, one per
// warp: region cp holds the 16-byte elements (row k, columns 2cp, 2cp+1) at padded index k + (k >> 3).  Every butterfly of
// a column stays inside one region, so a warp transforms its two columns with __syncwarp only -- no CTA barrier between
// the radix steps -- and every shared-memory access is conflict-free (the pad makes rows 8, 64, ...
// apart land in different bank groups; regions are an odd number of elements long so that the 8 column pairs of one row,
// written by one quarter-warp of cp.async, land in 8 different bank groups too).
#define NTT_CP (NTT_W / 2)
__host__ __device__ __forceinline__ int ntt_pad(int k) { return k + (k >> 3); }
__host__ __device__ __forceinline__ int ntt_region_elems(int t) { return ((1 << t) + ((1 << t) >> 3)) | 1; }

// One "step" handles R consecutive layers in registers; `low` is the lowest k-bit of the step.  Lane l works on column
// (l & 1) of the pair, so a half-warp touches 8 whole 16-byte elements per access: conflict-free 64-bit LDS/STS, and a
// thread keeps only 2^R values live (44-50 registers -> 4-5 CTAs per SM).
template <int R, bool DIF>
__device__ __forceinline__ void ntt_warp_step(ulonglong2* __restrict__ reg, const u64* __restrict__ TW, int t, int low) {
    const int items = 2 << (t - R);
    for (int item = threadIdx.x & 31; item < items; item += 32) {
        const int g = item >> 1;
        const int glow = g & ((1 << low) - 1);
        const int kbase = ((g >> low) << (low + R)) | glow;
        u64* __restrict__ p = reinterpret_cast<u64*>(reg) + (item & 1);
        u64 v[1 << R];
#pragma unroll
        for (int i = 0; i < (1 << R); i++) v[i] = p[ntt_pad(kbase + (i << low)) * 2];
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
        for (int i = 0; i < (1 << R); i++) p[ntt_pad(kbase + (i << low)) * 2] = v[i];
    }
    __syncwarp();
}

// Full transform of the warp's region (2^t points of two columns).  DIF: natural -> bit-reversed (layers from the top),
// canonical in/out; DIT: bit-reversed -> natural (layers from the bottom), any u64 in/out.  Steps are radix-8 when that
// keeps all 32 lanes busy (t >= 7), radix-4 / radix-2 for smaller tiles.
template <bool DIF, int R>
__device__ __forceinline__ void ntt_warp_transform_r(ulonglong2* reg, const u64* TW, int t) {
    if (DIF) {
        int top = t;
        while (top >= R) { ntt_warp_step<R, true>(reg, TW, t, top - R); top -= R; }
        if (R > 2 && top == 2) ntt_warp_step<2, true>(reg, TW, t, 0);
        if (R > 1 && top == 1) ntt_warp_step<1, true>(reg, TW, t, 0);
    } else {
        int low = 0;
        const int rem = t % R;
        if (R > 1 && rem == 1) { ntt_warp_step<1, false>(reg, TW, t, 0); low = 1; }
        if (R > 2 && rem == 2) { ntt_warp_step<2, false>(reg, TW, t, 0); low = 2; }
        while (low < t) { ntt_warp_step<R, false>(reg, TW, t, low); low += R; }
    }
}
// Compile-time chains for the tile sizes the planner actually produces (6..9 bits): with t and `low` constant every
// shared-memory offset and twiddle index folds into an immediate (measured: 40 -> ~31 instructions per butterfly).
template <int T, int TOP, int R>
__device__ __forceinline__ void ntt_dif_chain(ulonglong2* reg, const u64* TW) {
    if constexpr (TOP >= R) {
        ntt_warp_step<R, true>(reg, TW, T, TOP - R);
        ntt_dif_chain<T, TOP - R, R>(reg, TW);
    } else if constexpr (TOP == 2) {
        ntt_warp_step<2, true>(reg, TW, T, 0);
    } else if constexpr (TOP == 1) {
        ntt_warp_step<1, true>(reg, TW, T, 0);
    }
}
template <int T, int LOW, int R>
__device__ __forceinline__ void ntt_dit_chain(ulonglong2* reg, const u64* TW) {
    if constexpr (LOW < T) {
        ntt_warp_step<R, false>(reg, TW, T, LOW);
        ntt_dit_chain<T, LOW + R, R>(reg, TW);
    }
}
template <bool DIF, int T>
__device__ __forceinline__ void ntt_warp_transform_c(ulonglong2* reg, const u64* TW) {
    constexpr int R = (T >= 7) ? 3 : 2;
    if constexpr (DIF) {
        ntt_dif_chain<T, T, R>(reg, TW);
    } else {
        constexpr int rem = T % R;
        if constexpr (rem == 1) ntt_warp_step<1, false>(reg, TW, T, 0);
        if constexpr (rem == 2) ntt_warp_step<2, false>(reg, TW, T, 0);
        ntt_dit_chain<T, rem, R>(reg, TW);
    }
}
template <bool DIF>
__device__ __forceinline__ void ntt_warp_transform(ulonglong2* reg, const u64* TW, int t) {
    switch (t) {
        case 9: ntt_warp_transform_c<DIF, 9>(reg, TW); break;
        case 8: ntt_warp_transform_c<DIF, 8>(reg, TW); break;
        case 7: ntt_warp_transform_c<DIF, 7>(reg, TW); break;
        case 6: ntt_warp_transform_c<DIF, 6>(reg, TW); break;
        default:
            if (t >= 2) ntt_warp_transform_r<DIF, 2>(reg, TW, t);
            else if (t == 1) ntt_warp_transform_r<DIF, 1>(reg, TW, t);
    }
}

// Shared-memory layout: tile (NTT_CP regions) | [tile2] | TW (2^t) | G (16)
static inline size_t ntt_smem_bytes(int t, bool two_tiles) {
    const size_t words = (size_t)NTT_CP * ntt_region_elems(t) * 2 * (two_tiles ? 2 : 1) + ((size_t)1 << t) + 16;
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
    int in_z;          // coset passes: the input rows are interleaved by coset too (0: every coset reads the same rows)
    int unit_shift;    // coset passes: evaluate on w_E^r <w_N> (shift 1) instead of 7 w_E^r <w_N>
};

// Tile load: tile row k comes from buffer row (row0 + k * row_step), or, for the first DIT pass of a natural -> natural
// transform, from the bit-reversed position.  Full, aligned chunks use one 16-byte cp.async per (row, column pair) with a
// running pointer (no 64-bit multiplies in the loop); everything else takes scalar loads.  Callers wait with
// ntt_cp_async_wait + barrier when this returns true.
__device__ __forceinline__ bool ntt_tile_load(ulonglong2* __restrict__ tile, const u64* __restrict__ in, u64 C, u64 c0, int cw, int t, u64 row0,
                                              u64 row_step, int bitrev_n) {
    const int rows = 1 << t, RS = ntt_region_elems(t);
    const bool async_ok = bitrev_n < 0 && cw == NTT_W && (C % 2 == 0) && ((((size_t)in) & 15) == 0);
    if (async_ok) {
        const int cp = threadIdx.x % NTT_CP, k0 = threadIdx.x / NTT_CP;
        const u64* __restrict__ src = in + (row0 + (u64)k0 * row_step) * C + c0 + 2 * cp;
        const u64 step = row_step * C * (NTT_THREADS / NTT_CP);
        ulonglong2* __restrict__ dst = tile + cp * RS;
        for (int k = k0; k < rows; k += NTT_THREADS / NTT_CP, src += step) ntt_cp_async16(reinterpret_cast<u64*>(dst + ntt_pad(k)), src);
    } else {
        const int c = threadIdx.x % NTT_W;
        u64* __restrict__ tw = reinterpret_cast<u64*>(tile) + (c >> 1) * RS * 2 + (c & 1);
        for (int k = threadIdx.x / NTT_W; k < rows; k += NTT_THREADS / NTT_W) {
            u64 row = row0 + (u64)k * row_step;
            if (bitrev_n >= 0) row = (bitrev_n == 0) ? 0 : (u64)(__brevll(row) >> (64 - bitrev_n));   // row0/row_step describe positions here
            tw[ntt_pad(k) * 2] = (c < cw) ? in[row * C + c0 + c] : 0;
        }
    }
    return async_ok;
}

// Tile store to buffer rows (row0 + k * row_step); FIX: 0 = as is, 1 = canonicalise, 2 = multiply by `scale` (Montgomery form).
// SC: rows go to the peer receive buffers described by `sc` (see NttScatter) instead of `out`.
template <int FIX, bool SC>
__device__ __forceinline__ void ntt_tile_store_fix(const ulonglong2* __restrict__ tile, u64* __restrict__ out, u64 C, u64 c0, int cw, int t, u64 scale,
                                                   u64 row0, u64 row_step, const NttScatter& sc) {
    const int rows = 1 << t, RS = ntt_region_elems(t);
    const u64 rl_mask = SC ? (((u64)1 << sc.rl_bits) - 1) : 0;
    if (cw == NTT_W && (C % 2 == 0) && (SC ? (((sc.out_C | sc.col_off) & 1) == 0) : ((((size_t)out) & 15) == 0))) {
        const int cp = threadIdx.x % NTT_CP, k0 = threadIdx.x / NTT_CP;
        u64* __restrict__ dst = out + (row0 + (u64)k0 * row_step) * C + c0 + 2 * cp;
        const u64 step = row_step * C * (NTT_THREADS / NTT_CP);
        const ulonglong2* __restrict__ src = tile + cp * RS;
        for (int k = k0; k < rows; k += NTT_THREADS / NTT_CP, dst += step) {
            ulonglong2 v = src[ntt_pad(k)];
            if (FIX == 2) { v.x = gl_mmul(v.x, scale); v.y = gl_mmul(v.y, scale); }
            if (FIX == 1) { v.x = gl_canon(v.x); v.y = gl_canon(v.y); }
            if (SC) {
                const u64 row = row0 + (u64)k * row_step;
                u64* p = sc.peer[row >> sc.rl_bits] + sc.tile_off + (row & rl_mask) * sc.out_C + sc.col_off + c0 + 2 * cp;
                *reinterpret_cast<ulonglong2*>(p) = v;
            } else {
                *reinterpret_cast<ulonglong2*>(dst) = v;
            }
        }
    } else {
        const int c = threadIdx.x % NTT_W;
        const u64* __restrict__ tw = reinterpret_cast<const u64*>(tile) + (c >> 1) * RS * 2 + (c & 1);
        for (int k = threadIdx.x / NTT_W; k < rows; k += NTT_THREADS / NTT_W) {
            if (c < cw) {
                u64 v = tw[ntt_pad(k) * 2];
                if (FIX == 2) v = gl_mmul(v, scale);
                if (FIX == 1) v = gl_canon(v);
                const u64 row = row0 + (u64)k * row_step;
                if (SC) sc.peer[row >> sc.rl_bits][sc.tile_off + (row & rl_mask) * sc.out_C + sc.col_off + c0 + c] = v;
                else out[row * C + c0 + c] = v;
            }
        }
    }
}
template <bool SC>
__device__ __forceinline__ void ntt_tile_store(const ulonglong2* __restrict__ tile, u64* __restrict__ out, u64 C, u64 c0, int cw, int t, int fix, u64 scale,
                                               u64 row0, u64 row_step, const NttScatter& sc) {
    if (fix == 2) ntt_tile_store_fix<2, SC>(tile, out, C, c0, cw, t, scale, row0, row_step, sc);
    else if (fix == 1) ntt_tile_store_fix<1, SC>(tile, out, C, c0, cw, t, scale, row0, row_step, sc);
    else ntt_tile_store_fix<0, SC>(tile, out, C, c0, cw, t, scale, row0, row_step, sc);
}

// ---- generic pass ------------------------------------------------------------------------------------------
// One pass over bits [lo, lo+t) of a 2^n-point transform of every column.
//   position(k) = (base_hi << (lo+t)) | (k << lo) | base_lo,   tile id = (base_hi << lo) | base_lo
// Work items are (tile id, column chunk) pairs, tile-major: item = tile_id * ychunks + y.  A CTA owns `ipc` consecutive items
// (gridDim.x = ceil(items / ipc), gridDim.z = cosets; default ipc = 1).  With two tile buffers (PIPE) the cp.async loads of item
// i+1 are in flight while item i is transformed and the twiddle table is built once per tile id -- an A/B variant that measured
// slower than one item per CTA at 5 CTAs / SM (see ntt_launch_pass).  Per item: barrier (loads landed, previous store finished)
// -> [prefetch next] -> warp-private butterflies -> barrier -> cooperative store.
struct NttItems {
    u32 ychunks;   // column chunks per tile id
    u32 ipc;       // items per CTA
    u32 total;     // tiles * ychunks
    u32 tiles;     // tile ids; ipc == 1 numbers the items chunk-major (item = y * tiles + tile_id: neighbouring CTAs work on
                   // neighbouring rows of the same column chunk, the order r01 measured fastest), ipc > 1 tile-major
};
__device__ __forceinline__ void ntt_item_decode(const NttItems& it, u32 item, u32& tile_id, u32& y) {
    if (it.ipc == 1) { y = item / it.tiles; tile_id = item - y * it.tiles; }
    else { tile_id = item / it.ychunks; y = item - tile_id * it.ychunks; }
}
template <bool DIF, bool INVERSE, bool SC = false, bool PIPE = true, bool NARROW = false>
__global__ void __launch_bounds__(NTT_THREADS, PIPE ? 3 : NTT_PASS_MIN_CTAS) ntt_pass_kernel(const u64* in, u64* out, NttPass P, NttTables tb,   // in == out for the in-place passes: no __restrict__
                                                               const __grid_constant__ NttScatter sc, NttItems it) {
    extern __shared__ __align__(16) u64 ntt_smem[];
    const int t = P.t, lo = P.lo;
    const int RS = ntt_region_elems(t);
    ulonglong2* tiles[2];
    tiles[0] = reinterpret_cast<ulonglong2*>(ntt_smem);
    tiles[1] = PIPE ? tiles[0] + (size_t)NTT_CP * RS : tiles[0];
    u64* TW = ntt_smem + (size_t)NTT_CP * RS * 2 * (PIPE ? 2 : 1);
    u64* G = TW + ((size_t)1 << t);
    const u64 z = blockIdx.z;
    const u32 item0 = blockIdx.x * it.ipc;
    const u32 item_end = (item0 + it.ipc < it.total) ? item0 + it.ipc : it.total;

    auto issue_load = [&](u32 item, ulonglong2* tile) -> bool {
        u32 tile_id, y;
        ntt_item_decode(it, item, tile_id, y);
        const u32 base_lo = tile_id & ((1u << lo) - 1), base_hi = tile_id >> lo;
        const u64 c0 = (u64)y * NTT_W;
        const int cw = (int)((P.C - c0 < NTT_W) ? (P.C - c0) : NTT_W);
        const u64 pos0 = ((u64)base_hi << (lo + t)) | base_lo;
        // position(k) = pos0 + (k << lo); buffer row = position * mul + z (bit-reversed gather: the helper reverses the
        // position itself, mul is 1 and z is 0 there)
        const bool async = P.bitrev_in ? ntt_tile_load(tile, in, P.C, c0, cw, t, pos0, (u64)1 << lo, P.n)
                                       : ntt_tile_load(tile, in, P.C, c0, cw, t, pos0 * P.in_mul + (P.in_z ? z : 0), P.in_mul << lo, -1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        return async;
    };

    u32 cur_tile = 0xFFFFFFFFu;
    if (item0 < item_end) issue_load(item0, tiles[0]);
    for (u32 item = item0; item < item_end; item++) {
        const int bsel = PIPE ? (int)((item - item0) & 1) : 0;
        ulonglong2* tile = tiles[bsel];
        u32 tile_id, y;
        ntt_item_decode(it, item, tile_id, y);
        const u32 base_lo = tile_id & ((1u << lo) - 1), base_hi = tile_id >> lo;
        if (tile_id != cur_tile) {   // every warp is past the butterflies of the previous item (barrier B below): the table is free
            ntt_build_tw<INVERSE>(TW, G, t, lo, base_lo, P.coset ? (int)z : -1, P.n, P.ext_bits, tb, P.unit_shift != 0);   // ends with a barrier
            cur_tile = tile_id;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // A: this item's rows are in shared memory; the previous store is finished
        if (PIPE && item + 1 < item_end) issue_load(item + 1, tiles[bsel ^ 1]);
        ulonglong2* reg = tile + (threadIdx.x >> 5) * RS;      // this warp's column pair
        // NARROW: buffers narrower than a tile (quotient, LEv: 3-6 columns) -- warps whose column pair is padding skip the butterflies.
        // A separate instantiation: the test costs the wide kernels registers they do not have (48 at 5 CTAs / SM; measured +1.5 ms on cfg3's LDE).
        if (!NARROW || 2 * (u64)(threadIdx.x >> 5) < P.C) {
            if (DIF && P.canon_in) {   // the caller's buffer may hold non-canonical words; DIF butterflies need canonical inputs
                for (int k = threadIdx.x & 31; k < (1 << t); k += 32) {
                    ulonglong2 v = reg[ntt_pad(k)];
                    v.x = gl_canon(v.x); v.y = gl_canon(v.y);
                    reg[ntt_pad(k)] = v;
                }
                __syncwarp();
            }
            ntt_warp_transform<DIF>(reg, TW, t);
        }
        __syncthreads();                                   // B
        const u64 c0 = (u64)y * NTT_W;
        const int cw = (int)((P.C - c0 < NTT_W) ? (P.C - c0) : NTT_W);
        const u64 pos0 = ((u64)base_hi << (lo + t)) | base_lo;
        ntt_tile_store<SC>(tile, out, P.C, c0, cw, t, P.scale ? 2 : ((!DIF && P.canon_out) ? 1 : 0), P.scale, pos0 * P.out_mul + z, P.out_mul << lo, sc);
        if (!PIPE && item + 1 < item_end) {
            __syncthreads();                               // single buffer: the store must finish before the next load overwrites it
            issue_load(item + 1, tile);
        }
    }
}

// ---- fused LDE middle kernel ---------------------------------------------------------------------------------
// Tile = 2^t contiguous transform positions q (lo = 0).  Finishes the INTT (last t DIF layers, inverse roots), scales by
// 1/N, then for every coset r < B runs the first t DIT layers of the size-N forward NTT on 7 w_E^r <w_N> (coset folded into
// the twiddles) and stores to row (B*q + r).  Input row = q*in_mul (src: in_mul = 1; dst: in_mul = B).
template <bool SC>
__global__ void __launch_bounds__(NTT_THREADS, NTT_FUSED_MIN_CTAS) ntt_lde_fused_kernel(const u64* __restrict__ in, u64* __restrict__ out, u64 C, int n,
                                                                    int ext_bits, int t, u64 in_mul, u64 n_inv_mont, int canon_in,
                                                                    int canon_out, NttTables tb, const __grid_constant__ NttScatter sc) {
    extern __shared__ __align__(16) u64 ntt_smem[];
    const int RS = ntt_region_elems(t);
    ulonglong2* tile = reinterpret_cast<ulonglong2*>(ntt_smem);
    ulonglong2* tile2 = tile + (size_t)NTT_CP * RS;
    u64* TW = ntt_smem + (size_t)NTT_CP * RS * 4;
    u64* G = TW + ((size_t)1 << t);
    const int B = 1 << (ext_bits - n);
    const u64 q0 = (u64)blockIdx.x << t;
    const u64 c0 = (u64)blockIdx.y * NTT_W;
    const int cw = (int)((C - c0 < NTT_W) ? (C - c0) : NTT_W);
    const int rows = 1 << t;
    const bool async = ntt_tile_load(tile, in, C, c0, cw, t, q0 * in_mul, in_mul, -1);
    ntt_build_tw<true>(TW, G, t, 0, 0, -1, n, ext_bits, tb);
    if (async) { ntt_cp_async_wait(); __syncthreads(); }
    ulonglong2* reg = tile + (threadIdx.x >> 5) * RS;
    ulonglong2* reg2 = tile2 + (threadIdx.x >> 5) * RS;
    if (canon_in) {
        for (int k = threadIdx.x & 31; k < rows; k += 32) {
            ulonglong2 v = reg[ntt_pad(k)];
            v.x = gl_canon(v.x); v.y = gl_canon(v.y);
            reg[ntt_pad(k)] = v;
        }
        __syncwarp();
    }
    ntt_warp_transform<true>(reg, TW, t);          // coefficients, bit-reversed: position q holds a_{bitrev_n(q)}
    for (int k = threadIdx.x & 31; k < rows; k += 32) {   // * 1/N, once (warp-private region)
        ulonglong2 v = reg[ntt_pad(k)];
        v.x = gl_mmul(v.x, n_inv_mont); v.y = gl_mmul(v.y, n_inv_mont);
        reg[ntt_pad(k)] = v;
    }
    __syncwarp();
    for (int r = 0; r < B; r++) {
        __syncthreads();                                           // every warp is done with the previous twiddle table
        ntt_build_tw<false>(TW, G, t, 0, 0, r, n, ext_bits, tb);   // ends with a barrier
        ulonglong2* work = reg;                                    // the last coset transforms the coefficients in place
        if (r + 1 < B) {
            work = reg2;
            for (int k = threadIdx.x & 31; k < rows; k += 32) reg2[ntt_pad(k)] = reg[ntt_pad(k)];
            __syncwarp();
        }
        ntt_warp_transform<false>(work, TW, t);
        __syncthreads();
        ntt_tile_store<SC>((r + 1 < B) ? tile2 : tile, out, C, c0, cw, t, canon_out ? 1 : 0, 0, (q0 << (ext_bits - n)) + r, (u64)B, sc);
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

// Pass launch.  Default: one item per CTA, single tile buffer, 5 CTAs / SM -- co-resident CTAs hide the load latency.
// Measured on B200 at cfg3 (r02, profiles/r02_lde_pipeline_ab.md): the software-pipelined variant (PIL2GPU_NTT_PIPE=1: two tile
// buffers, cp.async prefetch of item i+1 under the butterflies of item i, 3 CTAs / SM at 80 registers, PIL2GPU_NTT_IPC items
// per CTA) runs the LDE in 137-139 ms against 125 ms, and even the unpipelined item loop (IPC=16) costs 132 ms: many small
// independent CTAs mix the alu- and fmaheavy-heavy phases of the butterfly better than fewer, longer-lived ones.  Both variants stay
// selectable for A/B runs.
#ifndef NTT_ITEMS_PER_CTA
#define NTT_ITEMS_PER_CTA 1
#endif
template <bool DIF, bool INVERSE, bool SC>
static inline int ntt_launch_pass(const u64* in, u64* out, const NttPass& P, unsigned cosets, const NttTables& tb, const NttScatter& sc, cudaStream_t st) {
    const u32 tiles = 1u << (P.n - P.t);
    const u32 ychunks = (u32)((P.C + NTT_W - 1) / NTT_W);
    const u64 total64 = (u64)tiles * ychunks;
    if (total64 > 0xFFFFFFF0ull) return -1;
    NttItems it;
    it.ychunks = ychunks;
    it.tiles = tiles;
    it.total = (u32)total64;
    // whole tile ids per CTA (the twiddle table is per tile id), about NTT_ITEMS_PER_CTA items, but keep >= ~8 waves of CTAs
    static const int env_ipc = getenv("PIL2GPU_NTT_IPC") ? atoi(getenv("PIL2GPU_NTT_IPC")) : 0;       // tuning / A-B knobs
    static const int env_pipe = getenv("PIL2GPU_NTT_PIPE") ? atoi(getenv("PIL2GPU_NTT_PIPE")) : 0;
    const u32 want = env_ipc > 0 ? (u32)env_ipc : NTT_ITEMS_PER_CTA;
    u32 tpc = ychunks >= want ? 1 : want / ychunks;
    while (tpc > 1 && (u64)(tiles / tpc) * cosets < 148u * 3u * 8u) tpc >>= 1;
    it.ipc = tpc * ychunks;
    if (env_ipc > 0 && (u32)env_ipc < it.ipc) it.ipc = (u32)env_ipc;      // below one tile id per CTA: the table is rebuilt per CTA
    const bool pipe = env_pipe && P.t <= 8 && it.ipc >= 2;
    dim3 grid((it.total + it.ipc - 1) / it.ipc, 1, cosets);
    const size_t smem = ntt_smem_bytes(P.t, pipe);
    if (pipe) {
        if (ntt_set_smem(ntt_pass_kernel<DIF, INVERSE, SC, true>, smem) != cudaSuccess) return -1;
        ntt_pass_kernel<DIF, INVERSE, SC, true><<<grid, NTT_THREADS, smem, st>>>(in, out, P, tb, sc, it);
    } else if (P.C < NTT_W && !SC) {
        if (ntt_set_smem(ntt_pass_kernel<DIF, INVERSE, false, false, true>, smem) != cudaSuccess) return -1;
        ntt_pass_kernel<DIF, INVERSE, false, false, true><<<grid, NTT_THREADS, smem, st>>>(in, out, P, tb, sc, it);
    } else {
        if (ntt_set_smem(ntt_pass_kernel<DIF, INVERSE, SC, false>, smem) != cudaSuccess) return -1;
        ntt_pass_kernel<DIF, INVERSE, SC, false><<<grid, NTT_THREADS, smem, st>>>(in, out, P, tb, sc, it);
    }
    return 1;
}

static inline NttPass ntt_pass_defaults(u64 C, int n, int ext_bits) {
    NttPass P;
    P.C = C; P.n = n; P.lo = 0; P.t = 0; P.in_mul = 1; P.out_mul = 1;
    P.bitrev_in = 0; P.canon_in = 0; P.canon_out = 0; P.coset = 0; P.ext_bits = ext_bits; P.scale = 0; P.in_z = 1; P.unit_shift = 0;
    return P;
}
static inline NttScatter ntt_no_scatter() {
    NttScatter sc;
    for (int i = 0; i < NTT_MAX_PEERS; i++) sc.peer[i] = nullptr;
    sc.rl_bits = 0;
    sc.tile_off = 0;
    sc.out_C = 0;
    sc.col_off = 0;
    return sc;
}

// natural -> natural transform of every column; src != dst.  DIT passes: the first one gathers its rows from src in
// bit-reversed order, every later pass runs in place in dst.  Returns the number of kernel launches or -1.
static int ntt_launch_transform(const u64* src, u64* dst, u64 C, int n, bool inverse, const NttTables& tb, cudaStream_t st) {
    NttPlan p = ntt_plan(n, NTT_TMAX);
    const u64 n_inv = glh_to_mont(glh_inv((1ULL << n) % GL_P));
    const NttScatter nosc = ntt_no_scatter();
    int lo = 0;
    int launches = 0;
    for (int i = p.npass - 1; i >= 0; i--) {
        const int t = p.bits[i];
        const bool first = (lo == 0), last = (i == 0);
        NttPass P = ntt_pass_defaults(C, n, n);
        P.lo = lo; P.t = t;
        P.bitrev_in = first ? 1 : 0; P.canon_out = last ? 1 : 0;
        P.scale = (last && inverse) ? n_inv : 0;
        const u64* in = first ? src : dst;
        if ((inverse ? ntt_launch_pass<false, true, false>(in, dst, P, 1, tb, nosc, st) : ntt_launch_pass<false, false, false>(in, dst, P, 1, tb, nosc, st)) < 0)
            return -1;
        launches++;
        lo += t;
    }
    return launches;
}

// LDE src (2^n rows) -> dst (2^ext rows), all in dst after the first pass.  With `sc` the stores of the last kernel go to
// the peer receive buffers instead of dst (multi-GPU row exchange fused into the LDE).  Returns the number of kernel
// launches or -1.
static int ntt_launch_lde(const u64* src, u64* dst, u64 C, int n, int ext_bits, const NttTables& tb, cudaStream_t st,
                          const NttScatter* sc = nullptr) {
    NttPlan p = ntt_plan(n, NTT_TMAX);
    const int B = 1 << (ext_bits - n);
    const u64 n_inv = glh_to_mont(glh_inv((1ULL << n) % GL_P));
    const NttScatter nosc = ntt_no_scatter();
    const unsigned ychunks = (unsigned)((C + NTT_W - 1) / NTT_W);
    int launches = 0;
    // INTT passes (DIF, inverse roots) except the last: src/dst rows q -> dst rows B*q
    int hi = n;
    for (int i = 0; i + 1 < p.npass; i++) {
        const int t = p.bits[i];
        const int lo = hi - t;
        NttPass P = ntt_pass_defaults(C, n, ext_bits);
        P.lo = lo; P.t = t; P.in_mul = (i == 0) ? 1 : (u64)B; P.out_mul = (u64)B;
        P.canon_in = (i == 0) ? 1 : 0;
        if (ntt_launch_pass<true, true, false>(i == 0 ? src : dst, dst, P, 1, tb, nosc, st) < 0) return -1;
        launches++;
        hi = lo;
    }
    // fused middle pass
    const int tf = p.bits[p.npass - 1];
    {
        dim3 grid(1u << (n - tf), ychunks, 1);
        const size_t smem = ntt_smem_bytes(tf, true);
        const bool only = (p.npass == 1);
        if (only && sc) {
            if (ntt_set_smem(ntt_lde_fused_kernel<true>, smem) != cudaSuccess) return -1;
            ntt_lde_fused_kernel<true><<<grid, NTT_THREADS, smem, st>>>(src, dst, C, n, ext_bits, tf, 1, n_inv, 1, 1, tb, *sc);
        } else {
            if (ntt_set_smem(ntt_lde_fused_kernel<false>, smem) != cudaSuccess) return -1;
            ntt_lde_fused_kernel<false><<<grid, NTT_THREADS, smem, st>>>(only ? src : dst, dst, C, n, ext_bits, tf, only ? 1 : (u64)B, n_inv,
                                                                         only ? 1 : 0, only ? 1 : 0, tb, nosc);
        }
        launches++;
    }
    // remaining forward DIT passes, in place on the interleaved layout, one grid.z slice per coset
    int lo = tf;
    for (int i = p.npass - 2; i >= 0; i--) {
        const int t = p.bits[i];
        NttPass P = ntt_pass_defaults(C, n, ext_bits);
        P.lo = lo; P.t = t; P.in_mul = (u64)B; P.out_mul = (u64)B;
        P.canon_out = (i == 0) ? 1 : 0; P.coset = 1;
        if (((i == 0 && sc) ? ntt_launch_pass<false, false, true>(dst, dst, P, (unsigned)B, tb, *sc, st)
                            : ntt_launch_pass<false, false, false>(dst, dst, P, (unsigned)B, tb, nosc, st)) < 0) return -1;
        launches++;
        lo += t;
    }
    return launches;
}

// INTT of every column WITHOUT the 1/2^n scale, natural -> bit-reversed: dst row q holds 2^n * a_{bitrev_n(q)}.  DIF passes,
// the first one reads src, the others run in place in dst (src == dst is allowed).  Returns launches or -1.
static int ntt_launch_intt_bitrev(const u64* src, u64* dst, u64 C, int n, const NttTables& tb, cudaStream_t st) {
    NttPlan p = ntt_plan(n, NTT_TMAX);
    const NttScatter nosc = ntt_no_scatter();
    int launches = 0, hi = n;
    for (int i = 0; i < p.npass; i++) {
        const int t = p.bits[i];
        const int lo = hi - t;
        NttPass P = ntt_pass_defaults(C, n, n);
        P.lo = lo; P.t = t; P.canon_in = (i == 0) ? 1 : 0;
        if (ntt_launch_pass<true, true, false>(i == 0 ? src : dst, dst, P, 1, tb, nosc, st) < 0) return -1;
        launches++;
        hi = lo;
    }
    return launches;
}

// Evaluate polynomials given by their coefficients in bit-reversed row order (coef row q = a_{bitrev_n(q)}, 2^n rows, any
// u64 representatives) on the 2^ext_bits points  s * w_E^j  (s = 7, or 1 with unit_shift), natural order, canonical out:
// B = 2^(ext_bits - n) coset NTTs of size 2^n (DIT), the first pass reads coef, the others run in place in dst.
static int ntt_launch_coset_eval(const u64* coef, u64* dst, u64 C, int n, int ext_bits, bool unit_shift, const NttTables& tb, cudaStream_t st) {
    NttPlan p = ntt_plan(n, NTT_TMAX);
    const int B = 1 << (ext_bits - n);
    const NttScatter nosc = ntt_no_scatter();
    int launches = 0, lo = 0;
    for (int i = p.npass - 1; i >= 0; i--) {
        const int t = p.bits[i];
        const bool first = (lo == 0);
        NttPass P = ntt_pass_defaults(C, n, ext_bits);
        P.lo = lo; P.t = t; P.in_mul = first ? 1 : (u64)B; P.out_mul = (u64)B; P.in_z = first ? 0 : 1;
        P.canon_out = (i == 0) ? 1 : 0; P.coset = 1; P.unit_shift = unit_shift ? 1 : 0;
        if (ntt_launch_pass<false, false, false>(first ? coef : dst, dst, P, (unsigned)B, tb, nosc, st) < 0) return -1;
        launches++;
        lo += t;
    }
    return launches;
}
