// Poseidon-GL linear hash, Merkle tree and group-proof gather for sm_100a.
//
// Replaces (semantics, not structure):
//   linear hash        src/helpers/hash/linearhash/linearhash.js:8-42, linearhash_gpu.js:31-67,
//                      width<=4 passthrough src/helpers/hash/merklehash/merklehash_worker.js:42-49
//   merkelize          src/helpers/hash/merklehash/merklehash_p.js:44-133 (node layout _getNNodes :28-42)
//   getGroupProof      src/helpers/hash/merklehash/merklehash_p.js:142-168
// One row (or one batch of a row in split mode) per thread; the tree is reduced three levels per launch (private depth-3 subtrees per thread) with the
// last <= MERKLE_TAIL pairs finished inside a single CTA.  `nodes` uses the reference layout exactly: level l starts
// right after level l-1, every level is padded with a zero node to an even node count, the root is the last 4
// words; a tree of height 1 is L(row) followed by four zero words (reference quirk, root = 0).
#pragma once
#include "poseidon.cuh"
#include "poseidon_tc.cuh"

#define MERKLE_THREADS 128
#ifndef MERKLE_MIN_CTAS
#define MERKLE_MIN_CTAS 5      // occupancy hint of the hashing kernels: 96 registers with the FP64 partial rounds (measured best of 4/5/6)
#endif
#define MERKLE_FULL_WAVE (1u << 17)   // threads that fill 148 SMs x 640 resident threads
#define MERKLE_TAIL 512    // pairs handled by the single-CTA tail kernel (512 threads keeps 128 regs/thread)

// _getNNodes(height*4) of merklehash_p.js:28-42, in words.
static inline u64 merkle_nnodes_words(u64 height) {
    u64 n = height * 4;
    u64 next = ((n - 1) / 8 + 1) * 4;
    u64 acc = next * 2;
    while (n > 4) {
        n = next;
        next = ((n - 1) / 8 + 1) * 4;
        if (n > 4) acc += next * 2; else acc += 4;
    }
    return acc;
}
static inline int merkle_depth(u64 height) {
    int d = 0;
    u64 n = height * 4;
    while (n > 4) { n = ((n - 1) / 8 + 1) * 4; d++; }
    return d;
}

// Sponge over `w` words at `v` (stride 1): st = H(chunk, st) per 8 words, last chunk zero padded; w <= 4 is copied.
// Works for global or shared pointers.  Output canonical.
GL_D void merkle_sponge(const u64* __restrict__ v, u64 w, u64 out[4]) {
    if (w <= 4) {
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = (u64)i < w ? v[i] : 0;
        return;
    }
    u64 x[12];
#pragma unroll
    for (int i = 8; i < 12; i++) x[i] = 0;
    for (u64 off = 0; off < w; off += 8) {
        if (off + 8 <= w) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = gl_to_mont(v[off + i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (off + i < w) ? gl_to_mont(v[off + i]) : 0;
        }
        poseidon_permute_mont(x);      // the capacity words stay in Montgomery form between absorptions
#pragma unroll
        for (int i = 0; i < 4; i++) x[8 + i] = x[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl_from_mont(x[8 + i]);
}

// Column-tiled row storage: element (row, c) lives at base + (c / tile_cols) * tile_stride + row * tile_cols + c % tile_cols.
// A plain row-major buffer is the one-tile case (tile_cols = width).  Tiled buffers are what the multi-GPU all-to-all
// delivers (one tile per source rank); hashing them in place saves a repacking pass.
struct RowTiles {
    const u64* base;
    u64 tile_cols;     // columns per tile (multiple of 8 unless there is a single tile)
    u64 tile_stride;   // words between consecutive tiles
};
GL_D const u64* tiles_ptr(const RowTiles& t, u64 row, u64 c) {
    const u64 tile = c / t.tile_cols;
    return t.base + tile * t.tile_stride + row * t.tile_cols + (c - tile * t.tile_cols);
}

// Sponge over columns [c0, c0 + w) of one row of a tiled buffer (every 8-word chunk lies inside one tile).
GL_D void merkle_sponge_tiled(const RowTiles& t, u64 row, u64 c0, u64 w, u64 out[4]) {
    if (w <= 4) {
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = (u64)i < w ? *tiles_ptr(t, row, c0 + i) : 0;
        return;
    }
    u64 x[12];
#pragma unroll
    for (int i = 8; i < 12; i++) x[i] = 0;
    for (u64 off = 0; off < w; off += 8) {
        const u64* v = tiles_ptr(t, row, c0 + off);
        if (off + 8 <= w) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = gl_to_mont(v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (off + i < w) ? gl_to_mont(v[i]) : 0;
        }
        poseidon_permute_mont(x);
#pragma unroll
        for (int i = 0; i < 4; i++) x[8 + i] = x[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl_from_mont(x[8 + i]);
}

// Standard linear hash of every row: nodes[4*row ..] = L(row).
__global__ void __launch_bounds__(MERKLE_THREADS, MERKLE_MIN_CTAS) merkle_leaf_kernel(RowTiles t, u64 width, u64 height, u64* __restrict__ nodes) {
    const u64 row = (u64)blockIdx.x * MERKLE_THREADS + threadIdx.x;
    if (row >= height) return;
    u64 d[4];
    merkle_sponge_tiled(t, row, 0, width, d);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(nodes + 4 * row);
    o[0] = make_ulonglong2(d[0], d[1]);
    o[1] = make_ulonglong2(d[2], d[3]);
}

// The same leaf hash with the partial rounds of every permutation in tensor-core form (poseidon_tc.cuh): persistent CTAs of
// MERKLE_TC_THREADS threads (one per SM: the fragment tables, 66 KB, and a 288-byte row per thread live in shared memory), each warp
// hashes 32 rows in lock step; rows past the end are clamped and not stored.  width > 4 only.
#define MERKLE_TC_THREADS 512
#ifndef MERKLE_TC_DEFAULT
#define MERKLE_TC_DEFAULT 0         // PIL2GPU_LEAF_TC=1 / =0 in the environment overrides
#endif
#ifndef MERKLE_TC_MIN_ROWS
#define MERKLE_TC_MIN_ROWS 4096     // below this the table copy per CTA is not worth it
#endif
__global__ void __launch_bounds__(MERKLE_TC_THREADS, 1) merkle_leaf_tc_kernel(RowTiles t, u64 width, u64 height, u64* __restrict__ nodes) {
    extern __shared__ __align__(16) unsigned char ptc_smem[];
    poseidon_tc_setup(ptc_smem, MERKLE_TC_THREADS, threadIdx.x);
    const u64* tab = reinterpret_cast<const u64*>(ptc_smem);
    const int lane = threadIdx.x & 31;
    unsigned char* wrows = ptc_smem + PTC_TABLE_BYTES + (threadIdx.x >> 5) * PTC_WARP_BYTES;
    const u64 ntiles = (height + MERKLE_TC_THREADS - 1) / MERKLE_TC_THREADS;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 row = tile * MERKLE_TC_THREADS + threadIdx.x;
        const u64 r = row < height ? row : height - 1;
        u64 x[12];
#pragma unroll
        for (int i = 8; i < 12; i++) x[i] = 0;
        for (u64 off = 0; off < width; off += 8) {
            const u64* v = tiles_ptr(t, r, off);
            if (off + 8 <= width) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = gl_to_mont(v[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = (off + i < width) ? gl_to_mont(v[i]) : 0;
            }
            poseidon_permute_mont_tc(x, wrows, tab, lane);
#pragma unroll
            for (int i = 0; i < 4; i++) x[8 + i] = x[i];
        }
        if (row < height) {
            ulonglong2* o = reinterpret_cast<ulonglong2*>(nodes + 4 * row);
            o[0] = make_ulonglong2(gl_from_mont(x[8]), gl_from_mont(x[9]));
            o[1] = make_ulonglong2(gl_from_mont(x[10]), gl_from_mont(x[11]));
        }
    }
}
// Launch helper: returns false when the tensor-core kernel does not apply (the caller uses merkle_leaf_kernel).
// Measured on B200 (profiles/r02_poseidon_tc.md): 107 ms against 97 ms for the FP64-resident kernel at 2^22 x 256 -- the tensor-core form
// has 27 % fewer issue slots but only 4 warps per scheduler fit next to its tables, so it stays opt-in (PIL2GPU_LEAF_TC=1).
static bool merkle_launch_leaf_tc(RowTiles t, u64 width, u64 height, u64* nodes, cudaStream_t st) {
    if (width <= 4 || height < MERKLE_TC_MIN_ROWS) return false;
    static const int on = [] { const char* env = getenv("PIL2GPU_LEAF_TC"); return env ? (env[0] == '1') : (MERKLE_TC_DEFAULT != 0); }();
    if (!on) return false;
    int dev = 0, sms = 0;                     // per device, on every launch: a process may drive several GPUs
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0 ||
        cudaFuncSetAttribute(merkle_leaf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POSEIDON_TC_SMEM(MERKLE_TC_THREADS)) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    const u64 ntiles = (height + MERKLE_TC_THREADS - 1) / MERKLE_TC_THREADS;
    const unsigned grid = (unsigned)(ntiles < (u64)sms ? ntiles : (u64)sms);
    merkle_leaf_tc_kernel<<<grid, MERKLE_TC_THREADS, POSEIDON_TC_SMEM(MERKLE_TC_THREADS), st>>>(t, width, height, nodes);
    return true;
}

// Incremental standard linear hash for the column-slab pipeline (pil2gpu.cu: pipelined extend-and-merkelize): absorbs
// `cols` more columns of every row (a row-major slab [height][cols], cols % 8 == 0 except for the last slab) into the
// running sponge state.  state[row*4 ..] holds the capacity words in Montgomery form between slabs; the last slab
// writes the canonical digest to nodes[row*4 ..].  Equivalent to merkle_sponge over the concatenated slabs (total > 4).
__global__ void __launch_bounds__(MERKLE_THREADS, MERKLE_MIN_CTAS) merkle_absorb_kernel(const u64* __restrict__ slab, u64 cols, u64 height,
                                                                       u64* __restrict__ state, int first, int last, u64* __restrict__ nodes) {
    const u64 row = (u64)blockIdx.x * MERKLE_THREADS + threadIdx.x;
    if (row >= height) return;
    u64 x[12];
    if (first) {
#pragma unroll
        for (int i = 8; i < 12; i++) x[i] = 0;
    } else {
        const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(state + 4 * row);
        const ulonglong2 a = sp[0], b = sp[1];
        x[8] = a.x; x[9] = a.y; x[10] = b.x; x[11] = b.y;
    }
    const u64* __restrict__ v = slab + row * cols;
    for (u64 off = 0; off < cols; off += 8) {
        if (off + 8 <= cols) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = gl_to_mont(v[off + i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (off + i < cols) ? gl_to_mont(v[off + i]) : 0;
        }
        poseidon_permute_mont(x);
#pragma unroll
        for (int i = 0; i < 4; i++) x[8 + i] = x[i];
    }
    if (last) {
        ulonglong2* o = reinterpret_cast<ulonglong2*>(nodes + 4 * row);
        o[0] = make_ulonglong2(gl_from_mont(x[8]), gl_from_mont(x[9]));
        o[1] = make_ulonglong2(gl_from_mont(x[10]), gl_from_mont(x[11]));
    } else {
        ulonglong2* o = reinterpret_cast<ulonglong2*>(state + 4 * row);
        o[0] = make_ulonglong2(x[8], x[9]);
        o[1] = make_ulonglong2(x[10], x[11]);
    }
}

// Split linear hash, stage 1: one thread per (row, batch): digests[(row*nb + b)*4 ..] = L(row[b*batch .. ]).
__global__ void __launch_bounds__(MERKLE_THREADS, MERKLE_MIN_CTAS) merkle_batch_kernel(RowTiles t, u64 width, u64 height, u64 batch, u64 nb,
                                                                      u64* __restrict__ digests) {
    const u64 id = (u64)blockIdx.x * MERKLE_THREADS + threadIdx.x;
    if (id >= height * nb) return;
    const u64 row = id / nb, b = id % nb;
    const u64 off = b * batch;
    const u64 sz = (width - off < batch) ? (width - off) : batch;
    u64 d[4];
    merkle_sponge_tiled(t, row, off, sz, d);
#pragma unroll
    for (int i = 0; i < 4; i++) digests[id * 4 + i] = d[i];
}

// One tree level: out[i] = H(in[2i] || in[2i+1], cap = 0), i < pairs (merkelizeLevel, glwasm.js:1220-1254).
GL_D void merkle_pair(const u64* __restrict__ in, u64* __restrict__ out, u64 i) {
    u64 x[12];
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + 8 * i);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ulonglong2 v = p[k];
        x[2 * k] = gl_to_mont(v.x);
        x[2 * k + 1] = gl_to_mont(v.y);
    }
#pragma unroll
    for (int k = 8; k < 12; k++) x[k] = 0;
    poseidon_permute_mont(x);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(out + 4 * i);
    o[0] = make_ulonglong2(gl_from_mont(x[0]), gl_from_mont(x[1]));
    o[1] = make_ulonglong2(gl_from_mont(x[2]), gl_from_mont(x[3]));
}
// All remaining levels inside one CTA (pairs <= MERKLE_TAIL at entry).  nodes + p_in is the current level.
__global__ void __launch_bounds__(MERKLE_TAIL) merkle_tail_kernel(u64* __restrict__ nodes, u64 p_in, u64 n64) {
    u64 next = ((n64 - 1) / 8 + 1) * 4;
    u64 p_out = p_in + next * 2;
    while (n64 > 4) {
        const u64 pairs = next / 4;
        if (threadIdx.x < pairs) merkle_pair(nodes + p_in, nodes + p_out, threadIdx.x);
        __syncthreads();
        n64 = next;
        next = ((n64 - 1) / 8 + 1) * 4;
        p_in = p_out;
        p_out = p_in + next * 2;
    }
}

// Zero the padding node of every odd level (and the height-1 quirk node).
__global__ void merkle_pad_kernel(u64* __restrict__ nodes, u64 height) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u64 p_in = 0, n64 = height * 4;
    u64 next = ((n64 - 1) / 8 + 1) * 4;
    if (n64 <= 4) {   // height 1: nodes = L(row) || 0^4
        for (int i = 0; i < 4; i++) nodes[4 + i] = 0;
        return;
    }
    while (n64 > 4) {
        if (next * 2 != n64) for (u64 i = n64; i < next * 2; i++) nodes[p_in + i] = 0;
        p_in += next * 2;
        n64 = next;
        next = ((n64 - 1) / 8 + 1) * 4;
    }
}

// Up to three tree levels per launch (merkelize, merklehash_p.js:87-132): thread i owns the 2^LV input nodes [2^LV i, 2^LV (i+1)) and
// walks its private subtree depth-first -- 4 + 2 + 1 permutations for LV = 3 -- so every lane stays busy, nothing crosses threads, and
// each level is still written to its place in the reference layout.  A node that does not exist counts as the zero padding node (the
// stored pads of odd levels are zeroed up front by merkle_pad_kernel; power-of-two heights have none).
GL_D void merkle_hash2(const u64 a[4], const u64 b[4], u64 out[4]) {
    u64 x[12];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = gl_to_mont(a[k]);
        x[4 + k] = gl_to_mont(b[k]);
        x[8 + k] = 0;
    }
    poseidon_permute_mont(x);
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = gl_from_mont(x[k]);
}
GL_D void merkle_load4(const u64* __restrict__ p, u64 v[4]) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
    const ulonglong2 a = q[0], b = q[1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
GL_D void merkle_store4(u64* __restrict__ p, const u64 v[4]) {
    ulonglong2* q = reinterpret_cast<ulonglong2*>(p);
    q[0] = make_ulonglong2(v[0], v[1]);
    q[1] = make_ulonglong2(v[2], v[3]);
}
// Thread i walks the depth-LV subtree over the input nodes [2^LV i, 2^LV (i+1)) in post-order -- leaf pair, leaf pair, their parent, ...
// -- as ONE rolled loop of 2^LV - 1 steps with a single copy of the permutation (seven inlined copies would be 250 KB of code):
// a step either hashes the next pair of input nodes (level 1) or the pending left sibling with the node just produced.
template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, MERKLE_MIN_CTAS) merkle_levels_kernel(u64* __restrict__ nodes, u64 p_in, u64 n_in) {
    const u64 i = (u64)blockIdx.x * MERKLE_THREADS + threadIdx.x;
    const u64 first = i << LV;
    if (first >= n_in) return;
    u64 off[LV + 1];                       // word offset of level l above the input (reference layout: even-padded levels)
    off[0] = p_in;
    {
        u64 n = n_in;
#pragma unroll
        for (int l = 1; l <= LV; l++) { off[l] = off[l - 1] + 4 * (n + (n & 1)); n = (n + 1) / 2; }
    }
    u64 left[LV][4];                       // left[l]: finished level-l node waiting for its right sibling (l = 1..LV-1)
    u64 cur[4] = {0, 0, 0, 0};
    int k = 0, pending = 0;                // k: input pairs consumed; pending: level of the waiting left sibling whose parent is next, 0 = none
#pragma unroll 1
    for (int step = 0; step < (1 << LV) - 1; step++) {
        u64 a[4], b[4];
        int lvl;
        u64 idx;                           // the node produced by this step: level lvl, index idx (global)
        if (pending == 0) {
            lvl = 1;
            idx = (first >> 1) + k;
            const u64 n0 = 2 * idx;        // first + 2k
            if (n0 < n_in) {
                merkle_load4(nodes + p_in + 4 * n0, a);
                merkle_load4(nodes + p_in + 4 * (n0 + 1), b);      // n0 + 1 == n_in (odd level): the stored zero pad
            }
            k++;
        } else {
            lvl = pending + 1;
            idx = (first >> lvl) + ((k - 1) >> pending);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                u64 v = left[1 % LV][c];
#pragma unroll
                for (int l = 2; l < LV; l++) v = (pending == l) ? left[l][c] : v;
                a[c] = v;
                b[c] = cur[c];
            }
        }
        if ((idx << lvl) < n_in) {         // the node exists iff its first input node does; a missing node counts as the zero pad
            merkle_hash2(a, b, cur);
            merkle_store4(nodes + off[(lvl < LV) ? lvl : LV] + 4 * idx, cur);
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) cur[c] = 0;
        }
        // odd index within the subtree: the left sibling is waiting, its parent comes next; even: park it and go back to the input level
        const int j = (int)((k - 1) >> (lvl - 1));
        if ((j & 1) && lvl < LV) {
            pending = lvl;
        } else {
#pragma unroll
            for (int l = 1; l < LV; l++)
                if (lvl == l) {
#pragma unroll
                    for (int c = 0; c < 4; c++) left[l][c] = cur[c];
                }
            pending = 0;
        }
    }
}

// Reduce leaf digests already stored at nodes[0 .. 4*height) to the root: three levels per launch while the level above still has more
// than MERKLE_TAIL pairs, then the single-CTA tail.  Returns launches or -1.
static int merkle_launch_tree(u64* nodes, u64 height, cudaStream_t st) {
    int launches = 0;
    if ((height & (height - 1)) != 0 || height == 1) {        // power-of-two heights > 1 have no padding nodes
        merkle_pad_kernel<<<1, 32, 0, st>>>(nodes, height);
        launches++;
    }
    u64 p_in = 0, n = height;                                // n = nodes of the current level
    while (n > 1) {
        if ((n + 1) / 2 <= MERKLE_TAIL) {
            merkle_tail_kernel<<<1, MERKLE_TAIL, 0, st>>>(nodes, p_in, n * 4);
            launches++;
            break;
        }
        // levels this launch: as many as leave a full wave of threads (a thread walks its subtree sequentially, so on a level that
        // no longer fills the GPU one launch per level is faster: measured +0.4 ms on the 2^24-leaf tree with 3 levels everywhere)
        int lv = 1;
        for (u64 m = (n + 1) / 2; lv < 3 && (m + 1) / 2 > MERKLE_TAIL && (n >> (lv + 1)) >= MERKLE_FULL_WAVE; lv++) m = (m + 1) / 2;
        const u64 threads = (n + (1ull << lv) - 1) >> lv;
        const unsigned blocks = (unsigned)((threads + MERKLE_THREADS - 1) / MERKLE_THREADS);
        if (lv == 3) merkle_levels_kernel<3><<<blocks, MERKLE_THREADS, 0, st>>>(nodes, p_in, n);
        else if (lv == 2) merkle_levels_kernel<2><<<blocks, MERKLE_THREADS, 0, st>>>(nodes, p_in, n);
        else merkle_levels_kernel<1><<<blocks, MERKLE_THREADS, 0, st>>>(nodes, p_in, n);
        launches++;
        for (int l = 0; l < lv; l++) { p_in += 4 * (n + (n & 1)); n = (n + 1) / 2; }
    }
    return launches;
}

// Full merkelize of device-resident rows.  `scratch` (split mode only) holds height*nb*4 words.
static inline u64 merkle_split_batch(u64 width) { u64 b = (width + 3) / 4; return b < 8 ? 8 : b; }   // linearhash_gpu.js:42-44
static inline u64 merkle_split_scratch_words(u64 width, u64 height) {
    if (width <= 4) return 0;
    u64 batch = merkle_split_batch(width);
    return height * ((width + batch - 1) / batch) * 4;
}
static int merkle_launch(RowTiles t, u64 width, u64 height, int split, u64* nodes, u64* scratch, cudaStream_t st) {
    int launches = 0;
    if (height == 0) return 0;
    const unsigned blocks = (unsigned)((height + MERKLE_THREADS - 1) / MERKLE_THREADS);
    if (!split || width <= 4) {
        if (!merkle_launch_leaf_tc(t, width, height, nodes, st)) merkle_leaf_kernel<<<blocks, MERKLE_THREADS, 0, st>>>(t, width, height, nodes);
        launches++;
    } else {
        const u64 batch = merkle_split_batch(width);
        const u64 nb = (width + batch - 1) / batch;
        const u64 total = height * nb;
        merkle_batch_kernel<<<(unsigned)((total + MERKLE_THREADS - 1) / MERKLE_THREADS), MERKLE_THREADS, 0, st>>>(t, width, height, batch, nb, scratch);
        RowTiles d = {scratch, nb * 4, 0};
        merkle_leaf_kernel<<<blocks, MERKLE_THREADS, 0, st>>>(d, nb * 4, height, nodes);
        launches += 2;
    }
    int tl = merkle_launch_tree(nodes, height, st);
    return tl < 0 ? -1 : launches + tl;
}

// Group proofs for a batch of leaf indices: row values + one 4-word sibling per level (merklehash_p.js:142-168).
// One CTA per query; rows_out[q*width ..], sib_out[q*depth*4 ..].
__global__ void merkle_group_proof_kernel(RowTiles t, const u64* __restrict__ nodes, u64 width, u64 height,
                                          const u64* __restrict__ idxs, int depth, u64* __restrict__ rows_out, u64* __restrict__ sib_out) {
    const u64 q = blockIdx.x;
    u64 idx = idxs[q];
    if (idx >= height) {   // "not mine" marker of the multi-GPU gather (host entry points reject out-of-range indices): zero fill
        for (u64 i = threadIdx.x; i < width; i += blockDim.x) rows_out[q * width + i] = 0;
        for (u64 i = threadIdx.x; i < (u64)depth * 4; i += blockDim.x) sib_out[q * depth * 4 + i] = 0;
        return;
    }
    for (u64 i = threadIdx.x; i < width; i += blockDim.x) rows_out[q * width + i] = *tiles_ptr(t, idx, i);
    if (threadIdx.x < 4) {
        u64 off = 0, n = height * 4;
        for (int d = 0; d < depth; d++) {
            sib_out[(q * depth + d) * 4 + threadIdx.x] = nodes[off + (idx ^ 1) * 4 + threadIdx.x];
            const u64 next = ((n - 1) / 8 + 1) * 4;
            idx >>= 1;
            off += next * 2;
            n = next;
        }
    }
}

// Single permutation / single row hash (test hooks and the host-side transcript).
__global__ void poseidon_single_kernel(const u64* __restrict__ in12, u64* __restrict__ out12) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u64 x[12];
    for (int i = 0; i < 12; i++) x[i] = in12[i];
    poseidon_permute(x);
    for (int i = 0; i < 12; i++) out12[i] = x[i];
}
