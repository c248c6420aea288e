// Multi-GPU commit group without a collective library (the reference has no counterpart: it is single-process).
//
// Every rank owns a receive buffer (its rows of the extended trace, as one column tile per source rank) and a small MAILBOX; both
// are mapped into every peer (CUDA IPC between processes, plain peer access inside one process).  Everything the ranks tell each
// other travels as peer stores issued by kernels on the rank's own stream:
//   * the column -> row exchange rides on the last butterfly pass of the LDE (ntt.cuh, NttScatter);
//   * ordering is a flag barrier: a one-CTA kernel writes this rank's epoch into every peer's mailbox (release, system scope) and
//     spins until every peer's epoch has arrived in its own (acquire, system scope) -- compute and signal are fused on the stream,
//     no host round trip and no NCCL call on the critical path;
//   * the sub-roots and the opened rows / sibling paths of the queries a rank owns are stored into every peer's mailbox.
// The host side (Node workers, Python processes, or one thread driving several GPUs) only moves two 64-byte handles per rank once.
#pragma once
#include "merkle.cuh"

#define SHARD_MAX_RANKS NTT_MAX_PEERS
// mailbox layout (u64 words)
#define SHARD_FLAGS 0                                   // [SHARD_MAX_RANKS] epoch written by rank r
#define SHARD_ERR (SHARD_FLAGS + SHARD_MAX_RANKS)       // [1] non-zero: a barrier timed out
#define SHARD_SUB (SHARD_ERR + 1 + 15)                  // [4 * SHARD_MAX_RANKS] sub-roots (16-byte aligned)
#define SHARD_TOP (SHARD_SUB + 4 * SHARD_MAX_RANKS)     // [256] tree over the sub-roots (reference layout of a height-world tree)
#define SHARD_STAGE (SHARD_TOP + 256)                   // [stage_words] opened rows + sibling paths, one slot per query

struct ShardPeers {
    u64* mail[SHARD_MAX_RANKS];     // mail[r]: rank r's mailbox as seen from this device
};

GL_D void shard_st_release(u64* p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
GL_D u64 shard_ld_acquire(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Signal + wait.  Everything this stream did before (peer stores of earlier kernels included) is ordered before the signal; whatever
// the peers did before THEIR signal is visible after the wait.  A peer that never arrives trips the timeout instead of hanging the
// GPU: the error word is set (pil2gpu_shard_status reports it) and the kernel returns.
__global__ void shard_barrier_kernel(ShardPeers P, u32 rank, u32 world, u64 epoch, u64 timeout_ns) {
    const u32 t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    shard_st_release(P.mail[t] + SHARD_FLAGS + rank, epoch);
    const u64* mine = P.mail[rank] + SHARD_FLAGS + t;
    u64 t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (shard_ld_acquire(mine) < epoch) {
        __nanosleep(64);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {
            P.mail[rank][SHARD_ERR] = epoch;
            break;
        }
    }
    __threadfence_system();
}

// src4 (this rank's sub-root) -> sub[rank] of every mailbox
__global__ void shard_publish_kernel(ShardPeers P, u32 rank, u32 world, const u64* __restrict__ src4) {
    const u32 peer = threadIdx.x >> 2, w = threadIdx.x & 3;
    if (peer < world) P.mail[peer][SHARD_SUB + 4 * rank + w] = src4[w];
}

// getGroupProof (merklehash_p.js:142-168) of a tree whose leaves are spread over the ranks: one CTA per query; the rank that owns leaf
// idx stores the row, the siblings inside its subtree and the siblings of the top tree (which every rank holds) into slot q of EVERY
// mailbox; the other ranks do nothing.  Slot = width + 4 * (depth_local + depth_top) words.
__global__ void shard_group_proof_kernel(RowTiles t, const u64* __restrict__ nodes, u64 width, u64 rows_local, const u64* __restrict__ idxs,
                                         int depth_local, int depth_top, ShardPeers P, u32 rank, u32 world) {
    const u64 q = blockIdx.x;
    const u64 idx = idxs[q];
    if (idx / rows_local != rank) return;
    u64 local = idx - (u64)rank * rows_local;
    const u64 slot = width + 4 * (u64)(depth_local + depth_top);
    const u64* __restrict__ top = P.mail[rank] + SHARD_TOP;
    for (u32 p = 0; p < world; p++) {
        u64* __restrict__ dst = P.mail[p] + SHARD_STAGE + q * slot;
        for (u64 i = threadIdx.x; i < width; i += blockDim.x) dst[i] = *tiles_ptr(t, local, i);
    }
    if (threadIdx.x < 4) {
        u64 off = 0, n = rows_local * 4, li = local;
        for (int d = 0; d < depth_local; d++) {
            const u64 v = nodes[off + (li ^ 1) * 4 + threadIdx.x];
            for (u32 p = 0; p < world; p++) P.mail[p][SHARD_STAGE + q * slot + width + 4 * d + threadIdx.x] = v;
            const u64 next = ((n - 1) / 8 + 1) * 4;
            li >>= 1;
            off += next * 2;
            n = next;
        }
        off = 0; n = (u64)world * 4; li = rank;
        for (int d = 0; d < depth_top; d++) {
            const u64 v = top[off + (li ^ 1) * 4 + threadIdx.x];
            for (u32 p = 0; p < world; p++) P.mail[p][SHARD_STAGE + q * slot + width + 4 * (depth_local + d) + threadIdx.x] = v;
            const u64 next = ((n - 1) / 8 + 1) * 4;
            li >>= 1;
            off += next * 2;
            n = next;
        }
    }
}
// staging slots -> rows_out[q * width ..], sib_out[(q * depth + level) * 4 ..]
__global__ void shard_unpack_kernel(const u64* __restrict__ stage, u64 width, int depth, u64* __restrict__ rows_out, u64* __restrict__ sib_out) {
    const u64 q = blockIdx.x, slot = width + 4 * (u64)depth;
    const u64* __restrict__ s = stage + q * slot;
    for (u64 i = threadIdx.x; i < width; i += blockDim.x) rows_out[q * width + i] = s[i];
    for (u64 i = threadIdx.x; i < 4 * (u64)depth; i += blockDim.x) sib_out[q * 4 * depth + i] = s[width + i];
}
