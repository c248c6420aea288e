/*
 * pil2gpu -- C ABI of the B200-native commit-phase library (libpil2gpu.so).
 *
 * Drop-in boundary for the three JavaScript modules of pil2-stark-js that make up the STARK commit phase.  Every
 * entry point names the reference interface it replaces (paths relative to the reference repo root); the N-API
 * addon (napi/pil2gpu_addon.cc), the JS shims (js/ *.js) and the Python mirror (pil2_stark_js_b200/) all bind
 * exactly these symbols.
 *
 * Conventions
 *   - All field data are little-endian u64 Goldilocks elements (p = 2^64 - 2^32 + 1); outputs are always canonical
 *     (< p).  Inputs should be canonical (the reference requires it: src/helpers/f3g.js:61-66); values >= p are
 *     interpreted mod p.
 *   - Buffers are row-major: buff[row * nPols + col]  (pilcom BigBuffer / BigUint64Array layout).
 *   - Every function returns 0 on success and a negative PIL2GPU_E_* code on failure; pil2gpu_last_error() returns
 *     a thread-local message.  Nothing throws or aborts across the boundary.
 *   - Pointers without a _dev suffix are HOST pointers; *_dev functions take DEVICE pointers (on ctx's device) and
 *     only enqueue work on ctx's stream (call pil2gpu_sync, or sync the stream you supplied, before reading).
 *   - Host-pointer functions are synchronous.  A ctx is bound to one GPU and one stream; use one ctx per thread.
 *   - There is no CPU fallback: without a CUDA device pil2gpu_create fails with PIL2GPU_E_CUDA.
 */
#ifndef PIL2GPU_H
#define PIL2GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIL2GPU_OK 0
#define PIL2GPU_E_INVALID (-1)   /* bad argument (sizes, aliasing, null pointers) */
#define PIL2GPU_E_RANGE (-2)     /* "Out of range" (merklehash_p.js:143) */
#define PIL2GPU_E_CUDA (-3)      /* CUDA runtime error (no device, launch failure, ...) */
#define PIL2GPU_E_NOMEM (-4)     /* device or pinned-host allocation failed */
#define PIL2GPU_E_UNSUPPORTED (-5)

typedef struct pil2gpu_ctx pil2gpu_ctx;
typedef struct pil2gpu_tree pil2gpu_tree;   /* device-resident {elements, nodes, width, height} */

/* ---- lifetime ------------------------------------------------------------------------------------------- */
/* stream: a cudaStream_t to enqueue on (e.g. torch.cuda.current_stream().cuda_stream; pass cudaStreamLegacy = 0x1 for
 * the default stream, whose handle is 0) or NULL to let the ctx create its own non-blocking stream.  Replaces the per-call workerpool.pool()/terminate() of fft_p.js:121,175
 * and merklehash_p.js:53,105 (no threads are left alive between calls; the ctx only owns twiddle tables). */
int pil2gpu_create(int device, void* stream, pil2gpu_ctx** out);
void pil2gpu_destroy(pil2gpu_ctx* ctx);
const char* pil2gpu_last_error(void);
const char* pil2gpu_version(void);
int pil2gpu_sync(pil2gpu_ctx* ctx);
/* The host-pointer entry points keep a grow-only device workspace in the ctx between calls (allocating and freeing tens
 * of GiB per call costs more than the kernels); this returns it to the driver. */
int pil2gpu_release_workspace(pil2gpu_ctx* ctx);
/* Number of kernels of this library launched through ctx since creation (bench.py "gpu_launches"). */
uint64_t pil2gpu_launch_count(const pil2gpu_ctx* ctx);

/* ---- memory helpers (so JS / Python hosts need no CUDA binding of their own) ---------------------------- */
int pil2gpu_dev_alloc(pil2gpu_ctx* ctx, size_t bytes, void** dptr);
int pil2gpu_dev_free(pil2gpu_ctx* ctx, void* dptr);
int pil2gpu_host_alloc(size_t bytes, void** hptr);   /* pinned host memory for BigBuffer pages */
int pil2gpu_host_free(void* hptr);
int pil2gpu_h2d(pil2gpu_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);   /* async on ctx stream */
int pil2gpu_d2h(pil2gpu_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);   /* async on ctx stream */

/* ---- NTT: src/helpers/fft/fft_p.js ------------------------------------------------------------------------ */
/* fft(buffSrc,nPols,nBits,buffDst) / ifft(...)  fft_p.js:178-184.  dst = per-column NTT (inverse != 0: INTT) of
 * the 2^nBits x nPols buffer, natural order in and out.  src and dst must not overlap. */
int pil2gpu_ntt(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t* dst, uint64_t nPols, uint32_t nBits, int inverse);
int pil2gpu_ntt_dev(pil2gpu_ctx* ctx, const uint64_t* src_dev, uint64_t* dst_dev, uint64_t nPols, uint32_t nBits, int inverse);
/* interpolate(buffSrc,nPols,nBits,buffDst,nBitsExt)  fft_p.js:187-297: dst[j*nPols+c] = P_c(7 * w_ext^j) where P_c
 * interpolates column c of src on <w_n>.  dst holds 2^nBitsExt rows and is fully overwritten (its prior contents are
 * ignored, unlike the reference which needs a zeroed dst: fft_p.js:265).  No scratch buffer is used. */
int pil2gpu_lde(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t* dst, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt);
int pil2gpu_lde_dev(pil2gpu_ctx* ctx, const uint64_t* src_dev, uint64_t* dst_dev, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt);
/* BigBuffer twins (pilcom BigBuffer = list of BigUint64Array pages; the reference passes BigBuffers of any size to fft / ifft /
 * interpolate, fft_p.js:178-297): page p holds page_words[p] u64, pages are concatenated in order and need not hold whole
 * rows.  Every host-pointer entry point of this library accepts pinned memory (pil2gpu_host_alloc: copied by DMA directly)
 * and ordinary pageable memory (an ordinary BigUint64Array: staged through a ring of pinned slots by a few copy threads,
 * PIL2GPU_COPY_THREADS) -- detected per page with cudaPointerGetAttributes. */
int pil2gpu_ntt_paged(pil2gpu_ctx* ctx, const uint64_t* const* src_pages, const uint64_t* src_page_words, uint32_t n_src_pages,
                      uint64_t* const* dst_pages, const uint64_t* dst_page_words, uint32_t n_dst_pages, uint64_t nPols, uint32_t nBits,
                      int inverse);
int pil2gpu_lde_paged(pil2gpu_ctx* ctx, const uint64_t* const* src_pages, const uint64_t* src_page_words, uint32_t n_src_pages,
                      uint64_t* const* dst_pages, const uint64_t* dst_page_words, uint32_t n_dst_pages, uint64_t nPols,
                      uint32_t nBits, uint32_t nBitsExt);

/* ---- multi-GPU row exchange fused into the LDE (no reference counterpart: the reference is single-process) ------- */
/* CUDA IPC plumbing so that one process per GPU can map its peers' receive buffers (cudaIpcGetMemHandle /
 * cudaIpcOpenMemHandle).  dptr must be the base of an allocation made with pil2gpu_dev_alloc. */
int pil2gpu_ipc_export(pil2gpu_ctx* ctx, const void* dptr, uint8_t handle_out[64]);
int pil2gpu_ipc_open(pil2gpu_ctx* ctx, const uint8_t handle[64], void** dptr_out);
int pil2gpu_ipc_close(pil2gpu_ctx* ctx, void* dptr);
/* interpolate (fft_p.js:187-297) of a column slab (2^nBits x nPols) of this rank whose LAST butterfly pass stores extended
 * row R directly into the receive buffer of the rank that hashes it: peer_recv_dev[R / rows_local] (a HOST array of
 * n_ranks device pointers valid on this device, own buffer included), at word offset
 *     rank * rows_local * tile_cols + (R % rows_local) * tile_cols + col_off + column,   rows_local = 2^nBitsExt / n_ranks
 * -- the tile layout pil2gpu_merkelize_tiled_dev hashes in place (tile_cols = columns this rank contributes in total,
 * 0 = nPols; col_off = first column of this slab inside the tile).  dst_dev (2^nBitsExt x nPols) is the in-place workspace
 * of the earlier passes.  The caller orders the ranks (a barrier before the receive buffers are read or rewritten). */
int pil2gpu_lde_scatter_dev(pil2gpu_ctx* ctx, const uint64_t* src_dev, uint64_t* dst_dev, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt,
                            uint64_t* const* peer_recv_dev, uint32_t n_ranks, uint32_t rank, uint64_t tile_cols, uint64_t col_off);
/* Same with the slab in HOST memory (row pitch src_pitch_cols words, pinned for full speed): uploads and transforms it
 * in sub-slabs of 32 columns so that the PCIe upload overlaps the LDE and the peer stores.  Asynchronous: returns once
 * the work is enqueued on the ctx stream. */
int pil2gpu_lde_scatter(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t src_pitch_cols, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt,
                        uint64_t* const* peer_recv_dev, uint32_t n_ranks, uint32_t rank);

/* ---- multi-GPU commit group: peer-mapped buffers, flag barriers, no collective library (no reference counterpart) --------------
 * The choreography of SURVEY 8(e) behind the C ABI, so that ANY host can drive it -- Node workers (one process per GPU, the handles
 * travel over the fork channel), one Node / Python thread driving several GPUs (pil2gpu_shard_connect_local), or the torch.distributed
 * launcher bench.py uses.  rank g owns columns [g*nPols/world, (g+1)*nPols/world) of the trace and, after the exchange, rows
 * [g*E/world, (g+1)*E/world) of the extended trace (E = 2^nBitsExt) as `world` column tiles in its receive buffer.  Every call below
 * only ENQUEUES work on the context's stream (no host synchronisation), and every rank must issue the same sequence of calls: the
 * ranks order themselves with flag barriers -- a one-CTA kernel stores this rank's epoch into every peer's mailbox (release.sys) and
 * spins on its own (acquire.sys) -- so compute and signal share the stream.  A peer that never arrives trips a timeout
 * (PIL2GPU_SHARD_TIMEOUT_MS, default 20000) reported by pil2gpu_shard_status instead of hanging the GPU. */
typedef struct pil2gpu_shard pil2gpu_shard;
/* recv_words >= (nPols/world) << nBitsExt of the largest commit; stage_words >= n_idx * (nPols + 4*nBitsExt) of the largest
 * pil2gpu_shard_open_dev (0: no queries).  world: power of two <= 16. */
int pil2gpu_shard_create(pil2gpu_ctx* ctx, uint32_t rank, uint32_t world, uint64_t recv_words, uint64_t stage_words, pil2gpu_shard** out);
int pil2gpu_shard_destroy(pil2gpu_shard* sh);
/* One process per GPU: export this rank's 2 x 64-byte CUDA IPC handles (receive buffer, mailbox); gather the pairs of all ranks in rank
 * order by any means; connect. */
int pil2gpu_shard_handles(pil2gpu_shard* sh, uint8_t handles_out[128]);
int pil2gpu_shard_connect(pil2gpu_shard* sh, const uint8_t* handles /* world x 128 bytes */, uint32_t n_ranks);
/* One process, several GPUs (or several contexts of one GPU): wire group[r] (rank r of n_ranks) to each other directly, enabling peer
 * access between the devices where needed. */
int pil2gpu_shard_connect_local(pil2gpu_shard* const* group, uint32_t n_ranks);
uint64_t* pil2gpu_shard_recv_dev(pil2gpu_shard* sh);               /* this rank's rows: tile t at t * (E/world) * (nPols/world) words */
uint64_t* const* pil2gpu_shard_peer_recv(pil2gpu_shard* sh);       /* host array of `world` device pointers for pil2gpu_lde_scatter[_dev] */
/* After a commit / hash: the `world` sub-roots (4 words each, rank order) and the tree over them (pil2gpu_merkle_nnodes(world) words,
 * reference layout of a height-`world` tree; for world == 1 the root is sub-root 0), both in this rank's mailbox. */
const uint64_t* pil2gpu_shard_sub_roots_dev(pil2gpu_shard* sh);
const uint64_t* pil2gpu_shard_top_nodes_dev(pil2gpu_shard* sh);
int pil2gpu_shard_barrier(pil2gpu_shard* sh);                      /* enqueue one flag barrier (all ranks) */
int pil2gpu_shard_status(pil2gpu_shard* sh);                       /* synchronises the stream; PIL2GPU_E_CUDA if a barrier timed out */
/* extendAndMerkelize (stark_gen_helpers.js:388-412) over the group: barrier -> LDE of this rank's 2^nBits x nPols/world column slab
 * whose last pass stores every row into its owner's receive buffer (work_dev: (nPols/world) << nBitsExt words of scratch) -> barrier
 * -> hashing + subtree of the local rows in place (nodes_dev: pil2gpu_merkle_nnodes(E/world) words, reference layout of a
 * height-E/world tree) -> sub-roots stored into every mailbox -> barrier -> top log2(world) levels.  root_out_dev (device, 4 words,
 * may be NULL) receives the root of the whole tree on every rank; it equals the single-GPU root. */
int pil2gpu_shard_commit_dev(pil2gpu_shard* sh, const uint64_t* src_slab_dev, uint64_t* work_dev, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt,
                             int split, uint64_t* nodes_dev, uint64_t* root_out_dev);
/* The hashing half alone, for rows delivered by other means (pil2gpu_lde_scatter from host slabs + pil2gpu_shard_barrier). */
int pil2gpu_shard_hash_dev(pil2gpu_shard* sh, uint64_t nPols, uint32_t nBitsExt, int split, uint64_t* nodes_dev, uint64_t* root_out_dev);
/* getGroupProof (merklehash_p.js:142-168) for n_idx global leaf indices (device array, identical on every rank): the owner of each leaf
 * stores its row, the siblings inside its subtree and those of the top tree into every rank's mailbox; after the barrier every rank
 * holds rows_out_dev[q*nPols ..] and siblings_out_dev[(q*nBitsExt + level)*4 ..] of the single-GPU tree. */
int pil2gpu_shard_open_dev(pil2gpu_shard* sh, const uint64_t* nodes_dev, uint64_t nPols, uint32_t nBitsExt, const uint64_t* idxs_dev, uint32_t n_idx,
                           uint64_t* rows_out_dev, uint64_t* siblings_out_dev);

/* ---- quotient polynomial: computeQStark, src/stark/stark_gen_helpers.js:168-208 ----------------------------------- */
/* q_ext: 2^nBitsExt x qDim evaluations of Q on 7<w_ext>.  cmq_ext (2^nBitsExt x qDim*qDeg): column p*qDim + k holds the
 * evaluations on 7<w_ext> of the p-th degree-<2^nBits chunk of Q (ifft :177, shift-split :179-190, fft :192).  qDeg must
 * not exceed the blowup factor.  The host-pointer form also merkelizes cmq_ext (:198): cmq_ext_out / nodes_out may be NULL. */
int pil2gpu_compute_q_dev(pil2gpu_ctx* ctx, const uint64_t* q_ext_dev, uint64_t qDim, uint64_t qDeg, uint32_t nBits, uint32_t nBitsExt,
                          uint64_t* cmq_ext_dev);
int pil2gpu_compute_q(pil2gpu_ctx* ctx, const uint64_t* q_ext, uint64_t qDim, uint64_t qDeg, uint32_t nBits, uint32_t nBitsExt, int split,
                      uint64_t* cmq_ext_out, uint64_t* nodes_out, uint64_t root_out[4]);
/* BigBuffer twin: ctx.q_ext and ctx.cm<Q>_ext are BigBuffers (stark_gen_helpers.js:174-176); cmq_pages may be NULL / 0 pages. */
int pil2gpu_compute_q_paged(pil2gpu_ctx* ctx, const uint64_t* const* q_pages, const uint64_t* q_page_words, uint32_t n_q_pages, uint64_t qDim,
                            uint64_t qDeg, uint32_t nBits, uint32_t nBitsExt, int split, uint64_t* const* cmq_pages,
                            const uint64_t* cmq_page_words, uint32_t n_cmq_pages, uint64_t* nodes_out, uint64_t root_out[4]);

/* ---- evaluations at the challenge point: computeEvalsStark, src/stark/stark_gen_helpers.js:210-273 ------------------ */
/* LEv vector of one opening point (:216-231): lev_dev (2^nBits x 3 words) = F3 ifft of the powers of
 * xi_challenge * w_n^opening / 7.  xi_challenge is an F3 element (3 words). */
int pil2gpu_compute_lev_dev(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], int32_t opening, uint32_t nBits, uint64_t* lev_dev);
/* One entry of pilInfo.evMap inside one extended buffer (:236-249). */
typedef struct pil2gpu_eval_desc {
    uint64_t offset;   /* first column of the polynomial inside a row (p.offset) */
    uint32_t dim;      /* 1 or 3 (p.dim) */
    uint32_t lev;      /* index of the opening point (openingPoints.findIndex(...), :262) */
} pil2gpu_eval_desc;
/* The evaluation loop (:250-264) over one device-resident extended buffer buf_dev (2^nBitsExt rows of `size` words):
 * evals_out[3e ..] = sum_k buf[(k << (nBitsExt - nBits)) * size + offset_e (..+2)] * LEv_{lev_e}[k].  lev_dev holds n_lev
 * vectors back to back (n_lev x 2^nBits x 3).  evals_out is a HOST array of n_evals x 3 words; the call synchronises. */
int pil2gpu_compute_evals_dev(pil2gpu_ctx* ctx, const uint64_t* buf_dev, uint64_t size, uint32_t nBits, uint32_t nBitsExt,
                              const pil2gpu_eval_desc* desc, uint32_t n_evals, const uint64_t* lev_dev, uint32_t n_lev, uint64_t* evals_out);
/* Host-buffer form of the whole of computeEvalsStark's arithmetic for one extended buffer `buf` (2^nBitsExt rows of `size`
 * words, host): builds the LEv vectors of all opening points and uploads only the 2^nBits rows the loop reads (:252-259). */
int pil2gpu_compute_evals(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], const int32_t* openings, uint32_t n_open, uint32_t nBits,
                          uint32_t nBitsExt, const uint64_t* buf, uint64_t size, const pil2gpu_eval_desc* desc, uint32_t n_evals,
                          uint64_t* evals_out);
/* xDivXSubXi_ext of computeFRIStark (:289-323): out_dev[3 * (k * n_open + i) ..] = x_k / (x_k - xi_challenge * w_n^openings[i]),
 * x_k = 7 * w_ext^k, k < 2^nBitsExt.  (A point where x_k equals the shifted challenge makes the reference throw
 * "Division by zero"; here the affected words are unspecified.) */
int pil2gpu_x_div_x_sub_xi_dev(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], const int32_t* openings, uint32_t n_open, uint32_t nBits,
                               uint32_t nBitsExt, uint64_t* out_dev);
int pil2gpu_x_div_x_sub_xi(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], const int32_t* openings, uint32_t n_open, uint32_t nBits,
                           uint32_t nBitsExt, uint64_t* out);

/* ---- constraint expressions over a domain: calculateExps / callCalculateExps, src/prover/prover_helpers.js:23-110 ------------ */
/* The reference evaluates a straight-line program of {op, dest, src} records (add / sub / mul / copy on base-field or F3 operands,
 * f3g.js:47-104) at every row i of the domain; operands are rows (i + next) mod N of the stage buffers, uniform values (numbers,
 * publics, challenges, evals, subproof values), x_i and temporaries (getRef / setRef, prover_helpers.js:112-265).  Here the host
 * compiles the program into fixed records of 16 u32 words:
 *     [ opcode, n_src, 0, 0,  dest[3],  src0[3],  src1[3],  src2[3] ]        opcode: 0 add, 1 sub, 2 mul, 3 copy (1 source), 4 muladd (3)
 * an operand is  [ kind | dim << 8 | buffer << 16,  a,  row_offset ]  with dim 1 (base field) or 3 (F3) and
 *     kind 0 tmp      a = slot < 64 (slots are assigned by liveness on the host; a temporary is 3 words per thread)
 *     kind 1 const    a = index into consts (n_consts entries of 3 words; a base-field value is (v, 0, 0))          [source only]
 *     kind 2 buffer   a = first column inside a row of bufs[buffer]; the row is (i + row_offset) mod 2^domain_bits
 *     kind 3 x        x_i = w^i of the domain, times the coset shift 7 when x_shift != 0 (ctx.x_n / ctx.x_ext)      [source only]
 * (Zi_ext, xDivXSubXi_ext, q_ext, f_ext and the stage buffers are all "buffer" operands).  A program must not read, at a non-zero
 * row offset, a column it also writes (the reference runs the rows in order; here they run concurrently).  Asynchronous on the
 * ctx stream.  The Python mirror (pil2_stark_js_b200/prover_helpers.py) compiles the reference's code objects into this form. */
typedef struct pil2gpu_expr_buffer {
    uint64_t* ptr_dev;
    uint64_t row_words;
} pil2gpu_expr_buffer;
int pil2gpu_calculate_exps_dev(pil2gpu_ctx* ctx, const uint32_t* ops, uint32_t n_ops, const uint64_t* consts, uint32_t n_consts,
                               const pil2gpu_expr_buffer* bufs, uint32_t n_bufs, uint32_t domain_bits, int x_shift);
/* Host-buffer form (the reference's ctx arrays, prover_helpers.js:33-76): buffers with `read` != 0 are uploaded, those with `written` != 0
 * come back after the program has run; every buffer holds 2^domain_bits rows of row_words words.  Synchronous. */
typedef struct pil2gpu_expr_host_buffer {
    uint64_t* ptr;
    uint64_t row_words;
    int32_t read;
    int32_t written;
} pil2gpu_expr_host_buffer;
int pil2gpu_calculate_exps(pil2gpu_ctx* ctx, const uint32_t* ops, uint32_t n_ops, const uint64_t* consts, uint32_t n_consts,
                           const pil2gpu_expr_host_buffer* bufs, uint32_t n_bufs, uint32_t domain_bits, int x_shift);
/* The records are compiled at run time into a straight-line kernel (NVRTC, once per device and program -- what compileCode,
 * prover_helpers.js:87-110, does with `new Function`); the interpreter remains as the fallback when libnvrtc is absent
 * (PIL2GPU_EXPR=interp forces it, =jit turns a missing compiler into PIL2GPU_E_UNSUPPORTED; PIL2GPU_JIT_CACHE=<dir> keeps the compiled
 * kernels across processes).  pil2gpu_expr_jit_check returns the generated
 * source (up to cap - 1 characters) and 0 when NVRTC compiles it for sm_100a; it needs no device. */
int pil2gpu_expr_jit_check(const uint32_t* ops, uint32_t n_ops, const uint64_t* row_words, uint32_t n_bufs, uint32_t domain_bits, int x_shift,
                           char* source_out, uint64_t cap);

/* ---- FRI polynomial: computeFRIStark after the xDivXSubXi table, src/stark/stark_gen_helpers.js:325-334 ------------- */
/* One entry of pilInfo.evMap, in evMap order (the Horner order of friPolinomial.js:26-40 depends on it): the polynomial's
 * extended buffer on the device (2^nBitsExt rows of `size` words), its first column, dim 1 or 3, and the opening (prime). */
typedef struct pil2gpu_fri_term {
    const uint64_t* buf_dev;
    uint64_t size;
    uint64_t offset;
    uint32_t dim;
    int32_t prime;
} pil2gpu_fri_term;
/* f_dev[3k ..] = friExp at row k (src/pil_info/helpers/polynomials/friPolinomial.js:26-56, what callCalculateExps evaluates into
 * f_ext): per opening o the Horner combination in vf2 of (p_i(x_k) - evals[i]) over the evMap entries opened at o, times
 * xDivXSubXi_ext[k][index of o], combined across the openings by Horner in vf1 in the order JavaScript enumerates the keys of
 * friExps.  evals: HOST array n_terms x 3 (ctx.evals); xdiv_dev: 2^nBitsExt x n_open x 3 on the device; at most 4 distinct
 * openings, rows of at most 4096 columns.  Asynchronous on the ctx stream. */
int pil2gpu_fri_pol_dev(pil2gpu_ctx* ctx, const pil2gpu_fri_term* terms, uint32_t n_terms, const uint64_t* evals, const int32_t* openings,
                        uint32_t n_open, const uint64_t* xdiv_dev, const uint64_t vf1[3], const uint64_t vf2[3], uint32_t nBitsExt,
                        uint64_t* f_dev);

/* Host-buffer form of all of computeFRIStark's arithmetic (:289-334): terms[i].buf_dev are HOST pointers here (each distinct
 * buffer is uploaded once), the xDivXSubXi table is built on the device from xi_challenge, f_out (2^nBitsExt x 3) and, if
 * xdiv_out != NULL, the table (2^nBitsExt x n_open x 3) come back.  Synchronous. */
int pil2gpu_fri_pol(pil2gpu_ctx* ctx, const pil2gpu_fri_term* terms, uint32_t n_terms, const uint64_t* evals, const int32_t* openings,
                    uint32_t n_open, const uint64_t xi_challenge[3], const uint64_t vf1[3], const uint64_t vf2[3], uint32_t nBits,
                    uint32_t nBitsExt, uint64_t* f_out, uint64_t* xdiv_out);

/* ---- hashing: src/helpers/hash/poseidon/poseidon.js, linearhash/ *.js -------------------------------------- */
/* poseidon(inputs[8], capacity[4], nOuts) poseidon.js:57-108: full 12-word permutation of in12 = inputs || capacity. */
int pil2gpu_poseidon(pil2gpu_ctx* ctx, const uint64_t in12[12], uint64_t out12[12]);
/* LinearHash.hash linearhash.js:8-42 (split == 0) / LinearHashGPU.hash linearhash_gpu.js:31-67 (split != 0). */
int pil2gpu_linear_hash(pil2gpu_ctx* ctx, const uint64_t* vals, uint64_t width, int split, uint64_t out4[4]);

/* ---- Merkle: src/helpers/hash/merklehash/merklehash_p.js --------------------------------------------------- */
/* _getNNodes(height*4) merklehash_p.js:28-42: number of u64 words in tree.nodes. */
uint64_t pil2gpu_merkle_nnodes(uint64_t height);
/* Number of sibling levels of a proof (length of getGroupProof(...)[1]). */
uint32_t pil2gpu_merkle_depth(uint64_t height);
/* merkelize(buff,width,height) merklehash_p.js:44-133 with the ctor flag splitLinearHash (:20-26).  Writes
 * pil2gpu_merkle_nnodes(height) words to nodes (reference layout, root = last 4 words). */
int pil2gpu_merkelize(pil2gpu_ctx* ctx, const uint64_t* elems, uint64_t width, uint64_t height, int split, uint64_t* nodes);
int pil2gpu_merkelize_dev(pil2gpu_ctx* ctx, const uint64_t* elems_dev, uint64_t width, uint64_t height, int split, uint64_t* nodes_dev);
/* Multi-GPU building blocks.  After the column->row all-to-all a rank holds its rows as n_tiles column
 * tiles (one per source rank): element (row, c) at tiles + (c / tile_cols) * tile_stride + row * tile_cols + c % tile_cols.
 * merkelize_tiled hashes them in place (width = n_tiles * tile_cols; tile_cols % 8 == 0); tree_from_digests builds the
 * levels above `height` leaf digests already stored at nodes[0 .. 4*height) (the gathered sub-roots). */
int pil2gpu_merkelize_tiled_dev(pil2gpu_ctx* ctx, const uint64_t* tiles_dev, uint32_t n_tiles, uint64_t tile_cols, uint64_t tile_stride,
                                uint64_t height, int split, uint64_t* nodes_dev);
int pil2gpu_merkle_tree_from_digests_dev(pil2gpu_ctx* ctx, uint64_t* nodes_dev, uint64_t height);
int pil2gpu_merkelize_paged(pil2gpu_ctx* ctx, const uint64_t* const* elem_pages, const uint64_t* page_words, uint32_t n_pages,
                            uint64_t width, uint64_t height, int split, uint64_t* nodes);

/* ---- device-resident commit (extendAndMerkelize, src/stark/stark_gen_helpers.js:388-412; buildConstTree,
 *      src/stark/stark_buildConstTree.js:17-33): interpolate + merkelize without the LDE ever leaving HBM ---- */
int pil2gpu_commit(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split,
                   pil2gpu_tree** tree_out, uint64_t root_out[4]);
int pil2gpu_commit_dev(pil2gpu_ctx* ctx, const uint64_t* src_dev, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split,
                       pil2gpu_tree** tree_out, uint64_t root_out[4]);
/* extendAndMerkelize (src/stark/stark_gen_helpers.js:388-412 = interpolate :397 + merkelize :400) with HOST buffers in
 * and out: uploads src, extends and hashes on the device, and writes the extended buffer (dst_out, 2^nBitsExt x nPols,
 * may be NULL) and tree.nodes (nodes_out, may be NULL) back; the download of dst overlaps the hashing.  root_out = root. */
int pil2gpu_extend_and_merkelize(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split,
                                 uint64_t* dst_out, uint64_t* nodes_out, uint64_t root_out[4]);

/* BigBuffer twin of extendAndMerkelize: cm<stage>_n (16 GiB at 2^23 x 256) and cm<stage>_ext are BigBuffers.  Pages that hold
 * whole rows (always the case for power-of-two nPols) go through the column-slab pipeline -- upload | LDE + hashing | download
 * overlapped, one strided copy per (slab, page); other page cuts upload everything, extend, and download under the hashing.
 * dst_pages may be NULL / 0 pages (root and nodes only). */
int pil2gpu_extend_and_merkelize_paged(pil2gpu_ctx* ctx, const uint64_t* const* src_pages, const uint64_t* src_page_words, uint32_t n_src_pages,
                                       uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split, uint64_t* const* dst_pages,
                                       const uint64_t* dst_page_words, uint32_t n_dst_pages, uint64_t* nodes_out, uint64_t root_out[4]);

int pil2gpu_tree_wrap_tiled_dev(pil2gpu_ctx* ctx, const uint64_t* tiles_dev, uint32_t n_tiles, uint64_t tile_cols, uint64_t tile_stride,
                                const uint64_t* nodes_dev, uint64_t height, pil2gpu_tree** tree_out);

/* ---- bench / test utilities (not part of the reference surface) ------------------------------------------------ */
/* dst_dev[i] = splitmix64(seed ^ (first_index + i)) mod p : the synthetic trace generator of SURVEY 8(d). */
int pil2gpu_synth_dev(pil2gpu_ctx* ctx, uint64_t* dst_dev, uint64_t n_words, uint64_t seed, uint64_t first_index);
/* Column slab of the same generator: dst[r*cols + c] = splitmix64(seed ^ (r*row_stride + col0 + c)) mod p. */
int pil2gpu_synth2d_dev(pil2gpu_ctx* ctx, uint64_t* dst_dev, uint64_t rows, uint64_t cols, uint64_t row_stride, uint64_t col0, uint64_t seed);
/* Integer-pipe roofline denominators measured live: standalone Goldilocks mulmod/s and IMAD.WIDE.U32/s on this GPU. */
int pil2gpu_bench_int_pipes(pil2gpu_ctx* ctx, double* mulmod_per_s, double* imad_wide_per_s);

/* Wrap / build trees over existing data.  tree_from_host uploads elements and merkelizes them on the device. */
int pil2gpu_tree_from_host(pil2gpu_ctx* ctx, const uint64_t* elems, uint64_t width, uint64_t height, int split, pil2gpu_tree** tree_out);
/* A tree whose nodes are already known -- readFromFile (merklehash_p.js:249-278), the const tree of a setup -- goes to the
 * device without re-hashing: allocate, then fill elements (which = 0) and nodes (which = 1) piecewise from host chunks. */
int pil2gpu_tree_alloc(pil2gpu_ctx* ctx, uint64_t width, uint64_t height, pil2gpu_tree** tree_out);
int pil2gpu_tree_fill(pil2gpu_ctx* ctx, pil2gpu_tree* t, int which, uint64_t offset_words, const uint64_t* src, uint64_t n_words);
int pil2gpu_tree_width(const pil2gpu_tree* t, uint64_t* width, uint64_t* height);
const uint64_t* pil2gpu_tree_elements_dev(const pil2gpu_tree* t);
const uint64_t* pil2gpu_tree_nodes_dev(const pil2gpu_tree* t);
/* root(tree) merklehash_p.js:224 */
int pil2gpu_tree_root(pil2gpu_ctx* ctx, const pil2gpu_tree* t, uint64_t root_out[4]);
/* getGroupProof(tree, idx) merklehash_p.js:142-168 for n_idx leaves at once: rows_out[q*width ..] = row values,
 * siblings_out[(q*depth + level)*4 ..] = sibling at each level.  Any idx >= height -> PIL2GPU_E_RANGE ("Out of range"). */
int pil2gpu_tree_group_proofs(pil2gpu_ctx* ctx, const pil2gpu_tree* t, const uint64_t* idxs, uint32_t n_idx, uint64_t* rows_out,
                              uint64_t* siblings_out);
/* Same with device-resident indices and outputs (asynchronous, no range check: an index >= height marks a slot that
 * belongs to another rank and is zero-filled, so that the per-rank results of a sharded tree combine with one sum-reduce). */
int pil2gpu_tree_group_proofs_dev(pil2gpu_ctx* ctx, const pil2gpu_tree* t, const uint64_t* idxs_dev, uint32_t n_idx, uint64_t* rows_out_dev,
                                  uint64_t* siblings_out_dev);
/* Copy back to the host layout of the reference tree object {elements, nodes}; either pointer may be NULL. */
int pil2gpu_tree_download(pil2gpu_ctx* ctx, const pil2gpu_tree* t, uint64_t* elems_out, uint64_t* nodes_out);
void pil2gpu_tree_free(pil2gpu_ctx* ctx, pil2gpu_tree* t);

/* ---- FRI: src/stark/fri.js --------------------------------------------------------------------------------- */
/* One FRI.fold(step > 0, pol, challenge) fri.js:22-81.  pol: 2^prevBits F3 values (3 words each); pol_out: 2^curBits.
 * step0Bits = steps[0].nBits (fixes the coset shift, fri.js:31-36).  nextBits >= 0: also produce the transposed rows
 * (getTransposedBuffer fri.js:187-202; rows_out: 2^nextBits x 3*2^(curBits-nextBits) words) and their Merkle nodes
 * (nodes_out: pil2gpu_merkle_nnodes(2^nextBits) words); nextBits < 0 is the last step (rows_out/nodes_out ignored).
 * Step 0 (identity fold, fri.js:48-49, followed by the first layer commit :63-71) is prevBits == curBits.
 * The _dev form accepts nodes_out_dev == NULL with nextBits >= 0: rows only, the caller hashes them (sharded layer trees). */
int pil2gpu_fri_fold(pil2gpu_ctx* ctx, const uint64_t* pol, uint32_t prevBits, uint32_t curBits, int32_t nextBits, uint32_t step0Bits,
                     const uint64_t challenge[3], int split, uint64_t* pol_out, uint64_t* rows_out, uint64_t* nodes_out);
int pil2gpu_fri_fold_dev(pil2gpu_ctx* ctx, const uint64_t* pol_dev, uint32_t prevBits, uint32_t curBits, int32_t nextBits,
                         uint32_t step0Bits, const uint64_t challenge[3], int split, uint64_t* pol_out_dev, uint64_t* rows_out_dev,
                         uint64_t* nodes_out_dev);
/* One rank's share of a fold in a sharded FRI chain (no reference counterpart: the reference is single-process): computes rows
 * [row0, row0 + n_rows) of the next layer's transposed buffer (n_rows == 0: all) at their absolute positions in rows_out_dev, and
 * pol_out_dev[g] for the same outputs when pol_out_dev != NULL.  in_layout 0: in_dev is the polynomial in fri.js order;
 * in_layout 1: in_dev is the previous layer's transposed buffer, whose row g holds the 2^(prevBits-curBits) inputs of output g
 * contiguously, so a chain that all-gathers each layer's rows never needs the polynomial order again.  No hashing: the caller
 * hashes its rows.  n_rows must be a multiple of 32 (or the whole layer). */
int pil2gpu_fri_fold_range_dev(pil2gpu_ctx* ctx, const uint64_t* in_dev, int in_layout, uint32_t prevBits, uint32_t curBits, int32_t nextBits,
                               uint32_t step0Bits, const uint64_t challenge[3], uint64_t row0, uint64_t n_rows, uint64_t* pol_out_dev,
                               uint64_t* rows_out_dev);
/* Paged twin (a 2^27-point first layer is 3 GiB, more than one BigUint64Array page): pol / pol_out / rows_out as page lists;
 * rows_pages may be NULL / 0 pages. */
int pil2gpu_fri_fold_paged(pil2gpu_ctx* ctx, const uint64_t* const* pol_pages, const uint64_t* pol_page_words, uint32_t n_pol_pages,
                           uint32_t prevBits, uint32_t curBits, int32_t nextBits, uint32_t step0Bits, const uint64_t challenge[3], int split,
                           uint64_t* const* out_pages, const uint64_t* out_page_words, uint32_t n_out_pages, uint64_t* const* rows_pages,
                           const uint64_t* rows_page_words, uint32_t n_rows_pages, uint64_t* nodes_out);
/* Wrap device buffers produced by the *_dev FRI calls into a tree handle (not owned: tree_free leaves them alone). */
int pil2gpu_tree_wrap_dev(pil2gpu_ctx* ctx, const uint64_t* elems_dev, const uint64_t* nodes_dev, uint64_t width, uint64_t height,
                          pil2gpu_tree** tree_out);

#ifdef __cplusplus
}
#endif
#endif /* PIL2GPU_H */
