// libpil2gpu.so -- C ABI over the sm_100a kernels (see include/pil2gpu.h for the contract and the reference
// interfaces each entry point replaces).  Unity build: the kernels live in the .cuh files included below.
// There is deliberately no CPU implementation behind any entry point: every failure to reach the GPU is an error.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pil2gpu.h"
#include "fri.cuh"
#include "gl.cuh"
#include "merkle.cuh"
#include "ntt.cuh"
#include "qpath.cuh"
#include "evals.cuh"
#include "shard.cuh"
#include "fripol.cuh"
#include "expr.cuh"
#include "expr_jit.cuh"

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? PIL2GPU_E_NOMEM : PIL2GPU_E_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                         \
    } while (0)

// Inside the enqueue loops of the multi-stream pipelines an early `return` would skip the drain of the copy streams (which
// still target the workspace and the caller's host buffers): record the failure in `rc` and leave the loop instead.
#define CU_BREAK(call)                                                                                        \
    {                                                                                                         \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            rc = fail(e__ == cudaErrorMemoryAllocation ? PIL2GPU_E_NOMEM : PIL2GPU_E_CUDA, "%s: %s (%s:%d)", #call, \
                      cudaGetErrorString(e__), __FILE__, __LINE__);                                           \
            break;                                                                                            \
        }                                                                                                     \
    }

// ---- host staging for PAGEABLE caller memory ------------------------------------------------------------------------
// A JS BigBuffer page is an ordinary (pageable) BigUint64Array unless it came from the addon's pinned allocator.  The driver
// stages pageable cudaMemcpyAsync copies through one internal bounce buffer on one thread (measured ~6-12 GB/s).  Here a
// small pool of host threads copies between the caller's page and a ring of pinned slots while the DMA engine moves the
// previous slot, so pageable pages travel at several times that rate and pinned pages (detected with
// cudaPointerGetAttributes) skip the ring entirely.
struct CopyPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::function<void(int)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    explicit CopyPool(int n) {
        for (int i = 0; i < n; i++) th.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> l(m); stop = true; }
        cv_work.notify_all();
        for (std::thread& t : th) t.join();
    }
    void run(int id) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int)> f;
            {
                std::unique_lock<std::mutex> l(m);
                cv_work.wait(l, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                f = job;
            }
            f(id);
            { std::lock_guard<std::mutex> l(m); if (--pending == 0) cv_done.notify_all(); }
        }
    }
    // runs fn(worker) on every worker and waits
    void all(const std::function<void(int)>& fn) {
        std::unique_lock<std::mutex> l(m);
        job = fn;
        pending = (int)th.size();
        gen++;
        cv_work.notify_all();
        cv_done.wait(l, [&] { return pending == 0; });
    }
    // rows x row_bytes block copy with independent pitches (pitch == row_bytes: one flat memcpy), split over the workers
    void copy2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t row_bytes, size_t rows) {
        const int n = (int)th.size();
        if (rows * row_bytes < (1u << 20) || n <= 1) {
            if (dpitch == row_bytes && spitch == row_bytes) memcpy(dst, src, rows * row_bytes);
            else for (size_t r = 0; r < rows; r++) memcpy(dst + r * dpitch, src + r * spitch, row_bytes);
            return;
        }
        if (dpitch == row_bytes && spitch == row_bytes) {   // flat: split by bytes
            const size_t total = rows * row_bytes, per = ((total + n - 1) / n + 63) & ~(size_t)63;
            all([&](int w) {
                const size_t a = (size_t)w * per, b = a + per < total ? a + per : total;
                if (a < b) memcpy(dst + a, src + a, b - a);
            });
        } else {
            const size_t per = (rows + n - 1) / n;
            all([&](int w) {
                const size_t a = (size_t)w * per, b = a + per < rows ? a + per : rows;
                for (size_t r = a; r < b; r++) memcpy(dst + r * dpitch, src + r * spitch, row_bytes);
            });
        }
    }
};
#define STAGE_SLOTS 4
#define STAGE_SLOT_BYTES ((size_t)32 << 20)
struct StageRing {
    char* slot[STAGE_SLOTS] = {};
    cudaEvent_t ev[STAGE_SLOTS] = {};
    CopyPool* pool = nullptr;
    ~StageRing() {
        delete pool;
        for (int i = 0; i < STAGE_SLOTS; i++) { if (slot[i]) cudaFreeHost(slot[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    }
};

struct pil2gpu_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    u64* tables;      // bytepow[1024] | tw_fwd[2^TMAX] | tw_inv[2^TMAX] | pow7[32]  (Montgomery form, see ntt.cuh) | small[256]
    u64* small;       // 256 words of per-call constants (quotient chunk factors)
    uint64_t launches;
    NttTables tb;
    cudaStream_t copy_stream;   // D2H stream: downloads of finished slabs overlap the compute (extend_and_merkelize)
    cudaStream_t in_stream;     // H2D stream: uploads of the next slab overlap the compute
    cudaEvent_t ev;
    cudaMemPool_t pool;         // stream-ordered scratch (cudaMallocFromPoolAsync); keeps what it frees (release threshold = max):
                                // with the default threshold of 0 every synchronisation hands the scratch back to the driver and the
                                // next call pays for mapping hundreds of MiB again (measured: 2x on the quotient commit)
    u64* ws;                    // grow-only device workspace of the host-pointer entry points (cudaMalloc/cudaFree of tens of GiB
    size_t ws_words;            // per call cost ~0.3 s at cfg3); released by pil2gpu_destroy / pil2gpu_release_workspace
    StageRing* stage;           // pinned slots + copy threads for pageable caller pages (created on first use)
};

struct pil2gpu_tree {
    u64* elems;
    u64* nodes;
    uint64_t width, height;
    uint64_t tile_cols, tile_stride;   // column tiling of elems (tile_cols == width: plain row-major)
    bool own_elems, own_nodes;
};

struct EventPool {
    std::vector<cudaEvent_t> ev;
    ~EventPool() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
    bool timing = false;
    cudaError_t make(cudaEvent_t* out) {
        cudaError_t e = cudaEventCreateWithFlags(out, timing ? cudaEventDefault : cudaEventDisableTiming);
        if (e == cudaSuccess) ev.push_back(*out);
        return e;
    }
};

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        int cur;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define ENTER(ctx)                                                        \
    if (!(ctx)) return fail(PIL2GPU_E_INVALID, "null context");           \
    DeviceGuard guard__((ctx)->device);                                   \
    if (!guard__.ok) return fail(PIL2GPU_E_CUDA, "cudaSetDevice(%d) failed", (ctx)->device)

static int check_launch(pil2gpu_ctx* ctx, int launches, const char* what) {
    if (launches < 0) return fail(PIL2GPU_E_CUDA, "%s: kernel configuration failed: %s", what, cudaGetErrorString(cudaGetLastError()));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
    ctx->launches += (uint64_t)launches;
    return PIL2GPU_OK;
}

// ---- utility kernels (bench / test) ----
__global__ void synth_kernel(u64* __restrict__ dst, u64 n, u64 seed, u64 first) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u64 z = (seed ^ (first + i)) + 0x9E3779B97F4A7C15ULL;       // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        dst[i] = gl_canon(z);
    }
}
__global__ void synth2d_kernel(u64* __restrict__ dst, u64 rows, u64 cols, u64 row_stride, u64 col0, u64 seed) {
    const u64 n = rows * cols, stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 r = i / cols, c = i - r * cols;
        u64 z = (seed ^ (r * row_stride + col0 + c)) + 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        dst[i] = gl_canon(z);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) pipe_probe_kernel(u64* out, int iters) {
    u64 a0 = threadIdx.x + 1, a1 = blockIdx.x + 3, a2 = a0 ^ 0x1234567, a3 = a0 + a1 + 77;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {
                a0 = gl_mmul(a0, a1); a1 = gl_mmul(a1, a2); a2 = gl_mmul(a2, a3); a3 = gl_mmul(a3, a0);
            } else {
                a0 += (u64)(u32)a0 * 0x9E3779B9u; a1 += (u64)(u32)a1 * 0x85EBCA6Bu;
                a2 += (u64)(u32)a2 * 0xC2B2AE35u; a3 += (u64)(u32)a3 * 0x27D4EB2Fu;
            }
        }
    }
    out[(u64)blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3;
}

extern "C" {

const char* pil2gpu_last_error(void) { return g_last_error.c_str(); }
const char* pil2gpu_version(void) { return "pil2gpu 0.1 (sm_100a)"; }

int pil2gpu_create(int device, void* stream, pil2gpu_ctx** out) {
    if (!out) return fail(PIL2GPU_E_INVALID, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PIL2GPU_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(PIL2GPU_E_INVALID, "device %d out of range (have %d)", device, ndev);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(PIL2GPU_E_CUDA, "cudaSetDevice(%d) failed", device);
    pil2gpu_ctx* ctx = new (std::nothrow) pil2gpu_ctx();
    if (!ctx) return fail(PIL2GPU_E_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->launches = 0;
    ctx->copy_stream = nullptr;
    ctx->in_stream = nullptr;
    ctx->ev = nullptr;
    ctx->ws = nullptr;
    ctx->ws_words = 0;
    ctx->stage = nullptr;
    ctx->pool = nullptr;
    ctx->tables = nullptr;
    ctx->stream = nullptr;
    ctx->own_stream = false;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(PIL2GPU_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        ctx->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->in_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming) != cudaSuccess) {
        pil2gpu_destroy(ctx);
        return fail(PIL2GPU_E_CUDA, "stream/event creation failed");
    }
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        unsigned long long keep = ~0ULL;
        if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess ||
            cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) {
            pil2gpu_destroy(ctx);
            return fail(PIL2GPU_E_CUDA, "memory pool creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    const size_t words = NTT_TABLE_WORDS + 256;
    e = cudaMalloc(&ctx->tables, words * sizeof(u64));
    if (e != cudaSuccess) { pil2gpu_destroy(ctx); return fail(PIL2GPU_E_NOMEM, "cudaMalloc(tables): %s", cudaGetErrorString(e)); }
    u64* tw_fwd = ctx->tables + 1024;
    u64* tw_inv = tw_fwd + (1u << NTT_TMAX);
    u64* pow7 = tw_inv + (1u << NTT_TMAX);
    ntt_setup_tables<<<4, 256, 0, ctx->stream>>>(ctx->tables, tw_fwd, tw_inv, pow7);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        pil2gpu_destroy(ctx);
        return fail(PIL2GPU_E_CUDA, "table setup kernel failed: %s (is this an sm_100a device?)", cudaGetErrorString(e));
    }
    ctx->launches++;
    ctx->tb.bytepow = ctx->tables;
    ctx->tb.tw_fwd = tw_fwd;
    ctx->tb.tw_inv = tw_inv;
    ctx->tb.pow7 = pow7;
    ctx->small = ctx->tables + NTT_TABLE_WORDS;
    *out = ctx;
    return PIL2GPU_OK;
}

void pil2gpu_destroy(pil2gpu_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->in_stream) { cudaStreamSynchronize(ctx->in_stream); cudaStreamDestroy(ctx->in_stream); }
    if (ctx->ev) cudaEventDestroy(ctx->ev);
    if (ctx->tables) cudaFree(ctx->tables);
    if (ctx->ws) cudaFree(ctx->ws);
    delete ctx->stage;
    if (ctx->pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(ctx->pool); }
    delete ctx;
}

int pil2gpu_sync(pil2gpu_ctx* ctx) {
    ENTER(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}
uint64_t pil2gpu_launch_count(const pil2gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

// The pipelined entry points also enqueue on the two copy streams; whatever frees or reuses the workspace waits for all three.
static cudaError_t sync_all_streams(pil2gpu_ctx* ctx) {
    cudaError_t e1 = cudaStreamSynchronize(ctx->stream), e2 = ctx->in_stream ? cudaStreamSynchronize(ctx->in_stream) : cudaSuccess,
                e3 = ctx->copy_stream ? cudaStreamSynchronize(ctx->copy_stream) : cudaSuccess;
    return e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
}
static inline size_t ev2(size_t w) { return (w + 1) & ~(size_t)1; }   // keep 16-byte alignment of workspace carve-outs
static int ensure_ws(pil2gpu_ctx* ctx, size_t words) {
    if (ctx->ws_words >= words) return PIL2GPU_OK;
    if (ctx->ws) {
        CU(sync_all_streams(ctx));
        CU(cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_words = 0;
    }
    CU(cudaMalloc(&ctx->ws, words * sizeof(u64)));
    ctx->ws_words = words;
    return PIL2GPU_OK;
}
int pil2gpu_release_workspace(pil2gpu_ctx* ctx) {
    ENTER(ctx);
    CU(sync_all_streams(ctx));
    if (ctx->pool) CU(cudaMemPoolTrimTo(ctx->pool, 0));
    if (ctx->ws) {
        CU(cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_words = 0;
    }
    return PIL2GPU_OK;
}

int pil2gpu_dev_alloc(pil2gpu_ctx* ctx, size_t bytes, void** dptr) {
    ENTER(ctx);
    if (!dptr) return fail(PIL2GPU_E_INVALID, "null dptr");
    CU(cudaMalloc(dptr, bytes ? bytes : 8));
    return PIL2GPU_OK;
}
int pil2gpu_dev_free(pil2gpu_ctx* ctx, void* dptr) {
    ENTER(ctx);
    if (dptr) CU(cudaFree(dptr));
    return PIL2GPU_OK;
}
int pil2gpu_host_alloc(size_t bytes, void** hptr) {
    if (!hptr) return fail(PIL2GPU_E_INVALID, "null hptr");
    CU(cudaHostAlloc(hptr, bytes ? bytes : 8, cudaHostAllocPortable));
    return PIL2GPU_OK;
}
int pil2gpu_host_free(void* hptr) {
    if (hptr) CU(cudaFreeHost(hptr));
    return PIL2GPU_OK;
}
int pil2gpu_h2d(pil2gpu_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    ENTER(ctx);
    if (bytes) CU(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PIL2GPU_OK;
}
int pil2gpu_d2h(pil2gpu_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    ENTER(ctx);
    if (bytes) CU(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PIL2GPU_OK;
}

// ------------------------------------------------------------------------------------------------------------
// NTT / LDE
// ------------------------------------------------------------------------------------------------------------
static int check_ntt_args(const void* src, const void* dst, uint64_t nPols, uint32_t nBits) {
    if (!src || !dst) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (nPols == 0) return fail(PIL2GPU_E_INVALID, "nPols must be > 0");
    if (nBits > 32) return fail(PIL2GPU_E_INVALID, "nBits %u exceeds the 2-adicity of the field (32)", nBits);
    return PIL2GPU_OK;
}

int pil2gpu_ntt_dev(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t* dst, uint64_t nPols, uint32_t nBits, int inverse) {
    ENTER(ctx);
    int rc = check_ntt_args(src, dst, nPols, nBits);
    if (rc) return rc;
    const size_t words = (size_t)nPols << nBits;
    if ((const u64*)src < (const u64*)dst + words && (const u64*)dst < (const u64*)src + words)
        return fail(PIL2GPU_E_INVALID, "src and dst overlap");
    int l = ntt_launch_transform((const u64*)src, (u64*)dst, nPols, (int)nBits, inverse != 0, ctx->tb, ctx->stream);
    return check_launch(ctx, l, "ntt");
}

int pil2gpu_lde_dev(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t* dst, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt) {
    ENTER(ctx);
    int rc = check_ntt_args(src, dst, nPols, nBitsExt);
    if (rc) return rc;
    if (nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "nBitsExt (%u) < nBits (%u)", nBitsExt, nBits);
    if (nBitsExt - nBits > 8) return fail(PIL2GPU_E_UNSUPPORTED, "blowup 2^%u not supported (max 2^8)", nBitsExt - nBits);
    const size_t sw = (size_t)nPols << nBits, dw = (size_t)nPols << nBitsExt;
    if ((const u64*)src < (const u64*)dst + dw && (const u64*)dst < (const u64*)src + sw) return fail(PIL2GPU_E_INVALID, "src and dst overlap");
    int l = ntt_launch_lde((const u64*)src, (u64*)dst, nPols, (int)nBits, (int)nBitsExt, ctx->tb, ctx->stream);
    return check_launch(ctx, l, "lde");
}

// ---- multi-GPU row exchange fused into the LDE ----
int pil2gpu_ipc_export(pil2gpu_ctx* ctx, const void* dptr, uint8_t handle_out[64]) {
    ENTER(ctx);
    if (!dptr || !handle_out) return fail(PIL2GPU_E_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void*>(dptr)));
    memcpy(handle_out, &h, 64);
    return PIL2GPU_OK;
}
int pil2gpu_ipc_open(pil2gpu_ctx* ctx, const uint8_t handle[64], void** dptr_out) {
    ENTER(ctx);
    if (!handle || !dptr_out) return fail(PIL2GPU_E_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return PIL2GPU_OK;
}
int pil2gpu_ipc_close(pil2gpu_ctx* ctx, void* dptr) {
    ENTER(ctx);
    if (dptr) CU(cudaIpcCloseMemHandle(dptr));
    return PIL2GPU_OK;
}

static int make_scatter(NttScatter& sc, uint64_t nPols, uint32_t nBitsExt, uint64_t* const* peer_recv_dev, uint32_t n_ranks, uint32_t rank,
                        uint64_t tile_cols, uint64_t col_off) {
    if (!peer_recv_dev || n_ranks == 0 || n_ranks > NTT_MAX_PEERS || (n_ranks & (n_ranks - 1)) || rank >= n_ranks)
        return fail(PIL2GPU_E_INVALID, "bad rank description (n_ranks must be a power of two <= %d)", NTT_MAX_PEERS);
    if (col_off + nPols > tile_cols) return fail(PIL2GPU_E_INVALID, "slab columns [%llu, %llu) outside a tile of %llu columns",
                                                 (unsigned long long)col_off, (unsigned long long)(col_off + nPols), (unsigned long long)tile_cols);
    uint32_t gb = 0;
    while ((1u << gb) < n_ranks) gb++;
    if (gb > nBitsExt) return fail(PIL2GPU_E_INVALID, "more ranks than extended rows");
    sc = ntt_no_scatter();
    for (uint32_t h = 0; h < n_ranks; h++) {
        if (!peer_recv_dev[h]) return fail(PIL2GPU_E_INVALID, "null peer buffer %u", h);
        sc.peer[h] = (u64*)peer_recv_dev[h];
    }
    sc.rl_bits = (int)(nBitsExt - gb);
    sc.tile_off = (u64)rank * (tile_cols << sc.rl_bits);
    sc.out_C = tile_cols;
    sc.col_off = col_off;
    return PIL2GPU_OK;
}

int pil2gpu_lde_scatter_dev(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t* dst, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt,
                            uint64_t* const* peer_recv_dev, uint32_t n_ranks, uint32_t rank, uint64_t tile_cols, uint64_t col_off) {
    ENTER(ctx);
    int rc = check_ntt_args(src, dst, nPols, nBitsExt);
    if (rc) return rc;
    if (nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "nBitsExt (%u) < nBits (%u)", nBitsExt, nBits);
    if (nBitsExt - nBits > 8) return fail(PIL2GPU_E_UNSUPPORTED, "blowup 2^%u not supported (max 2^8)", nBitsExt - nBits);
    NttScatter sc;
    rc = make_scatter(sc, nPols, nBitsExt, peer_recv_dev, n_ranks, rank, tile_cols ? tile_cols : nPols, col_off);
    if (rc) return rc;
    int l = ntt_launch_lde((const u64*)src, (u64*)dst, nPols, (int)nBits, (int)nBitsExt, ctx->tb, ctx->stream, &sc);
    return check_launch(ctx, l, "lde_scatter");
}

// Host-source form: the slab is cut into sub-slabs of PIPE columns; the upload of sub-slab s+1 (in_stream) overlaps the
// LDE + peer stores of sub-slab s (ctx stream).  Returns once everything is enqueued; the work is complete when the ctx
// stream is (pil2gpu_sync, or a collective ordered after it).
int pil2gpu_lde_scatter(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t src_pitch_cols, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt,
                        uint64_t* const* peer_recv_dev, uint32_t n_ranks, uint32_t rank) {
    ENTER(ctx);
    if (!src) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (nPols == 0 || nBitsExt > 32 || nBitsExt < nBits || src_pitch_cols < nPols) return fail(PIL2GPU_E_INVALID, "bad LDE shape");
    if (nBitsExt - nBits > 8) return fail(PIL2GPU_E_UNSUPPORTED, "blowup 2^%u not supported (max 2^8)", nBitsExt - nBits);
    const u64 N = 1ULL << nBits, E = 1ULL << nBitsExt;
    const u64 cs = (nPols % 32 == 0 && nPols > 32) ? 32 : ((nPols % 16 == 0 && nPols > 16) ? 16 : nPols);
    const u64 nslabs = nPols / cs;
    int rc = ensure_ws(ctx, 2 * N * cs + 2 * E * cs);
    if (rc) return rc;
    u64* sbuf[2] = {ctx->ws, ctx->ws + N * cs};
    u64* dbuf[2] = {ctx->ws + 2 * N * cs, ctx->ws + 2 * N * cs + E * cs};
    EventPool pool;
    std::vector<cudaEvent_t> ev_in(nslabs), ev_done(nslabs);
    for (u64 s = 0; s < nslabs; s++) { CU(pool.make(&ev_in[s])); CU(pool.make(&ev_done[s])); }
    cudaEvent_t ev_start;
    CU(pool.make(&ev_start));
    CU(cudaEventRecord(ev_start, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->in_stream, ev_start, 0));
    rc = PIL2GPU_OK;
    for (u64 s = 0; s < nslabs; s++) {
        const int b = (int)(s & 1);
        if (s >= 2) CU_BREAK(cudaStreamWaitEvent(ctx->in_stream, ev_done[s - 2], 0));      // sbuf[b] is free again
        CU_BREAK(cudaMemcpy2DAsync(sbuf[b], cs * 8, src + s * cs, src_pitch_cols * 8, cs * 8, N, cudaMemcpyHostToDevice, ctx->in_stream));
        CU_BREAK(cudaEventRecord(ev_in[s], ctx->in_stream));
        CU_BREAK(cudaStreamWaitEvent(ctx->stream, ev_in[s], 0));
        NttScatter sc;
        rc = make_scatter(sc, cs, nBitsExt, peer_recv_dev, n_ranks, rank, nPols, s * cs);
        if (rc) break;
        int l = ntt_launch_lde(sbuf[b], dbuf[b], cs, (int)nBits, (int)nBitsExt, ctx->tb, ctx->stream, &sc);
        rc = check_launch(ctx, l, "lde_scatter");
        if (rc) break;
        CU_BREAK(cudaEventRecord(ev_done[s], ctx->stream));
    }
    // the events may be destroyed while still pending (CUDA keeps them alive until they complete); the in_stream must not
    // be left with work that outlives the workspace on the error path
    if (rc) { cudaStreamSynchronize(ctx->in_stream); cudaStreamSynchronize(ctx->stream); }
    return rc;
}

// ---- quotient polynomial path: computeQStark (stark_gen_helpers.js:168-208) ----
int pil2gpu_compute_q_dev(pil2gpu_ctx* ctx, const uint64_t* q_ext, uint64_t qDim, uint64_t qDeg, uint32_t nBits, uint32_t nBitsExt,
                          uint64_t* cmq_ext) {
    ENTER(ctx);
    if (!q_ext || !cmq_ext) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (qDim == 0 || qDeg == 0) return fail(PIL2GPU_E_INVALID, "qDim and qDeg must be > 0");
    if (nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad sizes");
    if (nBitsExt - nBits > 8) return fail(PIL2GPU_E_UNSUPPORTED, "blowup 2^%u not supported (max 2^8)", nBitsExt - nBits);
    const u64 B = 1ULL << (nBitsExt - nBits), N = 1ULL << nBits, E = 1ULL << nBitsExt;
    if (qDeg > B) return fail(PIL2GPU_E_INVALID, "qDeg (%llu) exceeds the blowup factor (%llu): the quotient does not fit the extended domain",
                              (unsigned long long)qDeg, (unsigned long long)B);
    // chunk factors shiftIn^p / E, shiftIn = 7^-N (stark_gen_helpers.js:178-190), Montgomery form
    u64 fac[256];
    const u64 shift_in = glh_pow(glh_inv(GL_SHIFT), N);
    u64 cur = glh_inv(E % GL_P);
    for (u64 p = 0; p < qDeg; p++) { fac[p] = glh_to_mont(cur); cur = glh_mul(cur, shift_in); }
    CU(cudaMemcpyAsync(ctx->small, fac, qDeg * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    u64 *S = nullptr, *T = nullptr;
    CU(cudaMallocFromPoolAsync(&S, E * qDim * sizeof(u64), ctx->pool, ctx->stream));
    cudaError_t e = cudaMallocFromPoolAsync(&T, N * qDim * qDeg * sizeof(u64), ctx->pool, ctx->stream);
    if (e != cudaSuccess) { cudaFreeAsync(S, ctx->stream); return fail(PIL2GPU_E_NOMEM, "scratch allocation failed: %s", cudaGetErrorString(e)); }
    int launches = 0;
    int l = ntt_launch_intt_bitrev((const u64*)q_ext, S, qDim, (int)nBitsExt, ctx->tb, ctx->stream);
    if (l >= 0) {
        launches += l;
        const u64 total = N * qDim * qDeg;
        const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        q_split_kernel<<<blocks, 256, 0, ctx->stream>>>(S, T, ctx->small, N, (int)(nBitsExt - nBits), (u32)qDim, (u32)qDeg);
        launches++;
        l = ntt_launch_coset_eval(T, (u64*)cmq_ext, qDim * qDeg, (int)nBits, (int)nBitsExt, true, ctx->tb, ctx->stream);
        if (l >= 0) launches += l;
    }
    cudaFreeAsync(S, ctx->stream);
    cudaFreeAsync(T, ctx->stream);
    return check_launch(ctx, l < 0 ? -1 : launches, "compute_q");
}

// ---- evaluations at xi and FRI denominators: computeEvalsStark / computeFRIStark (stark_gen_helpers.js:210-323) ----
static void host_opening_xi(u64 xi[3], const uint64_t xi_challenge[3], int32_t opening, uint32_t nBits) {   // :222-226
    u64 w = 1;
    const u64 wn = glh_root(nBits);
    for (int32_t j = 0; j < (opening < 0 ? -opening : opening); j++) w = glh_mul(w, wn);
    if (opening < 0) w = glh_inv(w);
    for (int c = 0; c < 3; c++) xi[c] = glh_mul(xi_challenge[c] % GL_P, w);
}

int pil2gpu_compute_lev_dev(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], int32_t opening, uint32_t nBits, uint64_t* lev_dev) {
    ENTER(ctx);
    if (!xi_challenge || !lev_dev) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nBits > 32) return fail(PIL2GPU_E_INVALID, "nBits %u exceeds the 2-adicity of the field (32)", nBits);
    const u64 N = 1ULL << nBits;
    u64 xi[3];
    host_opening_xi(xi, xi_challenge, opening, nBits);
    const u64 sinv = glh_inv(GL_SHIFT);
    for (int c = 0; c < 3; c++) xi[c] = glh_mul(xi[c], sinv);                       // :227
    u64* pw = nullptr;
    CU(cudaMallocFromPoolAsync(&pw, N * 3 * sizeof(u64), ctx->pool, ctx->stream));
    const u64 threads_needed = (N + EV_POW_CHUNK - 1) / EV_POW_CHUNK;
    f3_powers_kernel<<<(unsigned)((threads_needed + 255) / 256), 256, 0, ctx->stream>>>(xi[0], xi[1], xi[2], N, pw);   // :228-230
    int l = ntt_launch_transform(pw, (u64*)lev_dev, 3, (int)nBits, true, ctx->tb, ctx->stream);                        // :231
    cudaFreeAsync(pw, ctx->stream);
    return check_launch(ctx, l < 0 ? -1 : l + 1, "compute_lev");
}

int pil2gpu_compute_evals_dev(pil2gpu_ctx* ctx, const uint64_t* buf_dev, uint64_t size, uint32_t nBits, uint32_t nBitsExt,
                              const pil2gpu_eval_desc* desc, uint32_t n_evals, const uint64_t* lev_dev, uint32_t n_lev, uint64_t* evals_out) {
    ENTER(ctx);
    if (n_evals == 0) return PIL2GPU_OK;
    if (!buf_dev || !desc || !lev_dev || !evals_out) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad sizes");
    static_assert(sizeof(pil2gpu_eval_desc) == sizeof(EvalDesc), "descriptor layout");
    for (uint32_t e = 0; e < n_evals; e++) {
        if (desc[e].dim != 1 && desc[e].dim != 3) return fail(PIL2GPU_E_INVALID, "evaluation %u: dim must be 1 or 3", e);
        if (desc[e].offset + desc[e].dim > size) return fail(PIL2GPU_E_RANGE, "evaluation %u: columns [%llu, %llu) outside a row of %llu", e,
                                                             (unsigned long long)desc[e].offset, (unsigned long long)(desc[e].offset + desc[e].dim),
                                                             (unsigned long long)size);
        if (desc[e].lev >= n_lev) return fail(PIL2GPU_E_RANGE, "evaluation %u: opening index %u out of range", e, desc[e].lev);
    }
    const u64 N = 1ULL << nBits;
    const int eb = (int)(nBitsExt - nBits);
    // Tensor-core path (evals.cuh: byte-limb GEMM): the base rows are read once for all evaluations.  PIL2GPU_EVALS=scalar forces
    // the per-evaluation kernel (kept for small inputs and as the cross-check in the tests).
    const char* mode = getenv("PIL2GPU_EVALS");
    const bool use_mma = nBits >= 10 && n_lev <= 4 && !(mode && strcmp(mode, "scalar") == 0);
    // (measured crossover at 2^22 rows: 128 columns 1.92 vs 2.72 ms, 64 columns 1.86 vs 1.53 ms -- its A-fragment work is per row, whatever the width)
    if (use_mma && n_lev <= 2 && (size >= 96 || (mode && strcmp(mode, "mma2") == 0)) && !(mode && strcmp(mode, "mma1") == 0)) {
        // second formulation (evals_mma2_kernel): the limb weight sits in the LEv operand, the buffer rows are MMA fragments as they lie
        const u32 n_oc = 3 * n_lev;
        const u64 colgroups = (size + EV2_COLS - 1) / EV2_COLS;
        u64 chunks = (148 * 4) / colgroups;
        if (chunks < 1) chunks = 1;
        if (chunks > N / 32) chunks = N / 32;
        const u64 rpc = (((N + chunks - 1) / chunks) + 31) & ~(u64)31;
        chunks = (N + rpc - 1) / rpc;
        const size_t desc_words = ((size_t)n_evals * sizeof(EvalDesc) + 7) / 8, part_words = chunks * size * n_oc, out_words = (size_t)n_evals * 3;
        u64* scratch = nullptr;
        CU(cudaMallocFromPoolAsync(&scratch, (desc_words + part_words + out_words) * sizeof(u64), ctx->pool, ctx->stream));
        EvalDesc* ddesc = reinterpret_cast<EvalDesc*>(scratch);
        u64 *partial = scratch + desc_words, *dout = partial + part_words;
        cudaError_t e = cudaMemcpyAsync(ddesc, desc, (size_t)n_evals * sizeof(EvalDesc), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            dim3 grid((unsigned)colgroups, (unsigned)chunks, 1);
            e = cudaFuncSetAttribute(evals_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EV2_SMEM);   // per device: set on every launch
            if (e == cudaSuccess) {
                evals_mma2_kernel<<<grid, EV2_THREADS, EV2_SMEM, ctx->stream>>>((const u64*)buf_dev, size, eb, N, rpc, (const u64*)lev_dev, n_lev, partial);
                evals_gather_kernel<<<(n_evals + 3) / 4, 128, 0, ctx->stream>>>(partial, (u32)chunks, size, n_lev, ddesc, n_evals, dout);
                e = cudaMemcpyAsync(evals_out, dout, out_words * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
            }
        }
        cudaFreeAsync(scratch, ctx->stream);
        if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "compute_evals: %s", cudaGetErrorString(e));
        int rc = check_launch(ctx, 2, "compute_evals");
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        return PIL2GPU_OK;
    }
    if (use_mma) {
        const u32 n_oc = 3 * n_lev, MT = (24 * n_lev + 15) / 16, M = MT * 16;
        u64 rpc = (N / 64) & ~(u64)(EVM_KSTEP - 1);
        if (rpc < EVM_KSTEP) rpc = EVM_KSTEP;
        if (rpc > EVM_MAX_CHUNK) rpc = EVM_MAX_CHUNK;
        const u64 chunks = (N + rpc - 1) / rpc;
        const size_t lt_words = (N * M + 7) / 8, desc_words = ((size_t)n_evals * sizeof(EvalDesc) + 7) / 8,
                     part_words = chunks * size * n_oc, out_words = (size_t)n_evals * 3;
        u64* scratch = nullptr;
        CU(cudaMallocFromPoolAsync(&scratch, (lt_words + desc_words + part_words + out_words) * sizeof(u64), ctx->pool, ctx->stream));
        unsigned char* LT = reinterpret_cast<unsigned char*>(scratch);
        EvalDesc* ddesc = reinterpret_cast<EvalDesc*>(scratch + lt_words);
        u64 *partial = scratch + lt_words + desc_words, *dout = partial + part_words;
        cudaError_t e = cudaMemcpyAsync(ddesc, desc, (size_t)n_evals * sizeof(EvalDesc), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            lev_bytes_kernel<<<148 * 8, 256, 0, ctx->stream>>>((const u64*)lev_dev, N, n_lev, M, LT);
            dim3 grid((unsigned)((size + EVM_COLS - 1) / EVM_COLS), (unsigned)chunks, 1);
            switch (MT) {
                case 2: evals_mma_kernel<2><<<grid, EVM_THREADS, 0, ctx->stream>>>((const u64*)buf_dev, size, eb, N, rpc, LT, n_lev, partial); break;
                case 3: evals_mma_kernel<3><<<grid, EVM_THREADS, 0, ctx->stream>>>((const u64*)buf_dev, size, eb, N, rpc, LT, n_lev, partial); break;
                case 5: evals_mma_kernel<5><<<grid, EVM_THREADS, 0, ctx->stream>>>((const u64*)buf_dev, size, eb, N, rpc, LT, n_lev, partial); break;
                default: evals_mma_kernel<6><<<grid, EVM_THREADS, 0, ctx->stream>>>((const u64*)buf_dev, size, eb, N, rpc, LT, n_lev, partial); break;
            }
            evals_gather_kernel<<<(n_evals + 3) / 4, 128, 0, ctx->stream>>>(partial, (u32)chunks, size, n_lev, ddesc, n_evals, dout);
            e = cudaMemcpyAsync(evals_out, dout, out_words * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
        }
        cudaFreeAsync(scratch, ctx->stream);
        if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "compute_evals: %s", cudaGetErrorString(e));
        int rc = check_launch(ctx, 3, "compute_evals");
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        return PIL2GPU_OK;
    }
    u32 ew = 1;
    while (ew < n_evals && ew < EV_THREADS) ew <<= 1;
    u64 chunks = N / 64 ? N / 64 : 1;                 // >= 64 rows per CTA, at most 8 CTAs per SM
    if (chunks > 148 * 8) chunks = 148 * 8;
    u64* scratch = nullptr;                           // desc | partial | out
    const size_t desc_words = ((size_t)n_evals * sizeof(EvalDesc) + 7) / 8, part_words = chunks * n_evals * 3, out_words = (size_t)n_evals * 3;
    CU(cudaMallocFromPoolAsync(&scratch, (desc_words + part_words + out_words) * sizeof(u64), ctx->pool, ctx->stream));
    EvalDesc* ddesc = reinterpret_cast<EvalDesc*>(scratch);
    u64 *partial = scratch + desc_words, *dout = partial + part_words;
    cudaError_t e = cudaMemcpyAsync(ddesc, desc, (size_t)n_evals * sizeof(EvalDesc), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        evals_partial_kernel<<<(unsigned)chunks, EV_THREADS, EV_THREADS * 3 * sizeof(u64), ctx->stream>>>(
            (const u64*)buf_dev, size, eb, N, ddesc, n_evals, (const u64*)lev_dev, ew, partial);
        evals_reduce_kernel<<<(n_evals * 3 + 127) / 128, 128, 0, ctx->stream>>>(partial, (u32)chunks, n_evals, dout);
        e = cudaMemcpyAsync(evals_out, dout, out_words * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaFreeAsync(scratch, ctx->stream);
    if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "compute_evals: %s", cudaGetErrorString(e));
    int rc = check_launch(ctx, 2, "compute_evals");
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

int pil2gpu_x_div_x_sub_xi_dev(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], const int32_t* openings, uint32_t n_open, uint32_t nBits,
                               uint32_t nBitsExt, uint64_t* out_dev) {
    ENTER(ctx);
    if (n_open == 0) return PIL2GPU_OK;
    if (!xi_challenge || !openings || !out_dev) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad sizes");
    if (n_open > 64) return fail(PIL2GPU_E_UNSUPPORTED, "at most 64 opening points");
    u64 xi[64 * 3];
    for (uint32_t i = 0; i < n_open; i++) host_opening_xi(xi + 3 * i, xi_challenge, openings[i], nBits);   // :291-300
    u64* dxi = nullptr;
    CU(cudaMallocFromPoolAsync(&dxi, (size_t)n_open * 3 * sizeof(u64), ctx->pool, ctx->stream));
    cudaError_t e = cudaMemcpyAsync(dxi, xi, (size_t)n_open * 3 * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const u64 E = 1ULL << nBitsExt, per = (u64)XDIV_THREADS * XDIV_BATCH;
        dim3 grid((unsigned)((E + per - 1) / per), n_open, 1);
        xdiv_kernel<<<grid, XDIV_THREADS, 0, ctx->stream>>>(dxi, n_open, (int)nBitsExt, ctx->tb, (u64*)out_dev);
    }
    cudaFreeAsync(dxi, ctx->stream);
    if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "x_div_x_sub_xi: %s", cudaGetErrorString(e));
    return check_launch(ctx, 1, "x_div_x_sub_xi");
}

// Host-buffer forms (what the JS shims call): only the 2^nBits base rows of the extended buffer travel over PCIe.
int pil2gpu_compute_evals(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], const int32_t* openings, uint32_t n_open, uint32_t nBits,
                          uint32_t nBitsExt, const uint64_t* buf, uint64_t size, const pil2gpu_eval_desc* desc, uint32_t n_evals,
                          uint64_t* evals_out) {
    ENTER(ctx);
    if (n_evals == 0) return PIL2GPU_OK;
    if (!xi_challenge || !openings || !buf || !desc || !evals_out || n_open == 0) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nBitsExt > 32 || nBitsExt < nBits || size == 0) return fail(PIL2GPU_E_INVALID, "bad sizes");
    const u64 N = 1ULL << nBits;
    int rc = ensure_ws(ctx, ev2(N * size) + (size_t)n_open * N * 3);
    if (rc) return rc;
    u64 *rows = ctx->ws, *lev = ctx->ws + ev2(N * size);
    // rows k << extendBits of the extended buffer, compacted (stark_gen_helpers.js:252-259 reads nothing else)
    CU(cudaMemcpy2DAsync(rows, size * 8, buf, (size << (nBitsExt - nBits)) * 8, size * 8, N, cudaMemcpyHostToDevice, ctx->stream));
    for (uint32_t i = 0; i < n_open; i++) {
        rc = pil2gpu_compute_lev_dev(ctx, xi_challenge, openings[i], nBits, lev + (size_t)i * N * 3);
        if (rc) return rc;
    }
    return pil2gpu_compute_evals_dev(ctx, rows, size, nBits, nBits, desc, n_evals, lev, n_open, evals_out);
}

int pil2gpu_x_div_x_sub_xi(pil2gpu_ctx* ctx, const uint64_t xi_challenge[3], const int32_t* openings, uint32_t n_open, uint32_t nBits,
                           uint32_t nBitsExt, uint64_t* out) {
    ENTER(ctx);
    if (n_open == 0) return PIL2GPU_OK;
    if (!out) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nBitsExt > 32) return fail(PIL2GPU_E_INVALID, "bad sizes");
    const size_t words = ((size_t)3 * n_open) << nBitsExt;
    int rc = ensure_ws(ctx, words);
    if (rc) return rc;
    rc = pil2gpu_x_div_x_sub_xi_dev(ctx, xi_challenge, openings, n_open, nBits, nBitsExt, ctx->ws);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, ctx->ws, words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

// ---- constraint expressions over a domain: calculateExps (prover_helpers.js:33-110) ----
int pil2gpu_calculate_exps_dev(pil2gpu_ctx* ctx, const uint32_t* ops, uint32_t n_ops, const uint64_t* consts, uint32_t n_consts,
                               const pil2gpu_expr_buffer* bufs, uint32_t n_bufs, uint32_t domain_bits, int x_shift) {
    ENTER(ctx);
    if (n_ops == 0) return PIL2GPU_OK;
    if (!ops || (!consts && n_consts) || (!bufs && n_bufs)) return fail(PIL2GPU_E_INVALID, "null argument");
    if (domain_bits > 32) return fail(PIL2GPU_E_INVALID, "domain of 2^%u rows exceeds the 2-adicity of the field", domain_bits);
    if (n_bufs > EXPR_MAX_BUFS) return fail(PIL2GPU_E_UNSUPPORTED, "more than %d buffers", EXPR_MAX_BUFS);
    ExprBufs eb;
    memset(&eb, 0, sizeof(eb));
    for (uint32_t b = 0; b < n_bufs; b++) {
        if (!bufs[b].ptr_dev || bufs[b].row_words == 0) return fail(PIL2GPU_E_INVALID, "buffer %u: null pointer or empty rows", b);
        eb.ptr[b] = (u64*)bufs[b].ptr_dev;
        eb.row_words[b] = bufs[b].row_words;
    }
    const u64 N = 1ULL << domain_bits;
    auto check_operand = [&](const uint32_t* o, bool is_dest, uint32_t k, const char* what) -> int {
        const uint32_t kind = o[0] & 255, dim = (o[0] >> 8) & 255, b = o[0] >> 16;
        if (dim != 1 && dim != 3) return fail(PIL2GPU_E_INVALID, "record %u: %s has dimension %u (1 or 3)", k, what, dim);
        switch (kind) {
            case EXPR_K_TMP:
                if (o[1] >= EXPR_MAX_SLOTS) return fail(PIL2GPU_E_UNSUPPORTED, "record %u: more than %d live temporaries", k, EXPR_MAX_SLOTS);
                return PIL2GPU_OK;
            case EXPR_K_CONST:
                if (is_dest) break;
                if (o[1] >= n_consts) return fail(PIL2GPU_E_RANGE, "record %u: constant %u out of range", k, o[1]);
                return PIL2GPU_OK;
            case EXPR_K_BUF:
                if (b >= n_bufs) return fail(PIL2GPU_E_RANGE, "record %u: buffer %u out of range", k, b);
                if ((u64)o[1] + dim > eb.row_words[b]) return fail(PIL2GPU_E_RANGE, "record %u: columns [%u, %u) outside a row of %llu words", k, o[1], o[1] + dim,
                                                                   (unsigned long long)eb.row_words[b]);
                if ((u64)o[2] >= N && N > 0 && o[2] != 0) return fail(PIL2GPU_E_RANGE, "record %u: row offset %u outside the domain", k, o[2]);
                return PIL2GPU_OK;
            case EXPR_K_X:
                if (is_dest) break;
                if (dim != 1) return fail(PIL2GPU_E_INVALID, "record %u: x is a base-field value", k);
                return PIL2GPU_OK;
            default: break;
        }
        return fail(PIL2GPU_E_INVALID, "record %u: invalid %s kind %u", k, what, kind);
    };
    for (uint32_t k = 0; k < n_ops; k++) {
        const uint32_t* o = ops + (size_t)k * EXPR_OP_WORDS;
        if (o[0] > EXPR_MULADD) return fail(PIL2GPU_E_INVALID, "record %u: invalid opcode %u", k, o[0]);
        const uint32_t nsrc = o[0] == EXPR_COPY ? 1 : (o[0] == EXPR_MULADD ? 3 : 2);
        if (o[1] != nsrc) return fail(PIL2GPU_E_INVALID, "record %u: opcode %u takes %u sources", k, o[0], nsrc);
        int rc = check_operand(o + 4, true, k, "destination");
        for (uint32_t j = 0; j < nsrc && !rc; j++) rc = check_operand(o + 7 + 3 * j, false, k, "source");
        if (rc) return rc;
    }
    // Compiled path (expr_jit.cuh): the records become a straight-line kernel, built once per (device, program) with NVRTC.
    // PIL2GPU_EXPR=interp forces the interpreter, =jit makes a missing / failing compiler an error instead of a fallback.
    const char* xmode = getenv("PIL2GPU_EXPR");
    if (!(xmode && strcmp(xmode, "interp") == 0)) {
        cudaKernel_t jk = expr_jit_get(ops, n_ops, eb.row_words, domain_bits, x_shift != 0);
        if (!jk && xmode && strcmp(xmode, "jit") == 0) return fail(PIL2GPU_E_UNSUPPORTED, "calculate_exps: %s", expr_jit_error.c_str());
        if (jk) {
            const size_t cb = (size_t)(n_consts ? n_consts : 1) * 3 * sizeof(u64);
            u64* d_c = nullptr;
            CU(cudaMallocFromPoolAsync(&d_c, cb, ctx->pool, ctx->stream));
            cudaError_t je = n_consts ? cudaMemcpyAsync(d_c, consts, (size_t)n_consts * 3 * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
            const u64* d_cc = d_c;
            const u64* bytepow = ctx->tb.bytepow;
            void* args[] = {(void*)&d_cc, (void*)&eb, (void*)&bytepow};
            if (je == cudaSuccess)
                je = cudaLaunchKernel((const void*)jk, dim3((unsigned)((N + EXPR_THREADS - 1) / EXPR_THREADS)), dim3(EXPR_THREADS), args, 0, ctx->stream);
            cudaFreeAsync(d_c, ctx->stream);
            if (je != cudaSuccess) return fail(PIL2GPU_E_CUDA, "calculate_exps (compiled): %s", cudaGetErrorString(je));
            return check_launch(ctx, 1, "calculate_exps");
        }
    }
    const size_t op_bytes = (size_t)n_ops * EXPR_OP_WORDS * sizeof(uint32_t), c_bytes = (size_t)(n_consts ? n_consts : 1) * 3 * sizeof(u64);
    char* scratch = nullptr;
    CU(cudaMallocFromPoolAsync(&scratch, ((op_bytes + 15) & ~(size_t)15) + c_bytes, ctx->pool, ctx->stream));
    u32* d_ops = reinterpret_cast<u32*>(scratch);
    u64* d_consts = reinterpret_cast<u64*>(scratch + ((op_bytes + 15) & ~(size_t)15));
    cudaError_t e = cudaMemcpyAsync(d_ops, ops, op_bytes, cudaMemcpyHostToDevice, ctx->stream);          // pageable sources: staged before return
    if (e == cudaSuccess && n_consts) e = cudaMemcpyAsync(d_consts, consts, (size_t)n_consts * 3 * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) expr_kernel<<<(unsigned)((N + EXPR_THREADS - 1) / EXPR_THREADS), EXPR_THREADS, 0, ctx->stream>>>(d_ops, n_ops, d_consts, eb, (int)domain_bits, x_shift != 0, ctx->tb);
    cudaFreeAsync(scratch, ctx->stream);
    if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "calculate_exps: %s", cudaGetErrorString(e));
    return check_launch(ctx, 1, "calculate_exps");
}

// ---- FRI polynomial: computeFRIStark after the xDivXSubXi table (stark_gen_helpers.js:325-334; friPolinomial.js:26-56) ----
int pil2gpu_fri_pol_dev(pil2gpu_ctx* ctx, const pil2gpu_fri_term* terms, uint32_t n_terms, const uint64_t* evals, const int32_t* openings,
                        uint32_t n_open, const uint64_t* xdiv_dev, const uint64_t vf1[3], const uint64_t vf2[3], uint32_t nBitsExt, uint64_t* f_dev) {
    ENTER(ctx);
    if (!terms || !evals || !openings || !xdiv_dev || !vf1 || !vf2 || !f_dev || n_terms == 0 || n_open == 0)
        return fail(PIL2GPU_E_INVALID, "null or empty argument");
    if (nBitsExt > 32) return fail(PIL2GPU_E_INVALID, "bad sizes");
    const u64 E = 1ULL << nBitsExt;
    // enumeration order of friExps' keys (friPolinomial.js:44): non-negative primes ascending, then the negative ones as first used
    std::vector<int32_t> first_use, order;
    for (uint32_t i = 0; i < n_terms; i++) {
        if (terms[i].dim != 1 && terms[i].dim != 3) return fail(PIL2GPU_E_INVALID, "term %u: dim must be 1 or 3", i);
        if (!terms[i].buf_dev || terms[i].offset + terms[i].dim > terms[i].size) return fail(PIL2GPU_E_RANGE, "term %u: columns outside its buffer", i);
        if (terms[i].size * 8 > 32768) return fail(PIL2GPU_E_UNSUPPORTED, "term %u: rows wider than 4096 columns", i);
        bool seen = false;
        for (int32_t p : first_use) seen |= (p == terms[i].prime);
        if (!seen) first_use.push_back(terms[i].prime);
    }
    for (int32_t p : first_use) if (p >= 0) order.push_back(p);
    for (size_t a = 0; a < order.size(); a++) for (size_t b = a + 1; b < order.size(); b++) if (order[b] < order[a]) std::swap(order[a], order[b]);
    for (int32_t p : first_use) if (p < 0) order.push_back(p);
    const int n_groups = (int)order.size();
    if (n_groups > 4) return fail(PIL2GPU_E_UNSUPPORTED, "more than 4 distinct opening points in the evaluation map");
    FriPolFinish fin;
    memset(&fin, 0, sizeof(fin));
    fin.n_groups = n_groups;
    fin.n_open = (int)n_open;
    std::vector<int> group(n_terms), count(n_groups, 0), rank(n_terms);
    for (int g = 0; g < n_groups; g++) {
        fin.xidx[g] = -1;
        for (uint32_t o = 0; o < n_open; o++) if (openings[o] == order[g]) { fin.xidx[g] = (int)o; break; }
        if (fin.xidx[g] < 0) return fail(PIL2GPU_E_INVALID, "opening %d of the evaluation map is not in openingPoints", order[g]);
    }
    for (uint32_t i = 0; i < n_terms; i++) {
        for (int g = 0; g < n_groups; g++) if (order[g] == terms[i].prime) group[i] = g;
        rank[i] = count[group[i]]++;
    }
    const h3 one = {{1, 0, 0}}, v1 = {{vf1[0] % GL_P, vf1[1] % GL_P, vf1[2] % GL_P}}, v2 = {{vf2[0] % GL_P, vf2[1] % GL_P, vf2[2] % GL_P}};
    int max_count = 0;
    for (int g = 0; g < n_groups; g++) max_count = count[g] > max_count ? count[g] : max_count;
    std::vector<h3> pow2(max_count > 0 ? max_count : 1, one);
    for (int k = 1; k < max_count; k++) pow2[k] = h3_mul(pow2[k - 1], v2);
    std::vector<h3> w(n_terms);
    h3 cg[4] = {{{0, 0, 0}}, {{0, 0, 0}}, {{0, 0, 0}}, {{0, 0, 0}}};
    for (uint32_t i = 0; i < n_terms; i++) {
        w[i] = pow2[count[group[i]] - 1 - rank[i]];                                    // Horner in vf2 (:33-39)
        const h3 ev = {{evals[3 * i] % GL_P, evals[3 * i + 1] % GL_P, evals[3 * i + 2] % GL_P}};
        cg[group[i]] = h3_add(cg[group[i]], h3_mul(w[i], ev));
    }
    h3 ug = one;
    for (int g = n_groups - 1; g >= 0; g--) {                                          // Horner in vf1 (:48-52)
        for (int c = 0; c < 3; c++) { fin.u[g][c] = ug.c[c]; fin.c[g][c] = cg[g].c[c]; }
        ug = h3_mul(ug, v1);
    }
    const int NT = 3 * n_groups;
    // distinct buffers
    struct BufRef { const uint64_t* p; uint64_t size; };
    std::vector<BufRef> bufs;
    for (uint32_t i = 0; i < n_terms; i++) {
        bool seen = false;
        for (const BufRef& b : bufs) seen |= (b.p == terms[i].buf_dev && b.size == terms[i].size);
        if (!seen) bufs.push_back(BufRef{terms[i].buf_dev, terms[i].size});
    }
    u64* S = nullptr;
    CU(cudaMallocFromPoolAsync(&S, E * NT * sizeof(u64), ctx->pool, ctx->stream));
    int launches = 0;
    cudaError_t e = cudaSuccess;
    std::vector<uint2*> bfs;
    for (size_t bi = 0; bi < bufs.size() && e == cudaSuccess; bi++) {
        const u64 size = bufs[bi].size;
        const u32 dsteps = (u32)((size * 8 + 63) / 64);
        std::vector<h3> W((size_t)size * n_groups, h3{{0, 0, 0}});                     // W[col][g]: coefficient of column col in S_g
        for (uint32_t i = 0; i < n_terms; i++) {
            if (terms[i].buf_dev != bufs[bi].p || terms[i].size != size) continue;
            h3 t = w[i];
            for (uint32_t j = 0; j < terms[i].dim; j++) {                              // F3 column (a, b, c) = a + b x + c x^2
                h3& dst = W[(size_t)(terms[i].offset + j) * n_groups + group[i]];
                dst = h3_add(dst, t);
                t = h3_mulx(t);
            }
        }
        // byte limbs of W'[(col, b)][oc] = W[col][oc] * 2^(8b), in fragment order (see fripol.cuh)
        const size_t K = (size_t)dsteps * 64;
        std::vector<unsigned char> Wb(K * NT * 8, 0);
        for (u64 col = 0; col < size; col++)
            for (int oc = 0; oc < NT; oc++) {
                u64 v = W[(size_t)col * n_groups + oc / 3].c[oc % 3];
                for (int b = 0; b < 8; b++) {
                    for (int bp = 0; bp < 8; bp++) Wb[((size_t)col * 8 + b) * NT * 8 + oc * 8 + bp] = (unsigned char)(v >> (8 * bp));
                    v = glh_mul(v, 256);
                }
            }
        const char* fpmode = getenv("PIL2GPU_FRIPOL");
        if (n_groups <= 2 && !(fpmode && strcmp(fpmode, "mma1") == 0)) {
            // second formulation (fripol_mma2_kernel): rows as the B operand, A fragments in (limb pair, oc) order
            const u32 ksteps = (u32)((size + 3) / 4), npieces = (ksteps + 1) / 2;
            std::vector<uint4> AF((size_t)(npieces + 3) * 256, make_uint4(0, 0, 0, 0));
            for (u32 s = 0; s < ksteps; s++)
                for (int i = 0; i < 4; i++)
                    for (int lane = 0; lane < 32; lane++) {
                        const int gid = lane >> 2, tig = lane & 3;
                        const u64 col = (u64)4 * s + tig;
                        if (gid >= NT || col >= size) continue;
                        u32 a[4] = {0, 0, 0, 0};
                        for (int ib = 0; ib < 4; ib++) {
                            const unsigned char* lo = &Wb[((size_t)col * 8 + ib) * NT * 8 + gid * 8];
                            const unsigned char* hi = &Wb[((size_t)col * 8 + 4 + ib) * NT * 8 + gid * 8];
                            a[0] |= (u32)lo[2 * i] << (8 * ib);
                            a[1] |= (u32)lo[2 * i + 1] << (8 * ib);
                            a[2] |= (u32)hi[2 * i] << (8 * ib);
                            a[3] |= (u32)hi[2 * i + 1] << (8 * ib);
                        }
                        AF[((size_t)(s >> 1) * 2 + (s & 1)) * 128 + i * 32 + lane] = make_uint4(a[0], a[1], a[2], a[3]);
                    }
            uint4* dAF = nullptr;
            e = cudaMallocFromPoolAsync(&dAF, AF.size() * sizeof(uint4), ctx->pool, ctx->stream);
            if (e != cudaSuccess) break;
            bfs.push_back(reinterpret_cast<uint2*>(dAF));
            e = cudaMemcpyAsync(dAF, AF.data(), AF.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);   // pageable source: staged before return
            if (e != cudaSuccess) break;
            e = cudaFuncSetAttribute(fripol_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FP2_SMEM);   // per device: set on every launch
            if (e != cudaSuccess) break;
            fripol_mma2_kernel<<<(unsigned)((E + 255) / 256), FP2_THREADS, FP2_SMEM, ctx->stream>>>((const u64*)bufs[bi].p, size, E, dAF, npieces, NT, S, bi > 0);
            launches++;
            continue;
        }
        std::vector<uint2> BF((size_t)dsteps * 2 * NT * 32);
        for (u32 j = 0; j < dsteps; j++)
            for (int h = 0; h < 2; h++)
                for (int nt = 0; nt < NT; nt++)
                    for (int lane = 0; lane < 32; lane++) {
                        const int gid = lane >> 2, tig = lane & 3;
                        u32 b0 = 0, b1 = 0;
                        for (int i = 0; i < 4; i++) {
                            const size_t P0 = (size_t)j * 64 + 16 * tig + 8 * h + i;
                            b0 |= (u32)Wb[P0 * NT * 8 + nt * 8 + gid] << (8 * i);
                            b1 |= (u32)Wb[(P0 + 4) * NT * 8 + nt * 8 + gid] << (8 * i);
                        }
                        BF[((size_t)(2 * j + h) * NT + nt) * 32 + lane] = make_uint2(b0, b1);
                    }
        uint2* dBF = nullptr;
        e = cudaMallocFromPoolAsync(&dBF, BF.size() * sizeof(uint2), ctx->pool, ctx->stream);
        if (e != cudaSuccess) break;
        bfs.push_back(dBF);
        e = cudaMemcpyAsync(dBF, BF.data(), BF.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream);   // pageable source: staged before return
        if (e != cudaSuccess) break;
        const unsigned blocks = (unsigned)((E + FP_WARPS * 32 - 1) / (FP_WARPS * 32));
        const int accum = bi > 0;
        switch (NT) {
            case 3: fripol_mma_kernel<3><<<blocks, FP_WARPS * 32, 0, ctx->stream>>>((const u64*)bufs[bi].p, size, E, dBF, dsteps, S, accum); break;
            case 6: fripol_mma_kernel<6><<<blocks, FP_WARPS * 32, 0, ctx->stream>>>((const u64*)bufs[bi].p, size, E, dBF, dsteps, S, accum); break;
            case 9: fripol_mma_kernel<9><<<blocks, FP_WARPS * 32, 0, ctx->stream>>>((const u64*)bufs[bi].p, size, E, dBF, dsteps, S, accum); break;
            default: fripol_mma_kernel<12><<<blocks, FP_WARPS * 32, 0, ctx->stream>>>((const u64*)bufs[bi].p, size, E, dBF, dsteps, S, accum); break;
        }
        launches++;
    }
    if (e == cudaSuccess) {
        fripol_finish_kernel<<<(unsigned)((E + 255) / 256), 256, 0, ctx->stream>>>(S, (const u64*)xdiv_dev, fin, E, (u64*)f_dev);
        launches++;
    }
    for (uint2* b : bfs) cudaFreeAsync(b, ctx->stream);
    cudaFreeAsync(S, ctx->stream);
    if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "fri_pol: %s", cudaGetErrorString(e));
    return check_launch(ctx, launches, "fri_pol");
}

// Host-buffer form of calculateExps: every buffer is a host array of 2^domain_bits rows; buffers with `read` set are uploaded, the
// program runs on the device, buffers with `written` set come back.  (The device-resident form is what a prover that keeps its extended
// buffers in HBM calls; this one is the drop-in for the reference's host arrays, prover_helpers.js:33-76.)
int pil2gpu_calculate_exps(pil2gpu_ctx* ctx, const uint32_t* ops, uint32_t n_ops, const uint64_t* consts, uint32_t n_consts,
                           const pil2gpu_expr_host_buffer* bufs, uint32_t n_bufs, uint32_t domain_bits, int x_shift) {
    ENTER(ctx);
    if (n_ops == 0) return PIL2GPU_OK;
    if (!bufs && n_bufs) return fail(PIL2GPU_E_INVALID, "null argument");
    if (domain_bits > 32) return fail(PIL2GPU_E_INVALID, "domain of 2^%u rows exceeds the 2-adicity of the field", domain_bits);
    if (n_bufs > EXPR_MAX_BUFS) return fail(PIL2GPU_E_UNSUPPORTED, "more than %d buffers", EXPR_MAX_BUFS);
    const u64 N = 1ULL << domain_bits;
    size_t words = 0;
    for (uint32_t b = 0; b < n_bufs; b++) {
        if (!bufs[b].ptr || bufs[b].row_words == 0) return fail(PIL2GPU_E_INVALID, "buffer %u: null pointer or empty rows", b);
        words += ev2(N * bufs[b].row_words);
    }
    int rc = ensure_ws(ctx, words);
    if (rc) return rc;
    pil2gpu_expr_buffer dev[EXPR_MAX_BUFS];
    size_t off = 0;
    for (uint32_t b = 0; b < n_bufs; b++) {
        dev[b].ptr_dev = ctx->ws + off;
        dev[b].row_words = bufs[b].row_words;
        if (bufs[b].read) CU(cudaMemcpyAsync(dev[b].ptr_dev, bufs[b].ptr, N * bufs[b].row_words * 8, cudaMemcpyHostToDevice, ctx->stream));
        off += ev2(N * bufs[b].row_words);
    }
    rc = pil2gpu_calculate_exps_dev(ctx, ops, n_ops, consts, n_consts, dev, n_bufs, domain_bits, x_shift);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    for (uint32_t b = 0; b < n_bufs; b++)
        if (bufs[b].written) CU(cudaMemcpyAsync(bufs[b].ptr, dev[b].ptr_dev, N * bufs[b].row_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

// The CUDA source the compiled path generates for a program, and whether NVRTC accepts it for sm_100a (no device needed: tests / tooling).
// source_out (may be NULL) receives up to cap - 1 characters + NUL; returns 0 = compiles, > 0 = NVRTC error code (message in
// pil2gpu_last_error), PIL2GPU_E_UNSUPPORTED = no NVRTC on this machine.
int pil2gpu_expr_jit_check(const uint32_t* ops, uint32_t n_ops, const uint64_t* row_words, uint32_t n_bufs, uint32_t domain_bits, int x_shift,
                           char* source_out, uint64_t cap) {
    if (!ops || n_ops == 0 || (!row_words && n_bufs) || n_bufs > EXPR_MAX_BUFS || domain_bits > 32) return fail(PIL2GPU_E_INVALID, "bad argument");
    u64 rw[EXPR_MAX_BUFS] = {0};
    for (uint32_t b = 0; b < n_bufs; b++) rw[b] = row_words[b];
    for (uint32_t k = 0; k < n_ops; k++) {
        const uint32_t* o = ops + (size_t)k * EXPR_OP_WORDS;
        if (o[0] > EXPR_MULADD || o[1] > 3) return fail(PIL2GPU_E_INVALID, "record %u: invalid opcode / source count", k);
        for (uint32_t j = 0; j <= o[1]; j++) {
            const uint32_t* w = j == 0 ? o + 4 : o + 4 + 3 * j;
            const uint32_t kind = w[0] & 255;
            if (kind > EXPR_K_X || (kind == EXPR_K_TMP && w[1] >= EXPR_MAX_SLOTS) || (kind == EXPR_K_BUF && (w[0] >> 16) >= n_bufs))
                return fail(PIL2GPU_E_INVALID, "record %u: invalid operand", k);
        }
    }
    std::string src, log;
    const int rc = expr_jit_compile_only(ops, n_ops, rw, domain_bits, x_shift != 0, &src, &log);
    if (source_out && cap) {
        const size_t n = src.size() < cap - 1 ? src.size() : (size_t)cap - 1;
        memcpy(source_out, src.data(), n);
        source_out[n] = 0;
    }
    if (rc == -1) return fail(PIL2GPU_E_UNSUPPORTED, "%s", log.c_str());
    if (rc != 0) return fail(rc > 0 ? rc : PIL2GPU_E_CUDA, "NVRTC: %s", log.substr(0, 800).c_str());
    return PIL2GPU_OK;
}

// Host-buffer form of the whole of computeFRIStark's arithmetic (:289-334): uploads every distinct extended buffer once, builds
// the xDivXSubXi table on the device and returns f_ext (and the table, if wanted).  terms[i].buf_dev holds HOST pointers here.
int pil2gpu_fri_pol(pil2gpu_ctx* ctx, const pil2gpu_fri_term* terms, uint32_t n_terms, const uint64_t* evals, const int32_t* openings,
                    uint32_t n_open, const uint64_t xi_challenge[3], const uint64_t vf1[3], const uint64_t vf2[3], uint32_t nBits,
                    uint32_t nBitsExt, uint64_t* f_out, uint64_t* xdiv_out) {
    ENTER(ctx);
    if (!terms || !f_out || n_terms == 0 || n_open == 0) return fail(PIL2GPU_E_INVALID, "null or empty argument");
    if (nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad sizes");
    const u64 E = 1ULL << nBitsExt;
    struct BufRef { const uint64_t* host; uint64_t size; size_t off; };
    std::vector<BufRef> bufs;
    size_t words = 0;
    for (uint32_t i = 0; i < n_terms; i++) {
        if (!terms[i].buf_dev) return fail(PIL2GPU_E_INVALID, "term %u: null buffer", i);
        bool seen = false;
        for (const BufRef& b : bufs) seen |= (b.host == terms[i].buf_dev && b.size == terms[i].size);
        if (!seen) { bufs.push_back(BufRef{terms[i].buf_dev, terms[i].size, words}); words += ev2(E * terms[i].size); }
    }
    const size_t xw = ev2((size_t)3 * n_open * E), fw = 3 * E;
    int rc = ensure_ws(ctx, words + xw + fw);
    if (rc) return rc;
    u64 *xd = ctx->ws + words, *f = ctx->ws + words + xw;
    for (const BufRef& b : bufs) CU(cudaMemcpyAsync(ctx->ws + b.off, b.host, E * b.size * 8, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<pil2gpu_fri_term> dterms(terms, terms + n_terms);
    for (uint32_t i = 0; i < n_terms; i++)
        for (const BufRef& b : bufs) if (b.host == terms[i].buf_dev && b.size == terms[i].size) dterms[i].buf_dev = ctx->ws + b.off;
    rc = pil2gpu_x_div_x_sub_xi_dev(ctx, xi_challenge, openings, n_open, nBits, nBitsExt, xd);
    if (rc) return rc;
    rc = pil2gpu_fri_pol_dev(ctx, dterms.data(), n_terms, evals, openings, n_open, xd, vf1, vf2, nBitsExt, f);
    if (rc) return rc;
    CU(cudaMemcpyAsync(f_out, f, fw * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (xdiv_out) CU(cudaMemcpyAsync(xdiv_out, xd, (size_t)3 * n_open * E * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

// ---- host <-> device copies that accept pinned OR pageable host memory ------------------------------------------------
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
static int ensure_stage(pil2gpu_ctx* ctx) {
    if (ctx->stage) return PIL2GPU_OK;
    StageRing* r = new (std::nothrow) StageRing();
    if (!r) return fail(PIL2GPU_E_NOMEM, "out of host memory");
    for (int i = 0; i < STAGE_SLOTS; i++) {
        if (cudaHostAlloc((void**)&r->slot[i], STAGE_SLOT_BYTES, cudaHostAllocPortable) != cudaSuccess ||
            cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            delete r;
            return fail(PIL2GPU_E_NOMEM, "pinned staging ring allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    int n = (int)std::thread::hardware_concurrency();
    n = n >= 16 ? 8 : (n >= 4 ? n / 2 : 1);
    if (const char* e = getenv("PIL2GPU_COPY_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) n = v; }
    r->pool = new (std::nothrow) CopyPool(n);
    if (!r->pool) { delete r; return fail(PIL2GPU_E_NOMEM, "out of host memory"); }
    ctx->stage = r;
    return PIL2GPU_OK;
}
// A copy is cut into chunks that fit one staging slot: a flat copy (both pitches == row_bytes) into byte ranges, a strided one
// into row ranges.
struct StageChunks {
    bool flat;
    size_t row_bytes, rows, per, n;
    StageChunks(size_t row_bytes_, size_t rows_, size_t hpitch, size_t dpitch) : row_bytes(row_bytes_), rows(rows_) {
        flat = (hpitch == row_bytes && dpitch == row_bytes);
        if (flat) { const size_t total = rows * row_bytes; per = STAGE_SLOT_BYTES; n = (total + per - 1) / per; }
        else { per = STAGE_SLOT_BYTES / row_bytes; n = per ? (rows + per - 1) / per : 0; }
    }
    // chunk c: offsets in rows (strided) or bytes (flat), and its extent as (row_bytes, rows)
    void get(size_t c, size_t& first, size_t& rb, size_t& nr) const {
        if (flat) { const size_t total = rows * row_bytes; first = c * per; rb = total - first < per ? total - first : per; nr = 1; }
        else { first = c * per; rb = row_bytes; nr = rows - first < per ? rows - first : per; }
    }
};
// rows x row_bytes, host (pitch hpitch) -> device (pitch dpitch), on stream st.  Pinned host memory: one asynchronous 2-D copy.
// Pageable: through the pinned ring -- the host threads fill slot k+1 while the DMA engine drains slot k.
static int h2d_2d(pil2gpu_ctx* ctx, char* dev, size_t dpitch, const char* host, size_t hpitch, size_t row_bytes, size_t rows, cudaStream_t st) {
    if (rows == 0 || row_bytes == 0) return PIL2GPU_OK;
    if (host_is_pinned(host)) {
        if (dpitch == row_bytes && hpitch == row_bytes) CU(cudaMemcpyAsync(dev, host, rows * row_bytes, cudaMemcpyHostToDevice, st));
        else CU(cudaMemcpy2DAsync(dev, dpitch, host, hpitch, row_bytes, rows, cudaMemcpyHostToDevice, st));
        return PIL2GPU_OK;
    }
    int rc = ensure_stage(ctx);
    if (rc) return rc;
    StageRing* r = ctx->stage;
    if (row_bytes > STAGE_SLOT_BYTES && !(dpitch == row_bytes && hpitch == row_bytes))
        return fail(PIL2GPU_E_UNSUPPORTED, "row segment of %zu bytes exceeds the staging slot", row_bytes);
    const StageChunks ch(row_bytes, rows, hpitch, dpitch);
    for (size_t c = 0; c < ch.n; c++) {
        const int k = (int)(c % STAGE_SLOTS);
        size_t first, rb, nr;
        ch.get(c, first, rb, nr);
        const size_t ho = ch.flat ? first : first * hpitch, dofs = ch.flat ? first : first * dpitch;
        CU(cudaEventSynchronize(r->ev[k]));                       // the last DMA that used this slot is done
        r->pool->copy2d(r->slot[k], rb, host + ho, ch.flat ? rb : hpitch, rb, nr);
        CU(cudaMemcpy2DAsync(dev + dofs, ch.flat ? rb : dpitch, r->slot[k], rb, rb, nr, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(r->ev[k], st));
    }
    return PIL2GPU_OK;
}
// device -> host twin.  Pinned: asynchronous.  Pageable: returns when the data is in the caller's memory.
static int d2h_2d(pil2gpu_ctx* ctx, char* host, size_t hpitch, const char* dev, size_t dpitch, size_t row_bytes, size_t rows, cudaStream_t st) {
    if (rows == 0 || row_bytes == 0) return PIL2GPU_OK;
    if (host_is_pinned(host)) {
        if (dpitch == row_bytes && hpitch == row_bytes) CU(cudaMemcpyAsync(host, dev, rows * row_bytes, cudaMemcpyDeviceToHost, st));
        else CU(cudaMemcpy2DAsync(host, hpitch, dev, dpitch, row_bytes, rows, cudaMemcpyDeviceToHost, st));
        return PIL2GPU_OK;
    }
    int rc = ensure_stage(ctx);
    if (rc) return rc;
    StageRing* r = ctx->stage;
    if (row_bytes > STAGE_SLOT_BYTES && !(dpitch == row_bytes && hpitch == row_bytes))
        return fail(PIL2GPU_E_UNSUPPORTED, "row segment of %zu bytes exceeds the staging slot", row_bytes);
    const StageChunks ch(row_bytes, rows, hpitch, dpitch);
    for (size_t c = 0; c < ch.n + STAGE_SLOTS; c++) {
        size_t first, rb, nr;
        if (c >= STAGE_SLOTS && c - STAGE_SLOTS < ch.n) {          // drain chunk c - SLOTS into the caller's memory
            const size_t j = c - STAGE_SLOTS;
            const int k = (int)(j % STAGE_SLOTS);
            ch.get(j, first, rb, nr);
            CU(cudaEventSynchronize(r->ev[k]));
            r->pool->copy2d(host + (ch.flat ? first : first * hpitch), ch.flat ? rb : hpitch, r->slot[k], rb, rb, nr);
        }
        if (c < ch.n) {
            const int k = (int)(c % STAGE_SLOTS);
            ch.get(c, first, rb, nr);
            if (c < STAGE_SLOTS) CU(cudaEventSynchronize(r->ev[k]));   // a slot last used by an earlier call (later ones were drained above)
            CU(cudaMemcpy2DAsync(r->slot[k], rb, dev + (ch.flat ? first : first * dpitch), ch.flat ? rb : dpitch, rb, nr, cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(r->ev[k], st));
        }
    }
    return PIL2GPU_OK;
}

// A paged host buffer (pilcom BigBuffer = list of BigUint64Array pages): page p holds words[p] u64.
struct PageList {
    const uint64_t* const* pages;
    const uint64_t* words;
    uint32_t n;
};
static int check_pages(const PageList& pl, size_t expect, const char* what) {
    if ((!pl.pages || !pl.words) && pl.n) return fail(PIL2GPU_E_INVALID, "%s: null page list", what);
    size_t off = 0;
    for (uint32_t p = 0; p < pl.n; p++) {
        if (pl.words[p] && !pl.pages[p]) return fail(PIL2GPU_E_INVALID, "%s: page %u is null", what, p);
        if (pl.words[p] > expect - off) return fail(PIL2GPU_E_INVALID, "%s: pages hold more than %zu words", what, expect);
        off += pl.words[p];
    }
    if (off != expect) return fail(PIL2GPU_E_INVALID, "%s: pages hold %zu words, expected %zu", what, off, expect);
    return PIL2GPU_OK;
}
// Gather a paged host buffer into device memory / scatter back, on stream st (see h2d_2d / d2h_2d for pinned vs pageable pages).
static int pages_to_dev(pil2gpu_ctx* ctx, u64* dev, const PageList& pl, size_t expect, cudaStream_t st) {
    int rc = check_pages(pl, expect, "source");
    size_t off = 0;
    for (uint32_t p = 0; p < pl.n && !rc; p++) {
        rc = h2d_2d(ctx, (char*)(dev + off), pl.words[p] * 8, (const char*)pl.pages[p], pl.words[p] * 8, pl.words[p] * 8, pl.words[p] ? 1 : 0, st);
        off += pl.words[p];
    }
    return rc;
}
static int dev_to_pages(pil2gpu_ctx* ctx, const u64* dev, const PageList& pl, size_t expect, cudaStream_t st) {
    int rc = check_pages(pl, expect, "destination");
    size_t off = 0;
    for (uint32_t p = 0; p < pl.n && !rc; p++) {
        rc = d2h_2d(ctx, (char*)pl.pages[p], pl.words[p] * 8, (const char*)(dev + off), pl.words[p] * 8, pl.words[p] * 8, pl.words[p] ? 1 : 0, st);
        off += pl.words[p];
    }
    return rc;
}
// Columns [c0, c0 + w) of every row of a row-major paged buffer (row = nPols words) <-> a dense device slab [rows][w].
// Needs pages that hold whole rows.  All rows of a page go in one strided copy.
static bool pages_row_aligned(const PageList& pl, size_t nPols) {
    for (uint32_t p = 0; p < pl.n; p++) if (pl.words[p] % nPols) return false;
    return true;
}
static int slab_from_pages(pil2gpu_ctx* ctx, u64* slab, const PageList& pl, size_t nPols, size_t c0, size_t w, cudaStream_t st) {
    size_t row = 0;
    for (uint32_t p = 0; p < pl.n; p++) {
        const size_t nr = pl.words[p] / nPols;
        int rc = h2d_2d(ctx, (char*)(slab + row * w), w * 8, (const char*)(pl.pages[p] + c0), nPols * 8, w * 8, nr, st);
        if (rc) return rc;
        row += nr;
    }
    return PIL2GPU_OK;
}
static int slab_to_pages(pil2gpu_ctx* ctx, const u64* slab, const PageList& pl, size_t nPols, size_t c0, size_t w, cudaStream_t st) {
    size_t row = 0;
    for (uint32_t p = 0; p < pl.n; p++) {
        const size_t nr = pl.words[p] / nPols;
        int rc = d2h_2d(ctx, (char*)(const_cast<uint64_t*>(pl.pages[p]) + c0), nPols * 8, (const char*)(slab + row * w), w * 8, w * 8, nr, st);
        if (rc) return rc;
        row += nr;
    }
    return PIL2GPU_OK;
}


struct DevBuf {   // RAII device allocation for the host-pointer entry points
    u64* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t words) { return cudaMalloc(&p, (words ? words : 1) * sizeof(u64)); }
};
struct PoolBuf {  // stream-ordered scratch from the ctx pool: no device-wide synchronisation on allocation or release (cudaMalloc /
    u64* p = nullptr;   // cudaFree cost ~0.1 ms each and serialise against every stream -- measured on the query-opening path)
    pil2gpu_ctx* ctx;
    explicit PoolBuf(pil2gpu_ctx* c) : ctx(c) {}
    ~PoolBuf() { if (p) cudaFreeAsync(p, ctx->stream); }
    cudaError_t alloc(size_t words) { return cudaMallocFromPoolAsync(&p, (words ? words : 1) * sizeof(u64), ctx->pool, ctx->stream); }
};

int pil2gpu_ntt_paged(pil2gpu_ctx* ctx, const uint64_t* const* src_pages, const uint64_t* src_page_words, uint32_t n_src_pages,
                      uint64_t* const* dst_pages, const uint64_t* dst_page_words, uint32_t n_dst_pages, uint64_t nPols, uint32_t nBits,
                      int inverse) {
    ENTER(ctx);
    if (!src_pages || !dst_pages || !src_page_words || !dst_page_words) return fail(PIL2GPU_E_INVALID, "null page list");
    if (nPols == 0) return fail(PIL2GPU_E_INVALID, "nPols must be > 0");
    if (nBits > 32) return fail(PIL2GPU_E_INVALID, "nBits %u exceeds the 2-adicity of the field (32)", nBits);
    const size_t words = (size_t)nPols << nBits;
    const PageList sp = {src_pages, src_page_words, n_src_pages}, dp = {(const uint64_t* const*)dst_pages, dst_page_words, n_dst_pages};
    int rc = check_pages(sp, words, "source");
    if (!rc) rc = check_pages(dp, words, "destination");
    if (!rc) rc = ensure_ws(ctx, 2 * ev2(words));
    if (rc) return rc;
    u64 *a = ctx->ws, *b = ctx->ws + ev2(words);
    rc = pages_to_dev(ctx, a, sp, words, ctx->stream);
    if (!rc) rc = pil2gpu_ntt_dev(ctx, a, b, nPols, nBits, inverse);
    if (!rc) rc = dev_to_pages(ctx, b, dp, words, ctx->stream);
    cudaError_t es = sync_all_streams(ctx);
    if (rc) return rc;
    if (es != cudaSuccess) return fail(PIL2GPU_E_CUDA, "ntt: %s", cudaGetErrorString(es));
    return PIL2GPU_OK;
}

int pil2gpu_ntt(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t* dst, uint64_t nPols, uint32_t nBits, int inverse) {
    if (!src || !dst) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (nBits > 32) return fail(PIL2GPU_E_INVALID, "nBits %u exceeds the 2-adicity of the field (32)", nBits);
    const uint64_t words = nPols << nBits;
    const uint64_t* sp[1] = {src};
    uint64_t* dp[1] = {dst};
    return pil2gpu_ntt_paged(ctx, sp, &words, 1, dp, &words, 1, nPols, nBits, inverse);
}

int pil2gpu_lde(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t* dst, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt) {
    if (!src || !dst) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad LDE shape");
    const uint64_t sw = nPols << nBits, dw = nPols << nBitsExt;
    const uint64_t* sp[1] = {src};
    uint64_t* dp[1] = {dst};
    return pil2gpu_lde_paged(ctx, sp, &sw, 1, dp, &dw, 1, nPols, nBits, nBitsExt);
}

int pil2gpu_lde_paged(pil2gpu_ctx* ctx, const uint64_t* const* src_pages, const uint64_t* src_page_words, uint32_t n_src_pages,
                      uint64_t* const* dst_pages, const uint64_t* dst_page_words, uint32_t n_dst_pages, uint64_t nPols, uint32_t nBits,
                      uint32_t nBitsExt) {
    ENTER(ctx);
    if (!src_pages || !dst_pages || !src_page_words || !dst_page_words) return fail(PIL2GPU_E_INVALID, "null page list");
    if (nPols == 0 || nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad LDE shape");
    const size_t sw = (size_t)nPols << nBits, dw = (size_t)nPols << nBitsExt;
    const PageList sp = {src_pages, src_page_words, n_src_pages}, dp = {(const uint64_t* const*)dst_pages, dst_page_words, n_dst_pages};
    int rc = check_pages(sp, sw, "source");
    if (!rc) rc = check_pages(dp, dw, "destination");
    if (!rc) rc = ensure_ws(ctx, ev2(sw) + dw);
    if (rc) return rc;
    u64 *a = ctx->ws, *b = ctx->ws + ev2(sw);
    rc = pages_to_dev(ctx, a, sp, sw, ctx->stream);
    if (!rc) rc = pil2gpu_lde_dev(ctx, a, b, nPols, nBits, nBitsExt);
    if (!rc) rc = dev_to_pages(ctx, b, dp, dw, ctx->stream);
    cudaError_t es = sync_all_streams(ctx);
    if (rc) return rc;
    if (es != cudaSuccess) return fail(PIL2GPU_E_CUDA, "lde: %s", cudaGetErrorString(es));
    return PIL2GPU_OK;
}

// computeQStark with host buffers (stark_gen_helpers.js:168-208): q_ext up, cmQ_ext and tree.nodes down; the download of cmQ_ext
// runs under the hashing.
int pil2gpu_compute_q_paged(pil2gpu_ctx* ctx, const uint64_t* const* q_pages, const uint64_t* q_page_words, uint32_t n_q_pages, uint64_t qDim,
                            uint64_t qDeg, uint32_t nBits, uint32_t nBitsExt, int split, uint64_t* const* cmq_pages, const uint64_t* cmq_page_words,
                            uint32_t n_cmq_pages, uint64_t* nodes_out, uint64_t root_out[4]) {
    ENTER(ctx);
    if (!q_pages || !q_page_words) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (qDim == 0 || qDeg == 0 || nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad sizes");
    const u64 E = 1ULL << nBitsExt;
    const size_t qw = E * qDim, cw = E * qDim * qDeg, nw = merkle_nnodes_words(E);
    const PageList qp = {q_pages, q_page_words, n_q_pages}, cp = {(const uint64_t* const*)cmq_pages, cmq_page_words, n_cmq_pages};
    const bool want_ext = cmq_pages != nullptr && n_cmq_pages > 0;
    int rc = check_pages(qp, qw, "q_ext");
    if (!rc && want_ext) rc = check_pages(cp, cw, "cmQ_ext");
    if (!rc) rc = ensure_ws(ctx, ev2(qw) + ev2(cw) + nw);
    if (rc) return rc;
    u64 *a = ctx->ws, *b = ctx->ws + ev2(qw), *n = ctx->ws + ev2(qw) + ev2(cw);
    rc = pages_to_dev(ctx, a, qp, qw, ctx->stream);
    if (!rc) rc = pil2gpu_compute_q_dev(ctx, a, qDim, qDeg, nBits, nBitsExt, b);
    for (int once = 0; once < 1 && rc == PIL2GPU_OK; once++) {
        CU_BREAK(cudaEventRecord(ctx->ev, ctx->stream));
        rc = pil2gpu_merkelize_dev(ctx, b, qDim * qDeg, E, split, n);
        if (rc) break;
        if (want_ext) {
            CU_BREAK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev, 0));
            rc = dev_to_pages(ctx, b, cp, cw, ctx->copy_stream);
            if (rc) break;
        }
        if (nodes_out) rc = d2h_2d(ctx, (char*)nodes_out, nw * 8, (const char*)n, nw * 8, nw * 8, 1, ctx->stream);
        if (rc) break;
        if (root_out) CU_BREAK(cudaMemcpyAsync(root_out, n + nw - 4, 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    cudaError_t es = sync_all_streams(ctx);       // both streams drain on every path: the download targets the caller's buffer
    if (rc) return rc;
    if (es != cudaSuccess) return fail(PIL2GPU_E_CUDA, "compute_q: %s", cudaGetErrorString(es));
    return PIL2GPU_OK;
}

int pil2gpu_compute_q(pil2gpu_ctx* ctx, const uint64_t* q_ext, uint64_t qDim, uint64_t qDeg, uint32_t nBits, uint32_t nBitsExt, int split,
                      uint64_t* cmq_ext_out, uint64_t* nodes_out, uint64_t root_out[4]) {
    if (!q_ext) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (qDim == 0 || qDeg == 0 || nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad sizes");
    const uint64_t qw = qDim << nBitsExt, cw = (qDim * qDeg) << nBitsExt;
    const uint64_t* qp[1] = {q_ext};
    uint64_t* cp[1] = {cmq_ext_out};
    return pil2gpu_compute_q_paged(ctx, qp, &qw, 1, qDim, qDeg, nBits, nBitsExt, split, cmq_ext_out ? cp : nullptr, &cw, cmq_ext_out ? 1 : 0, nodes_out,
                                   root_out);
}

// ------------------------------------------------------------------------------------------------------------
// Hashing / Merkle
// ------------------------------------------------------------------------------------------------------------
uint64_t pil2gpu_merkle_nnodes(uint64_t height) { return height == 0 ? 0 : merkle_nnodes_words(height); }
uint32_t pil2gpu_merkle_depth(uint64_t height) { return height == 0 ? 0 : (uint32_t)merkle_depth(height); }

int pil2gpu_poseidon(pil2gpu_ctx* ctx, const uint64_t in12[12], uint64_t out12[12]) {
    ENTER(ctx);
    if (!in12 || !out12) return fail(PIL2GPU_E_INVALID, "null buffer");
    PoolBuf d(ctx);
    CU(d.alloc(24));
    CU(cudaMemcpyAsync(d.p, in12, 96, cudaMemcpyHostToDevice, ctx->stream));
    poseidon_single_kernel<<<1, 32, 0, ctx->stream>>>(d.p, d.p + 12);
    int rc = check_launch(ctx, 1, "poseidon");
    if (rc) return rc;
    CU(cudaMemcpyAsync(out12, d.p + 12, 96, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

static int merkelize_tiles(pil2gpu_ctx* ctx, RowTiles t, uint64_t width, uint64_t height, int split, uint64_t* nodes) {
    u64* scratch = nullptr;
    const u64 sw = (split && width > 4) ? merkle_split_scratch_words(width, height) : 0;
    if (sw) CU(cudaMallocFromPoolAsync(&scratch, sw * sizeof(u64), ctx->pool, ctx->stream));
    int l = merkle_launch(t, width, height, split, (u64*)nodes, scratch, ctx->stream);
    if (scratch) CU(cudaFreeAsync(scratch, ctx->stream));
    return check_launch(ctx, l, "merkelize");
}

int pil2gpu_merkelize_dev(pil2gpu_ctx* ctx, const uint64_t* elems, uint64_t width, uint64_t height, int split, uint64_t* nodes) {
    ENTER(ctx);
    if (!nodes || (!elems && width * height != 0)) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (height == 0) return fail(PIL2GPU_E_INVALID, "height must be > 0");
    RowTiles t = {(const u64*)elems, width ? width : 1, 0};
    return merkelize_tiles(ctx, t, width, height, split, nodes);
}

int pil2gpu_merkelize_tiled_dev(pil2gpu_ctx* ctx, const uint64_t* tiles, uint32_t n_tiles, uint64_t tile_cols, uint64_t tile_stride,
                                uint64_t height, int split, uint64_t* nodes) {
    ENTER(ctx);
    if (!nodes || !tiles) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (height == 0 || n_tiles == 0 || tile_cols == 0) return fail(PIL2GPU_E_INVALID, "bad tile description");
    if (n_tiles > 1 && (tile_cols % 8) != 0) return fail(PIL2GPU_E_UNSUPPORTED, "tile_cols must be a multiple of 8 (got %llu)", (unsigned long long)tile_cols);
    if (n_tiles > 1 && split) {
        const u64 width = tile_cols * n_tiles, batch = merkle_split_batch(width);
        if (batch % 8) return fail(PIL2GPU_E_UNSUPPORTED, "split hash over tiles needs a batch size that is a multiple of 8");
    }
    RowTiles t = {(const u64*)tiles, tile_cols, tile_stride};
    return merkelize_tiles(ctx, t, tile_cols * n_tiles, height, split, nodes);
}

int pil2gpu_merkle_tree_from_digests_dev(pil2gpu_ctx* ctx, uint64_t* nodes, uint64_t height) {
    ENTER(ctx);
    if (!nodes || height == 0) return fail(PIL2GPU_E_INVALID, "bad arguments");
    int l = merkle_launch_tree((u64*)nodes, height, ctx->stream);
    return check_launch(ctx, l, "tree_from_digests");
}

int pil2gpu_linear_hash(pil2gpu_ctx* ctx, const uint64_t* vals, uint64_t width, int split, uint64_t out4[4]) {
    ENTER(ctx);
    if (!out4 || (!vals && width)) return fail(PIL2GPU_E_INVALID, "null buffer");
    PoolBuf e(ctx), n(ctx);
    CU(e.alloc(width));
    CU(n.alloc(8));
    if (width) CU(cudaMemcpyAsync(e.p, vals, width * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = pil2gpu_merkelize_dev(ctx, e.p, width, 1, split, n.p);   // height-1 tree: nodes[0..4) is the leaf digest
    if (rc) return rc;
    CU(cudaMemcpyAsync(out4, n.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

int pil2gpu_merkelize_paged(pil2gpu_ctx* ctx, const uint64_t* const* elem_pages, const uint64_t* page_words, uint32_t n_pages,
                            uint64_t width, uint64_t height, int split, uint64_t* nodes) {
    ENTER(ctx);
    if (!nodes || !elem_pages || !page_words) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (height == 0) return fail(PIL2GPU_E_INVALID, "height must be > 0");
    const size_t ew = (size_t)width * height, nw = merkle_nnodes_words(height);
    const PageList ep = {elem_pages, page_words, n_pages};
    int rc = check_pages(ep, ew, "elements");
    if (!rc) rc = ensure_ws(ctx, ev2(ew) + nw);
    if (rc) return rc;
    u64 *e = ctx->ws, *n = ctx->ws + ev2(ew);
    rc = pages_to_dev(ctx, e, ep, ew, ctx->stream);
    if (!rc) rc = pil2gpu_merkelize_dev(ctx, e, width, height, split, n);
    if (!rc) rc = d2h_2d(ctx, (char*)nodes, nw * 8, (const char*)n, nw * 8, nw * 8, 1, ctx->stream);
    cudaError_t es = sync_all_streams(ctx);
    if (rc) return rc;
    if (es != cudaSuccess) return fail(PIL2GPU_E_CUDA, "merkelize: %s", cudaGetErrorString(es));
    return PIL2GPU_OK;
}

int pil2gpu_merkelize(pil2gpu_ctx* ctx, const uint64_t* elems, uint64_t width, uint64_t height, int split, uint64_t* nodes) {
    const uint64_t ew = width * height;
    const uint64_t* pages[1] = {elems};
    if (!elems && ew) return fail(PIL2GPU_E_INVALID, "null buffer");
    return pil2gpu_merkelize_paged(ctx, pages, &ew, 1, width, height, split, nodes);
}

// ------------------------------------------------------------------------------------------------------------
// Trees / commit
// ------------------------------------------------------------------------------------------------------------
static pil2gpu_tree* new_tree() {
    pil2gpu_tree* t = new (std::nothrow) pil2gpu_tree();
    if (t) { t->elems = t->nodes = nullptr; t->width = t->height = 0; t->tile_cols = t->tile_stride = 0; t->own_elems = t->own_nodes = false; }
    return t;
}

void pil2gpu_tree_free(pil2gpu_ctx* ctx, pil2gpu_tree* t) {
    if (!t) return;
    if (ctx) {
        DeviceGuard guard(ctx->device);
        sync_all_streams(ctx);
        if (t->own_elems && t->elems) cudaFree(t->elems);
        if (t->own_nodes && t->nodes) cudaFree(t->nodes);
    }
    delete t;
}

int pil2gpu_tree_wrap_dev(pil2gpu_ctx* ctx, const uint64_t* elems_dev, const uint64_t* nodes_dev, uint64_t width, uint64_t height,
                          pil2gpu_tree** tree_out) {
    if (!ctx || !tree_out || !nodes_dev || height == 0) return fail(PIL2GPU_E_INVALID, "bad tree description");
    pil2gpu_tree* t = new_tree();
    if (!t) return fail(PIL2GPU_E_NOMEM, "out of host memory");
    t->elems = (u64*)elems_dev;
    t->nodes = (u64*)nodes_dev;
    t->width = width;
    t->height = height;
    t->tile_cols = width ? width : 1;
    *tree_out = t;
    return PIL2GPU_OK;
}

int pil2gpu_tree_wrap_tiled_dev(pil2gpu_ctx* ctx, const uint64_t* tiles_dev, uint32_t n_tiles, uint64_t tile_cols, uint64_t tile_stride,
                                const uint64_t* nodes_dev, uint64_t height, pil2gpu_tree** tree_out) {
    int rc = pil2gpu_tree_wrap_dev(ctx, tiles_dev, nodes_dev, tile_cols * n_tiles, height, tree_out);
    if (rc) return rc;
    (*tree_out)->tile_cols = tile_cols;
    (*tree_out)->tile_stride = tile_stride;
    return PIL2GPU_OK;
}

int pil2gpu_tree_width(const pil2gpu_tree* t, uint64_t* width, uint64_t* height) {
    if (!t) return fail(PIL2GPU_E_INVALID, "null tree");
    if (width) *width = t->width;
    if (height) *height = t->height;
    return PIL2GPU_OK;
}
const uint64_t* pil2gpu_tree_elements_dev(const pil2gpu_tree* t) { return t ? (const uint64_t*)t->elems : nullptr; }
const uint64_t* pil2gpu_tree_nodes_dev(const pil2gpu_tree* t) { return t ? (const uint64_t*)t->nodes : nullptr; }

int pil2gpu_tree_root(pil2gpu_ctx* ctx, const pil2gpu_tree* t, uint64_t root_out[4]) {
    ENTER(ctx);
    if (!t || !root_out) return fail(PIL2GPU_E_INVALID, "null argument");
    const u64 nw = merkle_nnodes_words(t->height);
    CU(cudaMemcpyAsync(root_out, t->nodes + nw - 4, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

static int commit_common(pil2gpu_ctx* ctx, u64* src_dev_owned, const u64* src_dev, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt,
                         int split, pil2gpu_tree** tree_out, uint64_t root_out[4]) {
    (void)src_dev_owned;
    pil2gpu_tree* t = new_tree();
    if (!t) return fail(PIL2GPU_E_NOMEM, "out of host memory");
    t->width = nPols;
    t->tile_cols = nPols;
    t->height = 1ULL << nBitsExt;
    t->own_elems = t->own_nodes = true;
    cudaError_t e = cudaMalloc(&t->elems, ((size_t)nPols << nBitsExt) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&t->nodes, merkle_nnodes_words(t->height) * 8);
    if (e != cudaSuccess) {
        pil2gpu_tree_free(ctx, t);
        return fail(PIL2GPU_E_NOMEM, "device allocation for the committed tree failed: %s", cudaGetErrorString(e));
    }
    int rc = pil2gpu_lde_dev(ctx, src_dev, t->elems, nPols, nBits, nBitsExt);
    if (!rc) rc = pil2gpu_merkelize_dev(ctx, t->elems, nPols, t->height, split, t->nodes);
    if (!rc && root_out) rc = pil2gpu_tree_root(ctx, t, root_out);
    if (rc) { pil2gpu_tree_free(ctx, t); return rc; }
    *tree_out = t;
    return PIL2GPU_OK;
}

int pil2gpu_commit_dev(pil2gpu_ctx* ctx, const uint64_t* src_dev, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split,
                       pil2gpu_tree** tree_out, uint64_t root_out[4]) {
    ENTER(ctx);
    if (!src_dev || !tree_out) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nPols == 0 || nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad commit shape");
    return commit_common(ctx, nullptr, (const u64*)src_dev, nPols, nBits, nBitsExt, split, tree_out, root_out);
}

int pil2gpu_commit(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split,
                   pil2gpu_tree** tree_out, uint64_t root_out[4]) {
    ENTER(ctx);
    if (!src || !tree_out) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nPols == 0 || nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad commit shape");
    DevBuf a;
    const size_t sw = (size_t)nPols << nBits;
    CU(a.alloc(sw));
    CU(cudaMemcpyAsync(a.p, src, sw * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = commit_common(ctx, nullptr, a.p, nPols, nBits, nBitsExt, split, tree_out, root_out);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

// Column-slab pipeline behind pil2gpu_extend_and_merkelize: the trace is cut into slabs of PIPE_COLS columns (columns are
// independent NTTs and the standard linear hash absorbs columns left to right), and three streams overlap
//     H2D of slab s+1   |   LDE + sponge absorption of slab s   |   D2H of the extended slab s-1
// so the call costs max(PCIe down, PCIe up, compute) instead of their sum.  Strided 2D copies of >= 256-byte row
// segments run at the full PCIe rate (measured 55.6 / 57.2 GB/s up / down, 98.7 GB/s both ways: tools/probe/pcie_probe.cu).

static int extend_and_merkelize_pipelined(pil2gpu_ctx* ctx, const PageList& src, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, uint64_t cs,
                                          const PageList* dst_out, uint64_t* nodes_out, uint64_t root_out[4]) {
    const u64 N = 1ULL << nBits, E = 1ULL << nBitsExt;
    const size_t nw = merkle_nnodes_words(E);
    // Slab widths.  The download is the longest leg, and a strided device->host copy only reaches the full PCIe rate with
    // row segments of >= 512 bytes (measured: 256-byte segments 44-50 GB/s, 512-byte segments 57 GB/s), so wide traces use
    // 64-column slabs -- after two 32-column ones that get the first download going early.
    std::vector<u64> col0, width;
    const bool wide = dst_out && cs == 32 && nPols % 64 == 0 && nPols >= 256;
    for (u64 c = 0; c < nPols;) {
        const u64 w = (wide && c >= 64) ? 64 : cs;
        col0.push_back(c); width.push_back(w);
        c += w;
    }
    const u64 nslabs = col0.size(), wmax = wide ? 64 : cs;
    // workspace: the whole trace (so the upload runs ahead at full rate and then leaves PCIe to the download, which is the
    // longer of the two), two extended slabs, the sponge states and the nodes
    int wrc = ensure_ws(ctx, N * nPols + 2 * E * wmax + 4 * E + nw);
    if (wrc) return wrc;
    struct { u64* p; } sall = {ctx->ws}, dbuf[2] = {{ctx->ws + N * nPols}, {ctx->ws + N * nPols + E * wmax}}, state = {ctx->ws + N * nPols + 2 * E * wmax},
                       nodes = {ctx->ws + N * nPols + 2 * E * wmax + 4 * E};
    EventPool pool;
    const bool trace = getenv("PIL2GPU_TRACE") != nullptr;   // diagnostic: print the slab timeline after the call
    pool.timing = trace;
    std::vector<cudaEvent_t> ev_in(nslabs), ev_lde(nslabs), ev_out(nslabs), ev_abs(nslabs);
    for (u64 s = 0; s < nslabs; s++) { CU(pool.make(&ev_in[s])); CU(pool.make(&ev_lde[s])); CU(pool.make(&ev_out[s])); CU(pool.make(&ev_abs[s])); }
    cudaEvent_t ev_start;
    CU(pool.make(&ev_start));
    CU(cudaEventRecord(ev_start, ctx->stream));                // the copy streams must not run ahead of prior work on ctx->stream
    CU(cudaStreamWaitEvent(ctx->in_stream, ev_start, 0));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ev_start, 0));
    const unsigned blocks = (unsigned)((E + MERKLE_THREADS - 1) / MERKLE_THREADS);
    // Enqueue order per slab: LDE(s), absorb(s), upload(s+1), download(s).  With pinned pages every call is asynchronous and the
    // three streams overlap freely; with pageable pages the staged copies block this thread, and this order keeps the GPU busy
    // with slab s while slab s+1 is staged up and lets the download of slab s run under the LDE of slab s+1.
    int rc = slab_from_pages(ctx, sall.p, src, nPols, col0[0], width[0], ctx->in_stream);
    for (int once = 0; once < 1 && rc == PIL2GPU_OK; once++) CU_BREAK(cudaEventRecord(ev_in[0], ctx->in_stream));
    for (u64 s = 0; s < nslabs && rc == PIL2GPU_OK; s++) {
        const int b = (int)(s & 1);
        const u64 w = width[s];
        u64* sslab = sall.p + N * col0[s];
        CU_BREAK(cudaStreamWaitEvent(ctx->stream, ev_in[s], 0));
        if (s >= 2 && dst_out) CU_BREAK(cudaStreamWaitEvent(ctx->stream, ev_out[s - 2], 0));   // dbuf[b] was downloaded
        rc = pil2gpu_lde_dev(ctx, sslab, dbuf[b].p, w, nBits, nBitsExt);
        if (rc) break;
        CU_BREAK(cudaEventRecord(ev_lde[s], ctx->stream));
        merkle_absorb_kernel<<<blocks, MERKLE_THREADS, 0, ctx->stream>>>(dbuf[b].p, w, E, state.p, s == 0, s + 1 == nslabs, nodes.p);
        rc = check_launch(ctx, 1, "absorb");
        if (rc) break;
        if (trace) CU_BREAK(cudaEventRecord(ev_abs[s], ctx->stream));
        if (s + 1 < nslabs) {
            rc = slab_from_pages(ctx, sall.p + N * col0[s + 1], src, nPols, col0[s + 1], width[s + 1], ctx->in_stream);
            if (rc) break;
            CU_BREAK(cudaEventRecord(ev_in[s + 1], ctx->in_stream));
        }
        if (dst_out) {
            CU_BREAK(cudaStreamWaitEvent(ctx->copy_stream, ev_lde[s], 0));
            rc = slab_to_pages(ctx, dbuf[b].p, *dst_out, nPols, col0[s], w, ctx->copy_stream);
            if (rc) break;
            CU_BREAK(cudaEventRecord(ev_out[s], ctx->copy_stream));
        }
    }
    if (rc == PIL2GPU_OK) {
        int l = merkle_launch_tree(nodes.p, E, ctx->stream);
        rc = check_launch(ctx, l, "tree");
    }
    if (rc == PIL2GPU_OK && nodes_out) rc = d2h_2d(ctx, (char*)nodes_out, nw * 8, (const char*)nodes.p, nw * 8, nw * 8, 1, ctx->stream);
    for (int once = 0; once < 1 && rc == PIL2GPU_OK; once++)
        if (root_out) CU_BREAK(cudaMemcpyAsync(root_out, nodes.p + nw - 4, 32, cudaMemcpyDeviceToHost, ctx->stream));
    // every stream must drain before the workspace is reused, on the error paths too
    cudaError_t e1 = cudaStreamSynchronize(ctx->in_stream), e2 = cudaStreamSynchronize(ctx->stream), e3 = cudaStreamSynchronize(ctx->copy_stream);
    if (rc) return rc;
    if (trace && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess) {
        for (u64 s = 0; s < nslabs; s++) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a, ev_start, ev_in[s]); cudaEventElapsedTime(&b, ev_start, ev_lde[s]); cudaEventElapsedTime(&c, ev_start, ev_abs[s]);
            if (dst_out) cudaEventElapsedTime(&d, ev_start, ev_out[s]);
            fprintf(stderr, "[pil2gpu] slab %llu (%llu cols): h2d done %.1f ms, lde done %.1f, absorb done %.1f, d2h done %.1f\n", (unsigned long long)s,
                    (unsigned long long)width[s], a, b, c, d);
        }
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(PIL2GPU_E_CUDA, "pipelined commit failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return PIL2GPU_OK;
}

// extendAndMerkelize over BigBuffer pages (stark_gen_helpers.js:388-412 with cm<stage>_n / cm<stage>_ext as pilcom BigBuffers).
//  - pages that hold whole rows (any power-of-two nPols): the column-slab pipeline above, one strided copy per (slab, page);
//  - anything else: pages -> device, LDE, then the page downloads run under the hashing.
int pil2gpu_extend_and_merkelize_paged(pil2gpu_ctx* ctx, const uint64_t* const* src_pages, const uint64_t* src_page_words, uint32_t n_src_pages,
                                       uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split, uint64_t* const* dst_pages,
                                       const uint64_t* dst_page_words, uint32_t n_dst_pages, uint64_t* nodes_out, uint64_t root_out[4]) {
    ENTER(ctx);
    if (!src_pages || !src_page_words) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nPols == 0 || nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad commit shape");
    if (nBitsExt - nBits > 8) return fail(PIL2GPU_E_UNSUPPORTED, "blowup 2^%u not supported (max 2^8)", nBitsExt - nBits);
    const size_t sw = (size_t)nPols << nBits, dw = (size_t)nPols << nBitsExt;
    const PageList sp = {src_pages, src_page_words, n_src_pages};
    const PageList dp = {(const uint64_t* const*)dst_pages, dst_page_words, n_dst_pages};
    const bool want_dst = dst_pages != nullptr && n_dst_pages > 0;
    int rc = check_pages(sp, sw, "source");
    if (!rc && want_dst) rc = check_pages(dp, dw, "destination");
    if (rc) return rc;
    // slab pipeline: standard hash only (split batches do not align with slabs), wide traces only, whole rows per page
    const uint64_t cs = (nPols % 32 == 0) ? 32 : 16;
    if (!split && nPols % 16 == 0 && nPols / cs >= 4 && nBits >= 12 && pages_row_aligned(sp, nPols) && (!want_dst || pages_row_aligned(dp, nPols)))
        return extend_and_merkelize_pipelined(ctx, sp, nPols, nBits, nBitsExt, cs, want_dst ? &dp : nullptr, nodes_out, root_out);
    const u64 height = 1ULL << nBitsExt;
    const size_t nw = merkle_nnodes_words(height);
    rc = ensure_ws(ctx, ev2(sw) + ev2(dw) + nw);
    if (rc) return rc;
    u64 *a = ctx->ws, *b = ctx->ws + ev2(sw), *n = ctx->ws + ev2(sw) + ev2(dw);
    rc = pages_to_dev(ctx, a, sp, sw, ctx->stream);
    if (!rc) rc = pil2gpu_lde_dev(ctx, a, b, nPols, nBits, nBitsExt);
    for (int once = 0; once < 1 && rc == PIL2GPU_OK; once++) {
        // hash first (asynchronous), then download the extended buffer on the copy stream: with pageable pages the staged
        // download blocks this thread while the main stream hashes
        CU_BREAK(cudaEventRecord(ctx->ev, ctx->stream));
        rc = pil2gpu_merkelize_dev(ctx, b, nPols, height, split, n);
        if (rc) break;
        if (want_dst) {
            CU_BREAK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev, 0));
            rc = dev_to_pages(ctx, b, dp, dw, ctx->copy_stream);
            if (rc) break;
        }
        if (nodes_out) rc = d2h_2d(ctx, (char*)nodes_out, nw * 8, (const char*)n, nw * 8, nw * 8, 1, ctx->stream);
        if (rc) break;
        if (root_out) CU_BREAK(cudaMemcpyAsync(root_out, n + nw - 4, 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    cudaError_t es = sync_all_streams(ctx);
    if (rc) return rc;
    if (es != cudaSuccess) return fail(PIL2GPU_E_CUDA, "extend_and_merkelize: %s", cudaGetErrorString(es));
    return PIL2GPU_OK;
}

int pil2gpu_extend_and_merkelize(pil2gpu_ctx* ctx, const uint64_t* src, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt, int split,
                                 uint64_t* dst_out, uint64_t* nodes_out, uint64_t root_out[4]) {
    if (!src) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nPols == 0 || nBitsExt > 32 || nBitsExt < nBits) return fail(PIL2GPU_E_INVALID, "bad commit shape");
    const uint64_t sw = nPols << nBits, dw = nPols << nBitsExt;
    const uint64_t* sp[1] = {src};
    uint64_t* dp[1] = {dst_out};
    return pil2gpu_extend_and_merkelize_paged(ctx, sp, &sw, 1, nPols, nBits, nBitsExt, split, dst_out ? dp : nullptr, &dw, dst_out ? 1 : 0, nodes_out,
                                              root_out);
}

// ------------------------------------------------------------------------------------------------------------
// Bench / test utilities
// ------------------------------------------------------------------------------------------------------------
int pil2gpu_synth_dev(pil2gpu_ctx* ctx, uint64_t* dst_dev, uint64_t n_words, uint64_t seed, uint64_t first_index) {
    ENTER(ctx);
    if (!dst_dev && n_words) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (n_words == 0) return PIL2GPU_OK;
    synth_kernel<<<148 * 16, 256, 0, ctx->stream>>>((u64*)dst_dev, n_words, seed, first_index);
    return check_launch(ctx, 1, "synth");
}

int pil2gpu_synth2d_dev(pil2gpu_ctx* ctx, uint64_t* dst_dev, uint64_t rows, uint64_t cols, uint64_t row_stride, uint64_t col0, uint64_t seed) {
    ENTER(ctx);
    if (!dst_dev && rows * cols != 0) return fail(PIL2GPU_E_INVALID, "null buffer");
    if (rows * cols == 0) return PIL2GPU_OK;
    synth2d_kernel<<<148 * 16, 256, 0, ctx->stream>>>((u64*)dst_dev, rows, cols, row_stride, col0, seed);
    return check_launch(ctx, 1, "synth2d");
}

int pil2gpu_bench_int_pipes(pil2gpu_ctx* ctx, double* mulmod_per_s, double* imad_wide_per_s) {
    ENTER(ctx);
    const int blocks = 148 * 8, threads = 256, iters = 2048;
    DevBuf o;
    CU(o.alloc((size_t)blocks * threads));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double res[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CU(cudaEventRecord(e0, ctx->stream));
            if (mode == 0) pipe_probe_kernel<0><<<blocks, threads, 0, ctx->stream>>>(o.p, iters);
            else pipe_probe_kernel<1><<<blocks, threads, 0, ctx->stream>>>(o.p, iters);
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
            ctx->launches++;
        }
        res[mode] = (double)blocks * threads * iters * 32.0 / (best * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (mulmod_per_s) *mulmod_per_s = res[0];
    if (imad_wide_per_s) *imad_wide_per_s = res[1];
    return PIL2GPU_OK;
}

int pil2gpu_tree_from_host(pil2gpu_ctx* ctx, const uint64_t* elems, uint64_t width, uint64_t height, int split, pil2gpu_tree** tree_out) {
    ENTER(ctx);
    if (!tree_out || height == 0 || (!elems && width)) return fail(PIL2GPU_E_INVALID, "bad tree description");
    pil2gpu_tree* t = new_tree();
    if (!t) return fail(PIL2GPU_E_NOMEM, "out of host memory");
    t->width = width;
    t->tile_cols = width ? width : 1;
    t->height = height;
    t->own_elems = t->own_nodes = true;
    cudaError_t e = cudaMalloc(&t->elems, (width * height != 0 ? width * height : 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&t->nodes, merkle_nnodes_words(height) * 8);
    if (e != cudaSuccess) { pil2gpu_tree_free(ctx, t); return fail(PIL2GPU_E_NOMEM, "device allocation failed: %s", cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(t->elems, elems, width * height * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { pil2gpu_tree_free(ctx, t); return fail(PIL2GPU_E_CUDA, "upload failed: %s", cudaGetErrorString(e)); }
    int rc = pil2gpu_merkelize_dev(ctx, t->elems, width, height, split, t->nodes);
    if (!rc) { e = cudaStreamSynchronize(ctx->stream); if (e != cudaSuccess) rc = fail(PIL2GPU_E_CUDA, "merkelize failed: %s", cudaGetErrorString(e)); }
    if (rc) { pil2gpu_tree_free(ctx, t); return rc; }
    *tree_out = t;
    return PIL2GPU_OK;
}

// A tree whose nodes are already known (read back from a const-tree file, merklehash_p.js:249-278): device allocation only.
// The caller fills it piecewise with pil2gpu_tree_fill (so that a 30+ GiB file can stream through a small host buffer).
int pil2gpu_tree_alloc(pil2gpu_ctx* ctx, uint64_t width, uint64_t height, pil2gpu_tree** tree_out) {
    ENTER(ctx);
    if (!tree_out || height == 0) return fail(PIL2GPU_E_INVALID, "bad tree description");
    pil2gpu_tree* t = new_tree();
    if (!t) return fail(PIL2GPU_E_NOMEM, "out of host memory");
    t->width = width;
    t->tile_cols = width ? width : 1;
    t->height = height;
    t->own_elems = t->own_nodes = true;
    cudaError_t e = cudaMalloc(&t->elems, (width * height != 0 ? width * height : 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&t->nodes, merkle_nnodes_words(height) * 8);
    if (e != cudaSuccess) { pil2gpu_tree_free(ctx, t); return fail(PIL2GPU_E_NOMEM, "device allocation failed: %s", cudaGetErrorString(e)); }
    *tree_out = t;
    return PIL2GPU_OK;
}
// which: 0 = elements, 1 = nodes; copies n_words host words to word offset `offset` of that array (synchronous).
int pil2gpu_tree_fill(pil2gpu_ctx* ctx, pil2gpu_tree* t, int which, uint64_t offset, const uint64_t* src, uint64_t n_words) {
    ENTER(ctx);
    if (!t || (!src && n_words)) return fail(PIL2GPU_E_INVALID, "null argument");
    const u64 cap = which == 0 ? t->width * t->height : merkle_nnodes_words(t->height);
    if (which != 0 && which != 1) return fail(PIL2GPU_E_INVALID, "which must be 0 (elements) or 1 (nodes)");
    if (offset > cap || n_words > cap - offset) return fail(PIL2GPU_E_RANGE, "fill of %llu words at %llu exceeds %llu", (unsigned long long)n_words,
                                                            (unsigned long long)offset, (unsigned long long)cap);
    if (n_words == 0) return PIL2GPU_OK;
    u64* dst = (which == 0 ? t->elems : t->nodes) + offset;
    CU(cudaMemcpyAsync(dst, src, n_words * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

int pil2gpu_tree_group_proofs(pil2gpu_ctx* ctx, const pil2gpu_tree* t, const uint64_t* idxs, uint32_t n_idx, uint64_t* rows_out,
                              uint64_t* siblings_out) {
    ENTER(ctx);
    if (!t || !idxs || !rows_out || !siblings_out) return fail(PIL2GPU_E_INVALID, "null argument");
    if (!t->elems) return fail(PIL2GPU_E_INVALID, "tree has no device-resident elements");
    for (uint32_t i = 0; i < n_idx; i++)
        if (idxs[i] >= t->height) return fail(PIL2GPU_E_RANGE, "Out of range");
    if (n_idx == 0) return PIL2GPU_OK;
    const int depth = merkle_depth(t->height);
    PoolBuf di(ctx), dr(ctx), ds(ctx);
    CU(di.alloc(n_idx));
    CU(dr.alloc((size_t)n_idx * t->width));
    CU(ds.alloc((size_t)n_idx * depth * 4));
    CU(cudaMemcpyAsync(di.p, idxs, (size_t)n_idx * 8, cudaMemcpyHostToDevice, ctx->stream));
    RowTiles rt = {t->elems, t->tile_cols ? t->tile_cols : 1, t->tile_stride};
    merkle_group_proof_kernel<<<n_idx, 128, 0, ctx->stream>>>(rt, t->nodes, t->width, t->height, di.p, depth, dr.p, ds.p);
    int rc = check_launch(ctx, 1, "group_proofs");
    if (rc) return rc;
    if (t->width) CU(cudaMemcpyAsync(rows_out, dr.p, (size_t)n_idx * t->width * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (depth) CU(cudaMemcpyAsync(siblings_out, ds.p, (size_t)n_idx * depth * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

int pil2gpu_tree_group_proofs_dev(pil2gpu_ctx* ctx, const pil2gpu_tree* t, const uint64_t* idxs_dev, uint32_t n_idx, uint64_t* rows_out_dev,
                                  uint64_t* siblings_out_dev) {
    ENTER(ctx);
    if (!t || !idxs_dev || !rows_out_dev || !siblings_out_dev) return fail(PIL2GPU_E_INVALID, "null argument");
    if (!t->elems) return fail(PIL2GPU_E_INVALID, "tree has no device-resident elements");
    if (n_idx == 0) return PIL2GPU_OK;
    RowTiles rt = {t->elems, t->tile_cols ? t->tile_cols : 1, t->tile_stride};
    merkle_group_proof_kernel<<<n_idx, 128, 0, ctx->stream>>>(rt, t->nodes, t->width, t->height, (const u64*)idxs_dev, merkle_depth(t->height),
                                                              (u64*)rows_out_dev, (u64*)siblings_out_dev);
    return check_launch(ctx, 1, "group_proofs_dev");
}

int pil2gpu_tree_download(pil2gpu_ctx* ctx, const pil2gpu_tree* t, uint64_t* elems_out, uint64_t* nodes_out) {
    ENTER(ctx);
    if (!t) return fail(PIL2GPU_E_INVALID, "null tree");
    if (elems_out && t->elems && t->tile_cols != t->width) return fail(PIL2GPU_E_UNSUPPORTED, "download of a column-tiled tree is not supported");
    if (elems_out && t->elems) CU(cudaMemcpyAsync(elems_out, t->elems, t->width * t->height * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nodes_out) CU(cudaMemcpyAsync(nodes_out, t->nodes, merkle_nnodes_words(t->height) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PIL2GPU_OK;
}

// ------------------------------------------------------------------------------------------------------------
// FRI
// ------------------------------------------------------------------------------------------------------------
// One kernel folds by at most 2^FRI_MAX_FOLD_BITS (the interpolant of a group lives in registers).  Wider folds -- fri.js:38-41 puts
// no bound on steps[s-1].nBits - steps[s].nBits -- are split: the fold evaluates the group's interpolant by radix-2 halvings with
// beta, beta^2, beta^4, ... (fri.cuh), and the first a halvings of the group of output g are exactly the standard fold by 2^a of
// the whole polynomial (same pairs (j, j + 2^(prev-1)), same twiddles shift_inv * w_prev^-(g + j 2^cur)), so
//     fold_{2^(a+b)}(P, alpha) = fold_{2^b}( fold_{2^a}(P, alpha), alpha^(2^a) )        with the standard parameters of each size.
static int fri_fold_one(pil2gpu_ctx* ctx, const u64* pol, uint32_t prevBits, uint32_t curBits, int32_t nextBits, uint32_t step0Bits, const u64 challenge[3],
                        int split, u64* pol_out, u64* rows_out, u64* nodes_out) {
    FriParams P;
    P.prev_bits = (int)prevBits;
    P.cur_bits = (int)curBits;
    P.next_bits = nextBits;
    P.write_rows = nextBits >= 0;
    P.in_rows = 0;
    P.write_pol = 1;
    P.row0 = 0;
    P.n_rows = 0;
    u64 si = glh_inv(GL_SHIFT);                                    // fri.js:31-36
    for (uint32_t j = 0; j < step0Bits - prevBits; j++) si = glh_mul(si, si);
    P.shift_inv = si;
    P.nx_inv = glh_inv(1ULL << (prevBits - curBits));
    for (int k = 0; k < 3; k++) P.challenge[k] = challenge[k];
    const u64 gs = nextBits >= 0 ? (1ULL << (curBits - nextBits)) : 1;
    // The in-kernel leaf hash runs on FRI_ROWS_PER_CTA threads of each CTA: worth it only for small layers, where it saves
    // a launch; big layers (the first FRI tree has 2^20 leaves at cfg3) go through the full-width leaf kernel instead.
    P.fuse_leaf_hash = (nextBits >= 0) && nodes_out && (!split || 3 * gs <= 4) && (gs * 3 * 8 * FRI_ROWS_PER_CTA <= 160 * 1024) && nextBits <= 12;
    int l = fri_launch_fold(pol, pol_out, rows_out, nodes_out, P, ctx->tb, ctx->stream);
    int rc = check_launch(ctx, l, "fri_fold");
    if (rc) return rc;
    if (nextBits >= 0 && nodes_out) {
        const u64 height = 1ULL << nextBits;
        if (P.fuse_leaf_hash) {
            l = merkle_launch_tree(nodes_out, height, ctx->stream);
            rc = check_launch(ctx, l, "fri_tree");
        } else {
            rc = pil2gpu_merkelize_dev(ctx, rows_out, 3 * gs, height, split, nodes_out);
        }
    }
    return rc;
}

int pil2gpu_fri_fold_dev(pil2gpu_ctx* ctx, const uint64_t* pol, uint32_t prevBits, uint32_t curBits, int32_t nextBits, uint32_t step0Bits,
                         const uint64_t challenge[3], int split, uint64_t* pol_out, uint64_t* rows_out, uint64_t* nodes_out) {
    ENTER(ctx);
    if (!pol || !pol_out || !challenge) return fail(PIL2GPU_E_INVALID, "null argument");
    if (curBits > prevBits || prevBits > step0Bits || step0Bits > 32) return fail(PIL2GPU_E_INVALID, "bad FRI step sizes");
    if (nextBits >= 0 && ((uint32_t)nextBits > curBits || !rows_out)) return fail(PIL2GPU_E_INVALID, "bad next-layer description");
    {   // aliasing: the identity step (fri.js:48-49) may run in place (pol_out == pol: every thread rewrites the three words it
        // read); any other overlap of the output with the input is a cross-CTA race and is rejected
        const u64 *a = (const u64*)pol, *b = (const u64*)pol_out;
        const bool overlap = a < b + ((size_t)3 << curBits) && b < a + ((size_t)3 << prevBits);
        if (overlap && !(prevBits == curBits && a == b)) return fail(PIL2GPU_E_INVALID, "pol_out overlaps pol (only the identity step may run in place)");
    }
    u64 ch[3] = {challenge[0] % GL_P, challenge[1] % GL_P, challenge[2] % GL_P};
    if (prevBits - curBits <= FRI_MAX_FOLD_BITS)
        return fri_fold_one(ctx, (const u64*)pol, prevBits, curBits, nextBits, step0Bits, ch, split, (u64*)pol_out, (u64*)rows_out, (u64*)nodes_out);
    // wide fold: chunks of FRI_MAX_FOLD_BITS through two ping-pong scratch polynomials, the challenge squared once per halving
    u64 *sa = nullptr, *sb = nullptr;
    const size_t sw = (size_t)3 << (prevBits - FRI_MAX_FOLD_BITS);
    CU(cudaMallocFromPoolAsync(&sa, sw * sizeof(u64), ctx->pool, ctx->stream));
    cudaError_t e = cudaMallocFromPoolAsync(&sb, (sw >> FRI_MAX_FOLD_BITS ? sw >> FRI_MAX_FOLD_BITS : 3) * sizeof(u64), ctx->pool, ctx->stream);
    if (e != cudaSuccess) { cudaFreeAsync(sa, ctx->stream); return fail(PIL2GPU_E_NOMEM, "scratch allocation failed: %s", cudaGetErrorString(e)); }
    const u64* in = (const u64*)pol;
    uint32_t at = prevBits;
    int rc = PIL2GPU_OK, turn = 0;
    while (rc == PIL2GPU_OK && at - curBits > FRI_MAX_FOLD_BITS) {
        u64* out = turn ? sb : sa;
        rc = fri_fold_one(ctx, in, at, at - FRI_MAX_FOLD_BITS, -1, step0Bits, ch, split, out, nullptr, nullptr);
        h3 c = {{ch[0], ch[1], ch[2]}};
        for (int k = 0; k < FRI_MAX_FOLD_BITS; k++) c = h3_mul(c, c);
        ch[0] = c.c[0]; ch[1] = c.c[1]; ch[2] = c.c[2];
        in = out;
        at -= FRI_MAX_FOLD_BITS;
        turn ^= 1;
    }
    if (rc == PIL2GPU_OK) rc = fri_fold_one(ctx, in, at, curBits, nextBits, step0Bits, ch, split, (u64*)pol_out, (u64*)rows_out, (u64*)nodes_out);
    cudaFreeAsync(sa, ctx->stream);
    cudaFreeAsync(sb, ctx->stream);
    return rc;
}

// Sharded FRI chains: one rank's share of a fold.  Computes rows [row0, row0 + n_rows) of the next layer's
// transposed buffer (= the outputs g = i + 2^nextBits * j of those rows i) at their absolute positions in rows_out_dev, and
// pol_out_dev[g] for the same outputs when pol_out_dev != NULL.  in_layout 0: in_dev is the polynomial (fri.js order);
// in_layout 1: in_dev is the previous layer's transposed buffer (its nextBits == this prevBits - ... == curBits), whose row g holds
// the 2^(prevBits-curBits) inputs of output g contiguously -- so a chain that all-gathers the rows of every layer never
// needs the polynomial order again.  No hashing here: the caller hashes its rows (pil2gpu_merkelize_dev on the slice).
int pil2gpu_fri_fold_range_dev(pil2gpu_ctx* ctx, const uint64_t* in_dev, int in_layout, uint32_t prevBits, uint32_t curBits, int32_t nextBits,
                               uint32_t step0Bits, const uint64_t challenge[3], uint64_t row0, uint64_t n_rows, uint64_t* pol_out_dev,
                               uint64_t* rows_out_dev) {
    ENTER(ctx);
    if (!in_dev || !challenge || (!pol_out_dev && !rows_out_dev)) return fail(PIL2GPU_E_INVALID, "null argument");
    if (curBits > prevBits || prevBits > step0Bits || step0Bits > 32) return fail(PIL2GPU_E_INVALID, "bad FRI step sizes");
    if (nextBits >= 0 && (uint32_t)nextBits > curBits) return fail(PIL2GPU_E_INVALID, "bad next-layer description");
    if (prevBits - curBits > FRI_MAX_FOLD_BITS) return fail(PIL2GPU_E_UNSUPPORTED, "fold by 2^%u not supported (max 2^%d)", prevBits - curBits, FRI_MAX_FOLD_BITS);
    if (in_layout != 0 && in_layout != 1) return fail(PIL2GPU_E_INVALID, "in_layout must be 0 (polynomial) or 1 (rows)");
    const uint32_t nb = nextBits >= 0 ? (uint32_t)nextBits : curBits;
    const u64 all_rows = 1ULL << nb;
    if (n_rows == 0) { row0 = 0; n_rows = all_rows; }
    if (row0 > all_rows || n_rows > all_rows - row0) return fail(PIL2GPU_E_RANGE, "row range outside the layer");
    if (nextBits >= 0 && !rows_out_dev) return fail(PIL2GPU_E_INVALID, "rows_out_dev is required when nextBits >= 0");
    FriParams P;
    P.prev_bits = (int)prevBits;
    P.cur_bits = (int)curBits;
    P.next_bits = nextBits;
    P.write_rows = nextBits >= 0;
    P.fuse_leaf_hash = 0;
    P.in_rows = in_layout;
    P.write_pol = pol_out_dev != nullptr;
    P.row0 = row0;
    P.n_rows = n_rows;
    u64 si = glh_inv(GL_SHIFT);                                    // fri.js:31-36
    for (uint32_t j = 0; j < step0Bits - prevBits; j++) si = glh_mul(si, si);
    P.shift_inv = si;
    P.nx_inv = glh_inv(1ULL << (prevBits - curBits));
    for (int k = 0; k < 3; k++) P.challenge[k] = challenge[k];
    int l = fri_launch_fold((const u64*)in_dev, (u64*)pol_out_dev, (u64*)rows_out_dev, nullptr, P, ctx->tb, ctx->stream);
    if (l < 0) return fail(PIL2GPU_E_UNSUPPORTED, "row range must be a multiple of %d rows", FRI_ROWS_PER_CTA);
    return check_launch(ctx, l, "fri_fold_range");
}

int pil2gpu_fri_fold_paged(pil2gpu_ctx* ctx, const uint64_t* const* pol_pages, const uint64_t* pol_page_words, uint32_t n_pol_pages,
                           uint32_t prevBits, uint32_t curBits, int32_t nextBits, uint32_t step0Bits, const uint64_t challenge[3], int split,
                           uint64_t* const* out_pages, const uint64_t* out_page_words, uint32_t n_out_pages, uint64_t* const* rows_pages,
                           const uint64_t* rows_page_words, uint32_t n_rows_pages, uint64_t* nodes_out) {
    ENTER(ctx);
    if (!pol_pages || !pol_page_words || !out_pages || !out_page_words || !challenge) return fail(PIL2GPU_E_INVALID, "null argument");
    if (prevBits > 32 || curBits > prevBits || prevBits > step0Bits || step0Bits > 32) return fail(PIL2GPU_E_INVALID, "bad FRI step sizes");
    if (nextBits >= 0 && (uint32_t)nextBits > curBits) return fail(PIL2GPU_E_INVALID, "bad next-layer description");
    const size_t pw = (size_t)3 << prevBits, cw = (size_t)3 << curBits;
    const u64 height = nextBits >= 0 ? (1ULL << nextBits) : 0;
    const size_t nw = nextBits >= 0 ? merkle_nnodes_words(height) : 0;
    const PageList pp = {pol_pages, pol_page_words, n_pol_pages}, op = {(const uint64_t* const*)out_pages, out_page_words, n_out_pages},
                   rp = {(const uint64_t* const*)rows_pages, rows_page_words, n_rows_pages};
    const bool want_rows = nextBits >= 0 && rows_pages != nullptr && n_rows_pages > 0;
    int rc = check_pages(pp, pw, "pol");
    if (!rc) rc = check_pages(op, cw, "pol_out");
    if (!rc && want_rows) rc = check_pages(rp, cw, "rows_out");
    if (!rc) rc = ensure_ws(ctx, ev2(pw) + 2 * ev2(cw) + nw);
    if (rc) return rc;
    u64 *a = ctx->ws, *b = ctx->ws + ev2(pw), *r = ctx->ws + ev2(pw) + ev2(cw), *n = ctx->ws + ev2(pw) + 2 * ev2(cw);
    rc = pages_to_dev(ctx, a, pp, pw, ctx->stream);
    if (!rc) rc = pil2gpu_fri_fold_dev(ctx, a, prevBits, curBits, nextBits, step0Bits, challenge, split, b, r, n);
    if (!rc) rc = dev_to_pages(ctx, b, op, cw, ctx->stream);
    if (!rc && want_rows) rc = dev_to_pages(ctx, r, rp, cw, ctx->stream);
    if (!rc && nextBits >= 0 && nodes_out) rc = d2h_2d(ctx, (char*)nodes_out, nw * 8, (const char*)n, nw * 8, nw * 8, 1, ctx->stream);
    cudaError_t es = sync_all_streams(ctx);
    if (rc) return rc;
    if (es != cudaSuccess) return fail(PIL2GPU_E_CUDA, "fri_fold: %s", cudaGetErrorString(es));
    return PIL2GPU_OK;
}

int pil2gpu_fri_fold(pil2gpu_ctx* ctx, const uint64_t* pol, uint32_t prevBits, uint32_t curBits, int32_t nextBits, uint32_t step0Bits,
                     const uint64_t challenge[3], int split, uint64_t* pol_out, uint64_t* rows_out, uint64_t* nodes_out) {
    if (!pol || !pol_out) return fail(PIL2GPU_E_INVALID, "null argument");
    if (prevBits > 32 || curBits > prevBits) return fail(PIL2GPU_E_INVALID, "bad FRI step sizes");
    const uint64_t pw = (uint64_t)3 << prevBits, cw = (uint64_t)3 << curBits;
    const uint64_t* pp[1] = {pol};
    uint64_t *op[1] = {pol_out}, *rp[1] = {rows_out};
    return pil2gpu_fri_fold_paged(ctx, pp, &pw, 1, prevBits, curBits, nextBits, step0Bits, challenge, split, op, &cw, 1, rows_out ? rp : nullptr, &cw,
                                  rows_out ? 1 : 0, nodes_out);
}

// ---- multi-GPU commit group (shard.cuh): peer-mapped receive buffers + mailboxes, flag barriers, no collective library ----
struct pil2gpu_shard {
    pil2gpu_ctx* ctx;
    uint32_t rank, world;
    uint64_t recv_words, stage_words;
    u64 *recv, *mail;                                    // own allocations
    u64 *peer_recv[SHARD_MAX_RANKS], *peer_mail[SHARD_MAX_RANKS];   // every rank's buffers as seen from this device (own included)
    bool ipc_recv[SHARD_MAX_RANKS], ipc_mail[SHARD_MAX_RANKS];      // opened with cudaIpcOpenMemHandle (to be closed)
    bool connected;
    uint64_t epoch, timeout_ns;
    ShardPeers peers() const {
        ShardPeers P;
        for (int i = 0; i < SHARD_MAX_RANKS; i++) P.mail[i] = peer_mail[i];
        return P;
    }
};

// CUDA loads kernels lazily, and loading one synchronises with the kernels that are running.  When one host thread drives several ranks,
// a rank spinning in a flag barrier would therefore block the load of a kernel its peer still has to launch -- until the barrier times
// out.  Every kernel of the commit / open path is loaded when a group member is created, before any barrier can be in flight.
extern "C++" {
template <typename K>
static void shard_preload_one(K kernel) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, kernel) != cudaSuccess) cudaGetLastError();
}
}
static void shard_preload_kernels() {
    shard_preload_one(ntt_pass_kernel<true, true, false, false>);
    shard_preload_one(ntt_pass_kernel<true, true, false, true>);
    shard_preload_one(ntt_pass_kernel<false, false, false, false>);
    shard_preload_one(ntt_pass_kernel<false, false, false, true>);
    shard_preload_one(ntt_pass_kernel<false, false, true, false>);
    shard_preload_one(ntt_pass_kernel<false, false, true, true>);
    shard_preload_one(ntt_lde_fused_kernel<false>);
    shard_preload_one(ntt_lde_fused_kernel<true>);
    shard_preload_one(merkle_leaf_kernel);
    shard_preload_one(merkle_batch_kernel);
    shard_preload_one(merkle_levels_kernel<1>);
    shard_preload_one(merkle_levels_kernel<2>);
    shard_preload_one(merkle_levels_kernel<3>);
    shard_preload_one(merkle_tail_kernel);
    shard_preload_one(merkle_pad_kernel);
    shard_preload_one(shard_barrier_kernel);
    shard_preload_one(shard_publish_kernel);
    shard_preload_one(shard_group_proof_kernel);
    shard_preload_one(shard_unpack_kernel);
}

static int shard_check(const pil2gpu_shard* sh, bool need_connected) {
    if (!sh || !sh->ctx) return fail(PIL2GPU_E_INVALID, "null shard");
    if (need_connected && !sh->connected) return fail(PIL2GPU_E_INVALID, "shard is not connected to its peers (pil2gpu_shard_connect[_local])");
    return PIL2GPU_OK;
}

int pil2gpu_shard_create(pil2gpu_ctx* ctx, uint32_t rank, uint32_t world, uint64_t recv_words, uint64_t stage_words, pil2gpu_shard** out) {
    ENTER(ctx);
    if (!out) return fail(PIL2GPU_E_INVALID, "null out");
    if (world == 0 || world > SHARD_MAX_RANKS || (world & (world - 1)) || rank >= world)
        return fail(PIL2GPU_E_INVALID, "bad rank description (world must be a power of two <= %d)", SHARD_MAX_RANKS);
    pil2gpu_shard* sh = new (std::nothrow) pil2gpu_shard();
    if (!sh) return fail(PIL2GPU_E_NOMEM, "out of host memory");
    shard_preload_kernels();
    sh->ctx = ctx; sh->rank = rank; sh->world = world; sh->recv_words = recv_words; sh->stage_words = stage_words;
    sh->recv = sh->mail = nullptr; sh->connected = false; sh->epoch = 0;
    const char* tm = getenv("PIL2GPU_SHARD_TIMEOUT_MS");
    sh->timeout_ns = (uint64_t)(tm ? atoll(tm) : 20000) * 1000000ull;
    for (int i = 0; i < SHARD_MAX_RANKS; i++) { sh->peer_recv[i] = sh->peer_mail[i] = nullptr; sh->ipc_recv[i] = sh->ipc_mail[i] = false; }
    const size_t mail_bytes = (SHARD_STAGE + stage_words) * sizeof(u64);
    cudaError_t e = cudaMalloc(&sh->recv, (recv_words ? recv_words : 2) * sizeof(u64));
    if (e == cudaSuccess) e = cudaMalloc(&sh->mail, mail_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(sh->mail, 0, mail_bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        if (sh->recv) cudaFree(sh->recv);
        if (sh->mail) cudaFree(sh->mail);
        delete sh;
        return fail(e == cudaErrorMemoryAllocation ? PIL2GPU_E_NOMEM : PIL2GPU_E_CUDA, "shard_create: %s", cudaGetErrorString(e));
    }
    sh->peer_recv[rank] = sh->recv;
    sh->peer_mail[rank] = sh->mail;
    if (world == 1) sh->connected = true;
    *out = sh;
    return PIL2GPU_OK;
}

int pil2gpu_shard_destroy(pil2gpu_shard* sh) {
    if (!sh) return PIL2GPU_OK;
    ENTER(sh->ctx);
    sync_all_streams(sh->ctx);
    for (uint32_t r = 0; r < sh->world; r++) {
        if (sh->ipc_recv[r]) cudaIpcCloseMemHandle(sh->peer_recv[r]);
        if (sh->ipc_mail[r]) cudaIpcCloseMemHandle(sh->peer_mail[r]);
    }
    cudaFree(sh->recv);
    cudaFree(sh->mail);
    delete sh;
    return PIL2GPU_OK;
}

int pil2gpu_shard_handles(pil2gpu_shard* sh, uint8_t handles_out[128]) {
    int rc = shard_check(sh, false);
    if (rc) return rc;
    ENTER(sh->ctx);
    if (!handles_out) return fail(PIL2GPU_E_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, sh->recv));
    memcpy(handles_out, &h, 64);
    CU(cudaIpcGetMemHandle(&h, sh->mail));
    memcpy(handles_out + 64, &h, 64);
    return PIL2GPU_OK;
}

int pil2gpu_shard_connect(pil2gpu_shard* sh, const uint8_t* handles, uint32_t n_ranks) {
    int rc = shard_check(sh, false);
    if (rc) return rc;
    ENTER(sh->ctx);
    if (!handles || n_ranks != sh->world) return fail(PIL2GPU_E_INVALID, "need the 128-byte handle pair of each of the %u ranks", sh->world);
    if (sh->connected && sh->world > 1) return fail(PIL2GPU_E_INVALID, "shard is already connected");
    for (uint32_t r = 0; r < sh->world; r++) {
        if (r == sh->rank) continue;
        cudaIpcMemHandle_t h;
        void* p = nullptr;
        memcpy(&h, handles + 128 * r, 64);
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        sh->peer_recv[r] = (u64*)p; sh->ipc_recv[r] = true;
        memcpy(&h, handles + 128 * r + 64, 64);
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        sh->peer_mail[r] = (u64*)p; sh->ipc_mail[r] = true;
    }
    sh->connected = true;
    return PIL2GPU_OK;
}

int pil2gpu_shard_connect_local(pil2gpu_shard* const* group, uint32_t n_ranks) {
    if (!group || n_ranks == 0 || n_ranks > SHARD_MAX_RANKS) return fail(PIL2GPU_E_INVALID, "bad group");
    for (uint32_t r = 0; r < n_ranks; r++)
        if (!group[r] || group[r]->rank != r || group[r]->world != n_ranks) return fail(PIL2GPU_E_INVALID, "group[%u] is not rank %u of %u", r, r, n_ranks);
    for (uint32_t a = 0; a < n_ranks; a++) {
        pil2gpu_shard* sa = group[a];
        DeviceGuard g(sa->ctx->device);
        if (!g.ok) return fail(PIL2GPU_E_CUDA, "cudaSetDevice(%d) failed", sa->ctx->device);
        for (uint32_t b = 0; b < n_ranks; b++) {
            pil2gpu_shard* sb = group[b];
            if (sb->ctx->device != sa->ctx->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, sa->ctx->device, sb->ctx->device));
                if (!can) return fail(PIL2GPU_E_UNSUPPORTED, "device %d cannot map device %d", sa->ctx->device, sb->ctx->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(sb->ctx->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return fail(PIL2GPU_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            }
            sa->peer_recv[b] = sb->recv;
            sa->peer_mail[b] = sb->mail;
        }
        sa->connected = true;
    }
    return PIL2GPU_OK;
}

uint64_t* pil2gpu_shard_recv_dev(pil2gpu_shard* sh) { return sh ? (uint64_t*)sh->recv : nullptr; }
uint64_t* const* pil2gpu_shard_peer_recv(pil2gpu_shard* sh) { return sh ? (uint64_t* const*)sh->peer_recv : nullptr; }
const uint64_t* pil2gpu_shard_sub_roots_dev(pil2gpu_shard* sh) { return sh ? (const uint64_t*)(sh->mail + SHARD_SUB) : nullptr; }
const uint64_t* pil2gpu_shard_top_nodes_dev(pil2gpu_shard* sh) { return sh ? (const uint64_t*)(sh->mail + SHARD_TOP) : nullptr; }

static int shard_barrier_enqueue(pil2gpu_shard* sh) {
    sh->epoch++;
    if (sh->world > 1) {
        shard_barrier_kernel<<<1, SHARD_MAX_RANKS, 0, sh->ctx->stream>>>(sh->peers(), sh->rank, sh->world, sh->epoch, sh->timeout_ns);
        return check_launch(sh->ctx, 1, "shard_barrier");
    }
    return PIL2GPU_OK;
}
int pil2gpu_shard_barrier(pil2gpu_shard* sh) {
    int rc = shard_check(sh, true);
    if (rc) return rc;
    ENTER(sh->ctx);
    return shard_barrier_enqueue(sh);
}

int pil2gpu_shard_status(pil2gpu_shard* sh) {
    int rc = shard_check(sh, false);
    if (rc) return rc;
    ENTER(sh->ctx);
    u64 err = 0;
    CU(cudaMemcpyAsync(&err, sh->mail + SHARD_ERR, sizeof(u64), cudaMemcpyDeviceToHost, sh->ctx->stream));
    CU(cudaStreamSynchronize(sh->ctx->stream));
    if (err) return fail(PIL2GPU_E_CUDA, "shard barrier %llu timed out on rank %u: a peer never arrived", (unsigned long long)err, sh->rank);
    return PIL2GPU_OK;
}

// after the barrier that follows shard_publish: tree over the sub-roots in this rank's mailbox -> root
static int shard_top_tree(pil2gpu_shard* sh, uint64_t* root_out_dev) {
    pil2gpu_ctx* ctx = sh->ctx;
    u64* top = sh->mail + SHARD_TOP;
    const u64 nt = merkle_nnodes_words(sh->world);
    if (sh->world > 1) {
        CU(cudaMemcpyAsync(top, sh->mail + SHARD_SUB, 4 * sh->world * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        int l = merkle_launch_tree(top, sh->world, ctx->stream);
        int rc = check_launch(ctx, l, "shard top tree");
        if (rc) return rc;
    }
    if (root_out_dev)
        CU(cudaMemcpyAsync(root_out_dev, sh->world > 1 ? top + nt - 4 : sh->mail + SHARD_SUB, 4 * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    return PIL2GPU_OK;
}

int pil2gpu_shard_hash_dev(pil2gpu_shard* sh, uint64_t nPols, uint32_t nBitsExt, int split, uint64_t* nodes_dev, uint64_t* root_out_dev) {
    int rc = shard_check(sh, true);
    if (rc) return rc;
    pil2gpu_ctx* ctx = sh->ctx;
    ENTER(ctx);
    if (!nodes_dev) return fail(PIL2GPU_E_INVALID, "null nodes");
    if (nPols == 0 || nPols % sh->world) return fail(PIL2GPU_E_INVALID, "nPols (%llu) must be a positive multiple of the number of ranks", (unsigned long long)nPols);
    if (nBitsExt > 40 || ((u64)1 << nBitsExt) < sh->world) return fail(PIL2GPU_E_INVALID, "fewer extended rows than ranks");
    const u64 cg = nPols / sh->world, rows_local = ((u64)1 << nBitsExt) / sh->world;
    if (cg * rows_local * sh->world > sh->recv_words) return fail(PIL2GPU_E_INVALID, "receive buffer too small: %llu words needed", (unsigned long long)(cg * rows_local * sh->world));
    rc = pil2gpu_merkelize_tiled_dev(ctx, sh->recv, sh->world, cg, rows_local * cg, rows_local, split, nodes_dev);
    if (rc) return rc;
    const u64 nn = merkle_nnodes_words(rows_local);
    shard_publish_kernel<<<1, 4 * SHARD_MAX_RANKS, 0, ctx->stream>>>(sh->peers(), sh->rank, sh->world, (const u64*)nodes_dev + nn - 4);
    rc = check_launch(ctx, 1, "shard_publish");
    if (!rc) rc = shard_barrier_enqueue(sh);
    if (!rc) rc = shard_top_tree(sh, root_out_dev);
    return rc;
}

int pil2gpu_shard_commit_dev(pil2gpu_shard* sh, const uint64_t* src_slab_dev, uint64_t* work_dev, uint64_t nPols, uint32_t nBits, uint32_t nBitsExt,
                             int split, uint64_t* nodes_dev, uint64_t* root_out_dev) {
    int rc = shard_check(sh, true);
    if (rc) return rc;
    pil2gpu_ctx* ctx = sh->ctx;
    ENTER(ctx);
    if (nPols == 0 || nPols % sh->world) return fail(PIL2GPU_E_INVALID, "nPols (%llu) must be a positive multiple of the number of ranks", (unsigned long long)nPols);
    const u64 cg = nPols / sh->world;
    if (sh->world > 1 && cg % 8) return fail(PIL2GPU_E_UNSUPPORTED, "columns per rank must be a multiple of 8 (sponge chunks must not straddle tiles)");
    if (nBitsExt > 40 || (cg << nBitsExt) > sh->recv_words) return fail(PIL2GPU_E_INVALID, "receive buffer too small");
    rc = shard_barrier_enqueue(sh);                 // every peer is done with its receive buffer (previous commit, queries, downloads)
    if (!rc) rc = pil2gpu_lde_scatter_dev(ctx, src_slab_dev, work_dev, cg, nBits, nBitsExt, (uint64_t* const*)sh->peer_recv, sh->world, sh->rank, 0, 0);
    if (!rc) rc = shard_barrier_enqueue(sh);        // every rank's rows have landed
    if (!rc) rc = pil2gpu_shard_hash_dev(sh, nPols, nBitsExt, split, nodes_dev, root_out_dev);
    return rc;
}

int pil2gpu_shard_open_dev(pil2gpu_shard* sh, const uint64_t* nodes_dev, uint64_t nPols, uint32_t nBitsExt, const uint64_t* idxs_dev, uint32_t n_idx,
                           uint64_t* rows_out_dev, uint64_t* siblings_out_dev) {
    int rc = shard_check(sh, true);
    if (rc) return rc;
    pil2gpu_ctx* ctx = sh->ctx;
    ENTER(ctx);
    if (!nodes_dev || !idxs_dev || !rows_out_dev || !siblings_out_dev) return fail(PIL2GPU_E_INVALID, "null argument");
    if (nPols == 0 || nPols % sh->world || nBitsExt > 40 || ((u64)1 << nBitsExt) < sh->world) return fail(PIL2GPU_E_INVALID, "bad shape");
    if (n_idx == 0) return PIL2GPU_OK;
    const u64 cg = nPols / sh->world, rows_local = ((u64)1 << nBitsExt) / sh->world;
    int dl = 0, dt = 0;
    while (((u64)1 << dl) < rows_local) dl++;
    while ((1u << dt) < sh->world) dt++;
    const u64 slot = nPols + 4 * (u64)(dl + dt);
    if (slot * n_idx > sh->stage_words) return fail(PIL2GPU_E_INVALID, "mailbox staging too small: %llu words needed, %llu available",
                                                    (unsigned long long)(slot * n_idx), (unsigned long long)sh->stage_words);
    rc = shard_barrier_enqueue(sh);                 // the peers have consumed the staging area of the previous call
    if (rc) return rc;
    RowTiles t = {sh->recv, cg, rows_local * cg};
    shard_group_proof_kernel<<<n_idx, 128, 0, ctx->stream>>>(t, (const u64*)nodes_dev, nPols, rows_local, (const u64*)idxs_dev, dl, dt, sh->peers(), sh->rank,
                                                             sh->world);
    rc = check_launch(ctx, 1, "shard_group_proof");
    if (!rc) rc = shard_barrier_enqueue(sh);        // every owner has delivered
    if (rc) return rc;
    shard_unpack_kernel<<<n_idx, 128, 0, ctx->stream>>>(sh->mail + SHARD_STAGE, nPols, dl + dt, (u64*)rows_out_dev, (u64*)siblings_out_dev);
    return check_launch(ctx, 1, "shard_unpack");
}

}   // extern "C"
