// Node N-API addon: marshals BigUint64Array / BigBuffer pages to the C ABI of include/pil2gpu.h.  No arithmetic here.
// Build (on a machine with Node >= 16 and node-gyp; this repo's build container has neither, so this file is
// type-checked against tests/stubs/node_api.h only):   cd napi && node-gyp configure build
//
// Conventions
//   * Every exported function maps to one pil2gpu_* entry point.  Failures become JS Errors carrying pil2gpu_last_error()
//     ("Out of range" for PIL2GPU_E_RANGE, as merklehash_p.js:143 throws).
//   * Every typed-array / page-list length is checked against the shape arguments BEFORE the C call: a wrong-sized buffer is a
//     RangeError in JS, never an out-of-bounds host access inside a copy.  Wrong argument types are TypeErrors; N-API status
//     codes are never ignored.
//   * The heavy calls (anything that moves a buffer) run on a libuv worker through napi_async_work and return a Promise, so the
//     Node main thread is not blocked for the seconds a 2^23 x 256 commit takes; the arguments are kept alive by references until
//     the call completes.  The reference awaits these calls one at a time (proofGen), which is also what a ctx supports.
//   * allocPinnedPage(words) returns a BigUint64Array over page-locked memory (pil2gpu_host_alloc): BigBuffer pages allocated
//     this way are copied by DMA directly; ordinary pages work too (staged through the library's pinned ring).
//   * Device-resident trees (commit / treeRoot / treeGroupProofs / treeDownload / treeFree) let proofQueries run against HBM:
//     only the opened rows and siblings cross PCIe.
#include <node_api.h>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include "../include/pil2gpu.h"

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// argument access with status checks
// ---------------------------------------------------------------------------------------------------------------------
struct Args {
    napi_env env;
    size_t argc;
    napi_value v[20];
    bool ok = true;

    Args(napi_env e, napi_callback_info info, size_t want) : env(e), argc(20) {
        if (napi_get_cb_info(env, info, &argc, v, nullptr, nullptr) != napi_ok) { type_error("could not read the arguments"); return; }
        if (argc < want) type_error("too few arguments");
    }
    void type_error(const char* m) { if (ok) napi_throw_type_error(env, nullptr, m); ok = false; }
    void range_error(const char* m) { if (ok) napi_throw_range_error(env, nullptr, m); ok = false; }
    void need(bool cond, const char* m) { if (!cond) range_error(m); }

    bool is_nullish(size_t i) {
        if (i >= argc) return true;
        napi_valuetype t;
        if (napi_typeof(env, v[i], &t) != napi_ok) { type_error("napi_typeof failed"); return true; }
        return t == napi_undefined || t == napi_null;
    }
    pil2gpu_ctx* ctx(size_t i) {
        void* p = nullptr;
        if (!ok) return nullptr;
        if (napi_get_value_external(env, v[i], &p) != napi_ok || !p) { type_error("expected a pil2gpu context (addon.create())"); return nullptr; }
        return (pil2gpu_ctx*)p;
    }
    void* external(size_t i, const char* what) {
        void* p = nullptr;
        if (!ok) return nullptr;
        if (napi_get_value_external(env, v[i], &p) != napi_ok || !p) { type_error(what); return nullptr; }
        return p;
    }
    uint32_t u32(size_t i) {
        uint32_t x = 0;
        if (ok && napi_get_value_uint32(env, v[i], &x) != napi_ok) type_error("expected an unsigned 32-bit number");
        return x;
    }
    int32_t i32(size_t i) {
        int32_t x = 0;
        if (ok && napi_get_value_int32(env, v[i], &x) != napi_ok) type_error("expected a 32-bit number");
        return x;
    }
    // sizes arrive as Numbers (exact below 2^53) or BigInts
    uint64_t u64(size_t i) {
        if (!ok) return 0;
        napi_valuetype t;
        if (napi_typeof(env, v[i], &t) != napi_ok) { type_error("napi_typeof failed"); return 0; }
        if (t == napi_bigint) {
            uint64_t x = 0; bool lossless = false;
            if (napi_get_value_bigint_uint64(env, v[i], &x, &lossless) != napi_ok || !lossless) { range_error("BigInt out of the u64 range"); return 0; }
            return x;
        }
        double d = 0;
        if (napi_get_value_double(env, v[i], &d) != napi_ok) { type_error("expected a Number or BigInt size"); return 0; }
        if (!(d >= 0) || d > 9007199254740992.0 || d != (double)(uint64_t)d) { range_error("expected a non-negative integer"); return 0; }
        return (uint64_t)d;
    }
    bool u64_array(size_t i, uint64_t** data, size_t* len) {
        napi_typedarray_type ty; napi_value ab; size_t off; bool is_ta = false;
        if (!ok) return false;
        if (napi_is_typedarray(env, v[i], &is_ta) != napi_ok || !is_ta ||
            napi_get_typedarray_info(env, v[i], &ty, len, (void**)data, &ab, &off) != napi_ok || ty != napi_biguint64_array) {
            type_error("expected a BigUint64Array");
            return false;
        }
        return true;
    }
    // null / undefined -> (nullptr, 0)
    bool opt_u64_array(size_t i, uint64_t** data, size_t* len) {
        *data = nullptr; *len = 0;
        if (is_nullish(i)) return ok;
        return u64_array(i, data, len);
    }
    bool i32_array(size_t i, int32_t** data, size_t* len) {
        napi_typedarray_type ty; napi_value ab; size_t off; bool is_ta = false;
        if (!ok) return false;
        if (napi_is_typedarray(env, v[i], &is_ta) != napi_ok || !is_ta ||
            napi_get_typedarray_info(env, v[i], &ty, len, (void**)data, &ab, &off) != napi_ok || ty != napi_int32_array) {
            type_error("expected an Int32Array");
            return false;
        }
        return true;
    }
    bool i64_array(size_t i, int64_t** data, size_t* len) {
        napi_typedarray_type ty; napi_value ab; size_t off; bool is_ta = false;
        if (!ok) return false;
        if (napi_is_typedarray(env, v[i], &is_ta) != napi_ok || !is_ta ||
            napi_get_typedarray_info(env, v[i], &ty, len, (void**)data, &ab, &off) != napi_ok || ty != napi_bigint64_array) {
            type_error("expected a BigInt64Array");
            return false;
        }
        return true;
    }
    // Array of BigUint64Array pages; total = sum of the page lengths.  null / undefined -> no pages when optional.
    bool pages(size_t i, std::vector<uint64_t*>& p, std::vector<uint64_t>& w, uint64_t* total, bool optional = false) {
        *total = 0;
        if (!ok) return false;
        if (optional && is_nullish(i)) return true;
        uint32_t n = 0; bool is_arr = false;
        if (napi_is_array(env, v[i], &is_arr) != napi_ok || !is_arr || napi_get_array_length(env, v[i], &n) != napi_ok) {
            type_error("expected an Array of BigUint64Array pages");
            return false;
        }
        for (uint32_t k = 0; k < n; k++) {
            napi_value e; napi_typedarray_type ty; napi_value ab; size_t off, len; uint64_t* d; bool is_ta = false;
            if (napi_get_element(env, v[i], k, &e) != napi_ok || napi_is_typedarray(env, e, &is_ta) != napi_ok || !is_ta ||
                napi_get_typedarray_info(env, e, &ty, &len, (void**)&d, &ab, &off) != napi_ok || ty != napi_biguint64_array) {
                type_error("expected an Array of BigUint64Array pages");
                return false;
            }
            p.push_back(d); w.push_back(len); *total += len;
        }
        return true;
    }
};

bool shl_fits(uint64_t a, uint32_t bits, uint64_t* out) {   // a << bits without overflow
    if (bits >= 64 || (bits && (a >> (64 - bits)))) return false;
    *out = a << bits;
    return true;
}

napi_value make_u64_array(napi_env env, const uint64_t* src, size_t n) {
    napi_value ab, ta; void* data;
    if (napi_create_arraybuffer(env, n * 8, &data, &ab) != napi_ok) return nullptr;
    if (n) memcpy(data, src, n * 8);
    if (napi_create_typedarray(env, napi_biguint64_array, n, ab, 0, &ta) != napi_ok) return nullptr;
    return ta;
}
napi_value fail_now(napi_env env, int rc) {
    napi_throw_error(env, nullptr, rc == PIL2GPU_E_RANGE ? "Out of range" : pil2gpu_last_error());
    return nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// asynchronous calls: run() on a worker thread, result() back on the main thread, Promise to the caller
// ---------------------------------------------------------------------------------------------------------------------
struct Job {
    napi_async_work work = nullptr;
    napi_deferred deferred = nullptr;
    std::function<int()> run;
    std::function<napi_value(napi_env)> result;      // may be empty: resolves with undefined
    std::vector<napi_ref> keep;                      // arguments (typed arrays, page lists, handles) alive until completion
    int rc = 0;
    std::string err;
};
void job_execute(napi_env, void* data) {
    Job* j = (Job*)data;
    j->rc = j->run();
    if (j->rc) j->err = pil2gpu_last_error();        // thread-local: read it on the thread that made the call
}
void job_complete(napi_env env, napi_status status, void* data) {
    Job* j = (Job*)data;
    napi_value val = nullptr;
    bool rejected = true;
    if (status == napi_ok && j->rc == 0) {
        if (j->result) val = j->result(env);
        else napi_get_undefined(env, &val);
        rejected = (val == nullptr);
    }
    if (!rejected) {
        napi_resolve_deferred(env, j->deferred, val);
    } else {
        const char* m = status != napi_ok ? "pil2gpu addon: the call was cancelled"
                        : (j->rc == PIL2GPU_E_RANGE ? "Out of range" : (j->rc ? j->err.c_str() : "pil2gpu addon: could not build the result"));
        napi_value msg, err;
        if (napi_create_string_utf8(env, m, NAPI_AUTO_LENGTH, &msg) == napi_ok && napi_create_error(env, nullptr, msg, &err) == napi_ok)
            napi_reject_deferred(env, j->deferred, err);
    }
    for (napi_ref r : j->keep) napi_delete_reference(env, r);
    napi_delete_async_work(env, j->work);
    delete j;
}
// Takes ownership of j.  keep[]: the argument values to hold on to.
napi_value launch(napi_env env, Job* j, const napi_value* keep, size_t n_keep, const char* name) {
    napi_value promise = nullptr, rname;
    bool ok = napi_create_promise(env, &j->deferred, &promise) == napi_ok && napi_create_string_utf8(env, name, NAPI_AUTO_LENGTH, &rname) == napi_ok;
    for (size_t i = 0; ok && i < n_keep; i++) {
        napi_valuetype t;
        if (napi_typeof(env, keep[i], &t) != napi_ok) { ok = false; break; }
        if (t != napi_object && t != napi_external && t != napi_function) continue;
        napi_ref r;
        if (napi_create_reference(env, keep[i], 1, &r) != napi_ok) { ok = false; break; }
        j->keep.push_back(r);
    }
    ok = ok && napi_create_async_work(env, nullptr, rname, job_execute, job_complete, j, &j->work) == napi_ok;
    ok = ok && napi_queue_async_work(env, j->work) == napi_ok;
    if (!ok) {
        for (napi_ref r : j->keep) napi_delete_reference(env, r);
        if (j->work) napi_delete_async_work(env, j->work);
        delete j;
        napi_throw_error(env, nullptr, "pil2gpu addon: could not queue the call");
        return nullptr;
    }
    return promise;
}

struct Pages {   // owned copy of a page list for the worker thread
    std::vector<uint64_t*> p;
    std::vector<uint64_t> w;
    uint64_t total = 0;
    const uint64_t* const* cp() const { return (const uint64_t* const*)p.data(); }
    uint64_t* const* mp() const { return p.data(); }
    uint32_t n() const { return (uint32_t)p.size(); }
};

void ctx_finalize(napi_env, void* data, void*) { pil2gpu_destroy((pil2gpu_ctx*)data); }
void pinned_finalize(napi_env, void* data, void*) { pil2gpu_host_free(data); }

// ---------------------------------------------------------------------------------------------------------------------
// lifetime, memory
// ---------------------------------------------------------------------------------------------------------------------
// create(device) -> external
napi_value Create(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    const int32_t dev = a.i32(0);
    if (!a.ok) return nullptr;
    pil2gpu_ctx* ctx = nullptr;
    int rc = pil2gpu_create(dev, nullptr, &ctx);
    if (rc) return fail_now(env, rc);
    napi_value ext;
    if (napi_create_external(env, ctx, ctx_finalize, nullptr, &ext) != napi_ok) { pil2gpu_destroy(ctx); napi_throw_error(env, nullptr, "napi_create_external failed"); return nullptr; }
    return ext;
}
// allocPinnedPage(words) -> BigUint64Array over page-locked host memory (released when the array is collected)
napi_value AllocPinnedPage(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    const uint64_t words = a.u64(0);
    a.need(words <= ((uint64_t)1 << 40), "page too large");
    if (!a.ok) return nullptr;
    void* p = nullptr;
    int rc = pil2gpu_host_alloc((size_t)words * 8, &p);
    if (rc) return fail_now(env, rc);
    memset(p, 0, (size_t)words * 8);                       // a fresh BigUint64Array reads as zeros
    napi_value ab, ta;
    if (napi_create_external_arraybuffer(env, p, (size_t)words * 8, pinned_finalize, nullptr, &ab) != napi_ok ||
        napi_create_typedarray(env, napi_biguint64_array, (size_t)words, ab, 0, &ta) != napi_ok) {
        pil2gpu_host_free(p);
        napi_throw_error(env, nullptr, "could not wrap the pinned page (external ArrayBuffers disabled?)");
        return nullptr;
    }
    return ta;
}
// releaseWorkspace(ctx): hand the grow-only device workspace back to the driver
napi_value ReleaseWorkspace(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    pil2gpu_ctx* ctx = a.ctx(0);
    if (!a.ok) return nullptr;
    int rc = pil2gpu_release_workspace(ctx);
    return rc ? fail_now(env, rc) : nullptr;
}
// merkleNNodes(height) -> BigInt
napi_value MerkleNNodes(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    const uint64_t h = a.u64(0);
    if (!a.ok) return nullptr;
    napi_value r;
    if (napi_create_bigint_uint64(env, pil2gpu_merkle_nnodes(h), &r) != napi_ok) { napi_throw_error(env, nullptr, "napi_create_bigint_uint64 failed"); return nullptr; }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// fft_p.js: fft / ifft / interpolate over BigBuffer pages
// ---------------------------------------------------------------------------------------------------------------------
// nttPaged(ctx, srcPages, dstPages, nPols, nBits, inverse) -> Promise          fft_p.js:178-184
napi_value NttPaged(napi_env env, napi_callback_info info) {
    Args a(env, info, 6);
    pil2gpu_ctx* ctx = a.ctx(0);
    auto s = std::make_shared<Pages>(), d = std::make_shared<Pages>();
    a.pages(1, s->p, s->w, &s->total); a.pages(2, d->p, d->w, &d->total);
    const uint64_t nPols = a.u64(3); const uint32_t nBits = a.u32(4); const int32_t inverse = a.i32(5);
    uint64_t words = 0;
    a.need(nBits <= 32 && shl_fits(nPols, nBits, &words), "nPols * 2^nBits does not fit");
    a.need(s->total == words, "buffSrc does not hold nPols * 2^nBits elements");
    a.need(d->total == words, "buffDst does not hold nPols * 2^nBits elements");
    if (!a.ok) return nullptr;
    Job* j = new Job;
    j->run = [=] { int rc = pil2gpu_ntt_paged(ctx, s->cp(), s->w.data(), s->n(), d->mp(), d->w.data(), d->n(), nPols, nBits, inverse); return rc; };
    return launch(env, j, a.v, 3, "pil2gpu.ntt");
}
// ldePaged(ctx, srcPages, dstPages, nPols, nBits, nBitsExt) -> Promise         fft_p.js:187-297
napi_value LdePaged(napi_env env, napi_callback_info info) {
    Args a(env, info, 6);
    pil2gpu_ctx* ctx = a.ctx(0);
    auto s = std::make_shared<Pages>(), d = std::make_shared<Pages>();
    a.pages(1, s->p, s->w, &s->total); a.pages(2, d->p, d->w, &d->total);
    const uint64_t nPols = a.u64(3); const uint32_t nBits = a.u32(4), nBitsExt = a.u32(5);
    uint64_t sw = 0, dw = 0;
    a.need(nBitsExt <= 32 && nBits <= nBitsExt && shl_fits(nPols, nBits, &sw) && shl_fits(nPols, nBitsExt, &dw), "bad interpolate shape");
    a.need(s->total == sw, "buffSrc does not hold nPols * 2^nBits elements");
    a.need(d->total == dw, "buffDst does not hold nPols * 2^nBitsExt elements");
    if (!a.ok) return nullptr;
    Job* j = new Job;
    j->run = [=] { int rc = pil2gpu_lde_paged(ctx, s->cp(), s->w.data(), s->n(), d->mp(), d->w.data(), d->n(), nPols, nBits, nBitsExt); return rc; };
    return launch(env, j, a.v, 3, "pil2gpu.lde");
}

// ---------------------------------------------------------------------------------------------------------------------
// merklehash_p.js
// ---------------------------------------------------------------------------------------------------------------------
// merkelizePaged(ctx, elemPages, width, height, split, nodes) -> Promise       merklehash_p.js:44-133
napi_value MerkelizePaged(napi_env env, napi_callback_info info) {
    Args a(env, info, 6);
    pil2gpu_ctx* ctx = a.ctx(0);
    auto e = std::make_shared<Pages>();
    a.pages(1, e->p, e->w, &e->total);
    const uint64_t width = a.u64(2), height = a.u64(3); const int32_t split = a.i32(4);
    uint64_t* nodes = nullptr; size_t nlen = 0;
    a.u64_array(5, &nodes, &nlen);
    a.need(height > 0 && height <= ((uint64_t)1 << 40) && width <= ((uint64_t)1 << 32), "bad tree shape");
    a.need(e->total == width * height, "buff does not hold width * height elements");
    a.need(nlen == pil2gpu_merkle_nnodes(height), "nodes must hold _getNNodes(height * 4) words");
    if (!a.ok) return nullptr;
    Job* j = new Job;
    j->run = [=] { int rc = pil2gpu_merkelize_paged(ctx, e->cp(), e->w.data(), e->n(), width, height, split, nodes); return rc; };
    napi_value keep[3] = {a.v[0], a.v[1], a.v[5]};
    return launch(env, j, keep, 3, "pil2gpu.merkelize");
}
// poseidon(ctx, in12) -> BigUint64Array(12);  linearHash(ctx, vals, split) -> BigUint64Array(4)   (small, synchronous)
napi_value Poseidon(napi_env env, napi_callback_info info) {
    Args a(env, info, 2);
    pil2gpu_ctx* ctx = a.ctx(0);
    uint64_t* in = nullptr; size_t n = 0; uint64_t out[12];
    a.u64_array(1, &in, &n);
    a.need(n == 12, "expected BigUint64Array(12)");
    if (!a.ok) return nullptr;
    int rc = pil2gpu_poseidon(ctx, in, out);
    if (rc) return fail_now(env, rc);
    return make_u64_array(env, out, 12);
}
napi_value LinearHash(napi_env env, napi_callback_info info) {
    Args a(env, info, 3);
    pil2gpu_ctx* ctx = a.ctx(0);
    uint64_t* in = nullptr; size_t n = 0; uint64_t out[4];
    a.u64_array(1, &in, &n);
    const int32_t split = a.i32(2);
    if (!a.ok) return nullptr;
    int rc = pil2gpu_linear_hash(ctx, in, n, split, out);
    if (rc) return fail_now(env, rc);
    return make_u64_array(env, out, 4);
}

// ---------------------------------------------------------------------------------------------------------------------
// device-resident trees: commit / treeFromPages / treeRoot / treeGroupProofs / treeDownload / treeFree
// ---------------------------------------------------------------------------------------------------------------------
struct TreeBox {   // what a JS tree handle points to
    pil2gpu_ctx* ctx;
    pil2gpu_tree* tree;
    uint64_t width, height;
};
void tree_finalize(napi_env, void* data, void*) {
    TreeBox* b = (TreeBox*)data;
    if (b->tree) pil2gpu_tree_free(b->ctx, b->tree);
    delete b;
}
napi_value wrap_tree(napi_env env, TreeBox* b) {
    napi_value ext;
    if (napi_create_external(env, b, tree_finalize, nullptr, &ext) != napi_ok) { tree_finalize(env, b, nullptr); return nullptr; }
    return ext;
}
// commit(ctx, srcPages, nPols, nBits, nBitsExt, split) -> Promise<{ tree: handle, root: BigUint64Array(4) }>
// interpolate + merkelize with the extended buffer and the nodes kept in HBM (stark_gen_helpers.js:388-412 without the download)
napi_value Commit(napi_env env, napi_callback_info info) {
    Args a(env, info, 6);
    pil2gpu_ctx* ctx = a.ctx(0);
    auto s = std::make_shared<Pages>();
    a.pages(1, s->p, s->w, &s->total);
    const uint64_t nPols = a.u64(2); const uint32_t nBits = a.u32(3), nBitsExt = a.u32(4); const int32_t split = a.i32(5);
    uint64_t sw = 0, dw = 0;
    a.need(nBitsExt <= 32 && nBits <= nBitsExt && nPols > 0 && shl_fits(nPols, nBits, &sw) && shl_fits(nPols, nBitsExt, &dw), "bad commit shape");
    a.need(s->total == sw, "buffer does not hold nPols * 2^nBits elements");
    if (!a.ok) return nullptr;
    TreeBox* box = new TreeBox{ctx, nullptr, nPols, (uint64_t)1 << nBitsExt};
    auto root = std::make_shared<std::vector<uint64_t>>(4);
    Job* j = new Job;
    j->run = [=] {
        // upload the pages into a device buffer, then commit from there
        void* d = nullptr;
        int rc = pil2gpu_dev_alloc(ctx, (size_t)sw * 8, &d);
        if (rc) return rc;
        size_t off = 0;
        for (uint32_t k = 0; k < s->n() && !rc; k++) { rc = pil2gpu_h2d(ctx, (uint64_t*)d + off, s->p[k], s->w[k] * 8); off += s->w[k]; }
        if (!rc) rc = pil2gpu_commit_dev(ctx, (const uint64_t*)d, nPols, nBits, nBitsExt, split, &box->tree, root->data());
        pil2gpu_sync(ctx);
        pil2gpu_dev_free(ctx, d);
        return rc;
    };
    j->result = [=](napi_env e) -> napi_value {
        napi_value obj, t = wrap_tree(e, box), r = make_u64_array(e, root->data(), 4);
        if (!t || !r || napi_create_object(e, &obj) != napi_ok || napi_set_named_property(e, obj, "tree", t) != napi_ok ||
            napi_set_named_property(e, obj, "root", r) != napi_ok) return nullptr;
        return obj;
    };
    return launch(env, j, a.v, 2, "pil2gpu.commit");
}
// treeFromPages(ctx, elemPages, width, height, nodes) -> handle: a tree whose nodes are known (readFromFile, the const tree of a
// setup) goes to the device without re-hashing (pil2gpu_tree_alloc + pil2gpu_tree_fill).  Synchronous.
napi_value TreeFromPages(napi_env env, napi_callback_info info) {
    Args a(env, info, 5);
    pil2gpu_ctx* ctx = a.ctx(0);
    Pages e;
    a.pages(1, e.p, e.w, &e.total);
    const uint64_t width = a.u64(2), height = a.u64(3);
    uint64_t* nodes = nullptr; size_t nlen = 0;
    a.u64_array(4, &nodes, &nlen);
    a.need(height > 0 && e.total == width * height, "elements do not hold width * height words");
    a.need(nlen == pil2gpu_merkle_nnodes(height), "nodes must hold _getNNodes(height * 4) words");
    if (!a.ok) return nullptr;
    TreeBox* box = new TreeBox{ctx, nullptr, width, height};
    int rc = pil2gpu_tree_alloc(ctx, width, height, &box->tree);
    uint64_t off = 0;
    for (uint32_t k = 0; k < e.n() && !rc; k++) { rc = pil2gpu_tree_fill(ctx, box->tree, 0, off, e.p[k], e.w[k]); off += e.w[k]; }
    if (!rc) rc = pil2gpu_tree_fill(ctx, box->tree, 1, 0, nodes, nlen);
    if (rc) { tree_finalize(env, box, nullptr); return fail_now(env, rc); }
    return wrap_tree(env, box);
}
TreeBox* tree_arg(Args& a, size_t i) {
    TreeBox* b = (TreeBox*)a.external(i, "expected a device tree handle");
    if (b && !b->tree) { a.type_error("the device tree was freed"); return nullptr; }
    return b;
}
// treeRoot(tree) -> BigUint64Array(4)                                          merklehash_p.js:224
napi_value TreeRoot(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    TreeBox* b = tree_arg(a, 0);
    if (!a.ok) return nullptr;
    uint64_t r[4];
    int rc = pil2gpu_tree_root(b->ctx, b->tree, r);
    if (rc) return fail_now(env, rc);
    return make_u64_array(env, r, 4);
}
// treeInfo(tree) -> { width, height, depth }
napi_value TreeInfo(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    TreeBox* b = tree_arg(a, 0);
    if (!a.ok) return nullptr;
    napi_value obj, w, h, d;
    if (napi_create_object(env, &obj) != napi_ok || napi_create_double(env, (double)b->width, &w) != napi_ok ||
        napi_create_double(env, (double)b->height, &h) != napi_ok || napi_create_double(env, (double)pil2gpu_merkle_depth(b->height), &d) != napi_ok ||
        napi_set_named_property(env, obj, "width", w) != napi_ok || napi_set_named_property(env, obj, "height", h) != napi_ok ||
        napi_set_named_property(env, obj, "depth", d) != napi_ok) { napi_throw_error(env, nullptr, "could not build the result"); return nullptr; }
    return obj;
}
// treeGroupProofs(tree, idxs: BigUint64Array) -> { rows: BigUint64Array(n * width), siblings: BigUint64Array(n * depth * 4) }
// getGroupProof (merklehash_p.js:142-168) for a batch of leaves; an index >= height throws Error("Out of range").
napi_value TreeGroupProofs(napi_env env, napi_callback_info info) {
    Args a(env, info, 2);
    TreeBox* b = tree_arg(a, 0);
    uint64_t* idx = nullptr; size_t n = 0;
    a.u64_array(1, &idx, &n);
    a.need(n <= ((size_t)1 << 24), "too many indices");
    if (!a.ok) return nullptr;
    const uint32_t depth = pil2gpu_merkle_depth(b->height);
    std::vector<uint64_t> rows((size_t)n * b->width + 1), sib((size_t)n * depth * 4 + 1);
    int rc = pil2gpu_tree_group_proofs(b->ctx, b->tree, idx, (uint32_t)n, rows.data(), sib.data());
    if (rc) return fail_now(env, rc);
    napi_value obj, r = make_u64_array(env, rows.data(), (size_t)n * b->width), s = make_u64_array(env, sib.data(), (size_t)n * depth * 4);
    if (!r || !s || napi_create_object(env, &obj) != napi_ok || napi_set_named_property(env, obj, "rows", r) != napi_ok ||
        napi_set_named_property(env, obj, "siblings", s) != napi_ok) { napi_throw_error(env, nullptr, "could not build the result"); return nullptr; }
    return obj;
}
// treeDownload(tree, elemPages | null, nodes | null) -> Promise: back to the reference's host tree object on demand
napi_value TreeDownload(napi_env env, napi_callback_info info) {
    Args a(env, info, 3);
    TreeBox* b = tree_arg(a, 0);
    auto e = std::make_shared<Pages>();
    a.pages(1, e->p, e->w, &e->total, true);
    uint64_t* nodes = nullptr; size_t nlen = 0;
    a.opt_u64_array(2, &nodes, &nlen);
    if (a.ok) {
        a.need(e->n() == 0 || e->total == b->width * b->height, "elements pages do not hold width * height words");
        a.need(!nodes || nlen == pil2gpu_merkle_nnodes(b->height), "nodes must hold _getNNodes(height * 4) words");
    }
    if (!a.ok) return nullptr;
    Job* j = new Job;
    j->run = [=] {
        int rc = 0;
        if (nodes) rc = pil2gpu_tree_download(b->ctx, b->tree, nullptr, nodes);
        const uint64_t* src = pil2gpu_tree_elements_dev(b->tree);
        size_t off = 0;
        for (uint32_t k = 0; k < e->n() && !rc; k++) { rc = pil2gpu_d2h(b->ctx, e->p[k], src + off, e->w[k] * 8); off += e->w[k]; }
        if (!rc) rc = pil2gpu_sync(b->ctx);
        return rc;
    };
    return launch(env, j, a.v, 3, "pil2gpu.treeDownload");
}
// treeFree(tree): release the device memory now (otherwise when the handle is collected)
napi_value TreeFree(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    TreeBox* b = (TreeBox*)a.external(0, "expected a device tree handle");
    if (!a.ok) return nullptr;
    if (b->tree) { pil2gpu_tree_free(b->ctx, b->tree); b->tree = nullptr; }
    return nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// fri.js
// ---------------------------------------------------------------------------------------------------------------------
// friFoldPaged(ctx, polPages, prevBits, curBits, nextBits, step0Bits, challenge(3), split, polOutPages, rowsPages|null, nodes|null) -> Promise
napi_value FriFoldPaged(napi_env env, napi_callback_info info) {
    Args a(env, info, 11);
    pil2gpu_ctx* ctx = a.ctx(0);
    auto p = std::make_shared<Pages>(), o = std::make_shared<Pages>(), r = std::make_shared<Pages>();
    a.pages(1, p->p, p->w, &p->total);
    const uint32_t prevBits = a.u32(2), curBits = a.u32(3); const int32_t nextBits = a.i32(4); const uint32_t step0Bits = a.u32(5);
    uint64_t* ch = nullptr; size_t chl = 0;
    a.u64_array(6, &ch, &chl);
    const int32_t split = a.i32(7);
    a.pages(8, o->p, o->w, &o->total);
    a.pages(9, r->p, r->w, &r->total, true);
    uint64_t* nodes = nullptr; size_t nlen = 0;
    a.opt_u64_array(10, &nodes, &nlen);
    if (a.ok) {
        a.need(chl == 3, "the challenge is an F3 element (3 words)");
        a.need(prevBits <= 32 && curBits <= prevBits && prevBits <= step0Bits && step0Bits <= 32, "bad FRI step sizes");
        a.need(nextBits < 0 || (uint32_t)nextBits <= curBits, "bad next-layer size");
    }
    if (a.ok) {
        a.need(p->total == ((uint64_t)3 << prevBits), "pol does not hold 3 * 2^prevBits words");
        a.need(o->total == ((uint64_t)3 << curBits), "polOut does not hold 3 * 2^curBits words");
        a.need(r->n() == 0 || r->total == ((uint64_t)3 << curBits), "rows do not hold 3 * 2^curBits words");
        a.need(!nodes || (nextBits >= 0 && nlen == pil2gpu_merkle_nnodes((uint64_t)1 << nextBits)), "nodes must hold _getNNodes(2^nextBits * 4) words");
    }
    if (!a.ok) return nullptr;
    auto chal = std::make_shared<std::vector<uint64_t>>(ch, ch + 3);
    Job* j = new Job;
    j->run = [=] {
        int rc = pil2gpu_fri_fold_paged(ctx, p->cp(), p->w.data(), p->n(), prevBits, curBits, nextBits, step0Bits, chal->data(), split, o->mp(), o->w.data(), o->n(),
                                        r->n() ? r->mp() : nullptr, r->w.data(), r->n(), nodes);
        return rc;
    };
    napi_value keep[5] = {a.v[0], a.v[1], a.v[8], a.v[9], a.v[10]};
    return launch(env, j, keep, 5, "pil2gpu.friFold");
}

// ---------------------------------------------------------------------------------------------------------------------
// prover-side callers (src/stark/stark_gen_helpers.js)
// ---------------------------------------------------------------------------------------------------------------------
// extendAndMerkelizePaged(ctx, srcPages, nPols, nBits, nBitsExt, split, dstPages|null, nodes|null) -> Promise<root BigUint64Array(4)>   :388-412
napi_value ExtendAndMerkelizePaged(napi_env env, napi_callback_info info) {
    Args a(env, info, 8);
    pil2gpu_ctx* ctx = a.ctx(0);
    auto s = std::make_shared<Pages>(), d = std::make_shared<Pages>();
    a.pages(1, s->p, s->w, &s->total);
    const uint64_t nPols = a.u64(2); const uint32_t nBits = a.u32(3), nBitsExt = a.u32(4); const int32_t split = a.i32(5);
    a.pages(6, d->p, d->w, &d->total, true);
    uint64_t* nodes = nullptr; size_t nlen = 0;
    a.opt_u64_array(7, &nodes, &nlen);
    uint64_t sw = 0, dw = 0;
    a.need(nBitsExt <= 32 && nBits <= nBitsExt && nPols > 0 && shl_fits(nPols, nBits, &sw) && shl_fits(nPols, nBitsExt, &dw), "bad commit shape");
    if (a.ok) {
        a.need(s->total == sw, "cm_n does not hold nPols * 2^nBits elements");
        a.need(d->n() == 0 || d->total == dw, "cm_ext does not hold nPols * 2^nBitsExt elements");
        a.need(!nodes || nlen == pil2gpu_merkle_nnodes((uint64_t)1 << nBitsExt), "nodes must hold _getNNodes(extN * 4) words");
    }
    if (!a.ok) return nullptr;
    auto root = std::make_shared<std::vector<uint64_t>>(4);
    Job* j = new Job;
    j->run = [=] {
        int rc = pil2gpu_extend_and_merkelize_paged(ctx, s->cp(), s->w.data(), s->n(), nPols, nBits, nBitsExt, split, d->n() ? d->mp() : nullptr, d->w.data(),
                                                    d->n(), nodes, root->data());
        return rc;
    };
    j->result = [=](napi_env e) { return make_u64_array(e, root->data(), 4); };
    napi_value keep[4] = {a.v[0], a.v[1], a.v[6], a.v[7]};
    return launch(env, j, keep, 4, "pil2gpu.extendAndMerkelize");
}
// computeQPaged(ctx, qExtPages, qDim, qDeg, nBits, nBitsExt, split, cmqExtPages|null, nodes|null) -> Promise<root>   computeQStark :168-208
napi_value ComputeQPaged(napi_env env, napi_callback_info info) {
    Args a(env, info, 9);
    pil2gpu_ctx* ctx = a.ctx(0);
    auto q = std::make_shared<Pages>(), c = std::make_shared<Pages>();
    a.pages(1, q->p, q->w, &q->total);
    const uint64_t qDim = a.u64(2), qDeg = a.u64(3); const uint32_t nBits = a.u32(4), nBitsExt = a.u32(5); const int32_t split = a.i32(6);
    a.pages(7, c->p, c->w, &c->total, true);
    uint64_t* nodes = nullptr; size_t nlen = 0;
    a.opt_u64_array(8, &nodes, &nlen);
    uint64_t qw = 0, cw = 0;
    a.need(nBitsExt <= 32 && nBits <= nBitsExt && qDim > 0 && qDeg > 0 && qDim <= 64 && qDeg <= 256 && shl_fits(qDim, nBitsExt, &qw) &&
               shl_fits(qDim * qDeg, nBitsExt, &cw), "bad quotient shape");
    if (a.ok) {
        a.need(q->total == qw, "q_ext does not hold qDim * 2^nBitsExt elements");
        a.need(c->n() == 0 || c->total == cw, "cmQ_ext does not hold qDim * qDeg * 2^nBitsExt elements");
        a.need(!nodes || nlen == pil2gpu_merkle_nnodes((uint64_t)1 << nBitsExt), "nodes must hold _getNNodes(extN * 4) words");
    }
    if (!a.ok) return nullptr;
    auto root = std::make_shared<std::vector<uint64_t>>(4);
    Job* j = new Job;
    j->run = [=] {
        int rc = pil2gpu_compute_q_paged(ctx, q->cp(), q->w.data(), q->n(), qDim, qDeg, nBits, nBitsExt, split, c->n() ? c->mp() : nullptr, c->w.data(), c->n(),
                                         nodes, root->data());
        return rc;
    };
    j->result = [=](napi_env e) { return make_u64_array(e, root->data(), 4); };
    napi_value keep[4] = {a.v[0], a.v[1], a.v[7], a.v[8]};
    return launch(env, j, keep, 4, "pil2gpu.computeQ");
}
// computeEvals(ctx, xi(3), openings Int32Array, nBits, nBitsExt, buf, size, descs BigUint64Array(2*n: offset, dim | lev << 32)) -> BigUint64Array(3n)
// (only the 2^nBits base rows of buf are read: a single typed array of the extended buffer, or its first page when the buffer is
// one page; multi-page extended buffers are opened from a device tree instead -- computeEvalsStark :210-273)
napi_value ComputeEvals(napi_env env, napi_callback_info info) {
    Args a(env, info, 8);
    pil2gpu_ctx* ctx = a.ctx(0);
    uint64_t *xi = nullptr, *buf = nullptr, *d = nullptr; size_t xl = 0, bl = 0, nd = 0; int32_t* op = nullptr; size_t nop = 0;
    a.u64_array(1, &xi, &xl); a.i32_array(2, &op, &nop);
    const uint32_t nBits = a.u32(3), nBitsExt = a.u32(4);
    a.u64_array(5, &buf, &bl);
    const uint64_t size = a.u64(6);
    a.u64_array(7, &d, &nd);
    uint64_t bw = 0;
    a.need(xl == 3, "xi is an F3 element (3 words)");
    a.need(nBitsExt <= 32 && nBits <= nBitsExt && size > 0 && shl_fits(size, nBitsExt, &bw) && bl == bw, "buf does not hold size * 2^nBitsExt elements");
    a.need(nd % 2 == 0 && nop >= 1 && nop <= 64, "bad descriptor / opening lists");
    if (!a.ok) return nullptr;
    const uint32_t n = (uint32_t)(nd / 2);
    std::vector<pil2gpu_eval_desc> desc(n);
    for (uint32_t i = 0; i < n; i++) { desc[i].offset = d[2 * i]; desc[i].dim = (uint32_t)d[2 * i + 1]; desc[i].lev = (uint32_t)(d[2 * i + 1] >> 32); }
    std::vector<uint64_t> out((size_t)3 * n + 1);
    int rc = pil2gpu_compute_evals(ctx, xi, op, (uint32_t)nop, nBits, nBitsExt, buf, size, desc.data(), n, out.data());
    if (rc) return fail_now(env, rc);
    return make_u64_array(env, out.data(), (size_t)3 * n);
}
// xDivXSubXi(ctx, xi(3), openings Int32Array, nBits, nBitsExt, out)            computeFRIStark :289-323
napi_value XDivXSubXi(napi_env env, napi_callback_info info) {
    Args a(env, info, 6);
    pil2gpu_ctx* ctx = a.ctx(0);
    uint64_t *xi = nullptr, *out = nullptr; size_t xl = 0, ol = 0; int32_t* op = nullptr; size_t nop = 0;
    a.u64_array(1, &xi, &xl); a.i32_array(2, &op, &nop);
    const uint32_t nBits = a.u32(3), nBitsExt = a.u32(4);
    a.u64_array(5, &out, &ol);
    uint64_t ow = 0;
    a.need(xl == 3, "xi is an F3 element (3 words)");
    a.need(nBitsExt <= 32 && nBits <= nBitsExt && nop >= 1 && nop <= 64 && shl_fits(3 * (uint64_t)nop, nBitsExt, &ow) && ol == ow,
           "out does not hold 3 * nOpenings * 2^nBitsExt words");
    if (!a.ok) return nullptr;
    int rc = pil2gpu_x_div_x_sub_xi(ctx, xi, op, (uint32_t)nop, nBits, nBitsExt, out);
    return rc ? fail_now(env, rc) : nullptr;
}
// friPol(ctx, bufs: BigUint64Array[], meta: BigInt64Array(5 per term: buffer index, row size, offset, dim, prime), evals(3 per term),
//        openings Int32Array, xi(3), vf1(3), vf2(3), nBits, nBitsExt, fOut, xdivOut|null)      computeFRIStark :289-334
napi_value FriPol(napi_env env, napi_callback_info info) {
    Args a(env, info, 12);
    pil2gpu_ctx* ctx = a.ctx(0);
    Pages b;
    a.pages(1, b.p, b.w, &b.total);
    int64_t* meta = nullptr; size_t nmeta = 0; int32_t* op = nullptr; size_t nop = 0;
    uint64_t *ev = nullptr, *xi = nullptr, *v1 = nullptr, *v2 = nullptr, *fout = nullptr, *xdout = nullptr; size_t nev = 0, l1 = 0, l2 = 0, l3 = 0, fl = 0, xl = 0;
    a.i64_array(2, &meta, &nmeta); a.u64_array(3, &ev, &nev); a.i32_array(4, &op, &nop);
    a.u64_array(5, &xi, &l1); a.u64_array(6, &v1, &l2); a.u64_array(7, &v2, &l3);
    const uint32_t nBits = a.u32(8), nBitsExt = a.u32(9);
    a.u64_array(10, &fout, &fl);
    a.opt_u64_array(11, &xdout, &xl);
    a.need(l1 == 3 && l2 == 3 && l3 == 3, "xi, vf1 and vf2 are F3 elements (3 words)");
    a.need(nBitsExt <= 32 && nBits <= nBitsExt && nmeta % 5 == 0 && nmeta > 0 && nop >= 1 && nop <= 64, "bad term / opening lists");
    if (!a.ok) return nullptr;
    const uint32_t n = (uint32_t)(nmeta / 5);
    const uint64_t E = (uint64_t)1 << nBitsExt;
    a.need(nev == (size_t)3 * n, "evals must hold 3 words per term");
    a.need(fl == 3 * E, "f_ext does not hold 3 * 2^nBitsExt words");
    a.need(!xdout || xl == 3 * (uint64_t)nop * E, "xDivXSubXi_ext does not hold 3 * nOpenings * 2^nBitsExt words");
    std::vector<pil2gpu_fri_term> terms(n);
    for (uint32_t i = 0; a.ok && i < n; i++) {
        const int64_t bi = meta[5 * i], size = meta[5 * i + 1], off = meta[5 * i + 2], dim = meta[5 * i + 3];
        a.need(bi >= 0 && (size_t)bi < b.p.size(), "buffer index out of range");
        if (!a.ok) break;
        a.need(size > 0 && (dim == 1 || dim == 3) && off >= 0 && off + dim <= size, "term outside its row");
        a.need(b.w[(size_t)bi] == (uint64_t)size * E, "a term's buffer does not hold size * 2^nBitsExt words");
        terms[i].buf_dev = b.p[(size_t)bi]; terms[i].size = (uint64_t)size; terms[i].offset = (uint64_t)off;
        terms[i].dim = (uint32_t)dim; terms[i].prime = (int32_t)meta[5 * i + 4];
    }
    if (!a.ok) return nullptr;
    int rc = pil2gpu_fri_pol(ctx, terms.data(), n, ev, op, (uint32_t)nop, xi, v1, v2, nBits, nBitsExt, fout, xdout);
    return rc ? fail_now(env, rc) : nullptr;
}

// calculateExps(ctx, ops: Int32Array (16 words per record, the u32 bit patterns of pil2gpu_calculate_exps), consts: BigUint64Array (3 per
//               entry), bufs: BigUint64Array[] (2^domainBits rows each), meta: BigInt64Array (3 per buffer: row words, read, written),
//               domainBits, xShift)                                                  calculateExps, prover_helpers.js:33-76
// js/prover_helpers.js compiles the reference's {op, dest, src} records into `ops`; buffers flagged `written` are updated in place.
napi_value CalculateExps(napi_env env, napi_callback_info info) {
    Args a(env, info, 7);
    pil2gpu_ctx* ctx = a.ctx(0);
    int32_t* ops = nullptr; size_t nops = 0; uint64_t* consts = nullptr; size_t nc = 0; int64_t* meta = nullptr; size_t nmeta = 0;
    Pages b;
    a.i32_array(1, &ops, &nops); a.u64_array(2, &consts, &nc);
    a.pages(3, b.p, b.w, &b.total);
    a.i64_array(4, &meta, &nmeta);
    const uint32_t domainBits = a.u32(5), xShift = a.u32(6);
    a.need(nops > 0 && nops % 16 == 0, "ops holds 16 words per record");
    a.need(nc % 3 == 0, "consts holds 3 words per entry");
    a.need(domainBits <= 32 && nmeta == 3 * b.p.size() && b.p.size() <= 24, "bad buffer list (3 meta words per buffer, at most 24 buffers)");
    if (!a.ok) return nullptr;
    std::vector<pil2gpu_expr_host_buffer> hb(b.p.size());
    for (size_t i = 0; a.ok && i < hb.size(); i++) {
        const int64_t rw = meta[3 * i];
        uint64_t want = 0;
        a.need(rw > 0 && shl_fits((uint64_t)rw, domainBits, &want) && b.w[i] == want, "a buffer does not hold rowWords * 2^domainBits words");
        hb[i].ptr = b.p[i]; hb[i].row_words = (uint64_t)rw; hb[i].read = meta[3 * i + 1] != 0; hb[i].written = meta[3 * i + 2] != 0;
    }
    if (!a.ok) return nullptr;
    int rc = pil2gpu_calculate_exps(ctx, reinterpret_cast<const uint32_t*>(ops), (uint32_t)(nops / 16), consts, (uint32_t)(nc / 3), hb.data(),
                                    (uint32_t)hb.size(), domainBits, (int)xShift);
    return rc ? fail_now(env, rc) : nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-GPU commit group (pil2gpu_shard_*): one handle per GPU, wired across worker processes (shardHandles / shardConnect: the
// 128-byte handle pairs travel over the fork channel as BigUint64Array(16)) or inside one process (shardConnectLocal).
// shardCommit / shardOpen only ENQUEUE on the rank's stream and return at once -- the ranks meet in flag barriers on the GPUs, so a
// single thread can drive every GPU and no libuv worker ever waits for another; shardRoot / shardProofs are the asynchronous reads.
// ---------------------------------------------------------------------------------------------------------------------
struct ShardBox {
    pil2gpu_ctx* ctx;
    pil2gpu_shard* sh;
    uint32_t rank, world;
    uint64_t recv_words, stage_words;
    void *src = nullptr, *work = nullptr, *nodes = nullptr, *root = nullptr, *idx = nullptr, *rows = nullptr, *sib = nullptr;
    uint64_t src_words = 0, work_words = 0, nodes_words = 0, idx_words = 0, rows_words = 0, sib_words = 0;
    uint64_t nPols = 0; uint32_t nBitsExt = 0, n_idx = 0;
    int grow(void** p, uint64_t* have, uint64_t want) {
        if (*have >= want) return PIL2GPU_OK;
        if (*p) pil2gpu_dev_free(ctx, *p);
        *p = nullptr; *have = 0;
        int rc = pil2gpu_dev_alloc(ctx, (size_t)want * 8, p);
        if (!rc) *have = want;
        return rc;
    }
    void release() {
        if (sh) pil2gpu_shard_destroy(sh);
        sh = nullptr;
        void** all[] = {&src, &work, &nodes, &root, &idx, &rows, &sib};
        for (void** q : all) { if (*q) pil2gpu_dev_free(ctx, *q); *q = nullptr; }
    }
};
void shard_finalize(napi_env, void* data, void*) {
    ShardBox* b = (ShardBox*)data;
    b->release();
    delete b;
}
ShardBox* shard_arg(Args& a, size_t i) {
    ShardBox* b = (ShardBox*)a.external(i, "expected a shard handle (addon.shardCreate())");
    if (b && !b->sh) { a.type_error("the shard handle has been freed"); return nullptr; }
    return b;
}
// shardCreate(ctx, rank, world, recvWords, stageWords) -> handle
napi_value ShardCreate(napi_env env, napi_callback_info info) {
    Args a(env, info, 5);
    pil2gpu_ctx* ctx = a.ctx(0);
    const uint32_t rank = a.u32(1), world = a.u32(2);
    const uint64_t recv = a.u64(3), stage = a.u64(4);
    a.need(world >= 1 && world <= 16 && (world & (world - 1)) == 0 && rank < world, "world must be a power of two <= 16 and rank < world");
    a.need(recv <= ((uint64_t)1 << 40) && stage <= ((uint64_t)1 << 36), "buffer sizes out of range");
    if (!a.ok) return nullptr;
    ShardBox* b = new ShardBox{ctx, nullptr, rank, world, recv, stage};
    int rc = pil2gpu_shard_create(ctx, rank, world, recv, stage, &b->sh);
    if (!rc) rc = pil2gpu_dev_alloc(ctx, 32, &b->root);
    if (rc) { b->release(); delete b; return fail_now(env, rc); }
    napi_value ext;
    if (napi_create_external(env, b, shard_finalize, nullptr, &ext) != napi_ok) { shard_finalize(env, b, nullptr); napi_throw_error(env, nullptr, "napi_create_external failed"); return nullptr; }
    return ext;
}
// shardHandles(shard) -> BigUint64Array(16): the CUDA IPC handles of this rank's receive buffer and mailbox
napi_value ShardHandles(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    ShardBox* b = shard_arg(a, 0);
    if (!a.ok) return nullptr;
    uint64_t h[16];
    int rc = pil2gpu_shard_handles(b->sh, (uint8_t*)h);
    if (rc) return fail_now(env, rc);
    return make_u64_array(env, h, 16);
}
// shardConnect(shard, handles): handles = BigUint64Array(16 * world), the pairs of all ranks in rank order
napi_value ShardConnect(napi_env env, napi_callback_info info) {
    Args a(env, info, 2);
    ShardBox* b = shard_arg(a, 0);
    uint64_t* h = nullptr; size_t n = 0;
    a.u64_array(1, &h, &n);
    if (a.ok) a.need(n == 16 * (size_t)b->world, "handles must hold 16 words per rank");
    if (!a.ok) return nullptr;
    int rc = pil2gpu_shard_connect(b->sh, (const uint8_t*)h, b->world);
    return rc ? fail_now(env, rc) : nullptr;
}
// shardConnectLocal([shard0, shard1, ...]): every rank lives in this process
napi_value ShardConnectLocal(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    uint32_t n = 0; bool is_arr = false;
    if (a.ok && (napi_is_array(env, a.v[0], &is_arr) != napi_ok || !is_arr || napi_get_array_length(env, a.v[0], &n) != napi_ok)) a.type_error("expected an Array of shard handles");
    a.need(n >= 1 && n <= 16, "between 1 and 16 shards");
    std::vector<pil2gpu_shard*> g;
    for (uint32_t k = 0; a.ok && k < n; k++) {
        napi_value e; void* q = nullptr;
        if (napi_get_element(env, a.v[0], k, &e) != napi_ok || napi_get_value_external(env, e, &q) != napi_ok || !q || !((ShardBox*)q)->sh) { a.type_error("expected an Array of shard handles"); break; }
        g.push_back(((ShardBox*)q)->sh);
    }
    if (!a.ok) return nullptr;
    int rc = pil2gpu_shard_connect_local(g.data(), n);
    return rc ? fail_now(env, rc) : nullptr;
}
// shardCommit(shard, slabPages, nPols, nBits, nBitsExt, split): extendAndMerkelize over the group (stark_gen_helpers.js:388-412).  slabPages
// hold THIS rank's column slab, 2^nBits rows x nPols/world columns; nPols is the width of the whole trace.  Enqueues and returns.
napi_value ShardCommit(napi_env env, napi_callback_info info) {
    Args a(env, info, 6);
    ShardBox* b = shard_arg(a, 0);
    Pages s;
    a.pages(1, s.p, s.w, &s.total);
    const uint64_t nPols = a.u64(2); const uint32_t nBits = a.u32(3), nBitsExt = a.u32(4); const int32_t split = a.i32(5);
    uint64_t sw = 0, dw = 0;
    if (a.ok) {
        a.need(nBitsExt <= 32 && nBits <= nBitsExt && nPols > 0 && nPols % b->world == 0, "bad commit shape (nPols must be a multiple of the number of ranks)");
        a.need(shl_fits(nPols / b->world, nBits, &sw) && shl_fits(nPols / b->world, nBitsExt, &dw), "bad commit shape");
        a.need(s.total == sw, "slab does not hold (nPols / world) * 2^nBits elements");
        a.need(dw <= b->recv_words, "the shard's receive buffer is too small for this commit");
        a.need(((uint64_t)1 << nBitsExt) >= b->world, "fewer extended rows than ranks");
    }
    if (!a.ok) return nullptr;
    const uint64_t nn = pil2gpu_merkle_nnodes(((uint64_t)1 << nBitsExt) / b->world);
    int rc = b->grow(&b->src, &b->src_words, sw);
    if (!rc) rc = b->grow(&b->work, &b->work_words, dw);
    if (!rc) rc = b->grow(&b->nodes, &b->nodes_words, nn);
    size_t off = 0;
    for (uint32_t k = 0; k < s.n() && !rc; k++) { rc = pil2gpu_h2d(b->ctx, (uint64_t*)b->src + off, s.p[k], s.w[k] * 8); off += s.w[k]; }
    if (!rc) rc = pil2gpu_shard_commit_dev(b->sh, (const uint64_t*)b->src, (uint64_t*)b->work, nPols, nBits, nBitsExt, split, (uint64_t*)b->nodes, (uint64_t*)b->root);
    if (rc) return fail_now(env, rc);
    b->nPols = nPols; b->nBitsExt = nBitsExt;
    return nullptr;
}
// shardRoot(shard) -> Promise<BigUint64Array(4)>: waits for this rank's stream; rejects if a peer never reached a barrier
napi_value ShardRoot(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    ShardBox* b = shard_arg(a, 0);
    if (a.ok) a.need(b->nPols != 0, "no commit has been enqueued on this shard");
    if (!a.ok) return nullptr;
    auto root = std::make_shared<std::vector<uint64_t>>(4);
    Job* j = new Job;
    j->run = [=] {
        int rc = pil2gpu_d2h(b->ctx, root->data(), b->root, 32);
        if (!rc) rc = pil2gpu_shard_status(b->sh);
        return rc;
    };
    j->result = [=](napi_env e) -> napi_value { return make_u64_array(e, root->data(), 4); };
    return launch(env, j, a.v, 1, "pil2gpu.shardRoot");
}
// shardOpen(shard, idxs): getGroupProof (merklehash_p.js:142-168) of the last commit for global leaf indices (the same on every rank).  Enqueues.
napi_value ShardOpen(napi_env env, napi_callback_info info) {
    Args a(env, info, 2);
    ShardBox* b = shard_arg(a, 0);
    uint64_t* idx = nullptr; size_t n = 0;
    a.u64_array(1, &idx, &n);
    if (a.ok) {
        a.need(b->nPols != 0, "no commit has been enqueued on this shard");
        a.need(n >= 1 && n <= (1u << 20), "between 1 and 2^20 indices");
        a.need(n * (b->nPols + 4 * (uint64_t)b->nBitsExt) <= b->stage_words, "the shard's staging area is too small for this many proofs");
        for (size_t k = 0; a.ok && k < n; k++) if (idx[k] >> b->nBitsExt) { napi_throw_error(env, nullptr, "Out of range"); return nullptr; }
    }
    if (!a.ok) return nullptr;
    int rc = b->grow(&b->idx, &b->idx_words, n);
    if (!rc) rc = b->grow(&b->rows, &b->rows_words, n * b->nPols);
    if (!rc) rc = b->grow(&b->sib, &b->sib_words, n * 4 * (uint64_t)b->nBitsExt);
    if (!rc) rc = pil2gpu_h2d(b->ctx, b->idx, idx, n * 8);
    if (!rc) rc = pil2gpu_sync(b->ctx);                       // idx is the caller's array: do not let it go while the copy is pending
    if (!rc) rc = pil2gpu_shard_open_dev(b->sh, (const uint64_t*)b->nodes, b->nPols, b->nBitsExt, (const uint64_t*)b->idx, (uint32_t)n, (uint64_t*)b->rows, (uint64_t*)b->sib);
    if (rc) return fail_now(env, rc);
    b->n_idx = (uint32_t)n;
    return nullptr;
}
// shardProofs(shard) -> Promise<{ rows: BigUint64Array(n * nPols), siblings: BigUint64Array(n * nBitsExt * 4) }>
napi_value ShardProofs(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    ShardBox* b = shard_arg(a, 0);
    if (a.ok) a.need(b->n_idx != 0, "no shardOpen has been enqueued on this shard");
    if (!a.ok) return nullptr;
    const size_t nr = (size_t)b->n_idx * b->nPols, ns = (size_t)b->n_idx * 4 * b->nBitsExt;
    auto rows = std::make_shared<std::vector<uint64_t>>(nr), sib = std::make_shared<std::vector<uint64_t>>(ns);
    Job* j = new Job;
    j->run = [=] {
        int rc = pil2gpu_d2h(b->ctx, rows->data(), b->rows, nr * 8);
        if (!rc) rc = pil2gpu_d2h(b->ctx, sib->data(), b->sib, ns * 8);
        if (!rc) rc = pil2gpu_shard_status(b->sh);
        return rc;
    };
    j->result = [=](napi_env e) -> napi_value {
        napi_value obj, r = make_u64_array(e, rows->data(), nr), sv = make_u64_array(e, sib->data(), ns);
        if (!r || !sv || napi_create_object(e, &obj) != napi_ok || napi_set_named_property(e, obj, "rows", r) != napi_ok ||
            napi_set_named_property(e, obj, "siblings", sv) != napi_ok) return nullptr;
        return obj;
    };
    return launch(env, j, a.v, 1, "pil2gpu.shardProofs");
}
// shardFree(shard)
napi_value ShardFree(napi_env env, napi_callback_info info) {
    Args a(env, info, 1);
    ShardBox* b = (ShardBox*)a.external(0, "expected a shard handle (addon.shardCreate())");
    if (!a.ok) return nullptr;
    b->release();
    return nullptr;
}

napi_value Init(napi_env env, napi_value exports) {
    const napi_property_descriptor props[] = {
        {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"allocPinnedPage", nullptr, AllocPinnedPage, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"releaseWorkspace", nullptr, ReleaseWorkspace, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"merkleNNodes", nullptr, MerkleNNodes, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"nttPaged", nullptr, NttPaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"ldePaged", nullptr, LdePaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"merkelizePaged", nullptr, MerkelizePaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"poseidon", nullptr, Poseidon, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"linearHash", nullptr, LinearHash, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"commit", nullptr, Commit, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"treeFromPages", nullptr, TreeFromPages, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"treeRoot", nullptr, TreeRoot, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"treeInfo", nullptr, TreeInfo, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"treeGroupProofs", nullptr, TreeGroupProofs, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"treeDownload", nullptr, TreeDownload, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"treeFree", nullptr, TreeFree, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"friFoldPaged", nullptr, FriFoldPaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"extendAndMerkelizePaged", nullptr, ExtendAndMerkelizePaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeQPaged", nullptr, ComputeQPaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeEvals", nullptr, ComputeEvals, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"xDivXSubXi", nullptr, XDivXSubXi, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"friPol", nullptr, FriPol, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"calculateExps", nullptr, CalculateExps, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardCreate", nullptr, ShardCreate, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardHandles", nullptr, ShardHandles, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardConnect", nullptr, ShardConnect, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardConnectLocal", nullptr, ShardConnectLocal, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardCommit", nullptr, ShardCommit, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardRoot", nullptr, ShardRoot, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardOpen", nullptr, ShardOpen, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardProofs", nullptr, ShardProofs, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"shardFree", nullptr, ShardFree, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    if (napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props) != napi_ok) napi_throw_error(env, nullptr, "pil2gpu addon: registration failed");
    return exports;
}

}   // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
