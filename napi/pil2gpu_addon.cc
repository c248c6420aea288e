// Node N-API addon: marshals BigUint64Array / BigBuffer pages to the C ABI of include/pil2gpu.h.  No arithmetic here.
// Build (on a machine with Node >= 16 and node-gyp; this repo's build container has neither, so this file is
// compile-checked only where node_api.h exists):   cd napi && node-gyp configure build
// Every exported function maps to exactly one pil2gpu_* entry point; failures become JS Errors carrying
// pil2gpu_last_error() ("Out of range" for PIL2GPU_E_RANGE, as merklehash_p.js:143 throws).
#include <node_api.h>
#include <cstdint>
#include <vector>
#include "../include/pil2gpu.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, nullptr, "pil2gpu addon: N-API call failed: " #call); return nullptr; } } while (0)

static napi_value fail(napi_env env) { napi_throw_error(env, nullptr, pil2gpu_last_error()); return nullptr; }

static bool get_u64_array(napi_env env, napi_value v, uint64_t** data, size_t* len) {
    napi_typedarray_type ty; napi_value ab; size_t off;
    bool is_ta = false;
    if (napi_is_typedarray(env, v, &is_ta) != napi_ok || !is_ta) return false;
    if (napi_get_typedarray_info(env, v, &ty, len, (void**)data, &ab, &off) != napi_ok) return false;
    return ty == napi_biguint64_array;
}
static bool get_pages(napi_env env, napi_value arr, std::vector<uint64_t*>& pages, std::vector<uint64_t>& words) {
    uint32_t n = 0;
    if (napi_get_array_length(env, arr, &n) != napi_ok) return false;
    for (uint32_t i = 0; i < n; i++) {
        napi_value e; uint64_t* d; size_t l;
        if (napi_get_element(env, arr, i, &e) != napi_ok || !get_u64_array(env, e, &d, &l)) return false;
        pages.push_back(d); words.push_back(l);
    }
    return true;
}
static pil2gpu_ctx* get_ctx(napi_env env, napi_value v) { void* p = nullptr; napi_get_value_external(env, v, &p); return (pil2gpu_ctx*)p; }
static uint32_t u32_of(napi_env env, napi_value v) { uint32_t x = 0; napi_get_value_uint32(env, v, &x); return x; }
static int32_t i32_of(napi_env env, napi_value v) { int32_t x = 0; napi_get_value_int32(env, v, &x); return x; }
static uint64_t u64_of(napi_env env, napi_value v) { double d = 0; napi_get_value_double(env, v, &d); return (uint64_t)d; }

static void ctx_finalize(napi_env, void* data, void*) { pil2gpu_destroy((pil2gpu_ctx*)data); }

// create(device) -> external
static napi_value Create(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    pil2gpu_ctx* ctx = nullptr;
    if (pil2gpu_create(i32_of(env, a[0]), nullptr, &ctx)) return fail(env);
    napi_value ext; NAPI_OK(napi_create_external(env, ctx, ctx_finalize, nullptr, &ext));
    return ext;
}
// nttPaged(ctx, srcPages, dstPages, nPols, nBits, inverse): single-page buffers go straight to pil2gpu_ntt
static napi_value NttPaged(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    std::vector<uint64_t*> sp, dp; std::vector<uint64_t> sw, dw;
    if (!get_pages(env, a[1], sp, sw) || !get_pages(env, a[2], dp, dw)) { napi_throw_type_error(env, nullptr, "expected arrays of BigUint64Array"); return nullptr; }
    if (sp.size() != 1 || dp.size() != 1) { napi_throw_error(env, nullptr, "fft/ifft on multi-page BigBuffers: use interpolate or concatenate (pages > 2^28 words)"); return nullptr; }
    if (pil2gpu_ntt(get_ctx(env, a[0]), sp[0], dp[0], u64_of(env, a[3]), u32_of(env, a[4]), i32_of(env, a[5]))) return fail(env);
    return nullptr;
}
// ldePaged(ctx, srcPages, dstPages, nPols, nBits, nBitsExt) -> pil2gpu_lde_paged
static napi_value LdePaged(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    std::vector<uint64_t*> sp, dp; std::vector<uint64_t> sw, dw;
    if (!get_pages(env, a[1], sp, sw) || !get_pages(env, a[2], dp, dw)) { napi_throw_type_error(env, nullptr, "expected arrays of BigUint64Array"); return nullptr; }
    if (pil2gpu_lde_paged(get_ctx(env, a[0]), (const uint64_t* const*)sp.data(), sw.data(), (uint32_t)sp.size(), dp.data(), dw.data(),
                          (uint32_t)dp.size(), u64_of(env, a[3]), u32_of(env, a[4]), u32_of(env, a[5]))) return fail(env);
    return nullptr;
}
// merkleNNodes(heightBigInt) -> BigInt
static napi_value MerkleNNodes(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t h = 0; bool lossless = false; NAPI_OK(napi_get_value_bigint_uint64(env, a[0], &h, &lossless));
    napi_value r; NAPI_OK(napi_create_bigint_uint64(env, pil2gpu_merkle_nnodes(h), &r));
    return r;
}
// merkelizePaged(ctx, elemPages, width, height, split, nodes) -> pil2gpu_merkelize_paged
static napi_value MerkelizePaged(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    std::vector<uint64_t*> ep; std::vector<uint64_t> ew; uint64_t* nodes; size_t nlen;
    if (!get_pages(env, a[1], ep, ew) || !get_u64_array(env, a[5], &nodes, &nlen)) { napi_throw_type_error(env, nullptr, "expected BigUint64Array buffers"); return nullptr; }
    const uint64_t height = u64_of(env, a[3]);
    if (nlen < pil2gpu_merkle_nnodes(height)) { napi_throw_range_error(env, nullptr, "nodes buffer too small"); return nullptr; }
    if (pil2gpu_merkelize_paged(get_ctx(env, a[0]), (const uint64_t* const*)ep.data(), ew.data(), (uint32_t)ep.size(), u64_of(env, a[2]), height,
                                i32_of(env, a[4]), nodes)) return fail(env);
    return nullptr;
}
// poseidon(ctx, in12) -> BigUint64Array(12);  linearHash(ctx, vals, split) -> BigUint64Array(4)
static napi_value make_u64_array(napi_env env, const uint64_t* src, size_t n) {
    napi_value ab, ta; void* data;
    if (napi_create_arraybuffer(env, n * 8, &data, &ab) != napi_ok) return nullptr;
    for (size_t i = 0; i < n; i++) ((uint64_t*)data)[i] = src[i];
    if (napi_create_typedarray(env, napi_biguint64_array, n, ab, 0, &ta) != napi_ok) return nullptr;
    return ta;
}
static napi_value Poseidon(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t* in; size_t n; uint64_t out[12];
    if (!get_u64_array(env, a[1], &in, &n) || n != 12) { napi_throw_type_error(env, nullptr, "expected BigUint64Array(12)"); return nullptr; }
    if (pil2gpu_poseidon(get_ctx(env, a[0]), in, out)) return fail(env);
    return make_u64_array(env, out, 12);
}
static napi_value LinearHash(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t* in; size_t n; uint64_t out[4];
    if (!get_u64_array(env, a[1], &in, &n)) { napi_throw_type_error(env, nullptr, "expected BigUint64Array"); return nullptr; }
    if (pil2gpu_linear_hash(get_ctx(env, a[0]), in, n, i32_of(env, a[2]), out)) return fail(env);
    return make_u64_array(env, out, 4);
}
// friFold(ctx, pol, prevBits, curBits, nextBits, step0Bits, challenge, split, polOut, rowsOut|null, nodesOut|null)
static napi_value FriFold(napi_env env, napi_callback_info info) {
    size_t argc = 11; napi_value a[11]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t *pol, *ch, *po, *rows = nullptr, *nodes = nullptr; size_t l;
    if (!get_u64_array(env, a[1], &pol, &l) || !get_u64_array(env, a[6], &ch, &l) || l != 3 || !get_u64_array(env, a[8], &po, &l)) {
        napi_throw_type_error(env, nullptr, "expected BigUint64Array buffers"); return nullptr;
    }
    get_u64_array(env, a[9], &rows, &l);     // null on the last step
    get_u64_array(env, a[10], &nodes, &l);
    if (pil2gpu_fri_fold(get_ctx(env, a[0]), pol, u32_of(env, a[2]), u32_of(env, a[3]), i32_of(env, a[4]), u32_of(env, a[5]), ch, i32_of(env, a[7]),
                         po, rows, nodes)) return fail(env);
    return nullptr;
}

// ---- prover-side callers (src/stark/stark_gen_helpers.js) ----
// extendAndMerkelize(ctx, src, nPols, nBits, nBitsExt, split, dst, nodes) -> root BigUint64Array(4)   (:388-412; flat typed arrays)
static napi_value ExtendAndMerkelize(napi_env env, napi_callback_info info) {
    size_t argc = 8; napi_value a[8]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t *src, *dst = nullptr, *nodes = nullptr, root[4]; size_t l;
    if (!get_u64_array(env, a[1], &src, &l)) { napi_throw_type_error(env, nullptr, "expected BigUint64Array"); return nullptr; }
    get_u64_array(env, a[6], &dst, &l);      // null: keep the extended buffer off the host
    get_u64_array(env, a[7], &nodes, &l);
    if (pil2gpu_extend_and_merkelize(get_ctx(env, a[0]), src, u64_of(env, a[2]), u32_of(env, a[3]), u32_of(env, a[4]), i32_of(env, a[5]), dst, nodes,
                                     root)) return fail(env);
    return make_u64_array(env, root, 4);
}
// computeQ(ctx, qExt, qDim, qDeg, nBits, nBitsExt, split, cmqExt, nodes) -> root   (computeQStark :168-208)
static napi_value ComputeQ(napi_env env, napi_callback_info info) {
    size_t argc = 9; napi_value a[9]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t *q, *ext = nullptr, *nodes = nullptr, root[4]; size_t l;
    if (!get_u64_array(env, a[1], &q, &l)) { napi_throw_type_error(env, nullptr, "expected BigUint64Array"); return nullptr; }
    get_u64_array(env, a[7], &ext, &l);
    get_u64_array(env, a[8], &nodes, &l);
    if (pil2gpu_compute_q(get_ctx(env, a[0]), q, u64_of(env, a[2]), u64_of(env, a[3]), u32_of(env, a[4]), u32_of(env, a[5]), i32_of(env, a[6]), ext,
                          nodes, root)) return fail(env);
    return make_u64_array(env, root, 4);
}
// computeEvals(ctx, xi(3), openings Int32Array, nBits, nBitsExt, buf, size, descs BigUint64Array(2*n: offset, dim | lev << 32)) -> BigUint64Array(3n)
static napi_value ComputeEvals(napi_env env, napi_callback_info info) {
    size_t argc = 8; napi_value a[8]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t *xi, *buf, *d; size_t l, nd; int32_t* op; size_t nop; napi_typedarray_type ty; napi_value ab; size_t off;
    if (!get_u64_array(env, a[1], &xi, &l) || l != 3 || !get_u64_array(env, a[5], &buf, &l) || !get_u64_array(env, a[7], &d, &nd) ||
        napi_get_typedarray_info(env, a[2], &ty, &nop, (void**)&op, &ab, &off) != napi_ok || ty != napi_int32_array) {
        napi_throw_type_error(env, nullptr, "bad argument types"); return nullptr;
    }
    const uint32_t n = (uint32_t)(nd / 2);
    std::vector<pil2gpu_eval_desc> desc(n);
    for (uint32_t i = 0; i < n; i++) { desc[i].offset = d[2 * i]; desc[i].dim = (uint32_t)d[2 * i + 1]; desc[i].lev = (uint32_t)(d[2 * i + 1] >> 32); }
    std::vector<uint64_t> out((size_t)3 * n);
    if (pil2gpu_compute_evals(get_ctx(env, a[0]), xi, op, (uint32_t)nop, u32_of(env, a[3]), u32_of(env, a[4]), buf, u64_of(env, a[6]), desc.data(), n,
                              out.data())) return fail(env);
    return make_u64_array(env, out.data(), out.size());
}
// xDivXSubXi(ctx, xi(3), openings Int32Array, nBits, nBitsExt, out)   (computeFRIStark :289-323)
static napi_value XDivXSubXi(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint64_t *xi, *out; size_t l; int32_t* op; size_t nop; napi_typedarray_type ty; napi_value ab; size_t off;
    if (!get_u64_array(env, a[1], &xi, &l) || l != 3 || !get_u64_array(env, a[5], &out, &l) ||
        napi_get_typedarray_info(env, a[2], &ty, &nop, (void**)&op, &ab, &off) != napi_ok || ty != napi_int32_array) {
        napi_throw_type_error(env, nullptr, "bad argument types"); return nullptr;
    }
    if (pil2gpu_x_div_x_sub_xi(get_ctx(env, a[0]), xi, op, (uint32_t)nop, u32_of(env, a[3]), u32_of(env, a[4]), out)) return fail(env);
    return nullptr;
}
// friPol(ctx, bufs: BigUint64Array[], meta: BigInt64Array(5 per term: buffer index, row size, offset, dim, prime), evals(3 per term),
//        openings Int32Array, xi(3), vf1(3), vf2(3), nBits, nBitsExt, fOut, xdivOut|null)      (computeFRIStark :289-334)
static napi_value FriPol(napi_env env, napi_callback_info info) {
    size_t argc = 13; napi_value a[13]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    std::vector<uint64_t*> bp; std::vector<uint64_t> bw;
    uint64_t *ev, *xi, *v1, *v2, *fout, *xdout = nullptr; size_t l, nev; int64_t* meta; size_t nmeta; int32_t* op; size_t nop;
    napi_typedarray_type ty; napi_value ab; size_t off;
    if (!get_pages(env, a[1], bp, bw) || napi_get_typedarray_info(env, a[2], &ty, &nmeta, (void**)&meta, &ab, &off) != napi_ok || ty != napi_bigint64_array ||
        !get_u64_array(env, a[3], &ev, &nev) || napi_get_typedarray_info(env, a[4], &ty, &nop, (void**)&op, &ab, &off) != napi_ok || ty != napi_int32_array ||
        !get_u64_array(env, a[5], &xi, &l) || l != 3 || !get_u64_array(env, a[6], &v1, &l) || l != 3 || !get_u64_array(env, a[7], &v2, &l) || l != 3 ||
        !get_u64_array(env, a[10], &fout, &l)) {
        napi_throw_type_error(env, nullptr, "bad argument types"); return nullptr;
    }
    get_u64_array(env, a[11], &xdout, &l);
    const uint32_t n = (uint32_t)(nmeta / 5);
    if (nev != (size_t)3 * n) { napi_throw_range_error(env, nullptr, "evals must hold 3 words per term"); return nullptr; }
    std::vector<pil2gpu_fri_term> terms(n);
    for (uint32_t i = 0; i < n; i++) {
        const int64_t bi = meta[5 * i];
        if (bi < 0 || (size_t)bi >= bp.size()) { napi_throw_range_error(env, nullptr, "buffer index out of range"); return nullptr; }
        terms[i].buf_dev = bp[bi]; terms[i].size = (uint64_t)meta[5 * i + 1]; terms[i].offset = (uint64_t)meta[5 * i + 2];
        terms[i].dim = (uint32_t)meta[5 * i + 3]; terms[i].prime = (int32_t)meta[5 * i + 4];
    }
    if (pil2gpu_fri_pol(get_ctx(env, a[0]), terms.data(), n, ev, op, (uint32_t)nop, xi, v1, v2, u32_of(env, a[8]), u32_of(env, a[9]), fout, xdout))
        return fail(env);
    return nullptr;
}

static napi_value Init(napi_env env, napi_value exports) {
    const napi_property_descriptor props[] = {
        {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"nttPaged", nullptr, NttPaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"ldePaged", nullptr, LdePaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"merkleNNodes", nullptr, MerkleNNodes, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"merkelizePaged", nullptr, MerkelizePaged, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"poseidon", nullptr, Poseidon, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"linearHash", nullptr, LinearHash, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"friFold", nullptr, FriFold, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"extendAndMerkelize", nullptr, ExtendAndMerkelize, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeQ", nullptr, ComputeQ, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeEvals", nullptr, ComputeEvals, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"xDivXSubXi", nullptr, XDivXSubXi, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"friPol", nullptr, FriPol, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
