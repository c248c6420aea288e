// A small in-process model of the Node-API surface napi/pil2gpu_addon.cc uses, plus a driver that EXECUTES the addon's argument
// checking without Node.js: the addon is compiled against tests/stubs/node_api.h, linked with this file and libpil2gpu.so, and
// called with crafted arguments.  Test infrastructure (tests/test_boundary_cpu.py builds and runs it); not a Node replacement:
// values are plain structs, async work runs synchronously at queue time, nothing is garbage collected.
#include <node_api.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

struct napi_value__ {
    enum Kind { UNDEF, NUL, NUMBER, BIGINT, STRING, OBJECT, ARRAY, TYPEDARRAY, ARRAYBUFFER, EXTERNAL, ERROR_, PROMISE, FUNCTION } kind = UNDEF;
    double num = 0;
    uint64_t big = 0;
    std::string str;
    std::vector<napi_value> elems;
    std::map<std::string, napi_value> props;
    napi_typedarray_type ta_type = napi_uint8_array;
    void* data = nullptr;
    size_t length = 0;
    void* ext = nullptr;
    napi_callback cb = nullptr;
    int promise_state = 0;          // 0 pending, 1 resolved, 2 rejected
    napi_value settled = nullptr;
};
struct napi_env__ {
    std::string exc_type, exc_msg;   // pending exception ("" = none)
    std::map<std::string, napi_value> exports;
};
struct napi_callback_info__ { std::vector<napi_value> args; };
struct napi_ref__ { napi_value v; };
struct napi_deferred__ { napi_value promise; };
struct napi_async_work__ { napi_async_execute_callback ex; napi_async_complete_callback done; void* data; };

static napi_value mk(napi_value__::Kind k) { napi_value v = new napi_value__(); v->kind = k; return v; }
static napi_status thrown(napi_env env, const char* type, const char* msg) { if (env->exc_type.empty()) { env->exc_type = type; env->exc_msg = msg ? msg : ""; } return napi_ok; }

extern "C" {
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value*, void**) {
    const size_t cap = *argc;
    for (size_t i = 0; i < cap; i++) argv[i] = i < info->args.size() ? info->args[i] : mk(napi_value__::UNDEF);
    *argc = info->args.size();
    return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char*, const char* msg) { return thrown(env, "Error", msg); }
napi_status napi_throw_type_error(napi_env env, const char*, const char* msg) { return thrown(env, "TypeError", msg); }
napi_status napi_throw_range_error(napi_env env, const char*, const char* msg) { return thrown(env, "RangeError", msg); }
napi_status napi_is_typedarray(napi_env, napi_value v, bool* r) { *r = v && v->kind == napi_value__::TYPEDARRAY; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* type, size_t* length, void** data, napi_value* ab, size_t* off) {
    if (!v || v->kind != napi_value__::TYPEDARRAY) return napi_invalid_arg;
    if (type) *type = v->ta_type;
    if (length) *length = v->length;
    if (data) *data = v->data;
    if (ab) *ab = nullptr;
    if (off) *off = 0;
    return napi_ok;
}
napi_status napi_is_array(napi_env, napi_value v, bool* r) { *r = v && v->kind == napi_value__::ARRAY; return napi_ok; }
napi_status napi_get_array_length(napi_env, napi_value v, uint32_t* r) { if (!v || v->kind != napi_value__::ARRAY) return napi_invalid_arg; *r = (uint32_t)v->elems.size(); return napi_ok; }
napi_status napi_get_element(napi_env, napi_value v, uint32_t i, napi_value* r) { if (!v || v->kind != napi_value__::ARRAY || i >= v->elems.size()) return napi_invalid_arg; *r = v->elems[i]; return napi_ok; }
napi_status napi_get_value_external(napi_env, napi_value v, void** r) { if (!v || v->kind != napi_value__::EXTERNAL) return napi_invalid_arg; *r = v->ext; return napi_ok; }
napi_status napi_get_value_uint32(napi_env, napi_value v, uint32_t* r) { if (!v || v->kind != napi_value__::NUMBER) return napi_invalid_arg; *r = (uint32_t)(int64_t)v->num; return napi_ok; }
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t* r) { if (!v || v->kind != napi_value__::NUMBER) return napi_invalid_arg; *r = (int32_t)(int64_t)v->num; return napi_ok; }
napi_status napi_get_value_double(napi_env, napi_value v, double* r) { if (!v || v->kind != napi_value__::NUMBER) return napi_invalid_arg; *r = v->num; return napi_ok; }
napi_status napi_get_value_bigint_uint64(napi_env, napi_value v, uint64_t* r, bool* lossless) { if (!v || v->kind != napi_value__::BIGINT) return napi_invalid_arg; *r = v->big; *lossless = true; return napi_ok; }
napi_status napi_create_bigint_uint64(napi_env, uint64_t x, napi_value* r) { *r = mk(napi_value__::BIGINT); (*r)->big = x; return napi_ok; }
napi_status napi_create_external(napi_env, void* data, napi_finalize, void*, napi_value* r) { *r = mk(napi_value__::EXTERNAL); (*r)->ext = data; return napi_ok; }
napi_status napi_create_arraybuffer(napi_env, size_t n, void** data, napi_value* r) { *r = mk(napi_value__::ARRAYBUFFER); (*r)->data = calloc(n ? n : 1, 1); (*r)->length = n; *data = (*r)->data; return napi_ok; }
napi_status napi_create_external_arraybuffer(napi_env, void* p, size_t n, napi_finalize, void*, napi_value* r) { *r = mk(napi_value__::ARRAYBUFFER); (*r)->data = p; (*r)->length = n; return napi_ok; }
napi_status napi_create_typedarray(napi_env, napi_typedarray_type t, size_t len, napi_value ab, size_t off, napi_value* r) {
    *r = mk(napi_value__::TYPEDARRAY); (*r)->ta_type = t; (*r)->length = len; (*r)->data = (char*)ab->data + off; return napi_ok;
}
napi_status napi_define_properties(napi_env env, napi_value, size_t n, const napi_property_descriptor* p) {
    for (size_t i = 0; i < n; i++) { napi_value f = mk(napi_value__::FUNCTION); f->cb = p[i].method; env->exports[p[i].utf8name] = f; }
    return napi_ok;
}
napi_status napi_typeof(napi_env, napi_value v, napi_valuetype* r) {
    switch (v ? v->kind : napi_value__::UNDEF) {
        case napi_value__::UNDEF: *r = napi_undefined; break;
        case napi_value__::NUL: *r = napi_null; break;
        case napi_value__::NUMBER: *r = napi_number; break;
        case napi_value__::BIGINT: *r = napi_bigint; break;
        case napi_value__::STRING: *r = napi_string; break;
        case napi_value__::EXTERNAL: *r = napi_external; break;
        case napi_value__::FUNCTION: *r = napi_function; break;
        default: *r = napi_object; break;
    }
    return napi_ok;
}
napi_status napi_get_undefined(napi_env, napi_value* r) { *r = mk(napi_value__::UNDEF); return napi_ok; }
napi_status napi_create_object(napi_env, napi_value* r) { *r = mk(napi_value__::OBJECT); return napi_ok; }
napi_status napi_create_double(napi_env, double x, napi_value* r) { *r = mk(napi_value__::NUMBER); (*r)->num = x; return napi_ok; }
napi_status napi_set_named_property(napi_env, napi_value o, const char* k, napi_value v) { o->props[k] = v; return napi_ok; }
napi_status napi_create_string_utf8(napi_env, const char* s, size_t, napi_value* r) { *r = mk(napi_value__::STRING); (*r)->str = s; return napi_ok; }
napi_status napi_create_error(napi_env, napi_value, napi_value msg, napi_value* r) { *r = mk(napi_value__::ERROR_); (*r)->str = msg->str; return napi_ok; }
napi_status napi_create_reference(napi_env, napi_value v, uint32_t, napi_ref* r) { *r = new napi_ref__{v}; return napi_ok; }
napi_status napi_delete_reference(napi_env, napi_ref r) { delete r; return napi_ok; }
napi_status napi_create_promise(napi_env, napi_deferred* d, napi_value* p) { *p = mk(napi_value__::PROMISE); *d = new napi_deferred__{*p}; return napi_ok; }
napi_status napi_resolve_deferred(napi_env, napi_deferred d, napi_value v) { d->promise->promise_state = 1; d->promise->settled = v; delete d; return napi_ok; }
napi_status napi_reject_deferred(napi_env, napi_deferred d, napi_value v) { d->promise->promise_state = 2; d->promise->settled = v; delete d; return napi_ok; }
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback ex, napi_async_complete_callback done, void* data, napi_async_work* r) {
    *r = new napi_async_work__{ex, done, data}; return napi_ok;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work w) { w->ex(env, w->data); w->done(env, napi_ok, w->data); return napi_ok; }   // synchronous
napi_status napi_delete_async_work(napi_env, napi_async_work w) { delete w; return napi_ok; }
napi_value napi_register_module_v1(napi_env env, napi_value exports);
}

// ---- driver ------------------------------------------------------------------------------------------------------------
static napi_value num(double x) { napi_value v = mk(napi_value__::NUMBER); v->num = x; return v; }
static napi_value big(uint64_t x) { napi_value v = mk(napi_value__::BIGINT); v->big = x; return v; }
static napi_value nul() { return mk(napi_value__::NUL); }
static napi_value ext(void* p) { napi_value v = mk(napi_value__::EXTERNAL); v->ext = p; return v; }
static napi_value u64arr(size_t n) { napi_value v = mk(napi_value__::TYPEDARRAY); v->ta_type = napi_biguint64_array; v->length = n; v->data = calloc(n ? n : 1, 8); return v; }
static napi_value i64arr(size_t n) { napi_value v = mk(napi_value__::TYPEDARRAY); v->ta_type = napi_bigint64_array; v->length = n; v->data = calloc(n ? n : 1, 8); return v; }
static napi_value i32arr(size_t n) { napi_value v = mk(napi_value__::TYPEDARRAY); v->ta_type = napi_int32_array; v->length = n; v->data = calloc(n ? n : 1, 4); return v; }
static napi_value arr(std::vector<napi_value> e) { napi_value v = mk(napi_value__::ARRAY); v->elems = e; return v; }

static int failures = 0;
struct Result { std::string exc_type, exc_msg; napi_value value; };
static Result call(napi_env env, const char* name, std::vector<napi_value> args) {
    env->exc_type.clear(); env->exc_msg.clear();
    napi_callback_info__ info{args};
    auto it = env->exports.find(name);
    if (it == env->exports.end()) { printf("FAIL: %s is not exported\n", name); failures++; return Result{"missing", "", nullptr}; }
    napi_value v = it->second->cb(env, &info);
    return Result{env->exc_type, env->exc_msg, v};
}
static void expect_throw(napi_env env, const char* what, const char* name, std::vector<napi_value> args, const char* type, const char* needle) {
    Result r = call(env, name, args);
    const bool ok = r.exc_type == type && r.exc_msg.find(needle) != std::string::npos;
    printf("%s: %s -> %s(\"%s\")\n", ok ? "ok  " : "FAIL", what, r.exc_type.c_str(), r.exc_msg.c_str());
    if (!ok) failures++;
}

int main() {
    napi_env env = new napi_env__();
    napi_register_module_v1(env, mk(napi_value__::OBJECT));
    printf("registered %zu functions\n", env->exports.size());
    const char* must[] = {"create", "allocPinnedPage", "nttPaged", "ldePaged", "merkelizePaged", "extendAndMerkelizePaged", "computeQPaged", "friFoldPaged", "commit",
                          "treeRoot", "treeGroupProofs", "treeDownload", "treeFree", "treeFromPages", "poseidon", "linearHash", "merkleNNodes", "computeEvals",
                          "xDivXSubXi", "friPol", "calculateExps", "shardCreate", "shardHandles", "shardConnect", "shardConnectLocal", "shardCommit", "shardRoot", "shardOpen",
                          "shardProofs", "shardFree"};
    for (const char* m : must) if (!env->exports.count(m)) { printf("FAIL: %s missing\n", m); failures++; }
    napi_value fake_ctx = ext((void*)0x1);      // never dereferenced: every call below must be rejected before it reaches the library

    // merkleNNodes (pure): _getNNodes(8 * 4) = 60 words
    { Result r = call(env, "merkleNNodes", {num(8)}); const bool ok = r.exc_type.empty() && r.value && r.value->kind == napi_value__::BIGINT && r.value->big == 60;
      printf("%s: merkleNNodes(8) = %llu\n", ok ? "ok  " : "FAIL", r.value ? (unsigned long long)r.value->big : 0ULL); if (!ok) failures++; }
    { Result r = call(env, "merkleNNodes", {big(1 << 20)}); const bool ok = r.exc_type.empty() && r.value && r.value->big == 8u * (1 << 20) - 4;
      printf("%s: merkleNNodes(2^20n)\n", ok ? "ok  " : "FAIL"); if (!ok) failures++; }

    // wrong sizes are RangeErrors raised before the C call
    expect_throw(env, "nttPaged: dst one word short", "nttPaged", {fake_ctx, arr({u64arr(10), u64arr(6)}), arr({u64arr(15)}), num(2), num(3), num(0)}, "RangeError", "buffDst");
    expect_throw(env, "nttPaged: src too long", "nttPaged", {fake_ctx, arr({u64arr(17)}), arr({u64arr(16)}), num(2), num(3), num(0)}, "RangeError", "buffSrc");
    expect_throw(env, "nttPaged: nBits 40", "nttPaged", {fake_ctx, arr({u64arr(16)}), arr({u64arr(16)}), num(2), num(40), num(0)}, "RangeError", "does not fit");
    expect_throw(env, "ldePaged: dst for the wrong blowup", "ldePaged", {fake_ctx, arr({u64arr(16)}), arr({u64arr(32)}), num(2), num(3), num(5)}, "RangeError", "buffDst");
    expect_throw(env, "ldePaged: nBitsExt < nBits", "ldePaged", {fake_ctx, arr({u64arr(16)}), arr({u64arr(8)}), num(2), num(3), num(2)}, "RangeError", "shape");
    expect_throw(env, "merkelizePaged: nodes too small", "merkelizePaged", {fake_ctx, arr({u64arr(24)}), num(3), num(8), num(0), u64arr(59)}, "RangeError", "nodes");
    expect_throw(env, "merkelizePaged: elements short", "merkelizePaged", {fake_ctx, arr({u64arr(23)}), num(3), num(8), num(0), u64arr(60)}, "RangeError", "width * height");
    expect_throw(env, "extendAndMerkelizePaged: cm_ext wrong size", "extendAndMerkelizePaged",
                 {fake_ctx, arr({u64arr(16)}), num(2), num(3), num(4), num(0), arr({u64arr(31)}), nul()}, "RangeError", "cm_ext");
    expect_throw(env, "computeQPaged: q_ext wrong size", "computeQPaged", {fake_ctx, arr({u64arr(47)}), num(3), num(2), num(3), num(4), num(0), nul(), nul()}, "RangeError", "q_ext");
    expect_throw(env, "friFoldPaged: challenge of 2 words", "friFoldPaged",
                 {fake_ctx, arr({u64arr(48)}), num(4), num(2), num(-1), num(4), u64arr(2), num(0), arr({u64arr(12)}), nul(), nul()}, "RangeError", "challenge");
    expect_throw(env, "friFoldPaged: polOut wrong size", "friFoldPaged",
                 {fake_ctx, arr({u64arr(48)}), num(4), num(2), num(-1), num(4), u64arr(3), num(0), arr({u64arr(11)}), nul(), nul()}, "RangeError", "polOut");
    expect_throw(env, "xDivXSubXi: out wrong size", "xDivXSubXi", {fake_ctx, u64arr(3), i32arr(2), num(3), num(4), u64arr(95)}, "RangeError", "out");
    expect_throw(env, "computeEvals: buf wrong size", "computeEvals", {fake_ctx, u64arr(3), i32arr(2), num(3), num(4), u64arr(79), num(5), u64arr(2)}, "RangeError", "buf");
    expect_throw(env, "calculateExps: ops not a multiple of 16", "calculateExps", {fake_ctx, i32arr(17), u64arr(3), arr({u64arr(32)}), i64arr(3), num(4), num(1)}, "RangeError", "ops");
    { napi_value meta = i64arr(3); ((int64_t*)meta->data)[0] = 3; ((int64_t*)meta->data)[1] = 1;
      expect_throw(env, "calculateExps: buffer of the wrong size", "calculateExps", {fake_ctx, i32arr(16), u64arr(3), arr({u64arr(47)}), meta, num(4), num(1)}, "RangeError", "rowWords"); }
    expect_throw(env, "calculateExps: ops is a BigUint64Array", "calculateExps", {fake_ctx, u64arr(16), u64arr(3), arr({u64arr(32)}), i64arr(3), num(4), num(1)}, "TypeError", "Int32Array");
    expect_throw(env, "commit: source wrong size", "commit", {fake_ctx, arr({u64arr(15)}), num(2), num(3), num(4), num(0)}, "RangeError", "buffer");
    expect_throw(env, "nttPaged: fractional nPols", "nttPaged", {fake_ctx, arr({u64arr(16)}), arr({u64arr(16)}), num(2.5), num(3), num(0)}, "RangeError", "integer");
    // wrong types are TypeErrors; a null context is rejected
    expect_throw(env, "nttPaged: pages is a number", "nttPaged", {fake_ctx, num(5), arr({u64arr(16)}), num(2), num(3), num(0)}, "TypeError", "pages");
    expect_throw(env, "nttPaged: a page is an Int32Array", "nttPaged", {fake_ctx, arr({i32arr(16)}), arr({u64arr(16)}), num(2), num(3), num(0)}, "TypeError", "pages");
    expect_throw(env, "nttPaged: null context", "nttPaged", {ext(nullptr), arr({u64arr(16)}), arr({u64arr(16)}), num(2), num(3), num(0)}, "TypeError", "context");
    expect_throw(env, "nttPaged: context is a number", "nttPaged", {num(1), arr({u64arr(16)}), arr({u64arr(16)}), num(2), num(3), num(0)}, "TypeError", "context");
    expect_throw(env, "poseidon: 11 words", "poseidon", {fake_ctx, u64arr(11)}, "RangeError", "12");
    expect_throw(env, "nttPaged: nBits is a string-like object", "nttPaged", {fake_ctx, arr({u64arr(16)}), arr({u64arr(16)}), num(2), arr({}), num(0)}, "TypeError", "32-bit");
    expect_throw(env, "too few arguments", "ldePaged", {fake_ctx}, "TypeError", "too few");
    expect_throw(env, "treeRoot: not a tree handle", "treeRoot", {num(3)}, "TypeError", "tree");
    expect_throw(env, "shardCreate: world 3", "shardCreate", {fake_ctx, num(0), num(3), num(64), num(0)}, "RangeError", "power of two");
    expect_throw(env, "shardCreate: rank == world", "shardCreate", {fake_ctx, num(2), num(2), num(64), num(0)}, "RangeError", "rank < world");
    expect_throw(env, "shardCommit: not a shard", "shardCommit", {num(1), arr({u64arr(16)}), num(2), num(3), num(4), num(0)}, "TypeError", "shard");
    expect_throw(env, "shardConnectLocal: not an array", "shardConnectLocal", {num(1)}, "TypeError", "Array");
    // create(): without a CUDA device the library refuses (no CPU fallback) and the addon turns that into an Error
    { Result r = call(env, "create", {num(0)});
      if (!r.exc_type.empty()) { const bool ok = r.exc_type == "Error" && r.exc_msg.find("CUDA") != std::string::npos;
          printf("%s: create(0) without a GPU -> %s(\"%s\")\n", ok ? "ok  " : "FAIL", r.exc_type.c_str(), r.exc_msg.c_str()); if (!ok) failures++; }
      else {
          printf("ok  : create(0) succeeded (a GPU is present): functional checks through the addon\n");
          napi_value ctx = r.value;
          auto settled_ok = [&](const Result& q, const char* what) {
              const bool ok = q.exc_type.empty() && q.value && q.value->kind == napi_value__::PROMISE && q.value->promise_state == 1;
              if (!ok) { printf("FAIL: %s: %s %s %s\n", what, q.exc_type.c_str(), q.exc_msg.c_str(),
                                (q.value && q.value->settled) ? q.value->settled->str.c_str() : ""); failures++; }
              return ok;
          };
          // fft then ifft over ragged pages is the identity (fft_p.js:178-184)
          const uint32_t nBits = 10, nPols = 6; const size_t words = (size_t)nPols << nBits;
          napi_value s0 = u64arr(1000), s1 = u64arr(words - 1000), d0 = u64arr(words - 7), d1 = u64arr(7), b0 = u64arr(words);
          uint64_t x = 88172645463325252ULL;
          auto fill = [&](napi_value a) { for (size_t i = 0; i < a->length; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; ((uint64_t*)a->data)[i] = x % 0xFFFFFFFF00000001ULL; } };
          fill(s0); fill(s1);
          if (settled_ok(call(env, "nttPaged", {ctx, arr({s0, s1}), arr({d0, d1}), num(nPols), num(nBits), num(0)}), "nttPaged forward") &&
              settled_ok(call(env, "nttPaged", {ctx, arr({d0, d1}), arr({b0}), num(nPols), num(nBits), num(1)}), "nttPaged inverse")) {
              const bool same = !memcmp(b0->data, s0->data, 8000) && !memcmp((uint64_t*)b0->data + 1000, s1->data, (words - 1000) * 8);
              printf("%s: ifft(fft(x)) == x over ragged pages\n", same ? "ok  " : "FAIL"); if (!same) failures++;
          }
          // extendAndMerkelizePaged (host tree) and commit (device tree) agree on the root; a proof opened from the device tree carries the host row
          const uint32_t ext = 11; const size_t dw = (size_t)nPols << ext; const uint64_t nn = 8 * (1ull << ext) - 4;
          napi_value e0 = u64arr(dw / 2), e1 = u64arr(dw / 2), nodes = u64arr(nn);
          Result em = call(env, "extendAndMerkelizePaged", {ctx, arr({s0, s1}), num(nPols), num(nBits), num(ext), num(0), arr({e0, e1}), nodes});
          Result cm = call(env, "commit", {ctx, arr({s0, s1}), num(nPols), num(nBits), num(ext), num(0)});
          if (settled_ok(em, "extendAndMerkelizePaged") && settled_ok(cm, "commit")) {
              napi_value root = em.value->settled, obj = cm.value->settled;
              napi_value tree = obj->props["tree"], root2 = obj->props["root"];
              bool ok = root->length == 4 && root2->length == 4 && !memcmp(root->data, root2->data, 32) && !memcmp(root->data, (uint64_t*)nodes->data + nn - 4, 32);
              printf("%s: host tree root == device tree root == nodes.slice(-4)\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              Result tr = call(env, "treeRoot", {tree});
              ok = tr.exc_type.empty() && tr.value && !memcmp(tr.value->data, root->data, 32);
              printf("%s: treeRoot(handle)\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              napi_value idx = u64arr(2); ((uint64_t*)idx->data)[0] = 5; ((uint64_t*)idx->data)[1] = (1u << ext) - 1;
              Result gp = call(env, "treeGroupProofs", {tree, idx});
              ok = gp.exc_type.empty() && gp.value && gp.value->props["rows"]->length == 2 * nPols && gp.value->props["siblings"]->length == 2 * ext * 4 &&
                   !memcmp(gp.value->props["rows"]->data, (uint64_t*)e0->data + 5 * nPols, nPols * 8) &&
                   !memcmp(gp.value->props["siblings"]->data, (uint64_t*)nodes->data + 4 * 4, 32);      // sibling of leaf 5 at level 0 = digest of leaf 4
              printf("%s: treeGroupProofs(handle, [5, last]) rows and first sibling match the host tree\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              ((uint64_t*)idx->data)[0] = 1u << ext;
              Result oor = call(env, "treeGroupProofs", {tree, idx});
              ok = oor.exc_type == "Error" && oor.exc_msg == "Out of range";
              printf("%s: out-of-range index -> Error(\"Out of range\")\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              napi_value m_nodes = u64arr(nn);
              if (settled_ok(call(env, "merkelizePaged", {ctx, arr({e0, e1}), num(nPols), num(1u << ext), num(0), m_nodes}), "merkelizePaged")) {
                  ok = !memcmp(m_nodes->data, nodes->data, nn * 8);
                  printf("%s: merkelizePaged(cm_ext) == nodes of extendAndMerkelizePaged\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              }
              call(env, "treeFree", {tree});
              Result af = call(env, "treeRoot", {tree});
              ok = af.exc_type == "TypeError";
              printf("%s: a freed tree handle is rejected\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
          }
          // multi-GPU commit group, two ranks in this process (two contexts on device 0): shardCommit / shardOpen only enqueue, the ranks meet
          // in flag barriers on the GPU; root and proofs must equal the single-context device tree of the same trace
          {
              const uint32_t gB = 9, gE = 10, gC = 16, W = 2, cg = gC / W; const size_t gw = (size_t)gC << gB;
              napi_value full = u64arr(gw); fill(full);
              Result c1 = call(env, "create", {num(0)}), c2 = call(env, "create", {num(0)});
              const uint64_t recv = (uint64_t)cg << gE, stage = 4 * (gC + 4 * gE);
              Result s1 = call(env, "shardCreate", {c1.value, num(0), num(W), big(recv), big(stage)}), s2 = call(env, "shardCreate", {c2.value, num(1), num(W), big(recv), big(stage)});
              bool ok = s1.exc_type.empty() && s2.exc_type.empty() && s1.value && s2.value;
              if (ok) { Result cl = call(env, "shardConnectLocal", {arr({s1.value, s2.value})}); ok = cl.exc_type.empty(); }
              napi_value slab[2] = {u64arr((size_t)cg << gB), u64arr((size_t)cg << gB)};
              for (uint32_t g = 0; g < W; g++) for (size_t r = 0; r < ((size_t)1 << gB); r++)
                  memcpy((uint64_t*)slab[g]->data + r * cg, (uint64_t*)full->data + r * gC + g * cg, cg * 8);
              napi_value sh[2] = {s1.value, s2.value};
              for (uint32_t g = 0; ok && g < W; g++) { Result q = call(env, "shardCommit", {sh[g], arr({slab[g]}), num(gC), num(gB), num(gE), num(0)}); ok = q.exc_type.empty(); }
              Result whole = call(env, "commit", {ctx, arr({full}), num(gC), num(gB), num(gE), num(0)});
              ok = ok && settled_ok(whole, "commit (whole trace)");
              napi_value idx = u64arr(4); uint64_t qs[4] = {0, (1u << gE) - 1, 1u << (gE - 1), (1u << (gE - 1)) - 1}; memcpy(idx->data, qs, 32);
              for (uint32_t g = 0; ok && g < W; g++) {
                  Result rr = call(env, "shardRoot", {sh[g]});
                  ok = settled_ok(rr, "shardRoot") && !memcmp(rr.value->settled->data, whole.value->settled->props["root"]->data, 32);
              }
              printf("%s: shard group of 2 ranks: root on every rank == root of the whole-trace commit\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              for (uint32_t g = 0; ok && g < W; g++) { Result q = call(env, "shardOpen", {sh[g], idx}); ok = q.exc_type.empty(); }
              if (ok) {
                  Result gp = call(env, "treeGroupProofs", {whole.value->settled->props["tree"], idx});
                  for (uint32_t g = 0; ok && g < W; g++) {
                      Result pr = call(env, "shardProofs", {sh[g]});
                      ok = settled_ok(pr, "shardProofs") && gp.exc_type.empty() &&
                           pr.value->settled->props["rows"]->length == 4 * gC && pr.value->settled->props["siblings"]->length == 4 * gE * 4 &&
                           !memcmp(pr.value->settled->props["rows"]->data, gp.value->props["rows"]->data, 4 * gC * 8) &&
                           !memcmp(pr.value->settled->props["siblings"]->data, gp.value->props["siblings"]->data, 4 * gE * 4 * 8);
                  }
              }
              printf("%s: shard group: rows + sibling paths on every rank == treeGroupProofs of the whole tree\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              ((uint64_t*)idx->data)[1] = 1u << gE;
              Result oor = call(env, "shardOpen", {sh[0], idx});
              ok = oor.exc_type == "Error" && oor.exc_msg == "Out of range";
              printf("%s: shardOpen: out-of-range index -> Error(\"Out of range\")\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              Result big_commit = call(env, "shardCommit", {sh[0], arr({u64arr((size_t)cg << (gB + 1))}), num(gC), num(gB + 1), num(gE + 1), num(0)});
              ok = big_commit.exc_type == "RangeError" && big_commit.exc_msg.find("receive buffer") != std::string::npos;
              printf("%s: shardCommit larger than the receive buffer -> RangeError\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
              call(env, "shardFree", {sh[0]}); call(env, "shardFree", {sh[1]});
              Result af = call(env, "shardRoot", {sh[0]});
              ok = af.exc_type == "TypeError";
              printf("%s: a freed shard handle is rejected\n", ok ? "ok  " : "FAIL"); if (!ok) failures++;
          }
          napi_value pg = call(env, "allocPinnedPage", {num(4096)}).value;
          const bool pok = pg && pg->kind == napi_value__::TYPEDARRAY && pg->length == 4096 && ((uint64_t*)pg->data)[4095] == 0;
          printf("%s: allocPinnedPage(4096) is a zeroed BigUint64Array\n", pok ? "ok  " : "FAIL"); if (!pok) failures++;
      } }
    printf(failures ? "FAILED %d checks\n" : "ALL CHECKS PASSED\n", failures);
    return failures ? 1 : 0;
}
