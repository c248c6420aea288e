// Loader + marshalling helpers shared by the three drop-in modules (fft_p.js, merklehash_p.js, fri.js).
// The addon (napi/pil2gpu_addon.cc) exposes the C ABI of include/pil2gpu.h one-to-one; every call is synchronous on the
// C side and wrapped in a Promise here because the reference functions are async (fft_p.js:178-187, merklehash_p.js:44).
// There is no JavaScript fallback: if the addon cannot be loaded, requiring this file throws.
"use strict";
const path = require("path");

let addon;
try {
    addon = require(process.env.PIL2GPU_ADDON || path.join(__dirname, "..", "napi", "build", "Release", "pil2gpu_addon.node"));
} catch (e) {
    throw new Error("pil2gpu: native addon not found (build napi/ with node-gyp; there is no CPU fallback): " + e.message);
}

let ctx = null;
function context() {
    if (!ctx) ctx = addon.create(Number(process.env.PIL2GPU_DEVICE || 0));   // pil2gpu_create; throws Error(last_error) on failure
    return ctx;
}

// pilcom BigBuffer = list of BigUint64Array pages (`buffers`); a plain BigUint64Array is the one-page case.
function pagesOf(buff) {
    if (buff instanceof BigUint64Array) return [buff];
    if (Array.isArray(buff.buffers)) return buff.buffers;
    throw new Error("pil2gpu: expected a BigBuffer or BigUint64Array");
}

module.exports = { addon, context, pagesOf };
