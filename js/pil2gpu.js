// Loader + marshalling helpers shared by the drop-in modules (fft_p.js, merklehash_p.js, fri.js, stark_gen_helpers.js).
// The addon (napi/pil2gpu_addon.cc) exposes the C ABI of include/pil2gpu.h; calls that move buffers run on a libuv worker and
// return Promises (the reference functions are async too: fft_p.js:178-187, merklehash_p.js:44).
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

// ---- BigBuffer pages ---------------------------------------------------------------------------------------------------
// A pilcom BigBuffer keeps its data as a list of BigUint64Array pages.  The reference only uses its public surface
// (`length`, `slice`, `set`, `getElement`, `setElement`); the page list itself (`buffers`) is an implementation detail, so it is
// used zero-copy only when it is present AND consistent (typed-array pages whose lengths add up to `length`).  Otherwise the
// buffer is bridged through the public surface with pinned bounce pages: `slice` in, `set` out.
const BOUNCE_WORDS = 1 << 27;     // 1 GiB bounce pages

function zeroCopyPages(buff) {
    if (buff instanceof BigUint64Array) return [buff];
    const pages = buff && buff.buffers;
    if (!Array.isArray(pages)) return null;
    let total = 0;
    for (const p of pages) {
        if (!(p instanceof BigUint64Array)) return null;
        total += p.length;
    }
    return total === buff.length ? pages : null;
}

// Pages holding the CONTENTS of buff (an input of a call).
function inPages(buff) {
    const z = zeroCopyPages(buff);
    if (z) return z;
    if (typeof buff.slice !== "function" || typeof buff.length !== "number") throw new Error("pil2gpu: expected a BigBuffer or BigUint64Array");
    const pages = [];
    for (let o = 0; o < buff.length; o += BOUNCE_WORDS) {
        const n = Math.min(BOUNCE_WORDS, buff.length - o);
        const pg = addon.allocPinnedPage(n);
        pg.set(buff.slice(o, o + n));
        pages.push(pg);
    }
    return pages;
}

// Pages a call writes into, plus commit(): copies them back into buff when they are bounce pages (no-op when zero-copy).
function outPages(buff) {
    const z = zeroCopyPages(buff);
    if (z) return { pages: z, commit() {} };
    if (typeof buff.set !== "function" || typeof buff.length !== "number") throw new Error("pil2gpu: expected a BigBuffer or BigUint64Array");
    const pages = [];
    for (let o = 0; o < buff.length; o += BOUNCE_WORDS) pages.push(addon.allocPinnedPage(Math.min(BOUNCE_WORDS, buff.length - o)));
    return { pages, commit() { let o = 0; for (const pg of pages) { buff.set(pg, o); o += pg.length; } } };
}

// A BigBuffer-compatible buffer over PINNED pages (copied by DMA directly, no staging): same surface as pilcom's BigBuffer, so it
// can replace `new BigBuffer(n)` at the allocation sites of the prover context (stark_gen_helpers.js:104-109).
class PinnedBigBuffer {
    constructor(size, pageWords = 1 << 28) {
        this.length = size;
        this.pageWords = pageWords;
        this.buffers = [];
        for (let o = 0; o < size; o += pageWords) this.buffers.push(addon.allocPinnedPage(Math.min(pageWords, size - o)));
    }
    getElement(i) { return this.buffers[Math.floor(i / this.pageWords)][i % this.pageWords]; }
    setElement(i, v) { this.buffers[Math.floor(i / this.pageWords)][i % this.pageWords] = v; }
    slice(from = 0, to = this.length) {
        if (from < 0) from += this.length;
        if (to < 0) to += this.length;
        from = Math.max(0, Math.min(from, this.length)); to = Math.max(from, Math.min(to, this.length));
        const out = new BigUint64Array(to - from);
        for (let pos = from; pos < to;) {
            const p = Math.floor(pos / this.pageWords), o = pos % this.pageWords, n = Math.min(this.pageWords - o, to - pos);
            out.set(this.buffers[p].subarray(o, o + n), pos - from);
            pos += n;
        }
        return out;
    }
    set(arr, offset = 0) {
        for (let done = 0; done < arr.length;) {
            const pos = offset + done, p = Math.floor(pos / this.pageWords), o = pos % this.pageWords, n = Math.min(this.pageWords - o, arr.length - done);
            this.buffers[p].set(arr.subarray(done, done + n), o);
            done += n;
        }
    }
}

module.exports = { addon, context, inPages, outPages, zeroCopyPages, PinnedBigBuffer };
