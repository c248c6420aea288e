// GPU replacement for the expression evaluators of src/prover/prover_helpers.js: callCalculateExps / calculateExps (:23-76) over the
// domains "n" and "ext".  The reference turns a program of {op, dest, src} records into the body of a JavaScript function
// (compileCode :87-110) and calls it once per row; here compileCode resolves the same records (operand references as in getRef /
// setRef / evalMap :112-265) into the fixed 16-word records of pil2gpu_calculate_exps (include/pil2gpu.h) and the library compiles
// them into a kernel at run time (one thread per row).  Twin of pil2_stark_js_b200/prover_helpers.py, which is what the tests run.
// Usage: delegate from the reference's module, e.g.
//     module.exports.callCalculateExps = require(".../js/prover_helpers.js").callCalculateExps;
// ctx is the reference's prover context (pilInfo, nBits, nBitsExt, challenges, publics, evals, subproofValues, const_n / const_ext,
// cm<stage>_n / cm<stage>_ext, Zi_ext, xDivXSubXi_ext, q_ext, f_ext); buffers are single-page BigBuffers or BigUint64Arrays here
// (at sizes where they are multi-page keep them on the device and call pil2gpu_calculate_exps_dev on the tree buffers).
"use strict";
const { addon, context, zeroCopyPages } = require("./pil2gpu.js");

const P = 0xFFFFFFFF00000001n;
const OPCODES = { add: 0, sub: 1, mul: 2, copy: 3, muladd: 4 };
const K_TMP = 0, K_CONST = 1, K_BUF = 2, K_X = 3;
const OP_WORDS = 16, MAX_SLOTS = 64;

function f3(v) {
    if (Array.isArray(v)) return [v.map((x) => ((BigInt(x) % P) + P) % P), 3];
    return [[((BigInt(v) % P) + P) % P, 0n, 0n], 1];
}
// prover_helpers.js:201-215
function ziIndex(info, boundaryId) {
    const b = info.boundaries[boundaryId];
    for (let k = 0; k < info.boundaries.length; k++) {
        const o = info.boundaries[k];
        if (b.name === "everyFrame") {
            if (o.name === "everyFrame" && o.offsetMin === b.offsetMin && o.offsetMax === b.offsetMax) return k;
        } else if (o.name === b.name) return k;
    }
    throw new Error("Something went wrong");
}

// compileCode (prover_helpers.js:87-110) for the GPU: returns { ops: Uint32Array, consts: BigUint64Array, buffers: [[name, rowWords]],
// written: Set(name), nSlots }.  Temporaries are packed into slots by liveness (the reference's tmp ids are single-assignment).
function compileCode(ctx, code, dom) {
    if (dom !== "n" && dom !== "ext") throw new Error("Invalid dom");
    const info = ctx.pilInfo;
    const N = 2 ** (dom === "n" ? ctx.nBits : ctx.nBitsExt);
    const extendBits = ctx.nBitsExt - ctx.nBits;
    const consts = [], constIndex = new Map();
    const buffers = [], bufIndex = new Map(), written = new Set();
    const constOf = (value) => {
        const [v, dim] = f3(value);
        const key = v.join(",") + ":" + dim;
        if (!constIndex.has(key)) { constIndex.set(key, consts.length); consts.push(v); }
        return [constIndex.get(key), dim];
    };
    const bufOf = (name, rowWords) => {
        if (!bufIndex.has(name)) { bufIndex.set(name, buffers.length); buffers.push([name, Number(rowWords)]); }
        return bufIndex.get(name);
    };
    const rowOffset = (prime) => {
        if (!prime) return 0;
        const nxt = dom === "n" ? prime : prime * 2 ** extendBits;          // getRef :158-166 / evalMap :227-233, reduced mod N
        return ((nxt % N) + N) % N;
    };
    const polRef = (polId) => {
        const p = info.cmPolsMap[polId];
        const st = "cm" + p.stage;
        return [st + "_" + dom, p.stagePos, info.mapSectionsN[st], p.dim];
    };
    // validation first (on the program as given: dead records are checked too)
    const seen = new Set();
    for (const c of code) {
        if (!(c.op in OPCODES)) throw new Error("Invalid op:" + c.op);
        const nsrc = c.op === "copy" ? 1 : (c.op === "muladd" ? 3 : 2);
        if (c.src.length !== nsrc) throw new Error("op " + c.op + " takes " + nsrc + " sources");
        for (const s of c.src) if (s.type === "tmp" && !seen.has(s.id)) throw new Error("temporary " + s.id + " read before it is written");
        if (c.dest.type === "tmp") seen.add(c.dest.id);
    }
    // dead records: temporaries nobody reads (the code generator leaves some behind) are dropped, transitively
    for (;;) {
        const read = new Set();
        code.forEach((c) => c.src.forEach((s) => { if (s.type === "tmp") read.add(s.id); }));
        const live = code.filter((c) => c.dest.type !== "tmp" || read.has(c.dest.id));
        if (live.length === code.length) break;
        code = live;
    }
    const lastUse = new Map();
    code.forEach((c, k) => c.src.forEach((s) => { if (s.type === "tmp") lastUse.set(s.id, k); }));
    const slotOf = new Map(), free = [], tmpDim = new Map();
    let nSlots = 0;
    const operand = (r, isDest) => {
        const t = r.type;
        if (t === "tmp") {
            if (isDest) {
                let slot;
                if (slotOf.has(r.id)) slot = slotOf.get(r.id);               // re-assignment of a live temporary keeps its slot
                else {
                    slot = free.length ? free.pop() : nSlots;
                    if (slot === nSlots) nSlots++;
                    slotOf.set(r.id, slot);
                }
                tmpDim.set(r.id, r.dim === undefined ? 3 : r.dim);
                return [K_TMP | (tmpDim.get(r.id) << 8), slot, 0];
            }
            if (!slotOf.has(r.id)) throw new Error("temporary " + r.id + " read before it is written");
            return [K_TMP | (tmpDim.get(r.id) << 8), slotOf.get(r.id), 0];
        }
        if (t === "const") {
            const b = bufOf("const_" + dom, info.nConstants);
            return [K_BUF | (1 << 8) | (b << 16), r.id, rowOffset(r.prime || 0)];
        }
        if (t === "cm") {
            const [name, off, size, dim] = polRef(r.id);
            const b = bufOf(name, size);
            if (isDest) written.add(name);
            return [K_BUF | (dim << 8) | (b << 16), off, rowOffset(r.prime || 0)];
        }
        if (t === "q" || t === "f") {
            if (!isDest || dom !== "ext") throw new Error("Accessing " + t + " in domain " + dom);
            const dim = t === "f" ? 3 : (r.dim === undefined ? 3 : r.dim);
            const b = bufOf(t + "_ext", dim);
            written.add(t + "_ext");
            return [K_BUF | (dim << 8) | (b << 16), 0, 0];
        }
        if (isDest) throw new Error("Invalid reference type set: " + t);
        let i, dim;
        if (t === "number") [i, dim] = constOf(BigInt(r.value));
        else if (t === "public") [i, dim] = constOf(ctx.publics[r.id]);
        else if (t === "challenge") [i, dim] = constOf(ctx.challenges[r.stage - 1][r.stageId]);
        else if (t === "subproofValue") [i, dim] = constOf(ctx.subproofValues[r.id]);
        else if (t === "eval") [i, dim] = constOf(ctx.evals[r.id]);
        else if (t === "xDivXSubXi") {
            const b = bufOf("xDivXSubXi_ext", 3 * info.openingPoints.length);
            return [K_BUF | (3 << 8) | (b << 16), 3 * r.id, 0];
        } else if (t === "x") return [K_X | (1 << 8), 0, 0];
        else if (t === "Zi") {
            const b = bufOf("Zi_ext:" + ziIndex(info, r.boundaryId), 1);      // Zi_ext is [boundary][row]: one single-column buffer per table
            return [K_BUF | (1 << 8) | (b << 16), 0, 0];
        } else throw new Error("Invalid reference type get: " + t);
        return [K_CONST | (dim << 8), i, 0];
    };
    const ops = new Uint32Array(code.length * OP_WORDS);
    code.forEach((c, k) => {
        if (!(c.op in OPCODES)) throw new Error("Invalid op:" + c.op);
        const nsrc = c.op === "copy" ? 1 : (c.op === "muladd" ? 3 : 2);
        if (c.src.length !== nsrc) throw new Error("op " + c.op + " takes " + nsrc + " sources");
        const rec = ops.subarray(k * OP_WORDS, (k + 1) * OP_WORDS);
        rec[0] = OPCODES[c.op]; rec[1] = nsrc;
        const srcs = c.src.map((s) => operand(s, false));
        c.src.forEach((s) => {                                              // slots whose last reader is this record are free again
            if (s.type === "tmp" && lastUse.get(s.id) === k && slotOf.has(s.id)) { free.push(slotOf.get(s.id)); slotOf.delete(s.id); }
        });
        rec.set(operand(c.dest, true), 4);
        srcs.forEach((s, j) => rec.set(s, 7 + 3 * j));
    });
    if (nSlots > MAX_SLOTS) throw new Error("the program keeps " + nSlots + " temporaries alive at once (at most " + MAX_SLOTS + ")");
    const carr = new BigUint64Array(consts.length * 3);
    consts.forEach((v, i) => v.forEach((x, j) => { carr[3 * i + j] = BigInt(x); }));
    return { ops, consts: carr, buffers, written, nSlots, dom };
}

function hostBuffer(ctx, name) {
    if (name.startsWith("Zi_ext:")) {
        const k = Number(name.split(":")[1]), E = 2 ** ctx.nBitsExt;
        const z = ctx.Zi_ext instanceof BigUint64Array ? ctx.Zi_ext : zeroCopyPages(ctx.Zi_ext)[0];
        return z.subarray(k * E, (k + 1) * E);
    }
    const b = ctx[name];
    if (b instanceof BigUint64Array) return b;
    const p = zeroCopyPages(b);
    if (!p || p.length !== 1) throw new Error("pil2gpu: " + name + " must be a single-page buffer here (keep multi-page buffers on the device)");
    return p[0];
}

// calculateExps (prover_helpers.js:33-76) with ret == false: runs the program at every row of the domain and updates the destination
// buffers of ctx in place.
function calculateExps(ctx, code, dom, debug, ret) {
    if (debug || ret) throw new Error("pil2gpu: the debug / ret modes of calculateExps return per-row values; use the reference for them");
    const records = Array.isArray(code) ? code : code.code;
    const cc = compileCode(ctx, records, dom);
    const bufs = cc.buffers.map(([name]) => hostBuffer(ctx, name));
    const meta = new BigInt64Array(3 * cc.buffers.length);
    cc.buffers.forEach(([name, rowWords], i) => { meta[3 * i] = BigInt(rowWords); meta[3 * i + 1] = 1n; meta[3 * i + 2] = cc.written.has(name) ? 1n : 0n; });
    addon.calculateExps(context(), new Int32Array(cc.ops.buffer), cc.consts, bufs, meta, dom === "n" ? ctx.nBits : ctx.nBitsExt, dom === "ext" ? 1 : 0);
    return cc;
}
// callCalculateExps (prover_helpers.js:23-31): the parallel / threaded variants are scheduling choices of the CPU implementation
async function callCalculateExps(stage, code, dom, ctx, parallelExec, useThreads, debug) {
    return calculateExps(ctx, code, dom, debug, false);
}

module.exports = { compileCode, calculateExps, callCalculateExps };
