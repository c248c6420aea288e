// GPU replacements for the arithmetic of four functions of src/stark/stark_gen_helpers.js.  Usage: require this file from
// the reference's stark_gen_helpers.js and delegate, e.g. `module.exports.computeQStark = require(".../js/stark_gen_helpers.js").computeQStark`.
// ctx is the reference's prover context (same field names).  extendAndMerkelize and computeQStark take BigBuffers of any size
// (pil2gpu_extend_and_merkelize_paged / pil2gpu_compute_q_paged); the evaluation / FRI-polynomial helpers read single-page buffers
// (at sizes where the extended buffers are multi-page, keep them on the device: MerkleHash.commitDevice).
"use strict";
const { addon, context, inPages, outPages, zeroCopyPages } = require("./pil2gpu.js");

function flat(buff) {
    const p = zeroCopyPages(buff);
    if (!p || p.length !== 1) throw new Error("pil2gpu: this entry point needs a single-page buffer");
    return p[0];
}
const split = (ctx) => (ctx.MH.splitLinearHash ? 1 : 0);
const xiOf = (ctx) => BigUint64Array.from(ctx.challenges[ctx.pilInfo.nStages + 1][0]);
const openingsOf = (ctx) => Int32Array.from(ctx.pilInfo.openingPoints.map(Number));

// stark_gen_helpers.js:388-412: interpolate (:397) + merkelize (:400) as ONE call -- the extended buffer crosses PCIe once
async function extendAndMerkelize(stage, ctx) {
    const nPols = ctx.pilInfo.mapSectionsN["cm" + stage] || 0;
    const dst = outPages(ctx["cm" + stage + "_ext"]);
    const nodes = new BigUint64Array(Number(addon.merkleNNodes(ctx.extN)));
    const root = await addon.extendAndMerkelizePaged(context(), inPages(ctx["cm" + stage + "_n"]), nPols, ctx.nBits, ctx.nBitsExt, split(ctx), dst.pages, nodes);
    dst.commit();
    ctx.trees[stage] = { elements: ctx["cm" + stage + "_ext"], nodes, width: nPols, height: ctx.extN };
    return [Array.from(root)];
}
// stark_gen_helpers.js:168-208
async function computeQStark(ctx) {
    const qStage = ctx.pilInfo.nStages + 1;
    const nPolsQ = ctx.pilInfo.mapSectionsN["cm" + qStage] || 0;
    const dst = outPages(ctx["cm" + qStage + "_ext"]);
    const nodes = new BigUint64Array(Number(addon.merkleNNodes(ctx.extN)));
    const root = await addon.computeQPaged(context(), inPages(ctx.q_ext), ctx.pilInfo.qDim, ctx.pilInfo.qDeg, ctx.nBits, ctx.nBitsExt, split(ctx), dst.pages, nodes);
    dst.commit();
    ctx.trees[qStage] = { elements: ctx["cm" + qStage + "_ext"], nodes, width: nPolsQ, height: ctx.extN };
    return [Array.from(root)];
}
// stark_gen_helpers.js:210-273 (hashCommits == false)
async function computeEvalsStark(ctx) {
    const openings = ctx.pilInfo.openingPoints.map(Number);
    const groups = new Map();          // buffer name -> { size, items: [[evIndex, offset, dim, lev]] }
    ctx.pilInfo.evMap.forEach((ev, i) => {
        let name, size, offset, dim;
        if (ev.type === "const") { name = "const_ext"; size = ctx.pilInfo.nConstants; offset = ev.id; dim = 1; }
        else if (ev.type === "cm") { const p = ctx.pilInfo.cmPolsMap[ev.id]; name = "cm" + p.stage + "_ext"; size = ctx.pilInfo.mapSectionsN["cm" + p.stage]; offset = p.stagePos; dim = p.dim; }
        else throw new Error("Invalid ev type: " + ev.type);
        if (!groups.has(name)) groups.set(name, { size, items: [] });
        groups.get(name).items.push([i, offset, dim, openings.indexOf(Number(ev.prime))]);
    });
    ctx.evals = new Array(ctx.pilInfo.evMap.length);
    for (const [name, g] of groups) {
        const d = new BigUint64Array(2 * g.items.length);
        g.items.forEach(([, offset, dim, lev], k) => { d[2 * k] = BigInt(offset); d[2 * k + 1] = BigInt(dim) | (BigInt(lev) << 32n); });
        const out = addon.computeEvals(context(), xiOf(ctx), openingsOf(ctx), ctx.nBits, ctx.nBitsExt, flat(ctx[name]), g.size, d);
        g.items.forEach(([i], k) => { ctx.evals[i] = [out[3 * k], out[3 * k + 1], out[3 * k + 2]]; });
    }
    return ctx.evals;
}
// the xDivXSubXi_ext loops of computeFRIStark, stark_gen_helpers.js:289-323
function computeXDivXSubXi(ctx) {
    addon.xDivXSubXi(context(), xiOf(ctx), openingsOf(ctx), ctx.nBits, ctx.nBitsExt, flat(ctx.xDivXSubXi_ext));
}
// computeFRIStark, stark_gen_helpers.js:275-334: the xDivXSubXi table and friExp over the extended domain in one call
// (replaces the two BigInt loops and callCalculateExps(stage, friExp code, "ext")); fills ctx.xDivXSubXi_ext, ctx.f_ext, ctx.friPol[0].
async function computeFRIStark(ctx) {
    const stage = ctx.pilInfo.nStages + 3;
    const names = [], bufs = [];
    const meta = new BigInt64Array(5 * ctx.pilInfo.evMap.length);
    ctx.pilInfo.evMap.forEach((ev, i) => {
        let name, size, offset, dim;
        if (ev.type === "const") { name = "const_ext"; size = ctx.pilInfo.nConstants; offset = ev.id; dim = 1; }
        else if (ev.type === "cm") { const p = ctx.pilInfo.cmPolsMap[ev.id]; name = "cm" + p.stage + "_ext"; size = ctx.pilInfo.mapSectionsN["cm" + p.stage]; offset = p.stagePos; dim = p.dim; }
        else throw new Error("Invalid ev type: " + ev.type);
        let bi = names.indexOf(name);
        if (bi < 0) { bi = names.length; names.push(name); bufs.push(flat(ctx[name])); }
        meta.set([BigInt(bi), BigInt(size), BigInt(offset), BigInt(dim), BigInt(Number(ev.prime))], 5 * i);
    });
    const evals = new BigUint64Array(3 * ctx.evals.length);
    ctx.evals.forEach((e, i) => { const v = Array.isArray(e) ? e : [e, 0n, 0n]; evals.set(v.map(BigInt), 3 * i); });
    const fExt = flat(ctx.f_ext);
    addon.friPol(context(), bufs, meta, evals, openingsOf(ctx), xiOf(ctx), BigUint64Array.from(ctx.challenges[stage][0]),
        BigUint64Array.from(ctx.challenges[stage][1]), ctx.nBits, ctx.nBitsExt, fExt, flat(ctx.xDivXSubXi_ext));
    ctx.friPol = []; ctx.friProof = [{}]; ctx.friTrees = [];
    ctx.friPol[0] = new Array(ctx.extN);
    for (let i = 0; i < ctx.extN; i++) ctx.friPol[0][i] = [fExt[3 * i], fExt[3 * i + 1], fExt[3 * i + 2]];   // :327-334
}
module.exports = { extendAndMerkelize, computeQStark, computeEvalsStark, computeXDivXSubXi, computeFRIStark };
