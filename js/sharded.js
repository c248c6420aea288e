// Multi-GPU extendAndMerkelize / getGroupProof for the reference's prover (stark_gen_helpers.js:388-412, merklehash_p.js:142-168) over the
// commit group of the C ABI (include/pil2gpu.h, pil2gpu_shard_*).  Rank g owns columns [g*nPols/G, (g+1)*nPols/G) of the trace; the
// column -> row exchange rides on the last LDE pass as peer stores, the ranks order themselves with flag barriers on their streams, and
// sub-roots / opened rows reach every rank through peer-mapped mailboxes -- there is no collective library and nothing for the host to
// relay after the 128-byte handle pairs.  Two shapes:
//
//   LocalCommitGroup   every GPU driven from this process.  addon.shardCommit / shardOpen only enqueue (they never wait for a peer), so
//                      one thread enqueues all ranks and then awaits the asynchronous reads; nothing can deadlock in the libuv pool.
//   forkCommitGroup    one worker process per GPU (child_process.fork): each worker builds / loads its own column slab (a 16 GiB trace
//                      does not travel over a pipe), the parent only gathers and redistributes the handle pairs and fans commands out.
//
// Both return the root of the WHOLE tree (identical to MerkleHash.merkelize on the full extended trace) and whole-tree proofs.
"use strict";
const { addon } = require("./pil2gpu.js");

function stageWordsFor(nPols, nBitsExt, maxQueries) {
    return BigInt(maxQueries) * (BigInt(nPols) + 4n * BigInt(nBitsExt));
}

class LocalCommitGroup {
    // devices: CUDA device ordinals, a power of two of them; shapes bound the buffers: { nPols, nBitsExt, maxQueries }
    constructor(devices, { nPols, nBitsExt, maxQueries = 128 }) {
        this.world = devices.length;
        if (nPols % this.world) throw new Error(`nPols (${nPols}) must be a multiple of the number of GPUs (${this.world})`);
        this.ctxs = devices.map((d) => addon.create(d));
        const recvWords = (BigInt(nPols / this.world)) << BigInt(nBitsExt);
        this.shards = this.ctxs.map((c, r) => addon.shardCreate(c, r, this.world, recvWords, stageWordsFor(nPols, nBitsExt, maxQueries)));
        addon.shardConnectLocal(this.shards);
    }
    // slabs[g]: Array of BigUint64Array pages holding rank g's column slab (2^nBits rows x nPols/G columns, row-major; pinned pages
    // from addon.allocPinnedPage upload by DMA).  Resolves with the 4-word root.
    async extendAndMerkelize(slabs, nPols, nBits, nBitsExt, split = false) {
        if (slabs.length !== this.world) throw new Error("one column slab per GPU");
        for (let g = 0; g < this.world; g++) addon.shardCommit(this.shards[g], slabs[g], nPols, nBits, nBitsExt, split ? 1 : 0);
        const roots = await Promise.all(this.shards.map((s) => addon.shardRoot(s)));
        return roots[0];
    }
    // idxs: BigUint64Array of leaf indices of the extended trace.  Resolves with { rows, siblings } exactly as
    // MerkleHash.getGroupProof would give them for the whole tree, concatenated over the queries.
    async getGroupProofs(idxs) {
        for (const s of this.shards) addon.shardOpen(s, idxs);
        const all = await Promise.all(this.shards.map((s) => addon.shardProofs(s)));
        return all[0];
    }
    free() {
        for (const s of this.shards) addon.shardFree(s);
        this.shards = [];
    }
}

// ---- one worker process per GPU ------------------------------------------------------------------------------------------------------
// parent:  const group = await forkCommitGroup(8, require.resolve("./my_slab_loader.js"), { nPols, nBitsExt });
//          const root = await group.extendAndMerkelize({ nPols, nBits, nBitsExt, split: false, what: "cm1" });
//          const proofs = await group.getGroupProofs(idxs);   group.close();
// loader:  module.exports = async function loadSlab(rank, world, what) -> Array of BigUint64Array pages (this rank's column slab)
async function forkCommitGroup(world, loaderPath, shapes) {
    const { fork } = require("child_process");
    const workers = [];
    for (let r = 0; r < world; r++)
        workers.push(fork(__filename, ["--pil2gpu-worker", String(r), String(world), loaderPath, JSON.stringify(shapes)], { env: { ...process.env, PIL2GPU_DEVICE: String(r) } }));
    const once = (w, type) => new Promise((res, rej) => {
        const h = (m) => { if (m.type === type) { w.off("message", h); res(m); } else if (m.type === "error") { w.off("message", h); rej(new Error(m.message)); } };
        w.on("message", h);
    });
    const handles = await Promise.all(workers.map((w) => once(w, "handles")));
    const all = handles.map((m) => m.words);                                   // world x 16 decimal strings (BigInt does not serialise)
    await Promise.all(workers.map((w) => { const p = once(w, "connected"); w.send({ type: "connect", all }); return p; }));
    const broadcast = async (msg, reply) => {
        const ps = workers.map((w) => once(w, reply));
        for (const w of workers) w.send(msg);
        return (await Promise.all(ps))[0];
    };
    return {
        extendAndMerkelize: async (args) => BigUint64Array.from((await broadcast({ type: "commit", args }, "root")).words.map(BigInt)),
        getGroupProofs: async (idxs) => {
            const m = await broadcast({ type: "open", idxs: Array.from(idxs, String) }, "proofs");
            return { rows: BigUint64Array.from(m.rows.map(BigInt)), siblings: BigUint64Array.from(m.siblings.map(BigInt)) };
        },
        close: () => { for (const w of workers) w.send({ type: "exit" }); },
    };
}

async function workerMain(rank, world, loaderPath, shapes) {
    const loadSlab = require(loaderPath);
    const ctx = addon.create(rank);
    const recvWords = BigInt(shapes.nPols / world) << BigInt(shapes.nBitsExt);
    const shard = addon.shardCreate(ctx, rank, world, recvWords, stageWordsFor(shapes.nPols, shapes.nBitsExt, shapes.maxQueries || 128));
    process.send({ type: "handles", words: Array.from(addon.shardHandles(shard), String) });
    process.on("message", async (m) => {
        try {
            if (m.type === "connect") {
                addon.shardConnect(shard, BigUint64Array.from(m.all.flat().map(BigInt)));
                process.send({ type: "connected" });
            } else if (m.type === "commit") {
                const a = m.args;
                addon.shardCommit(shard, await loadSlab(rank, world, a.what), a.nPols, a.nBits, a.nBitsExt, a.split ? 1 : 0);
                process.send({ type: "root", words: Array.from(await addon.shardRoot(shard), String) });
            } else if (m.type === "open") {
                addon.shardOpen(shard, BigUint64Array.from(m.idxs.map(BigInt)));
                const p = await addon.shardProofs(shard);
                process.send(rank === 0 ? { type: "proofs", rows: Array.from(p.rows, String), siblings: Array.from(p.siblings, String) } : { type: "proofs" });
            } else if (m.type === "exit") {
                addon.shardFree(shard);
                process.exit(0);
            }
        } catch (e) {
            process.send({ type: "error", message: String(e && e.message || e) });
        }
    });
}

if (require.main === module && process.argv[2] === "--pil2gpu-worker")
    workerMain(Number(process.argv[3]), Number(process.argv[4]), process.argv[5], JSON.parse(process.argv[6]));

module.exports = { LocalCommitGroup, forkCommitGroup };
