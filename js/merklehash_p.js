// Drop-in for src/helpers/hash/merklehash/merklehash_p.js: `module.exports = async function buildMerkleHash(split)`.
// tree = { elements: buff (aliased, not copied), nodes: BigUint64Array, width, height } exactly as merklehash_p.js:46-51 -- or, for a
// tree kept in HBM (commitDevice / readFromFileToDevice), { device: handle, width, height }: getGroupProof / getElement / root read it
// in place and only the opened rows and siblings cross PCIe.
"use strict";
const fs = require("fs");
const { addon, context, inPages, zeroCopyPages } = require("./pil2gpu.js");

module.exports = async function buildMerkleHash(splitLinearHash = false) {
    return new MerkleHash(splitLinearHash);
};

const IO_CHUNK = 1 << 25;            // elements per file read / write (the reference's chunk: merklehash_p.js:239,264)

class MerkleHash {
    constructor(splitLinearHash) {
        this.splitLinearHash = !!splitLinearHash;
        this.useThreads = true;                         // fields kept for source compatibility (merklehash_p.js:20-26)
        const split = this.splitLinearHash ? 1 : 0;
        // `poseidon(inputs[8], capacity[4]) -> 4 words`, `lh.hash(vals) -> 4 words`, and the field description `F` (p only: the GPU
        // path needs no JS field arithmetic)
        this.poseidon = (inputs, capacity = [0n, 0n, 0n, 0n]) => {
            const st = new BigUint64Array(12);
            st.set(BigUint64Array.from(inputs.map(BigInt)), 0); st.set(BigUint64Array.from(capacity.map(BigInt)), 8);
            return Array.from(addon.poseidon(context(), st).subarray(0, 4));
        };
        this.lh = { splitLinearHash: this.splitLinearHash, hash: (vals) => Array.from(addon.linearHash(context(), BigUint64Array.from(vals.flat(Infinity).map(BigInt)), split)) };
        this.F = { p: 0xFFFFFFFF00000001n };
    }
    _getNNodes(n) { return Number(addon.merkleNNodes(n / 4)); }                                 // :28-42
    async merkelize(buff, width, height) {                                                      // :44-133
        const nodes = new BigUint64Array(this._getNNodes(height * 4));
        await addon.merkelizePaged(context(), inPages(buff), width, height, this.splitLinearHash ? 1 : 0, nodes);
        return { elements: buff, nodes, width, height };
    }
    // interpolate + merkelize with the result kept on the device (stark_gen_helpers.js:388-412 without the 32 GiB download)
    async commitDevice(buffSrc, nPols, nBits, nBitsExt) {
        const r = await addon.commit(context(), inPages(buffSrc), nPols, nBits, nBitsExt, this.splitLinearHash ? 1 : 0);
        return { device: r.tree, width: nPols, height: 2 ** nBitsExt, rootWords: r.root };
    }
    async toHost(tree, elementsBuff) {                                                          // device tree -> the reference's tree object
        const nodes = new BigUint64Array(this._getNNodes(tree.height * 4));
        const pages = zeroCopyPages(elementsBuff);
        if (!pages) throw new Error("pil2gpu: toHost needs a BigUint64Array or a page-backed BigBuffer");
        await addon.treeDownload(tree.device, pages, nodes);
        return { elements: elementsBuff, nodes, width: tree.width, height: tree.height };
    }
    free(tree) { if (tree.device) { addon.treeFree(tree.device); tree.device = null; } }
    getElement(tree, idx, subIdx) {                                                             // :136-139
        if (tree.device) return addon.treeGroupProofs(tree.device, BigUint64Array.of(BigInt(idx))).rows[subIdx];
        const e = tree.elements;
        return (e instanceof BigUint64Array) ? e[tree.width * idx + subIdx] : e.getElement(tree.width * idx + subIdx);
    }
    getGroupProof(tree, idx) {                                                                  // :142-168
        if ((idx < 0) || (idx >= tree.height)) throw new Error("Out of range");
        if (tree.device) {
            const r = addon.treeGroupProofs(tree.device, BigUint64Array.of(BigInt(idx)));      // throws Error("Out of range") itself too
            const mp = [];
            for (let d = 0; d * 4 < r.siblings.length; d++) mp.push(Array.from(r.siblings.subarray(4 * d, 4 * d + 4)));
            return [Array.from(r.rows), mp];
        }
        const v = new Array(tree.width);
        for (let i = 0; i < tree.width; i++) v[i] = this.getElement(tree, idx, i);
        const mp = [];
        let offset = 0, n = tree.height * 4;
        while (n > 4) {                      // same walk as the reference, without copying `nodes` at every level (:155)
            const si = (idx ^ 1) * 4;
            mp.push([tree.nodes[offset + si], tree.nodes[offset + si + 1], tree.nodes[offset + si + 2], tree.nodes[offset + si + 3]]);
            const nextN = (Math.floor((n - 1) / 8) + 1) * 4;
            idx >>= 1; offset += nextN * 2; n = nextN;
        }
        return [v, mp];
    }
    calculateRootFromGroupProof(mp, idx, vals) {                                                // :170-209
        const flat = BigUint64Array.from(vals.flat(Infinity));
        let value = addon.linearHash(context(), flat, this.splitLinearHash ? 1 : 0);            // BigUint64Array(4)
        for (const sib of mp) {
            const st = new BigUint64Array(12);
            if ((idx & 1) === 0) { st.set(value, 0); st.set(BigUint64Array.from(sib), 4); }
            else { st.set(BigUint64Array.from(sib), 0); st.set(value, 4); }
            value = addon.poseidon(context(), st).subarray(0, 4);
            idx >>= 1;
        }
        return Array.from(value);
    }
    eqRoot(r1, r2) { for (let k = 0; k < 4; k++) if (BigInt(r1[k]) !== BigInt(r2[k])) return false; return true; }   // :211-217
    verifyGroupProof(root, mp, idx, groupElements) {                                            // :219-222
        return this.eqRoot(this.calculateRootFromGroupProof(mp, idx, groupElements), root);
    }
    root(tree) {                                                                                // :224-226
        if (tree.device) return Array.from(addon.treeRoot(tree.device));
        return Array.from(tree.nodes.slice(-4));
    }
    async writeToFile(tree, fileName) {                                                         // :228-247 (raw LE u64, 2^25-element chunks)
        const fd = fs.openSync(fileName, "w");
        const put = (ta) => { for (let o = 0; o < ta.length; o += IO_CHUNK) { const c = ta.subarray(o, Math.min(ta.length, o + IO_CHUNK)); fs.writeSync(fd, Buffer.from(c.buffer, c.byteOffset, c.byteLength)); } };
        fs.writeSync(fd, Buffer.from(new BigUint64Array([BigInt(tree.width), BigInt(tree.height)]).buffer));
        const e = tree.elements;
        if (e instanceof BigUint64Array) put(e);
        else for (let o = 0; o < e.length; o += IO_CHUNK) put(e.slice(o, Math.min(e.length, o + IO_CHUNK)));    // public surface only (:243)
        put(tree.nodes);
        fs.closeSync(fd);
    }
    async readFromFile(fileName) {                                                              // :249-278
        const { BigBuffer } = require("pilcom");
        const fd = fs.openSync(fileName, "r");
        const hdr = new BigUint64Array(2);
        fs.readSync(fd, Buffer.from(hdr.buffer), 0, 16, 0);
        const width = Number(hdr[0]), height = Number(hdr[1]);
        const elements = new BigBuffer(width * height);
        let pos = 16;
        const chunk = new BigUint64Array(Math.min(IO_CHUNK, Math.max(1, width * height)));
        for (let o = 0; o < width * height; o += IO_CHUNK) {
            const n = Math.min(IO_CHUNK, width * height - o);
            fs.readSync(fd, Buffer.from(chunk.buffer, 0, n * 8), 0, n * 8, pos);
            elements.set(chunk.subarray(0, n), o);                                              // :272
            pos += n * 8;
        }
        const nodes = new BigUint64Array(this._getNNodes(height * 4));
        for (let o = 0; o < nodes.length; o += IO_CHUNK) {
            const c = nodes.subarray(o, Math.min(nodes.length, o + IO_CHUNK));
            fs.readSync(fd, Buffer.from(c.buffer, c.byteOffset, c.byteLength), 0, c.byteLength, pos);
            pos += c.byteLength;
        }
        fs.closeSync(fd);
        return { elements, nodes, width, height };
    }
    // readFromFile straight into a device tree (no re-hashing): the const tree of a setup, opened by proofQueries on the GPU
    async readFromFileToDevice(fileName) {
        const t = await this.readFromFile(fileName);
        const pages = zeroCopyPages(t.elements) || [t.elements.slice(0, t.elements.length)];
        return { device: addon.treeFromPages(context(), pages, t.width, t.height, t.nodes), width: t.width, height: t.height };
    }
}
