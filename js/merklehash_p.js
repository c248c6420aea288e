// Drop-in for src/helpers/hash/merklehash/merklehash_p.js: `module.exports = async function buildMerkleHash(split)`.
// tree = { elements: buff (aliased, not copied), nodes: BigUint64Array, width, height } exactly as merklehash_p.js:46-51.
"use strict";
const fs = require("fs");
const { addon, context, pagesOf } = require("./pil2gpu.js");

module.exports = async function buildMerkleHash(splitLinearHash = false) {
    return new MerkleHash(splitLinearHash);
};

class MerkleHash {
    constructor(splitLinearHash) {
        this.splitLinearHash = !!splitLinearHash;
        this.useThreads = true;                         // field kept for source compatibility (merklehash_p.js:24)
    }
    _getNNodes(n) { return Number(addon.merkleNNodes(BigInt(n / 4))); }                       // :28-42
    async merkelize(buff, width, height) {                                                      // :44-133
        const nodes = new BigUint64Array(this._getNNodes(height * 4));
        addon.merkelizePaged(context(), pagesOf(buff), width, height, this.splitLinearHash ? 1 : 0, nodes);
        return { elements: buff, nodes, width, height };
    }
    getElement(tree, idx, subIdx) {                                                             // :136-139
        const e = tree.elements;
        return (e instanceof BigUint64Array) ? e[tree.width * idx + subIdx] : e.getElement(tree.width * idx + subIdx);
    }
    getGroupProof(tree, idx) {                                                                  // :142-168
        if ((idx < 0) || (idx >= tree.height)) throw new Error("Out of range");
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
    root(tree) { return Array.from(tree.nodes.slice(-4)); }                                     // :224-226
    async writeToFile(tree, fileName) {                                                         // :228-247 (raw LE u64)
        const fd = fs.openSync(fileName, "w");
        fs.writeSync(fd, Buffer.from(new BigUint64Array([BigInt(tree.width), BigInt(tree.height)]).buffer));
        for (const page of pagesOf(tree.elements)) fs.writeSync(fd, Buffer.from(page.buffer, page.byteOffset, page.byteLength));
        fs.writeSync(fd, Buffer.from(tree.nodes.buffer, tree.nodes.byteOffset, tree.nodes.byteLength));
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
        for (const page of pagesOf(elements)) { fs.readSync(fd, Buffer.from(page.buffer, page.byteOffset, page.byteLength), 0, page.byteLength, pos); pos += page.byteLength; }
        const nodes = new BigUint64Array(this._getNNodes(height * 4));
        fs.readSync(fd, Buffer.from(nodes.buffer), 0, nodes.byteLength, pos);
        fs.closeSync(fd);
        return { elements, nodes, width, height };
    }
}
