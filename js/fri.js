// Drop-in for src/stark/fri.js: `module.exports = class FRI` with fold / proofQueries / verify.
// fold (fri.js:22-81) runs as one C call (pil2gpu_fri_fold_paged: fold + transposed rows + layer tree); proofQueries (:83-105)
// is unchanged host logic over MH.getGroupProof (which also accepts device trees); verify (:107-174) is the reference's own code
// path and is re-exported from the reference file so that verification semantics cannot drift.
"use strict";
const { addon, context } = require("./pil2gpu.js");
const RefFRI = require(process.env.PIL2_REF_FRI || "../src/stark/fri.reference.js");   // the original fri.js, renamed at install time

function log2(n) { let b = 0; while ((2 ** b) < n) b++; return b; }
const PAGE = 1 << 27;                                                                  // words per typed array (1 GiB)
function alloc(words) { const p = []; for (let o = 0; o < words; o += PAGE) p.push(new BigUint64Array(Math.min(PAGE, words - o))); return p; }
const at = (pages, i) => pages[Math.floor(i / PAGE)][i % PAGE];

module.exports = class FRI extends RefFRI {
    async fold(step, pol, challenge) {
        const polBits = log2(pol.length);
        if (step === 0) { if (polBits !== this.inNBits) throw new Error("Invalid polynomial size"); }
        else if ((2 ** polBits) !== pol.length) throw new Error("Invalid polynomial size");
        const last = step === this.steps.length - 1;
        const curBits = step === 0 ? polBits : this.steps[step].nBits;          // step 0 is the identity fold (:48-49)
        const nextBits = last ? -1 : this.steps[step + 1].nBits;
        const flat = alloc(pol.length * 3);                                      // JS array of [a0,a1,a2] -> 3 words each
        for (let i = 0; i < pol.length; i++) {
            for (let k = 0; k < 3; k++) { const w = 3 * i + k; flat[Math.floor(w / PAGE)][w % PAGE] = pol[i][k]; }
        }
        const n2 = 2 ** curBits;
        const pol2w = alloc(n2 * 3);
        const rows = last ? null : alloc(n2 * 3);
        const nodes = last ? null : new BigUint64Array(Number(addon.merkleNNodes(2 ** nextBits)));
        await addon.friFoldPaged(context(), flat, polBits, curBits, nextBits, this.steps[0].nBits, BigUint64Array.from(challenge),
            this.MH.splitLinearHash ? 1 : 0, pol2w, rows, nodes);
        const pol2 = new Array(n2);
        for (let i = 0; i < n2; i++) pol2[i] = [at(pol2w, 3 * i), at(pol2w, 3 * i + 1), at(pol2w, 3 * i + 2)];
        if (last) return { pol: pol2, tree: undefined, proof: pol2 };
        const nGroups = 2 ** nextBits, groupSize = n2 / nGroups;
        const elements = rows.length === 1 ? rows[0] : { length: n2 * 3, buffers: rows, getElement: (i) => at(rows, i) };
        const tree = { elements, nodes, width: 3 * groupSize, height: nGroups };
        return { pol: pol2, tree, proof: { root: this.MH.root(tree) } };
    }
};
