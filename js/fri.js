// Drop-in for src/stark/fri.js: `module.exports = class FRI` with fold / proofQueries / verify.
// fold (fri.js:22-81) runs as one C call (pil2gpu_fri_fold: fold + transposed rows + layer tree); proofQueries (:83-105)
// is unchanged host logic over MH.getGroupProof; verify (:107-174) is the reference's own code path and is re-exported
// from the reference file so that verification semantics cannot drift.
"use strict";
const { addon, context } = require("./pil2gpu.js");
const RefFRI = require(process.env.PIL2_REF_FRI || "../src/stark/fri.reference.js");   // the original fri.js, renamed at install time

function log2(n) { let b = 0; while ((1 << b) < n) b++; return b; }

module.exports = class FRI extends RefFRI {
    async fold(step, pol, challenge) {
        const polBits = log2(pol.length);
        if (step === 0) { if (polBits !== this.inNBits) throw new Error("Invalid polynomial size"); }
        else if ((1 << polBits) !== pol.length) throw new Error("Invalid polynomial size");
        const last = step === this.steps.length - 1;
        const curBits = step === 0 ? polBits : this.steps[step].nBits;          // step 0 is the identity fold (:48-49)
        const nextBits = last ? -1 : this.steps[step + 1].nBits;
        const flat = new BigUint64Array(pol.length * 3);                         // JS array of [a0,a1,a2] -> 3 words each
        for (let i = 0; i < pol.length; i++) { flat[3 * i] = pol[i][0]; flat[3 * i + 1] = pol[i][1]; flat[3 * i + 2] = pol[i][2]; }
        const n2 = 1 << curBits;
        const pol2w = new BigUint64Array(n2 * 3);
        const rows = last ? null : new BigUint64Array(n2 * 3);
        const nodes = last ? null : new BigUint64Array(Number(addon.merkleNNodes(BigInt(1 << nextBits))));
        addon.friFold(context(), flat, polBits, curBits, nextBits, this.steps[0].nBits, BigUint64Array.from(challenge),
            this.MH.splitLinearHash ? 1 : 0, pol2w, rows, nodes);
        const pol2 = new Array(n2);
        for (let i = 0; i < n2; i++) pol2[i] = [pol2w[3 * i], pol2w[3 * i + 1], pol2w[3 * i + 2]];
        if (last) return { pol: pol2, tree: undefined, proof: pol2 };
        const nGroups = 1 << nextBits, groupSize = n2 / nGroups;
        const tree = { elements: rows, nodes, width: 3 * groupSize, height: nGroups };
        return { pol: pol2, tree, proof: { root: this.MH.root(tree) } };
    }
};
