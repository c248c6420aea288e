"""Mirror of src/stark/fri.js -- class FRI(starkStruct, MH): fold / proofQueries / verify.

Polynomials are numpy uint64 arrays of shape (2^bits, 3) (the reference uses JS arrays of [a0,a1,a2] BigInts).  The fold,
the transposed rows and the layer tree are computed by one C call (pil2gpu_fri_fold) on the GPU."""
import numpy as np

from .context import default_context

P = 0xFFFFFFFF00000001
SHIFT = 7
W32 = 7277203076849721926


def _log2(n):
    return int(n).bit_length() - 1


class FRI:
    def __init__(self, starkStruct, MH):
        if not starkStruct:
            raise ValueError("stark struct not defined")
        self.inNBits = starkStruct["nBitsExt"]
        self.maxDegNBits = starkStruct["nBits"]
        self.nQueries = starkStruct["nQueries"]
        self.steps = starkStruct["steps"]
        self.MH = MH
        self.ctx = getattr(MH, "ctx", None) or default_context()

    def fold(self, step, pol, challenge):
        """fri.js:22-81.  Returns {"pol", "tree", "proof"} like the reference."""
        pol = np.ascontiguousarray(pol, dtype=np.uint64).reshape(-1, 3)
        polBits = _log2(pol.shape[0])
        if step == 0:
            assert polBits == self.inNBits, "Invalid polynomial size"
        else:
            assert 1 << polBits == pol.shape[0], "Invalid polynomial size"
        curBits = self.steps[step]["nBits"]
        last = step == len(self.steps) - 1
        nextBits = None if last else self.steps[step + 1]["nBits"]
        prevBits = polBits
        if step == 0:
            curBits = polBits                        # identity fold (fri.js:48-49)
        split = bool(getattr(self.MH, "splitLinearHash", False))
        pol2, rows, nodes = self.ctx.fri_fold(pol, prevBits, curBits, nextBits, self.steps[0]["nBits"],
                                              [int(c) % P for c in challenge], split)
        tree, proof = None, None
        if not last:
            nGroups = 1 << nextBits
            groupSize = (1 << curBits) // nGroups
            tree = {"elements": rows, "nodes": nodes, "width": 3 * groupSize, "height": nGroups}
            proof = {"root": self.MH.root(tree)}
        else:
            proof = [[int(x) for x in row] for row in pol2]
        return {"pol": pol2, "tree": tree, "proof": proof}

    def proofQueries(self, proof, trees, friQueries):
        """fri.js:83-105 (mutates friQueries like the reference)."""
        for step in range(len(self.steps)):
            proof[step]["polQueries"] = []
            if step == 0:
                for q in friQueries:
                    proof[step]["polQueries"].append([self.MH.getGroupProof(t, q) for t in trees[step]])
            else:
                for i in range(len(friQueries)):
                    friQueries[i] = friQueries[i] % (1 << self.steps[step]["nBits"])
                for q in friQueries:
                    proof[step]["polQueries"].append(self.MH.getGroupProof(trees[step], q))

    def verify(self, friChallenges, friQueries, proof, checkQuery):
        """fri.js:107-174.  Host-side verification arithmetic (tiny); Merkle checks go through MH (GPU hashing)."""
        assert len(proof) == len(self.steps) + 1, "Invalid proof size"
        friQueries = list(friQueries)
        polBits = self.inNBits
        shift = SHIFT
        for si in range(len(self.steps)):
            proofItem = proof[si]
            reductionBits = polBits - self.steps[si]["nBits"]
            for i in range(self.nQueries):
                pgroup_e = checkQuery(proofItem["polQueries"][i], friQueries[i])
                if not pgroup_e:
                    return False
                w = pow(W32, 1 << (32 - polBits), P)
                sinv = pow((shift * pow(w, friQueries[i], P)) % P, P - 2, P)
                ev = _eval_group(pgroup_e, [(c * sinv) % P for c in friChallenges[si]])
                if si < len(self.steps) - 1:
                    nextNGroups = 1 << self.steps[si + 1]["nBits"]
                    groupIdx = friQueries[i] // nextNGroups
                    query = proof[si + 1]["polQueries"][i][0]
                    if [int(x) % P for x in query[groupIdx * 3:groupIdx * 3 + 3]] != ev:
                        return False
                else:
                    if [int(x) % P for x in proof[si + 1][friQueries[i]]] != ev:
                        return False
            root_next = proof[si + 1]["root"] if si < len(self.steps) - 1 else None

            def checkQuery(query, idx, _root=root_next):
                if not self.MH.verifyGroupProof(_root, query[1], idx, query[0]):
                    return False
                return [[int(x) for x in query[0][k:k + 3]] for k in range(0, len(query[0]), 3)]

            polBits = self.steps[si]["nBits"]
            for _ in range(reductionBits):
                shift = (shift * shift) % P
            if si < len(self.steps) - 1:
                friQueries = [q % (1 << self.steps[si + 1]["nBits"]) for q in friQueries]
        lastPol_e = [[int(x) % P for x in e] for e in proof[-1]]
        if polBits - (self.inNBits - self.maxDegNBits) < 0:
            maxDeg = 0
        else:
            maxDeg = 1 << (polBits - (self.inNBits - self.maxDegNBits))
        lastPol_c = _intt3(lastPol_e)
        return all(c == [0, 0, 0] for c in lastPol_c[maxDeg + 1:])


# ---- tiny host-side F3 helpers used only by verify() (O(nQueries * groupSize) work) ----
def _f3mul(a, b):
    A = (a[0] + a[1]) * (b[0] + b[1]); B = (a[0] + a[2]) * (b[0] + b[2]); C = (a[1] + a[2]) * (b[1] + b[2])
    D = a[0] * b[0]; E = a[1] * b[1]; F = a[2] * b[2]; G = D - E
    return [(C + G - F) % P, (A + C - E - E - D) % P, (B - G) % P]


def _intt3(e):
    n = len(e)
    if n <= 1:
        return [list(x) for x in e]
    bits = _log2(n)
    w = pow(W32, 1 << (32 - bits), P)
    winv = pow(w, P - 2, P)
    ninv = pow(n, P - 2, P)
    return [[sum(e[j][k] * pow(winv, i * j, P) for j in range(n)) * ninv % P for k in range(3)] for i in range(n)]


def _eval_group(pgroup_e, x):
    c = _intt3([[int(v) % P for v in e] for e in pgroup_e])
    res = c[-1]
    for i in range(len(c) - 2, -1, -1):
        m = _f3mul(res, x)
        res = [(m[k] + c[i][k]) % P for k in range(3)]
    return res
