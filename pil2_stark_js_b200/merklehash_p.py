"""Mirror of src/helpers/hash/merklehash/merklehash_p.js -- buildMerkleHash(splitLinearHash) -> MerkleHash.

tree = {"elements": buff (the caller's array, aliased like the reference), "nodes": uint64 array in the reference layout,
"width", "height"}.  Hashing (leaves, levels, proof verification) runs on the GPU through the C ABI."""
import numpy as np

from ._lib import OutOfRange, E_RANGE
from .context import default_context

P = 0xFFFFFFFF00000001


def buildMerkleHash(splitLinearHash=False, ctx=None):
    """merklehash_p.js:12-16."""
    return MerkleHash(splitLinearHash, ctx)


class MerkleHash:
    def __init__(self, splitLinearHash=False, ctx=None):
        self.ctx = ctx or default_context()
        self.splitLinearHash = bool(splitLinearHash)

    def _getNNodes(self, n):
        """merklehash_p.js:28-42; n = height*4."""
        return self.ctx.merkle_nnodes(n // 4)

    def merkelize(self, buff, width, height):
        """merklehash_p.js:44-133."""
        nodes = self.ctx.merkelize(buff, width, height, self.splitLinearHash)
        return {"elements": buff, "nodes": nodes, "width": width, "height": height}

    def getElement(self, tree, idx, subIdx):
        """merklehash_p.js:136-139."""
        return int(tree["elements"][tree["width"] * idx + subIdx])

    def getGroupProof(self, tree, idx):
        """merklehash_p.js:142-168: [row values, [4-word sibling per level]]."""
        if idx < 0 or idx >= tree["height"]:
            raise OutOfRange(E_RANGE, "Out of range")
        w = tree["width"]
        v = [int(x) for x in tree["elements"][idx * w:(idx + 1) * w]]
        mp = []
        offset, n = 0, tree["height"] * 4
        nodes = tree["nodes"]
        while n > 4:
            si = (idx ^ 1) * 4
            mp.append([int(x) for x in nodes[offset + si:offset + si + 4]])
            next_n = ((n - 1) // 8 + 1) * 4
            idx >>= 1
            offset += next_n * 2
            n = next_n
        return [v, mp]

    def calculateRootFromGroupProof(self, mp, idx, vals):
        """merklehash_p.js:170-209 (linear hash + one Poseidon per level, on the GPU)."""
        flat = []
        for v in vals:
            flat.extend(v) if isinstance(v, (list, tuple)) else flat.append(v)
        value = [int(x) for x in self.ctx.linear_hash(np.array([x % P for x in flat], dtype=np.uint64), self.splitLinearHash)]
        for sib in mp:
            sib = [int(x) for x in sib]
            st = (value + sib if (idx & 1) == 0 else sib + value) + [0, 0, 0, 0]
            value = [int(x) for x in self.ctx.poseidon(np.array(st, dtype=np.uint64))[:4]]
            idx >>= 1
        return value

    def eqRoot(self, r1, r2):
        """merklehash_p.js:211-217."""
        return all(int(a) % P == int(b) % P for a, b in zip(r1, r2)) and len(r1) == len(r2) == 4

    def verifyGroupProof(self, root, mp, idx, groupElements):
        """merklehash_p.js:219-222."""
        return self.eqRoot(self.calculateRootFromGroupProof(mp, idx, groupElements), root)

    def root(self, tree):
        """merklehash_p.js:224-226."""
        return [int(x) for x in tree["nodes"][-4:]]

    def writeToFile(self, tree, fileName):
        """merklehash_p.js:228-247: u64 width, u64 height, elements, nodes (raw little-endian)."""
        with open(fileName, "wb") as f:
            np.array([tree["width"], tree["height"]], dtype="<u8").tofile(f)
            np.ascontiguousarray(tree["elements"], dtype="<u8").tofile(f)
            np.ascontiguousarray(tree["nodes"], dtype="<u8").tofile(f)

    def readFromFile(self, fileName):
        """merklehash_p.js:249-278."""
        with open(fileName, "rb") as f:
            width, height = (int(x) for x in np.fromfile(f, dtype="<u8", count=2))
            elements = np.fromfile(f, dtype="<u8", count=width * height).astype(np.uint64, copy=False)
            nodes = np.fromfile(f, dtype="<u8", count=self._getNNodes(height * 4)).astype(np.uint64, copy=False)
        return {"elements": elements, "nodes": nodes, "width": width, "height": height}
