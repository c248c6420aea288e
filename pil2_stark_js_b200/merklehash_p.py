"""Mirror of src/helpers/hash/merklehash/merklehash_p.js -- buildMerkleHash(splitLinearHash) -> MerkleHash.

tree = {"elements": buff (the caller's array, aliased like the reference), "nodes": uint64 array in the reference layout,
"width", "height"}.  Hashing (leaves, levels, proof verification) runs on the GPU through the C ABI."""
import numpy as np

from ._lib import OutOfRange, E_RANGE
from .context import default_context, DeviceTree

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
        if isinstance(tree, DeviceTree):
            return int(tree.group_proofs([idx])[0][0][subIdx])
        return int(tree["elements"][tree["width"] * idx + subIdx])

    def getGroupProof(self, tree, idx):
        """merklehash_p.js:142-168: [row values, [4-word sibling per level]].  `tree` is the reference's tree object (host
        arrays) or a DeviceTree (rows and siblings gathered on the GPU, only the proof crosses PCIe)."""
        if isinstance(tree, DeviceTree):
            rows, sib = tree.group_proofs([idx])            # raises OutOfRange like merklehash_p.js:143
            return [[int(x) for x in rows[0]], [[int(x) for x in s] for s in sib[0]]]
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
        if isinstance(tree, DeviceTree):
            return [int(x) for x in tree.root()]
        return [int(x) for x in tree["nodes"][-4:]]

    def writeToFile(self, tree, fileName):
        """merklehash_p.js:228-247: u64 width, u64 height, elements, nodes (raw little-endian)."""
        with open(fileName, "wb") as f:
            np.array([tree["width"], tree["height"]], dtype="<u8").tofile(f)
            np.ascontiguousarray(tree["elements"], dtype="<u8").tofile(f)
            np.ascontiguousarray(tree["nodes"], dtype="<u8").tofile(f)

    def readFromFileToDevice(self, fileName, chunk_words=1 << 25):
        """readFromFile (merklehash_p.js:249-278) straight into a device-resident tree: the file streams through one host
        chunk (the reference reads in 2^25-element chunks too, :262-276), nothing is re-hashed, and proofQueries over the
        result only move the opened rows and siblings back (getGroupProof / root accept the returned DeviceTree)."""
        with open(fileName, "rb") as f:
            width, height = (int(x) for x in np.fromfile(f, dtype="<u8", count=2))
            tree = self.ctx.tree_alloc(width, height)
            for which, total in ((0, width * height), (1, self._getNNodes(height * 4))):
                off = 0
                while off < total:
                    n = min(chunk_words, total - off)
                    chunk = np.fromfile(f, dtype="<u8", count=n)
                    if chunk.size != n:
                        tree.free()
                        raise ValueError("const tree file is truncated")
                    tree.fill(which, off, chunk)
                    off += n
        return tree

    def readFromFile(self, fileName):
        """merklehash_p.js:249-278."""
        with open(fileName, "rb") as f:
            width, height = (int(x) for x in np.fromfile(f, dtype="<u8", count=2))
            elements = np.fromfile(f, dtype="<u8", count=width * height).astype(np.uint64, copy=False)
            nodes = np.fromfile(f, dtype="<u8", count=self._getNNodes(height * 4)).astype(np.uint64, copy=False)
        return {"elements": elements, "nodes": nodes, "width": width, "height": height}
