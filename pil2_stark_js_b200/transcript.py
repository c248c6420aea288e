"""Mirror of src/helpers/transcript/transcript.js (Fiat-Shamir sponge; host logic, Poseidon on the GPU)."""
import numpy as np

from .context import default_context


class Transcript:
    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self.state = [0, 0, 0, 0]
        self.pending = []
        self.out = []

    def _absorb(self):
        st = np.array(self.pending + self.state, dtype=np.uint64)
        self.out = [int(x) for x in self.ctx.poseidon(st)]
        self.pending = []
        self.state = self.out[:4]

    def updateState(self):                     # transcript.js:39-46
        while len(self.pending) < 8:
            self.pending.append(0)
        self._absorb()

    def getState(self):                        # transcript.js:10-15
        if self.pending:
            self.updateState()
        return self.state

    def put(self, a):                          # transcript.js:29-37
        for x in (a if isinstance(a, (list, tuple, np.ndarray)) else [a]):
            if isinstance(x, (list, tuple, np.ndarray)):
                self.put(x)
            else:
                self._add1(int(x))

    def _add1(self, a):                        # transcript.js:48-56
        self.out = []
        self.pending.append(a)
        if len(self.pending) == 8:
            self._absorb()

    def getFields1(self):                      # transcript.js:21-27
        if not self.out:
            self.updateState()
        return self.out.pop(0)

    def getField(self):                        # transcript.js:17-19
        return [self.getFields1(), self.getFields1(), self.getFields1()]

    def getPermutations(self, n, nBits):       # transcript.js:59-84
        n_fields = (n * nBits - 1) // 63 + 1
        fields = [self.getFields1() for _ in range(n_fields)]
        res, cur_field, cur_bit = [], 0, 0
        for _ in range(n):
            a = 0
            for j in range(nBits):
                if (fields[cur_field] >> cur_bit) & 1:
                    a += 1 << j
                cur_bit += 1
                if cur_bit == 63:
                    cur_bit = 0
                    cur_field += 1
            res.append(a)
        return res
