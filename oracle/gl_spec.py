"""CPU restatement (pure-Python integers) of the pil2-stark-js commit-phase hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; it is the checker used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Parity status: PINNED -- every function
here is checked in tests/test_oracle_*.py against the reference's own known-answer vectors (3 Poseidon
KATs, F3 KAT) and against the committed sm_all proof fixture (root1, rootC, every Merkle path, both FRI
fold links, the transcript-derived query indices).

Each function cites the reference file:line (paths relative to the reference repo root) it restates.
Everything is exact integer arithmetic mod p = 2^64 - 2^32 + 1; there is no floating point.
"""
from .poseidon_rc import RC

P = 0xFFFFFFFF00000001            # src/helpers/f3g.js:18
SHIFT = 7                         # src/helpers/f3g.js:22  (coset generator)
W32 = 7277203076849721926         # src/helpers/f3g.js:40  (root of unity of order 2^32)
K_CONN = 12275445934081160404     # src/helpers/f3g.js:26


# ----------------------------------------------------------------------------------------------
# Field and cubic extension                                   src/helpers/f3g.js:47-104,136-188
# ----------------------------------------------------------------------------------------------
def inv(a):
    """Base-field inverse (f3g.js:174 _inv1); raises on zero like the reference."""
    a %= P
    if a == 0:
        raise ZeroDivisionError("Division by zero")
    return pow(a, P - 2, P)


def root_of_unity(s):
    """w[s]: generator of the order-2^s subgroup, w[s] = w[s+1]^2 (src/helpers/fft/fft.js:45-50)."""
    assert 0 <= s <= 32
    return pow(W32, 1 << (32 - s), P)


SHIFT_INV = inv(SHIFT)            # f3g.js:23


def f3_add(a, b):
    return [(a[0] + b[0]) % P, (a[1] + b[1]) % P, (a[2] + b[2]) % P]


def f3_sub(a, b):
    return [(a[0] - b[0]) % P, (a[1] - b[1]) % P, (a[2] - b[2]) % P]


def f3_mul(a, b):
    """F_p[x]/(x^3 - x - 1) product (f3g.js:94-102)."""
    A = (a[0] + a[1]) * (b[0] + b[1])
    B = (a[0] + a[2]) * (b[0] + b[2])
    C = (a[1] + a[2]) * (b[1] + b[2])
    D = a[0] * b[0]
    E = a[1] * b[1]
    F = a[2] * b[2]
    G = D - E
    return [(C + G - F) % P, (A + C - E - E - D) % P, (B - G) % P]


def f3_mul_scalar(a, s):
    return [(a[0] * s) % P, (a[1] * s) % P, (a[2] * s) % P]


def f3_inv(a):
    """f3g.js:136-172."""
    aa = a[0] * a[0]; ac = a[0] * a[2]; ba = a[1] * a[0]; bb = a[1] * a[1]; bc = a[1] * a[2]; cc = a[2] * a[2]
    aaa = aa * a[0]; aac = aa * a[2]; abc = ba * a[2]; abb = ba * a[1]; acc = ac * a[2]
    bbb = bb * a[1]; bcc = bc * a[2]; ccc = cc * a[2]
    t = (-aaa - aac - aac + abc + abc + abc + abb - acc - bbb + bcc - ccc) % P
    tinv = inv(t)
    return [((-aa - ac - ac + bc + bb - cc) * tinv) % P, ((ba - cc) * tinv) % P, ((-bb + ac + cc) * tinv) % P]


# ----------------------------------------------------------------------------------------------
# Serial NTT, the semantics every transform is tested against   src/helpers/fft/fft.js:118-174
# ----------------------------------------------------------------------------------------------
def _bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def ntt(p):
    """Natural-order radix-2 DIT (fft.js:118-163).  Elements are ints or F3 lists (component-wise)."""
    n = len(p)
    if n <= 1:
        return list(p)
    bits = n.bit_length() - 1
    assert 1 << bits == n, "Size must be multiple of 2"
    ext = isinstance(p[0], (list, tuple))
    buff = [None] * n
    for i in range(n):
        buff[_bitrev(i, bits)] = p[i]
    for s in range(1, bits + 1):
        m = 1 << s
        h = m >> 1
        winc = root_of_unity(s)
        for k in range(0, n, m):
            w = 1
            for j in range(h):
                if ext:
                    t = f3_mul_scalar(buff[k + j + h], w)
                    u = buff[k + j]
                    buff[k + j] = f3_add(u, t)
                    buff[k + j + h] = f3_sub(u, t)
                else:
                    t = (w * buff[k + j + h]) % P
                    u = buff[k + j]
                    buff[k + j] = (u + t) % P
                    buff[k + j + h] = (u - t) % P
                w = (w * winc) % P
    return buff


def intt(p):
    """fft.js:165-174: forward transform, index reversal (n-i)%n, scale by 1/n."""
    n = len(p)
    if n <= 1:
        return list(p)
    q = ntt(p)
    ninv = inv(n)
    ext = isinstance(p[0], (list, tuple))
    res = [None] * n
    for i in range(n):
        res[(n - i) % n] = f3_mul_scalar(q[i], ninv) if ext else (q[i] * ninv) % P
    return res


def pol_mul_axi(p, init, acc):
    """src/helpers/polutils.js:1-7: p[i] *= init*acc^i (in place)."""
    r = init
    ext = len(p) > 0 and isinstance(p[0], (list, tuple))
    for i in range(len(p)):
        p[i] = f3_mul_scalar(p[i], r) if ext else (p[i] * r) % P
        r = (r * acc) % P


def eval_pol(p, x):
    """src/helpers/polutils.js:9-16: Horner; p is a list of F3, x in F3."""
    if len(p) == 0:
        return [0, 0, 0]
    res = p[-1]
    for i in range(len(p) - 2, -1, -1):
        res = f3_add(f3_mul(res, x), p[i])
    return res


def extend_pol(p, extend_bits=1):
    """src/helpers/polutils.js:18-30 (shift=true): single-column LDE onto the coset 7*<w_ext>."""
    res = intt([x % P for x in p])
    pol_mul_axi(res, 1, SHIFT)
    res = res + [0] * ((len(p) << extend_bits) - len(p))
    return ntt(res)


# -- multi-column buffer entry points (row-major buff[row*nPols + col]) --------------------------
def fft_p(src, n_pols, n_bits, inverse=False):
    """Per-column transform of a row-major buffer == fft_p.js:178-184 (tested equal to F.fft/F.ifft per
    column by test/fft_p.test.js:47-190)."""
    n = 1 << n_bits
    dst = [0] * (n * n_pols)
    for c in range(n_pols):
        col = [src[r * n_pols + c] for r in range(n)]
        out = intt(col) if inverse else ntt(col)
        for r in range(n):
            dst[r * n_pols + c] = out[r]
    return dst


def interpolate(src, n_pols, n_bits, n_bits_ext):
    """fft_p.js:187-297 net effect: dst[j*C+c] = P_c(7*w_ext^j) (== extendPol per column,
    test/fft_p.test.js:82,193)."""
    n = 1 << n_bits
    ne = 1 << n_bits_ext
    dst = [0] * (ne * n_pols)
    for c in range(n_pols):
        col = [src[r * n_pols + c] for r in range(n)]
        out = extend_pol(col, n_bits_ext - n_bits)
        for r in range(ne):
            dst[r * n_pols + c] = out[r]
    return dst


# ----------------------------------------------------------------------------------------------
# Poseidon-GL, plain 30-round form                    src/helpers/glwasm.js:359-390,428-440
# ----------------------------------------------------------------------------------------------
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]       # glwasm.js:430
MDS_DIAG = [8] + [0] * 11                                       # glwasm.js:431
MDS = [[MDS_CIRC[(j - i) % 12] + (MDS_DIAG[i] if i == j else 0) for j in range(12)] for i in range(12)]


def poseidon_perm(state):
    """Full width-12 permutation, all 12 outputs canonical (glwasm.js:359-390; equals the optimised
    form of src/helpers/hash/poseidon/poseidon.js:57-108)."""
    x = [v % P for v in state]
    assert len(x) == 12
    for r in range(30):
        x = [(x[i] + RC[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            x = [pow(v, 7, P) for v in x]
        else:
            x[0] = pow(x[0], 7, P)
        x = [sum(MDS[i][j] * x[j] for j in range(12)) % P for i in range(12)]
    return x


def poseidon(inputs, capacity=None, n_outs=4):
    """poseidon.js:57-108 signature: 8 inputs + optional 4 capacity -> first n_outs words."""
    if len(inputs) != 8:
        raise ValueError("Invalid Input size (must be 8)")
    if capacity is None:
        capacity = [0, 0, 0, 0]
    if len(capacity) != 4:
        raise ValueError("Invalid Capacity size (must be 4)")
    return poseidon_perm(list(inputs) + list(capacity))[:n_outs]


# ----------------------------------------------------------------------------------------------
# Linear hash                  src/helpers/hash/linearhash/linearhash.js:8-42, linearhash_gpu.js:8-67
# ----------------------------------------------------------------------------------------------
def _sponge(vals):
    """linearhash.js:19-41 / linearhash_gpu.js:8-29."""
    st = [0, 0, 0, 0]
    if len(vals) <= 4:
        return [int(v) for v in vals] + [0] * (4 - len(vals))
    for i in range(0, len(vals), 8):
        chunk = list(vals[i:i + 8])
        chunk += [0] * (8 - len(chunk))
        st = poseidon(chunk, st)
    return st


def linear_hash(vals, split=False):
    """Row -> 4-word digest.  split=False: linearhash.js:8-42 (+ the width<=4 passthrough done in JS at
    merklehash_worker.js:42-49).  split=True: linearhash_gpu.js:31-67."""
    vals = [int(v) for v in vals]
    if not split:
        return _sponge(vals)
    if len(vals) <= 4:
        return vals + [0] * (4 - len(vals))
    batch = max(8, (len(vals) + 3) // 4)           # linearhash_gpu.js:42-44
    hashes = []
    for b in range(0, len(vals), batch):
        hashes += _sponge(vals[b:b + batch])
    if len(hashes) <= 4:
        return hashes + [0] * (4 - len(hashes))
    return _sponge(hashes)


# ----------------------------------------------------------------------------------------------
# Merkle tree            src/helpers/hash/merklehash/merklehash_p.js:28-42,44-133,142-226
# ----------------------------------------------------------------------------------------------
def merkle_n_nodes(n64):
    """_getNNodes (merklehash_p.js:28-42); argument is height*4 (words in the leaf level)."""
    next_n = ((n64 - 1) // 8 + 1) * 4
    acc = next_n * 2
    n = n64
    while n > 4:
        n = next_n
        next_n = ((n - 1) // 8 + 1) * 4
        if n > 4:
            acc += next_n * 2
        else:
            acc += 4
    return acc


def merkelize(buff, width, height, split=False):
    """merklehash_p.js:44-133.  Returns the tree dict {elements, nodes, width, height}; `nodes` is the flat
    word list in the reference layout (levels padded to an even node count, root = last 4 words)."""
    nodes = [0] * merkle_n_nodes(height * 4)
    for r in range(height):
        nodes[4 * r:4 * r + 4] = linear_hash(buff[r * width:(r + 1) * width], split)
    p_in = 0
    n64 = height * 4
    next_n64 = ((n64 - 1) // 8 + 1) * 4
    p_out = p_in + next_n64 * 2
    while n64 > 4:
        for i in range(next_n64 // 4):
            nodes[p_out + 4 * i:p_out + 4 * i + 4] = poseidon(nodes[p_in + 8 * i:p_in + 8 * i + 8])
        n64 = next_n64
        next_n64 = ((n64 - 1) // 8 + 1) * 4
        p_in = p_out
        p_out = p_in + next_n64 * 2
    return {"elements": buff, "nodes": nodes, "width": width, "height": height}


def merkle_root(tree):
    """merklehash_p.js:224."""
    return list(tree["nodes"][-4:])


def get_group_proof(tree, idx):
    """merklehash_p.js:142-168: ([row values], [[4-word sibling] per level])."""
    if idx < 0 or idx >= tree["height"]:
        raise IndexError("Out of range")
    w = tree["width"]
    v = list(tree["elements"][idx * w:(idx + 1) * w])
    mp = []
    offset, n = 0, tree["height"] * 4
    while n > 4:
        si = (idx ^ 1) * 4
        mp.append(list(tree["nodes"][offset + si:offset + si + 4]))
        next_n = ((n - 1) // 8 + 1) * 4
        idx >>= 1
        offset += next_n * 2
        n = next_n
    return v, mp


def root_from_group_proof(mp, idx, vals, split=False):
    """merklehash_p.js:170-209."""
    value = linear_hash(vals, split)
    for sib in mp:
        value = poseidon(value + list(sib)) if (idx & 1) == 0 else poseidon(list(sib) + value)
        idx >>= 1
    return value


def verify_group_proof(root, mp, idx, vals, split=False):
    """merklehash_p.js:219-222 (eqRoot: canonical word equality)."""
    return [int(x) % P for x in root_from_group_proof(mp, idx, vals, split)] == [int(x) % P for x in root]


# ----------------------------------------------------------------------------------------------
# Transcript                                    src/helpers/transcript/transcript.js:2-86
# ----------------------------------------------------------------------------------------------
class Transcript:
    def __init__(self, hash_fn=None):
        self.H = hash_fn or poseidon
        self.state = [0, 0, 0, 0]
        self.pending = []
        self.out = []

    def _update(self):                                  # transcript.js:39-46
        while len(self.pending) < 8:
            self.pending.append(0)
        self.out = list(self.H(self.pending, self.state, 12))
        self.pending = []
        self.state = self.out[:4]

    def put(self, a):                                   # transcript.js:29-37,48-56
        items = a if isinstance(a, (list, tuple)) else [a]
        for x in items:
            if isinstance(x, (list, tuple)):
                self.put(list(x))
                continue
            self.out = []
            self.pending.append(int(x))
            if len(self.pending) == 8:
                self.out = list(self.H(self.pending, self.state, 12))
                self.pending = []
                self.state = self.out[:4]

    def get_field1(self):                               # transcript.js:21-27
        if len(self.out) == 0:
            self._update()
        return self.out.pop(0)

    def get_field(self):                                # transcript.js:17-19
        return [self.get_field1(), self.get_field1(), self.get_field1()]

    def get_permutations(self, n, n_bits):              # transcript.js:59-84
        total = n * n_bits
        n_fields = (total - 1) // 63 + 1
        fields = [self.get_field1() for _ in range(n_fields)]
        res, cur_field, cur_bit = [], 0, 0
        for _ in range(n):
            a = 0
            for j in range(n_bits):
                if (fields[cur_field] >> cur_bit) & 1:
                    a += 1 << j
                cur_bit += 1
                if cur_bit == 63:
                    cur_bit = 0
                    cur_field += 1
            res.append(a)
        return res


# ----------------------------------------------------------------------------------------------
# FRI                                                        src/stark/fri.js:22-105,107-174,187-202
# ----------------------------------------------------------------------------------------------
def transposed_buffer(pol, transpose_bits):
    """fri.js:187-202: pol (list of F3) -> flat rows: row i (< 2^bits) = [pol[i + 2^bits*j] for j]."""
    n = len(pol)
    w = 1 << transpose_bits
    h = n // w
    res = [0] * (n * 3)
    for i in range(w):
        for j in range(h):
            fi = j * w + i
            di = i * h * 3 + j * 3
            res[di:di + 3] = pol[fi]
    return res


def fri_fold(steps, step, pol, challenge, split=False):
    """fri.js:22-81.  steps = [nBits,...]; pol = list of F3.  Returns {pol, tree, proof}."""
    pol_bits = len(pol).bit_length() - 1
    assert 1 << pol_bits == len(pol), "Invalid polynomial size"
    if step == 0:
        assert pol_bits == steps[0], "Invalid polynomial size"
    shift_inv = SHIFT_INV
    if step > 0:
        for _ in range(steps[0] - steps[step - 1]):
            shift_inv = (shift_inv * shift_inv) % P
    reduction_bits = pol_bits - steps[step]
    pol2n = 1 << (pol_bits - reduction_bits)
    nx = len(pol) // pol2n
    pol2 = [None] * pol2n
    sinv = shift_inv
    wi = inv(root_of_unity(pol_bits))
    for g in range(len(pol) // nx):
        if step == 0:
            pol2[g] = list(pol[g])
        else:
            ppar = [pol[i * pol2n + g] for i in range(nx)]
            c = intt(ppar)
            pol_mul_axi(c, 1, sinv)
            pol2[g] = eval_pol(c, challenge)
            sinv = (sinv * wi) % P
    tree, proof = None, None
    if step != len(steps) - 1:
        n_groups = 1 << steps[step + 1]
        group_size = (1 << steps[step]) // n_groups
        rows = transposed_buffer(pol2, steps[step + 1])
        tree = merkelize(rows, 3 * group_size, n_groups, split)
        proof = {"root": merkle_root(tree)}
    else:
        proof = [list(x) for x in pol2]
    return {"pol": pol2, "tree": tree, "proof": proof}


def fri_proof_queries(steps, proof, trees, fri_queries):
    """fri.js:83-105 (mutates fri_queries exactly like the reference)."""
    for step in range(len(steps)):
        proof[step]["polQueries"] = []
        if step == 0:
            for q in fri_queries:
                proof[step]["polQueries"].append([get_group_proof(t, q) for t in trees[step]])
        else:
            for i in range(len(fri_queries)):
                fri_queries[i] = fri_queries[i] % (1 << steps[step])
            for q in fri_queries:
                proof[step]["polQueries"].append(get_group_proof(trees[step], q))


def fri_verify_fold(pgroup, pol_bits, shift, challenge, query):
    """One link of fri.js:121-127: ifft the queried group, evaluate at challenge/(shift*w^query)."""
    c = intt([list(x) for x in pgroup])
    sinv = inv((shift * pow(root_of_unity(pol_bits), query, P)) % P)
    return eval_pol(c, f3_mul_scalar(challenge, sinv))


# ----------------------------------------------------------------------------------------------
# quotient commit, evaluations at xi, x/(x - xi)     src/stark/stark_gen_helpers.js
# ----------------------------------------------------------------------------------------------
def compute_q(q_ext, q_dim, q_deg, n_bits, n_bits_ext):
    """computeQStark, stark_gen_helpers.js:168-192 (up to the merkelize at :198): returns cmQ_ext as a flat row-major
    list of extN x (q_dim*q_deg).  qq2 rows >= N stay zero (a fresh BigBuffer)."""
    n, ne = 1 << n_bits, 1 << n_bits_ext
    qq1 = fft_p(q_ext, q_dim, n_bits_ext, inverse=True)                     # :177
    qq2 = [0] * (q_dim * q_deg * ne)                                       # :175
    cur_s = 1
    shift_in = pow(inv(SHIFT), n, P)                                       # :180
    for p in range(q_deg):                                                 # :181-190
        for i in range(n):
            for k in range(q_dim):
                qq2[i * q_dim * q_deg + q_dim * p + k] = (qq1[p * n * q_dim + i * q_dim + k] * cur_s) % P
        cur_s = (cur_s * shift_in) % P
    return fft_p(qq2, q_dim * q_deg, n_bits_ext)                           # :192


def opening_xi(xi_challenge, opening, n_bits):
    """xi * w^opening (stark_gen_helpers.js:222-226 / :291-300); xi_challenge in F3."""
    w = 1
    for _ in range(abs(opening)):
        w = (w * root_of_unity(n_bits)) % P
    if opening < 0:
        w = inv(w)
    return f3_mul_scalar(xi_challenge, w)


def compute_lev(xi_challenge, opening, n_bits):
    """LEv[i] of computeEvalsStark (stark_gen_helpers.js:216-231): ifft of the powers of xi*w^opening/shift (F3)."""
    n = 1 << n_bits
    xi = f3_mul_scalar(opening_xi(xi_challenge, opening, n_bits), SHIFT_INV)   # :227
    lev = [[1, 0, 0]]
    for _ in range(1, n):                                                   # :228-230
        lev.append(f3_mul(lev[-1], xi))
    return intt(lev)                                                        # :231


def compute_evals(buffers, ev_map, levs, n_bits, extend_bits):
    """Evaluation loop of computeEvalsStark (stark_gen_helpers.js:234-267).  buffers: name -> (flat list, row size);
    ev_map: list of (buffer name, column offset, dim in {1,3}, index into levs).  Returns a list of F3 values."""
    n = 1 << n_bits
    out = []
    for name, offset, dim, oi in ev_map:
        buf, size = buffers[name]
        acc = [0, 0, 0]
        for k in range(n):
            base = (k << extend_bits) * size + offset                      # :252-259
            if dim == 1:
                t = f3_mul_scalar(levs[oi][k], buf[base])
            else:
                t = f3_mul(buf[base:base + 3], levs[oi][k])
            acc = f3_add(acc, t)                                           # :261
        out.append(acc)
    return out


def x_div_x_sub_xi(xi_challenge, openings, n_bits, n_bits_ext):
    """xDivXSubXi_ext of computeFRIStark (stark_gen_helpers.js:289-323): flat list of extN x nOpenings F3 values,
    element (k, i) = x_k / (x_k - xi*w^opening_i), x_k = 7*w_ext^k.  (The reference's batch inverse is an exact field
    inverse, f3g.js:370-385.)"""
    ne = 1 << n_bits_ext
    no = len(openings)
    out = [0] * (3 * ne * no)
    w_ext = root_of_unity(n_bits_ext)
    for i, opening in enumerate(openings):
        xi = opening_xi(xi_challenge, opening, n_bits)
        x = SHIFT
        for k in range(ne):
            den = f3_sub([x, 0, 0], xi)                                    # :309
            v = f3_mul_scalar(f3_inv(den), x)                              # :312-315
            out[3 * (k * no + i):3 * (k * no + i) + 3] = v                 # :316-318
            x = (x * w_ext) % P
    return out


# ----------------------------------------------------------------------------------------------
# FRI polynomial: the fixed-form expression friExp             src/pil_info/helpers/polynomials/friPolinomial.js
# ----------------------------------------------------------------------------------------------
def js_object_key_order(keys):
    """Order in which JavaScript enumerates the own keys of `friExps` (friPolinomial.js:44): canonical non-negative integer
    keys ascending first, then the remaining (negative) keys in insertion order.  `keys` = primes in order of first use."""
    ints = sorted(k for k in keys if k >= 0)
    return ints + [k for k in keys if k < 0]


def fri_polynomial(buffers, ev_map, evals, openings, xdiv, vf1, vf2, n_bits_ext):
    """Row-by-row evaluation of friExp (friPolinomial.js:26-56) as callCalculateExps(..., "ext") does for computeFRIStark
    (stark_gen_helpers.js:325): buffers: name -> (flat list, row size); ev_map: list of (buffer name, column offset, dim, prime);
    evals: list of F3 (ctx.evals); xdiv: flat xDivXSubXi_ext; vf1, vf2 in F3.  Returns the 2^n_bits_ext F3 values of f_ext."""
    ne = 1 << n_bits_ext
    no = len(openings)
    first_use = []
    for _, _, _, prime in ev_map:
        if prime not in first_use:
            first_use.append(prime)
    order = js_object_key_order(first_use)
    out = []
    for k in range(ne):
        fri_exps = {}
        for i, (name, offset, dim, prime) in enumerate(ev_map):
            buf, size = buffers[name]
            base = k * size + offset
            e = [buf[base], 0, 0] if dim == 1 else list(buf[base:base + 3])
            term = f3_sub(e, evals[i])                                              # E.sub(e, E.eval(i, 3))          :36,38
            fri_exps[prime] = f3_add(f3_mul(fri_exps[prime], vf2), term) if prime in fri_exps else term
        acc = None
        for opening in order:                                                       # :44-53
            idx = openings.index(opening)
            f = f3_mul(fri_exps[opening], xdiv[3 * (k * no + idx):3 * (k * no + idx) + 3])
            acc = f3_add(f3_mul(vf1, acc), f) if acc is not None else f
        out.append(acc)
    return out
