"""Deterministic sm_all traces (2^10 rows) -- TEST INFRASTRUCTURE (fixture generator for the golden proof).

Restates the reference's test state machines so the committed proof fixture
(test/compressor/verifier.proof.zkin.json: root1; test/compressor/verifier.circom:802: rootC) can be
reproduced end to end: trace -> LDE x2 -> linear hash -> Merkle root.

Column order follows the PIL declaration order of test/state_machines/sm_all/all_main.pil (Global,
Fibonacci, Connection, Permutation, Plookup).
"""
from .gl_spec import P, K_CONN, root_of_unity

N_BITS = 10
N = 1 << N_BITS


def committed_trace(inp=(1, 2)):
    """Stage-1 witness, row-major N x 15.
    Fibonacci.l1,l2 (sm_fibonacci.js:12-22); Connection.a,b,c (sm_connection.js:31-53);
    Permutation.a,b,c,d,selC,selD (sm_permutation.js:6-24); Plookup.sel,a,b,cc (sm_plookup.js:23-60)."""
    n = N
    l1 = [0] * n; l2 = [0] * n
    l2[0] = inp[0]; l1[0] = inp[1]
    for i in range(1, n):
        l2[i] = l1[i - 1]
        l1[i] = (l2[i - 1] * l2[i - 1] + l1[i - 1] * l1[i - 1]) % P
    ca = list(range(n))
    cb = [ca[i * 2] if i < n // 2 else ca[(i - n // 2) * 2 + 1] for i in range(n)]
    cc_ = [cb[i * 2] if i < n // 2 else cb[(i - n // 2) * 2 + 1] for i in range(n)]
    pa = [0] * n; pb = [0] * n; pc = [0] * n; pd = [0] * n; selc = [0] * n; seld = [0] * n
    for i in range(n):
        pa[i] = i * i + i + 1
        pb[n - i - 1] = pa[i]
        if i % 2 == 0:
            selc[i] = 1; pc[i] = pa[i]; seld[i // 2] = 1; pd[i // 2] = pa[i]
        else:
            selc[i] = 0; pc[i] = 44; seld[n // 2 + (i - 1) // 2] = 0; pd[n // 2 + (i - 1) // 2] = 55
    sel = [0] * n; la = [0] * n; lb = [0] * n; lcc = [0] * n
    p = 0
    for i in range(16):
        for j in range(16):
            lcc[p] = i * j; p += 1
    while p < n:
        lcc[p] = p; p += 1
    p = 0
    for i in range(10):
        sel[p] = 1; la[p] = i; lb[p] = 55 if i == 0 else i + 3; p += 1
    sel[p] = 0; la[p] = 55; lb[p] = 10; p += 1
    while p < n:
        sel[p] = 0; la[p] = 55; lb[p] = 55; p += 1
    cols = [l1, l2, ca, cb, cc_, pa, pb, pc, pd, selc, seld, sel, la, lb, lcc]
    buff = [0] * (n * len(cols))
    for c, col in enumerate(cols):
        for r in range(n):
            buff[r * len(cols) + c] = col[r] % P
    return buff, len(cols)


def constant_trace():
    """Constant columns, row-major N x 9: Global.L1 (sm_global.js:1-6); Fibonacci.L1,LLAST
    (sm_fibonacci.js:1-9); Connection.S1,S2,S3 (sm_connection.js:4-27, ks = [k, k^2]);
    Plookup.SEL,A,B (sm_plookup.js:1-20)."""
    n = N
    g_l1 = [1 if i == 0 else 0 for i in range(n)]
    f_l1 = list(g_l1)
    f_last = [1 if i == n - 1 else 0 for i in range(n)]
    ks = [K_CONN, (K_CONN * K_CONN) % P]
    s1 = [0] * n; s2 = [0] * n; s3 = [0] * n
    w = 1
    wn = root_of_unity(N_BITS)
    for i in range(n):
        s1[i] = w; s2[i] = (w * ks[0]) % P; s3[i] = (w * ks[1]) % P
        w = (w * wn) % P

    def connect(p1, i1, p2, i2):          # src/helpers/polutils.js:166-168
        p1[i1], p2[i2] = p2[i2], p1[i1]

    for i in range(n):
        if i % 2 == 0:
            connect(s1, i, s2, i // 2); connect(s2, i, s3, i // 2)
        else:
            connect(s1, i, s2, n // 2 + (i - 1) // 2); connect(s2, i, s3, n // 2 + (i - 1) // 2)
    lsel = [0] * n; la = [0] * n; lb = [0] * n
    p = 0
    for i in range(16):
        for j in range(16):
            la[p] = i; lb[p] = j; lsel[p] = 1; p += 1
    cols = [g_l1, f_l1, f_last, s1, s2, s3, lsel, la, lb]
    buff = [0] * (n * len(cols))
    for c, col in enumerate(cols):
        for r in range(n):
            buff[r * len(cols) + c] = col[r] % P
    return buff, len(cols)
