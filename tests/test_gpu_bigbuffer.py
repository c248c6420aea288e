"""The BigBuffer (paged) twins of every host-pointer entry point: pilcom BigBuffers are lists of BigUint64Array pages, and the
reference hands them to fft / ifft / interpolate / merkelize / extendAndMerkelize / computeQStark at any size
(fft_p.js:178-297, stark_gen_helpers.js:168-208,388-412).  Pages here are pageable numpy arrays (staged through the pinned ring
by the copy threads) or pinned (pil2gpu_host_alloc, copied by DMA directly); page cuts are ragged, fall inside rows, and include
empty pages.  Every result is compared with the oracle, bit-exact."""
import ctypes

import numpy as np
import pytest

from oracle import gl_spec as S
from oracle import gl_oracle as C

pytestmark = pytest.mark.gpu
P = S.P


@pytest.fixture(scope="module")
def ctx():
    import pil2_stark_js_b200 as m
    return m.default_context(0)


def rnd_field(seed, n):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a)


def ragged(total, seed, n_pages=5):
    """page lengths adding up to `total`, with one empty page"""
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.integers(0, total + 1, size=n_pages - 2))
    lens = np.diff(np.concatenate([[0], cuts, [total]])).tolist()
    lens.insert(1, 0)
    return lens


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("bits,npols", [(9, 12), (13, 3), (0, 5)])
def test_ntt_bigbuffer(ctx, bits, npols, pinned):
    from pil2_stark_js_b200 import BigBuffer, fft_p
    src = rnd_field(bits + npols, npols << bits)
    bs = BigBuffer.from_array(src, pinned=pinned, page_lens=ragged(src.size, 1))
    bd = BigBuffer(src.size, pinned=pinned, page_lens=ragged(src.size, 2, 4))
    fft_p.fft(bs, npols, bits, bd)
    assert np.array_equal(bd.to_array(), C.ntt(src, npols, bits))
    assert np.array_equal(bs.to_array(), src)                     # buffSrc untouched
    fft_p.ifft(bd, npols, bits, bs)
    assert np.array_equal(bs.to_array(), src)


@pytest.mark.parametrize("pinned", [False, True])
def test_interpolate_and_merkelize_bigbuffer(ctx, pinned):
    from pil2_stark_js_b200 import BigBuffer, fft_p, buildMerkleHash
    bits, ext, npols = 10, 12, 20
    src = rnd_field(5, npols << bits)
    bs = BigBuffer.from_array(src, pinned=pinned, page_lens=ragged(src.size, 3))
    bd = BigBuffer(npols << ext, pinned=pinned, page_len=30001)    # pages cut inside rows
    bd.set(np.full(npols << ext, 0xDEADBEEF, dtype=np.uint64))      # prior contents must not matter
    fft_p.interpolate(bs, npols, bits, bd, ext)
    want = C.lde(src, npols, bits, ext)
    assert np.array_equal(bd.to_array(), want)
    for split in (False, True):
        MH = buildMerkleHash(split)
        tree = MH.merkelize(bd, npols, 1 << ext)
        assert tree["elements"] is bd
        assert np.array_equal(tree["nodes"], C.merkelize(want, npols, 1 << ext, split))


@pytest.mark.parametrize("pinned,page_rows", [(False, 1 << 10), (True, 1 << 11), (False, None), (True, None)])
@pytest.mark.parametrize("bits,ext,npols,split", [(12, 13, 128, False), (12, 13, 64, True), (12, 14, 48, False)])
def test_extend_and_merkelize_bigbuffer(ctx, bits, ext, npols, split, pinned, page_rows):
    """page_rows: pages hold whole rows (the slab pipeline: one strided copy per slab and page; shape 1 takes it);
    None: ragged pages (upload -> LDE -> download under the hashing)."""
    from pil2_stark_js_b200 import BigBuffer
    src = rnd_field(bits * 3 + npols, npols << bits)
    if page_rows:
        bs = BigBuffer.from_array(src, pinned=pinned, page_len=page_rows * npols)
        bd = BigBuffer(npols << ext, pinned=pinned, page_len=3 * page_rows * npols)
    else:
        bs = BigBuffer.from_array(src, pinned=pinned, page_lens=ragged(src.size, 7))
        bd = BigBuffer(npols << ext, pinned=pinned, page_lens=ragged(npols << ext, 8, 6))
    out, nodes, root = ctx.extend_and_merkelize(bs, npols, bits, ext, split, dst=bd)
    assert out is bd
    want = C.lde(src, npols, bits, ext)
    assert np.array_equal(bd.to_array(), want)
    want_nodes = C.merkelize(want, npols, 1 << ext, split)
    assert np.array_equal(nodes, want_nodes) and np.array_equal(root, want_nodes[-4:])
    _, _, root2 = ctx.extend_and_merkelize(bs, npols, bits, ext, split, want_dst=False, want_nodes=False)
    assert np.array_equal(root2, root)


def test_stage_helpers_with_bigbuffers(ctx):
    """extendAndMerkelize / computeQStark (stark_gen_helpers.js:388-412,168-208) on a prover context whose cm*_n / cm*_ext /
    q_ext are BigBuffers, as in the reference (stark_gen_helpers.js:104-109)."""
    import types
    from pil2_stark_js_b200 import BigBuffer, buildMerkleHash, stark_gen_helpers as H
    n_bits, ext, cols = 9, 10, 10
    pil = {"mapSectionsN": {"cm1": cols, "cm2": 6}, "nStages": 1, "qDim": 3, "qDeg": 2}
    tr = rnd_field(1, cols << n_bits)
    q = rnd_field(2, 3 << ext)
    pctx = types.SimpleNamespace(pilInfo=pil, nBits=n_bits, nBitsExt=ext, N=1 << n_bits, extN=1 << ext, trees={}, MH=buildMerkleHash(False), gpu=ctx,
                                 cm1_n=BigBuffer.from_array(tr, page_len=777), cm1_ext=BigBuffer(cols << ext, page_len=1001),
                                 q_ext=BigBuffer.from_array(q, page_len=500), cm2_ext=BigBuffer(6 << ext, page_len=4099))
    root1 = H.extendAndMerkelize(1, pctx)
    e1 = C.lde(tr, cols, n_bits, ext)
    n1 = C.merkelize(e1, cols, 1 << ext)
    assert np.array_equal(pctx.cm1_ext.to_array(), e1) and root1 == [[int(x) for x in n1[-4:]]]
    assert pctx.trees[1]["elements"] is pctx.cm1_ext and np.array_equal(pctx.trees[1]["nodes"], n1)
    rootq = H.computeQStark(pctx)
    eq = C.compute_q(q, 3, 2, n_bits, ext)
    nq = C.merkelize(eq, 6, 1 << ext)
    assert np.array_equal(pctx.cm2_ext.to_array(), eq) and rootq == [[int(x) for x in nq[-4:]]]


def test_fri_fold_paged(ctx):
    from pil2_stark_js_b200 import BigBuffer
    from pil2_stark_js_b200._lib import check
    prev, cur, nxt = 12, 8, 4
    pol = rnd_field(9, 3 << prev)
    ch = rnd_field(10, 3)
    bp = BigBuffer.from_array(pol, page_lens=ragged(pol.size, 11))
    bo = BigBuffer(3 << cur, page_len=100)
    br = BigBuffer(3 << cur, page_len=333, pinned=True)
    nodes = np.empty(ctx.merkle_nnodes(1 << nxt), dtype=np.uint64)
    pp, pw, pn = bp.pages(); op, ow, on = bo.pages(); rp, rw, rn = br.pages()
    check(ctx._L.pil2gpu_fri_fold_paged(ctx.handle, pp, pw, pn, prev, cur, nxt, prev, ch.ctypes.data, 0, op, ow, on, rp, rw, rn, nodes.ctypes.data))
    ep, erows = C.fri_fold(pol.reshape(-1, 3), prev, cur, nxt, prev, [int(x) for x in ch])
    assert np.array_equal(bo.to_array(), ep.reshape(-1)) and np.array_equal(br.to_array(), erows)
    assert np.array_equal(nodes, C.merkelize(erows, 3 << (cur - nxt), 1 << nxt))


def test_paged_argument_checks(ctx):
    import pil2_stark_js_b200 as m
    from pil2_stark_js_b200 import BigBuffer
    src = BigBuffer.from_array(rnd_field(1, 64), page_len=10)
    with pytest.raises(ValueError):
        ctx.ntt(src, 4, 5, BigBuffer(64))                          # 4 << 5 != 64
    short = BigBuffer(60, page_len=7)
    p, w, n = short.pages(); sp, sw, sn = src.pages()
    assert ctx._L.pil2gpu_ntt_paged(ctx.handle, sp, sw, sn, p, w, n, 4, 4, 0) == -1 and "pages hold" in ctx._L.pil2gpu_last_error().decode()
    nullp = (ctypes.c_void_p * 2)(0, 0)
    nw = (ctypes.c_uint64 * 2)(32, 32)
    assert ctx._L.pil2gpu_ntt_paged(ctx.handle, nullp, nw, 2, sp, sw, sn, 4, 4, 0) == -1 and "null" in ctx._L.pil2gpu_last_error().decode()
    # still healthy
    dst = BigBuffer(64, page_len=9)
    ctx.ntt(src, 4, 4, dst)
    assert np.array_equal(dst.to_array(), C.ntt(src.to_array(), 4, 4))


def test_staged_copy_large_pageable(ctx):
    """A pageable buffer larger than the whole staging ring (4 x 32 MiB): every slot is reused several times in both directions
    (flat chunks through pil2gpu_ntt, strided slabs through the pipelined commit)."""
    bits, npols = 20, 32                                            # 256 MiB in, 256 MiB out
    src = rnd_field(77, npols << bits)
    dst = np.empty_like(src)
    back = np.empty_like(src)
    ctx.ntt(src, npols, bits, dst)
    ctx.ntt(dst, npols, bits, back, inverse=True)
    assert np.array_equal(back, src)
    cols = 128                                                      # slab pipeline with pageable host memory
    tr = rnd_field(78, cols << 16)
    out, nodes, root = ctx.extend_and_merkelize(tr, cols, 16, 17)
    want = C.lde(tr, cols, 16, 17)
    assert np.array_equal(out, want)
    assert np.array_equal(root, C.merkelize(want, cols, 1 << 17)[-4:])


def test_addon_end_to_end_under_the_mock_napi(ctx, tmp_path):
    """With a GPU the same driver also runs REAL calls through the addon's marshalling (page lists, async work, device-tree
    handles): fft/ifft round trip over ragged pages, extendAndMerkelizePaged vs commit vs merkelizePaged roots, group proofs from
    the device tree, "Out of range", pinned page allocation."""
    import subprocess
    from test_boundary_cpu import _build_addon_driver
    exe = _build_addon_driver(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL CHECKS PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
    assert "functional checks through the addon" in r.stdout and "FAIL" not in r.stdout
