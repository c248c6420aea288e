#!/usr/bin/env python
"""The commit-side stages of a pil2-stark proof on one GPU, driven through the mirror of the reference's own callers
(pil2_stark_js_b200.stark_gen_helpers == src/stark/stark_gen_helpers.js):

    stage commits (extendAndMerkelize) -> quotient commit (computeQStark) -> evaluations at xi (computeEvalsStark)
    -> FRI polynomial (computeFRIStark) -> FRI folding + layer commits (computeFRIFolding) -> queries (computeFRIQueries)
    -> the verifier's Merkle / fold / final-degree checks on the opened queries

The trace, the quotient and the challenges are synthetic (constraint evaluation is AIR-specific and stays in the reference's
JavaScript); everything downstream of them is the real pipeline, so the run ends with the same checks stark_verify.js makes
on the FRI part of a proof.     usage: python examples/stage_flow.py [nBits=16] [columns=64] [host|device]
"device" keeps the committed buffers and trees in HBM (ctx.device_resident): only roots, evaluations, f_ext and the opened
rows cross PCIe; "host" (default) returns every extended buffer to the host like the reference's BigBuffers.
"""
import sys
import time
import types
import pathlib

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import pil2_stark_js_b200 as m
from pil2_stark_js_b200 import stark_gen_helpers as H

P = 0xFFFFFFFF00000001


def field(rng, n):
    a = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a)


def main():
    n_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    device = len(sys.argv) > 3 and sys.argv[3] == "device"
    ext_bits = n_bits + 1
    N, extN = 1 << n_bits, 1 << ext_bits
    steps = [ext_bits]
    while steps[-1] > 5:
        steps.append(max(5, steps[-1] - 4))
    rng = np.random.default_rng(1)
    gpu = m.default_context(0)
    pil = {"nStages": 1, "qDim": 3, "qDeg": 2, "mapSectionsN": {"cm1": cols, "cm2": 6}, "nConstants": 4, "openingPoints": [0, 1],
           "cmPolsMap": [{"stage": 1, "stagePos": c, "dim": 1} for c in range(cols)] + [{"stage": 2, "stagePos": 0, "dim": 3},
                                                                                          {"stage": 2, "stagePos": 3, "dim": 3}],
           "starkStruct": {"nBits": n_bits, "nBitsExt": ext_bits, "nQueries": 32, "steps": [{"nBits": b} for b in steps]}}
    pil["evMap"] = ([{"type": "cm", "id": c, "prime": 0} for c in range(cols)] + [{"type": "cm", "id": c, "prime": 1} for c in range(0, cols, 2)] +
                    [{"type": "cm", "id": cols, "prime": 0}, {"type": "cm", "id": cols + 1, "prime": 0}] +
                    [{"type": "const", "id": c, "prime": 0} for c in range(4)])
    ctx = types.SimpleNamespace(pilInfo=pil, nBits=n_bits, nBitsExt=ext_bits, N=N, extN=extN, extendBits=1, trees={}, gpu=gpu,
                                MH=m.buildMerkleHash(False, gpu), challenges={}, cm2_ext=None, device_resident=device, dev_buffers=None)
    ctx.cm1_n = field(rng, cols * N)
    ctx.cm1_ext = np.empty(cols * extN, dtype=np.uint64)
    const_n = field(rng, 4 * N)
    if device:
        ctx.constTree, _ = gpu.commit(const_n, 4, n_bits, ext_bits)
        ctx.dev_buffers = {"const_ext": ctx.constTree.elements_ptr}
    else:
        ctx.const_ext = np.empty(4 * extN, dtype=np.uint64)
        m.interpolate(const_n, 4, n_bits, ctx.const_ext, ext_bits, ctx=gpu)
        ctx.constTree = ctx.MH.merkelize(ctx.const_ext, 4, extN)
    # a synthetic quotient of degree < qDeg * N: its evaluations on the extended coset
    q_n = field(rng, 3 * extN)
    ctx.q_ext = np.empty(3 * extN, dtype=np.uint64)
    m.fft(q_n, 3, ext_bits, ctx.q_ext, ctx=gpu)

    t = m.Transcript(gpu)
    timings = []

    def timed(name, fn, *a):
        t0 = time.perf_counter()
        r = fn(*a)
        gpu.sync()
        timings.append((name, time.perf_counter() - t0))
        return r

    t.put(ctx.MH.root(ctx.constTree))
    t.put(timed("stage 1: extendAndMerkelize", H.extendAndMerkelize, 1, ctx)[0])
    t.put(timed("stage Q: computeQStark", H.computeQStark, ctx)[0])
    ctx.challenges[2] = [t.getField()]                                   # xi
    evals = timed("evaluations: computeEvalsStark", H.computeEvalsStark, ctx)
    for e in evals:
        t.put(e)
    ctx.challenges[4] = [t.getField(), t.getField()]                     # vf1, vf2
    f = timed("FRI polynomial: computeFRIStark", H.computeFRIPol, ctx)
    ctx.fri = m.FRI(pil["starkStruct"], ctx.MH)
    ctx.friPol, ctx.friProof, ctx.friTrees = {0: f}, {0: {}}, {}
    ctx.friTrees[0] = [ctx.trees[1], ctx.trees[2], ctx.constTree]
    fri_challenges = []
    for step in range(len(steps)):
        ch = t.getField()
        fri_challenges.append(ch)
        out = timed(f"FRI step {step}: fold + layer commit", H.computeFRIFolding, step, ctx, ch)
        t.put(out if step + 1 < len(steps) else [list(map(int, e)) for e in np.asarray(out).reshape(-1, 3)])
    queries = H.getPermutationsStark(ctx, t.getField())
    timed("queries: computeFRIQueries", H.computeFRIQueries, ctx, list(queries))

    # ---- what the verifier checks on the FRI part (fri.js:107-174) ----
    MH = ctx.MH
    ok_paths = 0
    for qi, q in enumerate(queries):
        for tree, (vals, mp) in zip(ctx.friTrees[0], ctx.friProof[0]["polQueries"][qi]):
            assert MH.verifyGroupProof(MH.root(tree), mp, q, vals)
            ok_paths += 1
    last = np.asarray(ctx.friPol[len(steps)]).reshape(-1, 3)
    coef = np.empty(last.size, dtype=np.uint64)
    gpu.ntt(np.ascontiguousarray(last).reshape(-1), 3, steps[-1], coef, inverse=True)
    max_deg = 1 << (steps[-1] - (ext_bits - n_bits))
    assert not coef.reshape(-1, 3)[max_deg:].any(), "final FRI polynomial exceeds the degree bound"
    print(f"2^{n_bits} rows x {cols} columns, blowup 2, FRI steps {steps}, {len(queries)} queries")
    print("  mode: " + ("device-resident buffers and trees" if device else
                        "host buffers (pageable numpy arrays: every stage pays its PCIe transfers; bench.py has the pinned numbers)"))
    for name, dt in timings:
        print(f"  {name:44s} {dt * 1e3:9.2f} ms")
    print(f"verified {ok_paths} Merkle paths of the stage/const trees; final polynomial has degree < {max_deg}: OK")


if __name__ == "__main__":
    main()
