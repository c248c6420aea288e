"""Python mirror of the prover-side callers of the commit path, src/stark/stark_gen_helpers.js (same function names,
same ctx fields, same results), so that the stage loop of the reference prover (src/prover/prover.js:42-122) can be driven
against the GPU library:

    extendAndMerkelize(stage, ctx)          :388-412   interpolate + merkelize
    computeQStark(ctx)                      :168-208   ifft -> shift-split -> fft -> merkelize
    computeEvalsStark(ctx)                  :210-273   LEv vectors + evaluation sums
    computeXDivXSubXi(ctx)                  :289-323   the xDivXSubXi_ext part of computeFRIStark
    computeFRIPol(ctx)                      :325-334   friExp over the extended domain -> f_ext, friPol[0]
    computeFRIFolding(step, ctx, challenge) :337-356
    computeFRIQueries(ctx, friQueries)      :358-360
    getPermutationsStark(ctx, challenge)    :474-493

With ctx.device_resident = True the committed buffers and trees stay in HBM (ctx.trees[stage] is a DeviceTree, ctx.dev_buffers
maps "cmK_ext" to its device address): evaluations, the FRI polynomial and the query openings then read them in place and
only roots, evaluations, f_ext and the opened rows cross PCIe -- the integration INTEGRATION.md section 3 describes.

`ctx` is any attribute bag (types.SimpleNamespace) carrying the reference's field names: pilInfo (dict with qDim, qDeg,
nStages, mapSectionsN, openingPoints, evMap, cmPolsMap, nConstants, starkStruct), nBits, nBitsExt, N, extN, extendBits,
cm<stage>_n / cm<stage>_ext / const_ext / q_ext (numpy uint64, row-major), trees, MH, challenges, fri, friPol, friProof,
friTrees, and `gpu` (a pil2_stark_js_b200.Context).  Expression evaluation (callCalculateExps) is not part of this path.
"""
import numpy as np

from .context import default_context
from .transcript import Transcript


def _gpu(ctx):
    g = getattr(ctx, "gpu", None) or getattr(getattr(ctx, "MH", None), "ctx", None)
    return g if g is not None else default_context()


def _split(ctx):
    return bool(getattr(ctx.MH, "splitLinearHash", False))


def _dev_buffers(ctx):
    if getattr(ctx, "dev_buffers", None) is None:
        ctx.dev_buffers = {}
    return ctx.dev_buffers


def _tree(buff, nodes, width, height):
    return {"elements": buff, "nodes": nodes, "width": width, "height": height}


def extendAndMerkelize(stage, ctx, options=None):
    """stark_gen_helpers.js:388-412.  One fused call: the extended buffer stays in HBM between the LDE and the hashing."""
    n_pols = ctx.pilInfo["mapSectionsN"].get("cm%d" % stage, 0)
    buff_from = getattr(ctx, "cm%d_n" % stage)
    if getattr(ctx, "device_resident", False):
        tree, root = _gpu(ctx).commit(buff_from, n_pols, ctx.nBits, ctx.nBitsExt, split=_split(ctx))
        ctx.trees[stage] = tree
        _dev_buffers(ctx)["cm%d_ext" % stage] = tree.elements_ptr
        return [[int(x) for x in root]]
    buff_to = getattr(ctx, "cm%d_ext" % stage)                 # caller-allocated (numpy array or BigBuffer), filled in place
    _, nodes, root = _gpu(ctx).extend_and_merkelize(buff_from, n_pols, ctx.nBits, ctx.nBitsExt, split=_split(ctx), dst=buff_to)
    ctx.trees[stage] = _tree(buff_to, nodes, n_pols, ctx.extN)
    return [[int(x) for x in root]]


def computeQStark(ctx, options=None):
    """stark_gen_helpers.js:168-208."""
    q_stage = ctx.pilInfo["nStages"] + 1
    q_dim, q_deg = ctx.pilInfo["qDim"], ctx.pilInfo["qDeg"]
    if getattr(ctx, "device_resident", False):
        if ctx.pilInfo["mapSectionsN"].get("cm%d" % q_stage, 0) != q_dim * q_deg:
            raise ValueError("mapSectionsN.cm%d != qDim*qDeg" % q_stage)
        tree, root = _gpu(ctx).compute_q_tree(ctx.q_ext, q_dim, q_deg, ctx.nBits, ctx.nBitsExt, split=_split(ctx))
        ctx.trees[q_stage] = tree
        _dev_buffers(ctx)["cm%d_ext" % q_stage] = tree.elements_ptr
        return [[int(x) for x in root]]
    name = "cm%d_ext" % q_stage
    ext, nodes, root = _gpu(ctx).compute_q(ctx.q_ext, q_dim, q_deg, ctx.nBits, ctx.nBitsExt, split=_split(ctx), ext=getattr(ctx, name, None))
    if getattr(ctx, name, None) is None:
        setattr(ctx, name, ext)
    n_pols_q = ctx.pilInfo["mapSectionsN"].get("cm%d" % q_stage, 0)
    if n_pols_q != q_dim * q_deg:
        raise ValueError("mapSectionsN.cm%d (%d) != qDim*qDeg (%d)" % (q_stage, n_pols_q, q_dim * q_deg))
    ctx.trees[q_stage] = _tree(getattr(ctx, name), nodes, n_pols_q, ctx.extN)
    return [[int(x) for x in root]]


def _pol_ref_ext(ctx, ev):
    """getPolRef(ctx, id, "ext") (prover_helpers.js:305-321) / the const branch of computeEvalsStark (:238-245):
    (buffer name, row size, column offset, dim)."""
    if ev["type"] == "const":
        return "const_ext", ctx.pilInfo["nConstants"], ev["id"], 1
    if ev["type"] == "cm":
        p = ctx.pilInfo["cmPolsMap"][ev["id"]]
        st = "cm%d" % p["stage"]
        return st + "_ext", ctx.pilInfo["mapSectionsN"][st], p["stagePos"], p["dim"]
    raise ValueError("Invalid ev type: " + str(ev["type"]))


def computeEvalsStark(ctx, options=None):
    """stark_gen_helpers.js:210-273 (hashCommits == false branch): sets and returns ctx.evals (list of F3 values)."""
    g = _gpu(ctx)
    evals_stage = ctx.pilInfo["nStages"] + 1
    xi = ctx.challenges[evals_stage][0]
    openings = [int(o) for o in ctx.pilInfo["openingPoints"]]
    ev_map = ctx.pilInfo["evMap"]
    by_buffer = {}
    for i, ev in enumerate(ev_map):
        name, size, offset, dim = _pol_ref_ext(ctx, ev)
        by_buffer.setdefault((name, size), []).append((i, offset, dim, openings.index(int(ev["prime"]))))
    out = [None] * len(ev_map)
    dev = getattr(ctx, "dev_buffers", None) or {}  # optional: name -> DeviceBuffer / device address of buffers already in HBM
    levs = g.compute_levs(xi, openings, ctx.nBits) if any(name in dev for name, _ in by_buffer) else None
    for (name, size), items in by_buffer.items():
        descs = [(o, d, l) for _, o, d, l in items]
        if name in dev:
            vals = g.compute_evals(dev[name], size, ctx.nBits, ctx.nBitsExt, descs, levs, len(openings))
        else:
            vals = g.compute_evals_host(xi, openings, ctx.nBits, ctx.nBitsExt, getattr(ctx, name), size, descs)
        for (i, _, _, _), v in zip(items, vals):
            out[i] = [int(x) for x in v]
    if levs is not None:
        levs.free()
    ctx.evals = out
    return ctx.evals


def computeXDivXSubXi(ctx, options=None):
    """The xDivXSubXi_ext loop of computeFRIStark (stark_gen_helpers.js:289-323)."""
    evals_stage = ctx.pilInfo["nStages"] + 1
    xi = ctx.challenges[evals_stage][0]
    openings = [int(o) for o in ctx.pilInfo["openingPoints"]]
    ctx.xDivXSubXi_ext = _gpu(ctx).x_div_x_sub_xi_host(xi, openings, ctx.nBits, ctx.nBitsExt).reshape(-1)
    return ctx.xDivXSubXi_ext


def computeFRIPol(ctx, options=None):
    """The rest of computeFRIStark (stark_gen_helpers.js:325-334): evaluates friExp (friPolinomial.js:26-56) over the extended
    domain into ctx.f_ext and ctx.friPol[0].  Needs ctx.evals, ctx.challenges[nStages + 3] = [vf1, vf2] and the extended buffers."""
    g = _gpu(ctx)
    stage = ctx.pilInfo["nStages"] + 3
    vf1, vf2 = ctx.challenges[stage][0], ctx.challenges[stage][1]
    xi = ctx.challenges[ctx.pilInfo["nStages"] + 1][0]
    openings = [int(o) for o in ctx.pilInfo["openingPoints"]]
    dev = getattr(ctx, "dev_buffers", None) or {}  # optional: name -> DeviceBuffer / device address of buffers already in HBM
    refs = [_pol_ref_ext(ctx, ev) + (int(ev["prime"]),) for ev in ctx.pilInfo["evMap"]]
    if all(name in dev for name, _, _, _, _ in refs):
        xdiv = g.x_div_x_sub_xi(xi, openings, ctx.nBits, ctx.nBitsExt, download=False)
        f = g.fri_pol([(dev[name], size, offset, dim, prime) for name, size, offset, dim, prime in refs], ctx.evals, openings, xdiv, vf1, vf2,
                      ctx.nBitsExt)
        ctx.xDivXSubXi_ext = xdiv.download()
        xdiv.free()
    else:
        f, ctx.xDivXSubXi_ext = g.fri_pol_host([(getattr(ctx, name), size, offset, dim, prime) for name, size, offset, dim, prime in refs],
                                               ctx.evals, openings, xi, vf1, vf2, ctx.nBits, ctx.nBitsExt, want_xdiv=True)
    ctx.f_ext = f.reshape(-1)
    if not hasattr(ctx, "friPol") or ctx.friPol is None:
        ctx.friPol = {}
    ctx.friPol[0] = f
    return f


def computeFRIFolding(step, ctx, challenge, options=None):
    """stark_gen_helpers.js:337-356 (hashCommits == false)."""
    step_proof = ctx.fri.fold(step, ctx.friPol[step], challenge)
    ctx.friPol[step + 1] = step_proof["pol"]
    ctx.friProof[step + 1] = step_proof["proof"]
    n_steps = len(ctx.pilInfo["starkStruct"]["steps"])
    if step < n_steps - 1:
        ctx.friTrees[step + 1] = step_proof["tree"]
    if step + 1 < n_steps:
        return [ctx.friProof[step + 1]["root"]]
    return ctx.friPol[step + 1]


def computeFRIQueries(ctx, friQueries):
    """stark_gen_helpers.js:358-360."""
    ctx.fri.proofQueries(ctx.friProof, ctx.friTrees, friQueries)


def getPermutationsStark(ctx, challenge):
    """stark_gen_helpers.js:474-493 (verificationHashType == "GL")."""
    t = Transcript(_gpu(ctx))
    t.put(challenge)
    ss = ctx.pilInfo["starkStruct"]
    return t.getPermutations(ss["nQueries"], ss["steps"][0]["nBits"])
