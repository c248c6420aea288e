#!/usr/bin/env python
"""Headline benchmark: seconds per STARK commit (Goldilocks LDE + Poseidon-GL Merkle + FRI chain) on synthetic traces.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg5|tiny]

One "step" = one commit of the workload: interpolate (LDE) of the 2^n x C trace to blowup B, merkelize of the extended
buffer, then the FRI chain of BASELINE.json configs[3] scaled to the workload (fold 2^4 per step down to 2^4 points, one
layer tree per step) and 128 query openings on every tree.  `value` times it with everything resident in HBM;
`e2e` times the same commit through the host-buffer C-ABI calls a JS caller would make (pinned host buffers in,
extended buffer + nodes + FRI layers back out).  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "s per commit (GL NTT LDE+Poseidon Merkle+FRI) @2^23 rows x 256 cols, blowup 2"
WORKLOADS = {   # name: (nBits, nCols, blowup bits)
    "tiny": (14, 32, 1),
    "cfg2": (20, 64, 1),
    "slab": (23, 32, 1),    # one column slab of cfg3 (same pass structure 8/8/7): used for ncu captures of the NTT kernels
    "cfg3": (23, 256, 1),
    "cfg5": (22, 512, 2),
}
N_QUERIES = 128
P = 0xFFFFFFFF00000001


def fri_steps(ext_bits):
    steps = [ext_bits]
    while steps[-1] > 4:
        steps.append(max(4, steps[-1] - 4))
    return steps


def splitmix_field(seed, first, n):
    """numpy twin of pil2gpu_synth_dev (used by the CPU arms only)."""
    with np.errstate(over="ignore"):
        i = np.arange(first, first + n, dtype=np.uint64)
        z = (np.uint64(seed) ^ i) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return np.where(z >= np.uint64(P), z - np.uint64(P), z)


def splitmix_field_idx(seed, idx):
    """splitmix64(seed ^ idx) mod p for an array of u64 indices."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) ^ idx) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return np.where(z >= np.uint64(P), z - np.uint64(P), z)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arms (oracle).  bench.py is one of the few places allowed to execute oracle/ -- as the CPU baseline, never as
# part of the product path.
# ------------------------------------------------------------------------------------------------------------------
def cpu_commit_sample(n_bits, cols, blow, sample_bits, threads, seed):
    """Run the C oracle on a 2^sample_bits-row sample and extrapolate to 2^n_bits rows.
    NTT time scales with rows*log2(rows) per transform, hashing and FRI with rows."""
    from oracle import gl_oracle as C
    src = splitmix_field(seed, 0, cols << sample_bits)
    t0 = time.perf_counter()
    ext = C.lde(src, cols, sample_bits, sample_bits + blow, threads=threads)
    t1 = time.perf_counter()
    C.merkelize(ext, cols, 1 << (sample_bits + blow), threads=threads)
    t2 = time.perf_counter()
    # FRI chain on a sample-sized evaluation vector
    steps = fri_steps(sample_bits + blow)
    pol = splitmix_field(seed + 1, 0, 3 << steps[0]).reshape(-1, 3)
    ch = [int(x) for x in splitmix_field(seed + 2, 0, 3)]
    rows0 = np.ascontiguousarray(pol.reshape(1 << (steps[0] - steps[1]), 1 << steps[1], 3).transpose(1, 0, 2)).reshape(-1)
    C.merkelize(rows0, 3 << (steps[0] - steps[1]), 1 << steps[1], threads=threads)
    cur = pol
    for s in range(1, len(steps)):
        nxt = steps[s + 1] if s + 1 < len(steps) else None
        cur, rows = C.fri_fold(cur, steps[s - 1], steps[s], nxt, steps[0], ch, threads=threads)
        if nxt is not None:
            C.merkelize(rows, 3 << (steps[s] - nxt), 1 << nxt, threads=threads)
    t3 = time.perf_counter()
    ratio = float(1 << (n_bits - sample_bits))
    log_ratio = (n_bits + (n_bits + blow)) / float(sample_bits + (sample_bits + blow))
    t_lde, t_mk, t_fri = t1 - t0, t2 - t1, t3 - t2
    est = t_lde * ratio * log_ratio + t_mk * ratio + t_fri * ratio
    return {"sample_s": t3 - t0, "lde_s": t_lde, "merkle_s": t_mk, "fri_s": t_fri, "estimate_full_s": est}


def pick_sample_bits(n_bits, cols, threads, blow=1, target_s=8.0):
    """Rows of the CPU sample: the port's permutation rate is measured on a small tree first, then the sample is sized for
    about `target_s` seconds per step (hashing is ~85% of the port's time), so that a whole --steps K --warmup W run of the
    reference arm stays within a few minutes whatever the host."""
    from oracle import gl_oracle as C
    probe = splitmix_field(1, 0, 8 << 15)
    C.merkelize(probe[: 8 << 10], 8, 1 << 10, threads=threads)              # warm (library load, thread start)
    t0 = time.perf_counter()
    C.merkelize(probe, 8, 1 << 15, threads=threads)
    rate = (2 << 15) / max(time.perf_counter() - t0, 1e-6)                   # permutations per second, all threads
    perms_per_row = (1 << blow) * ((cols + 7) // 8 + 1)
    bits = int(np.log2(max(2.0, 0.85 * target_s * rate / perms_per_row)))
    return max(10, min(n_bits, bits))


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_bits, cols, blow = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample_bits = pick_sample_bits(n_bits, cols, threads, blow)
    times, walls, detail = [], [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        detail = cpu_commit_sample(n_bits, cols, blow, sample_bits, threads, 0x5EED0003)
        if i >= args.warmup:
            times.append(detail["estimate_full_s"])
            walls.append(detail["sample_s"])
        if time.perf_counter() - t0 > 60 and i >= args.warmup:
            break
    val = statistics.mean(times)
    sample = (f"C restatement of the reference algorithm (oracle/gl_oracle.c: radix-2 NTT, plain Poseidon, per-level Merkle, serial-"
              f"formulation FRI fold) on {threads} host threads; each step commits a 2^{sample_bits}-row x {cols}-col sample "
              f"(blowup {1 << blow}) and is extrapolated to 2^{n_bits} rows (NTT ~ rows*log2 rows, hashing/FRI ~ rows). "
              "Node.js is absent on this box, so the JS worker-thread path itself cannot run.")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": config_dict(args.workload, world),
        "cpu_baseline": {"value": val, "unit": "s", "cores": threads, "kind": "port", "sample": sample,
                         "sample_wall_s": detail["sample_s"]},
        "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # `value` is NOT the wall time of a step: every step ran a 2^sample_bits-row sample and scaled it (see `sample`);
        # the law is validated against full-row-count runs of the port in profiles/r02_cpu_port_scaling.json
        "extrapolated": sample_bits < n_bits, "sample_rows": 1 << sample_bits, "full_rows": 1 << n_bits,
        "sample_wall_s_per_step": statistics.mean(walls), "extrapolation": "lde*rows_ratio*log_ratio + (merkle + fri)*rows_ratio",
    }
    print(json.dumps(line), flush=True)


def config_dict(workload, world):
    n_bits, cols, blow = WORKLOADS[workload]
    return {"workload": f"{workload}: 2^{n_bits} rows x {cols} cols, blowup {1 << blow}, standard linear hash, FRI steps "
                        f"{fri_steps(n_bits + blow)}, {N_QUERIES} queries", "rows": 1 << n_bits, "cols": cols, "blowup": 1 << blow,
            "l2": "inputs larger than L2 (no flush needed)" if (cols << (n_bits + 3)) > (256 << 20) else "L2 flushed between steps",
            "sharding": "single GPU" if world == 1 else f"columns/{world} for the LDE -> all-to-all -> rows/{world} for hashing -> gathered tree top"}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Gpu:
    """Thin ctypes driver around the C ABI with torch tensors as device memory."""

    def __init__(self, torch, device):
        from pil2_stark_js_b200 import _lib
        self.torch, self.L, self.check = torch, _lib.load(), _lib.check
        self.vp = ctypes.c_void_p
        h = self.vp()
        self.check(self.L.pil2gpu_create(device, self.vp(torch.cuda.current_stream().cuda_stream), ctypes.byref(h)))
        self.h = h

    def dev(self, words):
        return self.torch.empty(int(words), dtype=self.torch.int64, device="cuda")

    def ptr(self, t, off_words=0):
        return self.vp(t.data_ptr() + 8 * off_words)

    def launches(self):
        return int(self.L.pil2gpu_launch_count(self.h))

    def nnodes(self, h):
        return int(self.L.pil2gpu_merkle_nnodes(h))


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs next to its GPU (NVML's affinity mask) before any pinned buffer is allocated: pinned
    pages are placed on the node of the allocating thread, and a buffer on the far socket halves the PCIe copy rate when
    eight ranks move data at once.  PIL2GPU_NO_BIND=1 disables it."""
    if os.environ.get("PIL2GPU_NO_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",")]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def run_ours(args, rank, world, local_rank):
    bound = bind_to_gpu_numa(local_rank)
    if os.environ.get("PIL2GPU_TRACE"):
        print(f"[bench] rank {rank}: cpu affinity {'%d cpus %s..%s' % (len(bound), bound[0], bound[-1]) if bound else 'unchanged'}", file=sys.stderr, flush=True)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the commit path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world > 1:
        from pil2_stark_js_b200 import sharded
        return sharded.bench_main(args, rank, world, local_rank, dist, sys.modules[__name__])

    # a non-default torch stream: the ctx enqueues on it, so torch.cuda.Event timing sees every kernel of the library
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g = Gpu(torch, local_rank)
    L, check, vp = g.L, g.check, g.vp
    n_bits, cols, blow = WORKLOADS[args.workload]
    ext_bits = n_bits + blow
    free_b, _ = torch.cuda.mem_get_info()
    need = 8 * ((cols << n_bits) + (cols << ext_bits) + g.nnodes(1 << ext_bits) + 4 * (3 << ext_bits))
    if need > free_b * 0.9:
        raise SystemExit(f"bench.py: workload {args.workload} needs {need >> 30} GiB, device has {free_b >> 30} GiB free")
    seed = 0x5EED0000 + 3
    src = g.dev(cols << n_bits)
    dst = g.dev(cols << ext_bits)
    nodes = g.dev(g.nnodes(1 << ext_bits))
    check(L.pil2gpu_synth_dev(g.h, g.ptr(src), cols << n_bits, seed, 0))
    steps = fri_steps(ext_bits)
    fri_pol = [g.dev(3 << b) for b in steps]            # fri_pol[s] = evaluations after step s (fri_pol[0] = FRI polynomial)
    fri_rows = [g.dev(3 << steps[s]) for s in range(len(steps) - 1)]
    fri_nodes = [g.dev(g.nnodes(1 << steps[s + 1])) for s in range(len(steps) - 1)]
    check(L.pil2gpu_synth_dev(g.h, g.ptr(fri_pol[0]), 3 << steps[0], seed + 1, 0))
    chal = [np.ascontiguousarray(splitmix_field(seed + 2 + s, 0, 3)) for s in range(len(steps))]
    rng = np.random.default_rng(7)
    queries = rng.integers(0, 1 << ext_bits, size=N_QUERIES, dtype=np.uint64)
    root = np.zeros(4, dtype=np.uint64)
    tree_main = vp()
    check(L.pil2gpu_tree_wrap_dev(g.h, g.ptr(dst), g.ptr(nodes), cols, 1 << ext_bits, ctypes.byref(tree_main)))
    fri_trees = []
    for s in range(len(steps) - 1):
        t = vp()
        check(L.pil2gpu_tree_wrap_dev(g.h, g.ptr(fri_rows[s]), g.ptr(fri_nodes[s]), 3 << (steps[s] - steps[s + 1]), 1 << steps[s + 1],
                                      ctypes.byref(t)))
        fri_trees.append(t)
    depth_main = ext_bits
    q_rows = np.empty(N_QUERIES * cols, dtype=np.uint64)
    q_sib = np.empty(N_QUERIES * depth_main * 4, dtype=np.uint64)
    fq_rows = [np.empty(N_QUERIES * (3 << (steps[s] - steps[s + 1])), dtype=np.uint64) for s in range(len(steps) - 1)]
    fq_sib = [np.empty(N_QUERIES * max(1, steps[s + 1]) * 4, dtype=np.uint64) for s in range(len(steps) - 1)]
    npp = lambda a: vp(a.ctypes.data)

    def phase_lde():
        check(L.pil2gpu_lde_dev(g.h, g.ptr(src), g.ptr(dst), cols, n_bits, ext_bits))

    def phase_merkle():
        check(L.pil2gpu_merkelize_dev(g.h, g.ptr(dst), cols, 1 << ext_bits, 0, g.ptr(nodes)))

    def phase_fri():
        # step 0: identity fold (in place) + first layer tree; steps s >= 1 fold and commit the next layer
        check(L.pil2gpu_fri_fold_dev(g.h, g.ptr(fri_pol[0]), steps[0], steps[0], steps[1], steps[0], npp(chal[0]), 0, g.ptr(fri_pol[0]),
                                     g.ptr(fri_rows[0]), g.ptr(fri_nodes[0])))
        for s in range(1, len(steps)):
            last = s == len(steps) - 1
            check(L.pil2gpu_fri_fold_dev(g.h, g.ptr(fri_pol[s - 1]), steps[s - 1], steps[s], -1 if last else steps[s + 1], steps[0],
                                         npp(chal[s]), 0, g.ptr(fri_pol[s]), None if last else g.ptr(fri_rows[s]),
                                         None if last else g.ptr(fri_nodes[s])))

    def phase_queries():
        check(L.pil2gpu_tree_group_proofs(g.h, tree_main, npp(queries), N_QUERIES, npp(q_rows), npp(q_sib)))
        q = queries.copy()
        for s in range(len(steps) - 1):
            q = q % np.uint64(1 << steps[s + 1])
            check(L.pil2gpu_tree_group_proofs(g.h, fri_trees[s], npp(q), N_QUERIES, npp(fq_rows[s]), npp(fq_sib[s])))
        check(L.pil2gpu_tree_root(g.h, tree_main, npp(root)))

    def step():
        phase_lde(); phase_merkle(); phase_fri(); phase_queries()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = g.launches()
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    launches = g.launches() - l0
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    sec_per_commit = total_ms / 1e3 / args.steps
    root_dev = [int(x) for x in root]

    # per-phase device times (CUDA events on the launching stream, outside the headline region)
    def time_phase(fn, reps):
        a, b = ev(), ev()
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps / 1e3
    reps = max(1, min(args.steps, 3))
    t_lde, t_mk, t_fri = time_phase(phase_lde, reps), time_phase(phase_merkle, reps), time_phase(phase_fri, reps)

    # the dominant kernel (merkle_leaf_kernel, ~73% of the step): its own duration = merkelize - tree levels, both timed live
    def phase_tree():
        check(L.pil2gpu_merkle_tree_from_digests_dev(g.h, g.ptr(nodes), 1 << ext_bits))
    t_tree = time_phase(phase_tree, reps)
    phase_merkle()          # restore the nodes of the full merkelize (tree_from_digests rewrote only the upper levels, identically)
    t_leaf = max(t_mk - t_tree, 1e-9)

    # integer-pipe denominators, measured live
    mm, iw = ctypes.c_double(), ctypes.c_double()
    check(L.pil2gpu_bench_int_pipes(g.h, ctypes.byref(mm), ctypes.byref(iw)))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md, 6.65 TB/s)"
    E = 1 << ext_bits
    leaf_perms = E * ((cols + 7) // 8)
    leaf_bytes = 8 * cols * E + 32 * E                    # algorithmic: read every extended row once, write one digest per row
    traffic = TRAFFIC.get(args.workload, {})
    roofline = leaf_roofline(leaf_perms, leaf_bytes, t_leaf, sec_per_commit, mm.value, iw.value, hbm_peak, peak_src,
                             traffic.get("merkle_leaf_kernel"))
    lde_bytes = 8 * cols * (1 << n_bits) * (1 + (1 << blow))
    npass = (n_bits + 8) // 9
    roofline_lde = {"kernel": "ntt_pass_kernel x%d + ntt_lde_fused_kernel (whole LDE)" % (2 * npass - 2), "bound": "hbm",
                    "achieved": lde_bytes / t_lde / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": lde_bytes / t_lde / 1e9 / hbm_peak,
                    "peak_src": peak_src, "algorithmic_bytes": lde_bytes, "traffic": traffic.get("lde"),
                    "note": "integer-pipe bound as well: %.3g butterflies at the measured register-only butterfly rate is the floor"
                            % (3 * cols * (1 << n_bits) * n_bits / 2)}
    n_bfly = (1 + (1 << blow)) * cols * (1 << n_bits) * n_bits / 2          # INTT + one size-N NTT per coset, n/2 * log n butterflies per column each
    roofline_lde["int"] = {"achieved": n_bfly / t_lde / 1e12, "peak": mm.value / 1e12, "unit": "T butterflies/s", "frac": n_bfly / t_lde / mm.value,
                           "def": "one modular multiply per butterfly: peak = the live-measured register-only multiply rate (pil2gpu_bench_int_pipes); the "
                                  "register-only butterfly loop (multiply + add + sub) reaches 0.90 T/s in tools/probe, i.e. 0.64 of this peak",
                           "butterflies": n_bfly}

    # ---- parity spot check of the buffers the timed steps produced (outside every timed region) ----
    spot = None
    if not args.no_verify:
        step()                              # the buffers of one whole step, as timed above
        torch.cuda.synchronize()
        try:
            spot = parity_spot_check(torch, n_bits, cols, blow, seed, dst, nodes, root_dev, queries, q_rows, q_sib, steps, chal, fri_pol, fri_nodes,
                                     fq_rows, fq_sib)
        except Exception as ex:             # reported, never hidden: "parity_spot_check" then says why it did not run
            spot = {"status": "error: " + str(ex)[:200]}

    # ---- rows next to the commit: quotient commit, evaluations at xi, FRI denominators -- device-resident, timed
    # with CUDA events like the phases above; GB/s are ALGORITHMIC bytes over the measured time ----
    extras = None
    if not args.no_extras:
        try:
            extras = run_next_rows(g, L, check, vp, npp, torch, time_phase, reps, n_bits, cols, blow, ext_bits, seed, dst, hbm_peak)
        except RuntimeError as ex:          # reported next to the headline, never instead of it
            extras = {"error": str(ex)[:300]}

    # ---- e2e: the same commit through the host-buffer entry points (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(g, args, torch, n_bits, cols, blow, steps, src, fri_pol[0], chal, queries, root_dev)
    u64l = lambda t: [int(x) & 0xFFFFFFFFFFFFFFFF for x in t.cpu().tolist()]
    fri_out = {"fri_roots": [u64l(n[-4:]) for n in fri_nodes], "fri_final0": u64l(fri_pol[-1][:3])}
    return finish_line(args, n_bits, cols, blow, sec_per_commit, t_lde, t_mk, t_leaf, t_fri, roofline, roofline_lde, e2e, extras, launches, clocks,
                       root_dev, seed, spot, fri_out)


def parity_spot_check(torch, n_bits, cols, blow, seed, dst, nodes, root_dev, queries, q_rows, q_sib, steps, chal, fri_pol, fri_nodes,
                      fq_rows, fq_sib, n_cols_checked=16, n_leaves_checked=256):
    """Outside the timed region: the buffers the timed steps produced, against the CPU oracle (oracle/ is the checker here,
    never the thing measured).  (1) `n_cols_checked` random columns of the extended buffer re-derived with the oracle's LDE
    from the synthetic trace; (2) `n_leaves_checked` random leaf digests re-hashed from the extended rows; (3) every opened
    row + sibling path of the main tree verified up to the root (merklehash_p.js:170-209); (4) every opened FRI group
    verified to its layer root and folded with the step challenge into the next layer's opened value / the final polynomial
    (fri.js:121-127).  Returns a dict for the JSON line; raises nothing (a mismatch is reported, not hidden)."""
    from oracle import gl_oracle as C
    from oracle import gl_spec as S
    t0 = time.perf_counter()
    ext_bits = n_bits + blow
    N, E = 1 << n_bits, 1 << ext_bits
    rng = np.random.default_rng(20261018)
    res = {"columns": 0, "leaves": 0, "main_paths": 0, "fri_paths": 0, "fri_fold_links": 0, "failures": []}
    u64 = lambda t: t.cpu().numpy().view(np.uint64)
    # (1) columns
    pick = np.sort(rng.choice(cols, size=min(n_cols_checked, cols), replace=False))
    r = np.arange(N, dtype=np.uint64) * np.uint64(cols)
    src_cols = np.stack([splitmix_field_idx(seed, r + np.uint64(c)) for c in pick], axis=1).reshape(-1)
    want = C.lde(src_cols, len(pick), n_bits, ext_bits).reshape(E, len(pick))
    got = u64(dst.view(E, cols)[:, torch.as_tensor(pick, device=dst.device)].contiguous())
    bad = np.nonzero((got != want).any(axis=0))[0]
    res["columns"] = int(len(pick))
    if bad.size:
        res["failures"].append("extended columns %s differ from the oracle LDE" % [int(pick[b]) for b in bad])
    del want, got, src_cols
    # (2) leaf digests
    rows = np.unique(np.concatenate([rng.integers(0, E, size=n_leaves_checked), [0, E - 1]]))
    row_vals = u64(dst.view(E, cols)[torch.as_tensor(rows, device=dst.device)].contiguous())
    digs = u64(nodes.view(-1)[: 4 * E].view(E, 4)[torch.as_tensor(rows, device=dst.device)].contiguous())
    for k, rw in enumerate(rows):
        if not np.array_equal(C.linear_hash(row_vals[k]), digs[k]):
            res["failures"].append("leaf digest of row %d differs from the oracle linear hash" % int(rw))
    res["leaves"] = int(rows.size)

    def root_from_path(vals, sib, idx):
        v = C.linear_hash(vals)
        for s in sib:
            st = np.concatenate([v, s, np.zeros(4, dtype=np.uint64)]) if (idx & 1) == 0 else np.concatenate([s, v, np.zeros(4, dtype=np.uint64)])
            v = C.poseidon_perm(st)[:4]
            idx >>= 1
        return [int(x) for x in v]
    # (3) main-tree openings
    depth = ext_bits
    for k, q in enumerate(queries):
        if root_from_path(q_rows[k * cols:(k + 1) * cols], q_sib[k * depth * 4:(k + 1) * depth * 4].reshape(depth, 4), int(q)) != root_dev:
            res["failures"].append("opening %d of the main tree does not verify" % int(q))
    res["main_paths"] = int(len(queries))
    # (4) FRI layers
    final = u64(fri_pol[-1]).reshape(-1, 3)
    qs = [int(q) for q in queries]
    prev_groups = None
    for s in range(len(steps) - 1):
        gsz = 1 << (steps[s] - steps[s + 1])
        d = steps[s + 1]
        lroot = [int(x) for x in u64(fri_nodes[s][-4:])]
        qs = [q % (1 << steps[s + 1]) for q in qs]
        groups = []
        for k, q in enumerate(qs):
            vals = fq_rows[s][k * 3 * gsz:(k + 1) * 3 * gsz]
            if root_from_path(vals, fq_sib[s][k * d * 4:(k + 1) * d * 4].reshape(d, 4), q) != lroot:
                res["failures"].append("opening %d of FRI layer tree %d does not verify" % (q, s))
            groups.append([[int(x) for x in vals[3 * j:3 * j + 3]] for j in range(gsz)])
            res["fri_paths"] += 1
        # fold link: the group opened in layer tree s (values of fri_pol[s]) folds with chal[s + 1] into fri_pol[s + 1][q]
        shift = pow(7, 1 << (steps[0] - steps[s]), P)
        nxt_last = (s + 2 == len(steps))
        for k, q in enumerate(qs):
            ev = S.fri_verify_fold(groups[k], steps[s], shift, [int(x) for x in chal[s + 1]], q)
            if nxt_last:
                ref = [int(x) for x in final[q]]
            else:
                q2, j = q % (1 << steps[s + 2]), q >> steps[s + 2]
                gs2 = 1 << (steps[s + 1] - steps[s + 2])
                ref = [int(x) for x in fq_rows[s + 1][(k * gs2 + j) * 3:(k * gs2 + j) * 3 + 3]]
            if ev != ref:
                res["failures"].append("fold link layer %d -> %d at %d differs" % (s, s + 1, q))
            res["fri_fold_links"] += 1
    res["status"] = "ok" if not res["failures"] else "MISMATCH"
    res["failures"] = res["failures"][:8]
    res["seconds"] = round(time.perf_counter() - t0, 2)
    return res


def leaf_roofline(leaf_perms, leaf_bytes, t_leaf, sec_per_step, mulmod_per_s, imad_wide_per_s, hbm_peak, peak_src, traffic):
    """Roofline object of the dominant kernel (merkle_leaf_kernel).  The kernel is bound by the integer pipes, not by HBM
    ("= permutations/s / measured-IMAD-bound permutations/s for hash kernels"), so `bound` is "int": `achieved` =
    permutations/s, `peak` = the live-measured register-only Goldilocks multiply rate / 472 (the S-box multiplies of one
    permutation are irreducible; the MDS layers, constant additions and domain conversions count against the fraction).
    The HBM figures of the same launch ride along under "hbm"."""
    peak = mulmod_per_s / 472.0
    return {
        "kernel": "merkle_leaf_kernel", "bound": "int", "achieved": leaf_perms / t_leaf / 1e9, "peak": peak / 1e9, "unit": "Gperm/s",
        "frac": leaf_perms / t_leaf / peak, "traffic": traffic,
        "peak_src": "measured live: pil2gpu_bench_int_pipes mulmod/s / 472 S-box multiplies per permutation",
        "launch_s": t_leaf, "share_of_step": t_leaf / sec_per_step, "perms_per_launch": leaf_perms,
        "int_pipes": {"mulmod_per_s_measured": mulmod_per_s, "imad_wide_per_s_measured": imad_wide_per_s,
                      "frac_of_imad_wide_bound": leaf_perms / t_leaf / (imad_wide_per_s / 1888.0),
                      "def": "1888 = 472 multiplies x 4 IMAD.WIDE",
                      "ncu": "profiles/r02_ncu_cfg3.md: 16.7 k instructions per permutation of which 5.0 k FP64 (two issue slots each) -> 86 % of the "
                             "issue slots used; fmaheavy 67 %, alu 45 %, fp64 34 % busy, dram 1 % (r01, all-integer: fmaheavy 86 %)"},
        "hbm": {"achieved": leaf_bytes / t_leaf / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": leaf_bytes / t_leaf / 1e9 / hbm_peak,
                "algorithmic_bytes": leaf_bytes, "peak_src": peak_src},
    }


def run_next_rows(g, L, check, vp, npp, torch, time_phase, reps, n_bits, cols, blow, ext_bits, seed, dst, hbm_peak):
    """Device-resident times of the rows next to the commit at the workload's size; GB/s are ALGORITHMIC bytes
     over the measured time."""
    from pil2_stark_js_b200._lib import EvalDesc
    q_dim, q_deg = 3, 1 << blow
    q_ext = g.dev(q_dim << ext_bits)
    cmq = g.dev((q_dim * q_deg) << ext_bits)
    q_nodes = g.dev(g.nnodes(1 << ext_bits))
    check(L.pil2gpu_synth_dev(g.h, g.ptr(q_ext), q_dim << ext_bits, seed + 9, 0))

    def phase_q():
        check(L.pil2gpu_compute_q_dev(g.h, g.ptr(q_ext), q_dim, q_deg, n_bits, ext_bits, g.ptr(cmq)))
        check(L.pil2gpu_merkelize_dev(g.h, g.ptr(cmq), q_dim * q_deg, 1 << ext_bits, 0, g.ptr(q_nodes)))
    openings = [0, 1]
    xi = np.ascontiguousarray(splitmix_field(seed + 10, 0, 3))
    lev = g.dev(len(openings) * (3 << n_bits))

    def phase_lev():
        for i, o in enumerate(openings):
            check(L.pil2gpu_compute_lev_dev(g.h, npp(xi), o, n_bits, g.ptr(lev, i * (3 << n_bits))))
    n_ev = 2 * cols
    desc = (EvalDesc * n_ev)(*[EvalDesc(c, 1, o) for o in range(2) for c in range(cols)])
    ev_out = np.empty(3 * n_ev, dtype=np.uint64)

    def phase_evals():
        check(L.pil2gpu_compute_evals_dev(g.h, g.ptr(dst), cols, n_bits, ext_bits, desc, n_ev, g.ptr(lev), len(openings), npp(ev_out)))
    xd = g.dev(3 * len(openings) << ext_bits)
    op_arr = (ctypes.c_int32 * len(openings))(*openings)

    def phase_xdiv():
        check(L.pil2gpu_x_div_x_sub_xi_dev(g.h, npp(xi), op_arr, len(openings), n_bits, ext_bits, g.ptr(xd)))
    from pil2_stark_js_b200._lib import FriTerm
    fterms = (FriTerm * n_ev)(*[FriTerm(dst.data_ptr(), cols, c, 1, o) for o in openings for c in range(cols)])
    f_ext = g.dev(3 << ext_bits)
    vf = [np.ascontiguousarray(splitmix_field(seed + 11 + i, 0, 3)) for i in range(2)]

    def phase_fripol():
        check(L.pil2gpu_fri_pol_dev(g.h, fterms, n_ev, npp(ev_out), op_arr, len(openings), g.ptr(xd), npp(vf[0]), npp(vf[1]), ext_bits,
                                    g.ptr(f_ext)))
    # f4: the quotient-constraint program the reference generated for its sm_all AIR (169 records, tests/golden/sm_all_q_code.json)
    # run over 2^ext_bits rows of synthetic stage buffers by the expression interpreter
    expr = None
    try:
        import types
        from pil2_stark_js_b200 import prover_helpers as PH
        from pil2_stark_js_b200._lib import ExprBuffer
        qc = json.load(open(os.path.join(ROOT, "tests", "golden", "sm_all_q_code.json")))
        info = qc["starkInfo"]
        evm = info["evMap"]
        code = []
        for c in qc["qVerifier"]["code"]:
            srcs = [({"type": evm[s_["id"]]["type"], "id": evm[s_["id"]]["id"], "prime": evm[s_["id"]]["prime"]} if s_["type"] == "eval" else dict(s_))
                    for s_ in c["src"]]
            code.append({"op": c["op"], "dest": dict(c["dest"]), "src": srcs})
        code[-1]["dest"] = {"type": "q", "id": 0, "dim": 3}
        pctx = types.SimpleNamespace(pilInfo=info, nBits=n_bits, nBitsExt=ext_bits, publics=[1, 2, 3],
                                     challenges=[[], *[[[int(x) for x in splitmix_field(seed + 20 + k, 0, 3)] for _ in range(m)] for k, m in enumerate((2, 3, 1, 1))]])
        cc = PH.compile_code(pctx, code, "ext")
        ebufs, keep, words = (ExprBuffer * len(cc.buffers))(), [], 0
        for bi, (name, rw) in enumerate(cc.buffers):
            t = g.dev(rw << ext_bits)
            check(L.pil2gpu_synth_dev(g.h, g.ptr(t), rw << ext_bits, seed + 30 + bi, 0))
            keep.append(t)
            ebufs[bi] = ExprBuffer(t.data_ptr(), rw)
            words += rw << ext_bits

        def phase_expr():
            check(L.pil2gpu_calculate_exps_dev(g.h, vp(cc.ops.ctypes.data), cc.ops.size // 16, vp(cc.consts.ctypes.data), cc.consts.size // 3, ebufs, len(cc.buffers),
                                               ext_bits, 1))
        phase_expr()
        t_ex = time_phase(phase_expr, reps)
        expr = {"s": t_ex, "shape": f"{len(code)}-record quotient program of the sm_all AIR ({cc.ops.size // 16} records after dropping dead temporaries; 51 columns in 4 buffers, {cc.n_slots} live temporaries) over 2^{ext_bits} rows",
                "algorithmic_bytes": 8 * words, "GBps": 8 * words / t_ex / 1e9, "frac_hbm": 8 * words / t_ex / 1e9 / hbm_peak,
                "records_per_s": len(code) * float(1 << ext_bits) / t_ex}
        del keep
    except Exception as ex:       # reported next to the other rows, never instead of them
        expr = {"error": str(ex)[:200]}
    for f in (phase_q, phase_lev, phase_evals, phase_xdiv, phase_fripol):
        f()
    t_q, t_lev, t_ev, t_xd, t_fp = (time_phase(f, reps) for f in (phase_q, phase_lev, phase_evals, phase_xdiv, phase_fripol))
    Ew, Nw = 1 << ext_bits, 1 << n_bits
    q_bytes = 8 * Ew * q_dim * (1 + q_deg) + 8 * Ew * q_dim * q_deg + 64 * Ew
    ev_bytes = 8 * Nw * cols + 24 * Nw * len(openings)
    xd_bytes = 24 * Ew * len(openings)
    fp_bytes = 8 * Ew * cols + 24 * Ew * len(openings) + 24 * Ew
    extras = {
        "fri_pol": {"s": t_fp, "shape": f"{n_ev} evMap terms over the 2^{ext_bits} rows of the {cols}-column extended buffer (friExp -> f_ext)",
                    "algorithmic_bytes": fp_bytes, "GBps": fp_bytes / t_fp / 1e9, "frac_hbm": fp_bytes / t_fp / 1e9 / hbm_peak},
        "q_commit": {"s": t_q, "shape": f"qDim {q_dim}, qDeg {q_deg}, 2^{ext_bits} rows (INTT + split + {q_deg} coset NTTs + merkelize)",
                     "algorithmic_bytes": q_bytes, "GBps": q_bytes / t_q / 1e9, "frac_hbm": q_bytes / t_q / 1e9 / hbm_peak},
        "lev": {"s": t_lev, "shape": f"{len(openings)} openings x 2^{n_bits} F3 (powers + INTT)"},
        "evals": {"s": t_ev, "shape": f"{n_ev} evaluations over the 2^{n_bits} base rows of the {cols}-column extended buffer",
                  "algorithmic_bytes": ev_bytes, "GBps": ev_bytes / t_ev / 1e9, "frac_hbm": ev_bytes / t_ev / 1e9 / hbm_peak,
                  "mulmod_per_s": 3.0 * n_ev * Nw / t_ev},
        "expressions": expr,
        "x_div_x_sub_xi": {"s": t_xd, "shape": f"{len(openings)} openings x 2^{ext_bits} points", "algorithmic_bytes": xd_bytes,
                           "GBps": xd_bytes / t_xd / 1e9, "frac_hbm": xd_bytes / t_xd / 1e9 / hbm_peak},
    }
    del q_ext, cmq, q_nodes, lev, xd, f_ext
    return extras


def finish_line(args, n_bits, cols, blow, sec_per_commit, t_lde, t_mk, t_leaf, t_fri, roofline, roofline_lde, e2e, extras, launches, clocks, root_dev,
                seed, spot=None, fri_out=None):
    cpu = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        sb = pick_sample_bits(n_bits, cols, threads, blow, target_s=15.0)
        d = cpu_commit_sample(n_bits, cols, blow, sb, threads, seed)
        cpu = {"value": d["estimate_full_s"], "unit": "s", "cores": threads, "kind": "port",
               "sample": f"oracle/gl_oracle.c on {threads} host threads: 2^{sb}-row x {cols}-col sample ({d['sample_s']:.1f} s wall), "
                         f"extrapolated to 2^{n_bits} rows (NTT ~ rows*log2 rows, hashing/FRI ~ rows)"}
    line = {
        "metric": METRIC, "value": sec_per_commit, "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_per_commit * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": config_dict(args.workload, 1),
        "rows_per_s": (1 << n_bits) / sec_per_commit, "phases_s": {"lde": t_lde, "merkle": t_mk, "merkle_leaf": t_leaf, "fri": t_fri},
        "roofline": roofline, "roofline_lde": roofline_lde, "cpu_baseline": cpu, "e2e": e2e, "next_rows": extras, "gpu_launches": launches,
        "clocks": clocks, "root": root_dev,
        "parity_spot_check": (spot or {}).get("status", "skipped"), "parity_spot_check_detail": spot,
    }
    line.update(fri_out or {})
    print(json.dumps(line), flush=True)


# dram bytes per launch from the committed `ncu --set full` captures (profiles/), keyed by workload
TRAFFIC = {}
try:
    TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass


def run_e2e(g, args, torch, n_bits, cols, blow, steps, src_dev, fri0_dev, chal, queries, root_dev):
    """The same commit through the reference-facing host-buffer calls, host<->device copies inside the timed region.  Three
    variants, all on BigBuffer page lists (pages of 2^28 u64 = 2 GiB, what a JS BigBuffer holds at this size):
      pinned_fused          e2e.value: extendAndMerkelize as ONE call (pil2gpu_extend_and_merkelize_paged) on pinned pages (the addon's
                            page allocator), FRI folds through pil2gpu_fri_fold -- what js/stark_gen_helpers.js calls;
      pageable_fused        the same call on ordinary pageable pages (plain BigUint64Arrays): staged through the pinned ring;
      pageable_module_swap  only fft_p / merklehash_p swapped: interpolate (pil2gpu_lde_paged) then merkelize
                            (pil2gpu_merkelize_paged) on pageable pages -- the extended buffer crosses PCIe three times."""
    from pil2_stark_js_b200.bigbuffer import BigBuffer
    L, check, vp = g.L, g.check, g.vp
    ext_bits = n_bits + blow
    page = 1 << 28

    def pinned(words):
        p = vp()
        check(L.pil2gpu_host_alloc(int(words) * 8, ctypes.byref(p)))
        return p, np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(int(words),))

    sw, dw, nw = cols << n_bits, cols << ext_bits, g.nnodes(1 << ext_bits)
    bufs = []
    b_src = BigBuffer(sw, page_len=page, pinned=True, zero=False)
    b_dst = BigBuffer(dw, page_len=page, pinned=True, zero=False)
    hp_nodes, h_nodes = pinned(nw); bufs.append(hp_nodes)
    hp_pol, h_pol = pinned(3 << steps[0]); bufs.append(hp_pol)
    off = 0
    for pg in b_src.buffers:                       # synthetic trace: device -> the pinned pages (outside the timed region)
        check(L.pil2gpu_d2h(g.h, vp(pg.ctypes.data), g.ptr(src_dev, off), pg.size * 8))
        off += pg.size
    check(L.pil2gpu_d2h(g.h, hp_pol, g.ptr(fri0_dev), (3 << steps[0]) * 8))
    check(L.pil2gpu_sync(g.h))
    layer = []
    for s in range(len(steps)):
        pp, pa = pinned(3 << steps[s]); bufs.append(pp)
        if s + 1 < len(steps):
            rp, ra = pinned(3 << steps[s]); bufs.append(rp)
            np_, na = pinned(g.nnodes(1 << steps[s + 1])); bufs.append(np_)
        else:
            rp = np_ = None
        layer.append((pp, rp, np_))
    root = np.zeros(4, dtype=np.uint64)
    npp = lambda a: vp(a.ctypes.data)
    h2d = sw * 8
    d2h = (dw + nw) * 8 + 32
    fri_h2d = fri_d2h = 0
    for s in range(len(steps)):
        prev = steps[s - 1] if s else steps[0]
        fri_h2d += (3 << prev) * 8
        fri_d2h += (3 << steps[s]) * 8
        if s + 1 < len(steps):
            fri_d2h += ((3 << steps[s]) + g.nnodes(1 << steps[s + 1])) * 8

    def fri_chain():
        cur = hp_pol
        for s in range(len(steps)):
            last = s == len(steps) - 1
            prev = steps[s - 1] if s else steps[0]
            pp, rp, np_ = layer[s]
            check(L.pil2gpu_fri_fold(g.h, cur, prev, steps[s], -1 if last else steps[s + 1], steps[0], npp(chal[s]), 0, pp, rp, np_))
            cur = pp
        # query openings are host-side gathers on the downloaded trees in the drop-in JS path (merklehash_p.js:142-168)

    def fused(bs, bd, nodes_ptr):
        sp, spw, sn = bs.pages()
        dp, dpw, dn = bd.pages()
        check(L.pil2gpu_extend_and_merkelize_paged(g.h, sp, spw, sn, cols, n_bits, ext_bits, 0, dp, dpw, dn, nodes_ptr, npp(root)))

    def module_swap(bs, bd, nodes_ptr):
        sp, spw, sn = bs.pages()
        dp, dpw, dn = bd.pages()
        check(L.pil2gpu_lde_paged(g.h, sp, spw, sn, dp, dpw, dn, cols, n_bits, ext_bits))                       # fft_p.interpolate
        check(L.pil2gpu_merkelize_paged(g.h, dp, dpw, dn, cols, 1 << ext_bits, 0, nodes_ptr))                   # merklehash_p.merkelize
        root[:] = np.ctypeslib.as_array(ctypes.cast(nodes_ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(nw,))[-4:]

    def timed(fn, n):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n

    n = max(1, min(args.steps, 3))
    t = timed(lambda: (fused(b_src, b_dst, hp_nodes), fri_chain()), n)
    ok = [int(x) for x in root] == root_dev
    variants = {"pinned_fused": {"value": t, "h2d_bytes_per_step": int(h2d + fri_h2d), "d2h_bytes_per_step": int(d2h + fri_d2h),
                                 "root_matches_device_run": ok}}
    if not getattr(args, "no_e2e_variants", False):
        try:
            # ordinary (pageable) pages with the same cut; filled from the pinned ones; the warm-up call of timed() faults them in
            p_src = BigBuffer(sw, page_len=page, zero=False)
            for a, b in zip(p_src.buffers, b_src.buffers):
                a[:] = b
            p_dst = BigBuffer(dw, page_len=page, zero=False)
            p_nodes = np.empty(nw, dtype=np.uint64)
            t_pf = timed(lambda: (fused(p_src, p_dst, npp(p_nodes)), fri_chain()), max(1, n - 1))
            ok_pf = [int(x) for x in root] == root_dev and all(np.array_equal(a, b) for a, b in zip(p_dst.buffers[:2], b_dst.buffers[:2]))
            t_ms = timed(lambda: (module_swap(p_src, p_dst, npp(p_nodes)), fri_chain()), 1)
            ok_ms = [int(x) for x in root] == root_dev
            variants["pageable_fused"] = {"value": t_pf, "h2d_bytes_per_step": int(h2d + fri_h2d), "d2h_bytes_per_step": int(d2h + fri_d2h),
                                          "root_matches_device_run": ok_pf}
            variants["pageable_module_swap"] = {"value": t_ms, "h2d_bytes_per_step": int(h2d + dw * 8 + fri_h2d),
                                                "d2h_bytes_per_step": int(d2h + fri_d2h), "root_matches_device_run": ok_ms}
            del p_src, p_dst, p_nodes
        except (MemoryError, RuntimeError) as ex:
            variants["pageable_error"] = str(ex)[:200]
    for p in bufs:
        L.pil2gpu_host_free(p)
    b_src.free(); b_dst.free()
    return {"value": t, "unit": "s", "h2d_bytes_per_step": int(h2d + fri_h2d), "d2h_bytes_per_step": int(d2h + fri_d2h), "steps": n,
            "root_matches_device_run": ok, "variants": variants,
            "call": "pil2gpu_extend_and_merkelize_paged (BigBuffer pages of 2^28 words: host src -> host dst + nodes, one call) + pil2gpu_fri_fold "
                    "per step; `value` = pinned pages; variants: the same call on pageable pages, and interpolate + merkelize as two calls"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle spot check of the timed buffers")
    ap.add_argument("--no-e2e-variants", action="store_true", help="e2e on pinned pages only (skip the pageable-page variants)")
    args = ap.parse_args()
    if args.impl == "ours" and args.warmup < 3:
        args.warmup = 3                     # timing rule: at least 3 untimed warm-up steps; the JSON line reports what was run
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
