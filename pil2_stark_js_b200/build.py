"""Build libpil2gpu.so in-tree with nvcc for sm_100a.  The .so is git-ignored (history stays source-only) but NOT
gpurun-ignored, so a build made here travels to the GPU box with the working-tree snapshot; a fresh checkout has no .so and
`_lib.load()` builds it on first use when nvcc is present (and raises otherwise: there is no CPU fallback)."""
import os
import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libpil2gpu.so"
SOURCES = ["pil2gpu.cu"]
HEADERS = ["gl.cuh", "poseidon.cuh", "poseidon_rc.inc", "poseidon_rc_limbs.inc", "poseidon_rc_limbs_partial.inc", "poseidon_rc_f64p.inc", "poseidon_tc.cuh", "poseidon_tc_consts.inc", "ntt.cuh", "merkle.cuh", "fri.cuh", "qpath.cuh", "evals.cuh", "fripol.cuh", "expr.cuh", "shard.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-cudart", "static",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libpil2gpu.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [PKG.parent / "include" / "pil2gpu.h", pathlib.Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: an A/B build of the same sources under another name (e.g. defines=["POSEIDON_F64P"], out=PKG / "libpil2gpu_f64p.so")."""
    if out is None and not force and not needs_build():
        return LIB
    out = pathlib.Path(out) if out else LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(out)] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpil2gpu.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
