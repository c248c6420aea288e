#!/bin/bash
# scratch job for gpurun: A/B runs of bench.py under different environments / library builds
#   ENVS="A=1 A=2"   one quick bench per setting;   LIBS="x.so y.so"  one per library build;   PROBES="gl_probe" binaries under tools/bin
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
Q="--no-cpu --no-extras --no-e2e --no-verify"
for p in $PROBES; do echo "== probe $p"; tools/bin/$p | grep "^poseidon\|mismatch\|^pipe\|^butterfly" | grep -v " 0 mismatching"; done
for l in $LIBS; do
  echo "== $l"
  PIL2GPU_LIB=$PWD/pil2_stark_js_b200/$l TAG=${TAG:-ab}_$l STEPS=3 BENCH_ARGS="$Q" bash tools/gpu.sh bench | grep "^value" | sed 's/e2e None.*//'
done
i=0
for e in $ENVS; do
  echo "== $e"; i=$((i+1))
  env $e TAG=${TAG:-ab}_env$i STEPS=3 BENCH_ARGS="$Q" bash tools/gpu.sh bench | grep "^value" | sed 's/e2e None.*//'
done
