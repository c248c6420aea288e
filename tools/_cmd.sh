#!/bin/bash
# scratch job for gpurun (A/B of library builds): LIBS="a b c" -> bench each
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for p in $PROBES; do echo "== probe $p"; tools/bin/$p | grep "^poseidon\|mismatch" | grep -v " 0 mismatching"; done
for l in $LIBS; do
  export PIL2GPU_LIB=$PWD/pil2_stark_js_b200/$l; echo "== $l"
  TAG=${TAG:-ab}_$l STEPS=3 BENCH_ARGS="--no-cpu --no-extras --no-e2e --no-verify" bash tools/gpu.sh bench | grep "^value" | sed 's/e2e None.*//'
done
