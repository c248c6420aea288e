#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" > gpurun_out/tests2.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests2.log; tail -3 gpurun_out/tests2.log
NG=${NG:-2} bash tools/gpu_bench_multi.sh
