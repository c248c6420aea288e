#!/bin/bash
# 1-GPU check: full GPU test-suite, then the headline bench.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench exit $?"
cat gpurun_out/c1_bench.json | cut -c1-1500
