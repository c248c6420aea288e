#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f_rows.py -m gpu -x -q -k "two_gpus or sharded_rows" > gpurun_out/tests2.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests2.log; tail -25 gpurun_out/tests2.log
