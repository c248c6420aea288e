# scratch command file for one-off gpurun calls:  gpurun -- 'bash tools/_cmd.sh > gpurun_out/x.log 2>&1; cat gpurun_out/x.log'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_expressions.py -x -q -m gpu -k "persist or golden_quotient" 2>&1 | tail -6
