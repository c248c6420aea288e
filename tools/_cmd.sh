# scratch command file for one-off gpurun calls:  gpurun -- 'bash tools/_cmd.sh > gpurun_out/x.log 2>&1; cat gpurun_out/x.log'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
