cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f_rows.py tests/test_gpu_expressions.py -x -q -m gpu 2>&1 | tail -3
echo "--- fripol new"; timeout 200 python tools/fripol_probe.py 23 256 2 2>&1 | tail -2
echo "--- fripol old"; PIL2GPU_FRIPOL=mma1 timeout 200 python tools/fripol_probe.py 23 256 2 2>&1 | tail -2
