cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for c in 8 37 64 128; do
echo "cols $c new:"; timeout 200 python tools/evals_probe.py 22 $c 2 2>&1 | tail -1
echo "cols $c old:"; PIL2GPU_EVALS=mma1 timeout 200 python tools/evals_probe.py 22 $c 2 2>&1 | tail -1
done
