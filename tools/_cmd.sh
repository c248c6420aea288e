# scratch command file for one-off gpurun calls:  gpurun -- 'bash tools/_cmd.sh > gpurun_out/x.log 2>&1; cat gpurun_out/x.log'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for n in 5 4; do
PIL2GPU_EXPR_CTAS=$n timeout 600 python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/nq_bench.json 2> gpurun_out/nq_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/nq_bench.json") if l.startswith("{")][-1])
print("ctas $n expressions", d["next_rows"]["expressions"])
PY
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
