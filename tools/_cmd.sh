cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_f_rows.py tests/test_gpu_expressions.py tests/test_gpu_parity.py -x -q -m gpu -k "evals or fri_pol or stage_flow or golden or sharded" 2>&1 | tail -3
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/nq_bench.json 2> gpurun_out/nq_bench.err; tail -2 gpurun_out/nq_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/nq_bench.json") if l.startswith("{")][-1])
print(d["value"], d["phases_s"])
for k,v in d["next_rows"].items(): print(k, v.get("s"), v.get("frac_hbm"))
PY
