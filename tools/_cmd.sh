# scratch command file for one-off gpurun calls:  gpurun -- 'bash tools/_cmd.sh > gpurun_out/x.log 2>&1; cat gpurun_out/x.log'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/nq_bench.json 2> gpurun_out/nq_bench.err; tail -2 gpurun_out/nq_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/nq_bench.json") if l.startswith("{")][-1])
print(d["value"], d["phases_s"])
for k,v in d["next_rows"].items(): print(k, v.get("s"), v.get("frac_hbm"))
PY
