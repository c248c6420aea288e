#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref exit $?"
grep '^{' gpurun_out/final_bench.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["phases_s"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"]); print({k:v["s"] for k,v in d["next_rows"].items()})'
grep '^{' gpurun_out/final_ref.json | cut -c1-300
