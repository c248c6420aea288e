#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f_rows.py -m gpu -x -q > gpurun_out/rows_tests.log 2>&1; tail -3 gpurun_out/rows_tests.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/rows_bench.json 2> gpurun_out/rows_bench.err
grep '^{' gpurun_out/rows_bench.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"]); print(json.dumps(d["next_rows"], indent=1))'
