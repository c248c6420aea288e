#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c4_pytest.log
tail -5 gpurun_out/c4_pytest.log
PIL2GPU_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c4_bench1.json 2> gpurun_out/c4_bench1.err
echo "bench exit $?"
grep '^{' gpurun_out/c4_bench1.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]); print(json.dumps(d["next_rows"], indent=1))'
tail -8 gpurun_out/c4_bench1.err
