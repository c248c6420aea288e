#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c10_pytest.log
tail -5 gpurun_out/c10_pytest.log
timeout 300 tools/bin/gl_probe > gpurun_out/c10_probe.log 2>&1; grep -E "poseidon|variant" gpurun_out/c10_probe.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c10_bench1.json 2> gpurun_out/c10_bench1.err
echo "bench exit $?"
grep '^{' gpurun_out/c10_bench1.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["phases_s"], d["e2e"]["value"], d["roofline"]["int_pipes"]["perms_per_s"])'
