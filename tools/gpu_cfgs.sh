#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for w in cfg2 cfg5; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/cfg_$w.json 2> gpurun_out/cfg_$w.err
  echo "$w exit $?"; tail -2 gpurun_out/cfg_$w.err
  grep '^{' gpurun_out/cfg_$w.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["config"]["workload"]); print(d["value"], d["phases_s"], d["e2e"]["value"], d["roofline_lde"]["frac"], d["roofline"]["int_pipes"]["perms_per_s"]); print(d["next_rows"]["q_commit"]["s"], d["next_rows"]["evals"]["s"])'
done
