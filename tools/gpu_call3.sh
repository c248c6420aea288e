#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c3_pytest.log
tail -15 gpurun_out/c3_pytest.log
for mode in peer; do
  PIL2GPU_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c3_bench_$mode.json 2> gpurun_out/c3_bench_$mode.err
  echo "bench $mode exit $?"
  tail -3 gpurun_out/c3_bench_$mode.err
  grep '^{' gpurun_out/c3_bench_$mode.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"], d["exchange"], d["root"])'
done
PIL2GPU_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/c3_bench1.json 2> gpurun_out/c3_bench1.err
grep '^{' gpurun_out/c3_bench1.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"])'
tail -20 gpurun_out/c3_bench1.err
