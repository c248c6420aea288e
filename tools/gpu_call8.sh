#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NG=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/c8_bench_${NG}.json 2> gpurun_out/c8_bench_${NG}.err
echo "bench exit $?"
tail -5 gpurun_out/c8_bench_${NG}.err
grep '^{' gpurun_out/c8_bench_${NG}.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["value"], d["e2e"]["value"], d["exchange"], d["root"], d["gpu_launches"])'
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus $NG --steps 3 --warmup 3 --workload cfg5 > gpurun_out/c8_bench_cfg5_${NG}.json 2> gpurun_out/c8_bench_cfg5_${NG}.err
echo "bench cfg5 exit $?"
tail -5 gpurun_out/c8_bench_cfg5_${NG}.err
grep '^{' gpurun_out/c8_bench_cfg5_${NG}.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["value"], d["e2e"]["value"], d["exchange"], d["root"], d["gpu_launches"])'
