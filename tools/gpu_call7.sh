#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NG=2
PIL2GPU_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/c7_bench_${NG}.json 2> gpurun_out/c7_bench_${NG}.err
grep "pil2gpu" gpurun_out/c7_bench_${NG}.err | tail -8
grep '^{' gpurun_out/c7_bench_${NG}.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["value"], d["e2e"]["value"])'
numactl -H 2>/dev/null | head -5; nvidia-smi topo -m | head -8
