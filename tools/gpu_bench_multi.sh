#!/bin/bash
# Multi-GPU bench in both exchange modes:   gpurun --gpus N -- 'NG=N bash tools/gpu_bench_multi.sh'
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
NG=${NG:-2}
for mode in ${MODES:-peer nccl}; do
  if [ -n "$TRACE" ]; then export PIL2GPU_TRACE=1; fi
  PIL2GPU_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
     --master-port 29511 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/multi_${NG}_$mode.json 2> gpurun_out/multi_${NG}_$mode.err
  echo "bench $mode exit $?"; tail -3 gpurun_out/multi_${NG}_$mode.err
  grep '^{' gpurun_out/multi_${NG}_$mode.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["value"], d["e2e"]["value"], d["exchange"], d["root"])'
done
