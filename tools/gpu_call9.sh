#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NG=8
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc; free -g | head -2) > gpurun_out/c9_topo.log 2>&1
for bind in 1 0; do
  if [ $bind = 0 ]; then export PIL2GPU_NO_BIND=1; fi
  PIL2GPU_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2951$bind \
     bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/c9_bench_bind$bind.json 2> gpurun_out/c9_bench_bind$bind.err
  echo "bench bind=$bind exit $?"
  grep '^{' gpurun_out/c9_bench_bind$bind.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["value"], d["e2e"]["value"])'
  grep "pil2gpu\|bench\]" gpurun_out/c9_bench_bind$bind.err | tail -16
done
cat gpurun_out/c9_topo.log
