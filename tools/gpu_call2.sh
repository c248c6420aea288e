#!/bin/bash
# 2-GPU check: sharded commit test in both exchange modes, then the bench at N=2 with both exchanges.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c2_smi.log 2>&1
nvidia-smi topo -m >> gpurun_out/c2_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k two_gpus > gpurun_out/c2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c2_pytest.log
tail -15 gpurun_out/c2_pytest.log
for mode in peer nccl; do
  PIL2GPU_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c2_bench_$mode.json 2> gpurun_out/c2_bench_$mode.err
  echo "bench $mode exit $?"
  tail -3 gpurun_out/c2_bench_$mode.err
  cut -c1-900 gpurun_out/c2_bench_$mode.json
done
