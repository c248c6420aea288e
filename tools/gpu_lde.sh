#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lde or ntt or interpolate or fft or fri" > gpurun_out/lde_tests.log 2>&1; tail -2 gpurun_out/lde_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-extras > gpurun_out/lde_bench.json 2> gpurun_out/lde_bench.err
grep '^{' gpurun_out/lde_bench.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["phases_s"])'
