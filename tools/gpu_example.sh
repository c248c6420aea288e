#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f_rows.py -m gpu -x -q -k stage_flow 2>&1 | tail -3
python examples/stage_flow.py 20 128 host 2>&1 | tail -15
python examples/stage_flow.py 20 128 device 2>&1 | tail -15
