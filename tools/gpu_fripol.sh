#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python tools/fripol_probe.py 23 256 2 2>&1 | tail -2
python tools/evals_probe.py 22 256 2 2>&1 | tail -1
