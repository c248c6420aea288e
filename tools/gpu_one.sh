#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "${K:-paged}" > gpurun_out/one.log 2>&1; echo "exit $?"; tail -15 gpurun_out/one.log
