#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python tools/evals_probe.py 22 256 2 > gpurun_out/evals_probe.log 2>&1; cat gpurun_out/evals_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'evals_mma|lev_bytes|evals_gather' -c 3 -f -o gpurun_out/r01_prof_evals python tools/evals_probe.py 22 256 2 > gpurun_out/evals_ncu.log 2>&1
tail -2 gpurun_out/evals_ncu.log
