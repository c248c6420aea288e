#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python tools/fripol_probe.py 23 256 2 > gpurun_out/fripol_probe.log 2>&1; cat gpurun_out/fripol_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fripol' -c 2 -f -o gpurun_out/r01_prof_fripol python tools/fripol_probe.py 23 256 2 > gpurun_out/fripol_ncu.log 2>&1
tail -2 gpurun_out/fripol_ncu.log
