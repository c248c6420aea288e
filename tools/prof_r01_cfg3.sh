#!/bin/bash
# ncu evidence for the round: (1) the bench command without the profiler, (2) its launch list, (3) --set full of the LDE kernels and
# the leaf hash.  Summaries are built here by tools/make_profiles.py from the files this leaves in gpurun_out/.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
B="python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras"
$B > gpurun_out/r01_plain_cfg3.log 2>&1 || { tail -5 gpurun_out/r01_plain_cfg3.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_cfg3.csv $B > gpurun_out/r01_ncu_launch_cfg3.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|ntt_lde|merkle_leaf' -c 6 -f -o gpurun_out/r01_prof_cfg3 $B > gpurun_out/r01_ncu_full_cfg3.log 2>&1
tail -2 gpurun_out/r01_ncu_full_cfg3.log
ls -la gpurun_out/r01_*
