cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'evals_mma2|evals_gather' -c 2 -s 2 -f -o gpurun_out/r02_prof_evals python tools/evals_probe.py 23 256 2 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fripol_mma2|fripol_finish' -c 2 -s 2 -f -o gpurun_out/r02_prof_fripol python tools/fripol_probe.py 23 256 2 2>&1 | tail -2
