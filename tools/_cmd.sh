# scratch command file for one-off gpurun calls:  gpurun -- 'bash tools/_cmd.sh > gpurun_out/x.log 2>&1; cat gpurun_out/x.log'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'expr_jit|expr_kernel' -c 1 -f -o gpurun_out/r02_prof_expr python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify 2>&1 | tail -2
PIL2GPU_EXPR=interp timeout 900 ncu --set full --clock-control none -k regex:'expr_jit|expr_kernel' -c 1 -f -o gpurun_out/r02_prof_expr_interp python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify 2>&1 | tail -2
