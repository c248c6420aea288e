python bench.py --workload slab --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_plain_slab.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|ntt_lde' -c 5 -f -o gpurun_out/r01_prof_ntt_slab python bench.py --workload slab --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_ncu_ntt_slab.log 2>&1
tail -1 gpurun_out/r01_ncu_ntt_slab.log; tail -1 gpurun_out/r01_plain_slab.log | cut -c1-600
