python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_plain_cfg3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_ncu_launch_cfg3.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|ntt_lde|merkle_leaf' -c 6 -f -o gpurun_out/r01_prof_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_ncu_full_cfg3.log 2>&1
tail -2 gpurun_out/r01_ncu_full_cfg3.log
