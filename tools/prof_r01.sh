set -x
python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_plain_cfg3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_ncu_launch_cfg3.log 2>&1
python bench.py --workload cfg2 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_plain_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'ntt_|merkle_leaf|merkle_level|fri_fold' -s 17 -c 14 -f -o gpurun_out/r01_prof_cfg2 python bench.py --workload cfg2 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_ncu_full_cfg2.log 2>&1
tail -2 gpurun_out/r01_ncu_full_cfg2.log
