python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --workload cfg3 --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phases_s'])"
python bench.py --workload cfg2 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_plain_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|ntt_lde' -c 5 -f -o gpurun_out/r01_prof_ntt_cfg2 python bench.py --workload cfg2 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r01_ncu_ntt_cfg2.log 2>&1
tail -1 gpurun_out/r01_ncu_ntt_cfg2.log
