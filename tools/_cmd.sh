B="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --no-verify"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline_sizes.py tests/test_gpu_f_rows.py -x -q -k "ntt or lde or fft or interpolate or extend or cfg2 or compute_q or scatter or lev" 2>&1 | tail -4
for cfg in "1 0" "0 1" "1 8" "1 32" "0 16"; do set -- $cfg; echo "PIPE=$1 IPC=$2"; PIL2GPU_NTT_PIPE=$1 PIL2GPU_NTT_IPC=$2 $B 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["phases_s"], d["value"], d["root"][0])'; done
