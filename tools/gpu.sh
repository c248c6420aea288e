#!/bin/bash
# One script for every GPU-box job (run through gpurun; outputs under gpurun_out/<tag>_*):
#   gpurun --timeout 1500 -- 'bash tools/gpu.sh tests bench'            # pytest -m gpu, then the default bench line
#   gpurun --timeout 1500 -- 'TAG=r02a bash tools/gpu.sh bench ncu'     # bench, launch list + ncu --set full of the hot kernels
#   gpurun --gpus 8 --timeout 900 -- 'NG=8 bash tools/gpu.sh multi'     # sharded bench under torchrun
# jobs: tests | bench | ref | ncu | multi | probe | cmd (runs "$CMD")
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${TAG:-r02}; NG=${NG:-1}; WL=${WL:-cfg3}
for job in "$@"; do
  case $job in
    tests)
      timeout ${T_TESTS:-1200} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/${TAG}_tests.log 2>&1
      echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log; tail -5 gpurun_out/${TAG}_tests.log;;
    bench)
      timeout 900 python bench.py --workload $WL --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/${TAG}_bench_$WL.json 2> gpurun_out/${TAG}_bench_$WL.err
      echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench_$WL.err
      python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_$WL.json") if l.startswith("{")][-1])
    print("value", d["value"], "phases", d.get("phases_s"), "e2e", (d.get("e2e") or {}).get("value"), "spot", d.get("parity_spot_check"),
          "roofline", d["roofline"]["bound"], d["roofline"]["frac"], "lde", d["roofline_lde"]["frac"], "clocks", d["clocks"])
except Exception as ex:
    print("no bench line:", ex)
PY
      ;;
    ref)
      timeout 900 python bench.py --impl reference --workload $WL --steps ${REF_STEPS:-3} --warmup 1 > gpurun_out/${TAG}_ref_$WL.json 2> gpurun_out/${TAG}_ref_$WL.err
      echo "ref exit $?"; tail -c 600 gpurun_out/${TAG}_ref_$WL.json;;
    ncu)
      B="python bench.py --workload $WL --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras --no-verify"
      $B > gpurun_out/${TAG}_plain_$WL.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain_$WL.log; continue; }
      ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_$WL.csv $B > gpurun_out/${TAG}_ncu_launch_$WL.log 2>&1
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-ntt_pass|ntt_lde|merkle_leaf|merkle_levels}" -c ${NCU_C:-8} -f \
          -o gpurun_out/${TAG}_prof_$WL $B > gpurun_out/${TAG}_ncu_full_$WL.log 2>&1
      tail -2 gpurun_out/${TAG}_ncu_full_$WL.log; ls -la gpurun_out/${TAG}_*;;
    multi)
      for mode in ${MODES:-peer}; do
        [ -n "$TRACE" ] && export PIL2GPU_TRACE=1
        PIL2GPU_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
           --master-port 29511 bench.py --gpus $NG --workload $WL --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} \
           > gpurun_out/${TAG}_multi_${NG}_${WL}_$mode.json 2> gpurun_out/${TAG}_multi_${NG}_${WL}_$mode.err
        echo "multi $NG $mode exit $?"; tail -3 gpurun_out/${TAG}_multi_${NG}_${WL}_$mode.err
        grep '^{' gpurun_out/${TAG}_multi_${NG}_${WL}_$mode.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["value"], (d.get("e2e") or {}).get("value"), d.get("phases_s"), d["root"])'
      done;;
    probe)
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I pil2_stark_js_b200/csrc -o /tmp/gl_probe tools/probe/gl_probe.cu \
        && /tmp/gl_probe ${PROBE_ARGS} > gpurun_out/${TAG}_probe.log 2>&1; tail -30 gpurun_out/${TAG}_probe.log;;
    cmd)
      bash -c "$CMD" > gpurun_out/${TAG}_cmd.log 2>&1; echo "cmd exit $?"; tail -30 gpurun_out/${TAG}_cmd.log;;
  esac
done
