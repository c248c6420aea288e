cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for m in 1 2; do
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTC_MODE=$m -I pil2_stark_js_b200/csrc -o /tmp/ntt_tc_probe$m tools/probe/ntt_tc_probe.cu && echo "mode $m" && timeout 120 /tmp/ntt_tc_probe$m 200
done
