cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in tcj1h tcj2h tcj8h; do
  PIL2GPU_LIB=pil2_stark_js_b200/libpil2gpu_$v.so timeout 300 python tools/hash_probe.py 22 256 2>&1 | tail -4
done
