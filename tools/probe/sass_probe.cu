// Minimal kernel for offline SASS inspection of the shipped permutation (pipe balance of the full-round loop):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pil2_stark_js_b200/csrc -cubin -o /tmp/p.cubin tools/probe/sass_probe.cu
#include "poseidon.cuh"
extern "C" __global__ void __launch_bounds__(128, 5) k_chain(u64* out, int iters) {
    u64 x[12];
    u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = tid * 12 + i;
    for (int it = 0; it < iters; it++) poseidon_permute_mont(x);
#pragma unroll
    for (int i = 0; i < 12; i++) out[tid * 12 + i] = x[i];
}
