// Stand-alone probe: Poseidon-GL throughput + integer-pipe microbenchmarks on one GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I pil2_stark_js_b200/csrc tools/poseidon_probe.cu -o gpurun_out/poseidon_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "poseidon.cuh"

__global__ void __launch_bounds__(256) k_perm_chain(u64* out, int iters) {
    u64 x[12];
    u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = tid * 12 + i;
    for (int it = 0; it < iters; it++) poseidon_permute(x);
#pragma unroll
    for (int i = 0; i < 12; i++) out[tid * 12 + i] = gl_canon(x[i]);
}

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(u32* out, int iters) {
    u32 a = threadIdx.x + 1, b = blockIdx.x + 3;
    u64 acc0 = a, acc1 = b, acc2 = a ^ b, acc3 = a + b;
    u32 c0 = a, c1 = b, c2 = a * 3, c3 = b * 5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) {  // IMAD.WIDE.U32 with 64-bit accumulate, 4 independent chains
                acc0 += (u64)(u32)acc0 * 0x9E3779B9u; acc1 += (u64)(u32)acc1 * 0x85EBCA6Bu;
                acc2 += (u64)(u32)acc2 * 0xC2B2AE35u; acc3 += (u64)(u32)acc3 * 0x27D4EB2Fu;
            } else if (MODE == 1) {  // 32-bit IMAD
                c0 = c0 * 0x9E3779B9u + c1; c1 = c1 * 0x85EBCA6Bu + c2; c2 = c2 * 0xC2B2AE35u + c3; c3 = c3 * 0x27D4EB2Fu + c0;
            } else if (MODE == 2) {  // IADD3 / LOP3 mix on the alu pipe
                c0 = (c0 + c1) ^ c2; c1 = (c1 + c2) ^ c3; c2 = (c2 + c3) ^ c0; c3 = (c3 + c0) ^ c1;
            } else {  // full 64x64 -> 128 + reduce (gl_mul), 4 chains
                acc0 = gl_mul(acc0, acc1); acc1 = gl_mul(acc1, acc2); acc2 = gl_mul(acc2, acc3); acc3 = gl_mul(acc3, acc0);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)(acc0 ^ acc1 ^ acc2 ^ acc3) ^ c0 ^ c1 ^ c2 ^ c3;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("device %s SMs %d clock %d kHz\n", pr.name, pr.multiProcessorCount, clk);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = pr.multiProcessorCount * 8, threads = 256;
    u64* out; CK(cudaMalloc(&out, (size_t)blocks * threads * 12 * 8));
    // KAT: H(0..11)
    {
        k_perm_chain<<<1, 32>>>(out, 1); CK(cudaDeviceSynchronize());
        u64 h[12]; CK(cudaMemcpy(h, out, 96, cudaMemcpyDeviceToHost));
        printf("perm(0..11)[0..3] = %016llx %016llx %016llx %016llx (expect d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2)\n", h[0], h[1], h[2], h[3]);
    }
    for (int rep = 0; rep < 3; rep++) {
        int iters = 64;
        CK(cudaEventRecord(e0)); k_perm_chain<<<blocks, threads>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double perms = (double)blocks * threads * iters;
        printf("poseidon plain: %.3f ms, %.3f Gperm/s\n", ms, perms / ms / 1e6);
    }
    u32* o32 = (u32*)out;
    const char* names[4] = {"IMAD.WIDE.U32 acc", "IMAD 32", "IADD3/LOP3", "gl_mul"};
    for (int mode = 0; mode < 4; mode++) {
        int iters = 4096; float ms;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(e0));
            if (mode == 0) k_pipe<0><<<blocks, threads>>>(o32, iters);
            if (mode == 1) k_pipe<1><<<blocks, threads>>>(o32, iters);
            if (mode == 2) k_pipe<2><<<blocks, threads>>>(o32, iters);
            if (mode == 3) k_pipe<3><<<blocks, threads>>>(o32, iters);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        double ops = (double)blocks * threads * iters * 16 * 4 * (mode == 2 ? 2 : 1);
        printf("%-20s %.3f ms  %.1f Gop/s  (%.2f ops/clk/SM at %d MHz nominal)\n", names[mode], ms, ops / ms / 1e6,
               ops / (ms * 1e-3) / pr.multiProcessorCount / (clk * 1e3), clk / 1000);
    }
    return 0;
}
