// Memory-pattern probe: copy kernel with the NTT pass access pattern (tiles of 2^t rows, 2^lo rows apart, W columns of 8 B).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
typedef unsigned long long u64;
__global__ void __launch_bounds__(256) k_copy(const ulonglong2* __restrict__ in, ulonglong2* __restrict__ out, u64 C2, int lo, int t, int W2) {
    // C2, W2 in 16-byte units
    const unsigned tile = blockIdx.x;
    const unsigned base_lo = tile & ((1u << lo) - 1), base_hi = tile >> lo;
    const u64 pos0 = ((u64)base_hi << (lo + t)) | base_lo;
    const u64 c0 = (u64)blockIdx.y * W2;
    const int cp = threadIdx.x % W2, kstep = 256 / W2;
    for (int k = threadIdx.x / W2; k < (1 << t); k += kstep) {
        const u64 row = pos0 | ((u64)k << lo);
        out[row * C2 + c0 + cp] = in[row * C2 + c0 + cp];
    }
}
int main() {
    const int n = 22, C = 256;                 // 2^22 rows x 256 cols x 8 B = 8 GiB per buffer
    const size_t bytes = ((size_t)C << n) * 8;
    ulonglong2 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1)); printf("cudaMemcpy D2D: %.0f GB/s (r+w)\n", 2.0 * bytes / ms / 1e6);
    const int t = 8;
    for (int W = 16; W <= 256; W *= 2) {
        for (int lo : {0, 7, 14}) {
            dim3 grid(1u << (n - t), C / W, 1);
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0)); k_copy<<<grid, 256>>>(a, b, C / 2, lo, t, W / 2); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            printf("W=%3d cols (%4d B segments), lo=%2d: %.0f GB/s (r+w)\n", W, W * 8, lo, 2.0 * bytes / ms / 1e6);
        }
    }
    return 0;
}
