// PCIe probe: cudaMemcpy2DAsync throughput for column slabs of a row-major pinned host buffer (pitch 2048 B).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
int main() {
    const size_t rows = 1 << 22, pitch = 2048;   // 8 GiB host buffer
    char *h, *d;
    CK(cudaHostAlloc(&h, rows * pitch, cudaHostAllocPortable));
    CK(cudaMalloc(&d, rows * pitch));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(d, h, rows * pitch, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("H2D contiguous 8 GiB: %.1f GB/s\n", rows * pitch / ms / 1e6);
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(h, d, rows * pitch, cudaMemcpyDeviceToHost, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("D2H contiguous 8 GiB: %.1f GB/s\n", rows * pitch / ms / 1e6);
    }
    for (size_t w = 128; w <= 2048; w *= 2) {
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpy2DAsync(d, w, h, pitch, w, rows, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("H2D 2D width %4zu B (host pitch 2048): %.1f GB/s\n", w, rows * w / ms / 1e6);
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpy2DAsync(h, pitch, d, w, w, rows, cudaMemcpyDeviceToHost, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("D2H 2D width %4zu B (host pitch 2048): %.1f GB/s\n", w, rows * w / ms / 1e6);
    }
    // both directions at once (contiguous halves)
    CK(cudaEventRecord(e0, s1));
    CK(cudaMemcpyAsync(d, h, rows * pitch / 2, cudaMemcpyHostToDevice, s1));
    CK(cudaMemcpyAsync(h + rows * pitch / 2, d + rows * pitch / 2, rows * pitch / 2, cudaMemcpyDeviceToHost, s2));
    CK(cudaStreamSynchronize(s2)); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); printf("bidirectional 4+4 GiB: %.1f GB/s aggregate\n", rows * pitch / ms / 1e6);
    return 0;
}
