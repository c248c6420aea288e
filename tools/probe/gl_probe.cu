// Stand-alone probe: arithmetic / Poseidon candidates vs the shipped versions (correctness + throughput on one GPU).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I pil2_stark_js_b200/csrc -I tools/probe tools/probe/gl_probe.cu -o tools/bin/gl_probe
// Results of the rounds of experiments are kept in profiles/r01_probe_*.log and summarised in DESIGN.md section 4.0.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "poseidon.cuh"
#include "poseidon_mont.cuh"
#include "poseidon_fp64.cuh"
#include "poseidon_z.cuh"
#include "poseidon_lr.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ u64 splitmix(u64 z) {
    z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
}

// VAR: 0 shipped, 1 64-bit lanes, 2 FP64-pipe MDS, 3 u64-state loop with two-input adds forced to IADD3, 4 previous shipped loop (u64 state)
template <int VAR>
__device__ __forceinline__ void permute_mont_var(u64 x[12]) {
    if (VAR == 0) poseidon_permute_mont_limb(x);
    if (VAR == 1) poseidon_permute_lanes64(x);
    if (VAR == 2) poseidon_permute_fp64(x);
    if (VAR == 3) poseidon_permute_mont_z(x);
    if (VAR == 4) poseidon_permute_mont_u64state(x);
    if (VAR == 5) poseidon_permute_mont_f64p<22>(x);
    if (VAR == 6) poseidon_permute_mont_f64p<6>(x);
    if (VAR == 7) poseidon_permute_mont_f64p<8>(x);
    if (VAR == 8) poseidon_permute_mont_f64p<10>(x);
    if (VAR == 9) poseidon_permute_mont_f64p<12>(x);
    if (VAR == 10) poseidon_permute_mont_f64p<14>(x);
    if (VAR == 11) poseidon_permute_mont_f64p<16>(x);
    if (VAR == 12) poseidon_permute_mont_f64p<18>(x);
}

template <int VAR>
__global__ void __launch_bounds__(256) k_check(u64* out, u64 seed, int lazy) {
    u64 x[12];
    u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 12; i++) { u64 v = seed ? splitmix(seed + tid * 12 + i) : (u64)i; x[i] = lazy ? v : gl_canon(v); }
    if (seed && tid % 7 == 0) { x[3] = GL_P - 1; x[4] = 0; x[5] = 0xFFFFFFFFULL; x[6] = lazy ? 0xFFFFFFFFFFFFFFFFULL : 1; }
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_to_mont(x[i]);
    permute_mont_var<VAR>(x);
#pragma unroll
    for (int i = 0; i < 12; i++) out[tid * 12 + i] = gl_from_mont(x[i]);
}

template <int VAR>
__global__ void __launch_bounds__(256) k_chain(u64* out, int iters) {
    u64 x[12];
    u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = tid * 12 + i;
    for (int it = 0; it < iters; it++) permute_mont_var<VAR>(x);
#pragma unroll
    for (int i = 0; i < 12; i++) out[tid * 12 + i] = x[i];
}

// MODE 0: gl_mul chain, 1: gl_mmul chain, 2: legacy butterfly (gl_mul/gl_add/gl_sub), 3: Montgomery butterfly, 4 DFMA, 5 DADD
template <int MODE>
__global__ void __launch_bounds__(256) k_arith(u64* out, int iters, u64 w) {
    u64 a0 = threadIdx.x + 1, a1 = blockIdx.x + 3, a2 = a0 ^ 0x1234567, a3 = a0 + a1 + 77;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) { a0 = gl_mul(a0, w); a1 = gl_mul(a1, w); a2 = gl_mul(a2, w); a3 = gl_mul(a3, w); }
            if (MODE == 1) { a0 = gl_mmul(a0, w); a1 = gl_mmul(a1, w); a2 = gl_mmul(a2, w); a3 = gl_mmul(a3, w); }
            if (MODE == 2) {
                u64 t = gl_mul(a1, w); a1 = gl_sub(a0, t); a0 = gl_add(a0, t);
                t = gl_mul(a3, w); a3 = gl_sub(a2, t); a2 = gl_add(a2, t);
                t = gl_mul(a2, w); a2 = gl_sub(a0, t); a0 = gl_add(a0, t);
                t = gl_mul(a3, w); a3 = gl_sub(a1, t); a1 = gl_add(a1, t);
            }
            if (MODE == 3) {
                u64 t = gl_mmul(a1, w); a1 = gl_subc(a0, t); a0 = gl_addc(a0, t);
                t = gl_mmul(a3, w); a3 = gl_subc(a2, t); a2 = gl_addc(a2, t);
                t = gl_mmul(a2, w); a2 = gl_subc(a0, t); a0 = gl_addc(a0, t);
                t = gl_mmul(a3, w); a3 = gl_subc(a1, t); a1 = gl_addc(a1, t);
            }
            if (MODE == 4) {
                double d0 = __longlong_as_double(a0), d1 = __longlong_as_double(a1), d2 = __longlong_as_double(a2), d3 = __longlong_as_double(a3);
                d0 = fma(d0, 1.0000001, d1); d1 = fma(d1, 0.9999999, d2); d2 = fma(d2, 1.0000002, d3); d3 = fma(d3, 0.9999998, d0);
                a0 = __double_as_longlong(d0); a1 = __double_as_longlong(d1); a2 = __double_as_longlong(d2); a3 = __double_as_longlong(d3);
            }
            if (MODE == 5) {
                double d0 = __longlong_as_double(a0), d1 = __longlong_as_double(a1), d2 = __longlong_as_double(a2), d3 = __longlong_as_double(a3);
                d0 = d0 + d1; d1 = d1 + d2; d2 = d2 + d3; d3 = d3 + d0;
                a0 = __double_as_longlong(d0); a1 = __double_as_longlong(d1); a2 = __double_as_longlong(d2); a3 = __double_as_longlong(d3);
            }
        }
    }
    out[(u64)blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3;
}


// Pipe concurrency: MODE bits 1 = DFMA chains (FP64 pipe), 2 = IMAD chains (fmaheavy), 4 = IADD3/LOP3 chains (alu); 4 independent
// chains of each selected kind per thread, interleaved instruction by instruction.  If the pipes run side by side the time of a
// mixed mode is the maximum of its parts, if they share an issue port it is their sum.
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(u64* out, int iters, u32 m, double dm) {
    double d0 = threadIdx.x + 1.0, d1 = blockIdx.x + 3.0, d2 = d0 * 0.5, d3 = d1 * 0.25;
    u32 a0 = threadIdx.x + 1, a1 = blockIdx.x + 3, a2 = a0 ^ 0x1234567u, a3 = a0 + a1 + 77;
    u32 b0 = a0 * 3, b1 = a1 * 5, b2 = a2 * 7, b3 = a3 * 9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE & 1) { d0 = fma(d0, dm, d1); d1 = fma(d1, dm, d2); d2 = fma(d2, dm, d3); d3 = fma(d3, dm, d0); }
            if (MODE & 2) { a0 = a0 * m + a1; a1 = a1 * m + a2; a2 = a2 * m + a3; a3 = a3 * m + a0; }
            if (MODE & 4) { b0 = (b0 + b1 + m) ^ b2; b1 = (b1 + b2 + m) ^ b3; b2 = (b2 + b3 + m) ^ b0; b3 = (b3 + b0 + m) ^ b1; }
        }
    }
    out[(u64)blockIdx.x * blockDim.x + threadIdx.x] = (u64)__double_as_longlong(d0 + d1 + d2 + d3) ^ a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
}

__global__ void k_arith_check(const u64* a, const u64* b, u64* o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    o[4 * i] = gl_mmul(a[i], b[i]);
    o[4 * i + 1] = gl_addc(a[i], b[i]);
    o[4 * i + 2] = gl_subc(a[i], b[i]);
    o[4 * i + 3] = gl_mmul(a[i], a[i]);
}
static u64 hmul(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % GL_P); }
static u64 hpow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = hmul(r, a); a = hmul(a, a); e >>= 1; } return r; }

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s SMs %d\n", pr.name, pr.multiProcessorCount);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = pr.multiProcessorCount * 8, threads = 256;
    const size_t n = (size_t)blocks * threads;
    u64 *o0, *o1; CK(cudaMalloc(&o0, n * 12 * 8)); CK(cudaMalloc(&o1, n * 12 * 8));
    {
        const int N = 1 << 16;
        u64 *ha = (u64*)malloc(N * 8), *hb = (u64*)malloc(N * 8), *ho = (u64*)malloc(N * 32);
        u64 s = 12345;
        for (int i = 0; i < N; i++) {
            s = s * 6364136223846793005ULL + 1442695040888963407ULL; ha[i] = s ^ (s >> 29);
            s = s * 6364136223846793005ULL + 1442695040888963407ULL; hb[i] = (s ^ (s >> 31)) % GL_P;
        }
        ha[0] = ~0ULL; hb[0] = GL_P - 1; ha[1] = 0; hb[1] = 0; ha[2] = ~0ULL; hb[2] = 0; ha[3] = GL_P; hb[3] = GL_P - 1; ha[4] = 0xFFFFFFFFULL; hb[4] = GL_P - 1;
        ha[5] = 0; hb[5] = GL_P - 1; ha[6] = 1; hb[6] = GL_P - 1; ha[7] = ~0ULL; hb[7] = 1;
        u64 *da, *db, *dout; CK(cudaMalloc(&da, N * 8)); CK(cudaMalloc(&db, N * 8)); CK(cudaMalloc(&dout, N * 32));
        CK(cudaMemcpy(da, ha, N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, N * 8, cudaMemcpyHostToDevice));
        k_arith_check<<<N / 256, 256>>>(da, db, dout, N); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ho, dout, N * 32, cudaMemcpyDeviceToHost));
        const u64 rinv = hpow((u64)(((unsigned __int128)1 << 64) % GL_P), GL_P - 2);
        long bad = 0, noncanon = 0;
        for (int i = 0; i < N; i++) {
            u64 am = ha[i] % GL_P;
            if (ho[4 * i] % GL_P != hmul(hmul(am, hb[i]), rinv)) bad++;
            if (ho[4 * i] >= GL_P) noncanon++;
            if (ho[4 * i + 1] % GL_P != (u64)(((unsigned __int128)am + hb[i]) % GL_P)) bad++;
            if (ho[4 * i + 2] % GL_P != (u64)(((unsigned __int128)am + GL_P - hb[i]) % GL_P)) bad++;
            if (ho[4 * i + 3] % GL_P != hmul(hmul(am, am), rinv)) bad++;
        }
        printf("arith check: %ld mismatches, %ld non-canonical mmul outputs (of %d)\n", bad, noncanon, N);
    }
    {
        k_check<0><<<1, 32>>>(o0, 0, 0); CK(cudaDeviceSynchronize());
        unsigned long long h[12]; CK(cudaMemcpy(h, o0, 96, cudaMemcpyDeviceToHost));
        printf("shipped perm(0..11)[0..3] = %016llx %016llx %016llx %016llx (expect d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2)\n",
               h[0], h[1], h[2], h[3]);
        u64* h0 = (u64*)malloc(n * 96); u64* h1 = (u64*)malloc(n * 96);
        for (int lazy = 0; lazy < 2; lazy++) {
            k_check<0><<<blocks, threads>>>(o0, 777 + lazy, lazy); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h0, o0, n * 96, cudaMemcpyDeviceToHost));
            for (int var = 1; var <= 12; var++) {
                if (var == 1) k_check<1><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 2) k_check<2><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 3) k_check<3><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 4) k_check<4><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 5) k_check<5><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 6) k_check<6><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 7) k_check<7><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 8) k_check<8><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 9) k_check<9><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 10) k_check<10><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 11) k_check<11><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                if (var == 12) k_check<12><<<blocks, threads>>>(o1, 777 + lazy, lazy);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h1, o1, n * 96, cudaMemcpyDeviceToHost));
                long bad = 0;
                for (size_t i = 0; i < n * 12; i++) bad += h0[i] != h1[i];
                printf("variant %d vs shipped (lazy inputs %d): %ld mismatching words of %zu\n", var, lazy, bad, n * 12);
            }
        }
    }
    const char* pn[13] = {"limb-resident partial rounds", "64-bit lanes", "FP64-pipe MDS", "u64 state, adds forced to IADD3",
                          "u64 state at every round (previous)", "FP64 partial rounds x22", "FP64 partial rounds x6", "FP64 partial rounds x8",
                          "FP64 partial rounds x10", "FP64 partial rounds x12", "FP64 partial rounds x14", "FP64 partial rounds x16", "FP64 partial rounds x18"};
    for (int var = 0; var < 13; var++) {
        float best = 1e30f; int iters = 64;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            if (var == 0) k_chain<0><<<blocks, threads>>>(o0, iters);
            if (var == 1) k_chain<1><<<blocks, threads>>>(o0, iters);
            if (var == 2) k_chain<2><<<blocks, threads>>>(o0, iters);
            if (var == 3) k_chain<3><<<blocks, threads>>>(o0, iters);
            if (var == 4) k_chain<4><<<blocks, threads>>>(o0, iters);
            if (var == 5) k_chain<5><<<blocks, threads>>>(o0, iters);
            if (var == 6) k_chain<6><<<blocks, threads>>>(o0, iters);
            if (var == 7) k_chain<7><<<blocks, threads>>>(o0, iters);
            if (var == 8) k_chain<8><<<blocks, threads>>>(o0, iters);
            if (var == 9) k_chain<9><<<blocks, threads>>>(o0, iters);
            if (var == 10) k_chain<10><<<blocks, threads>>>(o0, iters);
            if (var == 11) k_chain<11><<<blocks, threads>>>(o0, iters);
            if (var == 12) k_chain<12><<<blocks, threads>>>(o0, iters);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("poseidon %-32s %.3f ms  %.3f Gperm/s\n", pn[var], best, (double)n * iters / best / 1e6);
    }
    const char* an[6] = {"gl_mul (2^64=2^32-1 fold)", "gl_mmul (Montgomery)", "butterfly legacy", "butterfly mont", "DFMA", "DADD"};
    for (int mode = 0; mode < 6; mode++) {
        float best = 1e30f; int iters = 2048;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            if (mode == 0) k_arith<0><<<blocks, threads>>>(o0, iters, 0x123456789ABCDEFULL);
            if (mode == 1) k_arith<1><<<blocks, threads>>>(o0, iters, 0x123456789ABCDEFULL);
            if (mode == 2) k_arith<2><<<blocks, threads>>>(o0, iters, 0x123456789ABCDEFULL);
            if (mode == 3) k_arith<3><<<blocks, threads>>>(o0, iters, 0x123456789ABCDEFULL);
            if (mode == 4) k_arith<4><<<blocks, threads>>>(o0, iters, 0x123456789ABCDEFULL);
            if (mode == 5) k_arith<5><<<blocks, threads>>>(o0, iters, 0x123456789ABCDEFULL);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        double ops = (double)n * iters * 8 * 4;
        printf("%-28s %.3f ms  %.1f Gop/s\n", an[mode], best, ops / best / 1e6);
    }
    for (int mode = 1; mode < 8; mode++) {
        float best = 1e30f; int iters = 1024;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            switch (mode) {
                case 1: k_mix<1><<<blocks, threads>>>(o0, iters, 0x9E3779B1u, 0.99999991); break;
                case 2: k_mix<2><<<blocks, threads>>>(o0, iters, 0x9E3779B1u, 0.99999991); break;
                case 3: k_mix<3><<<blocks, threads>>>(o0, iters, 0x9E3779B1u, 0.99999991); break;
                case 4: k_mix<4><<<blocks, threads>>>(o0, iters, 0x9E3779B1u, 0.99999991); break;
                case 5: k_mix<5><<<blocks, threads>>>(o0, iters, 0x9E3779B1u, 0.99999991); break;
                case 6: k_mix<6><<<blocks, threads>>>(o0, iters, 0x9E3779B1u, 0.99999991); break;
                case 7: k_mix<7><<<blocks, threads>>>(o0, iters, 0x9E3779B1u, 0.99999991); break;
            }
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("pipe mix %s%s%s  %.3f ms  (%.1f G chain-steps/s per kind)\n", (mode & 1) ? "DFMA " : "     ", (mode & 2) ? "IMAD " : "     ",
               (mode & 4) ? "IADD3+LOP3 " : "           ", best, (double)n * iters * 16 * 4 / best / 1e6);
    }
    return 0;
}
