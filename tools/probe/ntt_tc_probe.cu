// Probe: the 8 butterfly layers of a 2^8-row NTT tile as two radix-16 DFT stages on the tensor cores, against the shipped radix-8
// register steps (ntt_warp_transform) in the same shared-memory layout.  A DFT-16 is a constant 16 x 16 matrix over F_p times the 16
// elements of a "vector" -- a dense contraction: the elements are taken as their bytes as they lie in shared memory (K = 16 x 8), the
// matrix as the 8 byte limbs of w16^(jk) 2^(8b) mod p (M = 16 x 8), mma.sync.m16n8k32.u8.u8.s32 sums the byte products exactly
// (< 2^23), a thread recombines the 8 limb sums of an output and applies the inter-stage twiddle with one Montgomery multiply.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I pil2_stark_js_b200/csrc -o /tmp/ntt_tc_probe tools/probe/ntt_tc_probe.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "ntt.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

GL_D void tc_mma(u32 (&d)[4], const uint4& a, u32 b0, u32 b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// L + H 2^32 mod p, L, H < 2^52 (as poseidon_join_planes)
GL_D u64 tc_join(u64 L, u64 H) {
    const u32 H0 = (u32)H, H1 = (u32)(H >> 32);
    const u64 s = (u64)H1 * 0xFFFFFFFFu + L;
    const u32 s0 = (u32)s, s1 = (u32)(s >> 32);
    u32 r0, r1;
    asm("{\n\t.reg .u32 t1, c, m;\n\tadd.cc.u32 t1, %3, %4;\n\taddc.u32 c, 0, 0;\n\tneg.s32 m, c;\n\tadd.cc.u32 %0, %2, m;\n\taddc.u32 %1, t1, 0;\n\t}"
        : "=r"(r0), "=r"(r1) : "r"(s0), "r"(s1), "r"(H0));
    return ((u64)r1 << 32) | r0;
}
// limb sums < 2^23: pairs combine in 32 bits
GL_D u64 tc_recombine(u32 d0, u32 d1, u32 d2, u32 d3, u32 d4, u32 d5, u32 d6, u32 d7) {
    const u32 e0 = d0 + (d1 << 8), e1 = d2 + (d3 << 8), e2 = d4 + (d5 << 8), e3 = d6 + (d7 << 8);
    const u64 L = (u64)e1 * 65536u + e0, H = (u64)e3 * 65536u + e2;
    return tc_join(L, H);
}
GL_D int brev4(int k) { return (int)(__brev((unsigned)k) >> 28); }

// One radix-16 stage over the warp's region (2 columns x 256 rows).  STAGE_B = false: vectors are the 16 rows i0 + 16 j (stride 16),
// true: the 16 contiguous rows 16 beta + j.  Output k of a vector lands on the vector's row number brev4(k), times diag[row].
#ifndef TC_MODE
#define TC_MODE 0      // 0: full stage; 1: MMAs only (limb sums xor-ed into the output); 2: recombination + twiddle only (no MMA)
#endif
template <bool STAGE_B>
GL_D void tc_stage(ulonglong2* __restrict__ reg, const uint4* __restrict__ atab, const u64* __restrict__ diag, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int q = 0; q < 2; q++) {
        u32 d[8][2][4];
#pragma unroll
        for (int m = 0; m < 8; m++)
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int i = 0; i < 4; i++) d[m][c][i] = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int row = STAGE_B ? 16 * (2 * g + q) + 4 * s + t : (8 * q + g) + 16 * (4 * s + t);
            const ulonglong2 e = reg[ntt_pad(row)];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const uint4 a = atab[(s * 8 + m) * 32 + lane];
#if TC_MODE != 2
                tc_mma(d[m][0], a, (u32)e.x, (u32)(e.x >> 32));
                tc_mma(d[m][1], a, (u32)e.y, (u32)(e.y >> 32));
#else
                d[m][0][s] += a.x ^ (u32)e.x; d[m][1][s] += a.y ^ (u32)e.y;
#endif
            }
        }
        __syncwarp();
#pragma unroll
        for (int rg = 0; rg < 2; rg++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int k = 8 * rg + g, n = 2 * t + e;
                const int orow = STAGE_B ? 16 * (2 * n + q) + brev4(k) : (8 * q + n) + 16 * brev4(k);
                const u64 w = diag[orow];
                ulonglong2 v;
#if TC_MODE == 1
                v.x = 0; v.y = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { v.x ^= ((u64)d[4 * rg + i][0][e] << 8 * i) ^ d[4 * rg + i][0][2 + e]; v.y ^= ((u64)d[4 * rg + i][1][e] << 8 * i) ^ d[4 * rg + i][1][2 + e]; }
                v.x += w; v.y += w;
                reg[ntt_pad(orow)] = v;
                continue;
#endif
                v.x = gl_mmul(tc_recombine(d[4 * rg][0][e], d[4 * rg][0][2 + e], d[4 * rg + 1][0][e], d[4 * rg + 1][0][2 + e], d[4 * rg + 2][0][e],
                                           d[4 * rg + 2][0][2 + e], d[4 * rg + 3][0][e], d[4 * rg + 3][0][2 + e]), w);
                v.y = gl_mmul(tc_recombine(d[4 * rg][1][e], d[4 * rg][1][2 + e], d[4 * rg + 1][1][e], d[4 * rg + 1][1][2 + e], d[4 * rg + 2][1][e],
                                           d[4 * rg + 2][1][2 + e], d[4 * rg + 3][1][e], d[4 * rg + 3][1][2 + e]), w);
                reg[ntt_pad(orow)] = v;
            }
        __syncwarp();
    }
}

template <bool TC>
__global__ void __launch_bounds__(NTT_THREADS, TC ? 2 : 4) k_tile(u64* out, NttTables tb, const uint4* atab_g, const u64* diag_g, int iters, u64 seed) {
    extern __shared__ __align__(16) u64 sm[];
    const int RS = ntt_region_elems(8);
    ulonglong2* tile = reinterpret_cast<ulonglong2*>(sm);
    u64* TW = sm + (size_t)NTT_CP * RS * 2;
    u64* G = TW + 256;
    uint4* atab = reinterpret_cast<uint4*>(G + 16);
    u64* diag = reinterpret_cast<u64*>(atab + 1024);
    ntt_build_tw<true>(TW, G, 8, 0, 0, -1, 8, 9, tb);
    if (TC) {
        for (int i = threadIdx.x; i < 1024; i += NTT_THREADS) atab[i] = atab_g[i];
        for (int i = threadIdx.x; i < 512; i += NTT_THREADS) diag[i] = diag_g[i];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ulonglong2* reg = tile + warp * RS;
    for (int k = lane; k < 256; k += 32) {
        u64 a = seed + 0x9E3779B97F4A7C15ULL * (u64)(((blockIdx.x * 8 + warp) * 256 + k) * 2 + 1);
        a ^= a >> 29; a *= 0xBF58476D1CE4E5B9ULL; a ^= a >> 32;
        u64 b = a * 0x94D049BB133111EBULL + 12345;
        reg[ntt_pad(k)] = make_ulonglong2(gl_canon(a), gl_canon(b));
    }
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if (TC) {
            tc_stage<false>(reg, atab, diag, lane);
            tc_stage<true>(reg, atab, diag + 256, lane);
        } else {
            ntt_warp_transform<true>(reg, TW, 8);
        }
    }
    __syncwarp();
    for (int k = lane; k < 256; k += 32) {
        const ulonglong2 v = reg[ntt_pad(k)];
        u64* o = out + ((size_t)(blockIdx.x * 8 + warp) * 256 + k) * 2;
        o[0] = gl_canon(v.x);
        o[1] = gl_canon(v.y);
    }
}

static u64 hmont(u64 x) { return glh_to_mont(x); }

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 4;
    // tables of the library
    u64* tabs;
    CK(cudaMalloc(&tabs, NTT_TABLE_WORDS * 8));
    NttTables tb = {tabs, tabs + 1024, tabs + 1024 + (1 << NTT_TMAX), tabs + 1024 + 2 * (1 << NTT_TMAX)};
    ntt_setup_tables<<<4, 256>>>(tabs, tabs + 1024, tabs + 1024 + (1 << NTT_TMAX), tabs + 1024 + 2 * (1 << NTT_TMAX));
    CK(cudaDeviceSynchronize());
    // DFT-16 fragments for the inverse root (the INTT direction the DIF passes run in)
    const u64 w256i = glh_inv(glh_root(8)), w16i = glh_pow(w256i, 16);
    std::vector<u64> W(16 * 128);       // W[k][kappa = 8 j + b] = w16^(jk) 2^(8b) mod p
    for (int k = 0; k < 16; k++)
        for (int j = 0; j < 16; j++)
            for (int b = 0; b < 8; b++) W[k * 128 + 8 * j + b] = glh_mul(glh_pow(w16i, (u64)(j * k)), 1ULL << (8 * b));
    std::vector<u32> atab(1024 * 4);
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 8; m++)
            for (int lane = 0; lane < 32; lane++) {
                const int g = lane >> 2, t = lane & 3, rg = m >> 2, i = m & 3;
                u32 a[4] = {0, 0, 0, 0};
                for (int ib = 0; ib < 4; ib++) {
                    const u64 lo = W[(8 * rg + g) * 128 + 32 * s + 8 * t + ib], hi = W[(8 * rg + g) * 128 + 32 * s + 8 * t + 4 + ib];
                    a[0] |= (u32)((lo >> (8 * (2 * i))) & 0xFF) << (8 * ib);
                    a[1] |= (u32)((lo >> (8 * (2 * i + 1))) & 0xFF) << (8 * ib);
                    a[2] |= (u32)((hi >> (8 * (2 * i))) & 0xFF) << (8 * ib);
                    a[3] |= (u32)((hi >> (8 * (2 * i + 1))) & 0xFF) << (8 * ib);
                }
                memcpy(&atab[(((s * 8 + m) * 32) + lane) * 4], a, 16);
            }
    std::vector<u64> diag(512);
    auto br4 = [](int k) { return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3); };
    for (int i0 = 0; i0 < 16; i0++)
        for (int kl = 0; kl < 16; kl++) diag[i0 + 16 * br4(kl)] = hmont(glh_pow(w256i, (u64)(i0 * kl)));
    for (int i = 0; i < 256; i++) diag[256 + i] = GL_MONT_ONE;
    uint4* atab_d; u64* diag_d; u64 *o0, *o1;
    const size_t out_words = (size_t)grid * 8 * 512;
    CK(cudaMalloc(&atab_d, 16384)); CK(cudaMalloc(&diag_d, 4096)); CK(cudaMalloc(&o0, out_words * 8)); CK(cudaMalloc(&o1, out_words * 8));
    CK(cudaMemcpy(atab_d, atab.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(diag_d, diag.data(), 4096, cudaMemcpyHostToDevice));
    const size_t smem = ((size_t)NTT_CP * ntt_region_elems(8) * 2 + 256 + 16) * 8 + 16384 + 4096;
    CK(cudaFuncSetAttribute(k_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // correctness: one transform each, bit-exact
    k_tile<false><<<grid, NTT_THREADS, smem>>>(o0, tb, atab_d, diag_d, 1, 7);
    k_tile<true><<<grid, NTT_THREADS, smem>>>(o1, tb, atab_d, diag_d, 1, 7);
    CK(cudaDeviceSynchronize());
    std::vector<u64> h0(out_words), h1(out_words);
    CK(cudaMemcpy(h0.data(), o0, out_words * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), o1, out_words * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < out_words; i++) bad += h0[i] != h1[i];
    printf("tensor-core radix-16 x 2 vs radix-8 register steps, 2^8-row tiles: %zu of %zu words differ\n", bad, out_words);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int tc = 0; tc < 2; tc++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (tc) k_tile<true><<<grid, NTT_THREADS, smem>>>(o1, tb, atab_d, diag_d, iters, 7);
            else k_tile<false><<<grid, NTT_THREADS, smem>>>(o0, tb, atab_d, diag_d, iters, 7);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double bf = (double)grid * 8 * 1024 * 2 * iters;     // 2 columns x 256 rows x 8 layers / 2 per warp
        printf("%s: %.3f ms for %d transforms per warp -> %.3f T butterflies/s\n", tc ? "tensor-core radix-16 stages" : "radix-8 register steps     ", best, iters, bf / best / 1e9);
    }
    return 0;
}
