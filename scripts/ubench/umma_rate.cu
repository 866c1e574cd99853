// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operand layout and accumulator rotation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I torch-em_b200/csrc -o gpurun_out/umma_rate scripts/ubench/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace b200em::umma;

__device__ __forceinline__ uint64_t mk(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = make_desc(saddr, lbo, sbo);
    d |= (uint64_t)layout << 61;
    return d;
}

// layout: 0 none, 2 sw128, 4 sw64, 6 sw32
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int layout, int nacc, int iters, int amode, long long* out, int a_shift16 = 0, int sbo_a_override = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
        uint32_t lbo_a, sbo_a, lbo_b, sbo_b, kadv;
        if (layout == 0) { lbo_a = 2592; sbo_a = 128; lbo_b = N * 16; sbo_b = 128; kadv = 0; }
        else if (layout == 2) { lbo_a = 16; sbo_a = 1024; lbo_b = 16; sbo_b = 1024; kadv = 32; }
        else if (layout == 4) { lbo_a = 16; sbo_a = 512; lbo_b = 16; sbo_b = 512; kadv = 32; }
        else { lbo_a = 16; sbo_a = 256; lbo_b = 16; sbo_b = 256; kadv = 0; }
        if (sbo_a_override) sbo_a = sbo_a_override;
        const uint64_t ad0 = mk(a_addr + 16 * a_shift16, lbo_a, sbo_a, layout), bd0 = mk(b_addr, lbo_b, sbo_b, layout);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = i * 8 + u;
                // amode 0: same A every time; 1: rotate A over 4 row-shifted windows (tap-like)
                const uint64_t ad = ad0 + (uint64_t)(((amode ? (q & 3) * (layout == 0 ? 16 : (layout == 2 ? 8 : 4)) : 0)) + ((q & 1) * (kadv >> 4)));
                const uint64_t bd = bd0 + (uint64_t)((q & 1) * (kadv >> 4));
                umma_bf16(tbase + (q % nacc) * N, ad, bd, idesc, 1);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 500;
    int layouts[4] = {0, 6, 4, 2};
    const char* names[4] = {"none", "sw32", "sw64", "sw128"};
    int Ns[5] = {32, 64, 96, 128, 256};
    printf("cycles per MMA (M=128,K=16,bf16,cta_group::1), all 148 SMs busy; ideal = N/2\n");
    for (int amode = 0; amode < 1; ++amode)
        for (int li = 0; li < 1; ++li)
            for (int nacc = 1; nacc <= 1; nacc *= 2)
                for (int ni = 0; ni < 5; ++ni) {
                    int N = Ns[ni];
                    if (nacc * N > 512) continue;
                    rate_kernel<<<148, 128, 200 * 1024>>>(N, layouts[li], nacc, iters, amode, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s (layout %s N %d)\n", cudaGetErrorString(e), names[li], N); return 1; }
                    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("amode %d layout %-5s nacc %d N %3d : %7.1f cycles/MMA\n", amode, names[li], nacc, N, (double)c / (iters * 8));
                }
    // alignment of the 8-row x 16-byte core matrices of A (SWIZZLE_NONE): start shifted by k*16 B, 8-row-group pitch 128 / 160 B
    printf("A core-matrix alignment (layout none): shift16 = start offset in 16-B units, sbo = 8-row group pitch\n");
    for (int sbo = 128; sbo <= 160; sbo += 32)
        for (int sh = 0; sh < 8; ++sh)
            for (int ni = 0; ni < 4; ++ni) {
                int N = Ns[ni];
                rate_kernel<<<148, 128, 200 * 1024>>>(N, 0, 1, iters, 0, d, sh, sbo);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("sbo %3d shift16 %d N %3d : %7.1f cycles/MMA\n", sbo, sh, N, (double)c / (iters * 8));
            }
    return 0;
}
