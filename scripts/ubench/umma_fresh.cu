// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS form) when BOTH operands change from one MMA to the next
// (as in the conv kernels: a new tap window of A and a new filter tap of B every MMA), per layout.  umma_rate.cu reuses the
// same operands and therefore cannot see the shared-memory fetch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I torch-em_b200/csrc -o scripts/ubench/umma_fresh.bin scripts/ubench/umma_fresh.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace b200em::umma;

__device__ __forceinline__ uint64_t mk(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = make_desc(saddr, lbo, sbo);
    d |= (uint64_t)layout << 61;
    return d;
}

// layout: 0 none (A: planes of 2944 B, 8-row groups 160 B apart, like the conv tile), 2 sw128 (8-row x 128 B atoms)
// fresh: bit 0 = rotate A over 16 windows, bit 1 = rotate B over 16 filter taps
__global__ void __launch_bounds__(128, 1) fresh_kernel(int N, int layout, int fresh, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 96 * 1024);
        uint32_t lbo_a, sbo_a, lbo_b, sbo_b, a_step, b_step;
        if (layout == 0) { lbo_a = 2944; sbo_a = 160; lbo_b = N * 16; sbo_b = 128; a_step = 2944 * 2; b_step = N * 32; }
        else { lbo_a = 16; sbo_a = 1024; lbo_b = 16; sbo_b = 1024; a_step = 128 * 128; b_step = N * 128; }   // one 64-wide K block per step
        // descriptors precomputed: the timed loop only issues
        uint64_t ad[16], bd[16];
        const int nb = layout == 0 ? 16 : (96 * 1024 / (N * 128) > 16 ? 16 : 96 * 1024 / (N * 128));
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const uint32_t ao = (fresh & 1) ? (uint32_t)(u % (layout == 0 ? 16 : 5)) * a_step : 0u;
            const uint32_t bo = (fresh & 2) ? (uint32_t)(u % nb) * b_step : 0u;
            ad[u] = mk(a_addr + ao, lbo_a, sbo_a, layout);
            bd[u] = mk(b_addr + bo, lbo_b, sbo_b, layout);
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) umma_bf16(tbase + (u & 3) * N, ad[u], bd[u], idesc, 1);
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
    cudaFuncSetAttribute(fresh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 300;
    int Ns[4] = {32, 64, 96, 128};
    printf("cycles per MMA (M=128, K=16, bf16), operands rotating: fresh bit0 = A, bit1 = B\n");
    for (int layout = 0; layout <= 2; layout += 2)
        for (int fresh = 0; fresh < 4; ++fresh)
            for (int ni = 0; ni < 4; ++ni) {
                fresh_kernel<<<148, 128, 200 * 1024>>>(Ns[ni], layout, fresh, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("layout %-5s fresh A=%d B=%d  N %3d : %7.1f cycles/MMA\n", layout == 0 ? "none" : "sw128", fresh & 1, (fresh >> 1) & 1, Ns[ni],
                       (double)c / (iters * 16));
            }
    return 0;
}
