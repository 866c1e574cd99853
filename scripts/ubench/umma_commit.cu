// Microbenchmark: cost of tcgen05.commit / mbarrier try_wait / fence in the single-thread MMA issue loop.
// A "stage" = NM MMAs (M=128, N=96, K=16) followed by `ncommit` commits to (distinct) mbarriers and `nwait` waits on barriers
// that are already complete.  Reports cycles per stage; the MMA-only baseline is NM * ~68.6.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I torch-em_b200/csrc -o scripts/ubench/umma_commit.bin scripts/ubench/umma_commit.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace b200em::umma;

__global__ void __launch_bounds__(128, 1) commit_kernel(int NM, int ncommit, int nwait, int nfence, int stages, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint64_t done_bar;
    __shared__ uint64_t ready[4];
    __shared__ uint32_t tmem;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&ready[i], 1);
        mbar_init(&done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem;
    if (threadIdx.x == 32) for (int i = 0; i < 4; ++i) mbar_arrive(&ready[i]);   // phase 0 of the "ready" barriers completes
    __syncthreads();
    if (warp == 1 && elect_one()) {
        const int N = 96;
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t ad = make_desc(smem_u32(smem), 2896, 160), bd = make_desc(smem_u32(smem + 64 * 1024), N * 16, 128);
        const long long t0 = clock64();
        for (int s = 0; s < stages; ++s) {
            for (int w = 0; w < nwait; ++w) mbar_wait(&ready[w], 0);
            for (int f = 0; f < nfence; ++f) tc_fence_after();
            for (int m = 0; m < NM; ++m) umma_bf16(tbase + (m & 3) * N, ad + (uint64_t)(m & 7), bd, idesc, 1);
            for (int c = 0; c < ncommit; ++c) umma_commit(&bars[(s * 2 + c) & 7]);
        }
        umma_commit(&done_bar);
        mbar_wait(&done_bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int stages = 400;
    printf("cycles per stage (NM MMAs of N=96 + commits + waits), 148 CTAs\n");
    int NMs[3] = {0, 18, 24};
    for (int ni = 0; ni < 3; ++ni)
        for (int nc = 0; nc <= 2; ++nc)
            for (int nw = 0; nw <= 2; nw += 2)
                for (int nf = 0; nf <= 2; nf += 2) {
                    commit_kernel<<<148, 128, 200 * 1024>>>(NMs[ni], nc, nw, nf, stages, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("NM %2d commits %d waits %d fences %d : %8.1f cycles/stage\n", NMs[ni], nc, nw, nf, (double)c / stages);
                }
    return 0;
}
