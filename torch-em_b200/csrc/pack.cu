// Batched weight packing: every conv weight of a model -> its bf16 tcgen05 operand image, in ONE launch.
//
// The fp32 nn.Parameter stays the master (AdamW, checkpoints, DDP untouched); the operand images are derived data and are
// rebuilt at the start of every forward (forward operands) / backward (data-gradient operands) pass.  Rebuilding
// unconditionally -- instead of caching by tensor version -- means an in-place update the version counter does not see
// (``p.data.mul_()``, EMA, old-style optimizers) can never leave a stale operand behind; one launch over the job table
// costs less than the ~40 per-weight launches it replaces (the images are 2 bytes per weight element).
//
// Layouts (see conv_umma.cu / conv_umma_ds.cu):
//   plain          [nblk][chunk][tap][plane j][NPb][8]      (bf16, or IEEE fp16 for the h16 path: B200EM_PACK_PLAIN_F16)
//   depth-stacked  [chunk][tap (b,c)][plane j][a*Nc + n][8]
// dgrad = 1 packs the transposed, tap-flipped filter (the data gradient is a forward conv with it).
// One block per (8 output channels x 8 input channels) tile of one job: the 8 x 8 x taps fp32 sub-filter is read as 8
// contiguous runs, transposed in shared memory and written as 16-byte units.
#include "common.cuh"

namespace b200em {

bool umma_pack_layout(int Cin, int Cout, int kd, int kh, int kw, int* CC, int* NP, bool f32);   // conv_umma.cu
bool ds_pack_layout(int Cin, int Cout, int kd, int kh, int kw, int* CC);              // conv_umma_ds.cu

// ci channels per block tile: 32 (every layer but the first few: 3.4 KB contiguous per output channel row, 27 loads in flight per
// thread) or 8; b200em_pack_batch_prepare and the kernel agree through this function
__host__ __device__ inline int pack_tile_ci(int Cin, int layout) { return (Cin % 32 == 0 && layout != B200EM_PACK_PLAIN_TF32) ? 32 : 8; }

__global__ void __launch_bounds__(256) pack_batch_kernel(const b200em_pack_job* __restrict__ jobs, int njobs) {
    __shared__ float tile[8][32 * 27 + 1];             // +1: rows of 864 words would all start in the same bank
    __shared__ b200em_pack_job job;
    if (threadIdx.x == 0) {
        // jobs are sorted by block_begin: last job whose first block is <= blockIdx.x
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        job = jobs[lo];
    }
    __syncthreads();
    const int Cout = job.Cout, Cin = job.Cin, dgrad = job.dgrad, CC = job.CC;
    const int taps = job.kd * job.kh * job.kw;
    const int TCI = pack_tile_ci(Cin, job.layout);
    const int local = (int)blockIdx.x - job.block_begin;
    const int tiles_ci = Cin / TCI;
    const int co0 = (local / tiles_ci) * 8, ci0 = (local % tiles_ci) * TCI;
    const float* __restrict__ w = job.w;
    const int run = TCI * taps;                        // floats per output channel in this tile (contiguous in w)
    if (TCI == 32) {
        // one warp per output-channel row, 16-byte loads (the row is 32 * taps * 4 bytes from a 16-byte aligned start), no divisions
        const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const float4* __restrict__ src = reinterpret_cast<const float4*>(w + ((size_t)(co0 + c) * Cin + ci0) * taps);
        for (int q = lane; q < run / 4; q += 32) {
            const float4 v = __ldg(src + q);
            float* d = &tile[c][4 * q];
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int i = threadIdx.x; i < 8 * run; i += blockDim.x) {
            const int c = i / run, r = i % run;
            tile[c][r] = w[((size_t)(co0 + c) * Cin + ci0) * taps + r];
        }
    }
    __syncthreads();
    const int Kc = dgrad ? Cout : Cin;                 // reduction channels of the packed operand
    if (job.layout == B200EM_PACK_PLAIN_TF32) {
        // fp32 operand of the TF32 path: [nblk][chunk (16 ch)][tap][plane j (4 ch)][NPb][4 fp32]; a tile holds two 4-channel units
        float* __restrict__ outf = reinterpret_cast<float*>(job.packed);
        const int NPb = job.NPb, nchunks = Kc / CC, J = CC / 4;
        for (int u = threadIdx.x; u < 16 * taps; u += blockDim.x) {
            const int nl = u % 8, half = (u / 8) % 2, tp = u / 16;
            float4 v;
            float* vf = reinterpret_cast<float*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) vf[e] = dgrad ? tile[4 * half + e][nl * taps + tp] : tile[nl][(4 * half + e) * taps + tp];
            const int n_ = (dgrad ? ci0 : co0) + nl, k0 = (dgrad ? co0 : ci0) + 4 * half, t_ = dgrad ? taps - 1 - tp : tp;
            const int nb = n_ / NPb, nn = n_ % NPb, chunk = k0 / CC, j = (k0 % CC) / 4;
            *reinterpret_cast<float4*>(outf + ((((size_t)(nb * nchunks + chunk) * taps + t_) * J + j) * NPb + nn) * 4) = v;
        }
        return;
    }
    __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(job.packed);
    const int Nc = dgrad ? Cin : Cout;                 // output channels of the packed operand
    const int nchunks = Kc / CC, J = CC / 8;
    const int thw = job.kh * job.kw;
    const bool f16 = job.layout == B200EM_PACK_PLAIN_F16 || job.layout == B200EM_PACK_DEPTH_STACKED_F16;
    // 16-byte units (8 reduction channels of one operand row n and tap): forward n = co (8 rows) x TCI/8 ci groups, data gradient
    // n = ci (TCI rows) x the tile's one co group; consecutive threads take consecutive n (adjacent 16-byte units of the image)
    const int nrows = dgrad ? TCI : 8, kgroups = dgrad ? 1 : TCI / 8;
    for (int u = threadIdx.x; u < nrows * kgroups * taps; u += blockDim.x) {
        int nl, kg = 0, tp;                            // (rows x k-groups) per tap is 32 or 8: shifts, no divisions
        if (TCI == 32) {
            tp = u >> 5;
            if (dgrad) nl = u & 31; else { nl = u & 7; kg = (u >> 3) & 3; }
        } else {
            nl = u & 7; tp = u >> 3;
        }
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            // forward: n = co, k = ci;  dgrad: n = ci, k = co
            const float f = dgrad ? tile[e][nl * taps + tp] : tile[nl][(kg * 8 + e) * taps + tp];
            if (f16) reinterpret_cast<__half*>(v)[e] = __float2half_rn(f);
            else v[e] = __float2bfloat16_rn(f);
        }
        const int n_ = (dgrad ? ci0 : co0) + nl, k0 = dgrad ? co0 : ci0 + kg * 8, t_ = dgrad ? taps - 1 - tp : tp;
        const int chunk = k0 / CC, j = (k0 % CC) / 8;
        size_t o;
        if (job.layout == B200EM_PACK_PLAIN || job.layout == B200EM_PACK_PLAIN_F16) {
            const int NPb = job.NPb;
            const int nb = n_ / NPb, nn = n_ % NPb;
            o = ((((size_t)(nb * nchunks + chunk) * taps + t_) * J + j) * NPb + nn) * 8;
        } else {
            const int a = t_ / thw, bc = t_ % thw;
            o = ((((size_t)chunk * thw + bc) * J + j) * (3 * Nc) + (size_t)a * Nc + n_) * 8;
        }
        *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(v);
    }
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_pack_batch_prepare(b200em_pack_job* jobs, int njobs, int* total_blocks) {
    B2_CHECK_ARG(jobs && njobs > 0 && total_blocks, "pack_batch_prepare: bad arguments");
    int blocks = 0;
    for (int i = 0; i < njobs; ++i) {
        b200em_pack_job& j = jobs[i];
        B2_CHECK_ARG(aligned16(j.w) && aligned16(j.packed), "pack_batch_prepare: job %d: weight and image must be 16-byte aligned", i);
        B2_CHECK_ARG(j.Cout > 0 && j.Cin > 0 && j.Cout % 8 == 0 && j.Cin % 8 == 0 && j.kd * j.kh * j.kw <= 27,
                     "pack_batch_prepare: job %d: channels must be multiples of 8 and taps <= 27", i);
        const int n_ = j.dgrad ? j.Cin : j.Cout, k_ = j.dgrad ? j.Cout : j.Cin;   // operand's output / reduction channels
        bool ok;
        if (j.layout == B200EM_PACK_PLAIN || j.layout == B200EM_PACK_PLAIN_TF32 || j.layout == B200EM_PACK_PLAIN_F16) {
            ok = umma_pack_layout(k_, n_, j.kd, j.kh, j.kw, &j.CC, &j.NPb, j.layout == B200EM_PACK_PLAIN_TF32);
        } else if (j.layout == B200EM_PACK_DEPTH_STACKED || j.layout == B200EM_PACK_DEPTH_STACKED_F16) {
            j.NPb = 0;
            ok = ds_pack_layout(k_, n_, j.kd, j.kh, j.kw, &j.CC);
        } else {
            set_error("pack_batch_prepare: job %d: unknown layout %d", i, j.layout);
            return 1;
        }
        if (!ok) {
            set_error("pack_batch_prepare: job %d: shape (%d -> %d, %dx%dx%d) not supported by layout %d", i, k_, n_, j.kd, j.kh, j.kw,
                      j.layout);
            return 2;
        }
        j.block_begin = blocks;
        blocks += (j.Cout / 8) * (j.Cin / pack_tile_ci(j.Cin, j.layout));
    }
    *total_blocks = blocks;
    return 0;
}

int b200em_pack_batch(const b200em_pack_job* jobs_device, int njobs, int total_blocks, void* stream) {
    B2_CHECK_ARG(jobs_device && njobs > 0 && total_blocks > 0, "pack_batch: bad arguments");
    pack_batch_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_device, njobs);
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
