// Network head: out_conv (1x1x1 nn.Conv3d, unet.py:638,202-203) + final activation (unet.py:162-172,204-205),
// fused with the NDHWC -> NCDHW layout change, forward and backward.  HBM-bound (AI ~ 2-9 FLOP/B):
// one pass over x, one coalesced write of the fp32 NCDHW prediction.
#include <stdlib.h>

#include "common.cuh"

namespace b200em {

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == B200EM_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    if (act == B200EM_ACT_RELU) return fmaxf(v, 0.f);
    if (act == B200EM_ACT_TANH) return tanhf(v);
    return v;
}
// derivative expressed through the OUTPUT of the activation (what the forward keeps)
__device__ __forceinline__ float act_bwd(float out, int act) {
    if (act == B200EM_ACT_SIGMOID) return out * (1.f - out);
    if (act == B200EM_ACT_RELU) return out > 0.f ? 1.f : 0.f;
    if (act == B200EM_ACT_TANH) return 1.f - out * out;
    return 1.f;
}

constexpr int HEAD_COB = 8;  // output channels per register pass

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ out, int64_t S, int Cin, int Cout, int act, int64_t total) {
    for (int64_t vox = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; vox < total; vox += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = vox / S, s = vox % S;
        const T* xp = x + vox * x_ld;
        for (int co0 = 0; co0 < Cout; co0 += HEAD_COB) {
            float acc[HEAD_COB];
#pragma unroll
            for (int j = 0; j < HEAD_COB; ++j) acc[j] = (bias && co0 + j < Cout) ? __ldg(bias + co0 + j) : 0.f;
            for (int c = 0; c < Cin; c += VEC) {
                float xv[VEC];
                Vec<T, VEC>::load(xp + c, xv);
#pragma unroll
                for (int j = 0; j < HEAD_COB; ++j) {
                    if (co0 + j < Cout) {
                        const float* wr = w + (size_t)(co0 + j) * Cin + c;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) acc[j] = fmaf(__ldg(wr + k), xv[k], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < HEAD_COB; ++j)
                if (co0 + j < Cout) out[((size_t)n * Cout + co0 + j) * S + s] = act_fwd(acc[j], act);
        }
    }
}

constexpr int HB_MAXCO = 32;   // output channels handled per call
constexpr int HB_TILE = 256;   // voxels per tile (= threads)

// grad_out/out: (N, Cout_total, S) fp32; this call handles channels [co_begin, co_begin+Cout).
template <typename T, int VEC, int MAXP>
__global__ void __launch_bounds__(HB_TILE)
head_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ out, const T* __restrict__ x, int64_t x_ld,
                const float* __restrict__ w, T* __restrict__ dx, int64_t dx_ld, float* __restrict__ dw,
                float* __restrict__ db, int64_t S, int Cin, int Cout_total, int co_begin, int Cout, int act,
                int accumulate, int relu_mask, int64_t total) {
    __shared__ float dzs[HB_MAXCO][HB_TILE];
    const int tid = threadIdx.x;
    float accp[MAXP];
#pragma unroll
    for (int i = 0; i < MAXP; ++i) accp[i] = 0.f;
    float accb_slots[HB_MAXCO / 8];  // warp w accumulates db for co = w, w+8, ...
#pragma unroll
    for (int i = 0; i < HB_MAXCO / 8; ++i) accb_slots[i] = 0.f;
    const int npairs = Cout * Cin;
    int rep = 1;
    while (rep * 2 * npairs <= HB_TILE) rep *= 2;
    const int64_t ntiles = (total + HB_TILE - 1) / HB_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t vox = tile * HB_TILE + tid;
        const bool live = vox < total;
        const int64_t n = live ? vox / S : 0, s = live ? vox % S : 0;
        float dz[HB_MAXCO];
#pragma unroll
        for (int j = 0; j < HB_MAXCO; ++j) {
            float v = 0.f;
            if (live && j < Cout) {
                size_t o = ((size_t)n * Cout_total + co_begin + j) * S + s;
                v = grad_out[o] * act_bwd(out[o], act);
            }
            dz[j] = v;
            dzs[j][tid] = v;
        }
        // dx = W^T dz
        if (live && dx) {
            T* dxp = dx + vox * dx_ld;
            for (int c = 0; c < Cin; c += VEC) {
                float r[VEC], xv[VEC];
                if (relu_mask) Vec<T, VEC>::load(x + vox * x_ld + c, xv);
                if (accumulate) Vec<T, VEC>::load(dxp + c, r);
                else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) r[k] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < HB_MAXCO; ++j) {
                    if (j < Cout) {
                        const float* wr = w + (size_t)(co_begin + j) * Cin + c;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) r[k] = fmaf(__ldg(wr + k), dz[j], r[k]);
                    }
                }
                if (relu_mask) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (!(xv[k] > 0.f)) r[k] = 0.f;
                }
                Vec<T, VEC>::store(dxp + c, r);
            }
        }
        __syncthreads();
        // dw[co][ci] partial over this tile: thread p handles pairs p, p+256, ...; when there are fewer pairs than
        // threads, `rep` threads share a pair and split the tile's voxels (all 256 threads stay busy)
        const int64_t v0 = tile * HB_TILE;
        const int nv = (int)((total - v0) < HB_TILE ? (total - v0) : HB_TILE);
        if (rep > 1) {
            if (tid < rep * npairs) {
                const int p = tid % npairs, sub = tid / npairs;
                const int co = p / Cin, ci = p % Cin;
                const T* xc = x + v0 * x_ld + ci;
                float a = 0.f;
                for (int v = sub; v < nv; v += rep) a = fmaf(dzs[co][v], to_f<T>(xc[(size_t)v * x_ld]), a);
                accp[0] += a;
            }
        } else {
#pragma unroll
            for (int i = 0; i < MAXP; ++i) {
                const int p = tid + i * HB_TILE;
                if (p < npairs) {
                    const int co = p / Cin, ci = p % Cin;
                    const T* xc = x + v0 * x_ld + ci;
                    float a = 0.f;
                    for (int v = 0; v < nv; ++v) a = fmaf(dzs[co][v], to_f<T>(xc[(size_t)v * x_ld]), a);
                    accp[i] += a;
                }
            }
        }
        // db: warp wi sums rows co = wi + 8*k
        {
            const int wi = tid >> 5, lane = tid & 31;
#pragma unroll
            for (int k = 0; k < HB_MAXCO / 8; ++k) {
                const int co = wi + 8 * k;
                if (co < Cout) {
                    float a = 0.f;
                    for (int v = lane; v < HB_TILE; v += 32) a += dzs[co][v];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    accb_slots[k] += a;
                }
            }
        }
        __syncthreads();
    }
    if (rep > 1) {
        if (tid < rep * npairs) {
            const int p = tid % npairs;
            atomicAdd(dw + (size_t)(co_begin + p / Cin) * Cin + p % Cin, accp[0]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int p = tid + i * HB_TILE;
            if (p < npairs) atomicAdd(dw + (size_t)(co_begin + p / Cin) * Cin + p % Cin, accp[i]);
        }
    }
    if (db && (tid & 31) == 0) {
        const int wi = tid >> 5;
#pragma unroll
        for (int k = 0; k < HB_MAXCO / 8; ++k) {
            const int co = wi + 8 * k;
            if (co < Cout) atomicAdd(db + co_begin + co, accb_slots[k]);
        }
    }
}

// ---- small heads (Cin <= 32, Cout <= 2: the U-Net's out_conv on the first feature level) -------------------------------
// One thread per voxel, the Cout x Cin filter and the per-thread dW partials live in registers, x is read ONCE (it also
// carries the ReLU mask of the block below), dx is written once; dW / db are reduced over the block at the end.
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(256)
head_fwd_small_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, int64_t S, int act, int64_t total) {
    constexpr int V = FullVec<T>::value;
    float wr[COUT][CIN], br[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
        br[j] = bias ? bias[j] : 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) wr[j][c] = w[j * CIN + c];
    }
    for (int64_t vox = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; vox < total; vox += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = vox / S, s_ = vox % S;
        const T* xp = x + vox * x_ld;
        float acc[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = br[j];
#pragma unroll
        for (int c = 0; c < CIN; c += V) {
            float xv[V];
            Vec<T, V>::load(xp + c, xv);
#pragma unroll
            for (int j = 0; j < COUT; ++j)
#pragma unroll
                for (int k = 0; k < V; ++k) acc[j] = fmaf(wr[j][c + k], xv[k], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < COUT; ++j) out[((size_t)n * COUT + j) * S + s_] = act_fwd(acc[j], act);
    }
}

// cin_total = CIN * slices input channels: warp w of the grid handles channel slice (w % slices) of 32 consecutive voxels, so a
// thread still owns one voxel x CIN channels (the dW partials fit in registers) and wider heads (Cin = 64, 128) take the same path.
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(256, 2)
head_bwd_small_kernel(const float* __restrict__ grad_out, const float* __restrict__ out, const T* __restrict__ x, int64_t x_ld,
                      const float* __restrict__ w, T* __restrict__ dx, int64_t dx_ld, float* __restrict__ dw,
                      float* __restrict__ db, int64_t S, int act, int relu_mask, int64_t total, int cin_total, float* __restrict__ absmax) {
    unsigned am = 0;
    constexpr int V = FullVec<T>::value;
    constexpr int MAXC = 128;
    __shared__ float red[COUT * MAXC + COUT];        // block-level dW | db partial sums
    __shared__ float wr_s[COUT * MAXC];              // filter in shared memory (broadcast reads): two blocks per SM fit the register file
    float aw[COUT][CIN], ab[COUT];
    const int slices = cin_total / CIN;
    for (int i = threadIdx.x; i < COUT * cin_total; i += blockDim.x) wr_s[i] = w[i];
    for (int i = threadIdx.x; i < COUT * cin_total + COUT; i += blockDim.x) red[i] = 0.f;
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
        ab[j] = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) aw[j][c] = 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int sl = (int)(gwarp % slices);            // (gridDim.x * 8) % slices == 0: a warp keeps its slice over the grid-stride loop
    const int c0 = sl * CIN;
    for (int64_t vox = (gwarp / slices) * 32 + lane; vox < total; vox += (nwarps / slices) * 32) {
        const int64_t n = vox / S, s_ = vox % S;
        float dz[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            const size_t o = ((size_t)n * COUT + j) * S + s_;
            dz[j] = grad_out[o] * act_bwd(out[o], act);
            ab[j] += dz[j];
        }
        const T* xp = x + vox * x_ld + c0;
#pragma unroll
        for (int c = 0; c < CIN; c += V) {
            float xv[V], r[V];
            Vec<T, V>::load(xp + c, xv);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float g = 0.f;
#pragma unroll
                for (int j = 0; j < COUT; ++j) {
                    g = fmaf(wr_s[j * cin_total + c0 + c + k], dz[j], g);
                    aw[j][c + k] = fmaf(dz[j], xv[k], aw[j][c + k]);
                }
                r[k] = (relu_mask && !(xv[k] > 0.f)) ? 0.f : g;
                am = absmax_acc(am, r[k]);
            }
            if (dx) Vec<T, V>::store(dx + vox * dx_ld + c0 + c, r);
        }
    }
    absmax_flush(am, absmax);
    // block reduction: warp shuffles, one shared-memory atomic per warp and value, then COUT*cin_total + COUT atomics per block
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            float a = aw[j][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) atomicAdd(&red[j * cin_total + c0 + c], a);
        }
        float b = ab[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0 && sl == 0) atomicAdd(&red[COUT * cin_total + j], b);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < COUT * cin_total + COUT; i += blockDim.x) {
        if (i < COUT * cin_total) atomicAdd(dw + i, red[i]);
        else if (db) atomicAdd(db + (i - COUT * cin_total), red[i]);
    }
}

// ---- mid-size heads (Cout = 3 .. 12: affinity / multi-class outputs on the first feature level) ---------------------------------
// The filter of such a head (e.g. 12 x 32) and its dW partials do not fit one thread's registers, so a voxel is shared by several
// threads: FORWARD by groups of JG output channels (warp-uniform group: a warp reads 32 consecutive voxels' full x rows --
// coalesced, the other groups' warps hit L1 -- and writes JG coalesced NCDHW rows), BACKWARD by 8-channel slices of x (lane =
// 8 voxels x 4 slices: contiguous 64 / 128 bytes per voxel; a thread keeps COUT x 8 dW partials and produces dx for its 8
// channels from all COUT output gradients, so no cross-thread reduction is needed in the loop).
template <typename T, int COUT, int JG>
__global__ void __launch_bounds__(256)
head_fwd_mid_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, int64_t S, int Cin, int act, int64_t total) {
    constexpr int V = FullVec<T>::value, NG = COUT / JG, MAXC = 128;
    __shared__ __align__(16) float w_s[COUT * MAXC];
    for (int i = threadIdx.x; i < COUT * Cin; i += blockDim.x) w_s[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int jg = (int)(gwarp % NG);                  // (gridDim.x * 8) % NG == 0: a warp keeps its group
    float br[JG];
#pragma unroll
    for (int j = 0; j < JG; ++j) br[j] = bias ? bias[jg * JG + j] : 0.f;
    for (int64_t vox = (gwarp / NG) * 32 + lane; vox < total; vox += (nwarps / NG) * 32) {
        const int64_t n = vox / S, s_ = vox % S;
        const T* xp = x + vox * x_ld;
        float acc[JG];
#pragma unroll
        for (int j = 0; j < JG; ++j) acc[j] = br[j];
        for (int c = 0; c < Cin; c += V) {
            float xv[V];
            Vec<T, V>::load(xp + c, xv);
#pragma unroll
            for (int j = 0; j < JG; ++j) {
                const float* wr = w_s + (jg * JG + j) * Cin + c;     // warp-uniform address: broadcast
#pragma unroll
                for (int k = 0; k < V; k += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
                    acc[j] = fmaf(w4.x, xv[k], acc[j]); acc[j] = fmaf(w4.y, xv[k + 1], acc[j]);
                    acc[j] = fmaf(w4.z, xv[k + 2], acc[j]); acc[j] = fmaf(w4.w, xv[k + 3], acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < JG; ++j) out[((size_t)n * COUT + jg * JG + j) * S + s_] = act_fwd(acc[j], act);
    }
}

// (Tried and slower on the 12-channel affinity head: splitting the output channels over two lane halves to halve the registers --
// 2.6 ms; loading the output gradients once per voxel, coalesced, and handing them out by shuffles -- 1.9 ms; this form: 1.5 ms.
// The loop is bound by the latency of its loads at 8 resident warps per SM.)
template <typename T, int COUT>
__global__ void __launch_bounds__(128, COUT > 8 ? 2 : 3)
head_bwd_mid_kernel(const float* __restrict__ grad_out, const float* __restrict__ out, const T* __restrict__ x, int64_t x_ld,
                    const float* __restrict__ w, T* __restrict__ dx, int64_t dx_ld, float* __restrict__ dw, float* __restrict__ db,
                    int64_t S, int act, int relu_mask, int64_t total, int Cin, float* __restrict__ absmax) {
    constexpr int MAXC = 128;
    __shared__ float red[COUT * MAXC + COUT];          // block-level dW | db partial sums
    __shared__ float w_s[COUT * MAXC];                 // filter, transposed to [c][COUT]: a thread reads its 8 channels' weights
    const int nsl = Cin / 8;                           // 8-channel slices of x: 4 (Cin = 32), 8, 16; a warp covers 32 / nsl voxels
    for (int i = threadIdx.x; i < COUT * Cin; i += blockDim.x) w_s[(i % Cin) * COUT + i / Cin] = w[i];
    for (int i = threadIdx.x; i < COUT * Cin + COUT; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sl = lane % nsl, vl = lane / nsl, vpw = 32 / nsl;      // slice, voxel within the warp, voxels per warp
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float aw[COUT][8], ab[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
        ab[j] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) aw[j][c] = 0.f;
    }
    unsigned am = 0;
    for (int64_t vox = gwarp * vpw + vl; vox < total; vox += nwarps * vpw) {
        const int64_t n = vox / S, s_ = vox % S;
        float dz[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            const size_t o = ((size_t)n * COUT + j) * S + s_;
            dz[j] = grad_out[o] * act_bwd(out[o], act);
            ab[j] += dz[j];
        }
        float xv[8], r[8];
        const T* xp = x + vox * x_ld + sl * 8;
        if constexpr (sizeof(T) == 2) {
            Vec<T, 8>::load(xp, xv);
        } else {
            Vec<T, 4>::load(xp, *reinterpret_cast<float(*)[4]>(xv));
            Vec<T, 4>::load(xp + 4, *reinterpret_cast<float(*)[4]>(xv + 4));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float* wk = w_s + (sl * 8 + k) * COUT;
            float g = 0.f;
#pragma unroll
            for (int j = 0; j < COUT; ++j) {
                g = fmaf(wk[j], dz[j], g);
                aw[j][k] = fmaf(dz[j], xv[k], aw[j][k]);
            }
            r[k] = (relu_mask && !(xv[k] > 0.f)) ? 0.f : g;
            am = absmax_acc(am, r[k]);
        }
        if (dx) {
            T* dp = dx + vox * dx_ld + sl * 8;
            if constexpr (sizeof(T) == 2) {
                Vec<T, 8>::store(dp, r);
            } else {
                Vec<T, 4>::store(dp, *reinterpret_cast<float(*)[4]>(r));
                Vec<T, 4>::store(dp + 4, *reinterpret_cast<float(*)[4]>(r + 4));
            }
        }
    }
    absmax_flush(am, absmax);
    // reduce over the lanes that share a slice (the voxel bits of the lane index); the bias gradient is taken from slice 0's lanes
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float a = aw[j][c];
            for (int o = 16; o >= nsl; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (vl == 0) atomicAdd(&red[j * Cin + sl * 8 + c], a);
        }
        float b = ab[j];
        for (int o = 16; o >= nsl; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) atomicAdd(&red[COUT * Cin + j], b);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < COUT * Cin + COUT; i += blockDim.x) {
        if (i < COUT * Cin) atomicAdd(dw + i, red[i]);
        else if (db) atomicAdd(db + (i - COUT * Cin), red[i]);
    }
}

constexpr int HBT_TV = 128, HBT_NST = 3, HBT_CIN = 32;      // staged head kernels: voxels per tile, ring depth, input channels

// ---- mid-size head forward, shared-memory staged (Cin = 32) -------------------------------------------------------------------------
// The group form above reads every x row once per output-channel group through L1 with 64-byte lane strides (16 cache lines per
// warp load): ncu shows it bound by L1 throughput (79 %) at 1.7 TB/s.  Here the x rows of a 128-voxel tile arrive ONCE by cp.async
// (three-deep ring, rows padded to 80 / 144 bytes so that a quarter-warp's 16-byte reads hit distinct banks), a thread owns a
// voxel and all COUT outputs, the filter is read as 16-byte broadcasts and the FMAs are packed over channel pairs.
template <typename T> struct HftSmem {
    static constexpr int row_bytes = HBT_CIN * (int)sizeof(T) + 16;
    static constexpr int stage_bytes = HBT_TV * row_bytes;
};

template <typename T, int COUT>
__global__ void __launch_bounds__(128, 4)
head_fwd_mid_tiled_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ w, const float* __restrict__ bias,
                          float* __restrict__ out, int64_t S, int act, int N) {
    using L = HftSmem<T>;
    constexpr int TV = HBT_TV, Cin = HBT_CIN, XCH = Cin * (int)sizeof(T) / 16;
    extern __shared__ __align__(16) uint8_t hf_smem[];
    float4* w_s = reinterpret_cast<float4*>(hf_smem + HBT_NST * L::stage_bytes);     // [j][Cin / 4]
    for (int i = threadIdx.x; i < COUT * Cin / 4; i += blockDim.x) w_s[i] = reinterpret_cast<const float4*>(w)[i];
    const int tps = (int)((S + TV - 1) / TV);
    const int ntiles = N * tps;
    const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(hf_smem));
    auto issue = [&](int tile, int stage) {
        if (tile < ntiles) {
            const int n = tile / tps;
            const int64_t s0 = (int64_t)(tile % tps) * TV;
            const int vcnt = (int)min((int64_t)TV, S - s0);
            const T* xt = x + ((size_t)n * S + s0) * x_ld;
            const uint32_t sb = smem0 + stage * L::stage_bytes;
            for (int i = threadIdx.x; i < TV * XCH; i += 128) {
                const int v = i / XCH, q = i % XCH;
                const bool in = v < vcnt;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sb + (uint32_t)(v * L::row_bytes + q * 16)),
                             "l"(in ? reinterpret_cast<const uint8_t*>(xt + (size_t)v * x_ld) + q * 16 : reinterpret_cast<const uint8_t*>(x)),
                             "r"(in ? 16 : 0)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float br[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) br[j] = bias ? bias[j] : 0.f;
#pragma unroll
    for (int k = 0; k < HBT_NST - 1; ++k) issue(blockIdx.x + k * gridDim.x, k);
    int stage = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        { const int ps = stage + HBT_NST - 1; issue(tile + (HBT_NST - 1) * gridDim.x, ps >= HBT_NST ? ps - HBT_NST : ps); }
        asm volatile("cp.async.wait_group %0;" ::"n"(HBT_NST - 1) : "memory");
        __syncthreads();
        const int n = tile / tps;
        const int64_t s0 = (int64_t)(tile % tps) * TV;
        const int vcnt = (int)min((int64_t)TV, S - s0);
        const int v = threadIdx.x;
        float2 x2[Cin / 2];
        {
            const T* xs = reinterpret_cast<const T*>(hf_smem + stage * L::stage_bytes + v * L::row_bytes);
            constexpr int V = FullVec<T>::value;
#pragma unroll
            for (int c = 0; c < Cin; c += V) {
                float t[V];
                Vec<T, V>::load(xs + c, t);
#pragma unroll
                for (int e = 0; e < V; e += 2) x2[(c + e) / 2] = make_float2(t[e], t[e + 1]);
            }
        }
        float* op = out + (size_t)n * COUT * S + s0 + v;
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            float2 a = make_float2(br[j], 0.f);
#pragma unroll
            for (int q = 0; q < Cin / 4; ++q) {
                const float4 w4 = w_s[j * (Cin / 4) + q];
                a = __ffma2_rn(make_float2(w4.x, w4.y), x2[2 * q], a);
                a = __ffma2_rn(make_float2(w4.z, w4.w), x2[2 * q + 1], a);
            }
            if (v < vcnt) op[(size_t)j * S] = act_fwd(a.x + a.y, act);
        }
        __syncthreads();
        if (++stage == HBT_NST) stage = 0;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---- mid-size head backward, shared-memory staged (Cin = 32: the affinity head of cfg3) ------------------------------------------
// The register form above keeps 8 warps per SM (252 registers) with ~1.3 KB of loads in flight per warp: ~10 KB per SM against the
// ~30 KB that HBM's bandwidth x latency product asks for -- it runs at 1.1 TB/s whatever its instruction mix (packed FMAs, a
// shared-memory filter and a software-pipelined variant all measured the same or worse).  Here the operands of a 128-voxel tile
// (12 + 12 planar rows of the output gradient / output, the x rows) are staged by cp.async through a three-deep shared-memory
// ring, so ~40 KB per block are in flight independently of the registers, and the arithmetic reads shared memory: the dW
// partials (96 registers) stay thread-private as before, the FMAs are packed (fma.rn.f32x2 over channel pairs), the filter is
// read as 16-byte broadcast loads and the activation derivative is branch-free.
template <typename T, int COUT> struct HbtSmem {
    static constexpr int JP = (COUT + 1) / 2;
    static constexpr int go_bytes = 2 * COUT * HBT_TV * 4, x_bytes = HBT_TV * HBT_CIN * (int)sizeof(T);
    static constexpr int stage_bytes = go_bytes + x_bytes;
    static constexpr int WSL = 4 * JP + 1;           // float4 per 8-channel slice of the filter, +1: the four slices of a quarter-warp read distinct banks
    static constexpr int w_bytes = (HBT_CIN / 8) * WSL * 16, red_bytes = (COUT * HBT_CIN + COUT) * 4;
    static constexpr int total = HBT_NST * stage_bytes + w_bytes + red_bytes;
};

template <typename T, int COUT>
__global__ void __launch_bounds__(128, 2)
head_bwd_mid_tiled_kernel(const float* __restrict__ grad_out, const float* __restrict__ out, const T* __restrict__ x, int64_t x_ld,
                          const float* __restrict__ w, T* __restrict__ dx, int64_t dx_ld, float* __restrict__ dw, float* __restrict__ db,
                          int64_t S, int act, int relu_mask, int N, float* __restrict__ absmax) {
    using L = HbtSmem<T, COUT>;
    constexpr int JP = L::JP, TV = HBT_TV, Cin = HBT_CIN;
    constexpr int XCH = Cin * (int)sizeof(T) / 16;     // 16-byte chunks of one x row
    extern __shared__ __align__(16) uint8_t hb_smem[];
    float4* w_s = reinterpret_cast<float4*>(hb_smem + HBT_NST * L::stage_bytes);   // [channel pair][output pair] (w[2jp][2cp..], w[2jp+1][2cp..])
    float* red = reinterpret_cast<float*>(hb_smem + HBT_NST * L::stage_bytes + L::w_bytes);
    for (int i = threadIdx.x; i < (Cin / 2) * JP; i += blockDim.x) {
        const int cp = i / JP, j0 = 2 * (i % JP), j1 = j0 + 1;
        w_s[(cp / 4) * L::WSL + (cp % 4) * JP + i % JP] = make_float4(w[j0 * Cin + 2 * cp], w[j0 * Cin + 2 * cp + 1], j1 < COUT ? w[j1 * Cin + 2 * cp] : 0.f,
                                                                      j1 < COUT ? w[j1 * Cin + 2 * cp + 1] : 0.f);
    }
    for (int i = threadIdx.x; i < COUT * Cin + COUT; i += blockDim.x) red[i] = 0.f;
    const int tps = (int)((S + TV - 1) / TV);           // tiles per sample: a tile never straddles two samples
    const int ntiles = N * tps;
    const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(hb_smem));

    auto issue = [&](int tile, int stage) {
        if (tile < ntiles) {
            const int n = tile / tps;
            const int64_t s0 = (int64_t)(tile % tps) * TV;
            const int vcnt = (int)min((int64_t)TV, S - s0);                          // multiple of 4 (S % 4 == 0)
            const uint32_t sb = smem0 + stage * L::stage_bytes;
            for (int i = threadIdx.x; i < 2 * COUT * (TV / 4); i += 128) {           // planar rows: [array][j][TV] fp32
                const int ch = i % (TV / 4), row = i / (TV / 4), j = row % COUT;
                const float* src = (row < COUT ? grad_out : out) + ((size_t)n * COUT + j) * S + s0 + 4 * ch;
                const bool in = 4 * ch < vcnt;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sb + (uint32_t)(row * TV + 4 * ch) * 4), "l"(in ? src : grad_out),
                             "r"(in ? 16 : 0)
                             : "memory");
            }
            const T* xt = x + ((size_t)n * S + s0) * x_ld;
            for (int i = threadIdx.x; i < TV * XCH; i += 128) {                      // x rows: [voxel][Cin]
                const int v = i / XCH, q = i % XCH;
                const bool in = v < vcnt;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sb + (uint32_t)(L::go_bytes + i * 16)),
                             "l"(in ? reinterpret_cast<const uint8_t*>(xt + (size_t)v * x_ld) + q * 16 : reinterpret_cast<const uint8_t*>(x)),
                             "r"(in ? 16 : 0)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = lane & 3, vl = lane >> 2;            // 8-channel slice of x, voxel within the warp
    const float a0 = (act == B200EM_ACT_SIGMOID) ? 0.f : 1.f, a1 = (act == B200EM_ACT_SIGMOID) ? 1.f : 0.f,
                a2 = (act == B200EM_ACT_SIGMOID || act == B200EM_ACT_TANH) ? -1.f : 0.f;
    const bool step = act == B200EM_ACT_RELU;
    float2 aw[COUT][4];
    float ab[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
        ab[j] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) aw[j][c] = make_float2(0.f, 0.f);
    }
    unsigned am = 0;
    const float4* wsl = w_s + sl * L::WSL;
#pragma unroll
    for (int k = 0; k < HBT_NST - 1; ++k) issue(blockIdx.x + k * gridDim.x, k);
    int stage = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        { const int ps = stage + HBT_NST - 1; issue(tile + (HBT_NST - 1) * gridDim.x, ps >= HBT_NST ? ps - HBT_NST : ps); }
        asm volatile("cp.async.wait_group %0;" ::"n"(HBT_NST - 1) : "memory");
        __syncthreads();
        const int n = tile / tps;
        const int64_t s0 = (int64_t)(tile % tps) * TV;
        const int vcnt = (int)min((int64_t)TV, S - s0);
        const uint8_t* sb = hb_smem + stage * L::stage_bytes;
        const float* g_s = reinterpret_cast<const float*>(sb);
        const float* o_s = g_s + COUT * TV;
#pragma unroll 1
        for (int pass = 0; pass < TV / 64; ++pass) {
            // two voxels per thread and pass (v, v + 8): one filter read from shared memory serves both
            const int vb = pass * 64 + warp * 16;
            if (vb >= vcnt) break;                       // warp-uniform; zero-filled voxels beyond vcnt contribute nothing and are not stored
            float dz[2][2 * JP], xv[2][8], r[2][8];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = vb + 8 * u + vl;
#pragma unroll
                for (int j = 0; j < COUT; ++j) {
                    const float o_ = o_s[j * TV + v];
                    float d = fmaf(o_, fmaf(a2, o_, a1), a0);
                    d = step ? (o_ > 0.f ? 1.f : 0.f) : d;
                    dz[u][j] = g_s[j * TV + v] * d;
                    ab[j] += dz[u][j];
                }
                if (COUT & 1) dz[u][2 * JP - 1] = 0.f;
                const T* xs = reinterpret_cast<const T*>(sb + L::go_bytes) + v * Cin + sl * 8;
                if constexpr (sizeof(T) == 2) {
                    Vec<T, 8>::load(xs, xv[u]);
                } else {
                    Vec<T, 4>::load(xs, *reinterpret_cast<float(*)[4]>(xv[u]));
                    Vec<T, 4>::load(xs + 4, *reinterpret_cast<float(*)[4]>(xv[u] + 4));
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float2 x2[2] = {make_float2(xv[0][2 * c], xv[0][2 * c + 1]), make_float2(xv[1][2 * c], xv[1][2 * c + 1])};
                float2 g[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int jp = 0; jp < JP; ++jp) {
                    const float4 w4 = wsl[c * JP + jp];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float2 d0 = make_float2(dz[u][2 * jp], dz[u][2 * jp]), d1 = make_float2(dz[u][2 * jp + 1], dz[u][2 * jp + 1]);
                        g[u] = __ffma2_rn(make_float2(w4.x, w4.y), d0, g[u]);
                        aw[2 * jp][c] = __ffma2_rn(d0, x2[u], aw[2 * jp][c]);
                        if (2 * jp + 1 < COUT) {
                            g[u] = __ffma2_rn(make_float2(w4.z, w4.w), d1, g[u]);
                            aw[2 * jp + 1][c] = __ffma2_rn(d1, x2[u], aw[2 * jp + 1][c]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    r[u][2 * c] = (relu_mask && !(x2[u].x > 0.f)) ? 0.f : g[u].x;
                    r[u][2 * c + 1] = (relu_mask && !(x2[u].y > 0.f)) ? 0.f : g[u].y;
                    am = absmax_acc(am, r[u][2 * c]);
                    am = absmax_acc(am, r[u][2 * c + 1]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = vb + 8 * u + vl;
                if (dx && v < vcnt) {
                    T* dp = dx + ((size_t)n * S + s0 + v) * dx_ld + sl * 8;
                    if constexpr (sizeof(T) == 2) {
                        Vec<T, 8>::store(dp, r[u]);
                    } else {
                        Vec<T, 4>::store(dp, *reinterpret_cast<float(*)[4]>(r[u]));
                        Vec<T, 4>::store(dp + 4, *reinterpret_cast<float(*)[4]>(r[u] + 4));
                    }
                }
            }
        }
        __syncthreads();                                 // the stage is free for the load issued by the next iteration
        if (++stage == HBT_NST) stage = 0;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    absmax_flush(am, absmax);
    // reduce over the lanes that share a slice (the voxel bits of the lane index); the bias gradient is taken from slice 0's lanes
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float a = (c & 1) ? aw[j][c >> 1].y : aw[j][c >> 1].x;
            for (int o = 16; o >= 4; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (vl == 0) atomicAdd(&red[j * Cin + sl * 8 + c], a);
        }
        float b = ab[j];
        for (int o = 16; o >= 4; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) atomicAdd(&red[COUT * Cin + j], b);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < COUT * Cin + COUT; i += blockDim.x) {
        if (i < COUT * Cin) atomicAdd(dw + i, red[i]);
        else if (db) atomicAdd(db + (i - COUT * Cin), red[i]);
    }
}

template <typename T>
static bool head_mid_ok(const void* x, int64_t x_ld, const void* dx, int64_t dx_ld, int Cin, int Cout) {
    constexpr int V = FullVec<T>::value;
    return (Cin == 32 || Cin == 64 || Cin == 128) && (Cout == 3 || Cout == 4 || Cout == 6 || Cout == 8 || Cout == 12) && x_ld % V == 0 &&
           aligned16(x) && (!dx || (dx_ld % V == 0 && aligned16(dx)));
}

template <typename T>
static bool head_small_ok(const void* x, int64_t x_ld, const void* dx, int64_t dx_ld, int Cin, int Cout) {
    constexpr int V = FullVec<T>::value;
    return (Cin == 16 || Cin == 32) && (Cout == 1 || Cout == 2) && x_ld % V == 0 && aligned16(x) && (!dx || (dx_ld % V == 0 && aligned16(dx)));
}
// the backward kernel also takes Cin = 64 / 128 as 2 / 4 channel slices of 32
template <typename T>
static bool head_bwd_small_ok(const void* x, int64_t x_ld, const void* dx, int64_t dx_ld, int Cin, int Cout) {
    return head_small_ok<T>(x, x_ld, dx, dx_ld, (Cin == 64 || Cin == 128) ? 32 : Cin, Cout);
}

}  // namespace b200em

using namespace b200em;

template <typename T>
static void launch_head_bwd(unsigned blocks, cudaStream_t st, const float* grad_out, const float* out, const void* x,
                            int64_t x_ld, const float* w, void* dx, int64_t dx_ld, float* dw, float* db, int64_t S, int Cin,
                            int Cout, int co0, int cob, int act, int relu_mask, int64_t total) {
    constexpr int V = FullVec<T>::value;
    const bool vec = Cin % V == 0 && x_ld % V == 0 && aligned16(x) && (!dx || (dx_ld % V == 0 && aligned16(dx)));
    const int maxp = (cob * Cin + HB_TILE - 1) / HB_TILE;
    const int acc = co0 > 0 ? 1 : 0;
#define B2_HEAD_BWD(VV, MP)                                                                                          \
    head_bwd_kernel<T, VV, MP><<<blocks, HB_TILE, 0, st>>>(grad_out, out, (const T*)x, x_ld, w, (T*)dx, dx_ld, dw, db, S, \
                                                           Cin, Cout, co0, cob, act, acc, relu_mask, total)
    if (vec) {
        if (maxp <= 4) B2_HEAD_BWD(V, 4);
        else if (maxp <= 16) B2_HEAD_BWD(V, 16);
        else B2_HEAD_BWD(V, 64);
    } else {
        if (maxp <= 4) B2_HEAD_BWD(1, 4);
        else if (maxp <= 16) B2_HEAD_BWD(1, 16);
        else B2_HEAD_BWD(1, 64);
    }
#undef B2_HEAD_BWD
}


template <typename T>
static void launch_head_fwd_small(unsigned blocks, cudaStream_t st, const void* x, int64_t x_ld, const float* w, const float* bias,
                                  float* out, int64_t S, int Cin, int Cout, int act, int64_t total) {
#define B2_HEAD_FWD_SMALL(CI, CO) head_fwd_small_kernel<T, CI, CO><<<blocks, 256, 0, st>>>((const T*)x, x_ld, w, bias, out, S, act, total)
    if (Cin == 32 && Cout == 2) B2_HEAD_FWD_SMALL(32, 2);
    else if (Cin == 32) B2_HEAD_FWD_SMALL(32, 1);
    else if (Cout == 2) B2_HEAD_FWD_SMALL(16, 2);
    else B2_HEAD_FWD_SMALL(16, 1);
#undef B2_HEAD_FWD_SMALL
}

template <typename T>
static void launch_head_bwd_small(unsigned blocks, cudaStream_t st, const float* grad_out, const float* out, const void* x,
                                  int64_t x_ld, const float* w, void* dx, int64_t dx_ld, float* dw, float* db, int64_t S, int Cin,
                                  int Cout, int act, int relu_mask, int64_t total, float* absmax) {
#define B2_HEAD_BWD_SMALL(CI, CO)                                                                                             \
    head_bwd_small_kernel<T, CI, CO><<<blocks, 256, 0, st>>>(grad_out, out, (const T*)x, x_ld, w, (T*)dx, dx_ld, dw, db, S, act, \
                                                             relu_mask, total, Cin, absmax)
    if (Cin % 32 == 0 && Cout == 2) B2_HEAD_BWD_SMALL(32, 2);
    else if (Cin % 32 == 0) B2_HEAD_BWD_SMALL(32, 1);
    else if (Cout == 2) B2_HEAD_BWD_SMALL(16, 2);
    else B2_HEAD_BWD_SMALL(16, 1);
#undef B2_HEAD_BWD_SMALL
}

template <typename T>
static void launch_head_fwd_mid(int64_t blocks, cudaStream_t st, const void* x, int64_t x_ld, const float* w, const float* bias, float* out,
                                int64_t S, int Cin, int Cout, int act, int64_t total) {
    static const bool tiled_env = [] { const char* e = getenv("B200EM_HEAD_TILED"); return !(e && atoi(e) == 0); }();
    // (measured on (2, 64, 256, 256) x 32 channels, bf16: 12 outputs 548 -> 422 us; 8 outputs 298 vs 307, 4 outputs 155 vs 200: the group form keeps those)
    if (tiled_env && Cin == HBT_CIN && Cout > 8 && aligned16(w) && total / S * ((S + HBT_TV - 1) / HBT_TV) < (1LL << 31)) {
        const int N = (int)(total / S);
        const long long ntiles = (long long)N * ((S + HBT_TV - 1) / HBT_TV);
        const unsigned blt = (unsigned)(ntiles < 4LL * sm_count() ? ntiles : 4LL * sm_count());
#define B2_HEAD_FWD_TILED(CO)                                                                                                         \
    do {                                                                                                                              \
        const int smem = HBT_NST * HftSmem<T>::stage_bytes + CO * HBT_CIN * 4;                                                         \
        cudaFuncSetAttribute(head_fwd_mid_tiled_kernel<T, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                     \
        head_fwd_mid_tiled_kernel<T, CO><<<blt, 128, smem, st>>>((const T*)x, x_ld, w, bias, out, S, act, N);                          \
    } while (0)
        if (Cout == 12) { B2_HEAD_FWD_TILED(12); return; }
#undef B2_HEAD_FWD_TILED
    }
    // warps = voxel groups x output-channel groups of 3 (or 4): a block count divisible by 3 and 4 keeps a warp's group fixed
    unsigned bl = (unsigned)(((blocks * (Cout % 3 == 0 ? Cout / 3 : Cout / 4)) + 11) / 12 * 12);
    const unsigned cap_ = (unsigned)sm_count() * 16 / 12 * 12;
    if (bl > cap_) bl = cap_;
#define B2_HEAD_FWD_MID(CO, JG) head_fwd_mid_kernel<T, CO, JG><<<bl, 256, 0, st>>>((const T*)x, x_ld, w, bias, out, S, Cin, act, total)
    if (Cout <= 2) {
        bl = (unsigned)(blocks < (int64_t)sm_count() * 16 ? blocks : (int64_t)sm_count() * 16);
        if (Cout == 2) B2_HEAD_FWD_MID(2, 2); else B2_HEAD_FWD_MID(1, 1);
    } else if (Cout == 12) B2_HEAD_FWD_MID(12, 3);
    else if (Cout == 6) B2_HEAD_FWD_MID(6, 3);
    else if (Cout == 3) B2_HEAD_FWD_MID(3, 3);
    else if (Cout == 8) B2_HEAD_FWD_MID(8, 4);
    else B2_HEAD_FWD_MID(4, 4);
#undef B2_HEAD_FWD_MID
}

template <typename T>
static void launch_head_bwd_mid(cudaStream_t st, const float* grad_out, const float* out, const void* x, int64_t x_ld, const float* w, void* dx,
                                int64_t dx_ld, float* dw, float* db, int N, int64_t S, int Cin, int Cout, int act, int relu_mask, int64_t total,
                                float* absmax) {
    // shared-memory staged form: 32 input channels, planar rows that can be copied in 16-byte pieces
    static const bool tiled_env = [] { const char* e = getenv("B200EM_HEAD_TILED"); return !(e && atoi(e) == 0); }();
    if (tiled_env && Cin == HBT_CIN && S % 4 == 0 && aligned16(grad_out) && aligned16(out) && (N * ((S + HBT_TV - 1) / HBT_TV)) < (1LL << 31)) {
        const long long ntiles = (long long)N * ((S + HBT_TV - 1) / HBT_TV);
        const unsigned bl = (unsigned)(ntiles < 2LL * sm_count() ? ntiles : 2LL * sm_count());
#define B2_HEAD_BWD_TILED(CO)                                                                                                              \
    do {                                                                                                                                   \
        cudaFuncSetAttribute(head_bwd_mid_tiled_kernel<T, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, HbtSmem<T, CO>::total);         \
        head_bwd_mid_tiled_kernel<T, CO><<<bl, 128, HbtSmem<T, CO>::total, st>>>(grad_out, out, (const T*)x, x_ld, w, (T*)dx, dx_ld, dw, db, S, \
                                                                                  act, relu_mask, N, absmax);                             \
    } while (0)
        if (Cout == 12) B2_HEAD_BWD_TILED(12);
        else if (Cout == 8) B2_HEAD_BWD_TILED(8);
        else if (Cout == 6) B2_HEAD_BWD_TILED(6);
        else if (Cout == 4) B2_HEAD_BWD_TILED(4);
        else B2_HEAD_BWD_TILED(3);
#undef B2_HEAD_BWD_TILED
        return;
    }
    const unsigned bl = (unsigned)sm_count() * (Cout > 8 ? 2 : 3);      // resident blocks of 128 threads per SM, grid-stride
#define B2_HEAD_BWD_MID(CO) \
    head_bwd_mid_kernel<T, CO><<<bl, 128, 0, st>>>(grad_out, out, (const T*)x, x_ld, w, (T*)dx, dx_ld, dw, db, S, act, relu_mask, total, Cin, absmax)
    if (Cout == 12) B2_HEAD_BWD_MID(12);
    else if (Cout == 8) B2_HEAD_BWD_MID(8);
    else if (Cout == 6) B2_HEAD_BWD_MID(6);
    else if (Cout == 4) B2_HEAD_BWD_MID(4);
    else B2_HEAD_BWD_MID(3);
#undef B2_HEAD_BWD_MID
}

extern "C" {

int b200em_head_fwd(const void* x, int64_t x_ld, int dtype, const float* w, const float* bias, float* out, int N,
                    int64_t S, int Cin, int Cout, int act, void* stream) {
    B2_CHECK_ARG(x && w && out && N > 0 && S > 0 && Cin > 0 && Cout > 0 && x_ld >= Cin, "head_fwd: bad arguments");
    B2_CHECK_ARG(act >= 0 && act <= 3, "head_fwd: unknown activation code %d", act);
    int64_t total = (int64_t)N * S;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        if (head_small_ok<T>(x, x_ld, nullptr, 0, Cin, Cout)) {
            launch_head_fwd_small<T>((unsigned)blocks, (cudaStream_t)stream, x, x_ld, w, bias, out, S, Cin, Cout, act, total);
        } else if (head_mid_ok<T>(x, x_ld, nullptr, 0, Cin, Cout) || head_bwd_small_ok<T>(x, x_ld, nullptr, 0, Cin, Cout)) {
            // (the second condition: 1 / 2 output channels on 64 / 128 input channels -- too wide for the register-resident filter
            // of the small kernel, served by the shared-memory-filter kernel with one channel group)
            launch_head_fwd_mid<T>(blocks, (cudaStream_t)stream, x, x_ld, w, bias, out, S, Cin, Cout, act, total);
        } else if (Cin % V == 0 && x_ld % V == 0 && aligned16(x))
            head_fwd_kernel<T, V><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, w, bias, out, S, Cin, Cout, act, total);
        else
            head_fwd_kernel<T, 1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, w, bias, out, S, Cin, Cout, act, total);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_head_bwd(const float* grad_out, const float* out, const void* x, int64_t x_ld, int dtype, const float* w,
                    void* dx, int64_t dx_ld, float* dw, float* db, int N, int64_t S, int Cin, int Cout, int act,
                    int relu_mask, float* absmax, void* stream) {
    B2_CHECK_ARG(grad_out && out && x && w && dw && N > 0 && S > 0 && Cin > 0 && Cout > 0 && x_ld >= Cin, "head_bwd: bad arguments");
    B2_CHECK_ARG(!dx || dx_ld >= Cin, "head_bwd: dx pitch smaller than Cin");
    B2_CHECK_ARG(act >= 0 && act <= 3, "head_bwd: unknown activation code %d", act);
    B2_CHECK_ARG(Cin <= 512, "head_bwd: Cin %d > 512 not supported", Cin);
    int64_t total = (int64_t)N * S;
    int64_t ntiles = (total + HB_TILE - 1) / HB_TILE;
    int64_t blocks = (int64_t)sm_count() * 4;
    if (blocks > ntiles) blocks = ntiles;
    {
        bool done = false;
        B2_DISPATCH_DTYPE(dtype, T, {
            if (head_mid_ok<T>(x, x_ld, dx, dx_ld, Cin, Cout)) {
                launch_head_bwd_mid<T>((cudaStream_t)stream, grad_out, out, x, x_ld, w, dx, dx_ld, dw, db, N, S, Cin, Cout, act, relu_mask, total, absmax);
                done = true;
            } else if (head_bwd_small_ok<T>(x, x_ld, dx, dx_ld, Cin, Cout)) {
                launch_head_bwd_small<T>((unsigned)(blocks / 4 * 4 > 0 ? blocks / 4 * 4 : 4), (cudaStream_t)stream, grad_out, out, x, x_ld, w, dx, dx_ld, dw, db, S, Cin, Cout,
                                         act, relu_mask, total, absmax);
                done = true;
            }
        })
        if (done) {
            B2_LAUNCH_CHECK();
            return 0;
        }
    }
    for (int co0 = 0; co0 < Cout; co0 += HB_MAXCO) {
        int cob = Cout - co0 < HB_MAXCO ? Cout - co0 : HB_MAXCO;
        B2_DISPATCH_DTYPE(dtype, T, {
            launch_head_bwd<T>((unsigned)blocks, (cudaStream_t)stream, grad_out, out, x, x_ld, w, dx, dx_ld, dw, db, S, Cin, Cout,
                               co0, cob, act, relu_mask, total);
        })
        B2_LAUNCH_CHECK();
    }
    // the generic kernel does not track max |dx|: one more pass where the caller asked for it (fp32 heads wider than 128 channels)
    if (absmax && dx && dtype == B200EM_F32) return b200em_absmax_f32((const float*)dx, dx_ld, total, Cin, absmax, nullptr, stream);
    return 0;
}

}  // extern "C"
