// Fused DiceLoss [+ ApplyAndRemoveMask("multiply")] forward reductions and backward.
//
// Reference restated: flatten_samples + dice_score + DiceLoss (loss/dice.py:7-31, 34-93, 96-133) and the multiply
// masking of LossWrapper(…, ApplyAndRemoveMask("multiply")) (loss/wrapper.py:84-87, 129-152).  The reference
// materialises two transposed copies and ~8 elementwise/reduce launches; here the forward is ONE pass that reads
// prediction, target (and mask) once, and the backward is ONE pass that reads them once and writes the gradient:
//   per channel c over (n, voxels):  num = sum pm*tm,  den = sum pm^2 + sum tm^2   (pm = p*m, tm = t*m)
//   loss = reduce_c (1 - 2*num/max(den, eps));   dL/dp = (A_c*tm + B_c*pm) * m
//   A_c = -2/den, B_c = 4*num/den^2 when den > eps;  A_c = -2/eps, B_c = 0 below the clamp (clamp has zero slope).
// HBM-bound: algorithmic bytes fwd = (sizeof(p) + 4 [+4]) per element; bwd = the same + 4 written.
#include "common.cuh"

namespace b200em {

template <typename TP>
__device__ __forceinline__ void load4(const TP* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[wi] = v;
    __syncthreads();
    float r = 0.f;
    if (wi == 0) {
        r = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid in warp 0
}

// grid = (blocks over S, C, N)
template <typename TP, bool VEC>
__global__ void __launch_bounds__(256)
dice_sums_kernel(const TP* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ mask,
                 int64_t target_nstride, int C, int64_t S, float* __restrict__ sums) {
    __shared__ float sh[8];
    const int c = blockIdx.y, n = blockIdx.z;
    const TP* p = pred + ((size_t)n * C + c) * S;
    const float* t = target + (size_t)n * target_nstride + (size_t)c * S;
    const float* m = mask ? mask + (size_t)n * target_nstride + (size_t)c * S : nullptr;
    float a_pt = 0.f, a_pp = 0.f, a_tt = 0.f;
    if (VEC) {
        const int64_t S4 = S / 4;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < S4; i += (int64_t)gridDim.x * blockDim.x) {
            float pv[4], tv[4], mv[4];
            load4<TP>(p + 4 * i, pv);
            load4<float>(t + 4 * i, tv);
            if (m) load4<float>(m + 4 * i, mv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float pm = m ? pv[k] * mv[k] : pv[k], tm = m ? tv[k] * mv[k] : tv[k];
                a_pt = fmaf(pm, tm, a_pt);
                a_pp = fmaf(pm, pm, a_pp);
                a_tt = fmaf(tm, tm, a_tt);
            }
        }
    } else {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
            float pv = to_f<TP>(p[i]), tv = t[i];
            float mv = m ? m[i] : 1.f;
            float pm = pv * mv, tm = tv * mv;
            a_pt = fmaf(pm, tm, a_pt);
            a_pp = fmaf(pm, pm, a_pp);
            a_tt = fmaf(tm, tm, a_tt);
        }
    }
    a_pt = block_sum_256(a_pt, sh);
    a_pp = block_sum_256(a_pp, sh);
    a_tt = block_sum_256(a_tt, sh);
    if (threadIdx.x == 0) {
        atomicAdd(sums + c * 3 + 0, a_pt);
        atomicAdd(sums + c * 3 + 1, a_pp);
        atomicAdd(sums + c * 3 + 2, a_tt);
    }
}

// one thread: loss value(s) and backward coefficients
__global__ void dice_finalize_kernel(const float* __restrict__ sums, int C, float eps, int channelwise, int reduce,
                                     float* __restrict__ loss, float* __restrict__ coef) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    if (!channelwise) {
        double num = 0.0, den = 0.0;
        for (int c = 0; c < C; ++c) {
            num += sums[c * 3];
            den += (double)sums[c * 3 + 1] + sums[c * 3 + 2];
        }
        float numf = (float)num, denf = (float)den;
        bool live = denf > eps;
        float dc = live ? denf : eps;
        loss[0] = 1.f - 2.f * (numf / dc);
        float A = -2.f / dc, B = live ? 4.f * numf / (dc * dc) : 0.f;
        for (int c = 0; c < C; ++c) { coef[2 * c] = A; coef[2 * c + 1] = B; }
        return;
    }
    float acc = 0.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
        float num = sums[c * 3], den = sums[c * 3 + 1] + sums[c * 3 + 2];
        bool live = den > eps;
        float dc = live ? den : eps;
        float l = 1.f - 2.f * (num / dc);
        coef[2 * c] = -2.f / dc;
        coef[2 * c + 1] = live ? 4.f * num / (dc * dc) : 0.f;
        if (reduce == 4) loss[c] = l;
        else if (reduce == 0 || reduce == 1) acc += l;
        else if (c == 0) { acc = l; arg = 0; }
        else if (reduce == 2 && l > acc) { acc = l; arg = c; }
        else if (reduce == 3 && l < acc) { acc = l; arg = c; }
    }
    if (reduce == 4) return;
    if (reduce == 1) {
        acc /= (float)C;
        for (int c = 0; c < C; ++c) { coef[2 * c] /= (float)C; coef[2 * c + 1] /= (float)C; }
    } else if (reduce == 2 || reduce == 3) {
        // torch.max / torch.min over channels: gradient flows to the (first) arg-extremum only
        for (int c = 0; c < C; ++c)
            if (c != arg) { coef[2 * c] = 0.f; coef[2 * c + 1] = 0.f; }
    }
    loss[0] = acc;
}

template <typename TP, typename TG, bool VEC>
__global__ void __launch_bounds__(256)
dice_bwd_kernel(const TP* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ mask,
                int64_t target_nstride, const float* __restrict__ coef, const float* __restrict__ gout,
                int gout_per_channel, TG* __restrict__ grad, int C, int64_t S) {
    const int c = blockIdx.y, n = blockIdx.z;
    const float go = gout_per_channel ? gout[c] : gout[0];
    const float A = coef[2 * c] * go, B = coef[2 * c + 1] * go;
    const TP* p = pred + ((size_t)n * C + c) * S;
    const float* t = target + (size_t)n * target_nstride + (size_t)c * S;
    const float* m = mask ? mask + (size_t)n * target_nstride + (size_t)c * S : nullptr;
    TG* g = grad + ((size_t)n * C + c) * S;
    if (VEC) {
        const int64_t S4 = S / 4;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < S4; i += (int64_t)gridDim.x * blockDim.x) {
            float pv[4], tv[4], mv[4], r[4];
            load4<TP>(p + 4 * i, pv);
            load4<float>(t + 4 * i, tv);
            if (m) load4<float>(m + 4 * i, mv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float mm = m ? mv[k] * mv[k] : 1.f;
                r[k] = (A * tv[k] + B * pv[k]) * mm;
            }
            if (sizeof(TG) == 4) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(g) + 4 * i) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
                uint2 o;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
                h[0] = __floats2bfloat162_rn(r[0], r[1]);
                h[1] = __floats2bfloat162_rn(r[2], r[3]);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g) + 4 * i) = o;
            }
        }
    } else {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
            float mv = m ? m[i] : 1.f;
            g[i] = from_f<TG>((A * t[i] + B * to_f<TP>(p[i])) * mv * mv);
        }
    }
}

static inline unsigned slab_blocks(int64_t S, int C, int N, int per_thread) {
    int64_t b = (S / per_thread + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8 / ((int64_t)C * N) + 1;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_dice_sums(const void* pred, int pred_dtype, const float* target, const float* mask, int64_t target_nstride,
                     int N, int C, int64_t S, float* sums, void* stream) {
    B2_CHECK_ARG(pred && target && sums && N > 0 && C > 0 && S > 0, "dice_sums: bad arguments");
    B2_CHECK_ARG(C <= 65535 && N <= 65535, "dice_sums: C or N too large for the launch grid");
    dim3 grid(slab_blocks(S, C, N, 4), (unsigned)C, (unsigned)N);
    const bool vec = S % 4 == 0 && target_nstride % 4 == 0 && aligned16(pred) && aligned16(target) && (!mask || aligned16(mask));
    B2_DISPATCH_DTYPE(pred_dtype, TP, {
        if (vec) dice_sums_kernel<TP, true><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, target, mask, target_nstride, C, S, sums);
        else dice_sums_kernel<TP, false><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, target, mask, target_nstride, C, S, sums);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_dice_finalize(const float* sums, int C, float eps, int channelwise, int reduce, float* loss, float* coef,
                         void* stream) {
    B2_CHECK_ARG(sums && loss && coef && C > 0, "dice_finalize: bad arguments");
    B2_CHECK_ARG(reduce >= 0 && reduce <= 4, "dice_finalize: unknown channel reduction code %d", reduce);
    dice_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, C, eps, channelwise, reduce, loss, coef);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_dice_bwd(const void* pred, int pred_dtype, const float* target, const float* mask, int64_t target_nstride,
                    const float* coef, const float* gout, int gout_per_channel, void* grad_pred, int grad_dtype, int N,
                    int C, int64_t S, void* stream) {
    B2_CHECK_ARG(pred && target && coef && gout && grad_pred && N > 0 && C > 0 && S > 0, "dice_bwd: bad arguments");
    B2_CHECK_ARG(C <= 65535 && N <= 65535, "dice_bwd: C or N too large for the launch grid");
    dim3 grid(slab_blocks(S, C, N, 4), (unsigned)C, (unsigned)N);
    const bool vec = S % 4 == 0 && target_nstride % 4 == 0 && aligned16(pred) && aligned16(target) && (!mask || aligned16(mask)) &&
                     aligned16(grad_pred);
    B2_DISPATCH_DTYPE(pred_dtype, TP, {
        B2_DISPATCH_DTYPE(grad_dtype, TG, {
            if (vec) dice_bwd_kernel<TP, TG, true><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, target, mask, target_nstride, coef, gout, gout_per_channel, (TG*)grad_pred, C, S);
            else dice_bwd_kernel<TP, TG, false><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, target, mask, target_nstride, coef, gout, gout_per_channel, (TG*)grad_pred, C, S);
        })
    })
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
