// Fused Dice-family segmentation losses beyond plain DiceLoss: one forward pass that reads prediction, target (and mask)
// once, one backward pass that reads them once and writes the gradient.
//
// Reference restated (torch_em/loss):
//   DiceLossWithLogits       dice.py:136-173    dice_score(sigmoid(x), t)
//   BCEDiceLoss              dice.py:176-214    alpha * dice(p, t) + beta * F.binary_cross_entropy(p, t)
//   BCEDiceLossWithLogits    dice.py:217-256    alpha * dice(sigmoid(x), t) + beta * F.binary_cross_entropy_with_logits(x, t)
//   DistanceLoss             distance_based.py:7-57   Dice on channel 0 + MSE on channels 1, 2, the latter two optionally
//   DiceBasedDistanceLoss    distance_based.py:60-69  multiplied by the foreground target (channel 0 of the target)
//
// Per channel c the forward accumulates, over (n, voxels), with p = logits ? sigmoid(x) : x, m = channel uses the mask ?
// mask : 1, pm = p*m, tm = t*m:
//   sums[c] = (sum pm*tm, sum pm^2, sum tm^2, sum bce(p or x, t), sum (pm - tm)^2)
// The per-channel configuration chan[c] = (w_dice, w_bce, w_mse, use_mask) selects which terms a channel contributes:
//   loss = reduce_c w_dice[c] * (1 - 2 num_c / max(den_c, eps)) + sum_c w_bce[c] * bce_c / numel + sum_c w_mse[c] * mse_c / numel
// BCE follows ATen exactly: log clamped at -100 (binary_cross_entropy), gradient (p - t) / max(p (1-p), 1e-12); with logits
// max(x, 0) - x t + log1p(exp(-|x|)), gradient sigmoid(x) - t.
#include "common.cuh"

namespace b200em {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

}  // namespace

// grid = (blocks over S, C, N)
template <typename TP>
__global__ void __launch_bounds__(256)
segloss_sums_kernel(const TP* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ mask,
                    int64_t target_nstride, int64_t mask_nstride, int64_t mask_cstride, const float* __restrict__ chan, int logits,
                    int C, int64_t S, float* __restrict__ sums) {
    __shared__ float sh[8][5];
    const int c = blockIdx.y, n = blockIdx.z;
    const bool want_bce = chan[4 * c + 1] != 0.f, want_mse = chan[4 * c + 2] != 0.f;
    const bool use_mask = mask != nullptr && chan[4 * c + 3] != 0.f;
    const TP* p = pred + ((size_t)n * C + c) * S;
    const float* t = target + (size_t)n * target_nstride + (size_t)c * S;
    const float* m = use_mask ? mask + (size_t)n * mask_nstride + (size_t)c * mask_cstride : nullptr;
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = to_f<TP>(p[i]), tv = t[i];
        const float pv = logits ? sigmoidf_(x) : x;
        const float mv = m ? m[i] : 1.f;
        const float pm = pv * mv, tm = tv * mv;
        a[0] = fmaf(pm, tm, a[0]);
        a[1] = fmaf(pm, pm, a[1]);
        a[2] = fmaf(tm, tm, a[2]);
        if (want_bce) {
            if (logits) a[3] += fmaxf(x, 0.f) - x * tv + log1pf(__expf(-fabsf(x)));
            else a[3] -= tv * fmaxf(__logf(pv), -100.f) + (1.f - tv) * fmaxf(__logf(1.f - pv), -100.f);
        }
        if (want_mse) { const float d = pm - tm; a[4] = fmaf(d, d, a[4]); }
    }
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float v = warp_sum(a[k]);
        if (lane == 0) sh[wi][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
        atomicAdd(sums + c * 5 + threadIdx.x, v);
    }
}

// one thread: loss value(s) and backward coefficients coef[c] = (A_c, B_c, bce weight / numel, mse weight / numel)
__global__ void segloss_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ chan, int C, float eps, int channelwise,
                                        int reduce, float numel, float* __restrict__ loss, float* __restrict__ coef) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float extra = 0.f;                      // BCE and MSE terms (always summed over their channels)
    for (int c = 0; c < C; ++c) {
        const float wb = chan[4 * c + 1], wm = chan[4 * c + 2];
        coef[4 * c + 2] = wb / numel;
        coef[4 * c + 3] = wm / numel;
        extra += wb * (sums[c * 5 + 3] / numel) + wm * (sums[c * 5 + 4] / numel);
    }
    if (!channelwise) {
        double num = 0.0, den = 0.0;
        float wd = 0.f;
        for (int c = 0; c < C; ++c) {
            if (chan[4 * c] == 0.f) continue;
            wd = chan[4 * c];
            num += sums[c * 5];
            den += (double)sums[c * 5 + 1] + sums[c * 5 + 2];
        }
        const float numf = (float)num, denf = (float)den;
        const bool live = denf > eps;
        const float dc = live ? denf : eps;
        loss[0] = wd * (1.f - 2.f * (numf / dc)) + extra;
        const float A = -2.f / dc, B = live ? 4.f * numf / (dc * dc) : 0.f;
        for (int c = 0; c < C; ++c) {
            const bool on = chan[4 * c] != 0.f;
            coef[4 * c] = on ? wd * A : 0.f;
            coef[4 * c + 1] = on ? wd * B : 0.f;
        }
        return;
    }
    float acc = 0.f;
    int arg = -1, nd = 0;
    for (int c = 0; c < C; ++c) {
        const float wd = chan[4 * c];
        if (wd == 0.f) {
            coef[4 * c] = coef[4 * c + 1] = 0.f;
            if (reduce == 4) loss[c] = 0.f;
            continue;
        }
        ++nd;
        const float num = sums[c * 5], den = sums[c * 5 + 1] + sums[c * 5 + 2];
        const bool live = den > eps;
        const float dc = live ? den : eps;
        const float l = wd * (1.f - 2.f * (num / dc));
        coef[4 * c] = wd * (-2.f / dc);
        coef[4 * c + 1] = live ? wd * 4.f * num / (dc * dc) : 0.f;
        if (reduce == 4) loss[c] = l;
        else if (reduce == 0 || reduce == 1) acc += l;
        else if (arg < 0) { acc = l; arg = c; }
        else if (reduce == 2 && l > acc) { acc = l; arg = c; }
        else if (reduce == 3 && l < acc) { acc = l; arg = c; }
    }
    if (reduce == 4) return;
    if (reduce == 1 && nd > 0) {
        acc /= (float)nd;
        for (int c = 0; c < C; ++c) { coef[4 * c] /= (float)nd; coef[4 * c + 1] /= (float)nd; }
    } else if (reduce == 2 || reduce == 3) {
        for (int c = 0; c < C; ++c)
            if (c != arg) { coef[4 * c] = 0.f; coef[4 * c + 1] = 0.f; }
    }
    loss[0] = acc + extra;
}

template <typename TP, typename TG>
__global__ void __launch_bounds__(256)
segloss_bwd_kernel(const TP* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ mask,
                   int64_t target_nstride, int64_t mask_nstride, int64_t mask_cstride, const float* __restrict__ chan,
                   const float* __restrict__ coef, const float* __restrict__ gout, int gout_per_channel, int logits,
                   TG* __restrict__ grad, int C, int64_t S) {
    const int c = blockIdx.y, n = blockIdx.z;
    const float go = gout_per_channel ? gout[c] : gout[0];
    const float A = coef[4 * c] * go, B = coef[4 * c + 1] * go, cb = coef[4 * c + 2] * go, cm = coef[4 * c + 3] * go;
    const bool use_mask = mask != nullptr && chan[4 * c + 3] != 0.f;
    const TP* p = pred + ((size_t)n * C + c) * S;
    const float* t = target + (size_t)n * target_nstride + (size_t)c * S;
    const float* m = use_mask ? mask + (size_t)n * mask_nstride + (size_t)c * mask_cstride : nullptr;
    TG* g = grad + ((size_t)n * C + c) * S;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = to_f<TP>(p[i]), tv = t[i];
        const float pv = logits ? sigmoidf_(x) : x;
        const float mv = m ? m[i] : 1.f;
        const float pm = pv * mv, tm = tv * mv;
        float r = (A * tm + B * pm) * mv;                      // d dice / d p
        if (cm != 0.f) r = fmaf(2.f * cm * (pm - tm), mv, r);  // d mse / d p
        if (logits) {
            r *= pv * (1.f - pv);                              // d p / d x
            if (cb != 0.f) r = fmaf(cb, pv - tv, r);           // d bce_with_logits / d x
        } else if (cb != 0.f) {
            r = fmaf(cb, (pv - tv) / fmaxf((1.f - pv) * pv, 1e-12f), r);
        }
        g[i] = from_f<TG>(r);
    }
}

static inline unsigned seg_blocks(int64_t S, int C, int N) {
    int64_t b = (S + 1023) / 1024;
    int64_t cap = (int64_t)sm_count() * 8 / ((int64_t)C * N) + 1;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_segloss_sums(const void* pred, int pred_dtype, const float* target, const float* mask, int64_t target_nstride,
                        int64_t mask_nstride, int64_t mask_cstride, const float* chan, int logits, int N, int C, int64_t S,
                        float* sums, void* stream) {
    B2_CHECK_ARG(pred && target && chan && sums && N > 0 && C > 0 && S > 0, "segloss_sums: bad arguments");
    B2_CHECK_ARG(C <= 65535 && N <= 65535, "segloss_sums: C or N too large for the launch grid");
    dim3 grid(seg_blocks(S, C, N), (unsigned)C, (unsigned)N);
    B2_DISPATCH_DTYPE(pred_dtype, TP, {
        segloss_sums_kernel<TP><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, target, mask, target_nstride, mask_nstride,
                                                                     mask_cstride, chan, logits, C, S, sums);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_segloss_finalize(const float* sums, const float* chan, int C, float eps, int channelwise, int reduce, float numel,
                            float* loss, float* coef, void* stream) {
    B2_CHECK_ARG(sums && chan && loss && coef && C > 0 && numel > 0.f, "segloss_finalize: bad arguments");
    B2_CHECK_ARG(reduce >= 0 && reduce <= 4, "segloss_finalize: unknown channel reduction code %d", reduce);
    segloss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, chan, C, eps, channelwise, reduce, numel, loss, coef);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_segloss_bwd(const void* pred, int pred_dtype, const float* target, const float* mask, int64_t target_nstride,
                       int64_t mask_nstride, int64_t mask_cstride, const float* chan, const float* coef, const float* gout,
                       int gout_per_channel, int logits, void* grad_pred, int grad_dtype, int N, int C, int64_t S, void* stream) {
    B2_CHECK_ARG(pred && target && chan && coef && gout && grad_pred && N > 0 && C > 0 && S > 0, "segloss_bwd: bad arguments");
    B2_CHECK_ARG(C <= 65535 && N <= 65535, "segloss_bwd: C or N too large for the launch grid");
    dim3 grid(seg_blocks(S, C, N), (unsigned)C, (unsigned)N);
    B2_DISPATCH_DTYPE(pred_dtype, TP, {
        B2_DISPATCH_DTYPE(grad_dtype, TG, {
            segloss_bwd_kernel<TP, TG><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, target, mask, target_nstride, mask_nstride,
                                                                            mask_cstride, chan, coef, gout, gout_per_channel, logits,
                                                                            (TG*)grad_pred, C, S);
        })
    })
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
