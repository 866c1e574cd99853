// Instance labels -> affinity / boundary training targets, and the fused "target + masked Dice" reductions.
//
// Reference restated: AffinityTransform.__call__ (transform/label.py:290-327; arithmetic pinned by the brute force in
// test/transform/test_label_transforms.py:5-55) and BoundaryTransform (label.py:100-129, find_boundaries "thick").
// In the reference these run on CPU inside the dataloader and the 2C-channel fp32 target is shipped host->device
// every step (805 MB at cfg3); here they are integer stencils over the int64 label volume already on the GPU, and
// the fused variants never write the target tensor at all (HBM traffic: labels + prediction only).
//
//   q = p + offset_c;  oob = q outside the volume
//   no ignore label :  disaff = oob ? 1 : [lab[p] != lab[q]];                         mask = !oob
//   ignore label L  :  k = [lab[p]==L] + [lab[q]==L]
//                      oob or k==2 or (k==1 and !include_ignore_transitions) -> disaff = 1, mask = 0
//                      k==1 and include_ignore_transitions                   -> disaff = 1, mask = 1
//                      else                                                   -> disaff = [lab[p]!=lab[q]], mask = 1
#include "common.cuh"

namespace b200em {

constexpr int MAX_OFFSETS = 64;
struct OffsetTable {
    int n;
    int d[MAX_OFFSETS], h[MAX_OFFSETS], w[MAX_OFFSETS];
};

struct AffRule {
    int has_ignore;
    long long ignore_label;
    int include_ignore_transitions;
};

// The neighbour of voxel (d, h, w) at an offset: whether it lies inside the volume, and its linear index (the voxel's own index s
// when it does not, so that the label load below is unconditional and always in bounds).
__device__ __forceinline__ bool aff_neighbour(int d, int h, int w, int D, int H, int W, int od, int oh, int ow, unsigned s, unsigned& q) {
    const int qd = d + od, qh = h + oh, qw = w + ow;
    const bool inb = qd >= 0 && qd < D && qh >= 0 && qh < H && qw >= 0 && qw < W;
    q = inb ? ((unsigned)qd * (unsigned)H + (unsigned)qh) * (unsigned)W + (unsigned)qw : s;
    return inb;
}

// AffinityTransform's rule (label.py:277-327 / the brute force of test_label_transforms.py:5-55) on the two labels, branch-free:
// the loads of several voxels can then be issued together instead of one memory latency per voxel and channel.
__device__ __forceinline__ void aff_rule(long long lp, long long lq, bool inb, const AffRule& r, float& disaff, float& msk) {
    const int k = r.has_ignore ? (int)(lp == r.ignore_label) + (int)(lq == r.ignore_label) : 0;
    const bool off = !inb || k == 2 || (k == 1 && !r.include_ignore_transitions);     // masked out: (1, 0)
    const bool trans = k == 1 && r.include_ignore_transitions;                          // transition to the ignore label: (1, 1)
    disaff = (off || trans || lp != lq) ? 1.f : 0.f;
    msk = off ? 0.f : 1.f;
}

__device__ __forceinline__ void aff_eval(const long long* __restrict__ lab, long long lp, int d, int h, int w, int D, int H,
                                         int W, int od, int oh, int ow, const AffRule& r, float& disaff, float& msk) {
    unsigned q;
    const bool inb = aff_neighbour(d, h, w, D, H, W, od, oh, ow, ((unsigned)d * (unsigned)H + (unsigned)h) * (unsigned)W + (unsigned)w, q);
    aff_rule(lp, lab[q], inb, r, disaff, msk);
}

// out (N, channels, D,H,W): [fg?][n disaff][fg-mask?][n masks]
__global__ void __launch_bounds__(256)
affinity_targets_kernel(const long long* __restrict__ labels, float* __restrict__ out, int D, int H, int W, OffsetTable offs,
                        AffRule rule, int add_binary_target, int add_mask) {
    const int n = blockIdx.y;
    const int64_t S = (int64_t)D * H * W;
    const int noff = offs.n;
    const int channels = (noff + (add_binary_target ? 1 : 0)) * (add_mask ? 2 : 1);
    const long long* lab = labels + (size_t)n * S;
    float* o = out + (size_t)n * channels * S;
    const int c_aff = add_binary_target ? 1 : 0;
    const int c_mask = c_aff + noff + (add_binary_target ? 1 : 0);
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
        const long long lp = lab[s];
        if (add_binary_target) {
            o[s] = lp != 0 ? 1.f : 0.f;
            if (add_mask) o[(size_t)(c_aff + noff) * S + s] = (rule.has_ignore && lp == rule.ignore_label) ? 0.f : 1.f;
        }
        for (int c = 0; c < noff; ++c) {
            float a, m;
            aff_eval(lab, lp, d, h, w, D, H, W, offs.d[c], offs.h[c], offs.w[c], rule, a, m);
            o[(size_t)(c_aff + c) * S + s] = a;
            if (add_mask) o[(size_t)(c_mask + c) * S + s] = m;
        }
    }
}

// out (N, 1|2, D,H,W): [foreground?][boundary]; boundary = some in-bounds 6-neighbour carries a different label
__global__ void __launch_bounds__(256)
boundary_targets_kernel(const long long* __restrict__ labels, float* __restrict__ out, int D, int H, int W,
                        int add_binary_target) {
    const int n = blockIdx.y;
    const int64_t S = (int64_t)D * H * W;
    const long long* lab = labels + (size_t)n * S;
    float* o = out + (size_t)n * (add_binary_target ? 2 : 1) * S;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
        const long long lp = lab[s];
        bool b = false;
        if (w > 0) b |= lab[s - 1] != lp;
        if (w < W - 1) b |= lab[s + 1] != lp;
        if (h > 0) b |= lab[s - W] != lp;
        if (h < H - 1) b |= lab[s + W] != lp;
        if (d > 0) b |= lab[s - (int64_t)W * H] != lp;
        if (d < D - 1) b |= lab[s + (int64_t)W * H] != lp;
        if (add_binary_target) {
            o[s] = lp != 0 ? 1.f : 0.f;
            o[S + s] = b ? 1.f : 0.f;
        } else {
            o[s] = b ? 1.f : 0.f;
        }
    }
}

// Boundary targets with a masked transition class (label.py:133-244).  A voxel is a boundary of image f iff some in-bounds
// 6-neighbour q has f(q) != f(p) (find_boundaries "thick" on f).
//   mode 1  NoToBackgroundBoundaryTransform(bg_label, mask_label):   boundary of [lab != bg_label]  -> mask_label
//   mode 2  BoundaryTransformWithIgnoreLabel(ignore_label = aux):    boundary of [lab == aux]       -> aux
//   elsewhere the ordinary label boundary (0 / 1).  Optional channel 0: [lab != bg] with lab == aux -> aux.
// out (N, 1|2, D,H,W) fp32 holding -1 / 0 / 1 style values (the reference returns int8; targets are consumed as floats).
__global__ void __launch_bounds__(256)
boundary_targets_masked_kernel(const long long* __restrict__ labels, float* __restrict__ out, int D, int H, int W,
                               int add_binary_target, int mode, long long aux, long long bg) {
    const int n = blockIdx.y;
    const int64_t S = (int64_t)D * H * W;
    const long long* lab = labels + (size_t)n * S;
    float* o = out + (size_t)n * (add_binary_target ? 2 : 1) * S;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
        const long long lp = lab[s];
        const bool fp = mode == 1 ? lp != bg : lp == aux;      // the binary image whose boundaries are masked
        bool b = false, bm = false;
        auto visit = [&](long long lq) {
            b |= lq != lp;
            bm |= (mode == 1 ? lq != bg : lq == aux) != fp;
        };
        if (w > 0) visit(lab[s - 1]);
        if (w < W - 1) visit(lab[s + 1]);
        if (h > 0) visit(lab[s - W]);
        if (h < H - 1) visit(lab[s + W]);
        if (d > 0) visit(lab[s - (int64_t)W * H]);
        if (d < D - 1) visit(lab[s + (int64_t)W * H]);
        const float bv = bm ? (float)aux : (b ? 1.f : 0.f);
        if (add_binary_target) {
            const float bin = lp == aux ? (float)aux : (lp != (mode == 1 ? bg : 0) ? 1.f : 0.f);
            o[s] = bin;
            o[S + s] = bv;
        } else {
            o[s] = bv;
        }
    }
}

// OneHotTransform (label.py:330-353): out (N, n_classes, S) fp32, out[n][k][s] = [lab[n][s] == class_ids[k]]
__global__ void __launch_bounds__(256)
one_hot_kernel(const long long* __restrict__ labels, const long long* __restrict__ class_ids, int n_classes, int64_t S,
               float* __restrict__ out) {
    const int n = blockIdx.y;
    const long long* lab = labels + (size_t)n * S;
    float* o = out + (size_t)n * n_classes * S;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const long long lp = lab[s];
        for (int k = 0; k < n_classes; ++k) o[(size_t)k * S + s] = lp == class_ids[k] ? 1.f : 0.f;
    }
}

// segmentation_to_affinities (loss/affinity_side_loss.py:70-89): AFFINITIES (1 = same segment) against the segment at
// p + offset with REPLICATION at the border (the neighbour index is clamped into the volume), no masks.
__global__ void __launch_bounds__(256)
segmentation_affinities_kernel(const long long* __restrict__ labels, float* __restrict__ out, int D, int H, int W,
                               OffsetTable offs) {
    const int n = blockIdx.y;
    const int64_t S = (int64_t)D * H * W;
    const long long* lab = labels + (size_t)n * S;
    float* o = out + (size_t)n * offs.n * S;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
        const long long lp = lab[s];
        for (int c = 0; c < offs.n; ++c) {
            const int qd = min(max(d + offs.d[c], 0), D - 1), qh = min(max(h + offs.h[c], 0), H - 1), qw = min(max(w + offs.w[c], 0), W - 1);
            o[(size_t)c * S + s] = lab[((size_t)qd * H + qh) * W + qw] == lp ? 1.f : 0.f;
        }
    }
}

constexpr int AD_PER_THREAD = 8;                       // voxels per thread
constexpr int AD_CHUNK = 256 * AD_PER_THREAD;          // voxels per block

// Fused: Dice sums of pred (N, n_off, S) against AffinityTransform(offsets, add_mask=True) targets, target and mask
// computed on the fly.  grid = (chunks, N).  sums[c][3] += (sum pm*tm, sum pm^2, sum tm^2).
template <typename TP>
__global__ void __launch_bounds__(256)
affinity_dice_sums_kernel(const TP* __restrict__ pred, const long long* __restrict__ labels, int D, int H, int W,
                          OffsetTable offs, AffRule rule, float* __restrict__ sums) {
    __shared__ float sh[3][8];
    const int n = blockIdx.y;
    const unsigned S = (unsigned)D * H * W;                 // 32-bit voxel indices within a sample (checked by the launcher)
    const long long* lab = labels + (size_t)n * S;
    const int noff = offs.n;
    long long lp[AD_PER_THREAD];
    int pd[AD_PER_THREAD], ph[AD_PER_THREAD], pw[AD_PER_THREAD];
    const unsigned base = blockIdx.x * (unsigned)AD_CHUNK + threadIdx.x;
#pragma unroll
    for (int k = 0; k < AD_PER_THREAD; ++k) {
        const unsigned s = base + (unsigned)k * 256u;
        if (s < S) {
            lp[k] = lab[s];
            pw[k] = (int)(s % (unsigned)W); ph[k] = (int)((s / (unsigned)W) % (unsigned)H); pd[k] = (int)(s / ((unsigned)W * (unsigned)H));
        } else {
            lp[k] = 0; pw[k] = ph[k] = pd[k] = 0;
        }
    }
    for (int c = 0; c < noff; ++c) {
        const TP* p = pred + ((size_t)n * noff + c) * S;
        float a_pt = 0.f, a_pp = 0.f, a_tt = 0.f;
        // all loads of the thread's voxels first (16 independent requests), then the arithmetic
        const int od = offs.d[c], oh = offs.h[c], ow = offs.w[c];
        long long lq[AD_PER_THREAD];
        TP pv[AD_PER_THREAD];
        bool inb[AD_PER_THREAD];
#pragma unroll
        for (int k = 0; k < AD_PER_THREAD; ++k) {
            const unsigned s = base + (unsigned)k * 256u;
            const unsigned ss = s < S ? s : 0u;
            unsigned q;
            inb[k] = aff_neighbour(pd[k], ph[k], pw[k], D, H, W, od, oh, ow, ss, q);
            lq[k] = lab[q];
            pv[k] = p[ss];
        }
#pragma unroll
        for (int k = 0; k < AD_PER_THREAD; ++k) {
            const unsigned s = base + (unsigned)k * 256u;
            if (s < S) {
                float t, m;
                aff_rule(lp[k], lq[k], inb[k], rule, t, m);
                const float pm = to_f<TP>(pv[k]) * m, tm = t * m;
                a_pt = fmaf(pm, tm, a_pt);
                a_pp = fmaf(pm, pm, a_pp);
                a_tt = fmaf(tm, tm, a_tt);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a_pt += __shfl_xor_sync(0xffffffffu, a_pt, o);
            a_pp += __shfl_xor_sync(0xffffffffu, a_pp, o);
            a_tt += __shfl_xor_sync(0xffffffffu, a_tt, o);
        }
        const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();
        if (lane == 0) { sh[0][wi] = a_pt; sh[1][wi] = a_pp; sh[2][wi] = a_tt; }
        __syncthreads();
        if (threadIdx.x < 3) {
            float r = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) r += sh[threadIdx.x][j];
            atomicAdd(sums + c * 3 + threadIdx.x, r);
        }
    }
}

template <typename TP, typename TG>
__global__ void __launch_bounds__(256)
affinity_dice_bwd_kernel(const TP* __restrict__ pred, const long long* __restrict__ labels, int D, int H, int W,
                         OffsetTable offs, AffRule rule, const float* __restrict__ coef, const float* __restrict__ gout,
                         TG* __restrict__ grad) {
    const int n = blockIdx.y;
    const unsigned S = (unsigned)D * H * W;                 // 32-bit voxel indices within a sample (checked by the launcher)
    const long long* lab = labels + (size_t)n * S;
    const int noff = offs.n;
    const float go = gout[0];
    long long lp[AD_PER_THREAD];
    int pd[AD_PER_THREAD], ph[AD_PER_THREAD], pw[AD_PER_THREAD];
    const unsigned base = blockIdx.x * (unsigned)AD_CHUNK + threadIdx.x;
#pragma unroll
    for (int k = 0; k < AD_PER_THREAD; ++k) {
        const unsigned s = base + (unsigned)k * 256u;
        if (s < S) {
            lp[k] = lab[s];
            pw[k] = (int)(s % (unsigned)W); ph[k] = (int)((s / (unsigned)W) % (unsigned)H); pd[k] = (int)(s / ((unsigned)W * (unsigned)H));
        } else {
            lp[k] = 0; pw[k] = ph[k] = pd[k] = 0;
        }
    }
    for (int c = 0; c < noff; ++c) {
        const float A = coef[2 * c] * go, B = coef[2 * c + 1] * go;
        const TP* p = pred + ((size_t)n * noff + c) * S;
        TG* g = grad + ((size_t)n * noff + c) * S;
        const int od = offs.d[c], oh = offs.h[c], ow = offs.w[c];
        long long lq[AD_PER_THREAD];
        TP pv[AD_PER_THREAD];
        bool inb[AD_PER_THREAD];
#pragma unroll
        for (int k = 0; k < AD_PER_THREAD; ++k) {          // all loads first (16 independent requests per thread)
            const unsigned s = base + (unsigned)k * 256u;
            const unsigned ss = s < S ? s : 0u;
            unsigned q;
            inb[k] = aff_neighbour(pd[k], ph[k], pw[k], D, H, W, od, oh, ow, ss, q);
            lq[k] = lab[q];
            pv[k] = p[ss];
        }
#pragma unroll
        for (int k = 0; k < AD_PER_THREAD; ++k) {
            const unsigned s = base + (unsigned)k * 256u;
            if (s < S) {
                float t, m;
                aff_rule(lp[k], lq[k], inb[k], rule, t, m);
                g[s] = from_f<TG>((A * t + B * to_f<TP>(pv[k])) * m * m);
            }
        }
    }
}

static int fill_offsets(OffsetTable& t, const int* offsets, int n_off) {
    if (!offsets || n_off <= 0 || n_off > MAX_OFFSETS) {
        set_error("affinity: need 1..%d offsets, got %d", MAX_OFFSETS, n_off);
        return 1;
    }
    t.n = n_off;
    for (int i = 0; i < n_off; ++i) { t.d[i] = offsets[3 * i]; t.h[i] = offsets[3 * i + 1]; t.w[i] = offsets[3 * i + 2]; }
    return 0;
}

static inline unsigned flat_blocks(int64_t S, int N) {
    int64_t b = (S + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16 / (N > 0 ? N : 1) + 1;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_affinity_targets(const int64_t* labels, float* out, int N, int D, int H, int W, const int* offsets, int n_off,
                            int has_ignore, int64_t ignore_label, int add_binary_target, int add_mask,
                            int include_ignore_transitions, void* stream) {
    B2_CHECK_ARG(labels && out && N > 0 && D > 0 && H > 0 && W > 0, "affinity_targets: bad arguments");
    OffsetTable t;
    if (fill_offsets(t, offsets, n_off)) return 1;
    AffRule r{has_ignore, (long long)ignore_label, include_ignore_transitions};
    dim3 grid(flat_blocks((int64_t)D * H * W, N), (unsigned)N);
    affinity_targets_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, out, D, H, W, t, r, add_binary_target, add_mask);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_boundary_targets(const int64_t* labels, float* out, int N, int D, int H, int W, int add_binary_target,
                            void* stream) {
    B2_CHECK_ARG(labels && out && N > 0 && D > 0 && H > 0 && W > 0, "boundary_targets: bad arguments");
    dim3 grid(flat_blocks((int64_t)D * H * W, N), (unsigned)N);
    boundary_targets_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, out, D, H, W, add_binary_target);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_boundary_targets_masked(const int64_t* labels, float* out, int N, int D, int H, int W, int add_binary_target,
                                   int mode, int64_t aux_label, int64_t bg_label, void* stream) {
    B2_CHECK_ARG(labels && out && N > 0 && D > 0 && H > 0 && W > 0, "boundary_targets_masked: bad arguments");
    B2_CHECK_ARG(mode == 1 || mode == 2, "boundary_targets_masked: mode must be 1 (no-to-background) or 2 (ignore label)");
    dim3 grid(flat_blocks((int64_t)D * H * W, N), (unsigned)N);
    boundary_targets_masked_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, out, D, H, W, add_binary_target,
                                                                            mode, (long long)aux_label, (long long)bg_label);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_one_hot(const int64_t* labels, const int64_t* class_ids, int n_classes, float* out, int N, int64_t S, void* stream) {
    B2_CHECK_ARG(labels && class_ids && out && N > 0 && S > 0 && n_classes > 0, "one_hot: bad arguments");
    dim3 grid(flat_blocks(S, N), (unsigned)N);
    one_hot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, (const long long*)class_ids, n_classes, S, out);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_segmentation_affinities(const int64_t* labels, float* out, int N, int D, int H, int W, const int* offsets, int n_off,
                                   void* stream) {
    B2_CHECK_ARG(labels && out && N > 0 && D > 0 && H > 0 && W > 0, "segmentation_affinities: bad arguments");
    OffsetTable t;
    if (fill_offsets(t, offsets, n_off)) return 1;
    dim3 grid(flat_blocks((int64_t)D * H * W, N), (unsigned)N);
    segmentation_affinities_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, out, D, H, W, t);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_affinity_dice_sums(const void* pred, int pred_dtype, const int64_t* labels, int N, int D, int H, int W,
                              const int* offsets, int n_off, int has_ignore, int64_t ignore_label,
                              int include_ignore_transitions, float* sums, void* stream) {
    B2_CHECK_ARG(pred && labels && sums && N > 0 && D > 0 && H > 0 && W > 0, "affinity_dice_sums: bad arguments");
    OffsetTable t;
    if (fill_offsets(t, offsets, n_off)) return 1;
    AffRule r{has_ignore, (long long)ignore_label, include_ignore_transitions};
    int64_t S = (int64_t)D * H * W;
    B2_CHECK_ARG(S < (1LL << 31), "affinity_dice_sums: sample too large for 32-bit voxel indices");
    dim3 grid((unsigned)((S + AD_CHUNK - 1) / AD_CHUNK), (unsigned)N);
    B2_DISPATCH_DTYPE(pred_dtype, TP, {
        affinity_dice_sums_kernel<TP><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, (const long long*)labels, D, H, W, t, r, sums);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_affinity_dice_bwd(const void* pred, int pred_dtype, const int64_t* labels, int N, int D, int H, int W,
                             const int* offsets, int n_off, int has_ignore, int64_t ignore_label,
                             int include_ignore_transitions, const float* coef, const float* gout, void* grad_pred,
                             int grad_dtype, void* stream) {
    B2_CHECK_ARG(pred && labels && coef && gout && grad_pred && N > 0 && D > 0 && H > 0 && W > 0, "affinity_dice_bwd: bad arguments");
    OffsetTable t;
    if (fill_offsets(t, offsets, n_off)) return 1;
    AffRule r{has_ignore, (long long)ignore_label, include_ignore_transitions};
    int64_t S = (int64_t)D * H * W;
    B2_CHECK_ARG(S < (1LL << 31), "affinity_dice_bwd: sample too large for 32-bit voxel indices");
    dim3 grid((unsigned)((S + AD_CHUNK - 1) / AD_CHUNK), (unsigned)N);
    B2_DISPATCH_DTYPE(pred_dtype, TP, {
        B2_DISPATCH_DTYPE(grad_dtype, TG, {
            affinity_dice_bwd_kernel<TP, TG><<<grid, 256, 0, (cudaStream_t)stream>>>((const TP*)pred, (const long long*)labels, D, H, W, t, r, coef, gout, (TG*)grad_pred);
        })
    })
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
