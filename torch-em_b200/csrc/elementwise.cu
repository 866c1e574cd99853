// Memory-bound NDHWC kernels of the U-Net block plumbing: layout change, per-channel statistics, the
// InstanceNorm/GroupNorm apply and backward, max-pool, trilinear upsample.  All are HBM-roofline kernels:
// one pass over each operand, 16-byte vector accesses along the channel axis, fp32 math.
//
// Reference semantics restated (see include/b200em.h for the per-entry citations):
//   InstanceNorm3d / GroupNorm   unet.py:391-406      MaxPool3d   unet.py:645,316
//   F.interpolate(trilinear)     unet.py:456          ReLU'       unet.py:433,437
#include <algorithm>

#include "common.cuh"

namespace b200em {

// Thread layout shared by all kernels here: blockDim = (bx, by); x walks channel vectors (contiguous in
// memory -> coalesced), y walks voxels.  blockIdx.y = sample, blockIdx.x strides over voxel rows.
struct Launch2D {
    dim3 grid, block;
};
static Launch2D make_launch(int cvec, int64_t rows, int N, int max_blocks_per_sample = 0) {
    int bx = cvec < 256 ? cvec : 256;
    int by = 256 / bx;
    if (by < 1) by = 1;
    int64_t need = (rows + by - 1) / by;
    int64_t cap = max_blocks_per_sample > 0 ? max_blocks_per_sample : (int64_t)sm_count() * 8 / (N > 0 ? N : 1) + 1;
    if (need > cap) need = cap;
    if (need < 1) need = 1;
    Launch2D l;
    l.grid = dim3((unsigned)need, (unsigned)N, 1);
    l.block = dim3(bx, by, 1);
    return l;
}

template <int VEC, int K>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[VEC][K], float* out_nc, int cv, bool active) {
    __shared__ float red[256 * VEC * K];
    const int bx = blockDim.x, by = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
    float* mine = red + (ty * bx + tx) * (VEC * K);
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int k = 0; k < K; ++k) mine[v * K + k] = acc[v][k];
    __syncthreads();
    if (ty == 0 && active) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float s = 0.f;
                for (int j = 0; j < by; ++j) s += red[(j * bx + tx) * (VEC * K) + v * K + k];
                atomicAdd(out_nc + (size_t)(cv * VEC + v) * K + k, s);
            }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// layout: NCDHW fp32 -> NDHWC
template <typename T>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int64_t y_ld, int C, int64_t S,
                                      int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = i % S;
        int64_t nc = i / S;
        int c = (int)(nc % C);
        int64_t n = nc / C;
        y[(n * S + s) * y_ld + c] = from_f<T>(x[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-(n,c) sums: (sum x, sum x^2)  and  (sum g, sum g*x)
template <typename T, int VEC, bool DOT>
__global__ void channel_sums_kernel(const T* __restrict__ a, int64_t a_ld, const T* __restrict__ b, int64_t b_ld,
                                    int64_t S, int C, float* __restrict__ sums) {
    const int n = blockIdx.y, cvec = C / VEC;
    const T* an = a + (size_t)n * S * a_ld;
    const T* bn = DOT ? b + (size_t)n * S * b_ld : nullptr;
    for (int cvb = 0; cvb < cvec; cvb += blockDim.x) {
        const int cv = cvb + threadIdx.x;
        const bool active = cv < cvec;
        float acc[VEC][2];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v][0] = acc[v][1] = 0.f;
        if (active) {
            for (unsigned s = blockIdx.x * blockDim.y + threadIdx.y; s < (unsigned)S; s += gridDim.x * blockDim.y) {
                float va[VEC];
                Vec<T, VEC>::load(an + (size_t)s * a_ld + cv * VEC, va);
                if (DOT) {
                    float vb[VEC];
                    Vec<T, VEC>::load(bn + (size_t)s * b_ld + cv * VEC, vb);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { acc[v][0] += va[v]; acc[v][1] += va[v] * vb[v]; }
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { acc[v][0] += va[v]; acc[v][1] += va[v] * va[v]; }
                }
            }
        }
        block_channel_reduce<VEC, 2>(acc, sums + (size_t)n * C * 2, cv, active);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// norm finalize: one thread per (n, group)
__global__ void norm_finalize_kernel(const float* __restrict__ sums, int N, int C, float S, int groups,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                     float* __restrict__ scale_shift, float* __restrict__ mean_rstd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * groups) return;
    int n = i / groups, g = i % groups;
    int cpg = C / groups;
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
        int c = g * cpg + j;
        s1 += sums[((size_t)n * C + c) * 2];
        s2 += sums[((size_t)n * C + c) * 2 + 1];
    }
    double cnt = (double)S * cpg;
    double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    for (int j = 0; j < cpg; ++j) {
        int c = g * cpg + j;
        float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
        size_t o = ((size_t)n * C + c) * 2;
        scale_shift[o] = rstd * ga;
        scale_shift[o + 1] = be - (float)mean * rstd * ga;
        mean_rstd[o] = (float)mean;
        mean_rstd[o + 1] = rstd;
    }
}

// norm backward finalize: one thread per (n, group).
//   x_hat = (x - mean) * rstd;  dxh = gamma * g;  m1 = mean_grp(dxh), m2 = mean_grp(dxh * x_hat)
//   dx = rstd * (dxh - m1 - x_hat * m2) = coef0 * g + coef1 * x + coef2
__global__ void norm_bwd_finalize_kernel(const float* __restrict__ dsums, const float* __restrict__ mean_rstd,
                                         const float* __restrict__ gamma, int N, int C, float S, int groups,
                                         float* __restrict__ coef, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * groups) return;
    int n = i / groups, g = i % groups;
    int cpg = C / groups;
    size_t base = (size_t)n * C + (size_t)g * cpg;
    float mean = mean_rstd[base * 2], rstd = mean_rstd[base * 2 + 1];
    double a1 = 0.0, a2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
        int c = g * cpg + j;
        float ga = gamma ? gamma[c] : 1.f;
        float sg = dsums[(base + j) * 2], sgx = dsums[(base + j) * 2 + 1];
        float sgxh = rstd * (sgx - mean * sg);  // sum g * x_hat
        a1 += (double)ga * sg;
        a2 += (double)ga * sgxh;
        if (dgamma) atomicAdd(dgamma + c, sgxh);
        if (dbeta) atomicAdd(dbeta + c, sg);
    }
    double cnt = (double)S * cpg;
    float m1 = (float)(a1 / cnt), m2 = (float)(a2 / cnt);
    for (int j = 0; j < cpg; ++j) {
        int c = g * cpg + j;
        float ga = gamma ? gamma[c] : 1.f;
        size_t o = (base + j) * 3;
        coef[o] = rstd * ga;
        coef[o + 1] = -rstd * rstd * m2;
        coef[o + 2] = rstd * (mean * rstd * m2 - m1);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// y = scale*x + shift
template <typename T, int VEC>
__global__ void affine_apply_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ ss,
                                    T* __restrict__ y, int64_t y_ld, int64_t S, int C, int64_t total) {
    const unsigned cvec = C / VEC;
    const int64_t n = blockIdx.y;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {
        int cv = (int)(i % cvec);
        int64_t vox = n * S + i / cvec;
        float v[VEC];
        Vec<T, VEC>::load(x + vox * x_ld + cv * VEC, v);
        const float* p = ss + ((size_t)n * C + cv * VEC) * 2;
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = fmaf(v[k], p[2 * k], p[2 * k + 1]);
        Vec<T, VEC>::store(y + vox * y_ld + cv * VEC, v);
    }
}

// out = (c0*g + c1*x + c2 [+ add]) * (relu ? x>0 : 1)
template <typename T, int VEC>
__global__ void norm_bwd_apply_kernel(const T* __restrict__ g, int64_t g_ld, const T* __restrict__ x, int64_t x_ld,
                                      const float* __restrict__ coef, const T* __restrict__ add, int64_t add_ld,
                                      T* __restrict__ out, int64_t out_ld, int64_t S, int C, int relu_mask,
                                      int64_t total, float* __restrict__ absmax) {
    const unsigned cvec = C / VEC;
    const int64_t n = blockIdx.y;
    unsigned am = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {
        int cv = (int)(i % cvec);
        int64_t vox = n * S + i / cvec;
        float vg[VEC], vx[VEC], va[VEC], r[VEC];
        Vec<T, VEC>::load(g + vox * g_ld + cv * VEC, vg);
        const bool need_x = (coef != nullptr) || relu_mask;
        if (need_x) Vec<T, VEC>::load(x + vox * x_ld + cv * VEC, vx);
        if (add) Vec<T, VEC>::load(add + vox * add_ld + cv * VEC, va);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float t = vg[k];
            if (coef) {
                const float* p = coef + ((size_t)n * C + cv * VEC + k) * 3;
                t = fmaf(p[0], vg[k], fmaf(p[1], vx[k], p[2]));
            }
            if (add) t += va[k];
            if (relu_mask && !(vx[k] > 0.f)) t = 0.f;
            r[k] = t;
            am = absmax_acc(am, t);
        }
        Vec<T, VEC>::store(out + vox * out_ld + cv * VEC, r);
    }
    absmax_flush(am, absmax);
}

// fp32 -> (hi, lo) bf16 split of x_hat = scale*x + shift: hi = bf16(x_hat), lo = bf16(x_hat - hi)
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ ss, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, int64_t S, int C) {
    const int64_t n = blockIdx.y;
    const unsigned cvec = C / 4;
    const int64_t total = S * cvec;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % cvec);
        const int64_t vox = n * S + i / cvec;
        const float4 v = *reinterpret_cast<const float4*>(x + vox * x_ld + cv * 4);
        float f[4] = {v.x, v.y, v.z, v.w};
        if (ss) {
            const float* q = ss + ((size_t)n * C + cv * 4) * 2;
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = fmaf(f[e], q[2 * e], q[2 * e + 1]);
        }
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            h[e] = __float2bfloat16_rn(f[e]);
            l[e] = __float2bfloat16_rn(f[e] - __bfloat162float(h[e]));
        }
        *reinterpret_cast<uint2*>(hi + vox * C + cv * 4) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + vox * C + cv * 4) = *reinterpret_cast<const uint2*>(l);
    }
}

// max |x| of an fp32 tensor (rows x C, pitch x_ld): one atomicMax per block on the bit pattern (non-negative floats order like
// unsigned integers; a NaN ends up above +inf and switches the h16 scaling off, common.cuh h16_shift) -- and, in the same pass,
// the per-channel sums (the bias gradient, taken from the fp32 values like autograd does, not from the fp16 copies).
// Block = (bx channel vectors, by rows); a thread owns channel vectors tx, tx + bx, ... (at most 4).
__global__ void __launch_bounds__(256)
absmax_f32_kernel(const float* __restrict__ x, int64_t x_ld, int64_t rows, int C, unsigned* __restrict__ out, float* __restrict__ colsum) {
    const int cvec = C / 4, bx = blockDim.x, by = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
    unsigned m = 0;
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t row = (int64_t)blockIdx.x * by + ty; row < rows; row += (int64_t)gridDim.x * by) {
        const float* xr = x + row * x_ld;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cv = tx + k * bx;
            if (cv < cvec) {
                const float4 v = *reinterpret_cast<const float4*>(xr + cv * 4);
                m = max(max(m, __float_as_uint(v.x) & 0x7fffffffu), __float_as_uint(v.y) & 0x7fffffffu);
                m = max(max(m, __float_as_uint(v.z) & 0x7fffffffu), __float_as_uint(v.w) & 0x7fffffffu);
                acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
            }
        }
    }
    const int tid = ty * bx + tx;
    __shared__ unsigned s_m;
    extern __shared__ float s_col[];                       // [C] when colsum
    if (tid == 0) s_m = 0u;
    if (colsum)
        for (int i = tid; i < C; i += bx * by) s_col[i] = 0.f;
    __syncthreads();
    if (m) atomicMax(&s_m, m);
    if (colsum) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cv = tx + k * bx;
            if (cv < cvec) {
                atomicAdd(&s_col[cv * 4], acc[k].x); atomicAdd(&s_col[cv * 4 + 1], acc[k].y);
                atomicAdd(&s_col[cv * 4 + 2], acc[k].z); atomicAdd(&s_col[cv * 4 + 3], acc[k].w);
            }
        }
    }
    __syncthreads();
    if (tid == 0 && s_m) atomicMax(out, s_m);
    if (colsum)
        for (int i = tid; i < C; i += bx * by) atomicAdd(colsum + i, s_col[i]);
}

// fp32 -> fp16 operand copy of the h16 path: out = fp16(2^k * (scale*x + shift)), k from the device absmax (common.cuh).
// Block = (bx 8-channel vectors, by voxels): a thread keeps its channels over the loop, so the per-channel sums of x (the bias
// gradient sum(dz), wanted from the fp32 values like autograd computes it) accumulate in registers in the same pass.
__global__ void __launch_bounds__(256)
cvt_f16_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ ss, const float* __restrict__ absmax,
               __half* __restrict__ out, float* __restrict__ colsum, int64_t S, int C) {
    const int64_t n = blockIdx.y;
    const int cvec = C / 8, bx = blockDim.x, by = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
    const float mul = pow2i(h16_shift(absmax));
    extern __shared__ float s_col[];                   // [C] when colsum
    if (colsum) {
        for (int i = ty * bx + tx; i < C; i += bx * by) s_col[i] = 0.f;
        __syncthreads();
    }
    for (int cv = tx; cv < cvec; cv += bx) {
        float sc[8], sh[8], acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc[e] = 1.f; sh[e] = 0.f; acc[e] = 0.f; }
        if (ss) {
            const float* q = ss + ((size_t)n * C + cv * 8) * 2;
#pragma unroll
            for (int e = 0; e < 8; ++e) { sc[e] = q[2 * e]; sh[e] = q[2 * e + 1]; }
        }
        const int64_t vstep = (int64_t)gridDim.x * by;
        for (int64_t v0 = (int64_t)blockIdx.x * by + ty; v0 < S; v0 += 2 * vstep) {
            // two voxels per iteration: four 16-byte loads in flight per thread
            const int64_t v1 = v0 + vstep;
            const bool two = v1 < S;
            const float* p0 = x + (n * S + v0) * x_ld + cv * 8;
            const float* p1 = x + (n * S + (two ? v1 : v0)) * x_ld + cv * 8;
            const float4 a0 = *reinterpret_cast<const float4*>(p0), b0 = *reinterpret_cast<const float4*>(p0 + 4);
            const float4 a1 = *reinterpret_cast<const float4*>(p1), b1 = *reinterpret_cast<const float4*>(p1 + 4);
            float f[2][8] = {{a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, b0.z, b0.w}, {a1.x, a1.y, a1.z, a1.w, b1.x, b1.y, b1.z, b1.w}};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
#pragma unroll
                for (int e = 0; e < 8; ++e) { acc[e] += f[u][e]; f[u][e] = fmaf(f[u][e], sc[e], sh[e]); }
                uint4 o;
                __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[u][2 * e] * mul, f[u][2 * e + 1] * mul);
                *reinterpret_cast<uint4*>(out + (n * S + (u ? v1 : v0)) * C + cv * 8) = o;
            }
        }
        if (colsum) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_col[cv * 8 + e], acc[e]);
        }
    }
    if (colsum) {
        __syncthreads();
        for (int i = ty * bx + tx; i < C; i += bx * by) atomicAdd(colsum + i, s_col[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// max pool forward (+ statistics of the pooled output)
// F2 = 1: factor (2,2,2) or (1,2,2) known at compile time (FDC = depth factor): the window loads are unrolled and in flight together
template <typename T, int VEC, int F2 = 0, int FDC = 0>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, int64_t x_ld, T* __restrict__ y, int64_t y_ld, int D, int H,
                                   int W, int C, int fd_, int fh_, int fw_, float* __restrict__ sums) {
    const int fd = F2 ? FDC : fd_, fh = F2 ? 2 : fh_, fw = F2 ? 2 : fw_;
    const int n = blockIdx.y, cvec = C / VEC;
    const int Do = D / fd, Ho = H / fh, Wo = W / fw;
    const int64_t So = (int64_t)Do * Ho * Wo;
    const T* xn = x + (size_t)n * D * H * W * x_ld;
    T* yn = y + (size_t)n * So * y_ld;
    for (int cvb = 0; cvb < cvec; cvb += blockDim.x) {
        const int cv = cvb + threadIdx.x;
        const bool active = cv < cvec;
        float acc[VEC][2];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v][0] = acc[v][1] = 0.f;
        if (active) {
            for (unsigned s = blockIdx.x * blockDim.y + threadIdx.y; s < (unsigned)So; s += gridDim.x * blockDim.y) {
                int wo = (int)(s % (unsigned)Wo), ho = (int)((s / (unsigned)Wo) % (unsigned)Ho), d_o = (int)(s / (unsigned)(Wo * Ho));
                float m[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) m[v] = -INFINITY;
                if (F2) {
                    float t[FDC ? FDC * 4 : 1][VEC];
#pragma unroll
                    for (int q = 0; q < FDC * 4; ++q) {
                        const int a = q >> 2, b = (q >> 1) & 1, c = q & 1;
                        const int64_t vi = ((int64_t)(d_o * FDC + a) * H + (ho * 2 + b)) * W + (wo * 2 + c);
                        Vec<T, VEC>::load(xn + vi * x_ld + cv * VEC, t[q]);
                    }
#pragma unroll
                    for (int q = 0; q < FDC * 4; ++q)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) m[v] = fmaxf(m[v], t[q][v]);
                } else {
                    for (int a = 0; a < fd; ++a)
                        for (int b = 0; b < fh; ++b)
                            for (int c = 0; c < fw; ++c) {
                                int64_t vi = ((int64_t)(d_o * fd + a) * H + (ho * fh + b)) * W + (wo * fw + c);
                                float t[VEC];
                                Vec<T, VEC>::load(xn + vi * x_ld + cv * VEC, t);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) m[v] = fmaxf(m[v], t[v]);
                            }
                }
                Vec<T, VEC>::store(yn + (size_t)s * y_ld + cv * VEC, m);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { acc[v][0] += m[v]; acc[v][1] += m[v] * m[v]; }
            }
        }
        if (sums) block_channel_reduce<VEC, 2>(acc, sums + (size_t)n * C * 2, cv, active);
    }
}

// max pool backward: one thread per (window, channel vector); first maximum in (d,h,w) scan order wins
// (ATen max_pool3d_with_indices semantics).
// coef (nullable, per (n, c): c0, c1, c2 with sample stride coef_nstride): the term added to every voxel is then
// c0 * add + c1 * x + c2 -- the norm backward of the consuming decoder block applied to its raw data gradient `add` -- instead of
// `add` itself, so the skip gradient is never materialised.
template <typename T, int VEC>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, int64_t x_ld, const T* __restrict__ dp, int64_t dp_ld,
                                   const T* __restrict__ add, int64_t add_ld, const float* __restrict__ coef, int64_t coef_nstride,
                                   T* __restrict__ out, int64_t out_ld,
                                   int D, int H, int W, int C, int fd, int fh, int fw, int relu_mask, int64_t total, float* __restrict__ absmax) {
    unsigned am = 0;
    const unsigned cvec = C / VEC;
    const int Do = D / fd, Ho = H / fh, Wo = W / fw;
    const int64_t So = (int64_t)Do * Ho * Wo, Si = (int64_t)D * H * W;
    const int64_t n = blockIdx.y;
    extern __shared__ __align__(16) float s_coef[];     // [channel vector][c0 | c1 | c2][VEC] of this sample when coef is given
    if (coef) {
        for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
            const int c = i / 3, k = i % 3;
            s_coef[((c / VEC) * 3 + k) * VEC + c % VEC] = coef[n * coef_nstride + i];
        }
        __syncthreads();
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {
        int cv = (int)(i % cvec);
        const unsigned s = i / cvec;
        int64_t vox = n * So + s;
        int wo = (int)(s % (unsigned)Wo), ho = (int)((s / (unsigned)Wo) % (unsigned)Ho), d_o = (int)(s / (unsigned)(Wo * Ho));
        float m[VEC];
        int arg[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { m[v] = -INFINITY; arg[v] = 0; }
        int pos = 0;
        for (int a = 0; a < fd; ++a)
            for (int b = 0; b < fh; ++b)
                for (int c = 0; c < fw; ++c, ++pos) {
                    int64_t vi = n * Si + ((int64_t)(d_o * fd + a) * H + (ho * fh + b)) * W + (wo * fw + c);
                    float t[VEC];
                    Vec<T, VEC>::load(x + vi * x_ld + cv * VEC, t);
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        if (t[v] > m[v] || pos == 0) { m[v] = t[v]; arg[v] = pos; }
                }
        float gp[VEC];
        Vec<T, VEC>::load(dp + vox * dp_ld + cv * VEC, gp);
        pos = 0;
        for (int a = 0; a < fd; ++a)
            for (int b = 0; b < fh; ++b)
                for (int c = 0; c < fw; ++c, ++pos) {
                    int64_t vi = n * Si + ((int64_t)(d_o * fd + a) * H + (ho * fh + b)) * W + (wo * fw + c);
                    float r[VEC], t[VEC], va[VEC];
                    if (relu_mask || coef) Vec<T, VEC>::load(x + vi * x_ld + cv * VEC, t);
                    if (add) Vec<T, VEC>::load(add + vi * add_ld + cv * VEC, va);
                    if (coef) {
                        // shared memory, read as 16-byte vectors: no extra registers across the window loop, no global loads in it
                        float k0[VEC], k1[VEC], k2[VEC];
                        Vec<float, VEC == 1 ? 1 : 4>::load(s_coef + (cv * 3 + 0) * VEC, *reinterpret_cast<float(*)[VEC == 1 ? 1 : 4]>(k0));
                        Vec<float, VEC == 1 ? 1 : 4>::load(s_coef + (cv * 3 + 1) * VEC, *reinterpret_cast<float(*)[VEC == 1 ? 1 : 4]>(k1));
                        Vec<float, VEC == 1 ? 1 : 4>::load(s_coef + (cv * 3 + 2) * VEC, *reinterpret_cast<float(*)[VEC == 1 ? 1 : 4]>(k2));
                        if (VEC == 8) {
                            Vec<float, 4>::load(s_coef + (cv * 3 + 0) * VEC + 4, *reinterpret_cast<float(*)[4]>(k0 + (VEC == 8 ? 4 : 0)));
                            Vec<float, 4>::load(s_coef + (cv * 3 + 1) * VEC + 4, *reinterpret_cast<float(*)[4]>(k1 + (VEC == 8 ? 4 : 0)));
                            Vec<float, 4>::load(s_coef + (cv * 3 + 2) * VEC + 4, *reinterpret_cast<float(*)[4]>(k2 + (VEC == 8 ? 4 : 0)));
                        }
#pragma unroll
                        for (int v = 0; v < VEC; ++v) va[v] = fmaf(k0[v], va[v], fmaf(k1[v], t[v], k2[v]));
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        float q = (arg[v] == pos) ? gp[v] : 0.f;
                        if (add) q += va[v];
                        if (relu_mask && !(t[v] > 0.f)) q = 0.f;
                        r[v] = q;
                        am = absmax_acc(am, q);
                    }
                    Vec<T, VEC>::store(out + vi * out_ld + cv * VEC, r);
                }
    }
    absmax_flush(am, absmax);
}

// Max-pool backward for the windows the U-Net uses, (2|1, 2, 2), known at compile time: one thread per (window, channel vector)
// with every load of the window issued up front (the x values are kept as raw 16-byte words for the second pass), 32-bit byte
// offsets from block-uniform bases and an unrolled window loop.  Same semantics as the generic kernel above.
template <typename T, int VEC>
__device__ __forceinline__ void unpack16(const uint4& r, float (&v)[VEC]) {
    if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
    } else {
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
        v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
        v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
    }
}

template <typename T, int VEC, int FD>
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const T* __restrict__ x, int64_t x_ld, const T* __restrict__ dp, int64_t dp_ld, const T* __restrict__ add, int64_t add_ld,
                    const float* __restrict__ coef, int64_t coef_nstride, T* __restrict__ out, int64_t out_ld, int D, int H, int W, int C,
                    int relu_mask, unsigned total, float* __restrict__ absmax) {
    static_assert(VEC * sizeof(T) == 16, "16-byte channel vectors");
    unsigned am = 0;
    constexpr int NW = FD * 4;
    const unsigned cvec = C / VEC;
    const int Ho = H / 2, Wo = W / 2;
    const int n = blockIdx.y;
    extern __shared__ __align__(16) float s_coef[];     // [channel vector][c0 | c1 | c2][VEC] of this sample when coef is given
    if (coef) {
        for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
            const int c = i / 3, k = i % 3;
            s_coef[((c / VEC) * 3 + k) * VEC + c % VEC] = coef[(size_t)n * coef_nstride + i];
        }
        __syncthreads();
    }
    const size_t Si = (size_t)D * H * W, So = (size_t)(D / FD) * Ho * Wo;
    const char* xb = reinterpret_cast<const char*>(x + (size_t)n * Si * x_ld);
    const char* ab = reinterpret_cast<const char*>(add ? add + (size_t)n * Si * add_ld : nullptr);
    const char* pb = reinterpret_cast<const char*>(dp + (size_t)n * So * dp_ld);
    char* ob = reinterpret_cast<char*>(out + (size_t)n * Si * out_ld);
    const unsigned xl = (unsigned)x_ld * sizeof(T), al = (unsigned)add_ld * sizeof(T), ol = (unsigned)out_ld * sizeof(T), pl = (unsigned)dp_ld * sizeof(T);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned cv = i % cvec, s = i / cvec;
        const unsigned wo = s % (unsigned)Wo, ho = (s / (unsigned)Wo) % (unsigned)Ho, d_o = s / (unsigned)(Wo * Ho);
        const unsigned cvo = cv * 16u;
        unsigned vox[NW];                            // voxel index of the window positions, (a, b, c) scan order
#pragma unroll
        for (int q = 0; q < NW; ++q) vox[q] = ((d_o * FD + (q >> 2)) * (unsigned)H + (ho * 2 + ((q >> 1) & 1))) * (unsigned)W + (wo * 2 + (q & 1));
        uint4 xr[NW], ar[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) xr[q] = *reinterpret_cast<const uint4*>(xb + (vox[q] * xl + cvo));
        if (add) {
#pragma unroll
            for (int q = 0; q < NW; ++q) ar[q] = *reinterpret_cast<const uint4*>(ab + (vox[q] * al + cvo));
        }
        float gp[VEC];
        { const uint4 r = *reinterpret_cast<const uint4*>(pb + (s * pl + cvo)); unpack16<T, VEC>(r, gp); }
        float m[VEC];
        int arg[VEC];
        unpack16<T, VEC>(xr[0], m);
#pragma unroll
        for (int v = 0; v < VEC; ++v) arg[v] = 0;
#pragma unroll
        for (int q = 1; q < NW; ++q) {
            float t[VEC];
            unpack16<T, VEC>(xr[q], t);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (t[v] > m[v]) { m[v] = t[v]; arg[v] = q; }
        }
        float k0[VEC], k1[VEC], k2[VEC];
        if (coef) {
#pragma unroll
            for (int v = 0; v < VEC; v += 4) {
                *reinterpret_cast<float4*>(k0 + v) = *reinterpret_cast<const float4*>(s_coef + (cv * 3 + 0) * VEC + v);
                *reinterpret_cast<float4*>(k1 + v) = *reinterpret_cast<const float4*>(s_coef + (cv * 3 + 1) * VEC + v);
                *reinterpret_cast<float4*>(k2 + v) = *reinterpret_cast<const float4*>(s_coef + (cv * 3 + 2) * VEC + v);
            }
        }
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            float t[VEC], va[VEC], r[VEC];
            unpack16<T, VEC>(xr[q], t);
            if (add) unpack16<T, VEC>(ar[q], va);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float g = (arg[v] == q) ? gp[v] : 0.f;
                if (add) g += coef ? fmaf(k0[v], va[v], fmaf(k1[v], t[v], k2[v])) : va[v];
                if (relu_mask && !(t[v] > 0.f)) g = 0.f;
                r[v] = g;
                am = absmax_acc(am, g);
            }
            Vec<T, VEC>::store(reinterpret_cast<T*>(ob + (vox[q] * ol + cvo)), r);
        }
    }
    absmax_flush(am, absmax);
}

// ---------------------------------------------------------------------------------------------------------------
// trilinear, align_corners=False, integer scale f: src = max(0,(o+0.5)/f-0.5); i0=floor(src); i1=min(i0+1,n-1)
__device__ __forceinline__ void lerp_src(int o, int f, int n, int& i0, int& i1, float& lam) {
    if (f == 1) { i0 = i1 = o; lam = 0.f; return; }
    float src = ((float)o + 0.5f) / (float)f - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    i1 = i0 + 1 < n ? i0 + 1 : n - 1;
    lam = src - (float)i0;
}

template <typename T, int VEC>
__global__ void upsample_fwd_kernel(const T* __restrict__ x, int64_t x_ld, T* __restrict__ y, int64_t y_ld, int D, int H,
                                    int W, int C, int fd, int fh, int fw, float* __restrict__ sums) {
    const int n = blockIdx.y, cvec = C / VEC;
    const int Do = D * fd, Ho = H * fh, Wo = W * fw;
    const int64_t So = (int64_t)Do * Ho * Wo;
    const T* xn = x + (size_t)n * D * H * W * x_ld;
    T* yn = y + (size_t)n * So * y_ld;
    for (int cvb = 0; cvb < cvec; cvb += blockDim.x) {
        const int cv = cvb + threadIdx.x;
        const bool active = cv < cvec;
        float acc[VEC][2];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v][0] = acc[v][1] = 0.f;
        if (active) {
            for (unsigned s = blockIdx.x * blockDim.y + threadIdx.y; s < (unsigned)So; s += gridDim.x * blockDim.y) {
                int wo = (int)(s % (unsigned)Wo), ho = (int)((s / (unsigned)Wo) % (unsigned)Ho), d_o = (int)(s / (unsigned)(Wo * Ho));
                int d0, d1, h0, h1, w0, w1;
                float ld, lh, lw;
                lerp_src(d_o, fd, D, d0, d1, ld);
                lerp_src(ho, fh, H, h0, h1, lh);
                lerp_src(wo, fw, W, w0, w1, lw);
                float r[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) r[v] = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    int dd = (k & 4) ? d1 : d0, hh = (k & 2) ? h1 : h0, ww = (k & 1) ? w1 : w0;
                    float wt = ((k & 4) ? ld : 1.f - ld) * ((k & 2) ? lh : 1.f - lh) * ((k & 1) ? lw : 1.f - lw);
                    if (wt != 0.f) {
                        float t[VEC];
                        Vec<T, VEC>::load(xn + (((int64_t)dd * H + hh) * W + ww) * x_ld + cv * VEC, t);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) r[v] = fmaf(wt, t[v], r[v]);
                    }
                }
                Vec<T, VEC>::store(yn + (size_t)s * y_ld + cv * VEC, r);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float q = round_as<T>(r[v]);
                    acc[v][0] += q; acc[v][1] += q * q;
                }
            }
        }
        if (sums) block_channel_reduce<VEC, 2>(acc, sums + (size_t)n * C * 2, cv, active);
    }
}

// backward = transpose of the same sparse weights, written as a gather per low-res voxel (no atomics).
__device__ __forceinline__ float lerp_weight_to(int o, int f, int n, int i) {
    int i0, i1;
    float lam;
    lerp_src(o, f, n, i0, i1, lam);
    float w = 0.f;
    if (i0 == i) w += 1.f - lam;
    if (i1 == i && f != 1) w += lam;
    return w;
}

template <typename T, int VEC>
__global__ void upsample_bwd_kernel(const T* __restrict__ dy, int64_t dy_ld, T* __restrict__ dx, int64_t dx_ld, int D,
                                    int H, int W, int C, int fd, int fh, int fw, int64_t total) {
    const unsigned cvec = C / VEC;
    const int Do = D * fd, Ho = H * fh, Wo = W * fw;
    const int64_t Si = (int64_t)D * H * W, So = (int64_t)Do * Ho * Wo;
    const int64_t n = blockIdx.y;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {
        int cv = (int)(i % cvec);
        const unsigned s = i / cvec;
        int64_t vox = n * Si + s;
        int w = (int)(s % (unsigned)W), h = (int)((s / (unsigned)W) % (unsigned)H), d = (int)(s / (unsigned)(W * H));
        // candidate outputs o with src(o) in [i-1, i+1]
        int dlo = fd == 1 ? d : max(0, fd * d - fd), dhi = fd == 1 ? d : min(Do - 1, fd * d + 2 * fd - 1);
        int hlo = fh == 1 ? h : max(0, fh * h - fh), hhi = fh == 1 ? h : min(Ho - 1, fh * h + 2 * fh - 1);
        int wlo = fw == 1 ? w : max(0, fw * w - fw), whi = fw == 1 ? w : min(Wo - 1, fw * w + 2 * fw - 1);
        float r[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = 0.f;
        for (int od = dlo; od <= dhi; ++od) {
            float wd = lerp_weight_to(od, fd, D, d);
            if (wd == 0.f) continue;
            for (int oh = hlo; oh <= hhi; ++oh) {
                float wh = lerp_weight_to(oh, fh, H, h);
                if (wh == 0.f) continue;
                for (int ow = wlo; ow <= whi; ++ow) {
                    float ww = lerp_weight_to(ow, fw, W, w);
                    if (ww == 0.f) continue;
                    float t[VEC];
                    Vec<T, VEC>::load(dy + (n * So + ((int64_t)od * Ho + oh) * Wo + ow) * dy_ld + cv * VEC, t);
                    float wt = wd * wh * ww;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) r[v] = fmaf(wt, t[v], r[v]);
                }
            }
        }
        Vec<T, VEC>::store(dx + vox * dx_ld + cv * VEC, r);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// im2col of the network's first conv (Cin <= 4): out[vox][tap*Cin + ci] = x_hat[vox + tap][ci] (0 in the padding and in
// the channels >= taps*Cin), bf16, Kp channels per voxel.  Turns the K = 27*Cin conv -- too thin for an implicit GEMM
// -- into a 1x1x1 conv with K = Kp that the tcgen05 kernels take (forward AND weight gradient).  HBM-bound: reads x
// once (neighbours hit L1/L2), writes Kp bf16 per voxel.
// CIN/KD/KH/KW > 0: compile-time filter shape (tap decode becomes constant division); 0: runtime values.
template <typename T, int CIN, int KD, int KH, int KW>
__global__ void __launch_bounds__(256)
im2col_taps_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ in_ss, __nv_bfloat16* __restrict__ out,
                   int D, int H, int W, int Cin_, int kd_, int kh_, int kw_, int Kp, int64_t total) {
    const int Cin = CIN > 0 ? CIN : Cin_, kd = KD > 0 ? KD : kd_, kh = KH > 0 ? KH : kh_, kw = KW > 0 ? KW : kw_;
    const int pd = kd / 2, ph = kh / 2, pw = kw / 2;
    const int64_t S = (int64_t)D * H * W;
    const int groups = Kp / 8;
    const int kmax = kd * kh * kw * Cin;
    const int64_t n = blockIdx.y;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {
        const int g = (int)(i % (unsigned)groups);
        const unsigned s = i / (unsigned)groups;
        const int64_t vox = n * S + s;
        const int w = (int)(s % (unsigned)W), h = (int)((s / (unsigned)W) % (unsigned)H), d = (int)(s / (unsigned)(W * H));
        const T* xn = x + n * S * x_ld;
        float sc = 1.f, sh = 0.f;
        if (CIN == 1 && in_ss) { sc = in_ss[n * 2]; sh = in_ss[n * 2 + 1]; }
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = g * 8 + e;
            float r = 0.f;
            if (k < kmax) {
                const int ci = k % Cin, tp = k / Cin;
                const int a = tp / (kh * kw), b = (tp / kw) % kh, c = tp % kw;
                const int gd = d + a - pd, gh = h + b - ph, gw = w + c - pw;
                if (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
                    r = to_f<T>(xn[(((int64_t)gd * H + gh) * W + gw) * x_ld + ci]);
                    if (CIN == 1) {
                        r = fmaf(r, sc, sh);
                    } else if (in_ss) {
                        const float* p = in_ss + ((size_t)n * Cin + ci) * 2;
                        r = fmaf(r, p[0], p[1]);
                    }
                }
            }
            v[e] = r;
        }
        Vec<__nv_bfloat16, 8>::store(out + vox * Kp + g * 8, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Trilinear x2 / x1 per axis (the only factors the U-Net uses), specialised.
// Forward, block variant (the one launched for factors (2|1, 2, 2) with 4-channel vectors): one thread per LOW-resolution voxel
// and 4 channels produces its FD x 2 x 2 outputs, separably -- with clamped neighbour indices the weights are uniform, so per
// depth slice the 3 x 3 neighbourhood goes w-pass (3 rows -> 3 x 2) and h-pass (-> 2 x 2), and the d-pass combines three such
// slices: 27 loads and ~19 packed lerps (fma.rn.f32x2) per 8 outputs instead of 48 loads and 64 scalar FMAs per output vector.
// The kernel this replaces was issue-bound (ncu: 0.7 IPC per scheduler at 23 % of the HBM rate).
template <int VEC> struct FV { float2 v[VEC / 2]; };
template <int VEC>
__device__ __forceinline__ FV<VEC> lerp14(const FV<VEC>& far_, const FV<VEC>& near75) {      // 0.25 * far + near75 (near75 = 0.75 * near)
    const float2 q = make_float2(0.25f, 0.25f);
    FV<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) r.v[i] = __ffma2_rn(q, far_.v[i], near75.v[i]);
    return r;
}
template <int VEC>
__device__ __forceinline__ FV<VEC> scale75(const FV<VEC>& x) {
    const float2 q = make_float2(0.75f, 0.75f);
    FV<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) r.v[i] = __fmul2_rn(q, x.v[i]);
    return r;
}
template <typename T, int VEC>
__device__ __forceinline__ FV<VEC> load_fv(const T* p) {
    float v[VEC];
    Vec<T, VEC>::load(p, v);
    FV<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) r.v[i] = make_float2(v[2 * i], v[2 * i + 1]);
    return r;
}

template <typename T, int VEC, int FD>
__global__ void __launch_bounds__(256, 2)
upsample2_fwd_block_kernel(const T* __restrict__ x, int64_t x_ld, T* __restrict__ y, int64_t y_ld, int D, int H, int W, int C,
                           float* __restrict__ sums, unsigned total) {
    __shared__ float red[256 * 2];
    const int cvec = C / VEC;
    const int n = blockIdx.y;
    const int Ho = H * 2, Wo = W * 2;
    // block-uniform 64-bit bases + 32-bit BYTE offsets per access (the launch checks that a sample stays below 4 GiB): the
    // address arithmetic of 27 loads and 8 stores would otherwise cost more instructions than the interpolation itself
    const char* xb = reinterpret_cast<const char*>(x + (size_t)n * D * H * W * x_ld);
    char* yb = reinterpret_cast<char*>(y + (size_t)n * (D * FD) * Ho * Wo * y_ld);
    float2 acc_s[VEC / 2], acc_q[VEC / 2];
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) acc_s[i] = acc_q[i] = make_float2(0.f, 0.f);
    using F4 = FV<VEC>;
    const unsigned xl = (unsigned)x_ld * sizeof(T), yl = (unsigned)y_ld * sizeof(T);      // voxel pitches in bytes
    const unsigned xrow = (unsigned)W * xl, xslice = (unsigned)H * xrow, yrow = (unsigned)Wo * yl, yslice = (unsigned)Ho * yrow;
    // blockDim.x * gridDim.x is a multiple of cvec, so a thread keeps its channel vector over the grid-stride loop
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        unsigned t = idx;
        const unsigned cv = t % (unsigned)cvec; t /= (unsigned)cvec;
        const int w = (int)(t % (unsigned)W); t /= (unsigned)W;
        const int h = (int)(t % (unsigned)H); t /= (unsigned)H;
        const int d = (int)t;
        const unsigned cvo = cv * (unsigned)(VEC * sizeof(T));
        const unsigned col[3] = {(unsigned)max(w - 1, 0) * xl + cvo, (unsigned)w * xl + cvo, (unsigned)min(w + 1, W - 1) * xl + cvo};
        const unsigned hro[3] = {(unsigned)max(h - 1, 0) * xrow, (unsigned)h * xrow, (unsigned)min(h + 1, H - 1) * xrow};
        // one depth slice -> its 2 x 2 (oh, ow) outputs
        auto slice = [&](int dd, F4 (&P)[2][2]) {
            F4 R[3][2];
            const unsigned so = (unsigned)dd * xslice;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const unsigned ro = so + hro[b];
                const F4 m = load_fv<T, VEC>(reinterpret_cast<const T*>(xb + (ro + col[0]))), c = load_fv<T, VEC>(reinterpret_cast<const T*>(xb + (ro + col[1]))),
                         pl = load_fv<T, VEC>(reinterpret_cast<const T*>(xb + (ro + col[2])));
                const F4 c75 = scale75(c);
                R[b][0] = lerp14(m, c75);
                R[b][1] = lerp14(pl, c75);
            }
#pragma unroll
            for (int ow = 0; ow < 2; ++ow) {
                const F4 c75 = scale75(R[1][ow]);
                P[0][ow] = lerp14(R[0][ow], c75);
                P[1][ow] = lerp14(R[2][ow], c75);
            }
        };
        const unsigned o00 = (unsigned)(2 * h) * yrow + (unsigned)(2 * w) * yl + cvo;
        auto emit = [&](int od, const F4 (&O)[2][2]) {
            const unsigned ob = (unsigned)od * yslice + o00;
#pragma unroll
            for (int oh = 0; oh < 2; ++oh)
#pragma unroll
                for (int ow = 0; ow < 2; ++ow) {
                    float v[VEC];
#pragma unroll
                    for (int i = 0; i < VEC / 2; ++i) { v[2 * i] = O[oh][ow].v[i].x; v[2 * i + 1] = O[oh][ow].v[i].y; }
                    Vec<T, VEC>::store(reinterpret_cast<T*>(yb + (ob + (oh ? yrow : 0u) + (ow ? yl : 0u))), v);
#pragma unroll
                    for (int i = 0; i < VEC / 2; ++i) {
                        const float2 q = make_float2(round_as<T>(v[2 * i]), round_as<T>(v[2 * i + 1]));
                        acc_s[i] = __fadd2_rn(acc_s[i], q);
                        acc_q[i] = __ffma2_rn(q, q, acc_q[i]);
                    }
                }
        };
        if (FD == 1) {
            F4 P[2][2];
            slice(d, P);
            emit(d, P);
        } else {
            F4 P0[2][2], P1[2][2], O[2][2];
            slice(max(d - 1, 0), P0);
            slice(d, P1);
#pragma unroll
            for (int i = 0; i < 4; ++i) { P1[i >> 1][i & 1] = scale75(P1[i >> 1][i & 1]); O[i >> 1][i & 1] = lerp14(P0[i >> 1][i & 1], P1[i >> 1][i & 1]); }
            emit(2 * d, O);
            slice(min(d + 1, D - 1), P0);
#pragma unroll
            for (int i = 0; i < 4; ++i) O[i >> 1][i & 1] = lerp14(P0[i >> 1][i & 1], P1[i >> 1][i & 1]);
            emit(2 * d + 1, O);
        }
    }
    if (sums) {
        const int cvt = threadIdx.x % cvec;          // blockDim.x % cvec == 0
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            red[threadIdx.x * 2] = (v & 1) ? acc_s[v / 2].y : acc_s[v / 2].x;
            red[threadIdx.x * 2 + 1] = (v & 1) ? acc_q[v / 2].y : acc_q[v / 2].x;
            __syncthreads();
            if ((int)threadIdx.x < cvec) {
                float s1 = 0.f, s2 = 0.f;
                for (int j = threadIdx.x; j < 256; j += cvec) { s1 += red[j * 2]; s2 += red[j * 2 + 1]; }
                atomicAdd(sums + ((size_t)n * C + cvt * VEC + v) * 2, s1);
                atomicAdd(sums + ((size_t)n * C + cvt * VEC + v) * 2 + 1, s2);
            }
            __syncthreads();
        }
    }
}

// Backward, block variant (launched for factors (2|1, 2, 2)): one thread per ND x 2 x 2 block of LOW-resolution outputs (ND = 2 for
// a depth factor of 2, else 1) and channel vector.  With clamped indices the adjoint has uniform weights as well --
//   dx[i] = .25 g[2i-1] + .75 g[2i] + .75 g[2i+1] + .25 g[2i+2]     (g[-1] := g[0], g[2n] := g[2n-1]: the folded edge weights)
// -- so the 6 x 6 x 6 high-resolution neighbourhood of the block is reduced separably in registers (w, then h, then d): 27 loads
// per output instead of 64, ~50 packed multiply-adds instead of 64 x VEC scalar ones.  The fused norm backward of the consuming
// block (coef != null) is applied in the same pass: by linearity U^T (c0 g + c1 U z + c2) = c0 U^T g + c1 (U^T U) z + c2 U^T 1,
// where U^T U is the 3-tap stencil (.375, 1.25, .375) per x2 axis on the clamped low-resolution tensor z and U^T 1 = 2 per x2
// axis; the high-resolution `up` tensor is never read and the intermediate U^T g is never rounded or written.
template <int VEC>
__device__ __forceinline__ FV<VEC> fv_zero() {
    FV<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) r.v[i] = make_float2(0.f, 0.f);
    return r;
}
template <int VEC>
__device__ __forceinline__ FV<VEC> fv_add(const FV<VEC>& a, const FV<VEC>& b) {
    FV<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) r.v[i] = __fadd2_rn(a.v[i], b.v[i]);
    return r;
}
template <int VEC>
__device__ __forceinline__ void fv_axpy(FV<VEC>& acc, float w, const FV<VEC>& x) {          // acc += w * x
    const float2 q = make_float2(w, w);
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) acc.v[i] = __ffma2_rn(q, x.v[i], acc.v[i]);
}
template <int VEC>
__device__ __forceinline__ FV<VEC> fv_mix(float wa, const FV<VEC>& a, float wb, const FV<VEC>& b) {   // wa * a + wb * b
    const float2 qa = make_float2(wa, wa), qb = make_float2(wb, wb);
    FV<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) r.v[i] = __ffma2_rn(qa, a.v[i], __fmul2_rn(qb, b.v[i]));
    return r;
}

template <typename T, int VEC, int FD>
__global__ void __launch_bounds__(128)
upsample2_bwd_block_kernel(const T* __restrict__ dy, int64_t dy_ld, const T* __restrict__ zlow, int64_t zlow_ld, const float* __restrict__ coef,
                           int64_t coef_nstride, T* __restrict__ dx, int64_t dx_ld, int D, int H, int W, int C, unsigned total) {
    using F = FV<VEC>;
    constexpr int ND = FD == 2 ? 2 : 1;              // low-resolution output slices per thread
    constexpr int NS = FD == 2 ? 6 : 1;              // high-resolution slices feeding them
    const int cvec = C / VEC;
    const int n = blockIdx.y;
    const int Do = D * FD, Ho = H * 2, Wo = W * 2;
    const int Dp = (D + ND - 1) / ND, Hp = (H + 1) / 2, Wp = (W + 1) / 2;
    const char* gb = reinterpret_cast<const char*>(dy + (size_t)n * Do * Ho * Wo * dy_ld);
    const char* zb = reinterpret_cast<const char*>(zlow + (coef ? (size_t)n * D * H * W * zlow_ld : 0));
    char* xb = reinterpret_cast<char*>(dx + (size_t)n * D * H * W * dx_ld);
    const unsigned gl = (unsigned)dy_ld * sizeof(T), grow = (unsigned)Wo * gl, gslice = (unsigned)Ho * grow;
    const unsigned zl = (unsigned)zlow_ld * sizeof(T), zrow = (unsigned)W * zl, zslice = (unsigned)H * zrow;
    const unsigned xl = (unsigned)dx_ld * sizeof(T), xrow = (unsigned)W * xl, xslice = (unsigned)H * xrow;
    (void)Dp;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        unsigned t = idx;
        const unsigned cv = t % (unsigned)cvec; t /= (unsigned)cvec;
        const int w0 = 2 * (int)(t % (unsigned)Wp); t /= (unsigned)Wp;
        const int h0 = 2 * (int)(t % (unsigned)Hp); t /= (unsigned)Hp;
        const int d0 = ND * (int)t;
        const unsigned cvo = cv * (unsigned)(VEC * sizeof(T));
        unsigned gcol[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) gcol[k] = (unsigned)min(max(2 * w0 - 1 + k, 0), Wo - 1) * gl + cvo;
        F O[ND][2][2];
#pragma unroll
        for (int i = 0; i < ND * 4; ++i) O[i >> 2][(i >> 1) & 1][i & 1] = fv_zero<VEC>();
        // ---- U^T g: high-resolution slices 2 d0 - 1 ... 2 d0 + 4 (depth factor 2) or the slice d0 (depth factor 1) ----
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int od = FD == 2 ? min(max(2 * d0 - 1 + s, 0), Do - 1) : d0;
            const unsigned so = (unsigned)od * gslice;
            F S[2][2];
            S[0][0] = S[0][1] = S[1][0] = S[1][1] = fv_zero<VEC>();
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const unsigned ro = so + (unsigned)min(max(2 * h0 - 1 + r, 0), Ho - 1) * grow;
                F g[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) g[k] = load_fv<T, VEC>(reinterpret_cast<const T*>(gb + (ro + gcol[k])));
                const F a0 = fv_mix<VEC>(0.25f, fv_add<VEC>(g[0], g[3]), 0.75f, fv_add<VEC>(g[1], g[2]));
                const F a1 = fv_mix<VEC>(0.25f, fv_add<VEC>(g[2], g[5]), 0.75f, fv_add<VEC>(g[3], g[4]));
                constexpr float wt[6] = {0.25f, 0.75f, 0.75f, 0.25f, 0.f, 0.f};
                if (r < 4) { fv_axpy<VEC>(S[0][0], wt[r], a0); fv_axpy<VEC>(S[0][1], wt[r], a1); }
                if (r >= 2) { fv_axpy<VEC>(S[1][0], wt[r - 2], a0); fv_axpy<VEC>(S[1][1], wt[r - 2], a1); }
            }
            constexpr float wt[6] = {0.25f, 0.75f, 0.75f, 0.25f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (FD == 1) {
                    O[0][i >> 1][i & 1] = S[i >> 1][i & 1];
                } else {
                    if (s < 4) fv_axpy<VEC>(O[0][i >> 1][i & 1], wt[s], S[i >> 1][i & 1]);
                    if (s >= 2) fv_axpy<VEC>(O[ND - 1][i >> 1][i & 1], wt[s < 2 ? 0 : s - 2], S[i >> 1][i & 1]);
                }
            }
        }
        if (coef) {
            // ---- c0 * (U^T g) + c1 * (U^T U) z + c2 * (U^T 1) ----
            float c0[VEC], c1[VEC], c2[VEC];
            const float* cf = coef + (size_t)n * coef_nstride + (size_t)cv * VEC * 3;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { c0[v] = cf[3 * v]; c1[v] = cf[3 * v + 1]; c2[v] = cf[3 * v + 2] * (FD == 2 ? 8.f : 4.f); }
#pragma unroll
            for (int i = 0; i < ND * 4; ++i) {
                F& o = O[i >> 2][(i >> 1) & 1][i & 1];
#pragma unroll
                for (int q = 0; q < VEC / 2; ++q)
                    o.v[q] = __ffma2_rn(make_float2(c0[2 * q], c0[2 * q + 1]), o.v[q], make_float2(c2[2 * q], c2[2 * q + 1]));
            }
            unsigned zcol[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) zcol[k] = (unsigned)min(max(w0 - 1 + k, 0), W - 1) * zl + cvo;
            constexpr int NZ = FD == 2 ? 4 : 1;
            constexpr float st[4] = {0.375f, 1.25f, 0.375f, 0.f};
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                const int zd = FD == 2 ? min(max(d0 - 1 + a, 0), D - 1) : d0;
                const unsigned so = (unsigned)zd * zslice;
                F S[2][2];
                S[0][0] = S[0][1] = S[1][0] = S[1][1] = fv_zero<VEC>();
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const unsigned ro = so + (unsigned)min(max(h0 - 1 + r, 0), H - 1) * zrow;
                    F z[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) z[k] = load_fv<T, VEC>(reinterpret_cast<const T*>(zb + (ro + zcol[k])));
                    F a0 = fv_mix<VEC>(0.375f, fv_add<VEC>(z[0], z[2]), 1.25f, z[1]);
                    F a1 = fv_mix<VEC>(0.375f, fv_add<VEC>(z[1], z[3]), 1.25f, z[2]);
                    if (r < 3) { fv_axpy<VEC>(S[0][0], st[r], a0); fv_axpy<VEC>(S[0][1], st[r], a1); }
                    if (r >= 1) { fv_axpy<VEC>(S[1][0], st[r - 1], a0); fv_axpy<VEC>(S[1][1], st[r - 1], a1); }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    F sc = S[i >> 1][i & 1];             // c1 * S, then the depth taps
#pragma unroll
                    for (int q = 0; q < VEC / 2; ++q) sc.v[q] = __fmul2_rn(make_float2(c1[2 * q], c1[2 * q + 1]), sc.v[q]);
                    if (FD == 1) {
                        fv_axpy<VEC>(O[0][i >> 1][i & 1], 1.f, sc);
                    } else {
                        if (a < 3) fv_axpy<VEC>(O[0][i >> 1][i & 1], st[a], sc);
                        if (a >= 1) fv_axpy<VEC>(O[ND - 1][i >> 1][i & 1], st[a < 1 ? 0 : a - 1], sc);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < ND * 4; ++i) {
            const int dd = d0 + (i >> 2), hh = h0 + ((i >> 1) & 1), ww = w0 + (i & 1);
            if (dd < D && hh < H && ww < W) {
                const F& o = O[i >> 2][(i >> 1) & 1][i & 1];
                float v[VEC];
#pragma unroll
                for (int q = 0; q < VEC / 2; ++q) { v[2 * q] = o.v[q].x; v[2 * q + 1] = o.v[q].y; }
                Vec<T, VEC>::store(reinterpret_cast<T*>(xb + ((unsigned)dd * xslice + (unsigned)hh * xrow + (unsigned)ww * xl + cvo)), v);
            }
        }
    }
}

static inline int flat_grid(int64_t total, int threads, int N = 1) {
    int64_t b = (total + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 16 / (N > 0 ? N : 1) + 1;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <typename T>
static bool can_vec(int C, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
    const int V = FullVec<T>::value;
    if (C % V) return false;
    for (int64_t ld : lds)
        if (ld % V) return false;
    for (const void* p : ptrs)
        if (p && !aligned16(p)) return false;
    return true;
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_ncdhw_to_ndhwc(const float* x, void* y, int y_dtype, int64_t y_ld, int N, int C, int64_t S, void* stream) {
    B2_CHECK_ARG(x && y && N > 0 && C > 0 && S > 0 && y_ld >= C, "ncdhw_to_ndhwc: bad arguments");
    int64_t total = (int64_t)N * C * S;
    B2_DISPATCH_DTYPE(y_dtype, T, {
        ncdhw_to_ndhwc_kernel<T><<<flat_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (T*)y, y_ld, C, S, total);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_channel_sums(const void* x, int64_t x_ld, int dtype, int N, int64_t S, int C, float* sums, void* stream) {
    B2_CHECK_ARG(x && sums && N > 0 && C > 0 && S > 0 && x_ld >= C, "channel_sums: bad arguments");
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        if (can_vec<T>(C, {x_ld}, {x})) {
            Launch2D l = make_launch(C / V, S, N);
            channel_sums_kernel<T, V, false><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, nullptr, 0, S, C, sums);
        } else {
            Launch2D l = make_launch(C, S, N);
            channel_sums_kernel<T, 1, false><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, nullptr, 0, S, C, sums);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_channel_dot_sums(const void* g, int64_t g_ld, const void* x, int64_t x_ld, int dtype, int N, int64_t S, int C,
                            float* sums, void* stream) {
    B2_CHECK_ARG(g && x && sums && N > 0 && C > 0 && S > 0 && g_ld >= C && x_ld >= C, "channel_dot_sums: bad arguments");
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        if (can_vec<T>(C, {g_ld, x_ld}, {g, x})) {
            Launch2D l = make_launch(C / V, S, N);
            channel_sums_kernel<T, V, true><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)g, g_ld, (const T*)x, x_ld, S, C, sums);
        } else {
            Launch2D l = make_launch(C, S, N);
            channel_sums_kernel<T, 1, true><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)g, g_ld, (const T*)x, x_ld, S, C, sums);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_norm_finalize(const float* sums, int N, int C, int64_t S, int groups, const float* gamma, const float* beta,
                         float eps, float* scale_shift, float* mean_rstd, void* stream) {
    B2_CHECK_ARG(sums && scale_shift && mean_rstd && N > 0 && C > 0 && S > 0, "norm_finalize: bad arguments");
    B2_CHECK_ARG(groups > 0 && groups <= C && C % groups == 0, "norm_finalize: groups %d does not divide C %d", groups, C);
    int total = N * groups;
    norm_finalize_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, N, C, (float)S, groups, gamma, beta, eps,
                                                                                  scale_shift, mean_rstd);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_norm_bwd_finalize(const float* dsums, const float* mean_rstd, const float* gamma, int N, int C, int64_t S,
                             int groups, float* coef, float* dgamma, float* dbeta, void* stream) {
    B2_CHECK_ARG(dsums && mean_rstd && coef && N > 0 && C > 0 && S > 0, "norm_bwd_finalize: bad arguments");
    B2_CHECK_ARG(groups > 0 && groups <= C && C % groups == 0, "norm_bwd_finalize: groups %d does not divide C %d", groups, C);
    int total = N * groups;
    norm_bwd_finalize_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dsums, mean_rstd, gamma, N, C, (float)S,
                                                                                      groups, coef, dgamma, dbeta);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_affine_apply(const void* x, int64_t x_ld, const float* scale_shift, void* y, int64_t y_ld, int dtype, int N,
                        int64_t S, int C, void* stream) {
    B2_CHECK_ARG(x && y && scale_shift && N > 0 && C > 0 && S > 0 && x_ld >= C && y_ld >= C, "affine_apply: bad arguments");
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        if (can_vec<T>(C, {x_ld, y_ld}, {x, y})) {
            int64_t total = S * (C / V);
            affine_apply_kernel<T, V><<<dim3(flat_grid(total, 256, N), N), 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, scale_shift, (T*)y, y_ld, S, C, total);
        } else {
            int64_t total = S * C;
            affine_apply_kernel<T, 1><<<dim3(flat_grid(total, 256, N), N), 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, scale_shift, (T*)y, y_ld, S, C, total);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_norm_bwd_apply(const void* g, int64_t g_ld, const void* x, int64_t x_ld, const float* coef, const void* add,
                          int64_t add_ld, void* out, int64_t out_ld, int dtype, int N, int64_t S, int C, int relu_mask,
                          float* absmax, void* stream) {
    B2_CHECK_ARG(g && out && N > 0 && C > 0 && S > 0 && g_ld >= C && out_ld >= C, "norm_bwd_apply: bad arguments");
    B2_CHECK_ARG(S * C < (1LL << 31) && N <= 65535, "norm_bwd_apply: sample too large for 32-bit indexing");
    B2_CHECK_ARG(S * C < (1LL << 31) && N <= 65535, "norm_bwd_apply: sample too large for 32-bit indexing");
    B2_CHECK_ARG(x || (!coef && !relu_mask), "norm_bwd_apply: x is required with coef or relu_mask");
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        if (can_vec<T>(C, {g_ld, x ? x_ld : (int64_t)V, add ? add_ld : (int64_t)V, out_ld}, {g, x, add, out})) {
            int64_t total = S * (C / V);
            norm_bwd_apply_kernel<T, V><<<dim3(flat_grid(total, 256, N), N), 256, 0, (cudaStream_t)stream>>>(
                (const T*)g, g_ld, (const T*)x, x_ld, coef, (const T*)add, add_ld, (T*)out, out_ld, S, C, relu_mask, total, absmax);
        } else {
            int64_t total = S * C;
            norm_bwd_apply_kernel<T, 1><<<dim3(flat_grid(total, 256, N), N), 256, 0, (cudaStream_t)stream>>>(
                (const T*)g, g_ld, (const T*)x, x_ld, coef, (const T*)add, add_ld, (T*)out, out_ld, S, C, relu_mask, total, absmax);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_split_bf16(const float* x, int64_t x_ld, const float* in_scale_shift, void* hi, void* lo, int N, int64_t S, int C,
                      void* stream) {
    B2_CHECK_ARG(x && hi && lo && N > 0 && S > 0 && C > 0 && N <= 65535, "split_bf16: bad arguments");
    B2_CHECK_ARG(C % 4 == 0 && x_ld % 4 == 0 && aligned16(x) && aligned16(hi) && aligned16(lo), "split_bf16: needs C % 4 == 0 and 16-byte aligned tensors");
    split_bf16_kernel<<<dim3(flat_grid(S * (C / 4), 256, N), N), 256, 0, (cudaStream_t)stream>>>(x, x_ld, in_scale_shift, (__nv_bfloat16*)hi,
                                                                                                (__nv_bfloat16*)lo, S, C);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_absmax_f32(const float* x, int64_t x_ld, int64_t rows, int C, float* absmax, float* colsum, void* stream) {
    B2_CHECK_ARG(x && absmax && rows > 0 && C > 0, "absmax_f32: bad arguments");
    B2_CHECK_ARG(C % 4 == 0 && x_ld % 4 == 0 && aligned16(x) && C <= 4096, "absmax_f32: needs C % 4 == 0, C <= 4096 and a 16-byte aligned tensor");
    const int cvec = C / 4;
    const int bx = cvec < 256 ? cvec : 256;                   // a thread owns up to 4 channel vectors: cvec <= 1024
    const int by = 256 / bx;
    int64_t blocks = (rows + by - 1) / by;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    absmax_f32_kernel<<<(unsigned)blocks, dim3(bx, by), colsum ? C * sizeof(float) : 0, (cudaStream_t)stream>>>(
        x, x_ld, rows, C, reinterpret_cast<unsigned*>(absmax), colsum);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_cvt_f16(const float* x, int64_t x_ld, const float* in_scale_shift, const float* absmax, void* out, float* colsum, int N,
                   int64_t S, int C, void* stream) {
    B2_CHECK_ARG(x && out && N > 0 && S > 0 && C > 0 && N <= 65535, "cvt_f16: bad arguments");
    B2_CHECK_ARG(C % 8 == 0 && x_ld % 4 == 0 && aligned16(x) && aligned16(out), "cvt_f16: needs C % 8 == 0 and 16-byte aligned tensors");
    B2_CHECK_ARG(!colsum || C <= 8192, "cvt_f16: colsum needs C <= 8192");
    const int cvec = C / 8;
    int bx = 1;
    while (bx < cvec && bx < 256) bx <<= 1;           // power of two: 256 / bx voxels per block pass
    const int by = 256 / bx;
    int64_t blocks = (S + by - 1) / by;
    int64_t cap = (int64_t)sm_count() * 16 / N + 1;
    if (colsum && cap > (int64_t)sm_count() * 8 / N + 1) cap = (int64_t)sm_count() * 8 / N + 1;   // every block ends with C atomics
    if (blocks > cap) blocks = cap;
    cvt_f16_kernel<<<dim3((unsigned)blocks, N), dim3(bx, by), colsum ? C * sizeof(float) : 0, (cudaStream_t)stream>>>(
        x, x_ld, in_scale_shift, absmax, (__half*)out, colsum, S, C);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_maxpool3d_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int dtype, int N, int D, int H, int W, int C,
                         int fd, int fh, int fw, float* sums, void* stream) {
    B2_CHECK_ARG(x && y && N > 0 && C > 0 && fd > 0 && fh > 0 && fw > 0, "maxpool3d_fwd: bad arguments");
    // like nn.MaxPool3d, trailing voxels that do not fill a window are ignored (output dims = floor(dims / factor))
    B2_CHECK_ARG(D >= fd && H >= fh && W >= fw, "maxpool3d_fwd: (%d,%d,%d) smaller than the window (%d,%d,%d)", D, H, W, fd, fh, fw);
    int64_t So = (int64_t)(D / fd) * (H / fh) * (W / fw);
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        if (can_vec<T>(C, {x_ld, y_ld}, {x, y})) {
            Launch2D l = make_launch(C / V, So, N);
            if (fd == 2 && fh == 2 && fw == 2)
                maxpool_fwd_kernel<T, V, 1, 2><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, fd, fh, fw, sums);
            else if (fd == 1 && fh == 2 && fw == 2)
                maxpool_fwd_kernel<T, V, 1, 1><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, fd, fh, fw, sums);
            else
                maxpool_fwd_kernel<T, V><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, fd, fh, fw, sums);
        } else {
            Launch2D l = make_launch(C, So, N);
            maxpool_fwd_kernel<T, 1><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, fd, fh, fw, sums);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_maxpool3d_bwd(const void* x, int64_t x_ld, const void* dp, int64_t dp_ld, const void* add, int64_t add_ld,
                         const float* coef, int64_t coef_nstride, void* out, int64_t out_ld, int dtype, int N, int D, int H, int W,
                         int C, int fd, int fh, int fw, int relu_mask, float* absmax, void* stream) {
    B2_CHECK_ARG(x && dp && out && N > 0 && C > 0 && fd > 0 && fh > 0 && fw > 0, "maxpool3d_bwd: bad arguments");
    B2_CHECK_ARG(!coef || add, "maxpool3d_bwd: coef needs the raw gradient in `add`");
    // voxels beyond the last full window are NOT written (the caller initialises them: they receive no pooled gradient)
    B2_CHECK_ARG(D >= fd && H >= fh && W >= fw, "maxpool3d_bwd: dims smaller than the window");
    B2_CHECK_ARG((int64_t)D * H * W * C < (1LL << 31) && N <= 65535, "maxpool3d_bwd: sample too large for 32-bit indexing");
    int64_t So = (int64_t)(D / fd) * (H / fh) * (W / fw);
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        const int64_t si_ = (int64_t)D * H * W;
        const int64_t ldmax = std::max(std::max(x_ld, out_ld), std::max(add ? add_ld : (int64_t)0, dp_ld));
        if (can_vec<T>(C, {x_ld, dp_ld, add ? add_ld : (int64_t)V, out_ld}, {x, dp, add, out}) && fd <= 2 && fh == 2 && fw == 2 &&
            si_ * ldmax * (int64_t)sizeof(T) < (1LL << 32) && D % fd == 0 && H % 2 == 0 && W % 2 == 0) {
            int64_t total = So * (C / V);
            const dim3 grid(flat_grid(total, 256, N), N);
            const size_t sm = coef ? C * 3 * sizeof(float) : 0;
            if (fd == 2)
                maxpool2_bwd_kernel<T, V, 2><<<grid, 256, sm, (cudaStream_t)stream>>>((const T*)x, x_ld, (const T*)dp, dp_ld, (const T*)add, add_ld, coef, coef_nstride,
                                                                                      (T*)out, out_ld, D, H, W, C, relu_mask, (unsigned)total, absmax);
            else
                maxpool2_bwd_kernel<T, V, 1><<<grid, 256, sm, (cudaStream_t)stream>>>((const T*)x, x_ld, (const T*)dp, dp_ld, (const T*)add, add_ld, coef, coef_nstride,
                                                                                      (T*)out, out_ld, D, H, W, C, relu_mask, (unsigned)total, absmax);
        } else if (can_vec<T>(C, {x_ld, dp_ld, add ? add_ld : (int64_t)V, out_ld}, {x, dp, add, out})) {
            int64_t total = So * (C / V);
            maxpool_bwd_kernel<T, V><<<dim3(flat_grid(total, 256, N), N), 256, coef ? C * 3 * sizeof(float) : 0, (cudaStream_t)stream>>>(
                (const T*)x, x_ld, (const T*)dp, dp_ld, (const T*)add, add_ld, coef, coef_nstride, (T*)out, out_ld, D, H, W, C, fd, fh, fw,
                relu_mask, total, absmax);
        } else {
            int64_t total = So * C;
            maxpool_bwd_kernel<T, 1><<<dim3(flat_grid(total, 256, N), N), 256, coef ? C * 3 * sizeof(float) : 0, (cudaStream_t)stream>>>(
                (const T*)x, x_ld, (const T*)dp, dp_ld, (const T*)add, add_ld, coef, coef_nstride, (T*)out, out_ld, D, H, W, C, fd, fh, fw,
                relu_mask, total, absmax);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_upsample_trilinear_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int dtype, int N, int D, int H, int W,
                                  int C, int fd, int fh, int fw, float* sums, void* stream) {
    B2_CHECK_ARG(x && y && N > 0 && C > 0 && fd > 0 && fh > 0 && fw > 0, "upsample_fwd: bad arguments");
    int64_t So = (int64_t)D * fd * H * fh * W * fw;
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        const int cvec_ = C / V;
        const int64_t totv = (int64_t)D * H * W * cvec_, so_bytes = So * y_ld * (int64_t)sizeof(T);
        if (can_vec<T>(C, {x_ld, y_ld}, {x, y}) && cvec_ <= 256 && 256 % cvec_ == 0 && fd <= 2 && fh == 2 && fw == 2 && totv < (1LL << 31) &&
            so_bytes < (1LL << 32) && (int64_t)D * H * W * x_ld * (int64_t)sizeof(T) < (1LL << 32) && N <= 65535) {
            // >= 4 low-resolution voxels (32 outputs) per thread: the per-block statistics reduction and its atomics amortise
            // few, long-lived blocks: every block ends with 2 * C atomics on the same N * C * 2 statistics words
            int64_t blocks = (totv + 256 * 4 - 1) / (256 * 4);
            const int64_t cap = ((int64_t)sm_count() * 2 + N - 1) / N;    // one resident wave (2 blocks per SM)
            if (blocks > cap) blocks = cap;
            while ((blocks * 256) % cvec_) ++blocks;         // a thread keeps its channel vector over the grid-stride loop
            dim3 grid((unsigned)blocks, (unsigned)N, 1);
            if (fd == 2)
                upsample2_fwd_block_kernel<T, V, 2><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, sums, (unsigned)totv);
            else
                upsample2_fwd_block_kernel<T, V, 1><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, sums, (unsigned)totv);
        } else if (can_vec<T>(C, {x_ld, y_ld}, {x, y})) {
            Launch2D l = make_launch(C / V, So, N);
            upsample_fwd_kernel<T, V><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, fd, fh, fw, sums);
        } else {
            Launch2D l = make_launch(C, So, N);
            upsample_fwd_kernel<T, 1><<<l.grid, l.block, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, (T*)y, y_ld, D, H, W, C, fd, fh, fw, sums);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_upsample_trilinear_bwd(const void* dy, int64_t dy_ld, const void* zlow, int64_t zlow_ld, const float* coef,
                                  int64_t coef_nstride, void* dx, int64_t dx_ld, int dtype, int N, int D, int H,
                                  int W, int C, int fd, int fh, int fw, void* stream) {
    B2_CHECK_ARG(dy && dx && N > 0 && C > 0 && fd > 0 && fh > 0 && fw > 0, "upsample_bwd: bad arguments");
    B2_CHECK_ARG(!coef || zlow, "upsample_bwd: coef needs zlow");
    int64_t Si = (int64_t)D * H * W;
    B2_CHECK_ARG(Si * fd * fh * fw * C < (1LL << 31) && N <= 65535, "upsample_bwd: sample too large for 32-bit indexing");
    B2_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = FullVec<T>::value;
        const int cvec_ = C / V;
        if (can_vec<T>(C, {dy_ld, dx_ld, coef ? zlow_ld : (int64_t)V}, {dy, dx, coef ? zlow : nullptr}) && cvec_ <= 256 && 256 % cvec_ == 0 &&
            fd <= 2 && fh == 2 && fw == 2) {
            const int nd = fd == 2 ? 2 : 1;
            const int64_t totb = (int64_t)((D + nd - 1) / nd) * ((H + 1) / 2) * ((W + 1) / 2) * cvec_;
            if (Si * fd * 4 * dy_ld * (int64_t)sizeof(T) < (1LL << 32) && Si * dx_ld * (int64_t)sizeof(T) < (1LL << 32) &&
                (!coef || Si * zlow_ld * (int64_t)sizeof(T) < (1LL << 32))) {
                int64_t blocks = (totb + 127) / 128;
                const int64_t cap = (int64_t)sm_count() * 16;
                if (blocks > cap) blocks = cap;
                dim3 grid((unsigned)blocks, (unsigned)N, 1);
                if (fd == 2)
                    upsample2_bwd_block_kernel<T, V, 2><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)dy, dy_ld, (const T*)zlow, zlow_ld, coef, coef_nstride,
                                                                                                  (T*)dx, dx_ld, D, H, W, C, (unsigned)totb);
                else
                    upsample2_bwd_block_kernel<T, V, 1><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)dy, dy_ld, (const T*)zlow, zlow_ld, coef, coef_nstride,
                                                                                                  (T*)dx, dx_ld, D, H, W, C, (unsigned)totb);
                B2_LAUNCH_CHECK();
                return 0;
            }
            if (coef) {
                set_error("upsample_bwd: the fused norm backward needs tensors below 4 GiB per sample (32-bit byte offsets)");
                return 2;
            }
            // (larger samples: the generic kernel below)
            int64_t total = Si * (C / V);
            upsample_bwd_kernel<T, V><<<dim3(flat_grid(total, 256, N), N), 256, 0, (cudaStream_t)stream>>>((const T*)dy, dy_ld, (T*)dx, dx_ld, D, H, W, C, fd, fh, fw, total);
        } else if (coef) {
            set_error("upsample_bwd: the fused norm backward needs 16-byte aligned channel vectors (a power-of-two count of them) and factors (1|2, 2, 2)");
            return 2;
        } else if (can_vec<T>(C, {dy_ld, dx_ld}, {dy, dx})) {
            int64_t total = Si * (C / V);
            upsample_bwd_kernel<T, V><<<dim3(flat_grid(total, 256, N), N), 256, 0, (cudaStream_t)stream>>>((const T*)dy, dy_ld, (T*)dx, dx_ld, D, H, W, C, fd, fh, fw, total);
        } else {
            int64_t total = Si * C;
            upsample_bwd_kernel<T, 1><<<dim3(flat_grid(total, 256, N), N), 256, 0, (cudaStream_t)stream>>>((const T*)dy, dy_ld, (T*)dx, dx_ld, D, H, W, C, fd, fh, fw, total);
        }
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_im2col_taps(const void* x, int64_t x_ld, const float* in_scale_shift, int dtype, void* out, int N, int D, int H, int W,
                       int Cin, int kd, int kh, int kw, int Kp, void* stream) {
    B2_CHECK_ARG(x && out && N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0, "im2col_taps: bad arguments");
    B2_CHECK_ARG(Kp % 8 == 0 && Kp >= kd * kh * kw * Cin, "im2col_taps: Kp must be a multiple of 8 and >= taps*Cin");
    B2_CHECK_ARG(aligned16(out), "im2col_taps: output must be 16-byte aligned");
    int64_t total = (int64_t)D * H * W * (Kp / 8);
    B2_CHECK_ARG(total < (1LL << 31), "im2col_taps: sample too large for 32-bit indexing");
    B2_DISPATCH_DTYPE(dtype, T, {
        const dim3 grid(flat_grid(total, 256, N), N);
        cudaStream_t st = (cudaStream_t)stream;
        if (Cin == 1 && kd == 3 && kh == 3 && kw == 3)
            im2col_taps_kernel<T, 1, 3, 3, 3><<<grid, 256, 0, st>>>((const T*)x, x_ld, in_scale_shift, (__nv_bfloat16*)out, D, H, W, Cin, kd, kh, kw, Kp, total);
        else if (Cin == 1 && kd == 1 && kh == 3 && kw == 3)
            im2col_taps_kernel<T, 1, 1, 3, 3><<<grid, 256, 0, st>>>((const T*)x, x_ld, in_scale_shift, (__nv_bfloat16*)out, D, H, W, Cin, kd, kh, kw, Kp, total);
        else
            im2col_taps_kernel<T, 0, 0, 0, 0><<<grid, 256, 0, st>>>((const T*)x, x_ld, in_scale_shift, (__nv_bfloat16*)out, D, H, W, Cin, kd, kh, kw, Kp, total);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
