// Haloed NDHWC tile -> shared memory in the tcgen05 SWIZZLE_NONE core-matrix layout (shared by the conv kernels).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>
#include <stdint.h>

namespace b200em {

// Loads are issued in batches of LD_BATCH before any is consumed, so each thread keeps LD_BATCH 16-byte requests in
// flight (the tile load is latency-bound otherwise: one dependent global load per iteration).
constexpr int LD_BATCH = 8;

// Operand load for ONE 8-channel group (xn already points at the group's first channel): voxels v = v0, v0+step, ...
// of an (nslices x HP_ x WP_) haloed tile whose first voxel is (d0 - pd, h0 - 1, w0 - 1).  Element (s, hp, wp) goes to
// dst + s*slice_stride_bytes + (hp*WP_ + wp)*16.  Out-of-volume voxels are written as zeros; in-volume ones get
// scale*x + shift when `affine` (the fused InstanceNorm / GroupNorm apply of the preceding norm layer).
template <int HP_, int WP_>
__device__ __forceinline__ void load_halo_tile(const __nv_bfloat16* __restrict__ xn, long long x_ld, const float* sc, const float* sh,
                                               bool affine, uint8_t* dst, int slice_stride_bytes, int v0, int step, int units,
                                               int d0, int h0, int w0, int pd, int D, int H, int W) {
    for (int vb = v0; vb < units; vb += LD_BATCH * step) {
        uint4 val[LD_BATCH];
        int off[LD_BATCH];
        uint32_t inb = 0;
#pragma unroll
        for (int i = 0; i < LD_BATCH; ++i) {
            const int v = vb + i * step;
            val[i] = make_uint4(0, 0, 0, 0);
            off[i] = -1;
            if (v < units) {
                const int wp_ = v % WP_, hp_ = (v / WP_) % HP_, s = v / (WP_ * HP_);
                const int gd = d0 + s - pd, gh = h0 + hp_ - 1, gw = w0 + wp_ - 1;
                off[i] = s * slice_stride_bytes + (hp_ * WP_ + wp_) * 16;
                if (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
                    val[i] = __ldg(reinterpret_cast<const uint4*>(xn + (((size_t)gd * H + gh) * W + gw) * x_ld));
                    inb |= 1u << i;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < LD_BATCH; ++i) {
            if (off[i] >= 0) {
                if (affine && ((inb >> i) & 1)) {
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&val[i]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 f = __bfloat1622float2(h2[e]);
                        f.x = fmaf(f.x, sc[2 * e], sh[2 * e]);
                        f.y = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
                        h2[e] = __floats2bfloat162_rn(f.x, f.y);
                    }
                }
                *reinterpret_cast<uint4*>(dst + off[i]) = val[i];
            }
        }
    }
}

// In-place norm apply on one 16-byte unit: 8 bf16 or 4 fp32 channels, scale * x + shift in fp32.
template <typename T>
__device__ __forceinline__ void affine_unit(uint4& val, const float* sc, const float* sh) {
    if constexpr (sizeof(T) == 4) {
        float* f = reinterpret_cast<float*>(&val);
#pragma unroll
        for (int e = 0; e < 4; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
    } else if constexpr (std::is_same<T, __half>::value) {
        __half2* h2 = reinterpret_cast<__half2*>(&val);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 f = __half22float2(h2[e]);
            f.x = fmaf(f.x, sc[2 * e], sh[2 * e]);
            f.y = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
            h2[e] = __floats2half2_rn(f.x, f.y);
        }
    } else {
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 f = __bfloat1622float2(h2[e]);
            f.x = fmaf(f.x, sc[2 * e], sh[2 * e]);
            f.y = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
            h2[e] = __floats2bfloat162_rn(f.x, f.y);
        }
    }
}

// Asynchronous variant: every 16-byte unit of the thread is issued as one cp.async (LDGSTS, zero-filled when the voxel is
// outside the volume) with no register staging, so ALL of a thread's units are in flight at once and the tile costs one
// memory latency instead of one per register batch; the norm apply then runs as an in-place shared-memory pass over the
// in-volume units this thread copied.  At most 32 units per thread.  T = __nv_bfloat16 (8 channels per unit) or float (4).
template <int HP_, int WP_, typename T = __nv_bfloat16>
__device__ __forceinline__ void load_halo_tile_async(const T* __restrict__ xn, long long x_ld, const float* sc, const float* sh,
                                                     bool affine, uint8_t* dst, int slice_stride_bytes, int v0, int step, int units,
                                                     int d0, int h0, int w0, int pd, int D, int H, int W) {
    uint32_t inb = 0;
    const uint32_t dst32 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    int i = 0;
    for (int v = v0; v < units; v += step, ++i) {
        const int wp_ = v % WP_, hp_ = (v / WP_) % HP_, s = v / (WP_ * HP_);
        const int gd = d0 + s - pd, gh = h0 + hp_ - 1, gw = w0 + wp_ - 1;
        const uint32_t off = (uint32_t)(s * slice_stride_bytes + (hp_ * WP_ + wp_) * 16);
        const bool in = gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W;
        const T* src = in ? xn + (((size_t)gd * H + gh) * W + gw) * x_ld : xn;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst32 + off), "l"(src), "r"(in ? 16 : 0) : "memory");
        inb |= (in ? 1u : 0u) << i;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (affine) {
        i = 0;
        for (int v = v0; v < units; v += step, ++i) {
            if (!((inb >> i) & 1)) continue;
            const int wp_ = v % WP_, hp_ = (v / WP_) % HP_, s = v / (WP_ * HP_);
            uint4* q = reinterpret_cast<uint4*>(dst + s * slice_stride_bytes + (hp_ * WP_ + wp_) * 16);
            uint4 val = *q;
            affine_unit<T>(val, sc, sh);
            *q = val;
        }
    }
}

}  // namespace b200em
