// Shared helpers for libb200em (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200em.h"

namespace b200em {

// ---- error plumbing -----------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define B2_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            b200em::set_error(__VA_ARGS__);     \
            return 1;                           \
        }                                       \
    } while (0)

#define B2_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            b200em::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return 3;                                                                         \
        }                                                                                     \
    } while (0)

#define B2_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) {                                                             \
            b200em::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return 3;                                                                         \
        }                                                                                     \
        b200em::count_launch();                                                               \
    } while (0)

int sm_count();
// the caller-provided zero-initialised scratch of the current device (b200em_set_workspace), false if none is registered
constexpr int WS_COUNTER_BYTES = 1 << 20;
bool get_workspace(float** acc, int64_t* acc_floats, unsigned** counters, int* ncounters);

// ---- scalar conversion ----------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- 16-byte vectors of activations ----------------------------------------------------------------------------
// VEC elements per thread: 4 (f32) / 8 (bf16) when everything is 16 B aligned, else 1.
template <typename T, int VEC> struct Vec;
template <> struct Vec<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<float, 1> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
};
template <> struct Vec<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
};
template <> struct Vec<__nv_bfloat16, 4> {           // 8-byte accesses: half a 16-byte unit per thread (register-heavy stencils)
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        uint2 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
        h[0] = __floats2bfloat162_rn(v[0], v[1]);
        h[1] = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = t;
    }
};
template <> struct Vec<__nv_bfloat16, 1> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) { v[0] = __bfloat162float(*p); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) { *p = __float2bfloat16_rn(v[0]); }
};

template <typename T> struct FullVec;
template <> struct FullVec<float> { static constexpr int value = 4; };
template <> struct FullVec<__nv_bfloat16> { static constexpr int value = 8; };

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Round a stored value the way the activation dtype would (bf16 statistics are taken on the rounded values).
template <typename T> __device__ __forceinline__ float round_as(float v) { return to_f<T>(from_f<T>(v)); }

// ---- power-of-two operand scaling of the h16 path ---------------------------------------------------------------
// fp32 tensors without a bounded range (gradients, un-normalised activations) are multiplied by 2^k before they are rounded
// to fp16, k chosen from the tensor's max |x| so that the largest element lands in [2^14, 2^15); the consuming kernel undoes
// it (exactly) on the fp32 accumulator.  `absmax` is a DEVICE float written by b200em_absmax_f32 -- no host round trip.
__device__ __forceinline__ int h16_shift(const float* absmax) {
    if (!absmax) return 0;
    const uint32_t b = __float_as_uint(*absmax) & 0x7fffffffu;
    const int e = (int)(b >> 23) - 127;
    if (b == 0 || e == 128) return 0;                  // all zeros, or inf / nan somewhere: leave the values alone
    const int k = 14 - e;
    return k < -100 ? -100 : (k > 100 ? 100 : k);
}
__device__ __forceinline__ float pow2i(int k) { return __uint_as_float((uint32_t)(k + 127) << 23); }

// running max |v| of a thread (as the bit pattern) -> one atomicMax per warp into the device absmax word (nullable)
__device__ __forceinline__ unsigned absmax_acc(unsigned am, float v) { return max(am, __float_as_uint(v) & 0x7fffffffu); }
__device__ __forceinline__ void absmax_flush(unsigned am, float* absmax) {
    if (!absmax) return;
    am = __reduce_max_sync(0xffffffffu, am);
    // thousands of warps target ONE word: look first (a stale read only costs a redundant atomic), so that all but the first few skip it
    if ((threadIdx.x & 31) == 0 && am > *reinterpret_cast<volatile unsigned*>(absmax)) atomicMax(reinterpret_cast<unsigned*>(absmax), am);
}

// dispatch on dtype code
#define B2_DISPATCH_DTYPE(dtype, T, ...)                                   \
    if ((dtype) == B200EM_F32) { using T = float; __VA_ARGS__ }            \
    else if ((dtype) == B200EM_BF16) { using T = __nv_bfloat16; __VA_ARGS__ } \
    else { b200em::set_error("bad dtype code %d", (int)(dtype)); return 1; }

}  // namespace b200em
