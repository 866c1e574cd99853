// sm_100a building blocks: mbarrier, bulk async copy (TMA engine, UBLKCP), tcgen05 MMA / TMEM, descriptors.
// Inline PTX only -- no CUTLASS.  Everything here is cta_group::1 (one CTA per MMA).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <type_traits>

namespace b200em {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe (mbarrier.test_wait never suspends the thread).  A probe of an ALREADY complete barrier still costs
// ~200 cycles of latency (measured, scripts/ubench/umma_commit.cu), which a single-thread MMA issue loop cannot hide unless
// the probe for the NEXT stage is issued before the current stage's MMAs and its result is consumed after them.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must end as a trapped kernel (launch failure reported to the host), never as a hung GPU.
// The bound is a build-time constant (-DB200EM_WATCHDOG_CYCLES=n; 0 compiles the watchdog out, e.g. for runs under
// compute-sanitizer or a debugger whose slowdown could otherwise trip it).  Default: ~20 s at 2 GHz.
#ifndef B200EM_WATCHDOG_CYCLES
#define B200EM_WATCHDOG_CYCLES 40000000000LL
#endif
__device__ __forceinline__ void watchdog(long long t0) {
#if B200EM_WATCHDOG_CYCLES > 0
    if (clock64() - t0 > (long long)B200EM_WATCHDOG_CYCLES) __trap();
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) watchdog(t0);
}

// ---- plain shared-memory flags / counters for the hand-offs INTO the single-thread MMA issue loop ----------------------
// An mbarrier probe costs the issuing thread ~200 cycles even when the barrier is already complete (measured,
// scripts/ubench/umma_commit.cu) and the MMA thread cannot hide that; a volatile shared-memory load is ~30.  Producers
// that are ordinary threads (operand loaders, epilogue warps) therefore publish with a plain store / atomic add after a
// block-level fence, and only the hardware-signalled hand-offs (tcgen05.commit, cp.async.bulk) stay mbarriers.
// The storage is an 8-byte slot of the barrier array, zero-initialised instead of mbarrier.init.
__device__ __forceinline__ void flag_init(uint64_t* f) { *reinterpret_cast<volatile uint64_t*>(f) = 0ull; }
__device__ __forceinline__ void flag_store(uint64_t* f, uint32_t v) {
    __threadfence_block();
    *reinterpret_cast<volatile uint32_t*>(f) = v;
}
__device__ __forceinline__ void counter_add(uint64_t* f) {
    __threadfence_block();
    atomicAdd(reinterpret_cast<unsigned int*>(f), 1u);
}
__device__ __forceinline__ void flag_wait_eq(uint64_t* f, uint32_t v) {
    const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(f);
    if (*q == v) return;
    const long long t0 = clock64();
    while (*q != v) watchdog(t0);
}
__device__ __forceinline__ void counter_wait_ge(uint64_t* f, uint32_t v) {
    const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(f);
    if (*q >= v) return;
    const long long t0 = clock64();
    while (*q < v) watchdog(t0);
}

// ---- proxies / bulk copy -----------------------------------------------------------------------------------
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// contiguous global -> shared copy by the TMA engine, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------------------------
// whole warp; ncols power of two in [32, 512]; the allocated base address is written to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one thread: arrive on the mbarrier when all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 x bf16 -> fp32, M = 128 rows = TMEM lanes, N columns
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Same, accumulate flag known at compile time (no setp from a register in the single-thread issue loop).
template <bool ACC>
__device__ __forceinline__ void umma_bf16_c(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if (ACC) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.eq.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    }
}

// The same with TF32 operands (kind::tf32): A and B are 32-bit words of which the tensor core reads the upper 19 bits (fp32
// activations and weights are fed as they are), K = 8 per instruction = two 16-byte core-matrix columns of 4 elements.
template <bool ACC>
__device__ __forceinline__ void umma_tf32_c(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if (ACC) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.eq.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    }
}

// Operand-type dispatch of the issue loops: __nv_bfloat16 / __half -> kind::f16 (K = 16; the instruction descriptor names the
// format), float -> kind::tf32 (K = 8).
template <typename TA, bool ACC>
__device__ __forceinline__ void umma_c(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if constexpr (sizeof(TA) == 4) umma_tf32_c<ACC>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc);
    else umma_bf16_c<ACC>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave"), version 1 (Blackwell).
// K-major operand: 8-row x 16-byte core matrices, each 128 contiguous bytes (row r of the core matrix at +16*r);
//   lbo = byte distance between the two 16-byte K-chunks of one K=16 step, sbo = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor, kind::f16: bf16 A and B (both K-major), fp32 accumulator, M x N.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with IEEE fp16 A and B (format code 0): the "h16" path of fp32 activations -- fp16 carries the same 11-bit
// significand as TF32 (the precision torch / cuDNN use for fp32 convolutions), at the bf16 MMA rate (twice kind::tf32).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::tf32: TF32 A and B (format code 2 at bits [7,10) and [10,13)), fp32 accumulator.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <typename TA>
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return sizeof(TA) == 4 ? make_idesc_tf32(M, N, a_mn_major, b_mn_major)
                           : (std::is_same<TA, __half>::value ? make_idesc_f16(M, N, a_mn_major, b_mn_major)
                                                              : make_idesc_bf16(M, N, a_mn_major, b_mn_major));
}

// TMEM -> registers: this warp's 32 lanes (lane field of taddr = 32 * (warp_id % 4)), 32 / 16 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// Column sums over the 32 lanes of a warp of an NV-wide row vector held per lane (lane = row):
// after the call, lane L (< NV) holds sum over rows of column L in v[0].  Butterfly transpose-reduce: NV-1 (+1)
// shuffles instead of 5*NV.
template <int NV>
__device__ __forceinline__ float warp_column_sums(float (&v)[NV], int lane) {
    static_assert(NV == 32 || NV == 16, "NV must be 16 or 32");
    int n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n == 1) {  // NV == 16: one lane bit left over -> plain pairwise add
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        } else {
            const bool upper = (lane & off) != 0;
            const int h = n / 2;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < h) {
                    const float send = upper ? v[i] : v[i + h];
                    const float keep = upper ? v[i + h] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            n = h;
        }
    }
    return v[0];
}

}  // namespace umma
}  // namespace b200em
