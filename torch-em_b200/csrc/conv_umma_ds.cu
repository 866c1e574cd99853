// "depth-stacked" tcgen05 implicit-GEMM convolution for 3 x 3 x 3 filters with few output channels (Cout <= 80).
//
// Why (measured, scripts/ubench/umma_fresh.cu / umma_rate.cu on B200): one M=128 x N x K=16 bf16 tcgen05.mma
// with both operands in shared memory costs max(N/2, 32 + N/4) tensor-pipe cycles and ~60-70 cycles of the single issuing
// thread, so a layer issued with N = Cout = 32 (64) runs at <= 40 % (67 %) of the tensor peak however well it is fed, while
// N = 96 reaches 86 % and N >= 128 the full rate.  Here the three DEPTH taps of the filter share one MMA: for an input
// slice s it computes, for every (b, c) tap,
//     [ y[s+1] | y[s] | y[s-1] ]  +=  x_hat[s][v + (b, c)]  *  [ W[a=0,b,c] | W[a=1,b,c] | W[a=2,b,c] ]
// i.e. N = 3*Cout, and the three N-blocks land in the accumulators of three consecutive OUTPUT slices, which are
// contiguous TMEM column blocks of a ring (block of output d = NA-1 - (d mod NA)).  No shifted sums, no shuffles: the
// epilogue is the plain one.  Each input slice is staged in shared memory ONCE per work item (a column of DR output
// slices needs DR + 2 input slices), the whole packed filter stays resident in shared memory.
//
// Work item = 16 (h) x 8 (w) voxel tile x DR output depth slices; persistent grid = #SMs.
// Shared-memory operand layout as in conv_umma.cu (SWIZZLE_NONE K-major core matrices):
//   A stage = one input slice x one 32(16)-channel chunk: [plane j = 8-ch group][hp 0..17][wp 0..9][8 bf16], each plane one
//             TMA tiled load (cp.async.bulk.tensor 5-D; halo outside the volume zero-filled by the TMA unit)
//   B       = [chunk][tap (b,c)][plane j][n = a*Cout + co][8 bf16]   (resident, loaded once by cp.async.bulk)
// Warp roles (DsRoles): 4 epilogue warps + 8 loader warps, or for Cout = 32 / 64 two epilogue column groups (8 warps) + 6
// loader warps; each loader warp owns every nlw-th stage (issue the TMA loads, wait, norm apply in place, publish the
// stage with a plain shared-memory flag); one weight-loader warp; one MMA-issuing thread (branch-free burst per stage).
// The first convolution of the network (Cin = 1) and its weight gradient live at the end of this file.
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200em {

using namespace umma;

namespace {
constexpr int DS_TH = 16, DS_TW = 8;
constexpr int DS_HP = DS_TH + 2, DS_WP = DS_TW + 2;
constexpr int DS_PLANE = 2944;                      // 18 x 10 voxels x 16 B = 2880, padded to a multiple of 128 B (TMA destination alignment)
constexpr int DS_MAX_SMEM = 227 * 1024;
constexpr int DS_MAX_NS = 12;                       // A ring depth cap
constexpr int DS_MAX_NA = 16;                       // accumulator ring blocks cap
}  // namespace

struct ConvDsParams {
    const __nv_bfloat16* x; long long x_ld;
    const float* in_ss;
    const __nv_bfloat16* w;
    const float* bias;
    void* y; long long y_ld;                        // bf16, or fp32 for the h16 instantiation (TO = float)
    int fp16;                                       // operands (x, w) are IEEE fp16: the h16 path of fp32 activations (in_ss, dot_x unused)
    float* sums;
    const __nv_bfloat16* dot_x; long long dot_ld;   // non-null: sums = (sum y, sum y * dot_x) -- the norm-backward reductions
    int N, D, H, W, Cin, Cout;
    int kh, kw, relu;
    int DR, NS, CC, nchunks;
    int nlw;                         // active operand-loader warps = min(8, NS): a loader may never run two ring phases ahead
    int tiles_w, tiles_h, tiles_d;
    long long items;
    int wbytes;
    int use_tma;                     // 1: the haloed tile planes are loaded by cp.async.bulk.tensor (5-D tiled tensor map over x), else cp.async
    int xdepth;                      // dot_x prefetch ring depth (output slices in flight per epilogue thread), 0 without dot_x
    int debug;                       // bring-up switches (env B200EM_DEBUG): 1 no operand loads, 2 no epilogue math, 4 no MMAs, 16 no norm apply
};

__device__ __forceinline__ void ds_coords(const ConvDsParams& p, long long item_, int& n, int& d0, int& h0, int& w0) {
    unsigned item = (unsigned)item_;                 // 32-bit on purpose (64-bit division is a software routine)
    const unsigned tw = item % (unsigned)p.tiles_w; item /= (unsigned)p.tiles_w;
    const unsigned th = item % (unsigned)p.tiles_h; item /= (unsigned)p.tiles_h;
    const unsigned td = item % (unsigned)p.tiles_d; item /= (unsigned)p.tiles_d;
    n = (int)item; d0 = (int)td * p.DR; h0 = (int)th * DS_TH; w0 = (int)tw * DS_TW;
}

// One NV-wide block of output channels of one output slice: bias, ReLU, bf16 store, statistics.
// ACC: statistics go to per-thread accumulators (flushed once per work item); otherwise they are reduced over the
// warp's 32 rows right away and added to the shared-memory sums.
template <int NV, bool ACC, typename TO = __nv_bfloat16>
__device__ __forceinline__ void ds_epilogue_block(uint32_t taddr, const float* __restrict__ bias_s, int relu, bool valid,
                                                  TO* __restrict__ yp, const uint4* xv, bool has_x, float* acc_s,
                                                  float* acc_q, float* __restrict__ s_sums_blk, bool want_sums, int lane) {
    // xv: this row's dot_x values for these NV channels, already in registers (prefetched one slice ahead; zeros if !valid)
    uint32_t raw[NV];
    if constexpr (NV == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
    tmem_ld_wait();
    // bias: 16-byte shared-memory loads (or none at all for the data gradient).  The kernel runs at the shared-memory bandwidth
    // limit -- the tensor core's operand fetch takes ~half of it -- and one 4-byte broadcast load per column was ~9 % of the rest.
    float v[NV];
    if (bias_s) {
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + 4 * q);
            v[4 * q] = b4.x; v[4 * q + 1] = b4.y; v[4 * q + 2] = b4.z; v[4 * q + 3] = b4.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float f = __uint_as_float(raw[i]) + v[i];
        if (relu) f = fmaxf(f, 0.f);
        v[i] = round_as<TO>(f);
    }
    if (valid) {
        if constexpr (sizeof(TO) == 4) {                 // fp32 activations (the first conv of the h16 path)
#pragma unroll
            for (int q = 0; q < NV / 4; ++q)
                *reinterpret_cast<float4*>(yp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < NV / 8; ++q) {
                uint4 o;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
                *reinterpret_cast<uint4*>(yp + 8 * q) = o;
            }
        }
    }
    if (!want_sums) return;
    if constexpr (ACC) {
        // per-thread accumulators, 8 channels at a time (keeps the live register set small)
        if (valid) {
#pragma unroll
            for (int q = 0; q < NV / 8; ++q) {
                float m[8];
                if (has_x) {                           // second statistic = sum y * dot_x
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&xv[q]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __bfloat1622float2(h2[e]);
                        m[2 * e] = f.x;
                        m[2 * e + 1] = f.y;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) m[e] = v[8 * q + e];
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    acc_s[8 * q + e] += v[8 * q + e];
                    acc_q[8 * q + e] = fmaf(v[8 * q + e], m[e], acc_q[8 * q + e]);
                }
            }
        }
        return;
    }
    float s2[NV];
    if (has_x) {
#pragma unroll
        for (int q = 0; q < NV / 8; ++q) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&xv[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h2[e]);
                s2[8 * q + 2 * e] = f.x;
                s2[8 * q + 2 * e + 1] = f.y;
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) { v[i] = valid ? v[i] : 0.f; s2[i] *= v[i]; }
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) { v[i] = valid ? v[i] : 0.f; s2[i] = v[i] * v[i]; }
    }
    const float a1 = warp_column_sums<NV>(v, lane);
    const float a2 = warp_column_sums<NV>(s2, lane);
    if (NV == 32) {
        atomicAdd(&s_sums_blk[2 * lane], a1);
        atomicAdd(&s_sums_blk[2 * lane + 1], a2);
    } else if ((lane & 1) == 0) {
        atomicAdd(&s_sums_blk[2 * (lane >> 1)], a1);
        atomicAdd(&s_sums_blk[2 * (lane >> 1) + 1], a2);
    }
}

template <int NV>
__device__ __forceinline__ void ds_flush_stats(float* acc_s, float* acc_q, float* __restrict__ s_sums_blk, int lane) {
    float a[NV], q[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { a[i] = acc_s[i]; q[i] = acc_q[i]; acc_s[i] = 0.f; acc_q[i] = 0.f; }
    const float a1 = warp_column_sums<NV>(a, lane);
    const float a2 = warp_column_sums<NV>(q, lane);
    if (NV == 32) {
        atomicAdd(&s_sums_blk[2 * lane], a1);
        atomicAdd(&s_sums_blk[2 * lane + 1], a2);
    } else if ((lane & 1) == 0) {
        atomicAdd(&s_sums_blk[2 * (lane >> 1)], a1);
        atomicAdd(&s_sums_blk[2 * (lane >> 1) + 1], a2);
    }
}

// CO = Cout (16..80, multiple of 16), KC_ = K=16 steps per channel chunk (1 or 2): compile-time so that the epilogue's
// channel blocks are static and the single-thread MMA issue loop has immediate operand offsets.
// Warp roles: CO == 64 / 32 splits the epilogue over two groups of four warps (columns [0,32) and [32,64), each with per-thread
// statistics accumulators like the CO = 32 epilogue) and runs six loader warps; every other CO has four epilogue warps and
// eight loader warps.  16 resp. 14 warps, so 128 registers per thread either way.
template <int CO> struct DsRoles {
    static constexpr bool SPLIT = CO == 64 || CO == 32;    // two epilogue column groups
    static constexpr int NEPI = SPLIT ? 8 : 4;             // epilogue warps
    static constexpr int NLW = SPLIT ? 6 : 8;              // loader warps
    static constexpr int W_WLOAD = NEPI + NLW, W_MMA = W_WLOAD + 1;
    static constexpr int THREADS = (W_MMA + 1) * 32;
    static constexpr int CPT = SPLIT ? CO / 2 : CO;        // output columns per epilogue thread
};

template <int CO, int KC_, typename TO = __nv_bfloat16>
__global__ void __launch_bounds__(DsRoles<CO>::THREADS, 1) conv3d_umma_ds_kernel(const ConvDsParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem[];
    using RL = DsRoles<CO>;
    constexpr int DS_THREADS = RL::THREADS, DS_W_WLOAD = RL::W_WLOAD, DS_W_MMA = RL::W_MMA;
    constexpr int N3 = 3 * CO;
    constexpr int J = KC_ * 2;                                  // 8-channel planes per stage
    constexpr int A_STAGE = J * DS_PLANE;
    constexpr int NA = (512 / CO) < DS_MAX_NA ? (512 / CO) : DS_MAX_NA;   // accumulator ring blocks (CO columns each)
    uint8_t* smA = smem;
    uint8_t* smB = smA + p.NS * A_STAGE;
    float* s_bias = reinterpret_cast<float*>(smB + p.wbytes);
    float* s_sums = s_bias + CO;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_sums + 2 * CO);
    uint64_t* a_full = bars;                          // [NS]  plain flags: fill number of the slot (flag_store / flag_wait_eq)
    uint64_t* a_empty = a_full + DS_MAX_NS;           // [NS]  tcgen05.commit
    uint64_t* acc_full = a_empty + DS_MAX_NS;         // [NA]  tcgen05.commit
    uint64_t* acc_empty = acc_full + DS_MAX_NA;       // [NA] plain counters: epilogue-warp completions per ring block (counter_add / counter_wait_ge)
    uint64_t* b_full = acc_empty + DS_MAX_NA;         // [1]   expect_tx (resident filter)
    uint64_t* t_full = b_full + 1;                    // [NS]  TMA tile loads (expect_tx)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_full + DS_MAX_NS);
    uint32_t* s_tap = s_tmem + 2;                     // [9] operand start offset of each (b, c) tap, 16-byte units

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntap = p.kh * p.kw;
    const int ph = p.kh / 2, pw = p.kw / 2;
    const int DR = p.DR, NS = p.NS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) { flag_init(&a_full[i]); mbar_init(&a_empty[i], 1); mbar_init(&t_full[i], 1); }
        for (int i = 0; i < NA; ++i) { mbar_init(&acc_full[i], 1); flag_init(&acc_empty[i]); }
        mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == DS_W_MMA) tmem_alloc(s_tmem, 512);
    if (threadIdx.x < ntap) {
        const int b = threadIdx.x / p.kw, cc = threadIdx.x % p.kw;
        s_tap[threadIdx.x] = (uint32_t)((((b + 1 - ph) * DS_WP + (cc + 1 - pw)) * 16) >> 4);
    }
    for (int i = threadIdx.x; i < CO; i += DS_THREADS) {
        s_bias[i] = p.bias ? p.bias[i] : 0.f;
        s_sums[2 * i] = 0.f;
        s_sums[2 * i + 1] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp >= RL::NEPI && warp < DS_W_WLOAD) {
        // ===================== operand loaders: warp w8 owns the stages q = w8 (mod nlw) =====================
        // nlw <= NS: the ring's parity waits are only valid while a waiter is at most one phase ahead of its barrier, which
        // holds when a warp's next stage (q + nlw) reuses a slot whose previous use (q + nlw - NS) precedes q.
        // Per work item each lane precomputes the in-slice element offset of its units once (only the slice changes from
        // stage to stage); the norm apply is a short in-place shared-memory pass.
        const int w8 = warp - RL::NEPI;
        constexpr int UNITS = DS_HP * DS_WP * J;       // 16-byte units per stage
        constexpr int NU = (UNITS + 31) / 32;          // per lane
        constexpr int VSTEP = 32 / J;                  // voxels per 32 units
        const int j = lane / VSTEP;                    // this lane's 8-channel plane: a quarter (half) warp = consecutive voxels of
        const int vl = lane % VSTEP;                   // ONE plane = whole 128-byte shared-memory lines
        float sc[8], sh[8];
        int cur_n = -1, cur_c = -1;
        uint32_t slot = 0, phase = 1, lap = 1;         // ring position of the current stage (phase = parity to wait for on a_empty, lap = fill number)
        int owner = 0;                                 // loader warp that owns the current stage (round robin over p.nlw warps)
        const size_t slice_elems = (size_t)p.H * p.W * p.x_ld;
        const bool prof = (p.debug & 8) != 0;
        long long pf_wait = 0, pf_load = 0, pf_norm = 0, pf_t0 = clock64(), pf_n = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, d0, h0, w0;
            ds_coords(p, item, n, d0, h0, w0);
            int goff[NU];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const int v = vl + VSTEP * i;
                const int hp_ = v / DS_WP, wp_ = v % DS_WP;
                const int gh = h0 + hp_ - 1, gw = w0 + wp_ - 1;
                const bool in = v < DS_HP * DS_WP && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
                goff[i] = in ? (int)((gh * p.W + gw) * p.x_ld) : -1;
            }
            for (int s = 0; s < DR + 2; ++s) {
                const int gd = d0 - 1 + s;
                const bool in_d = gd >= 0 && gd < p.D;
                const __nv_bfloat16* xs = p.x + ((size_t)n * p.D + (in_d ? gd : 0)) * slice_elems + j * 8;
                for (int c = 0; c < p.nchunks; ++c) {
                    const uint32_t my_slot = slot, my_phase = phase, my_lap = lap;
                    const bool mine = owner == w8;
                    if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1; ++lap; }
                    if (++owner == p.nlw) owner = 0;
                    if (!mine) continue;
                    long long pf_a = 0, pf_b = 0, pf_c = 0;
                    if (prof) pf_a = clock64();
                    mbar_wait(&a_empty[my_slot], my_phase);
                    if (prof) pf_b = clock64();
                    uint8_t* dst = smA + my_slot * A_STAGE + j * DS_PLANE + vl * 16;
                    const uint32_t dst32 = smem_u32(dst);
                    const __nv_bfloat16* xc = xs + c * p.CC;
                    if (p.use_tma) {
                        // one tiled TMA load per 8-channel plane: box (8 ch, 10 w, 18 h, 1 d, 1 n) lands as [hp][wp][8 ch];
                        // coordinates outside the volume (halo, slices d0-1 / beyond D) are zero-filled by the TMA unit
                        if (lane == 0) {
                            mbar_arrive_expect_tx(&t_full[my_slot], (uint32_t)(J * DS_HP * DS_WP * 16));
                            const uint32_t bar32 = smem_u32(&t_full[my_slot]);
#pragma unroll
                            for (int jj = 0; jj < J; ++jj)
                                asm volatile(
                                    "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
                                        "r"(smem_u32(smA + my_slot * A_STAGE + jj * DS_PLANE)),
                                    "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c * p.CC + jj * 8), "r"(w0 - 1), "r"(h0 - 1), "r"(gd), "r"(n), "r"(bar32)
                                    : "memory");
                        }
                        mbar_wait(&t_full[my_slot], (my_lap - 1) & 1);
                    } else {
                        if (!(p.debug & 1)) {
#pragma unroll
                            for (int i = 0; i < NU; ++i) {
                                if (vl + VSTEP * i < DS_HP * DS_WP) {
                                    const bool in = in_d && goff[i] >= 0;
                                    const __nv_bfloat16* src = in ? xc + goff[i] : p.x;
                                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst32 + (uint32_t)(VSTEP * i * 16)),
                                                 "l"(src), "r"(in ? 16 : 0)
                                                 : "memory");
                                }
                            }
                        }
                        asm volatile("cp.async.wait_all;" ::: "memory");
                    }
                    if (prof) pf_c = clock64();
                    if (p.in_ss && in_d && !(p.debug & 16)) {
                        if (n != cur_n || c != cur_c) {
                            const float* qq = p.in_ss + ((size_t)n * p.Cin + c * p.CC + j * 8) * 2;
#pragma unroll
                            for (int e = 0; e < 8; ++e) { sc[e] = qq[2 * e]; sh[e] = qq[2 * e + 1]; }
                            cur_n = n; cur_c = c;
                        }
#pragma unroll
                        for (int i = 0; i < NU; ++i) {
                            if (goff[i] < 0) continue;
                            uint4* qd = reinterpret_cast<uint4*>(dst + VSTEP * i * 16);
                            uint4 val = *qd;
                            uint32_t* w32 = reinterpret_cast<uint32_t*>(&val);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {       // fp32 math on the unpacked pair, one rounding to bf16
                                const float lo = fmaf(__uint_as_float(w32[e] << 16), sc[2 * e], sh[2 * e]);
                                const float hi = fmaf(__uint_as_float(w32[e] & 0xffff0000u), sc[2 * e + 1], sh[2 * e + 1]);
                                const __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
                                w32[e] = *reinterpret_cast<const uint32_t*>(&r);
                            }
                            *qd = val;
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) flag_store(&a_full[my_slot], my_lap);
                    if (prof) { pf_wait += pf_b - pf_a; pf_load += pf_c - pf_b; pf_norm += clock64() - pf_c; ++pf_n; }
                }
            }
        }
        if (prof && blockIdx.x == 0 && lane == 0 && w8 < 2)
            printf("[ds prof] loader warp %d: total %lld cyc, %lld stages: wait a_empty %lld, load %lld, norm+arrive %lld\n", w8,
                   clock64() - pf_t0, pf_n, pf_wait, pf_load, pf_norm);
    } else if (warp == DS_W_WLOAD) {
        // ===================== weight loader: the whole packed filter, once =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(b_full, (uint32_t)p.wbytes);
            const uint32_t piece = (uint32_t)(J * N3 * 16);              // one (chunk, tap) block
            const int npieces = p.nchunks * ntap;
            for (int g = 0; g < npieces; ++g)
                bulk_g2s(smB + (size_t)g * piece, reinterpret_cast<const uint8_t*>(p.w) + (size_t)g * piece, piece, b_full);
        }
    } else if (warp == DS_W_MMA) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc1 = p.fp16 ? make_idesc_f16(128, CO) : make_idesc_bf16(128, CO),
                           idesc2 = p.fp16 ? make_idesc_f16(128, 2 * CO) : make_idesc_bf16(128, 2 * CO),
                           idesc3 = p.fp16 ? make_idesc_f16(128, N3) : make_idesc_bf16(128, N3);
            const uint64_t ad = make_desc(0, DS_PLANE, DS_WP * 16), bd = make_desc(0, (uint32_t)(N3 * 16), 128);
            const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
            const uint32_t a_lo_base = (uint32_t)(ad & 0xFFFFFFFFu) + (smem_u32(smA) >> 4);
            const uint32_t b_lo_base = (uint32_t)(bd & 0xFFFFFFFFu) + (smem_u32(smB) >> 4);
            constexpr uint32_t K16 = 2 * (DS_PLANE / 16), BK16 = 2 * N3, BTAP16 = J * N3, ASTAGE16 = A_STAGE / 16;
            auto idesc_of = [&](int nb) { return nb == 1 ? idesc1 : (nb == 2 ? idesc2 : idesc3); };
            auto TAPO = [](int t) { return (uint32_t)((t / 3) * DS_WP + (t % 3)); };   // (b, c) tap start offset, 16-byte units (3 x 3 taps)
            const int DRq = DR / NA, DRr = DR % NA;
            mbar_wait(b_full, 0);
            tc_fence_after();
            // Ring bookkeeping is incremental (no divisions in the issue loop): (slot, lap) = A stage and its fill number;
            // (r0, w0) = (oc % NA, (oc / NA) & 1) of the item's first output slice; the block of output counter oc is NA-1 - oc % NA.
            uint32_t slot = 0, lap = 1;
            int r0 = 0;
            uint32_t w0 = 0;
            uint32_t started = 0;                        // output slices this CTA has started accumulating
            const bool prof = (p.debug & 8) != 0;
            long long pf_wa = 0, pf_wacc = 0, pf_t0 = clock64(), pf_n = 0;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
                int r = r0;                              // ring position of output od = s (the newest output slice s feeds)
                uint32_t wpar = w0;
                for (int s = 0; s < DR + 2; ++s) {
                    // ---- fast path: an interior slice (feeds three outputs, starts a new one) whose three ring blocks do not wrap.
                    // The issuing thread's bookkeeping is NOT hidden behind the MMAs, so the common case is kept to
                    // a handful of instructions: one counter probe, one flag probe and one fence per stage, constants elsewhere.
                    if (s >= 2 && s <= DR - 1 && r >= 2 && !prof && !(p.debug & 4)) {
                        const int k0 = NA - 1 - r;       // blocks k0, k0+1, k0+2 = outputs s, s-1, s-2
                        if (started >= (uint32_t)NA) counter_wait_ge(&acc_empty[k0], (uint32_t)RL::NEPI * (started / (uint32_t)NA));
                        ++started;
                        const uint32_t col1 = tmem_base + (uint32_t)(k0 * CO);
                        uint32_t bch = b_lo_base;
                        for (int c = 0; c < p.nchunks; ++c, bch += 9u * BTAP16) {
                            flag_wait_eq(&a_full[slot], lap);
                            tc_fence_after();
                            const uint32_t abuf = a_lo_base + slot * ASTAGE16;
                            if (c == 0) {
                                umma_bf16_c<false>(col1, abuf + TAPO(0), a_hi, bch, b_hi, idesc1);
                                umma_bf16_c<true>(col1 + CO, abuf + TAPO(0), a_hi, bch + CO, b_hi, idesc2);
                            } else {
                                umma_bf16_c<true>(col1, abuf + TAPO(0), a_hi, bch, b_hi, idesc3);
                            }
#pragma unroll
                            for (int t = 0; t < 9; ++t)
#pragma unroll
                                for (int k = 0; k < KC_; ++k)
                                    if (t + k > 0) umma_bf16_c<true>(col1, abuf + TAPO(t) + k * K16, a_hi, bch + t * BTAP16 + k * BK16, b_hi, idesc3);
                            umma_commit(&a_empty[slot]);
                            if (++slot == (uint32_t)NS) { slot = 0; ++lap; }
                        }
                        umma_commit(&acc_full[k0 + 2]);  // output s - 2 = ring position r - 2
                        if (++r == NA) { r = 0; wpar ^= 1; }
                        continue;
                    }
                    const int a_lo = s - (DR - 1) > 0 ? s - (DR - 1) : 0;
                    const int nblk = (s < 2 ? s : 2) - a_lo + 1;
                    int rf = r - a_lo;                   // ring position of the first (a = a_lo) output this slice feeds
                    if (rf < 0) rf += NA;
                    const int k0 = NA - 1 - rf;          // a ascending = columns ascending
                    if (a_lo == 0) {                     // a new accumulator (output od = s) starts with this slice
                        long long t_ = 0;
                        if (prof) t_ = clock64();
                        // its ring block is free once every epilogue warp is done with the block's previous use
                        if (started >= (uint32_t)NA) counter_wait_ge(&acc_empty[k0], (uint32_t)RL::NEPI * (started / (uint32_t)NA));
                        ++started;
                        if (prof) pf_wacc += clock64() - t_;
                        tc_fence_after();
                    }
                    // segments of consecutive ring blocks: [k0, k0+n1) and, past the wrap, [0, n2)
                    const int n1 = nblk < NA - k0 ? nblk : NA - k0;
                    const int n2 = nblk - n1;
                    const uint32_t col1 = tmem_base + (uint32_t)(k0 * CO), col2 = tmem_base;
                    const uint32_t id1 = idesc_of(n1), id2 = idesc_of(n2 > 0 ? n2 : 1);
                    const uint32_t boff2 = (uint32_t)(n1 * CO);
                    uint32_t bch = b_lo_base + (uint32_t)(a_lo * CO);     // 16-byte units: n rows are 16 B apart
                    for (int c = 0; c < p.nchunks; ++c, bch += (uint32_t)ntap * BTAP16) {
                        long long t_ = 0;
                        if (prof) t_ = clock64();
                        flag_wait_eq(&a_full[slot], lap);
                        if (prof) { pf_wa += clock64() - t_; ++pf_n; }
                        tc_fence_after();
                        const uint32_t abuf = a_lo_base + slot * ASTAGE16;
                        if (!(p.debug & 4)) {
                            // (t = 0, k = 0) first: it overwrites block k0 when this slice starts a new accumulator
                            const uint32_t a0 = abuf + TAPO(0);
                            if (c == 0 && a_lo == 0) {
                                umma_bf16_c<false>(col1, a0, a_hi, bch, b_hi, idesc1);
                                if (nblk > 1) {
                                    // the other blocks a = 1..  accumulate; they start at ring block k0+1 (may wrap)
                                    const int rr = nblk - 1;
                                    const int kk = (k0 + 1 == NA) ? 0 : k0 + 1;
                                    const int r1 = rr < NA - kk ? rr : NA - kk;
                                    umma_bf16_c<true>(tmem_base + (uint32_t)(kk * CO), a0, a_hi, bch + CO, b_hi, idesc_of(r1));
                                    if (rr - r1 > 0)
                                        umma_bf16_c<true>(tmem_base, a0, a_hi, bch + CO + (uint32_t)(r1 * CO), b_hi, idesc_of(rr - r1));
                                }
                            } else {
                                umma_bf16_c<true>(col1, a0, a_hi, bch, b_hi, id1);
                                if (n2 > 0) umma_bf16_c<true>(col2, a0, a_hi, bch + boff2, b_hi, id2);
                            }
                            // the rest of the stage is branch-free straight-line code with immediate operand offsets: the descriptor
                            // halves, instruction descriptors and TMEM columns then stay in uniform registers across the burst
                            // (every branch inside the burst costs a re-materialisation of ~8 R2UR per tap, ~40 cycles per MMA)
                            if (n2 == 0) {
#pragma unroll
                                for (int t = 0; t < 9; ++t)
#pragma unroll
                                    for (int k = 0; k < KC_; ++k)
                                        if (t + k > 0) umma_bf16_c<true>(col1, abuf + TAPO(t) + k * K16, a_hi, bch + t * BTAP16 + k * BK16, b_hi, id1);
                            } else {
#pragma unroll
                                for (int t = 0; t < 9; ++t)
#pragma unroll
                                    for (int k = 0; k < KC_; ++k)
                                        if (t + k > 0) {
                                            umma_bf16_c<true>(col1, abuf + TAPO(t) + k * K16, a_hi, bch + t * BTAP16 + k * BK16, b_hi, id1);
                                            umma_bf16_c<true>(col2, abuf + TAPO(t) + k * K16, a_hi, bch + t * BTAP16 + k * BK16 + boff2, b_hi, id2);
                                        }
                            }
                        }
                        umma_commit(&a_empty[slot]);
                        if (++slot == (uint32_t)NS) { slot = 0; ++lap; }
                    }
                    if (s >= 2) {                        // output od = s - 2 has received its last contribution
                        int r2 = r - 2;
                        if (r2 < 0) r2 += NA;
                        umma_commit(&acc_full[NA - 1 - r2]);
                    }
                    if (++r == NA) { r = 0; wpar ^= 1; }
                }
                r0 += DRr; w0 ^= (uint32_t)(DRq & 1);
                if (r0 >= NA) { r0 -= NA; w0 ^= 1; }
            }
            if (prof && blockIdx.x == 0)
                printf("[ds prof] mma: total %lld cyc, %lld stages: wait a_full %lld, wait acc_empty %lld\n", clock64() - pf_t0, pf_n, pf_wa, pf_wacc);
        }
    } else {
        // ===================== epilogue (warps 0..NEPI-1; warp % 4 = TMEM lane quadrant, warp / 4 = column group) ============
        const int wq = warp & 3, grp = warp >> 2;
        const int row = wq * 32 + lane;              // GEMM row = TMEM lane
        const int hl = row / DS_TW, wl = row % DS_TW;
        constexpr int CPT = RL::CPT;
        const int col0 = grp * CPT;                  // first output column of this thread
        constexpr bool kAcc = CPT <= 32;
        constexpr int NAcc = kAcc ? CPT : 1;
        float acc_s[NAcc], acc_q[NAcc];
#pragma unroll
        for (int i = 0; i < NAcc; ++i) { acc_s[i] = 0.f; acc_q[i] = 0.f; }
        const bool ws = p.sums != nullptr;
        const bool has_x = p.dot_x != nullptr;
        constexpr int XV = CPT / 8;                  // 16-byte vectors of dot_x per thread
        int r = 0;                                   // (oc % NA, (oc / NA) & 1) of the next output slice
        uint32_t wpar = 0;
        const bool prof = (p.debug & 8) != 0;
        long long pf_w = 0, pf_t0 = clock64(), pf_n = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, d0, h0, w0;
            ds_coords(p, item, n, d0, h0, w0);
            const int gh = h0 + hl, gw = w0 + wl;
            const bool valid_hw = gh < p.H && gw < p.W;
            // this row's voxel in slice d0 (clamped when the row is outside the volume: never dereferenced then)
            const size_t vox0 = (((size_t)n * p.D + d0) * p.H + (valid_hw ? gh : 0)) * p.W + (valid_hw ? gw : 0);
            const size_t hw = (size_t)p.H * p.W;
            // dot_x rows are prefetched into registers two output slices ahead (one when a thread owns 32 columns or more:
            // register budget): a load issued when the accumulator is ready would expose its full latency on every slice, and
            // staging them through shared memory costs LSU wavefronts (8x the ideal for this 16-byte scatter) that the tensor
            // pipe's operand fetch needs.
            constexpr bool DBL = XV <= 2;
            uint4 xa[XV], xb[DBL ? XV : 1];
            auto load_x = [&](uint4* dst, int od_) {
                const bool ok = has_x && valid_hw && od_ < DR && d0 + od_ < p.D;
                const __nv_bfloat16* q_ = p.dot_x + (vox0 + (size_t)od_ * hw) * p.dot_ld + col0;
#pragma unroll
                for (int i = 0; i < XV; ++i) dst[i] = ok ? __ldg(reinterpret_cast<const uint4*>(q_ + 8 * i)) : make_uint4(0, 0, 0, 0);
            };
            auto process = [&](int od, const uint4* xcur) {
                const int blk = NA - 1 - r;
                long long t_ = 0;
                if (prof) t_ = clock64();
                mbar_wait(&acc_full[blk], wpar);
                if (prof) { pf_w += clock64() - t_; ++pf_n; }
                tc_fence_after();
                const int gd = d0 + od;
                const bool valid = valid_hw && gd < p.D;
                TO* yp = reinterpret_cast<TO*>(p.y) + (vox0 + (size_t)(gd < p.D ? od : 0) * hw) * p.y_ld + col0;
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(blk * CO + col0);
                const float* s_biasg = p.bias ? s_bias + col0 : nullptr;
                const float* const s_bias0 = p.bias ? s_bias : nullptr;     // (an absent bias costs no shared-memory loads)
                float* s_sumsg = s_sums + 2 * col0;
                if (p.debug & 2) {
                } else if constexpr (CPT == 16) {
                    ds_epilogue_block<16, true, TO>(taddr, s_biasg, p.relu, valid, yp, xcur, has_x, acc_s, acc_q, s_sumsg, ws, lane);
                } else if constexpr (CPT == 32) {
                    // two 16-column passes: half the live registers of one 32-column pass (the 64 statistics accumulators stay)
                    ds_epilogue_block<16, true, TO>(taddr, s_biasg, p.relu, valid, yp, xcur, has_x, acc_s, acc_q, s_sumsg, ws, lane);
                    ds_epilogue_block<16, true, TO>(taddr + 16, s_biasg ? s_biasg + 16 : nullptr, p.relu, valid, yp + 16, xcur + 2, has_x, acc_s + 16, acc_q + 16,
                                                s_sumsg, ws, lane);
                } else {
                    ds_epilogue_block<32, false, TO>(taddr, s_bias0, p.relu, valid, yp, xcur, has_x, acc_s, acc_q, s_sums, ws, lane);
                    if constexpr (CO == 48)
                        ds_epilogue_block<16, false, TO>(taddr + 32, s_bias0 ? s_bias0 + 32 : nullptr, p.relu, valid, yp + 32, xcur + 4, has_x, acc_s, acc_q,
                                                     s_sums + 64, ws, lane);
                    if constexpr (CO == 80) {
                        ds_epilogue_block<32, false, TO>(taddr + 32, s_bias0 ? s_bias0 + 32 : nullptr, p.relu, valid, yp + 32, xcur + 4, has_x, acc_s, acc_q,
                                                     s_sums + 64, ws, lane);
                        ds_epilogue_block<16, false, TO>(taddr + 64, s_bias0 ? s_bias0 + 64 : nullptr, p.relu, valid, yp + 64, xcur + 8, has_x, acc_s, acc_q,
                                                     s_sums + 128, ws, lane);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) counter_add(&acc_empty[blk]);
                if (++r == NA) { r = 0; wpar ^= 1; }
            };
            load_x(xa, 0);
            if constexpr (DBL) load_x(xb, 1);
            for (int od = 0; od < DR; od += (DBL ? 2 : 1)) {
                process(od, xa);
                load_x(xa, od + (DBL ? 2 : 1));
                if constexpr (DBL) {
                    if (od + 1 < DR) {
                        process(od + 1, xb);
                        load_x(xb, od + 3);
                    }
                }
            }
            if (ws) {
                // flush this item's per-channel partial sums (the sample n may change with the next item)
                if constexpr (CPT == 32) {
                    ds_flush_stats<16>(acc_s, acc_q, s_sums + 2 * col0, lane);
                    ds_flush_stats<16>(acc_s + 16, acc_q + 16, s_sums + 2 * col0 + 32, lane);
                } else if constexpr (kAcc) {
                    ds_flush_stats<CPT>(acc_s, acc_q, s_sums + 2 * col0, lane);
                }
                // each column group flushes its own channels behind its own named barrier (ids 1 and 3)
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
                for (int i = (int)(threadIdx.x & 127); i < 2 * CPT; i += 128) {
                    const int ii = 2 * col0 + i;
                    atomicAdd(p.sums + ((size_t)n * CO + (ii >> 1)) * 2 + (ii & 1), s_sums[ii]);
                    s_sums[ii] = 0.f;
                }
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
            }
        }
        if (prof && blockIdx.x == 0 && threadIdx.x == 0)
            printf("[ds prof] epilogue: total %lld cyc, %lld slabs: wait acc_full %lld\n", clock64() - pf_t0, pf_n, pf_w);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == DS_W_MMA) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// torch (Cout, Cin, 3, kh, kw) fp32 -> bf16 [chunk][tap (b,c)][plane j][a*Nc + n][8]; dgrad = 1: transposed, tap-flipped.
__global__ void pack_ds_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw, int dgrad, int CC,
                                       __nv_bfloat16* __restrict__ out) {
    const int thw = kh * kw, taps = 3 * thw;
    const int64_t total = (int64_t)Cout * Cin * taps;
    const int Nc = dgrad ? Cin : Cout;               // output channels of the packed operand
    const int J = CC / 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int tp = (int)(i % taps);
        const int ci = (int)((i / taps) % Cin);
        const int co = (int)(i / ((int64_t)taps * Cin));
        const int n_ = dgrad ? ci : co, k_ = dgrad ? co : ci, t_ = dgrad ? taps - 1 - tp : tp;
        const int a = t_ / thw, bc = t_ % thw;
        const int chunk = k_ / CC, j = (k_ % CC) / 8, e = k_ % 8;
        out[((((size_t)chunk * thw + bc) * J + j) * (3 * Nc) + (size_t)a * Nc + n_) * 8 + e] = __float2bfloat16_rn(w[i]);
    }
}


// ===============================================================================================================
// First convolution of the network (Cin = 1, 3x3x3: K = 27, too thin for the implicit GEMM above).  The im2col rows
// are built ON THE FLY in shared memory -- 27 neighbours of each voxel gathered from a small haloed 1-channel tile,
// zero-padded to K = 32 -- and consumed by two tcgen05.mma (K = 16 each) per 128-voxel slab; nothing but the input
// (2 B/voxel) is read and nothing but the output is written.  Same epilogue as the kernels above.  The weight-gradient
// kernel below builds the identical shared-memory image and uses it MN-major (K = voxels).
namespace {
constexpr int F1_NLW = 8;
constexpr int F1_THREADS = 128 + F1_NLW * 32 + 32;
constexpr int F1_W_MMA = 4 + F1_NLW;
constexpr int F1_NS = 8;                              // A ring depth (8 KB per stage)
constexpr int F1_ASTAGE = 4 * 128 * 16;               // [4 k-groups of 8 taps][128 voxels][8 bf16]
constexpr int F1_SCR = 3 * DS_HP * DS_WP;             // per-warp scratch: haloed 1-channel tile of three slices (bf16)
}  // namespace

struct ConvFirstParams {
    const void* x;                                    // (N, D, H, W, 1) bf16, or fp32 for the F32 instantiation
    const float* in_ss;                               // (N, 1, 2) or null
    const float* w;                                   // torch (Cout, 1, 3, 3, 3) fp32
    const float* bias;
    void* y; long long y_ld;                          // bf16, or fp32 for the F32 instantiation
    float* sums;
    const __nv_bfloat16* dz; long long dz_ld;         // weight-gradient kernel only
    float* dw; float* db;
    int N, D, H, W, Cout, relu, DR;
    int tiles_w, tiles_h, tiles_d;
    long long items;
};

__device__ __forceinline__ void f1_coords(const ConvFirstParams& p, long long item_, int& n, int& d0, int& h0, int& w0) {
    unsigned item = (unsigned)item_;
    const unsigned tw = item % (unsigned)p.tiles_w; item /= (unsigned)p.tiles_w;
    const unsigned th = item % (unsigned)p.tiles_h; item /= (unsigned)p.tiles_h;
    const unsigned td = item % (unsigned)p.tiles_d; item /= (unsigned)p.tiles_d;
    n = (int)item; d0 = (int)td * p.DR; h0 = (int)th * DS_TH; w0 = (int)tw * DS_TW;
}

// One loader warp builds the im2col image of one 128-voxel slab (output slice gd, tile origin h0, w0) at `dst`:
// [k-group j][voxel m = 8*hl + wl][8 taps], tap k = (a*3 + b)*3 + c, k >= 27 zero.  scr = this warp's scratch.
// F32: x is fp32 and the 16-bit im2col image is IEEE fp16 (the h16 path: same significand as TF32), else bf16 from bf16.
template <bool F32 = false>
__device__ __forceinline__ void f1_build_slab(const ConvFirstParams& p, int n, int gd, int h0, int w0, float sc, float sh,
                                              __nv_bfloat16* scr, uint8_t* dst, int lane) {
    using TX = typename std::conditional<F32, float, __nv_bfloat16>::type;
    const TX* xn = reinterpret_cast<const TX*>(p.x) + (size_t)n * p.D * p.H * p.W;
    // all of the lane's loads are issued before the first one is consumed (one global latency per slab, not 17)
    constexpr int NF = (F1_SCR + 31) / 32;
    TX raw[NF];
    bool inb[NF];
#pragma unroll
    for (int q = 0; q < NF; ++q) {
        const int i = lane + 32 * q;
        const int wp_ = i % DS_WP, hp_ = (i / DS_WP) % DS_HP, a = i / (DS_WP * DS_HP);
        const int sd = gd + a - 1, gh = h0 + hp_ - 1, gw = w0 + wp_ - 1;
        inb[q] = i < F1_SCR && sd >= 0 && sd < p.D && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
        raw[q] = inb[q] ? xn[((size_t)sd * p.H + gh) * p.W + gw] : from_f<TX>(0.f);
    }
#pragma unroll
    for (int q = 0; q < NF; ++q) {
        const int i = lane + 32 * q;
        const float xh = inb[q] ? fmaf(to_f<TX>(raw[q]), sc, sh) : 0.f;
        if (i < F1_SCR) {
            if constexpr (F32) reinterpret_cast<__half*>(scr)[i] = __float2half_rn(xh);
            else scr[i] = __float2bfloat16_rn(xh);
        }
    }
    __syncwarp();
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(scr);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int m = lane + 32 * q;
        const int hl = m / DS_TW, wl = m % DS_TW;
        uint32_t pk[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            uint32_t lo = 0, hi = 0;
            const int k0 = 2 * k2, k1 = 2 * k2 + 1;
            if (k0 < 27) lo = s16[((k0 / 9) * DS_HP + hl + (k0 / 3) % 3) * DS_WP + wl + k0 % 3];
            if (k1 < 27) hi = s16[((k1 / 9) * DS_HP + hl + (k1 / 3) % 3) * DS_WP + wl + k1 % 3];
            pk[k2] = lo | (hi << 16);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(dst + j * (128 * 16) + m * 16) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    __syncwarp();                                     // scratch is reused by this warp's next slab
}

// EG = epilogue warp groups: with EG = 2 the warps w and w + 4 share a TMEM lane quarter and split the CO output columns.  One
// warp's chain of ~330 dependent instructions per slab (TMEM load, bias, ReLU, rounding, statistics, store) is what bounds the
// kernel with EG = 1 (ncu source page: the loader warps sit in their a_empty wait); two groups halve it, the eight loader warps stay.
template <int CO, bool F32 = false, int EG = 1, int NLW = F1_NLW>
__global__ void __launch_bounds__(EG * 128 + NLW * 32 + 32, 1) conv3d_first_kernel(const ConvFirstParams p) {
    // (EG = 2: 17 warps put five on one SM sub-partition = 96 registers per thread and 32 bytes of spills; seven loader warps instead
    // -- 128 registers, no spills -- measured 412 us against 300: the loaders' round robin wants NLW == the ring depth)
    constexpr int W_LOAD0 = 4 * EG, W_MMA = 4 * EG + NLW;
    using TY = typename std::conditional<F32, float, __nv_bfloat16>::type;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NA = (512 / CO) < DS_MAX_NA ? (512 / CO) : DS_MAX_NA;
    uint8_t* smA = smem;                                              // [F1_NS][F1_ASTAGE]
    uint8_t* smB = smA + F1_NS * F1_ASTAGE;                           // [4 k-groups][CO][8 bf16]
    __nv_bfloat16* s_scr = reinterpret_cast<__nv_bfloat16*>(smB + 4 * CO * 16);   // [F1_NLW][F1_SCR]
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_scr) + ((F1_NLW * F1_SCR * 2 + 15) & ~15));
    float* s_sums = s_bias + CO;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_sums + 2 * CO);
    uint64_t* a_full = bars;                          // [F1_NS]
    uint64_t* a_empty = a_full + F1_NS;               // [F1_NS]
    uint64_t* acc_full = a_empty + F1_NS;             // [NA]
    uint64_t* acc_empty = acc_full + DS_MAX_NA;       // [NA]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + DS_MAX_NA);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int DR = p.DR;
    if (threadIdx.x == 0) {
        for (int i = 0; i < F1_NS; ++i) { flag_init(&a_full[i]); mbar_init(&a_empty[i], 1); }      // a_full: plain fill-number flags
        for (int i = 0; i < NA; ++i) { mbar_init(&acc_full[i], 1); flag_init(&acc_empty[i]); }    // acc_empty: plain completion counters
        fence_mbar_init();
    }
    if (warp == W_MMA) tmem_alloc(s_tmem, 512);
    // weight operand straight from the fp32 parameter: B[k-group j][co][8 taps], taps >= 27 zero
    for (int i = threadIdx.x; i < 4 * CO * 8; i += blockDim.x) {
        const int e = i % 8, co = (i / 8) % CO, j = i / (8 * CO);
        const int k = j * 8 + e;
        const float wv = k < 27 ? p.w[co * 27 + k] : 0.f;
        if constexpr (F32) reinterpret_cast<__half*>(smB)[i] = __float2half_rn(wv);
        else reinterpret_cast<__nv_bfloat16*>(smB)[i] = __float2bfloat16_rn(wv);
    }
    for (int i = threadIdx.x; i < CO; i += blockDim.x) {
        s_bias[i] = p.bias ? p.bias[i] : 0.f;
        s_sums[2 * i] = 0.f;
        s_sums[2 * i + 1] = 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp >= W_LOAD0 && warp < W_MMA) {
        // ===================== loaders: warp w8 builds every 8th slab =====================
        const int w8 = warp - W_LOAD0;
        __nv_bfloat16* scr = s_scr + w8 * F1_SCR;
        uint32_t slot = 0, phase = 1, lap = 1;
        int owner = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, d0, h0, w0;
            f1_coords(p, item, n, d0, h0, w0);
            const float sc = p.in_ss ? p.in_ss[n * 2] : 1.f, sh = p.in_ss ? p.in_ss[n * 2 + 1] : 0.f;
            for (int od = 0; od < DR; ++od) {
                const uint32_t my_slot = slot, my_phase = phase, my_lap = lap;
                const bool mine = owner == w8;
                if (++slot == F1_NS) { slot = 0; phase ^= 1; ++lap; }
                if (++owner == NLW) owner = 0;
                if (!mine) continue;
                mbar_wait(&a_empty[my_slot], my_phase);
                f1_build_slab<F32>(p, n, d0 + od, h0, w0, sc, sh, scr, smA + my_slot * F1_ASTAGE, lane);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) flag_store(&a_full[my_slot], my_lap);
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: two K = 16 MMAs per slab =====================
        if (elect_one()) {
            constexpr uint32_t idesc = F32 ? make_idesc_f16(128, CO) : make_idesc_bf16(128, CO);
            const uint64_t ad = make_desc(0, 128 * 16, 128), bd = make_desc(0, (uint32_t)(CO * 16), 128);
            const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
            const uint32_t a_lo_base = (uint32_t)(ad & 0xFFFFFFFFu) + (smem_u32(smA) >> 4);
            const uint32_t b_lo = (uint32_t)(bd & 0xFFFFFFFFu) + (smem_u32(smB) >> 4);
            uint32_t slot = 0, lap = 1, started = 0;
            int r = 0;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
                for (int od = 0; od < DR; ++od) {
                    if (started >= (uint32_t)NA) counter_wait_ge(&acc_empty[r], 4u * EG * (started / (uint32_t)NA));
                    ++started;
                    flag_wait_eq(&a_full[slot], lap);
                    tc_fence_after();
                    const uint32_t a0 = a_lo_base + slot * (F1_ASTAGE / 16);
                    const uint32_t tacc = tmem_base + (uint32_t)(r * CO);
                    umma_bf16_c<false>(tacc, a0, a_hi, b_lo, b_hi, idesc);
                    umma_bf16_c<true>(tacc, a0 + 2 * 128, a_hi, b_lo + 2 * CO, b_hi, idesc);
                    umma_commit(&a_empty[slot]);
                    umma_commit(&acc_full[r]);
                    if (++slot == F1_NS) { slot = 0; ++lap; }
                    if (++r == NA) r = 0;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 0 .. 4 * EG - 1) =====================
        const int wq = warp & 3, grp = warp >> 2;     // TMEM lane quarter of the warp, column group
        const int row = wq * 32 + lane;
        const int hl = row / DS_TW, wl = row % DS_TW;
        constexpr int CPT = CO / EG;                  // columns per epilogue thread
        constexpr bool kAcc = CPT <= 32 && CO <= 32;
        constexpr int NAcc = kAcc ? CPT : 1;
        const int col0 = grp * CPT;
        float acc_s[NAcc], acc_q[NAcc];
#pragma unroll
        for (int i = 0; i < NAcc; ++i) { acc_s[i] = 0.f; acc_q[i] = 0.f; }
        const bool ws = p.sums != nullptr;
        int r = 0;
        uint32_t wpar = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, d0, h0, w0;
            f1_coords(p, item, n, d0, h0, w0);
            const int gh = h0 + hl, gw = w0 + wl;
            const bool valid_hw = gh < p.H && gw < p.W;
            const size_t vox0 = (((size_t)n * p.D + d0) * p.H + (valid_hw ? gh : 0)) * p.W + (valid_hw ? gw : 0);
            const size_t hw = (size_t)p.H * p.W;
            for (int od = 0; od < DR; ++od) {
                mbar_wait(&acc_full[r], wpar);
                tc_fence_after();
                const int gd = d0 + od;
                const bool valid = valid_hw && gd < p.D;
                TY* yp = reinterpret_cast<TY*>(p.y) + (vox0 + (size_t)(gd < p.D ? od : 0) * hw) * p.y_ld;
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(r * CO + col0);
                if constexpr (CPT == 16) {
                    ds_epilogue_block<16, true, TY>(taddr, s_bias + col0, p.relu, valid, yp + col0, nullptr, false, acc_s, acc_q, s_sums, ws, lane);
                } else if constexpr (CO == 32) {
                    ds_epilogue_block<16, true, TY>(taddr, s_bias, p.relu, valid, yp, nullptr, false, acc_s, acc_q, s_sums, ws, lane);
                    ds_epilogue_block<16, true, TY>(taddr + 16, s_bias + 16, p.relu, valid, yp + 16, nullptr, false, acc_s + 16, acc_q + 16, s_sums,
                                                    ws, lane);
                } else {
                    ds_epilogue_block<32, false, TY>(taddr, s_bias, p.relu, valid, yp, nullptr, false, acc_s, acc_q, s_sums, ws, lane);
                    if constexpr (CO == 64)
                        ds_epilogue_block<32, false, TY>(taddr + 32, s_bias + 32, p.relu, valid, yp + 32, nullptr, false, acc_s, acc_q, s_sums + 64,
                                                         ws, lane);
                    if constexpr (CO == 48)
                        ds_epilogue_block<16, false, TY>(taddr + 32, s_bias + 32, p.relu, valid, yp + 32, nullptr, false, acc_s, acc_q, s_sums + 64,
                                                         ws, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) counter_add(&acc_empty[r]);
                if (++r == NA) { r = 0; wpar ^= 1; }
            }
            if (ws) {
                if constexpr (CPT == 32) {
                    ds_flush_stats<16>(acc_s, acc_q, s_sums + 2 * col0, lane);
                    ds_flush_stats<16>(acc_s + 16, acc_q + 16, s_sums + 2 * col0 + 32, lane);
                } else if constexpr (kAcc) {
                    ds_flush_stats<CPT>(acc_s, acc_q, s_sums + 2 * col0, lane);
                }
                // each column group flushes its own channels behind its own named barrier (ids 1 and 3)
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
                for (int i = (int)(threadIdx.x & 127); i < 2 * CPT; i += 128) {
                    const int ii = 2 * col0 + i;
                    atomicAdd(p.sums + ((size_t)n * CO + (ii >> 1)) * 2 + (ii & 1), s_sums[ii]);
                    s_sums[ii] = 0.f;
                }
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Weight (and bias) gradient of the first convolution: dW[co][tap] += sum_v dz[v][co] * x_hat[v + tap].
// A (MN-major) = the im2col image above, M = taps (32 rows used of 128: the rows beyond read whatever follows in shared
// memory and are ignored), B (MN-major) = the dz slab, N = Cout, K = the slab's 128 voxels (8 MMAs); one accumulator
// resident in TMEM for the CTA's whole life, one atomic epilogue.
template <int CO>
__global__ void __launch_bounds__(F1_THREADS, 1) conv3d_first_wgrad_kernel(const ConvFirstParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int ZSTAGE = (CO / 8) * 128 * 16;                       // [jo][128 voxels][8 bf16]
    uint8_t* smA = smem;                                              // [F1_NS][F1_ASTAGE]
    uint8_t* smZ = smA + F1_NS * F1_ASTAGE;                           // [F1_NS][ZSTAGE]  (>= 24 KB: absorbs the M = 128 over-read)
    __nv_bfloat16* s_scr = reinterpret_cast<__nv_bfloat16*>(smZ + F1_NS * ZSTAGE);
    float* s_db = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_scr) + ((F1_NLW * F1_SCR * 2 + 15) & ~15));
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_db + CO);
    uint64_t* full = bars;                            // [F1_NS]
    uint64_t* empty = full + F1_NS;                   // [F1_NS]
    uint64_t* acc_full = empty + F1_NS;               // [1]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int DR = p.DR;
    if (threadIdx.x == 0) {
        for (int i = 0; i < F1_NS; ++i) { flag_init(&full[i]); mbar_init(&empty[i], 1); }          // full: plain fill-number flags
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == F1_W_MMA) tmem_alloc(s_tmem, 512);
    for (int i = threadIdx.x; i < CO; i += F1_THREADS) s_db[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp >= 4 && warp < F1_W_MMA) {
        const int w8 = warp - 4;
        __nv_bfloat16* scr = s_scr + w8 * F1_SCR;
        constexpr int JO = CO / 8;
        float dbacc[JO][8];                            // this lane's voxels (m = lane + 32 q), all channels: reduced at the end
#pragma unroll
        for (int j = 0; j < JO; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) dbacc[j][e] = 0.f;
        uint32_t slot = 0, phase = 1, lap = 1;
        int owner = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, d0, h0, w0;
            f1_coords(p, item, n, d0, h0, w0);
            const float sc = p.in_ss ? p.in_ss[n * 2] : 1.f, sh = p.in_ss ? p.in_ss[n * 2 + 1] : 0.f;
            for (int od = 0; od < DR; ++od) {
                const uint32_t my_slot = slot, my_phase = phase, my_lap = lap;
                const bool mine = owner == w8;
                if (++slot == F1_NS) { slot = 0; phase ^= 1; ++lap; }
                if (++owner == F1_NLW) owner = 0;
                if (!mine) continue;
                mbar_wait(&empty[my_slot], my_phase);
                const int gd = d0 + od;
                uint8_t* zdst = smZ + my_slot * ZSTAGE;
                const uint32_t z32 = smem_u32(zdst);
                // dz slab: plain copies (zero outside the volume), issued first so they fly while the im2col image is built
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int m = lane + 32 * q;
                    const int gh = h0 + m / DS_TW, gw = w0 + m % DS_TW;
                    const bool in = gd < p.D && gh < p.H && gw < p.W;
                    const __nv_bfloat16* src = in ? p.dz + ((((size_t)n * p.D + gd) * p.H + gh) * p.W + gw) * p.dz_ld : p.dz;
#pragma unroll
                    for (int j = 0; j < JO; ++j)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(z32 + (uint32_t)(j * 128 * 16 + m * 16)), "l"(src + 8 * j),
                                     "r"(in ? 16 : 0)
                                     : "memory");
                }
                f1_build_slab(p, n, gd < p.D ? gd : p.D + 1, h0, w0, sc, sh, scr, smA + my_slot * F1_ASTAGE, lane);
                asm volatile("cp.async.wait_all;" ::: "memory");
                if (p.db) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int m = lane + 32 * q;
#pragma unroll
                        for (int j = 0; j < JO; ++j) {
                            const uint4 val = *reinterpret_cast<const uint4*>(zdst + j * 128 * 16 + m * 16);
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&val);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __bfloat1622float2(h2[e]);
                                dbacc[j][2 * e] += f.x;
                                dbacc[j][2 * e + 1] += f.y;
                            }
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) flag_store(&full[my_slot], my_lap);
            }
        }
        if (p.db) {
#pragma unroll
            for (int j = 0; j < JO; ++j)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float a = dbacc[j][e];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (lane == 0) atomicAdd(&s_db[j * 8 + e], a);
                }
            asm volatile("bar.sync 2, %0;" ::"n"(F1_NLW * 32) : "memory");
            const int t = threadIdx.x - 128;
            if (t < CO) atomicAdd(p.db + t, s_db[t]);
        }
    } else if (warp == F1_W_MMA) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, CO, 1, 1);
            // MN-major: LBO = next 8 voxels (128 B), SBO = next 8 taps / channels (one 2 KB plane)
            const uint64_t ad = make_desc(0, 128, 128 * 16), bd = make_desc(0, 128, 128 * 16);
            const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
            const uint32_t a_lo_base = (uint32_t)(ad & 0xFFFFFFFFu) + (smem_u32(smA) >> 4);
            const uint32_t b_lo_base = (uint32_t)(bd & 0xFFFFFFFFu) + (smem_u32(smZ) >> 4);
            uint32_t slot = 0, lap = 1;
            bool first = true;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
                for (int od = 0; od < DR; ++od) {
                    flag_wait_eq(&full[slot], lap);
                    tc_fence_after();
                    const uint32_t a0 = a_lo_base + slot * (F1_ASTAGE / 16), b0 = b_lo_base + slot * (ZSTAGE / 16);
                    if (first) umma_bf16_c<false>(tmem_base, a0, a_hi, b0, b_hi, idesc);
                    else umma_bf16_c<true>(tmem_base, a0, a_hi, b0, b_hi, idesc);
                    first = false;
#pragma unroll
                    for (int ks = 1; ks < 8; ++ks) umma_bf16_c<true>(tmem_base, a0 + ks * 16, a_hi, b0 + ks * 16, b_hi, idesc);
                    umma_commit(&empty[slot]);
                    if (++slot == F1_NS) { slot = 0; ++lap; }
                }
            }
            umma_commit(acc_full);
        }
    } else if (warp == 0) {
        // epilogue: accumulator rows 0..26 = taps (TMEM lanes of warp 0), columns = output channels
        mbar_wait(acc_full, 0);
        tc_fence_after();
        for (int cb = 0; cb < CO; cb += 16) {
            uint32_t raw[16];
            tmem_ld16(tmem_base + (uint32_t)cb, raw);
            tmem_ld_wait();
            if (lane < 27) {
#pragma unroll
                for (int i = 0; i < 16; ++i) atomicAdd(p.dw + (size_t)(cb + i) * 27 + lane, __uint_as_float(raw[i]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == F1_W_MMA) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static bool first_shape(int Cin, int Cout, int kd, int kh, int kw) {
    return Cin == 1 && kd == 3 && kh == 3 && kw == 3 && (Cout == 16 || Cout == 32 || Cout == 48 || Cout == 64);
}

struct DsShape {
    int CC, NS, wbytes, smem_bytes, xdepth;
};

static bool ds_shape(int Cin, int Cout, int kd, int kh, int kw, DsShape& s, bool with_dot = false) {
    if (kd != 3 || kh != 3 || kw != 3) return false;   // the MMA burst is compiled for the 3 x 3 (h, w) taps
    if (Cin % 16 || Cout % 16 || Cin < 16 || Cout < 16 || Cout > 80) return false;
    s.CC = (Cin % 32 == 0) ? 32 : 16;
    const int J = s.CC / 8;
    s.wbytes = 3 * kh * kw * Cin * Cout * 2;
    const int misc = Cout * 4 * 3 + (3 * DS_MAX_NS + 2 * DS_MAX_NA + 1) * 8 + 8 + 9 * 4 + 128;
    const int stage = J * DS_PLANE;
    int ns = (DS_MAX_SMEM - s.wbytes - misc) / stage;
    if (ns > DS_MAX_NS) ns = DS_MAX_NS;
    if (ns < 6) return false;                        // the filter must stay resident next to a useful operand ring
    s.xdepth = 0;
    const int xring = 0;                             // dot_x is prefetched into registers (no shared-memory ring)
    (void)with_dot;
    s.NS = ns;
    s.smem_bytes = ns * stage + s.wbytes + misc + xring;
    return true;
}

// Output slices per work item: few enough items per SM that the last wave is full, many enough slices that the DR + 2
// input slices (4 of them issued with a narrower N) amortise.
static int ds_pick_dr(int N, int D, int H, int W, int sms) {
    const long long cols = (long long)N * ((H + DS_TH - 1) / DS_TH) * ((W + DS_TW - 1) / DS_TW);
    int best = 1;
    double best_cost = 1e30;
    for (int dr = 1; dr <= (D < 64 ? D : 64); ++dr) {
        const long long items = cols * ((D + dr - 1) / dr);
        const long long waves = (items + sms - 1) / sms;
        const double cost = (double)waves * (dr + 1.8);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = dr; }
    }
    return best;
}

// layout parameter of the packed operand (pack.cu: batched packing of every weight of a model in one launch)
bool ds_pack_layout(int Cin, int Cout, int kd, int kh, int kw, int* CC) {
    DsShape s;
    if (!ds_shape(Cin, Cout, kd, kh, kw, s)) return false;
    *CC = s.CC;
    return true;
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_conv3d_umma_ds_supported(int Cin, int Cout, int kd, int kh, int kw) {
    DsShape s;
    return (ds_shape(Cin, Cout, kd, kh, kw, s, false) && ds_shape(Cin, Cout, kd, kh, kw, s, true)) ? 1 : 0;
}

int b200em_conv3d_umma_ds_pack(const float* w, int Cout, int Cin, int kd, int kh, int kw, int dgrad, void* packed, void* stream) {
    B2_CHECK_ARG(w && packed && Cout > 0 && Cin > 0, "conv3d_umma_ds_pack: bad arguments");
    DsShape s;
    const int n_ = dgrad ? Cin : Cout, k_ = dgrad ? Cout : Cin;
    if (!ds_shape(k_, n_, kd, kh, kw, s)) {
        set_error("conv3d_umma_ds_pack: shape (%d -> %d, %dx%dx%d) not supported by the depth-stacked tcgen05 path", k_, n_, kd, kh, kw);
        return 2;
    }
    int64_t total = (int64_t)Cout * Cin * kd * kh * kw;
    int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_ds_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, kh, kw, dgrad, s.CC, (__nv_bfloat16*)packed);
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// TO = __nv_bfloat16: bf16 activations in and out.  TO = float: the h16 path -- fp16 operand copies in, fp32 out (no fused norm
// apply: b200em_cvt_f16 did it; no dot_x: the fp32 data gradients take the plain kernel).
template <typename TO>
static int launch_conv_ds(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                          void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W,
                          int Cin, int Cout, int kd, int kh, int kw, int relu, void* stream) {
    constexpr bool H16 = sizeof(TO) == 4;
    B2_CHECK_ARG(!H16 || (!in_scale_shift && !dot_x), "conv3d_umma_ds_h16: takes neither a fused norm apply nor dot_x");
    B2_CHECK_ARG(x && w_packed && y && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_umma_ds: bad arguments");
    DsShape s;
    if (!ds_shape(Cin, Cout, kd, kh, kw, s, dot_x != nullptr)) {
        set_error("conv3d_umma_ds: shape (%d -> %d, %dx%dx%d) not supported by the depth-stacked tcgen05 path", Cin, Cout, kd, kh, kw);
        return 2;
    }
    B2_CHECK_ARG(x_ld % 8 == 0 && y_ld % (H16 ? 4 : 8) == 0 && aligned16(x) && aligned16(y) && aligned16(w_packed),
                 "conv3d_umma_ds: activations must be 16-byte aligned with pitch % 8 == 0");
    B2_CHECK_ARG(x_ld >= Cin && y_ld >= Cout, "conv3d_umma_ds: pitch smaller than channel count");
    B2_CHECK_ARG((long long)H * W * x_ld < (1LL << 31), "conv3d_umma_ds: slice too large for 32-bit in-slice offsets");
    B2_CHECK_ARG(!dot_x || (sums && dot_ld % 8 == 0 && aligned16(dot_x) && dot_ld >= Cout), "conv3d_umma_ds: dot_x needs sums, 16-byte alignment and pitch >= Cout");
    ConvDsParams p;
    p.x = (const __nv_bfloat16*)x; p.x_ld = x_ld; p.in_ss = in_scale_shift; p.w = (const __nv_bfloat16*)w_packed; p.bias = bias;
    p.y = y; p.y_ld = y_ld; p.fp16 = H16 ? 1 : 0; p.sums = sums; p.dot_x = (const __nv_bfloat16*)dot_x; p.dot_ld = dot_ld;
    p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kh = kh; p.kw = kw; p.relu = relu;
    p.CC = s.CC; p.nchunks = Cin / s.CC; p.NS = s.NS; p.wbytes = s.wbytes; p.xdepth = s.xdepth;
    { const int nlw_max = (Cout == 64 || Cout == 32) ? 6 : 8; p.nlw = s.NS < nlw_max ? s.NS : nlw_max; }
    { const char* e = getenv("B200EM_DEBUG"); p.debug = e ? atoi(e) : 0; }
    {
        const char* e = getenv("B200EM_DS_DR");
        p.DR = e ? atoi(e) : ds_pick_dr(N, D, H, W, sm_count());
        if (p.DR < 1) p.DR = 1;
    }
    p.tiles_w = (W + DS_TW - 1) / DS_TW; p.tiles_h = (H + DS_TH - 1) / DS_TH; p.tiles_d = (D + p.DR - 1) / p.DR;
    p.items = (long long)N * p.tiles_d * p.tiles_h * p.tiles_w;
    B2_CHECK_ARG(p.items < (1LL << 31), "conv: too many work items for 32-bit indexing");
    long long gx = p.items < sm_count() ? p.items : sm_count();
    // tiled tensor map over x: dims (C, W, H, D, N), box (8, 10, 18, 1, 1) -> one TMA load per 8-channel plane of a stage
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    p.use_tma = 0;
    {
        const char* e = getenv("B200EM_DS_TMA");
        if (!(e && atoi(e) == 0)) {
            typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            static EncodeFn encode = nullptr;
            static bool looked = false;
            if (!looked) {
                looked = true;
                void* fn = nullptr;
                cudaDriverEntryPointQueryResult qres;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
                    encode = (EncodeFn)fn;
                else
                    (void)cudaGetLastError();
            }
            if (encode) {
                const cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
                const cuuint64_t gstr[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2,
                                            (cuuint64_t)D * H * W * x_ld * 2};
                const cuuint32_t box[5] = {8, (cuuint32_t)DS_WP, (cuuint32_t)DS_HP, 1, 1};
                const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
                const CUresult r = encode(&tmap, H16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r == CUDA_SUCCESS) p.use_tma = 1;
            }
        }
    }
#define B2_DS_LAUNCH(CO_)                                                                                                          \
    case CO_:                                                                                                                      \
        if (s.CC == 32) {                                                                                                          \
            B2_CUDA(cudaFuncSetAttribute(conv3d_umma_ds_kernel<CO_, 2, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_MAX_SMEM)); \
            conv3d_umma_ds_kernel<CO_, 2, TO><<<(unsigned)gx, DsRoles<CO_>::THREADS, s.smem_bytes, (cudaStream_t)stream>>>(p, tmap);                    \
        } else {                                                                                                                   \
            B2_CUDA(cudaFuncSetAttribute(conv3d_umma_ds_kernel<CO_, 1, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_MAX_SMEM)); \
            conv3d_umma_ds_kernel<CO_, 1, TO><<<(unsigned)gx, DsRoles<CO_>::THREADS, s.smem_bytes, (cudaStream_t)stream>>>(p, tmap);                    \
        }                                                                                                                          \
        break;
    switch (Cout) {
        B2_DS_LAUNCH(16)
        B2_DS_LAUNCH(32)
        B2_DS_LAUNCH(48)
        B2_DS_LAUNCH(64)
        B2_DS_LAUNCH(80)
        default:
            set_error("conv3d_umma_ds: Cout %d not instantiated", Cout);
            return 2;
    }
#undef B2_DS_LAUNCH
    B2_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int b200em_conv3d_umma_ds(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                          void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W,
                          int Cin, int Cout, int kd, int kh, int kw, int relu, void* stream) {
    return launch_conv_ds<__nv_bfloat16>(x, x_ld, in_scale_shift, w_packed, bias, y, y_ld, sums, dot_x, dot_ld, N, D, H, W, Cin, Cout, kd, kh, kw,
                                         relu, stream);
}

int b200em_conv3d_umma_ds_h16(const void* x_f16, int64_t x_ld, const void* w_packed, const float* bias, float* y, int64_t y_ld, float* sums,
                              int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int relu, void* stream) {
    return launch_conv_ds<float>(x_f16, x_ld, nullptr, w_packed, bias, y, y_ld, sums, nullptr, 0, N, D, H, W, Cin, Cout, kd, kh, kw, relu, stream);
}

int b200em_conv3d_first_supported(int Cin, int Cout, int kd, int kh, int kw) { return first_shape(Cin, Cout, kd, kh, kw) ? 1 : 0; }

static int first_setup(ConvFirstParams& p, int N, int D, int H, int W, int Cout, int splits) {
    p.N = N; p.D = D; p.H = H; p.W = W; p.Cout = Cout;
    const char* e = getenv("B200EM_DS_DR");
    p.DR = e ? atoi(e) : ds_pick_dr(N, D, H, W, splits);
    if (p.DR < 1) p.DR = 1;
    p.tiles_w = (W + DS_TW - 1) / DS_TW; p.tiles_h = (H + DS_TH - 1) / DS_TH; p.tiles_d = (D + p.DR - 1) / p.DR;
    p.items = (long long)N * p.tiles_d * p.tiles_h * p.tiles_w;
    return 0;
}

}  // extern "C"

template <bool F32>
static int launch_conv_first(const void* x, const float* in_scale_shift, const float* w, const float* bias, void* y, int64_t y_ld,
                             float* sums, int N, int D, int H, int W, int Cout, int relu, void* stream) {
    B2_CHECK_ARG(x && w && y && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_first: bad arguments");
    B2_CHECK_ARG(first_shape(1, Cout, 3, 3, 3), "conv3d_first: Cout %d not supported", Cout);
    B2_CHECK_ARG(y_ld % (F32 ? 4 : 8) == 0 && aligned16(y) && y_ld >= Cout, "conv3d_first: output must be 16-byte aligned with a pitch that keeps it so");
    ConvFirstParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.in_ss = in_scale_shift; p.w = w; p.bias = bias; p.y = y; p.y_ld = y_ld;
    p.sums = sums; p.relu = relu;
    first_setup(p, N, D, H, W, Cout, sm_count());
    B2_CHECK_ARG(p.items < (1LL << 31), "conv: too many work items for 32-bit indexing");
    const long long gx = p.items < sm_count() ? p.items : sm_count();
    const int smem_bytes = F1_NS * F1_ASTAGE + 4 * Cout * 16 + ((F1_NLW * F1_SCR * 2 + 15) & ~15) + Cout * 12 + (2 * F1_NS + 2 * DS_MAX_NA) * 8 + 16 + 128;
#define B2_F1(CO_)                                                                                                     \
    case CO_:                                                                                                          \
        B2_CUDA(cudaFuncSetAttribute(conv3d_first_kernel<CO_, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_MAX_SMEM)); \
        conv3d_first_kernel<CO_, F32><<<(unsigned)gx, F1_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);                \
        break;
    static const bool eg1 = [] { const char* e = getenv("B200EM_FIRST_EG"); return e && atoi(e) == 1; }();
    if (Cout == 32 && !F32 && !eg1) {         // the bf16 nets of width 32: two epilogue warp groups of 16 columns each (369 -> 300 us on (4, 128^3))
        B2_CUDA(cudaFuncSetAttribute(conv3d_first_kernel<32, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_MAX_SMEM));
        conv3d_first_kernel<32, false, 2><<<(unsigned)gx, 2 * 128 + F1_NLW * 32 + 32, smem_bytes, (cudaStream_t)stream>>>(p);
        B2_LAUNCH_CHECK();
        return 0;
    }
    switch (Cout) { B2_F1(16) B2_F1(32) B2_F1(48) B2_F1(64) default: set_error("conv3d_first: Cout %d not instantiated", Cout); return 2; }
#undef B2_F1
    B2_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int b200em_conv3d_first(const void* x, const float* in_scale_shift, const float* w, const float* bias, void* y, int64_t y_ld,
                        float* sums, int N, int D, int H, int W, int Cout, int relu, void* stream) {
    return launch_conv_first<false>(x, in_scale_shift, w, bias, y, y_ld, sums, N, D, H, W, Cout, relu, stream);
}

int b200em_conv3d_first_f32(const float* x, const float* in_scale_shift, const float* w, const float* bias, float* y, int64_t y_ld,
                            float* sums, int N, int D, int H, int W, int Cout, int relu, void* stream) {
    return launch_conv_first<true>(x, in_scale_shift, w, bias, y, y_ld, sums, N, D, H, W, Cout, relu, stream);
}

int b200em_conv3d_first_wgrad(const void* x, const float* in_scale_shift, const void* dz, int64_t dz_ld, float* dw, float* db, int N,
                              int D, int H, int W, int Cout, void* stream) {
    B2_CHECK_ARG(x && dz && dw && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_first_wgrad: bad arguments");
    B2_CHECK_ARG(first_shape(1, Cout, 3, 3, 3), "conv3d_first_wgrad: Cout %d not supported", Cout);
    B2_CHECK_ARG(dz_ld % 8 == 0 && aligned16(dz) && dz_ld >= Cout, "conv3d_first_wgrad: dz must be 16-byte aligned with pitch % 8 == 0");
    ConvFirstParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.in_ss = in_scale_shift; p.dz = (const __nv_bfloat16*)dz; p.dz_ld = dz_ld; p.dw = dw; p.db = db;
    first_setup(p, N, D, H, W, Cout, sm_count());
    B2_CHECK_ARG(p.items < (1LL << 31), "conv: too many work items for 32-bit indexing");
    const long long gx = p.items < sm_count() ? p.items : sm_count();
    const int zstage = (Cout / 8) * 128 * 16;
    const int smem_bytes = F1_NS * F1_ASTAGE + F1_NS * zstage + ((F1_NLW * F1_SCR * 2 + 15) & ~15) + Cout * 4 + (2 * F1_NS + 1) * 8 + 16 + 128;
    B2_CHECK_ARG(F1_NS * zstage >= 24 * 1024 && smem_bytes <= DS_MAX_SMEM, "conv3d_first_wgrad: shared memory layout");
#define B2_F1W(CO_)                                                                                                          \
    case CO_:                                                                                                                \
        B2_CUDA(cudaFuncSetAttribute(conv3d_first_wgrad_kernel<CO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_MAX_SMEM)); \
        conv3d_first_wgrad_kernel<CO_><<<(unsigned)gx, F1_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);                     \
        break;
    switch (Cout) { B2_F1W(16) B2_F1W(32) B2_F1W(48) B2_F1W(64) default: set_error("conv3d_first_wgrad: Cout %d not instantiated", Cout); return 2; }
#undef B2_F1W
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
