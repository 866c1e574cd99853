// Implicit-GEMM 3-D convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), bf16
// operands, fp32 accumulation.  Forward ([Norm ->] Conv3d -> bias -> ReLU, unet.py:429-438) and, with flipped /
// transposed weights, the data gradient.  im2col-free: the A operand of every filter tap is the SAME haloed NDHWC
// tile in shared memory, addressed through the shared-memory matrix descriptor (start address + tap offset).
//
// Work item (one per CTA iteration, persistent grid = #SMs):  R consecutive depth slabs of a 16 (h) x 8 (w) voxel tile
//   -> R accumulators of 128 voxels x NP output channels in TMEM (double buffered when 2*R*NP <= 512 columns).
// K loop:  for each 32(16)-channel chunk of Cin:  for each filter tap:  for each slab r:  K/16 MMAs of 128 x NP x 16.
//
// Shared-memory operand layout (SWIZZLE_NONE K-major canonical form, 8-row x 16-byte core matrices):
//   A: [slice s = 0..R+kd-2][plane j = 8-channel group][hp = 0..17][wp = 0..9][8 bf16]   (PLANE bytes per plane)
//      GEMM row m = 8*hl + wl of slab r, tap (a,b,c) lives at slice r+a, row hp = hl+b(+1-ph), column wp = wl+c(+1-pw):
//      rows of one 8-group are 16 B apart, 8-groups are 10*16 B apart (SBO), the two K-chunks PLANE apart (LBO).
//   B: [tap in stage][plane j][n = 0..NP-1][8 bf16], filled by cp.async.bulk (TMA engine) from the pre-packed weights.
//
// Warp roles (448 threads): warps 0-3 epilogue (TMEM -> bias/ReLU -> bf16 store + per-(n,c) statistics for the next
// norm), warps 4-11 operand loaders (global -> [scale*x+shift of the preceding norm] -> shared; padding is written as
// zeros, i.e. in normalised space like the reference), warp 12 weight loader (one elected thread), warp 13 MMA issuer
// (one elected thread).  All hand-offs are mbarriers; the accumulator hand-off is tcgen05.commit.
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "umma.cuh"
#include "halo_tile.cuh"

namespace b200em {

using namespace umma;

namespace {
constexpr int TH = 16, TW = 8;             // voxel tile (h, w) = 128 GEMM rows
constexpr int HP = TH + 2, WP = TW + 2;     // haloed tile
constexpr int PLANE = HP * WP * 16 + 16;    // bytes per (slice, 8-channel group) plane; +16 staggers banks (cp.async loaders)
constexpr int PLANE_TMA = 2944;             // the same plane padded to a multiple of 128 B: destination of one tiled TMA load
constexpr int NSTAGE = 4;                   // maximum weight ring depth (the launch picks p.nstage <= NSTAGE)
constexpr int NLOAD = 256;                  // operand-loader threads (8 warps): the tile load is latency-bound, more threads = more bytes in flight
constexpr int THREADS = 128 + NLOAD + 64;   // 4 epilogue warps | 8 loader warps | weight-loader warp | MMA warp
constexpr int W_WLOAD = (128 + NLOAD) / 32, W_MMA = W_WLOAD + 1;
constexpr int MAX_SMEM = 227 * 1024;
}  // namespace


// Activations / packed weights are bf16 (kind::f16 MMAs, K = 16) or fp32 (kind::tf32 MMAs, K = 8: the tensor core reads the upper
// 19 bits of every fp32 word -- what torch / cuDNN do for fp32 convolutions by default, torch.backends.cudnn.allow_tf32).  Either
// way a 16-byte shared-memory unit holds EPU = 16 / sizeof(T) channels of one voxel and an MMA K step consumes two units.
struct ConvUmmaParams {
    const void* x; long long x_ld;
    const float* in_ss;
    const void* w;
    const float* bias;
    void* y; long long y_ld;
    float* sums;
    const void* dot_x; long long dot_ld;   // non-null: sums = (sum y, sum y * dot_x) -- the norm-backward reductions
    const float* x_absmax;                 // h16 path: device max |x| the fp16 operand was scaled by (common.cuh h16_shift), or null
    int prefetch;                          // bring-up switch (env B200EM_PREFETCH=0): no L1 prefetch of the dot_x rows
    int use_tma;                           // 1 (the TMA_ instantiation): the haloed tile planes are loaded by cp.async.bulk.tensor
    int ksplit;                            // > 1: gridDim.z CTAs share an output tile, each reducing a range of the Cin chunks (split-K)
    float* ws_acc; unsigned* ws_cnt;       // split-K: zeroed fp32 partial sums [voxel][Cout] and per-(item, nblk) arrival counters
    int N, D, H, W, Cin, Cout;
    int kd, kh, kw, relu;
    int R, NP, CC, nchunks, G, acc_bufs, nstage;
    int tiles_w, tiles_h, tiles_d;
    long long items;
    int a_bytes, b_stage_bytes;
};

// 32-bit arithmetic on purpose: 64-bit division is a ~100-instruction software routine and this runs per work item in
// every warp role.
__device__ __forceinline__ void item_coords(const ConvUmmaParams& p, long long item_, int& n, int& d0, int& h0, int& w0) {
    unsigned item = (unsigned)item_;
    const unsigned tw = item % (unsigned)p.tiles_w; item /= (unsigned)p.tiles_w;
    const unsigned th = item % (unsigned)p.tiles_h; item /= (unsigned)p.tiles_h;
    const unsigned td = item % (unsigned)p.tiles_d; item /= (unsigned)p.tiles_d;
    n = (int)item; d0 = (int)td * p.R; h0 = (int)th * TH; w0 = (int)tw * TW;
}

// NV consecutive channels of one voxel: registers (fp32) <-> global memory in the activation type, 16-byte accesses.
template <typename T, int NV>
__device__ __forceinline__ void store_row(T* __restrict__ dst, const float (&v)[NV]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < NV / 8; ++q) {
            uint4 o;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
            *reinterpret_cast<uint4*>(dst + 8 * q) = o;
        }
    }
}
template <typename T, int NV>
__device__ __forceinline__ void load_row(const T* __restrict__ src, bool valid, float (&v)[NV]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) {
            const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(src + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NV / 8; ++q) {
            const uint4 xv = valid ? __ldg(reinterpret_cast<const uint4*>(src + 8 * q)) : make_uint4(0, 0, 0, 0);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&xv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h2[e]);
                v[8 * q + 2 * e] = f.x;
                v[8 * q + 2 * e + 1] = f.y;
            }
        }
    }
}

// R_ = depth slabs per work item, KC_ = K=16 steps per channel chunk: compile-time so that the single-thread MMA issue
// loop is straight-line code with immediate operand offsets (it bounds the small-N layers otherwise).
// TA = operand type in shared memory (activations and packed weights), TO = type of y and dot_x in global memory:
// (bf16, bf16), (float, float) = TF32, (__half, float) = the h16 path (fp32 tensors, fp16 operand copies).
// (448 threads = 14 warps put 4 warps on two of the SM's four sub-partitions, each with 16 K registers: 128 registers per thread
// is the hardware limit for this block size, whatever __maxnreg__ says -- a launch with 144 fails.)
// TMA_: the operand tile arrives by tiled TMA loads (one per slice and 16-byte channel plane, halo zero-filled by the TMA unit)
// into planes of PLANE_TMA bytes; the loader warps only run the in-place norm apply.  A cp.async (LDGSTS) tile load writes shared
// memory sector by sector as the data returns -- measured 38 shared-memory wavefronts per warp instruction, ~15 % of the
// shared-memory pipe the tensor core's operand fetch saturates anyway; the TMA unit writes whole lines.
template <typename TA, typename TO, int R_, int KC_, bool TMA_>
__global__ void __launch_bounds__(THREADS, 1) conv3d_umma_kernel(const ConvUmmaParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int EPU = 16 / (int)sizeof(TA);       // channels per 16-byte unit: 8 (bf16) or 4 (fp32 / TF32)
    constexpr int PL = TMA_ ? PLANE_TMA : PLANE;    // bytes per (slice, 16-byte channel plane)
    // carve: A[2] | B[NSTAGE] | bias[NP] | sums[2*NP] | barriers | tmem ptr
    uint8_t* smA = smem;
    uint8_t* smB = smA + 2 * p.a_bytes;
    float* s_bias = reinterpret_cast<float*>(smB + p.nstage * p.b_stage_bytes);
    float* s_sums = s_bias + p.NP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_sums + 2 * p.NP);
    uint64_t* a_full = bars;             // [2]   128 loader arrivals
    uint64_t* a_empty = bars + 2;        // [2]   tcgen05.commit
    uint64_t* b_full = bars + 4;         // [NSTAGE] expect_tx
    uint64_t* b_empty = bars + 4 + NSTAGE;
    uint64_t* acc_full = bars + 4 + 2 * NSTAGE;   // [2] tcgen05.commit
    uint64_t* acc_empty = acc_full + 2;           // [2] 128 epilogue arrivals
    uint64_t* t_full = acc_empty + 2;             // [2] TMA tile loads (expect_tx)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_full + 2);
    uint32_t* s_tap = s_tmem + 2;                   // [27] operand start offset of each tap, in 16-byte units

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = blockIdx.y;                    // NP-wide output-channel block
    // split-K: this CTA reduces the Cin chunks [c_begin, c_end) of its tiles
    const int c_begin = p.ksplit > 1 ? (int)((blockIdx.z * p.nchunks) / p.ksplit) : 0;
    const int c_end = p.ksplit > 1 ? (int)(((blockIdx.z + 1) * p.nchunks) / p.ksplit) : p.nchunks;
    constexpr int J = KC_ * 2;                      // 8-channel planes per slice per chunk
    const int taps = p.kd * p.kh * p.kw;
    const int ngroups = taps / p.G;
    const int nslices = p.R + p.kd - 1;
    const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;
    const uint32_t acc_cols = (uint32_t)(p.R * p.NP);
    uint32_t tmem_cols = 32;
    while (tmem_cols < acc_cols * p.acc_bufs) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], NLOAD / 32); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); mbar_init(&t_full[i], 1); }
        fence_mbar_init();
    }
    if (warp == W_MMA) tmem_alloc(s_tmem, tmem_cols);
    if (threadIdx.x < taps) {
        const int tap = threadIdx.x;
        const int a = tap / (p.kh * p.kw), b = (tap / p.kw) % p.kh, cc = tap % p.kw;
        s_tap[tap] = (uint32_t)((a * J * PL + ((b + 1 - ph) * WP + (cc + 1 - pw)) * 16) >> 4);
    }
    for (int i = threadIdx.x; i < p.NP; i += THREADS) {
        const int co = nblk * p.NP + i;
        s_bias[i] = (p.bias && co < p.Cout) ? p.bias[co] : 0.f;
        s_sums[2 * i] = 0.f;
        s_sums[2 * i + 1] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp >= 4 && warp < W_WLOAD) {
        // ===================== operand loaders =====================
        const int t = threadIdx.x - 128;             // 0..NLOAD-1
        uint32_t fill = 0;
        if constexpr (TMA_) {
            // One elected thread issues the tile: a tiled TMA load per (slice, plane), box (EPU ch, 10 w, 18 h, 1 d, 1 n) -> [hp][wp][16 B];
            // coordinates outside the volume (the halo, slices before 0 / beyond D) are zero-filled by the TMA unit.  Every loader
            // warp then owns the (slice, plane) pairs w8, w8 + 8, ... (8 % J == 0: its plane j, hence its scale / shift registers,
            // is fixed) for the in-place norm apply: a lane takes voxels lane, lane + 32, ... of a plane, so the eight lanes of a
            // 16-byte shared-memory phase touch one 128-byte line.
            const int w8 = t >> 5;
            const int j = w8 % J;
            const int npairs = nslices * J;
            const uint32_t tile_bytes = (uint32_t)(npairs * HP * WP * 16);
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
                int n, d0, h0, w0;
                item_coords(p, item, n, d0, h0, w0);
                for (int c = c_begin; c < c_end; ++c, ++fill) {
                    const int buf = fill & 1;
                    uint8_t* abuf = smA + buf * p.a_bytes;
                    if (w8 == 0) mbar_wait(&a_empty[buf], ((fill >> 1) & 1) ^ 1);      // the whole warp waits: no divergent spinning
                    if (t == 0) {
                        mbar_arrive_expect_tx(&t_full[buf], tile_bytes);
                        const uint32_t bar32 = smem_u32(&t_full[buf]);
                        for (int s = 0; s < nslices; ++s)
#pragma unroll
                            for (int jj = 0; jj < J; ++jj)
                                asm volatile(
                                    "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
                                        "r"(smem_u32(abuf + (s * J + jj) * PL)),
                                    "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c * p.CC + jj * EPU), "r"(w0 - 1), "r"(h0 - 1), "r"(d0 + s - pd), "r"(n), "r"(bar32)
                                    : "memory");
                    }
                    __syncwarp();
                    mbar_wait(&t_full[buf], (fill >> 1) & 1);
                    if (p.in_ss) {
                        float sc[EPU], sh[EPU];
                        const float* q = p.in_ss + ((size_t)n * p.Cin + c * p.CC + j * EPU) * 2;
#pragma unroll
                        for (int e = 0; e < EPU; ++e) { sc[e] = q[2 * e]; sh[e] = q[2 * e + 1]; }
                        for (int pr = w8; pr < npairs; pr += NLOAD / 32) {
                            const int gd = d0 + pr / J - pd;
                            if (gd < 0 || gd >= p.D) continue;
                            uint8_t* pl = abuf + pr * PL;
#pragma unroll
                            for (int i = 0; i < (HP * WP + 31) / 32; ++i) {
                                const int v = lane + 32 * i;
                                const int hp_ = v / WP, wp_ = v % WP;
                                const int gh = h0 + hp_ - 1, gw = w0 + wp_ - 1;
                                if (v < HP * WP && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W) {
                                    uint4* qd = reinterpret_cast<uint4*>(pl + v * 16);
                                    uint4 val = *qd;
                                    affine_unit<TA>(val, sc, sh);
                                    *qd = val;
                                }
                            }
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[buf]);
                }
            }
        } else {
        const int j = t % J;                         // fixed 8-channel group of this thread (128 % J == 0)
        const int units = nslices * HP * WP;         // voxels of the haloed tile per chunk
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, d0, h0, w0;
            item_coords(p, item, n, d0, h0, w0);
            for (int c = c_begin; c < c_end; ++c, ++fill) {
                const int buf = fill & 1;
                mbar_wait(&a_empty[buf], ((fill >> 1) & 1) ^ 1);
                const int ch0 = c * p.CC + j * EPU;
                float sc[EPU], sh[EPU];
                if (p.in_ss) {
                    const float* q = p.in_ss + ((size_t)n * p.Cin + ch0) * 2;
#pragma unroll
                    for (int e = 0; e < EPU; ++e) { sc[e] = q[2 * e]; sh[e] = q[2 * e + 1]; }
                }
                uint8_t* dstbase = smA + buf * p.a_bytes + j * PL;
                const TA* xn = reinterpret_cast<const TA*>(p.x) + (size_t)n * p.D * p.H * p.W * p.x_ld + ch0;
                load_halo_tile_async<HP, WP, TA>(xn, p.x_ld, sc, sh, p.in_ss != nullptr, dstbase, J * PL, t / J, NLOAD / J, units, d0, h0, w0, pd,
                               p.D, p.H, p.W);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[buf]);
            }
        }
        }
    } else if (warp == W_WLOAD) {
        // ===================== weight loader =====================
        if (elect_one()) {
            const uint32_t bytes = (uint32_t)p.b_stage_bytes;
            const size_t tap_elems = (size_t)J * p.NP * 16;          // BYTES per (chunk, tap) block of the packed operand
            const uint8_t* wblk = reinterpret_cast<const uint8_t*>(p.w) + (size_t)nblk * p.nchunks * taps * tap_elems;
            int st = 0;
            uint32_t ph = 0;                              // ring position and lap parity
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x)
                for (int c = c_begin; c < c_end; ++c)
                    for (int g = 0; g < ngroups; ++g) {
                        mbar_wait(&b_empty[st], ph ^ 1);
                        mbar_arrive_expect_tx(&b_full[st], bytes);
                        bulk_g2s(smB + st * p.b_stage_bytes, wblk + ((size_t)c * taps + (size_t)g * p.G) * tap_elems, bytes, &b_full[st]);
                        if (++st == p.nstage) { st = 0; ph ^= 1; }
                    }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        // One thread issues every MMA, so the per-MMA instruction count is what bounds small-N layers: descriptors are
        // built once, only their 14-bit start-address field changes, and with R_/KC_ known at compile time the per-tap
        // body is R_*KC_ MMAs whose operand offsets are immediates.
        if (elect_one()) {
            const uint32_t idesc = make_idesc<TA>(128, p.NP);
            const uint64_t ad = make_desc(0, PL, WP * 16), bd = make_desc(0, (uint32_t)(p.NP * 16), 128);
            const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
            const uint32_t a_lo_base = (uint32_t)(ad & 0xFFFFFFFFu) + (smem_u32(smA) >> 4);
            const uint32_t b_lo_base = (uint32_t)(bd & 0xFFFFFFFFu) + (smem_u32(smB) >> 4);
            constexpr uint32_t SLAB16 = J * (PL / 16), K16 = 2 * (PL / 16);
            const uint32_t np = (uint32_t)p.NP, b_tap16 = (uint32_t)(J * p.NP), bk16 = 2 * np;
            const uint32_t a_bytes16 = (uint32_t)(p.a_bytes >> 4), bstage16 = (uint32_t)(p.b_stage_bytes >> 4);
            uint32_t fill = 0, it = 0;
            int st = 0;
            uint32_t ph = 0;                             // weight ring position and lap parity
            bool a_ok = false, b_ok = false;             // results of the early probes of the next a_full / b_full barriers
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
                const int slot = (p.acc_bufs == 2) ? (it & 1) : 0;
                const uint32_t use = (p.acc_bufs == 2) ? (it >> 1) : it;
                mbar_wait(&acc_empty[slot], (use & 1) ^ 1);
                tc_fence_after();
                uint32_t tcol[R_];
#pragma unroll
                for (int r = 0; r < R_; ++r) tcol[r] = tmem_base + slot * acc_cols + r * np;
                for (int c = c_begin; c < c_end; ++c, ++fill) {
                    const int buf = fill & 1;
                    if (!a_ok) mbar_wait(&a_full[buf], (fill >> 1) & 1);
                    tc_fence_after();
                    a_ok = mbar_test_wait(&a_full[(fill + 1) & 1], ((fill + 1) >> 1) & 1);
                    const uint32_t abuf = a_lo_base + buf * a_bytes16;
                    for (int g = 0; g < ngroups; ++g) {
                        if (!b_ok) mbar_wait(&b_full[st], ph);
                        tc_fence_after();
                        const int st_n = st + 1 == p.nstage ? 0 : st + 1;
                        const uint32_t ph_n = st_n == 0 ? ph ^ 1 : ph;
                        // early probe of the next weight stage: a probe costs ~200 cycles of latency even when the barrier is
                        // complete, so it is issued before this stage's MMAs and consumed after them
                        b_ok = mbar_test_wait(&b_full[st_n], ph_n);
                        const uint32_t bst = b_lo_base + st * bstage16;
                        for (int tg = 0; tg < p.G; ++tg) {
                            const int tap = g * p.G + tg;
                            const uint32_t a0 = abuf + s_tap[tap];
                            const uint32_t b0 = bst + tg * b_tap16;
                            // slabs beyond the volume (d0 + r >= D) read zero-filled slices: computed, never stored
                            if (c == c_begin && tap == 0) {
#pragma unroll
                                for (int r = 0; r < R_; ++r) {
                                    umma_c<TA, false>(tcol[r], a0 + r * SLAB16, a_hi, b0, b_hi, idesc);
                                    if (KC_ == 2) umma_c<TA, true>(tcol[r], a0 + r * SLAB16 + K16, a_hi, b0 + bk16, b_hi, idesc);
                                }
                            } else {
#pragma unroll
                                for (int r = 0; r < R_; ++r) {
                                    umma_c<TA, true>(tcol[r], a0 + r * SLAB16, a_hi, b0, b_hi, idesc);
                                    if (KC_ == 2) umma_c<TA, true>(tcol[r], a0 + r * SLAB16 + K16, a_hi, b0 + bk16, b_hi, idesc);
                                }
                            }
                        }
                        umma_commit(&b_empty[st]);
                        st = st_n; ph = ph_n;
                    }
                    umma_commit(&a_empty[buf]);
                }
                umma_commit(&acc_full[slot]);
            }
        }
    } else {
        // ===================== epilogue (warps 0-3) =====================
        const int row = warp * 32 + lane;            // GEMM row = TMEM lane
        const int hl = row / TW, wl = row % TW;
        const float osc = pow2i(-h16_shift(p.x_absmax));     // undo the power-of-two operand scaling (1 when there is none)
        const bool split = p.ksplit > 1;
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
            int n, d0, h0, w0;
            item_coords(p, item, n, d0, h0, w0);
            const int slot = (p.acc_bufs == 2) ? (it & 1) : 0;
            const uint32_t use = (p.acc_bufs == 2) ? (it >> 1) : it;
            mbar_wait(&acc_full[slot], use & 1);
            tc_fence_after();
            const int gh = h0 + hl, gw = w0 + wl;
            const bool valid_hw = gh < p.H && gw < p.W;
            const int rmax = min(p.R, p.D - d0);
            if (split) {
                // split-K, phase 1: add this CTA's partial accumulators into the fp32 workspace ([voxel][Cout], all zero between
                // launches), release the TMEM slot, and count the arrival; only the LAST CTA of the tile goes on to the epilogue
                // proper, reading the complete sums back from the workspace (and clearing them again).
                for (int r = 0; r < rmax; ++r) {
                    float* wrow = p.ws_acc + ((((size_t)n * p.D + d0 + r) * p.H + gh) * p.W + gw) * p.Cout + nblk * p.NP;
                    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + slot * acc_cols + r * p.NP;
                    for (int cb = 0; cb < p.NP; cb += 16) {
                        uint32_t raw[16];
                        tmem_ld16(taddr + cb, raw);
                        tmem_ld_wait();
                        if (valid_hw) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + cb + 4 * q), "f"(__uint_as_float(raw[4 * q])),
                                             "f"(__uint_as_float(raw[4 * q + 1])), "f"(__uint_as_float(raw[4 * q + 2])), "f"(__uint_as_float(raw[4 * q + 3]))
                                             : "memory");
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[slot]);
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 0) {
                    unsigned* cnt = p.ws_cnt + (size_t)item * gridDim.y + nblk;
                    const unsigned old = atomicAdd(cnt, 1u);
                    const bool last = old == (unsigned)p.ksplit - 1u;
                    if (last) *cnt = 0u;                 // all arrivals are in: leave the counter zero for the next launch
                    s_tmem[1] = last ? 1u : 0u;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const bool last = *reinterpret_cast<volatile uint32_t*>(&s_tmem[1]) != 0u;
                asm volatile("bar.sync 1, 128;" ::: "memory");     // everybody has read the flag before the next item overwrites it
                if (!last) continue;
                __threadfence();
            }
            // norm-backward reductions (data gradient): the dot_x row of a 32-column block is PREFETCHED into L1 one block ahead --
            // loaded only after the accumulators have arrived it would expose one global-memory latency per block and the epilogue,
            // not the MMAs, would bound the layer (fp32 rows are 128 bytes per thread).  A prefetch instruction, not a register
            // prefetch: the epilogue is at the register limit.
            const bool pre = p.dot_x != nullptr && p.sums != nullptr && valid_hw && p.prefetch;
            auto prefetch_x = [&](int r_, int cb_) {
                const size_t vox_ = (((size_t)n * p.D + d0 + r_) * p.H + gh) * p.W + gw;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const TO*>(p.dot_x) + vox_ * p.dot_ld + nblk * p.NP + cb_));
            };
            if (pre && rmax > 0) prefetch_x(0, 0);
            for (int r = 0; r < rmax; ++r) {
                const int gd = d0 + r;
                const size_t vox = (((size_t)n * p.D + gd) * p.H + gh) * p.W + gw;
                TO* yp = reinterpret_cast<TO*>(p.y) + vox * p.y_ld + nblk * p.NP;
                float* wrow = split ? p.ws_acc + vox * p.Cout + nblk * p.NP : nullptr;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + slot * acc_cols + r * p.NP;
                for (int cb = 0; cb < p.NP; cb += 32) {
                    if (p.NP - cb >= 32) {
                        uint32_t raw[32];
                        if (split) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 t4 = valid_hw ? __ldcg(reinterpret_cast<const float4*>(wrow + cb) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                                raw[4 * q] = __float_as_uint(t4.x); raw[4 * q + 1] = __float_as_uint(t4.y);
                                raw[4 * q + 2] = __float_as_uint(t4.z); raw[4 * q + 3] = __float_as_uint(t4.w);
                                if (valid_hw) __stcg(reinterpret_cast<float4*>(wrow + cb) + q, make_float4(0.f, 0.f, 0.f, 0.f));
                            }
                        } else {
                            tmem_ld32(taddr + cb, raw);
                            tmem_ld_wait();
                        }
                        float v[32];
                        if (p.bias) {                    // 16-byte shared-memory loads, none for the data gradient (no bias)
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cb + 4 * q);
                                v[4 * q] = b4.x; v[4 * q + 1] = b4.y; v[4 * q + 2] = b4.z; v[4 * q + 3] = b4.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float f = fmaf(__uint_as_float(raw[i]), osc, v[i]);
                            if (p.relu) f = fmaxf(f, 0.f);
                            v[i] = round_as<TO>(f);
                        }
                        if (valid_hw) store_row<TO, 32>(yp + cb, v);
                        if (p.sums) {
                            float s1[32], s2[32];
                            if (p.dot_x) {
                                if (pre) {               // next block of this item (same slab, or the first block of the next slab)
                                    if (cb + 32 < p.NP) prefetch_x(r, cb + 32);
                                    else if (r + 1 < rmax) prefetch_x(r + 1, 0);
                                }
                                const TO* xq = reinterpret_cast<const TO*>(p.dot_x) + vox * p.dot_ld + nblk * p.NP + cb;
                                load_row<TO, 32>(xq, valid_hw, s2);
#pragma unroll
                                for (int i = 0; i < 32; ++i) { s1[i] = valid_hw ? v[i] : 0.f; s2[i] *= s1[i]; }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i) { s1[i] = valid_hw ? v[i] : 0.f; s2[i] = s1[i] * s1[i]; }
                            }
                            const float a1 = warp_column_sums<32>(s1, lane);
                            const float a2 = warp_column_sums<32>(s2, lane);
                            atomicAdd(&s_sums[2 * (cb + lane)], a1);
                            atomicAdd(&s_sums[2 * (cb + lane) + 1], a2);
                        }
                    } else {   // 16-column tail (NP % 32 == 16)
                        uint32_t raw[16];
                        if (split) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 t4 = valid_hw ? __ldcg(reinterpret_cast<const float4*>(wrow + cb) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                                raw[4 * q] = __float_as_uint(t4.x); raw[4 * q + 1] = __float_as_uint(t4.y);
                                raw[4 * q + 2] = __float_as_uint(t4.z); raw[4 * q + 3] = __float_as_uint(t4.w);
                                if (valid_hw) __stcg(reinterpret_cast<float4*>(wrow + cb) + q, make_float4(0.f, 0.f, 0.f, 0.f));
                            }
                        } else {
                            tmem_ld16(taddr + cb, raw);
                            tmem_ld_wait();
                        }
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float f = fmaf(__uint_as_float(raw[i]), osc, s_bias[cb + i]);
                            if (p.relu) f = fmaxf(f, 0.f);
                            v[i] = round_as<TO>(f);
                        }
                        if (valid_hw) store_row<TO, 16>(yp + cb, v);
                        if (p.sums) {
                            float s1[16], s2[16];
                            if (p.dot_x) {
                                const TO* xq = reinterpret_cast<const TO*>(p.dot_x) + vox * p.dot_ld + nblk * p.NP + cb;
                                load_row<TO, 16>(xq, valid_hw, s2);
#pragma unroll
                                for (int i = 0; i < 16; ++i) { s1[i] = valid_hw ? v[i] : 0.f; s2[i] *= s1[i]; }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) { s1[i] = valid_hw ? v[i] : 0.f; s2[i] = s1[i] * s1[i]; }
                            }
                            const float a1 = warp_column_sums<16>(s1, lane);   // column = lane >> 1 (held twice)
                            const float a2 = warp_column_sums<16>(s2, lane);
                            if ((lane & 1) == 0) {
                                atomicAdd(&s_sums[2 * (cb + (lane >> 1))], a1);
                                atomicAdd(&s_sums[2 * (cb + (lane >> 1)) + 1], a2);
                            }
                        }
                    }
                }
            }
            if (!split) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[slot]);
            }
            if (p.sums) {
                // flush this item's per-channel partial sums (n may change with the next item)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = threadIdx.x; i < 2 * p.NP; i += 128) {
                    const int co = nblk * p.NP + (i >> 1);
                    if (co < p.Cout) atomicAdd(p.sums + ((size_t)n * p.Cout + co) * 2 + (i & 1), s_sums[i]);
                    s_sums[i] = 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: torch (Cout, Cin, taps) fp32 -> bf16 [nblk][chunk][tap][plane j][NPb][8]
// dgrad = 1 packs the transposed, tap-flipped filter: "Cout" of the packed operand is the conv's Cin and vice versa.
// One block per (8 output channels, 8 input channels): the 8 x 8 x taps fp32 sub-filter is read as 8 contiguous runs of
// 8*taps floats, transposed in shared memory, and written as 16-byte units (8 reduction channels each) -- 8 consecutive n
// share a 128-byte line of the packed image.  (One thread per element scatters 2-byte stores: ~4x slower on the wide layers.)
__global__ void __launch_bounds__(256) pack_umma_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int dgrad, int CC, int NPb,
                                                                __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[8][8 * 27];
    const int co0 = blockIdx.x * 8, ci0 = blockIdx.y * 8;
    const int run = 8 * taps;                          // floats per output channel in this tile (contiguous in w)
    for (int i = threadIdx.x; i < 8 * run; i += blockDim.x) {
        const int c = i / run, r = i % run;
        tile[c][r] = w[((size_t)(co0 + c) * Cin + ci0) * taps + r];
    }
    __syncthreads();
    const int Kc = dgrad ? Cout : Cin;                 // reduction channels of the packed operand
    const int nchunks = Kc / CC, J = CC / 8;
    for (int u = threadIdx.x; u < 8 * taps; u += blockDim.x) {
        const int nl = u % 8, tp = u / 8;              // n within the tile, filter tap (torch order)
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            // forward: n = co, k = ci;  dgrad: n = ci, k = co
            const float f = dgrad ? tile[e][nl * taps + tp] : tile[nl][e * taps + tp];
            v[e] = __float2bfloat16_rn(f);
        }
        const int n_ = (dgrad ? ci0 : co0) + nl, k0 = dgrad ? co0 : ci0, t_ = dgrad ? taps - 1 - tp : tp;
        const int nb = n_ / NPb, nn = n_ % NPb, chunk = k0 / CC, j = (k0 % CC) / 8;
        *reinterpret_cast<uint4*>(out + ((((size_t)(nb * nchunks + chunk) * taps + t_) * J + j) * NPb + nn) * 8) = *reinterpret_cast<const uint4*>(v);
    }
}

struct UmmaShape {
    int CC, NP, nblk, R, G, acc_bufs, nstage, a_bytes, b_stage_bytes, smem_bytes;
};

// Weight ring of a (R, NP, CC) configuration: G filter taps per stage and the ring depth.  The single issuing thread pays a fixed
// ~500 cycles per stage hand-off (barrier probe, commit, fence), so a stage should carry well over 500 cycles of MMAs: three taps
// per stage whenever >= 3 stages of them fit beside the two activation buffers, else one tap and four stages.
static void umma_ring(UmmaShape& s, int taps, int kd, int planes_per_slice) {
    const int fixed = s.NP * 4 * 3 + 18 * 8 + 16 + 27 * 4 + 128;
    s.a_bytes = (s.R + kd - 1) * planes_per_slice * PLANE_TMA;     // sized for the TMA planes (the cp.async planes are smaller)
    static const int g_env = [] { const char* e = getenv("B200EM_UMMA_G"); return e ? atoi(e) : 0; }();      // bring-up: force taps per stage
    for (int G = (taps % 3 == 0 && g_env != 1) ? 3 : 1; G >= 1; G -= 2) {
        s.G = G;
        s.b_stage_bytes = G * planes_per_slice * s.NP * 16;
        int ns = (MAX_SMEM - fixed - 2 * s.a_bytes) / s.b_stage_bytes;
        if (ns > NSTAGE) ns = NSTAGE;
        s.nstage = ns;
        s.smem_bytes = 2 * s.a_bytes + ns * s.b_stage_bytes + fixed;
        if (ns >= (G == 3 ? 3 : 2)) return;
    }
    s.smem_bytes = MAX_SMEM + 1;                        // does not fit
}

// Channel counts the tensor-core path takes; everything else goes to the direct kernel.
static bool umma_shape(int Cin, int Cout, int kd, int kh, int kw, UmmaShape& s, bool f32 = false) {
    if (Cin % 16 || Cout % 16 || Cin < 16 || Cout < 16) return false;
    if (Cout > 256 && Cout % 128) return false;
    // channels per chunk: four 16-byte planes per slice either way (32 bf16 channels, or 16 fp32 channels for the TF32 path)
    s.CC = f32 ? 16 : ((Cin % 32 == 0) ? 32 : 16);
    // Output-channel block of one CTA.  N = 128 already issues at the full tensor rate, so wide layers are cut
    // into 128-channel blocks: twice the CTAs of a 256-wide block for the deep levels (8^3 / 16^3 volumes have few voxel
    // tiles), half the weight bytes streamed per CTA, and room for double-buffered accumulators (2 * R * 128 <= 512).
    s.NP = (Cout % 128 == 0) ? 128 : Cout;
    s.nblk = Cout / s.NP;
    const int taps = kd * kh * kw;
    const int J = s.CC / (f32 ? 4 : 8);
    for (int R = 4; R >= 1; R >>= 1) {
        if (R * s.NP > 512) continue;
        s.R = R;
        s.acc_bufs = (2 * R * s.NP <= 512) ? 2 : 1;
        umma_ring(s, taps, kd, J);
        if (s.smem_bytes <= MAX_SMEM) return true;
    }
    return false;
}


// ===============================================================================================================
// Weight gradient on the tensor cores:  dW[co][ci][tap] += sum_{n,vox} dz[n,vox,co] * x_hat[n,vox+tap,ci]
//
// GEMM view per (32-channel chunk of Cin, NB-channel block of Cout):  K = voxels (16 per MMA: two 8-voxel w-rows),
//   A (MN-major) = x_hat: M = 128 rows = 16 consecutive planes of the haloed tile = 4 depth slices x 32 channels
//       -> rows [0,96) are the three depth taps a = 0,1,2 of slab r; rows [96,128) (slice r+3) are computed and ignored
//   B (MN-major) = dz slab: N = NB output channels
//   D[tap9 = (b,c)] in TMEM: 9 accumulators x NB columns, kept resident over ALL work items of the CTA;
//   one epilogue at the end adds them into the fp32 dW (torch layout) with atomics.
// The haloed x_hat tile is the same shared-memory image the forward kernel builds (and the same fused norm apply);
// the (b,c) tap shift is again just a start-address offset of the descriptor.
namespace {
constexpr int WG_DZ_PLANE = TH * TW * 16;      // bytes per 16-byte-unit plane (8 bf16 / 4 fp32 channels) of a dz slab
// depth slabs per work item: 2 with bf16 operands, 1 with fp32 (TF32) operands whose 32-channel chunk is 8 planes per slice.
// NOTE: only the bf16 instantiation is shipped.  With kind::tf32 and both operands MN-major (SWIZZLE_NONE) the accumulators came
// back as exact zeros on B200 (tests/test_gpu_tf32.py history) -- the no-swizzle MN-major canonical layout of 32-bit operands
// is not what this kernel assumes -- so the fp32 weight gradient runs as three bf16 MMAs on split operands instead
// (backend.wgrad: x = hi + lo, dz = hi + lo, hi*hi + hi*lo + lo*hi; ~16 mantissa bits, more than TF32's 10).
template <typename TA> struct WgR { static constexpr int value = sizeof(TA) == 4 ? 1 : 2; };
}  // namespace

struct WgradUmmaParams {
    const void* x; long long x_ld;
    const float* in_ss;
    const void* dz; long long dz_ld;
    float* dw;
    float* db;
    int N, D, H, W, Cin, Cout;
    int kd, kh, kw;
    int NB, nco;                               // Cout block, number of Cout blocks
    int pad_bytes;                             // zeroed shared memory behind the x buffers for the last slab's over-read
    int tiles_w, tiles_h, tiles_d;
    long long items;
    int x_bytes, dz_bytes;
    int debug;                                 // bring-up switches (env B200EM_DEBUG): 1 no operand loads, 4 no MMAs, 8 no epilogue atomics
    int CH;                                    // input channels per CTA: 32 (M rows = 4 depth slices x 32 channels: the depth taps), or for
                                               // 1x1x1 filters 32 / 64 / 128 (M rows = 128 / CH slices x CH channels, rows [0, CH) useful)
    int fp16;                                  // 2-byte operands are IEEE fp16 (the h16 path of fp32 activations), not bf16
    const float* x_absmax; const float* z_absmax;   // h16: device max |.| the operands were scaled by (common.cuh h16_shift), or null
};

template <typename TA>
__global__ void __launch_bounds__(THREADS, 1) conv3d_wgrad_umma_kernel(const WgradUmmaParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    // carve: X[2] | pad slices | DZ[2] | db sums[NB] | barriers | tmem ptr
    constexpr int EPU = 16 / (int)sizeof(TA);                  // channels per 16-byte unit
    const int J = p.CH / EPU;                                  // planes per slice: a 32-channel chunk = 4 (bf16) or 8 (fp32); up to 16 for 1x1x1 filters
    constexpr int WG_R = WgR<TA>::value;
    constexpr int KROWS = sizeof(TA) == 4 ? 1 : 2;             // tile rows (8 voxels each) per MMA K step: K = 8 (tf32) or 16
    const int nslices = WG_R + p.kd - 1;
    uint8_t* smX = smem;
    uint8_t* smZ = smX + 2 * p.x_bytes + p.pad_bytes;          // room for the over-read of the last slab (slices r+1 .. r+3)
    float* s_db = reinterpret_cast<float*>(smZ + 2 * p.dz_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_db + p.NB);
    uint64_t* full = bars;            // [2] 128 loader arrivals
    uint64_t* empty = bars + 2;       // [2] tcgen05.commit
    uint64_t* acc_full = bars + 4;    // [1]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 5);
    uint32_t* s_tap9 = s_tmem + 2;    // [9] (b, c) tap offset within the haloed tile, 16-byte units

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.y / p.nco, cob = blockIdx.y % p.nco;
    const int JO = p.NB / EPU;
    const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;
    const int tap9 = p.kh * p.kw;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(tap9 * p.NB)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&full[i], NLOAD / 32); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == W_MMA) tmem_alloc(s_tmem, tmem_cols);
    if (threadIdx.x < tap9) {
        const int b = threadIdx.x / p.kw, cc = threadIdx.x % p.kw;
        s_tap9[threadIdx.x] = (uint32_t)((b + 1 - ph) * WP + (cc + 1 - pw));
    }
    for (int i = threadIdx.x; i < p.NB; i += THREADS) s_db[i] = 0.f;
    // the over-read region must hold finite-or-not garbage only in rows that are ignored; zero it once anyway
    for (int i = threadIdx.x; i < p.pad_bytes / 16; i += THREADS)
        reinterpret_cast<uint4*>(smX + 2 * p.x_bytes)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    auto coords = [&](long long item_, int& n, int& d0, int& h0, int& w0) {      // 32-bit: see item_coords
        unsigned item = (unsigned)item_;
        const unsigned tw = item % (unsigned)p.tiles_w; item /= (unsigned)p.tiles_w;
        const unsigned th = item % (unsigned)p.tiles_h; item /= (unsigned)p.tiles_h;
        const unsigned td = item % (unsigned)p.tiles_d; item /= (unsigned)p.tiles_d;
        n = (int)item; d0 = (int)td * WG_R; h0 = (int)th * TH; w0 = (int)tw * TW;
    };

    if (warp >= 4 && warp < W_WLOAD) {
        // ===================== loaders: x_hat haloed tile + dz slabs =====================
        const int t = threadIdx.x - 128;
        const int j = t % J, jo = t % JO;
        const int xunits = nslices * HP * WP, zunits = WG_R * TH * TW;
        float dbacc[EPU], sc[EPU], sh[EPU];
#pragma unroll
        for (int e = 0; e < EPU; ++e) { dbacc[e] = 0.f; sc[e] = 1.f; sh[e] = 0.f; }
        int cur_n = -1;
        uint32_t fill = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++fill) {
            int n, d0, h0, w0;
            coords(item, n, d0, h0, w0);
            const int buf = fill & 1;
            mbar_wait(&empty[buf], ((fill >> 1) & 1) ^ 1);
            const int ch0 = chunk * p.CH + j * EPU;
            if (p.in_ss && n != cur_n) {           // scale/shift of this thread's channels: reload only when the sample changes
                const float* q = p.in_ss + ((size_t)n * p.Cin + ch0) * 2;
#pragma unroll
                for (int e = 0; e < EPU; ++e) { sc[e] = q[2 * e]; sh[e] = q[2 * e + 1]; }
                cur_n = n;
            }
            uint8_t* xdst = smX + buf * p.x_bytes + j * PLANE;
            const TA* xn = reinterpret_cast<const TA*>(p.x) + (size_t)n * p.D * p.H * p.W * p.x_ld + ch0;
            uint8_t* zdst = smZ + buf * p.dz_bytes + jo * WG_DZ_PLANE;
            const TA* zn = reinterpret_cast<const TA*>(p.dz) + (size_t)n * p.D * p.H * p.W * p.dz_ld + cob * p.NB + jo * EPU;
            const int zstep = NLOAD / JO;
            if (!(p.debug & 1)) {
                // dz slabs: plain copies, issued first; the wait inside load_halo_tile_async covers them as well
                const uint32_t z32 = smem_u32(zdst);
                for (int v = t / JO; v < zunits; v += zstep) {
                    const int wl = v % TW, hl = (v / TW) % TH, r = v / (TW * TH);
                    const int gd = d0 + r, gh = h0 + hl, gw = w0 + wl;
                    const bool in = gd < p.D && gh < p.H && gw < p.W;
                    const TA* src = in ? zn + (((size_t)gd * p.H + gh) * p.W + gw) * p.dz_ld : zn;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(z32 + (uint32_t)(r * JO * WG_DZ_PLANE + (hl * TW + wl) * 16)),
                                 "l"(src), "r"(in ? 16 : 0)
                                 : "memory");
                }
                load_halo_tile_async<HP, WP, TA>(xn, p.x_ld, sc, sh, p.in_ss != nullptr, xdst, J * PLANE, t / J, NLOAD / J, xunits, d0, h0, w0,
                                                 pd, p.D, p.H, p.W);
                if (p.db && chunk == 0) {
                    for (int v = t / JO; v < zunits; v += zstep) {
                        const int wl = v % TW, hl = (v / TW) % TH, r = v / (TW * TH);
                        const uint4 val = *reinterpret_cast<const uint4*>(zdst + r * JO * WG_DZ_PLANE + (hl * TW + wl) * 16);
                        if constexpr (sizeof(TA) == 4) {
                            const float* f = reinterpret_cast<const float*>(&val);
#pragma unroll
                            for (int e = 0; e < 4; ++e) dbacc[e] += f[e];
                        } else if (p.fp16) {
                            const __half2* h2 = reinterpret_cast<const __half2*>(&val);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 f = __half22float2(h2[e]);
                                dbacc[2 * e] += f.x;
                                dbacc[2 * e + 1] += f.y;
                            }
                        } else {
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&val);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 f = __bfloat1622float2(h2[e]);
                                dbacc[2 * e] += f.x;
                                dbacc[2 * e + 1] += f.y;
                            }
                        }
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[buf]);
        }
        if (p.db && chunk == 0) {
#pragma unroll
            for (int e = 0; e < EPU; ++e) atomicAdd(&s_db[jo * EPU + e], dbacc[e]);
            asm volatile("bar.sync 2, %0;" ::"n"(NLOAD) : "memory");
            if (t < p.NB) atomicAdd(p.db + cob * p.NB + t, s_db[t] * pow2i(-h16_shift(p.z_absmax)));
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = p.fp16 ? make_idesc_f16(128, p.NB, 1, 1) : make_idesc<TA>(128, p.NB, 1, 1);       // both operands MN-major
            // A: LBO = next 8 voxels (next tile row), SBO = next 8 channels (next plane); B likewise on the dz slab
            const uint64_t ad = make_desc(0, WP * 16, PLANE), bd = make_desc(0, TW * 16, WG_DZ_PLANE);
            const uint32_t a_hi = (uint32_t)(ad >> 32), a_lo_c = (uint32_t)(ad & 0xFFFFFFFFu);
            const uint32_t b_hi = (uint32_t)(bd >> 32), b_lo_c = (uint32_t)(bd & 0xFFFFFFFFu);
            const uint32_t x_base16 = smem_u32(smX) >> 4, z_base16 = smem_u32(smZ) >> 4;
            const uint32_t nb = (uint32_t)p.NB;
            uint32_t fill = 0;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++fill) {
                int n, d0, h0, w0;
                coords(item, n, d0, h0, w0);
                const int buf = fill & 1;
                mbar_wait(&full[buf], (fill >> 1) & 1);
                tc_fence_after();
                const int rmax = min(WG_R, p.D - d0);
                for (int r = 0; r < rmax; ++r) {
                    const uint32_t xs = a_lo_c + x_base16 + (uint32_t)((buf * p.x_bytes + r * J * PLANE) >> 4);
                    const uint32_t zs = b_lo_c + z_base16 + (uint32_t)((buf * p.dz_bytes + r * JO * WG_DZ_PLANE) >> 4);
                    const bool first = (fill | (uint32_t)r) == 0;      // very first slab of this CTA: overwrite the accumulators
                    if (p.debug & 4) continue;
                    for (int tp = 0; tp < tap9; ++tp) {
                        const uint32_t a0 = xs + s_tap9[tp];
                        const uint32_t tacc = tmem_base + tp * nb;
                        // K step ks = KROWS voxel rows of the tile (8 voxels each); offsets are immediates
                        if (first) umma_c<TA, false>(tacc, a0, a_hi, zs, b_hi, idesc);
                        else umma_c<TA, true>(tacc, a0, a_hi, zs, b_hi, idesc);
#pragma unroll
                        for (int ks = 1; ks < TH / KROWS; ++ks)
                            umma_c<TA, true>(tacc, a0 + ks * KROWS * WP, a_hi, zs + ks * KROWS * TW, b_hi, idesc);
                    }
                }
                umma_commit(&empty[buf]);
            }
            umma_commit(acc_full);
        }
    } else if (warp < 4) {
        // ===================== epilogue: TMEM -> atomics into dW =====================
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int a = warp;                              // depth tap of this warp's 32 lanes (a == 3: ignored rows)
        const int taps = p.kd * tap9;
        const float osc = pow2i(-h16_shift(p.x_absmax) - h16_shift(p.z_absmax));
        // rows [32 a, 32 a + 32): depth tap a of the CTA's 32 channels -- or, for a 1x1x1 filter, channels [32 a, 32 a + 32) of its CH
        const bool wide = p.CH > 32;
        if ((wide ? 32 * a < p.CH : a < p.kd) && !(p.debug & 8)) {
            const int ci = wide ? chunk * p.CH + 32 * a + lane : chunk * 32 + lane;
            for (int tp = 0; tp < tap9; ++tp) {
                const int tap = wide ? 0 : a * tap9 + tp;
                for (int cb = 0; cb < p.NB; cb += 16) {
                    uint32_t raw[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + tp * p.NB + cb, raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int co = cob * p.NB + cb + i;
                        atomicAdd(p.dw + ((size_t)co * p.Cin + ci) * taps + tap, __uint_as_float(raw[i]) * osc);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

static bool wgrad_umma_shape(int Cin, int Cout, int& NB) {
    if (Cin % 32 || Cout % 16 || Cin < 32 || Cout < 16) return false;
    NB = (Cout % 32 == 0) ? 32 : 16;
    return true;
}


// layout parameters of the packed operand (pack.cu: batched packing of every weight of a model in one launch)
bool umma_pack_layout(int Cin, int Cout, int kd, int kh, int kw, int* CC, int* NP, bool f32) {
    UmmaShape s;
    if (!umma_shape(Cin, Cout, kd, kh, kw, s, f32)) return false;
    *CC = s.CC; *NP = s.NP;
    return true;
}

}  // namespace b200em

using namespace b200em;

template <typename TA, typename TO = TA>
static int launch_conv_umma(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                            void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W, int Cin,
                            int Cout, int kd, int kh, int kw, int relu, void* stream, const float* x_absmax = nullptr) {
    constexpr bool F32 = sizeof(TA) == 4;
    constexpr int EPU = 16 / (int)sizeof(TA);
    constexpr int EPO = 16 / (int)sizeof(TO);
    B2_CHECK_ARG(x && w_packed && y && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_umma: bad arguments");
    B2_CHECK_ARG((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3), "conv3d_umma: kernel dims must be 1 or 3");
    UmmaShape s;
    if (!umma_shape(Cin, Cout, kd, kh, kw, s, F32)) {
        set_error("conv3d_umma: channel counts (%d -> %d) not supported by the tcgen05 path", Cin, Cout);
        return 2;
    }
    B2_CHECK_ARG(x_ld % EPU == 0 && y_ld % EPO == 0 && aligned16(x) && aligned16(y), "conv3d_umma: activations must be 16-byte aligned with a pitch that keeps them so");
    B2_CHECK_ARG(x_ld >= Cin && y_ld >= Cout, "conv3d_umma: pitch smaller than channel count");
    ConvUmmaParams p;
    p.x = x; p.x_ld = x_ld; p.in_ss = in_scale_shift; p.w = w_packed; p.bias = bias;
    p.y = y; p.y_ld = y_ld; p.sums = sums;
    p.dot_x = dot_x; p.dot_ld = dot_ld; p.x_absmax = x_absmax;
    B2_CHECK_ARG(!dot_x || (sums && dot_ld % EPO == 0 && aligned16(dot_x) && dot_ld >= Cout), "conv3d_umma: dot_x needs sums, 16-byte alignment and pitch >= Cout");
    p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = kh; p.kw = kw; p.relu = relu;
    // Depth slabs per work item and split-K factor.  The shape admits R <= s.R; fewer slabs = more work items (the deep levels
    // have fewer voxel tiles than SMs) but the weight block is streamed once per item.  With a registered workspace
    // (b200em_set_workspace) the Cin chunks of a tile can additionally be split over ks CTAs (partial sums through the workspace,
    // the last CTA runs the epilogue).  Cost model per CTA (cycles): items per CTA x (max(MMA issue, weight stream at ~48 B/clk
    // from L2) / ks + a fixed pipeline fill per item [+ the partial-sum round trip]).
    int ksplit = 1;
    float* ws_acc = nullptr;
    unsigned* ws_cnt = nullptr;
    {
        const int taps = kd * kh * kw;
        const int nchunks = Cin / s.CC;
        const long long cols = (long long)N * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
        int64_t ws_floats = 0;
        int ncnt = 0;
        const bool have_ws = get_workspace(&ws_acc, &ws_floats, &ws_cnt, &ncnt) && (int64_t)N * D * H * W * Cout <= ws_floats;
        int ks_env = 0;                                  // bring-up / tests: B200EM_KSPLIT=k forces the factor (1 forbids the split)
        { const char* e = getenv("B200EM_KSPLIT"); if (e) ks_env = atoi(e); }
        const double mma_cyc = s.NP / 2 > 32 + s.NP / 4 ? s.NP / 2 : 32 + s.NP / 4;
        const double wts = (double)taps * Cin * s.NP * sizeof(TA) / 48.0;
        const int ks_only = ks_env > 0 ? (ks_env < nchunks ? ks_env : nchunks) : 0;
        double best = 1e300, best_f = 1e300;             // best overall / best with the forced factor
        int best_r = s.R, best_ks = 1, best_fr = 0;
        for (int R = s.R; R >= 1; R >>= 1) {
            const long long items = cols * ((D + R - 1) / R);
            for (int ks = 1; ks <= nchunks && ks <= 16; ++ks) {
                if (ks > 1 && (!have_ws || items * s.nblk > ncnt || items * s.nblk * ks > sm_count())) break;
                const int cta_cap = sm_count() / (s.nblk * ks) > 0 ? sm_count() / (s.nblk * ks) : 1;
                const long long gx_ = items < cta_cap ? items : cta_cap;
                const long long per_cta = (items + gx_ - 1) / gx_;
                const double mma = (double)R * taps * (Cin / (2 * EPU)) * mma_cyc;
                const double epi = (2 * R * s.NP <= 512) ? 0.0 : (double)R * (s.NP / 32) * 700.0;   // single accumulator set: the epilogue is not overlapped
                const double part = ks > 1 ? (double)R * (s.NP / 32) * 1500.0 + 3000.0 : 0.0;      // partial sums out, arrival counter, sums back in
                const double cost = per_cta * ((mma > wts ? mma : wts) / ks + epi + part + 4000.0 + 1500.0 * (R + kd - 1));
                if (cost < best * 0.97) { best = cost; best_r = R; best_ks = ks; }
                if (ks == ks_only && cost < best_f * 0.97) { best_f = cost; best_fr = R; }
            }
        }
        if (ks_only && best_fr) { best_r = best_fr; best_ks = ks_only; }
        ksplit = best_ks;
        if (best_r != s.R) {
            s.R = best_r;
            s.acc_bufs = (2 * s.R * s.NP <= 512) ? 2 : 1;
            umma_ring(s, taps, kd, s.CC / EPU);
        }
    }
    p.ksplit = ksplit; p.ws_acc = ws_acc; p.ws_cnt = ws_cnt;
    { const char* e = getenv("B200EM_PREFETCH"); p.prefetch = e ? atoi(e) : 1; }
    p.R = s.R; p.NP = s.NP; p.CC = s.CC; p.nchunks = Cin / s.CC; p.G = s.G; p.acc_bufs = s.acc_bufs; p.nstage = s.nstage;
    p.tiles_w = (W + TW - 1) / TW; p.tiles_h = (H + TH - 1) / TH; p.tiles_d = (D + s.R - 1) / s.R;
    p.items = (long long)N * p.tiles_d * p.tiles_h * p.tiles_w;
    B2_CHECK_ARG(p.items < (1LL << 31), "conv: too many work items for 32-bit indexing");
    p.a_bytes = s.a_bytes; p.b_stage_bytes = s.b_stage_bytes;
    long long gx = p.items < sm_count() / (s.nblk * ksplit) ? p.items : sm_count() / (s.nblk * ksplit);
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)s.nblk, (unsigned)ksplit);
    // tiled tensor map over x: dims (C, W, H, D, N), box (one 16-byte channel unit, 10, 18, 1, 1) -> one TMA load per (slice, plane)
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    p.use_tma = 0;
    {
        const char* tma_e = getenv("B200EM_UMMA_TMA");      // "0": the cp.async fallback of the tile loader (read per launch: tests switch it)
        const bool tma_env = !(tma_e && atoi(tma_e) == 0);
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static const EncodeFn encode = [] {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
                return (EncodeFn)fn;
            (void)cudaGetLastError();
            return (EncodeFn) nullptr;
        }();
        if (tma_env && encode) {
            const cuuint64_t esz = sizeof(TA);
            const cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
            const cuuint64_t gstr[4] = {(cuuint64_t)x_ld * esz, (cuuint64_t)W * x_ld * esz, (cuuint64_t)H * W * x_ld * esz,
                                        (cuuint64_t)D * H * W * x_ld * esz};
            const cuuint32_t box[5] = {(cuuint32_t)EPU, (cuuint32_t)WP, (cuuint32_t)HP, 1, 1};
            const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            const CUtensorMapDataType dt = F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                               : (std::is_same<TA, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
            const CUresult r = encode(&tmap, dt, 5, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) p.use_tma = 1;
        }
    }
#define B2_UMMA_LAUNCH_(R_, KC_, TMA_)                                                                                             \
    do {                                                                                                                           \
        B2_CUDA(cudaFuncSetAttribute(conv3d_umma_kernel<TA, TO, R_, KC_, TMA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM)); \
        conv3d_umma_kernel<TA, TO, R_, KC_, TMA_><<<grid, THREADS, s.smem_bytes, (cudaStream_t)stream>>>(p, tmap);                       \
    } while (0)
#define B2_UMMA_LAUNCH(R_, KC_)                                                                                                    \
    do {                                                                                                                           \
        if (p.use_tma) B2_UMMA_LAUNCH_(R_, KC_, true);                                                                             \
        else B2_UMMA_LAUNCH_(R_, KC_, false);                                                                                      \
    } while (0)
    const int kc = s.CC / (2 * EPU);            // MMA K steps per chunk (two 16-byte planes each)
    if (s.R == 4 && kc == 2) B2_UMMA_LAUNCH(4, 2);
    else if (s.R == 4) B2_UMMA_LAUNCH(4, 1);
    else if (s.R == 2 && kc == 2) B2_UMMA_LAUNCH(2, 2);
    else if (s.R == 2) B2_UMMA_LAUNCH(2, 1);
    else if (kc == 2) B2_UMMA_LAUNCH(1, 2);
    else B2_UMMA_LAUNCH(1, 1);
#undef B2_UMMA_LAUNCH_
#undef B2_UMMA_LAUNCH
    B2_LAUNCH_CHECK();
    return 0;
}


template <typename TA>
static int launch_wgrad_umma(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld, float* dw,
                             float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, void* stream, int fp16 = 0,
                             const float* x_absmax = nullptr, const float* z_absmax = nullptr) {
    constexpr int EPU = 16 / (int)sizeof(TA);
    constexpr int WG_R = WgR<TA>::value;
    B2_CHECK_ARG(x && dz && dw && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_wgrad_umma: bad arguments");
    B2_CHECK_ARG((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3), "conv3d_wgrad_umma: kernel dims must be 1 or 3");
    int NB;
    if (!wgrad_umma_shape(Cin, Cout, NB)) {
        set_error("conv3d_wgrad_umma: channel counts (%d -> %d) not supported by the tcgen05 path", Cin, Cout);
        return 2;
    }
    B2_CHECK_ARG(x_ld % EPU == 0 && dz_ld % EPU == 0 && aligned16(x) && aligned16(dz), "conv3d_wgrad_umma: activations must be 16-byte aligned with a pitch that keeps them so");
    WgradUmmaParams p;
    p.x = x; p.x_ld = x_ld; p.in_ss = in_scale_shift; p.dz = dz; p.dz_ld = dz_ld;
    p.dw = dw; p.db = db;
    p.fp16 = fp16; p.x_absmax = x_absmax; p.z_absmax = z_absmax;
    B2_CHECK_ARG(!fp16 || !in_scale_shift, "conv3d_wgrad_umma: fp16 operands take no fused norm apply (b200em_cvt_f16 applies it)");
    p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = kh; p.kw = kw;
    p.NB = NB; p.nco = Cout / NB;
    p.tiles_w = (W + TW - 1) / TW; p.tiles_h = (H + TH - 1) / TH; p.tiles_d = (D + WG_R - 1) / WG_R;
    p.items = (long long)N * p.tiles_d * p.tiles_h * p.tiles_w;
    B2_CHECK_ARG(p.items < (1LL << 31), "conv: too many work items for 32-bit indexing");
    // 1x1x1 filters (the up-sampler convs): there are no depth taps to stack along M, so the 128 MMA rows hold up to 128 input
    // channels of ONE slice instead of 32 channels of four -- 4x fewer CTAs per layer (each re-reads the dz tile), 4x the split over
    // voxel tiles (these layers are a serial chain of small tile loads per CTA), up to all rows useful instead of a quarter
    p.CH = 32;
    if (kd == 1 && kh == 1 && kw == 1 && sizeof(TA) == 2 && !getenv("B200EM_WG_NARROW")) p.CH = Cin % 128 == 0 ? 128 : (Cin % 64 == 0 ? 64 : 32);
    const int J = p.CH / EPU;
    const int span = p.CH > 32 ? 128 / p.CH : 4;  // slices an MMA's 128 rows span
    p.x_bytes = (WG_R + kd - 1) * J * PLANE;
    p.dz_bytes = WG_R * (NB / EPU) * WG_DZ_PLANE;
    p.pad_bytes = (span - kd) * J * PLANE;       // slab r reads slices r .. r + span - 1; WG_R + kd - 1 are loaded
    { const char* e = getenv("B200EM_DEBUG"); p.debug = e ? atoi(e) : 0; }
    const int smem_bytes = 2 * p.x_bytes + p.pad_bytes + 2 * p.dz_bytes + NB * 4 + 8 * 8 + 16 + 9 * 4 + 128;
    B2_CHECK_ARG(smem_bytes <= MAX_SMEM, "conv3d_wgrad_umma: shared memory budget exceeded");
    B2_CUDA(cudaFuncSetAttribute(conv3d_wgrad_umma_kernel<TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    const int pairs = (Cin / p.CH) * p.nco;
    long long splits = sm_count() / pairs;
    if (splits < 1) splits = 1;
    if (splits > p.items) splits = p.items;
    dim3 grid((unsigned)splits, (unsigned)pairs, 1);
    conv3d_wgrad_umma_kernel<TA><<<grid, THREADS, smem_bytes, (cudaStream_t)stream>>>(p);
    B2_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int b200em_conv3d_umma_supported(int Cin, int Cout, int kd, int kh, int kw) {
    UmmaShape s;
    return umma_shape(Cin, Cout, kd, kh, kw, s) ? 1 : 0;
}

int b200em_conv3d_umma_pack(const float* w, int Cout, int Cin, int kd, int kh, int kw, int dgrad, void* packed, void* stream) {
    B2_CHECK_ARG(w && packed && Cout > 0 && Cin > 0, "conv3d_umma_pack: bad arguments");
    UmmaShape s;
    const int n_ = dgrad ? Cin : Cout, k_ = dgrad ? Cout : Cin;
    if (!umma_shape(k_, n_, kd, kh, kw, s)) {
        set_error("conv3d_umma_pack: channel counts (%d -> %d) not supported by the tcgen05 path", k_, n_);
        return 2;
    }
    const int taps = kd * kh * kw;
    B2_CHECK_ARG(taps <= 27 && Cout % 8 == 0 && Cin % 8 == 0 && aligned16(packed), "conv3d_umma_pack: needs taps <= 27, channels % 8 == 0 and a 16-byte aligned output");
    pack_umma_weights_kernel<<<dim3((unsigned)(Cout / 8), (unsigned)(Cin / 8)), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, taps, dgrad, s.CC, s.NP,
                                                                                                          (__nv_bfloat16*)packed);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_conv3d_umma(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                       void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W, int Cin,
                       int Cout, int kd, int kh, int kw, int relu, void* stream) {
    return launch_conv_umma<__nv_bfloat16>(x, x_ld, in_scale_shift, w_packed, bias, y, y_ld, sums, dot_x, dot_ld, N, D, H, W, Cin, Cout, kd,
                                           kh, kw, relu, stream);
}

int b200em_conv3d_umma_tf32_supported(int Cin, int Cout, int kd, int kh, int kw) {
    UmmaShape s;
    return umma_shape(Cin, Cout, kd, kh, kw, s, true) ? 1 : 0;
}

int b200em_conv3d_umma_tf32(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                            void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W, int Cin,
                            int Cout, int kd, int kh, int kw, int relu, void* stream) {
    return launch_conv_umma<float>(x, x_ld, in_scale_shift, w_packed, bias, y, y_ld, sums, dot_x, dot_ld, N, D, H, W, Cin, Cout, kd, kh, kw,
                                   relu, stream);
}

int b200em_conv3d_umma_h16(const void* x_f16, int64_t x_ld, const float* x_absmax, const void* w_packed, const float* bias, float* y,
                           int64_t y_ld, float* sums, const float* dot_x, int64_t dot_ld, int N, int D, int H, int W, int Cin, int Cout,
                           int kd, int kh, int kw, int relu, void* stream) {
    return launch_conv_umma<__half, float>(x_f16, x_ld, nullptr, w_packed, bias, y, y_ld, sums, dot_x, dot_ld, N, D, H, W, Cin, Cout, kd, kh,
                                           kw, relu, stream, x_absmax);
}

int b200em_conv3d_wgrad_umma_supported(int Cin, int Cout, int kd, int kh, int kw) {
    int NB;
    return wgrad_umma_shape(Cin, Cout, NB) ? 1 : 0;
}

int b200em_conv3d_wgrad_umma(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld, float* dw,
                             float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, void* stream) {
    return launch_wgrad_umma<__nv_bfloat16>(x, x_ld, in_scale_shift, dz, dz_ld, dw, db, N, D, H, W, Cin, Cout, kd, kh, kw, stream);
}

int b200em_conv3d_wgrad_umma_h16(const void* x_f16, int64_t x_ld, const float* x_absmax, const void* dz_f16, int64_t dz_ld,
                                 const float* dz_absmax, float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd,
                                 int kh, int kw, void* stream) {
    return launch_wgrad_umma<__nv_bfloat16>(x_f16, x_ld, nullptr, dz_f16, dz_ld, dw, db, N, D, H, W, Cin, Cout, kd, kh, kw, stream, 1,
                                            x_absmax, dz_absmax);
}

}  // extern "C"
