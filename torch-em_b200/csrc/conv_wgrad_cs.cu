// Weight gradient of the 3 x 3 x 3 convolutions on the tensor cores, "h-stacked" and depth-streamed:
//     dW[co][ci][a,b,c] += sum_{n,v} dz[n,v,co] * x_hat[n, v + (a-1, b-1, c-1), ci]        (+ db[co] += sum dz)
//
// GEMM view per (32-channel chunk of Cin, 32-channel block of Cout), K = voxels (16 per MMA: two 8-voxel w-rows H, H+1):
//   A (MN-major) = x_hat:  M = 128 rows = 4 consecutive depth slices x 32 channels at tile row H, w shifted by the tap c
//       (descriptor start address) -> rows [0,96) are the depth taps a = 0,1,2 of output slice r; rows [96,128) are ignored
//   B (MN-major) = dz:     N = 96 = 3 h-taps x 32 output channels.  The haloed dz slice is stored [row hp][8-ch group jo]
//       [w][8 ch]: N-group g = 4*b' + jo at stride 128 B is then 8-ch group jo of row hp + b' -- the three h-shifted
//       operands are ONE buffer, no copies (tap b = 2 - b').  One tcgen05.mma covers three taps: an MMA instruction costs
//       max(N/2, ~57 + N/8) cycles whatever N is, so N = 96 instead of 32 is 2.7x fewer tensor-pipe cycles per FLOP.
//   D[c] in TMEM: 3 accumulators x 96 columns, resident over ALL work items of the CTA; one epilogue at the end adds
//       them into the fp32 dW (torch layout) with atomics.
// Depth streaming: a work item is a 16 x 8 (h, w) tile x DR consecutive depth slices; stage t of an item holds input
// slice d0-1+t of x_hat (4 planes of 8 channels, loaded ONCE per item) and dz slice d0+t.  The x parts of consecutive
// ring slots are contiguous in shared memory and the first three slots are mirrored behind the last one, so the
// 4-slice operand window of any output slice is one contiguous block (descriptor stride = plane).
// Warp roles: warps 0-3 epilogue (once, at the end), warps 4-11 loaders (each warp owns every nlw-th stage), warp 12
// MMA issuer.  Replaces the autograd weight/bias gradient of nn.Conv3d (unet.py:429-438) for the 3x3x3 layers.
#include <cuda.h>      // CUtensorMap (types only; the encoder comes from cudaGetDriverEntryPoint)
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200em {

using namespace umma;

namespace {
constexpr int CS_TH = 16, CS_TW = 8;
constexpr int CS_HP = CS_TH + 2, CS_WP = CS_TW + 2;
constexpr int CS_XVOX = CS_TH * CS_WP;                 // x tile: 16 rows x 10 columns (w halo only)
constexpr int CS_PLANE = CS_XVOX * 16;                 // bytes per 8-channel plane of an x slice: 2560 = 20 x 128 (TMA destination alignment)
constexpr int CS_XSLOT = 4 * CS_PLANE;                 // one x slice of a 32-channel chunk
constexpr int CS_NB = 32;                              // output channels per CTA
constexpr int CS_ZROW = (CS_NB / 8) * CS_TW * 16;      // one haloed dz row: [jo][w][8 ch] = 512 B
constexpr int CS_ZSLOT = CS_HP * CS_ZROW;              // one dz slice with its h halo
constexpr int CS_NLW = 8;
constexpr int CS_THREADS = 128 + CS_NLW * 32 + 32;
constexpr int CS_W_MMA = 4 + CS_NLW;
constexpr int CS_MAX_NS = 8;
constexpr int CS_MAX_SMEM = 227 * 1024;
}  // namespace

struct WgradCsParams {
    const __nv_bfloat16* x; long long x_ld;
    const float* in_ss;
    const __nv_bfloat16* dz; long long dz_ld;
    float* dw;
    float* db;
    int N, D, H, W, Cin, Cout;
    int nco;                                   // number of 32-wide Cout blocks
    int DR, NS, nlw;
    int tiles_w, tiles_h, tiles_d;
    long long items;
    int use_tma;                               // 1: tiles loaded by cp.async.bulk.tensor (tensor maps over x and dz), else cp.async
    int debug;                                 // bring-up switches (env B200EM_DEBUG): 1 no operand loads, 4 no MMAs, 8 profile printf
    int fp16;                                  // operands are IEEE fp16 (the h16 path of fp32 activations), not bf16
    const float* x_absmax; const float* z_absmax;   // h16: device max |.| the operands were scaled by (common.cuh h16_shift), or null
};

__global__ void __launch_bounds__(CS_THREADS, 1) conv3d_wgrad_cs_kernel(const WgradCsParams p, const __grid_constant__ CUtensorMap tmx,
                                                                       const __grid_constant__ CUtensorMap tmz) {
    extern __shared__ __align__(128) uint8_t smem[];
    // carve: X[NS + 3] (slots 0..2 mirrored at NS..NS+2) | Z[NS] | db sums[32] | barriers | tmem ptr
    const int NS = p.NS, DR = p.DR;
    uint8_t* smX = smem;
    uint8_t* smZ = smX + (NS + 3) * CS_XSLOT;
    float* s_db = reinterpret_cast<float*>(smZ + NS * CS_ZSLOT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_db + CS_NB);
    uint64_t* full = bars;                     // [NS] plain flags: fill number of the slot
    uint64_t* empty = bars + CS_MAX_NS;        // [NS] tcgen05.commit
    uint64_t* acc_full = bars + 2 * CS_MAX_NS; // [1]
    uint64_t* t_full = acc_full + 1;           // [NS] TMA tile loads (expect_tx)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_full + CS_MAX_NS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.y / p.nco, cob = blockIdx.y % p.nco;
    constexpr int N3 = 3 * CS_NB;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) { flag_init(&full[i]); mbar_init(&empty[i], 1); mbar_init(&t_full[i], 1); }   // full: plain fill-number flags (umma.cuh)
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == CS_W_MMA) tmem_alloc(s_tmem, 512);
    for (int i = threadIdx.x; i < CS_NB; i += CS_THREADS) s_db[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    auto coords = [&](long long item_, int& n, int& d0, int& h0, int& w0) {      // 32-bit: 64-bit division is a software routine
        unsigned item = (unsigned)item_;
        const unsigned tw = item % (unsigned)p.tiles_w; item /= (unsigned)p.tiles_w;
        const unsigned th = item % (unsigned)p.tiles_h; item /= (unsigned)p.tiles_h;
        const unsigned td = item % (unsigned)p.tiles_d; item /= (unsigned)p.tiles_d;
        n = (int)item; d0 = (int)td * DR; h0 = (int)th * CS_TH; w0 = (int)tw * CS_TW;
    };

    if (warp >= 4 && warp < CS_W_MMA) {
        // ===================== loaders =====================
        const int w8 = warp - 4;
        constexpr int XU = CS_XVOX * 4 / 32;                 // x units (16 B) per lane and stage: 20
        constexpr int ZU = CS_HP * CS_TW * 4 / 32;           // dz units per lane and stage: 18 (unit i = haloed row i)
        const int j = lane >> 3, vl = lane & 7;              // 8-channel group of this lane, first voxel: a quarter warp = 8 consecutive voxels of one plane = one 128-byte shared-memory line
        float sc[8], sh[8];
        float dbacc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { dbacc[e] = 0.f; sc[e] = 1.f; sh[e] = 0.f; }
        int cur_n = -1;
        uint32_t slot = 0, phase = 1, lap = 1;
        int owner = 0;
        const size_t xslice = (size_t)p.H * p.W * p.x_ld, zslice = (size_t)p.H * p.W * p.dz_ld;
        const bool want_db = p.db != nullptr && chunk == 0;
        const bool prof = (p.debug & 8) != 0;
        long long pf_wait = 0, pf_load = 0, pf_rest = 0, pf_n = 0, pf_t0 = clock64();
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
            int n, d0, h0, w0;
            coords(item, n, d0, h0, w0);
            int goff[XU];                            // in-slice element offset of this lane's x units, -1 = outside the volume
#pragma unroll
            for (int i = 0; i < XU; ++i) {
                const int v = vl + 8 * i;
                const int hl = v / CS_WP, wp_ = v % CS_WP;
                const int gh = h0 + hl, gw = w0 + wp_ - 1;
                goff[i] = (gh < p.H && gw >= 0 && gw < p.W) ? (int)((gh * p.W + gw) * p.x_ld) : -1;
            }
            // dz unit i of this lane = haloed row hp = i (global row h0 + i - 1), column w0 + vl
            const int zgw = w0 + vl;
            const int zoff0 = (int)(((h0 - 1) * p.W + zgw) * p.dz_ld), zrow = (int)(p.W * p.dz_ld);
            uint32_t zmask = 0;                      // row i inside the volume (and the column)
#pragma unroll
            for (int i = 0; i < ZU; ++i) {
                const int gh = h0 + i - 1;
                if (gh >= 0 && gh < p.H && zgw < p.W) zmask |= 1u << i;
            }
            if (p.in_ss && n != cur_n) {
                const float* q = p.in_ss + ((size_t)n * p.Cin + chunk * 32 + j * 8) * 2;
#pragma unroll
                for (int e = 0; e < 8; ++e) { sc[e] = q[2 * e]; sh[e] = q[2 * e + 1]; }
                cur_n = n;
            }
            for (int t = 0; t < DR + 2; ++t) {
                const uint32_t my_slot = slot, my_phase = phase, my_lap = lap;
                const bool mine = owner == w8;
                if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1; ++lap; }
                if (++owner == p.nlw) owner = 0;
                if (!mine) continue;
                long long pa = 0, pb = 0, pc = 0;
                if (prof) pa = clock64();
                mbar_wait(&empty[my_slot], my_phase);
                if (prof) pb = clock64();
                const int gd = d0 - 1 + t;
                const bool in_d = gd >= 0 && gd < p.D;
                const int gz = d0 + t;
                const bool z_on = t < DR && gz < p.D;
                const __nv_bfloat16* xs = p.x + ((size_t)n * p.D + (in_d ? gd : 0)) * xslice + chunk * 32 + j * 8;
                const __nv_bfloat16* zs = p.dz + ((size_t)n * p.D + (z_on ? gz : 0)) * zslice + cob * CS_NB + j * 8;
                uint8_t* xdst = smX + my_slot * CS_XSLOT + j * CS_PLANE + vl * 16;
                uint8_t* zdst = smZ + my_slot * CS_ZSLOT + j * (CS_TW * 16) + vl * 16;
                const uint32_t xd32 = smem_u32(xdst), zd32 = smem_u32(zdst);
                const uint32_t mirror = my_slot < 3 ? (uint32_t)(NS * CS_XSLOT) : 0u;
                if (p.use_tma) {
                    // x slice: one tiled load per 8-channel plane, box (8 ch, 10 w, 16 h, 1 d, 1 n); dz slice: ONE load, box
                    // (8 ch, 8 w, 4 groups, 18 h, 1) over the view (8, W, Cout/8, H, N*D) -> lands as [row][group][w][8 ch];
                    // halo rows / columns and out-of-volume x slices are zero-filled by the TMA unit
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&t_full[my_slot], (uint32_t)(CS_XSLOT + (z_on ? CS_ZSLOT : 0)));
                        const uint32_t bar32 = smem_u32(&t_full[my_slot]);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            asm volatile(
                                "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
                                    "r"(smem_u32(smX + my_slot * CS_XSLOT + jj * CS_PLANE)),
                                "l"(reinterpret_cast<uint64_t>(&tmx)), "r"(chunk * 32 + jj * 8), "r"(w0 - 1), "r"(h0), "r"(gd), "r"(n), "r"(bar32)
                                : "memory");
                        if (z_on)
                            asm volatile(
                                "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
                                    "r"(smem_u32(smZ + my_slot * CS_ZSLOT)),
                                "l"(reinterpret_cast<uint64_t>(&tmz)), "r"(0), "r"(w0), "r"(cob * (CS_NB / 8)), "r"(h0 - 1), "r"(n * p.D + gz), "r"(bar32)
                                : "memory");
                    }
                    if (!z_on && t < DR) {           // dz slice beyond the volume (D % DR != 0): zeros
#pragma unroll
                        for (int i = 0; i < ZU; ++i) *reinterpret_cast<uint4*>(zdst + i * CS_ZROW) = make_uint4(0, 0, 0, 0);
                    }
                    mbar_wait(&t_full[my_slot], (my_lap - 1) & 1);
                } else if (!(p.debug & 1)) {
#pragma unroll
                    for (int i = 0; i < ZU; ++i) {
                        const bool in = z_on && ((zmask >> i) & 1);
                        const __nv_bfloat16* src = in ? zs + zoff0 + i * zrow : p.dz;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(zd32 + (uint32_t)(i * CS_ZROW)), "l"(src),
                                     "r"(in ? 16 : 0)
                                     : "memory");
                    }
#pragma unroll
                    for (int i = 0; i < XU; ++i) {
                        const bool in = in_d && goff[i] >= 0;
                        const __nv_bfloat16* src = in ? xs + goff[i] : p.x;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xd32 + (uint32_t)(8 * i * 16)), "l"(src),
                                     "r"(in ? 16 : 0)
                                     : "memory");
                    }
                }
                if (!p.use_tma) asm volatile("cp.async.wait_all;" ::: "memory");
                if (prof) pc = clock64();
                if (in_d && (p.in_ss || mirror)) {
                    // norm apply in place (fp32 math, one rounding to bf16) and the mirror copy of the first three slots
#pragma unroll
                    for (int i = 0; i < XU; ++i) {
                        uint4* qd = reinterpret_cast<uint4*>(xdst + 8 * i * 16);
                        uint4 val = *qd;
                        if (p.in_ss && goff[i] >= 0) {
                            uint32_t* w32 = reinterpret_cast<uint32_t*>(&val);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float lo = fmaf(__uint_as_float(w32[e] << 16), sc[2 * e], sh[2 * e]);
                                const float hi = fmaf(__uint_as_float(w32[e] & 0xffff0000u), sc[2 * e + 1], sh[2 * e + 1]);
                                const __nv_bfloat162 r2 = __floats2bfloat162_rn(lo, hi);
                                w32[e] = *reinterpret_cast<const uint32_t*>(&r2);
                            }
                            *qd = val;
                        }
                        if (mirror) *reinterpret_cast<uint4*>(xdst + mirror + 8 * i * 16) = val;
                    }
                } else if (mirror) {                 // slice outside the volume: zeros in the mirror slot as well
#pragma unroll
                    for (int i = 0; i < XU; ++i) *reinterpret_cast<uint4*>(xdst + mirror + 8 * i * 16) = make_uint4(0, 0, 0, 0);
                }
                if (want_db && z_on) {               // bias gradient from the tile's own rows (hp = 1..16): every in-volume dz element once
#pragma unroll
                    for (int i = 1; i <= CS_TH; ++i) {
                        const uint4 val = *reinterpret_cast<const uint4*>(zdst + i * CS_ZROW);
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&val);
                        const __half2* g2 = reinterpret_cast<const __half2*>(&val);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = p.fp16 ? __half22float2(g2[e]) : __bfloat1622float2(h2[e]);
                            dbacc[2 * e] += f.x;
                            dbacc[2 * e + 1] += f.y;
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) flag_store(&full[my_slot], my_lap);
                if (prof) { pf_wait += pb - pa; pf_load += pc - pb; pf_rest += clock64() - pc; ++pf_n; }
            }
        }
        if (want_db) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_db[j * 8 + e], dbacc[e]);
            asm volatile("bar.sync 2, %0;" ::"n"(CS_NLW * 32) : "memory");
            const int t = threadIdx.x - 128;
            if (t < CS_NB) atomicAdd(p.db + cob * CS_NB + t, s_db[t] * pow2i(-h16_shift(p.z_absmax)));
        }
        if (prof && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && w8 == 0)
            printf("[cs prof] loader warp 0: total %lld cyc, %lld stages: wait empty %lld, load %lld, norm/db/arrive %lld\n", clock64() - pf_t0,
                   pf_n, pf_wait, pf_load, pf_rest);
        asm volatile("bar.sync 4, %0;" ::"n"(CS_THREADS) : "memory");      // ring hand-over to the epilogue (see there)
    } else if (warp == CS_W_MMA) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = p.fp16 ? make_idesc_f16(128, N3, 1, 1) : make_idesc_bf16(128, N3, 1, 1);      // both operands MN-major
            // A: LBO = next 8 voxels (next tile row), SBO = next 8 channels (next plane; slices are consecutive planes)
            // B: LBO = next 8 voxels (next dz row), SBO = 128 B = next 8 output channels, and after four of them the next row
            const uint64_t ad = make_desc(0, CS_WP * 16, CS_PLANE), bd = make_desc(0, CS_ZROW, CS_TW * 16);
            const uint32_t a_hi = (uint32_t)(ad >> 32), a_lo_c = (uint32_t)(ad & 0xFFFFFFFFu) + (smem_u32(smX) >> 4);
            const uint32_t b_hi = (uint32_t)(bd >> 32), b_lo_c = (uint32_t)(bd & 0xFFFFFFFFu) + (smem_u32(smZ) >> 4);
            uint32_t slot = 0, lap = 1;              // ring position and fill number of the next stage to wait for
            uint32_t rslot = 0;                      // ring position of the output slice being issued (= its first x slice)
            bool first = true;                       // very first MMAs of this CTA overwrite the accumulators
            const bool prof = (p.debug & 8) != 0;
            long long pf_w = 0, pf_t0 = clock64(), pf_n = 0;
            for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
                for (int t = 0; t < DR + 2; ++t) {
                    long long t_ = 0;
                    if (prof) t_ = clock64();
                    flag_wait_eq(&full[slot], lap);
                    if (prof) { pf_w += clock64() - t_; ++pf_n; }
                    tc_fence_after();
                    if (++slot == (uint32_t)NS) { slot = 0; ++lap; }
                    if (t < 2) continue;
                    // output slice r = t - 2: x slices r, r+1, r+2 (+ one ignored) start at ring slot rslot, dz copies in slot rslot
                    const uint32_t xs = a_lo_c + rslot * (CS_XSLOT / 16);
                    const uint32_t zs = b_lo_c + rslot * (CS_ZSLOT / 16);
                    if (!(p.debug & 4)) {
                        // K step outermost, w tap innermost: consecutive MMAs share the B operand (the dz rows of this K step)
#pragma unroll
                        for (int ks = 0; ks < CS_TH / 2; ++ks) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {    // w tap: start column of the haloed x rows
                                const uint32_t a0 = xs + (uint32_t)c + ks * 2 * CS_WP;
                                const uint32_t tacc = tmem_base + (uint32_t)(c * N3);
                                if (ks == 0 && first) umma_bf16_c<false>(tacc, a0, a_hi, zs, b_hi, idesc);
                                else umma_bf16_c<true>(tacc, a0, a_hi, zs + ks * 2 * (CS_ZROW / 16), b_hi, idesc);
                            }
                        }
                        first = false;
                    }
                    umma_commit(&empty[rslot]);      // x slice r and dz slice r are free once these MMAs have completed
                    if (++rslot == (uint32_t)NS) rslot = 0;
                }
                // the item's last two x slices were only ever read as depth taps a = 1, 2: release their stages as well
                umma_commit(&empty[rslot]);
                if (++rslot == (uint32_t)NS) rslot = 0;
                umma_commit(&empty[rslot]);
                if (++rslot == (uint32_t)NS) rslot = 0;
            }
            umma_commit(acc_full);
            if (prof && blockIdx.x == 0 && blockIdx.y == 0)
                printf("[cs prof] mma: total %lld cyc, %lld stages: wait full %lld\n", clock64() - pf_t0, pf_n, pf_w);
        }
        __syncwarp();
        asm volatile("bar.sync 4, %0;" ::"n"(CS_THREADS) : "memory");
    } else if (warp < 4) {
        // ===================== epilogue: TMEM -> shared-memory transpose -> coalesced vector reductions into dW =====================
        // dW is (Cout, Cin, 27) fp32: for one output channel the CTA's 32 input channels x 27 taps are 864 CONTIGUOUS floats.  A
        // lane holds ONE input channel, so adding straight from the TMEM registers would scatter 4-byte atomics 108 bytes apart
        // (one L2 transaction each: the deep, weight-heavy levels were bound by exactly that).  Instead the accumulators are
        // transposed through the (now idle) operand ring -- every MMA has completed when acc_full fires, so the ring is free --
        // and leave as 16-byte red.global.add.v4.f32 over contiguous memory.
        mbar_wait(acc_full, 0);
        tc_fence_after();
        // acc_full (tcgen05.commit of the last MMA) already orders every loader write and every tensor-core read of the ring before
        // this point; the CTA-wide barrier states the same hand-over in terms compute-sanitizer's racecheck understands (the loader
        // and MMA warps arrive when their loops end -- nobody waits for long)
        asm volatile("bar.sync 4, %0;" ::"n"(CS_THREADS) : "memory");
        const int a = warp;                              // depth tap of this warp's 32 lanes (a == 3: ignored rows)
        constexpr int taps = 27, ROW = 32 * taps;        // floats per output channel of this CTA's (ci chunk, co block) tile
        float* stage = reinterpret_cast<float*>(smem);   // [32 co][32 ci][27 taps] = 110,592 B <= (NS + 3) * CS_XSLOT + NS * CS_ZSLOT
        const float osc = pow2i(-h16_shift(p.x_absmax) - h16_shift(p.z_absmax));
        if (a < 3) {
            for (int c = 0; c < 3; ++c) {
                for (int bq = 0; bq < 3; ++bq) {             // N block b' = h shift of the dz rows: tap b = 2 - b'
                    const int tap = (a * 3 + (2 - bq)) * 3 + c;
                    for (int cb = 0; cb < CS_NB; cb += 16) {
                        uint32_t raw[16];
                        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * N3 + bq * CS_NB + cb), raw);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) stage[(cb + i) * ROW + lane * taps + tap] = __uint_as_float(raw[i]) * osc;   // lane stride 27 words: conflict-free
                    }
                }
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int t = threadIdx.x;                       // 0..127
        float* dwb = p.dw + ((size_t)(cob * CS_NB) * p.Cin + chunk * 32) * taps;
        const size_t co_stride = (size_t)p.Cin * taps;
        // (fire-and-forget reductions also where this CTA is the tile's only contributor: a read-add-write of the 110 KB tile is a
        // chain of dependent round trips per thread and measured 3x slower on the 2048 -> 2048 layer)
        if ((reinterpret_cast<uintptr_t>(p.dw) & 15) == 0) {
            for (int i = t; i < CS_NB * (ROW / 4); i += 128) {
                const int co = i / (ROW / 4), q = i % (ROW / 4);
                const float4 v = *reinterpret_cast<const float4*>(stage + co * ROW + q * 4);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dwb + co * co_stride + q * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                             : "memory");
            }
        } else {
            for (int i = t; i < CS_NB * ROW; i += 128) atomicAdd(dwb + (i / ROW) * co_stride + i % ROW, stage[i]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CS_W_MMA) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static bool wgrad_cs_shape(int Cin, int Cout, int kd, int kh, int kw) {
    return kd == 3 && kw == 3 && kh == 3 && Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32;
}

}  // namespace b200em

using namespace b200em;

static int launch_wgrad_cs(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld, float* dw,
                           float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, void* stream, int fp16,
                           const float* x_absmax, const float* z_absmax) {
    B2_CHECK_ARG(x && dz && dw && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_wgrad_cs: bad arguments");
    if (!wgrad_cs_shape(Cin, Cout, kd, kh, kw)) {
        set_error("conv3d_wgrad_cs: shape (%d -> %d, %dx%dx%d) not supported by the w-stacked tcgen05 weight gradient", Cin, Cout, kd, kh, kw);
        return 2;
    }
    B2_CHECK_ARG(x_ld % 8 == 0 && dz_ld % 8 == 0 && aligned16(x) && aligned16(dz), "conv3d_wgrad_cs: activations must be 16-byte aligned with pitch % 8 == 0");
    B2_CHECK_ARG(x_ld >= Cin && dz_ld >= Cout, "conv3d_wgrad_cs: pitch smaller than channel count");
    B2_CHECK_ARG((long long)H * W * x_ld < (1LL << 31) && (long long)H * W * dz_ld < (1LL << 31), "conv3d_wgrad_cs: slice too large for 32-bit in-slice offsets");
    WgradCsParams p;
    p.x = (const __nv_bfloat16*)x; p.x_ld = x_ld; p.in_ss = in_scale_shift; p.dz = (const __nv_bfloat16*)dz; p.dz_ld = dz_ld;
    p.dw = dw; p.db = db;
    p.fp16 = fp16; p.x_absmax = x_absmax; p.z_absmax = z_absmax;
    B2_CHECK_ARG(!fp16 || !in_scale_shift, "conv3d_wgrad_cs: fp16 operands take no fused norm apply (b200em_cvt_f16 applies it)");
    p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    p.nco = Cout / CS_NB;
    const int misc = CS_NB * 4 + (3 * CS_MAX_NS + 1) * 8 + 8 + 128;
    int ns = (CS_MAX_SMEM - misc - 3 * CS_XSLOT) / (CS_XSLOT + CS_ZSLOT);
    if (ns > CS_MAX_NS) ns = CS_MAX_NS;
    B2_CHECK_ARG(ns >= 4 && (ns + 3) * CS_XSLOT + ns * CS_ZSLOT >= CS_NB * 32 * 27 * 4, "conv3d_wgrad_cs: shared memory budget exceeded");   // the ring doubles as the epilogue's transpose stage
    p.NS = ns;
    p.nlw = ns < CS_NLW ? ns : CS_NLW;
    const int smem_bytes = (ns + 3) * CS_XSLOT + ns * CS_ZSLOT + misc;
    const int pairs = (Cin / 32) * p.nco;
    long long splits = sm_count() / pairs;
    if (splits < 1) splits = 1;
    // depth slices per work item: enough items that every split has the same number of them, long enough that the two
    // extra input slices of an item amortise
    const long long cols = (long long)N * ((H + CS_TH - 1) / CS_TH) * ((W + CS_TW - 1) / CS_TW);
    {
        const char* e = getenv("B200EM_CS_DR");
        int best = 1;
        double best_cost = 1e30;
        for (int dr = 1; dr <= (D < 64 ? D : 64); ++dr) {
            const long long items = cols * ((D + dr - 1) / dr);
            const long long waves = (items + splits - 1) / splits;
            const double cost = (double)waves * (dr + 2.0);
            if (cost < best_cost - 1e-9) { best_cost = cost; best = dr; }
        }
        p.DR = e ? atoi(e) : best;
        if (p.DR < 1) p.DR = 1;
    }
    p.tiles_w = (W + CS_TW - 1) / CS_TW; p.tiles_h = (H + CS_TH - 1) / CS_TH; p.tiles_d = (D + p.DR - 1) / p.DR;
    p.items = (long long)N * p.tiles_d * p.tiles_h * p.tiles_w;
    B2_CHECK_ARG(p.items < (1LL << 31), "conv: too many work items for 32-bit indexing");
    if (splits > p.items) splits = p.items;
    { const char* e = getenv("B200EM_DEBUG"); p.debug = e ? atoi(e) : 0; }
    B2_CUDA(cudaFuncSetAttribute(conv3d_wgrad_cs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_MAX_SMEM));
    dim3 grid((unsigned)splits, (unsigned)pairs, 1);
    // tensor maps: x as (C, W, H, D, N) with box (8, 10, 16, 1, 1); dz as (8, W, Cout/8, H, N*D) with box (8, 8, 4, 18, 1)
    CUtensorMap tmx, tmz;
    memset(&tmx, 0, sizeof(tmx));
    memset(&tmz, 0, sizeof(tmz));
    p.use_tma = 0;
    {
        const char* e = getenv("B200EM_CS_TMA");
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
                const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
                const cuuint64_t xdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
                const cuuint64_t xstr[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2, (cuuint64_t)D * H * W * x_ld * 2};
                const cuuint32_t xbox[5] = {8, (cuuint32_t)CS_WP, (cuuint32_t)CS_TH, 1, 1};
                const cuuint64_t zdim[5] = {8, (cuuint64_t)W, (cuuint64_t)(Cout / 8), (cuuint64_t)H, (cuuint64_t)N * D};
                const cuuint64_t zstr[4] = {(cuuint64_t)dz_ld * 2, 16, (cuuint64_t)W * dz_ld * 2, (cuuint64_t)H * W * dz_ld * 2};
                const cuuint32_t zbox[5] = {8, (cuuint32_t)CS_TW, (cuuint32_t)(CS_NB / 8), (cuuint32_t)CS_HP, 1};
                const CUtensorMapDataType dt16 = fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
                const CUresult r1 = encode(&tmx, dt16, 5, const_cast<void*>(x), xdim, xstr, xbox, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                const CUresult r2 = encode(&tmz, dt16, 5, const_cast<void*>(dz), zdim, zstr, zbox, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) p.use_tma = 1;
            }
        }
    }
    conv3d_wgrad_cs_kernel<<<grid, CS_THREADS, smem_bytes, (cudaStream_t)stream>>>(p, tmx, tmz);
    B2_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int b200em_conv3d_wgrad_cs_supported(int Cin, int Cout, int kd, int kh, int kw) { return wgrad_cs_shape(Cin, Cout, kd, kh, kw) ? 1 : 0; }

int b200em_conv3d_wgrad_cs(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld, float* dw,
                           float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, void* stream) {
    return launch_wgrad_cs(x, x_ld, in_scale_shift, dz, dz_ld, dw, db, N, D, H, W, Cin, Cout, kd, kh, kw, stream, 0, nullptr, nullptr);
}

int b200em_conv3d_wgrad_cs_h16(const void* x_f16, int64_t x_ld, const float* x_absmax, const void* dz_f16, int64_t dz_ld,
                               const float* dz_absmax, float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd,
                               int kh, int kw, void* stream) {
    return launch_wgrad_cs(x_f16, x_ld, nullptr, dz_f16, dz_ld, dw, db, N, D, H, W, Cin, Cout, kd, kh, kw, stream, 1, x_absmax, dz_absmax);
}

}  // extern "C"
