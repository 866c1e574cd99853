// Direct (CUDA-core, fp32 accumulate) NDHWC convolution: forward / data-gradient and weight-gradient.
//
// This is the path for the layers that are NOT tensor-core shaped -- the network's first conv (Cin = 1: 27 MACs per
// output, HBM-bound, SURVEY.md 8a "enc0.conv1") and channel counts the tcgen05 kernel does not take -- and the
// fp32 arithmetic used by the exact-fp32 parity tests.  Same fused prologue/epilogue contract as the tcgen05 kernel:
//   prologue  x_hat = scale[n,c]*x + shift[n,c] on in-bounds voxels (the preceding InstanceNorm/GroupNorm apply;
//             padded zeros stay zero, i.e. they live in normalised space like nn.Conv3d(padding=1) sees them)
//   epilogue  + bias, ReLU, store, per-(n,c) sum / sum-of-squares of the stored values for the NEXT norm.
// Reference semantics: nn.Conv3d(k, padding=k//2) cross-correlation, unet.py:429-438.
#include "common.cuh"

namespace b200em {

namespace {
constexpr int TD = 4, TH = 8, TW = 8;   // output voxels per block (256 threads, one voxel each)
// input channels staged per round: CK = 4 (+1 pad: conflict-free voxel stride), CK = 1 for the network's first conv
constexpr int COT = 32;                  // output channels per block
constexpr int MAXHALO = (TD + 2) * (TH + 2) * (TW + 2);
}  // namespace

template <typename T, int CK>
__global__ void __launch_bounds__(256)
conv3d_direct_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ in_ss,
                     const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ y, int64_t y_ld,
                     float* __restrict__ sums, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw,
                     int relu, int tiles_w, int tiles_h) {
    constexpr int CKP = CK == 1 ? 1 : CK + 1;
    __shared__ float xs[MAXHALO * CKP];
    __shared__ __align__(16) float ws[27 * CK * COT];
    __shared__ float red[COT * 2];

    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int co0 = blockIdx.y * COT;
    int t = blockIdx.x;
    const int w0 = (t % tiles_w) * TW; t /= tiles_w;
    const int h0 = (t % tiles_h) * TH; t /= tiles_h;
    const int d0 = t * TD;
    const int pd = kd / 2, ph = kh / 2, pw = kw / 2;
    const int HD = TD + kd - 1, HH = TH + kh - 1, HW = TW + kw - 1;
    const int taps = kd * kh * kw;
    const int tx = tid % TW, ty = (tid / TW) % TH, tz = tid / (TW * TH);

    float acc[COT];
#pragma unroll
    for (int i = 0; i < COT; ++i) acc[i] = 0.f;

    const T* xn = x + (size_t)n * D * H * W * x_ld;
    for (int c0 = 0; c0 < Cin; c0 += CK) {
        __syncthreads();
        // stage the haloed input tile for channels [c0, c0+CK)
        const int hv_total = HD * HH * HW;
        for (int i = tid; i < hv_total * CK; i += 256) {
            int ci = i % CK, hv = i / CK;
            int hx = hv % HW, hy = (hv / HW) % HH, hz = hv / (HW * HH);
            int gd = d0 + hz - pd, gh = h0 + hy - ph, gw = w0 + hx - pw;
            float v = 0.f;
            int c = c0 + ci;
            if (c < Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
                v = to_f<T>(xn[(((size_t)gd * H + gh) * W + gw) * x_ld + c]);
                if (in_ss) {
                    const float* p = in_ss + ((size_t)n * Cin + c) * 2;
                    v = fmaf(v, p[0], p[1]);
                }
            }
            xs[hv * CKP + ci] = v;
        }
        // stage weights w[tap][c0+ci][co0+co]
        for (int i = tid; i < taps * CK * COT; i += 256) {
            int co = i % COT, ci = (i / COT) % CK, tp = i / (COT * CK);
            float v = 0.f;
            if (c0 + ci < Cin && co0 + co < Cout) v = w[((size_t)tp * Cin + c0 + ci) * Cout + co0 + co];
            ws[i] = v;
        }
        __syncthreads();
        int tp = 0;
        for (int a = 0; a < kd; ++a)
            for (int b = 0; b < kh; ++b)
                for (int c = 0; c < kw; ++c, ++tp) {
                    const float* xp = xs + (((tz + a) * HH + ty + b) * HW + tx + c) * CKP;
#pragma unroll
                    for (int ci = 0; ci < CK; ++ci) {
                        const float xv = xp[ci];
                        const float4* wr = reinterpret_cast<const float4*>(ws + (tp * CK + ci) * COT);
#pragma unroll
                        for (int q = 0; q < COT / 4; ++q) {
                            float4 wv = wr[q];
                            acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
                            acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
                            acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
                        }
                    }
                }
    }

    // epilogue
    const int gd = d0 + tz, gh = h0 + ty, gw = w0 + tx;
    const bool inb = gd < D && gh < H && gw < W;
    const int nco = min(COT, Cout - co0);
#pragma unroll
    for (int i = 0; i < COT; ++i) {
        float v = acc[i] + ((bias && i < nco) ? bias[co0 + i] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        acc[i] = (inb && i < nco) ? round_as<T>(v) : 0.f;
    }
    if (inb) {
        T* yp = y + ((size_t)n * D * H * W + ((size_t)gd * H + gh) * W + gw) * y_ld + co0;
        constexpr int V = FullVec<T>::value;
        if (nco == COT && (y_ld % V) == 0 && (co0 % V) == 0 && aligned16(y)) {
#pragma unroll
            for (int q = 0; q < COT / V; ++q) {
                float tmp[V];
#pragma unroll
                for (int k = 0; k < V; ++k) tmp[k] = acc[q * V + k];
                Vec<T, V>::store(yp + q * V, tmp);
            }
        } else {
            for (int i = 0; i < nco; ++i) yp[i] = from_f<T>(acc[i]);
        }
    }
    if (sums) {
        if (tid < COT * 2) red[tid] = 0.f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < COT; ++i) {
            float s = acc[i], q = acc[i] * acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if ((tid & 31) == 0) {
                atomicAdd(&red[2 * i], s);
                atomicAdd(&red[2 * i + 1], q);
            }
        }
        __syncthreads();
        if (tid < nco * 2) atomicAdd(sums + ((size_t)n * Cout + co0) * 2 + tid, red[tid]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: dw[co][ci][tap] += sum_{n,vox} dz[n,vox,co] * x_hat[n,vox+tap,ci]
namespace {
constexpr int GD = 2, GH = 4, GW = 8;       // voxel tile per iteration (64 voxels)
constexpr int GCO = 32, GCI = 32;           // channel tile per block; thread tile 2 x 2
constexpr int GHALO = (GD + 2) * (GH + 2) * (GW + 2);
}  // namespace

template <typename T>
__global__ void __launch_bounds__(256)
conv3d_wgrad_direct_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ in_ss,
                           const T* __restrict__ dz, int64_t dz_ld, float* __restrict__ dw, int N, int D, int H, int W,
                           int Cin, int Cout, int kd, int kh, int kw, int ci_tiles) {
    __shared__ __align__(16) float xs[GHALO * GCI];
    __shared__ __align__(16) float ds[GD * GH * GW * GCO];

    const int tid = threadIdx.x;
    const int ci0 = (blockIdx.y % ci_tiles) * GCI, co0 = (blockIdx.y / ci_tiles) * GCO;
    const int pd = kd / 2, ph = kh / 2, pw = kw / 2;
    const int HD = GD + kd - 1, HH = GH + kh - 1, HW = GW + kw - 1;
    const int taps = kd * kh * kw;
    const int tw_n = (W + GW - 1) / GW, th_n = (H + GH - 1) / GH, td_n = (D + GD - 1) / GD;
    const int64_t tiles = (int64_t)N * td_n * th_n * tw_n;
    const int cip = (tid % 16) * 2, cop = (tid / 16) * 2;

    float acc[3][3][3][2][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[a][b][c][0][0] = acc[a][b][c][0][1] = acc[a][b][c][1][0] = acc[a][b][c][1][1] = 0.f;

    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int64_t t = tile;
        const int w0 = (int)(t % tw_n) * GW; t /= tw_n;
        const int h0 = (int)(t % th_n) * GH; t /= th_n;
        const int d0 = (int)(t % td_n) * GD; t /= td_n;
        const int n = (int)t;
        const T* xn = x + (size_t)n * D * H * W * x_ld;
        const T* dn = dz + (size_t)n * D * H * W * dz_ld;
        __syncthreads();
        const int hv_total = HD * HH * HW;
        for (int i = tid; i < hv_total * GCI; i += 256) {
            int ci = i % GCI, hv = i / GCI;
            int hx = hv % HW, hy = (hv / HW) % HH, hz = hv / (HW * HH);
            int gd = d0 + hz - pd, gh = h0 + hy - ph, gw = w0 + hx - pw;
            int c = ci0 + ci;
            float v = 0.f;
            if (c < Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
                v = to_f<T>(xn[(((size_t)gd * H + gh) * W + gw) * x_ld + c]);
                if (in_ss) {
                    const float* p = in_ss + ((size_t)n * Cin + c) * 2;
                    v = fmaf(v, p[0], p[1]);
                }
            }
            xs[i] = v;
        }
        for (int i = tid; i < GD * GH * GW * GCO; i += 256) {
            int co = i % GCO, v = i / GCO;
            int vx = v % GW, vy = (v / GW) % GH, vz = v / (GW * GH);
            int gd = d0 + vz, gh = h0 + vy, gw = w0 + vx;
            float g = 0.f;
            if (co0 + co < Cout && gd < D && gh < H && gw < W)
                g = to_f<T>(dn[(((size_t)gd * H + gh) * W + gw) * dz_ld + co0 + co]);
            ds[i] = g;
        }
        __syncthreads();
        for (int v = 0; v < GD * GH * GW; ++v) {
            const int vx = v % GW, vy = (v / GW) % GH, vz = v / (GW * GH);
            const float2 g = *reinterpret_cast<const float2*>(ds + v * GCO + cop);
            // fully unrolled over the 3x3x3 slot grid so the accumulators stay in registers; absent taps
            // (1x3x3 kernels) are skipped by warp-uniform guards
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (a >= kd) continue;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    if (b >= kh) continue;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (c >= kw) continue;
                        const float2 xv = *reinterpret_cast<const float2*>(
                            xs + (((vz + a) * HH + vy + b) * HW + vx + c) * GCI + cip);
                        acc[a][b][c][0][0] = fmaf(g.x, xv.x, acc[a][b][c][0][0]);
                        acc[a][b][c][0][1] = fmaf(g.x, xv.y, acc[a][b][c][0][1]);
                        acc[a][b][c][1][0] = fmaf(g.y, xv.x, acc[a][b][c][1][0]);
                        acc[a][b][c][1][1] = fmaf(g.y, xv.y, acc[a][b][c][1][1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (a >= kd || b >= kh || c >= kw) continue;
                const int tp = (a * kh + b) * kw + c;
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        int co = co0 + cop + i, ci = ci0 + cip + j;
                        if (co < Cout && ci < Cin) atomicAdd(dw + ((size_t)co * Cin + ci) * taps + tp, acc[a][b][c][i][j]);
                    }
            }
}

// ---------------------------------------------------------------------------------------------------------------
// weight (+ bias) gradient of the network's FIRST conv (Cin <= 4): 27*Cin*Cout outputs reduced over all voxels.
// HBM-bound: one pass over dz (Cout channels) and x (Cin channels).  Block = 4x8x8 voxel tile, persistent over tiles;
// thread (co = tid % 32, slot group = tid / 32) owns the (tap, ci) slots {tid/32 + 8k}; dz is staged as fp32 in
// shared memory (lanes read consecutive co: conflict-free), x_hat values are warp-broadcast reads of the haloed tile.
namespace {
constexpr int SC_MAXCIN = 4;
constexpr int SC_MAXQ = (27 * SC_MAXCIN + 7) / 8;   // slots per thread
constexpr int SC_VOX = TD * TH * TW;                 // 256
}  // namespace

template <typename T>
__global__ void __launch_bounds__(256)
conv3d_wgrad_smallcin_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ in_ss,
                             const T* __restrict__ dz, int64_t dz_ld, float* __restrict__ dw, float* __restrict__ db,
                             int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw) {
    __shared__ float xs[MAXHALO * SC_MAXCIN];
    __shared__ float ds[SC_VOX * COT];
    const int tid = threadIdx.x;
    const int co_l = tid & 31, grp = tid >> 5;
    const int co0 = blockIdx.y * COT;
    const int pd = kd / 2, ph = kh / 2, pw = kw / 2;
    const int HH = TH + kh - 1, HW = TW + kw - 1, HD = TD + kd - 1;
    const int taps = kd * kh * kw;
    const int nslots = taps * Cin;
    const int tw_n = (W + TW - 1) / TW, th_n = (H + TH - 1) / TH, td_n = (D + TD - 1) / TD;
    const int64_t tiles = (int64_t)N * td_n * th_n * tw_n;

    float acc[SC_MAXQ];
    int off[SC_MAXQ];                                  // haloed-tile element offset of slot k relative to the voxel
#pragma unroll
    for (int k = 0; k < SC_MAXQ; ++k) {
        acc[k] = 0.f;
        const int q = grp + 8 * k;
        int o = 0;
        if (q < nslots) {
            const int tp = q / Cin, ci = q % Cin;
            const int a = tp / (kh * kw), b = (tp / kw) % kh, c = tp % kw;
            o = ((a * HH + b) * HW + c) * Cin + ci;
        }
        off[k] = o;
    }
    float dbacc = 0.f;

    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int64_t t = tile;
        const int w0 = (int)(t % tw_n) * TW; t /= tw_n;
        const int h0 = (int)(t % th_n) * TH; t /= th_n;
        const int d0 = (int)(t % td_n) * TD; t /= td_n;
        const int n = (int)t;
        const T* xn = x + (size_t)n * D * H * W * x_ld;
        const T* dn = dz + (size_t)n * D * H * W * dz_ld;
        __syncthreads();
        for (int i = tid; i < HD * HH * HW * Cin; i += 256) {
            const int ci = i % Cin, hv = i / Cin;
            const int hx = hv % HW, hy = (hv / HW) % HH, hz = hv / (HW * HH);
            const int gd = d0 + hz - pd, gh = h0 + hy - ph, gw = w0 + hx - pw;
            float v = 0.f;
            if (gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
                v = to_f<T>(xn[(((size_t)gd * H + gh) * W + gw) * x_ld + ci]);
                if (in_ss) {
                    const float* p = in_ss + ((size_t)n * Cin + ci) * 2;
                    v = fmaf(v, p[0], p[1]);
                }
            }
            xs[i] = v;
        }
        for (int i = tid; i < SC_VOX * COT; i += 256) {
            const int co = i % COT, v = i / COT;
            const int vx = v % TW, vy = (v / TW) % TH, vz = v / (TW * TH);
            const int gd = d0 + vz, gh = h0 + vy, gw = w0 + vx;
            float g = 0.f;
            if (co0 + co < Cout && gd < D && gh < H && gw < W)
                g = to_f<T>(dn[(((size_t)gd * H + gh) * W + gw) * dz_ld + co0 + co]);
            ds[i] = g;
        }
        __syncthreads();
        for (int v = 0; v < SC_VOX; ++v) {
            const int vx = v % TW, vy = (v / TW) % TH, vz = v / (TW * TH);
            const float g = ds[v * COT + co_l];
            const float* xb = xs + ((vz * HH + vy) * HW + vx) * Cin;
            if (grp == 0) dbacc += g;
#pragma unroll
            for (int k = 0; k < SC_MAXQ; ++k)
                if (grp + 8 * k < nslots) acc[k] = fmaf(g, xb[off[k]], acc[k]);
        }
    }
    const int co = co0 + co_l;
    if (co < Cout) {
#pragma unroll
        for (int k = 0; k < SC_MAXQ; ++k) {
            const int q = grp + 8 * k;
            if (q < nslots) atomicAdd(dw + ((size_t)co * Cin + q % Cin) * taps + q / Cin, acc[k]);
        }
        if (db && grp == 0) atomicAdd(db + co, dbacc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight re-packing: torch (Cout,Cin,taps) fp32 -> operand layouts
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, float* __restrict__ wf,
                                    float* __restrict__ wd) {
    int64_t total = (int64_t)Cout * Cin * taps;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int tp = (int)(i % taps);
        int ci = (int)((i / taps) % Cin);
        int co = (int)(i / ((int64_t)taps * Cin));
        float v = w[i];
        if (wf) wf[((size_t)tp * Cin + ci) * Cout + co] = v;
        if (wd) wd[((size_t)(taps - 1 - tp) * Cout + co) * Cin + ci] = v;
    }
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_pack_conv_weights(const float* w, int Cout, int Cin, int kd, int kh, int kw, float* w_fwd_f32,
                             float* w_dgrad_f32, void* stream) {
    B2_CHECK_ARG(w && Cout > 0 && Cin > 0 && kd > 0 && kh > 0 && kw > 0, "pack_conv_weights: bad arguments");
    int taps = kd * kh * kw;
    int64_t total = (int64_t)Cout * Cin * taps;
    int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, taps, w_fwd_f32, w_dgrad_f32);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_conv3d_direct(const void* x, int64_t x_ld, const float* in_scale_shift, const float* w, const float* bias,
                         void* y, int64_t y_ld, float* sums, int dtype, int N, int D, int H, int W, int Cin, int Cout,
                         int kd, int kh, int kw, int relu, void* stream) {
    B2_CHECK_ARG(x && w && y && N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3d_direct: bad arguments");
    B2_CHECK_ARG((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3), "conv3d_direct: kernel dims must be 1 or 3");
    B2_CHECK_ARG(x_ld >= Cin && y_ld >= Cout, "conv3d_direct: pitch smaller than channel count");
    int tw = (W + TW - 1) / TW, th = (H + TH - 1) / TH, td = (D + TD - 1) / TD;
    dim3 grid((unsigned)(tw * th * td), (unsigned)((Cout + COT - 1) / COT), (unsigned)N);
    B2_DISPATCH_DTYPE(dtype, T, {
        if (Cin < 4)
            conv3d_direct_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, in_scale_shift, w, bias, (T*)y, y_ld,
                                                                                sums, D, H, W, Cin, Cout, kd, kh, kw, relu, tw, th);
        else
            conv3d_direct_kernel<T, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, in_scale_shift, w, bias, (T*)y, y_ld,
                                                                                sums, D, H, W, Cin, Cout, kd, kh, kw, relu, tw, th);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_conv3d_wgrad_smallcin(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld,
                                 int dtype, float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh,
                                 int kw, void* stream) {
    B2_CHECK_ARG(x && dz && dw && N > 0 && D > 0 && H > 0 && W > 0 && Cout > 0, "conv3d_wgrad_smallcin: bad arguments");
    B2_CHECK_ARG(Cin >= 1 && Cin <= SC_MAXCIN, "conv3d_wgrad_smallcin: Cin must be in [1, %d], got %d", SC_MAXCIN, Cin);
    B2_CHECK_ARG((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3), "conv3d_wgrad_smallcin: kernel dims must be 1 or 3");
    int co_tiles = (Cout + COT - 1) / COT;
    int64_t tiles = (int64_t)N * ((D + TD - 1) / TD) * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
    int64_t per = (int64_t)sm_count() * 4 / co_tiles + 1;
    if (per > tiles) per = tiles;
    dim3 grid((unsigned)per, (unsigned)co_tiles, 1);
    B2_DISPATCH_DTYPE(dtype, T, {
        conv3d_wgrad_smallcin_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, in_scale_shift, (const T*)dz, dz_ld,
                                                                                 dw, db, N, D, H, W, Cin, Cout, kd, kh, kw);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_conv3d_wgrad_direct(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld,
                               int dtype, float* dw, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw,
                               void* stream) {
    B2_CHECK_ARG(x && dz && dw && N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3d_wgrad_direct: bad arguments");
    B2_CHECK_ARG((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3), "conv3d_wgrad_direct: kernel dims must be 1 or 3");
    int ci_tiles = (Cin + GCI - 1) / GCI, co_tiles = (Cout + GCO - 1) / GCO;
    int64_t tiles = (int64_t)N * ((D + GD - 1) / GD) * ((H + GH - 1) / GH) * ((W + GW - 1) / GW);
    int64_t per = (int64_t)sm_count() * 2 / ((int64_t)ci_tiles * co_tiles) + 1;
    if (per > tiles) per = tiles;
    dim3 grid((unsigned)per, (unsigned)(ci_tiles * co_tiles), 1);
    B2_DISPATCH_DTYPE(dtype, T, {
        conv3d_wgrad_direct_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, x_ld, in_scale_shift, (const T*)dz, dz_ld, dw,
                                                                               N, D, H, W, Cin, Cout, kd, kh, kw, ci_tiles);
    })
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
