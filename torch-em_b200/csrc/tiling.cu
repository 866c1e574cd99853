// Tiled-inference I/O kernels: the per-block host work of torch-em's predict_with_halo (util/prediction.py:98-142, 249-309)
// done on a volume that already lives in HBM.
//
//   gather_blocks      _load_block: the haloed bounding box [offset - halo, offset + block_shape + halo) clipped to the volume and
//                      completed by np.pad(mode="reflect") of the CLIPPED data (mirror without repeating the edge voxel; a
//                      triangular wave of period 2*(len-1) around the clipped range), any raw dtype -> fp32, plus the block's
//                      (sum, sum of squares) for ``standardize``
//   standardize_blocks transform/raw.py:40-65: x = (x - mean) / (std + eps), population statistics of the whole haloed block
//   scatter_blocks     the inner crop [halo, halo + block.shape) of the prediction written to output[:, block.begin:block.end],
//                      zeroed outside the mask (prediction.py:287-309)
// HBM-bound: gather reads sizeof(raw) and writes 4 bytes per haloed voxel, standardize reads + writes 4, scatter reads and writes
// 4 bytes per kept output element.  Several blocks per launch (one grid.y slice per block).
#include "common.cuh"

namespace b200em {

struct BlockList {
    int n;
    int begin[B200EM_MAX_TILE_BLOCKS][3];   // first voxel of the haloed box (may be negative) resp. of the inner block
    int shape[B200EM_MAX_TILE_BLOCKS][3];   // scatter: extent of the inner block (truncated last blocks)
};

__device__ __forceinline__ int reflect_index(int i, int start, int stop, int n) {
    // index i of the box [start, stop) into an axis of length n: clip the box to [lo, hi), mirror around the clipped range
    const int lo = start > 0 ? start : 0, hi = stop < n ? stop : n;
    const int len = hi - lo;
    if (len <= 1) return lo;
    const int period = 2 * (len - 1);
    int r = (i - lo) % period;
    if (r < 0) r += period;
    return lo + (r >= len ? period - r : r);
}

template <typename TI>
__device__ __forceinline__ float raw_to_f(TI v) { return (float)v; }
template <>
__device__ __forceinline__ float raw_to_f<__half>(__half v) { return __half2float(v); }

// vol (C, D, H, W) of TI -> out (nblocks, C, bd, bh, bw) fp32; stats[nblocks][2] += (sum, sum of squares) in double.
// One CTA = TILE_ROWS rows (c, d, h) of one block: the row's source (d, h) is resolved once per row, the threads run along w
// (coalesced loads and stores; only the w reflection is per element).  grid = (row groups, nblocks).
constexpr int TILE_ROWS = 8;

template <typename TI>
__global__ void __launch_bounds__(256)
gather_blocks_kernel(const TI* __restrict__ vol, int C, int D, int H, int W, BlockList bl, int bd, int bh, int bw,
                     float* __restrict__ out, double* __restrict__ stats) {
    __shared__ double sh[8][2];
    const int b = blockIdx.y;
    const int rows = C * bd * bh;
    const int d0 = bl.begin[b][0], h0 = bl.begin[b][1], w0 = bl.begin[b][2];
    float* o = out + (size_t)b * rows * bw;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 lanes along w, 8 rows per CTA
    const int row = blockIdx.x * TILE_ROWS + ty;
    float s = 0.f, q = 0.f;
    if (row < rows) {
        const int h = row % bh, d = (row / bh) % bd, c = row / (bh * bd);
        const int gd = reflect_index(d0 + d, d0, d0 + bd, D), gh = reflect_index(h0 + h, h0, h0 + bh, H);
        const TI* __restrict__ src = vol + (((size_t)c * D + gd) * H + gh) * W;
        float* __restrict__ dst = o + (size_t)row * bw;
        // interior of the row (no reflection along w) is a straight copy; the clipped range is [wlo, whi)
        const int wlo = w0 > 0 ? w0 : 0, whi = w0 + bw < W ? w0 + bw : W;
        for (int w = tx; w < bw; w += 32) {
            const int gw0 = w0 + w;
            const int gw = (gw0 >= wlo && gw0 < whi) ? gw0 : reflect_index(gw0, w0, w0 + bw, W);
            const float v = raw_to_f<TI>(src[gw]);
            dst[w] = v;
            s += v;
            q = fmaf(v, v, q);
        }
    }
    if (stats) {
        double ds = s, dq = q;                                   // <= bw / 32 fp32 terms per thread, double from here on
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            ds += __shfl_xor_sync(0xffffffffu, ds, off);
            dq += __shfl_xor_sync(0xffffffffu, dq, off);
        }
        if (tx == 0) { sh[ty][0] = ds; sh[ty][1] = dq; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double v = 0.0;
            for (int k = 0; k < 8; ++k) v += sh[k][threadIdx.x];
            atomicAdd(stats + 2 * b + threadIdx.x, v);
        }
    }
}

__global__ void __launch_bounds__(256)
standardize_blocks_kernel(float* __restrict__ x, int64_t per_block, const double* __restrict__ stats, float eps) {
    const int b = blockIdx.y;
    const double mean = stats[2 * b] / (double)per_block;
    double var = stats[2 * b + 1] / (double)per_block - mean * mean;
    if (var < 0.0) var = 0.0;
    const float m = (float)mean, inv = 1.f / ((float)sqrt(var) + eps);
    float* p = x + (size_t)b * per_block;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_block; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = (p[i] - m) * inv;
}

// pred (nblocks, Cp, bd, bh, bw) fp32 -> out (Co, D, H, W): channels [c0, c0 + nc) of the inner crop of every block.
// One CTA = TILE_ROWS rows (c, d, h) of one inner block, threads along w.  grid = (row groups of the largest block, nblocks).
__global__ void __launch_bounds__(256)
scatter_blocks_kernel(const float* __restrict__ pred, int Cp, int bd, int bh, int bw, int hd, int hh, int hw, BlockList bl,
                      float* __restrict__ out, int D, int H, int W, int c0, int nc, const unsigned char* __restrict__ mask) {
    const int b = blockIdx.y;
    const int sd = bl.shape[b][0], sh_ = bl.shape[b][1], sw = bl.shape[b][2];
    const int od = bl.begin[b][0], oh = bl.begin[b][1], ow = bl.begin[b][2];
    const int rows = nc * sd * sh_;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int row = blockIdx.x * TILE_ROWS + ty;
    if (row >= rows) return;
    const int h = row % sh_, d = (row / sh_) % sd, c = row / (sh_ * sd);
    const float* __restrict__ src = pred + ((((size_t)b * Cp + (c0 + c)) * bd + (hd + d)) * bh + (hh + h)) * bw + hw;
    const size_t dst_row = ((size_t)(od + d) * H + (oh + h)) * W + ow;
    float* __restrict__ dst = out + (size_t)c * D * H * W + dst_row;
    const unsigned char* __restrict__ m = mask ? mask + dst_row : nullptr;
    for (int w = tx; w < sw; w += 32) {
        float v = src[w];
        if (m && !m[w]) v = 0.f;
        dst[w] = v;
    }
}

static inline unsigned tile_grid(int64_t total) {
    int64_t b = (total + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

static int fill_blocks(BlockList& bl, const int* begins, const int* shapes, int n) {
    if (n < 1 || n > B200EM_MAX_TILE_BLOCKS) {
        set_error("tiling: between 1 and %d blocks per launch, got %d", B200EM_MAX_TILE_BLOCKS, n);
        return 1;
    }
    bl.n = n;
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            bl.begin[i][a] = begins[3 * i + a];
            bl.shape[i][a] = shapes ? shapes[3 * i + a] : 0;
        }
    return 0;
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_gather_blocks(const void* vol, int raw_dtype, int C, int D, int H, int W, const int* block_begins, int nblocks, int bd,
                         int bh, int bw, float* out, double* stats, void* stream) {
    B2_CHECK_ARG(vol && out && block_begins && C > 0 && D > 0 && H > 0 && W > 0 && bd > 0 && bh > 0 && bw > 0, "gather_blocks: bad arguments");
    BlockList bl;
    if (fill_blocks(bl, block_begins, nullptr, nblocks)) return 1;
    B2_CHECK_ARG((int64_t)C * bd * bh < (1LL << 31), "gather_blocks: block too large");
    dim3 grid((unsigned)(((int64_t)C * bd * bh + TILE_ROWS - 1) / TILE_ROWS), (unsigned)nblocks);
#define B2_GATHER(TI) gather_blocks_kernel<TI><<<grid, 256, 0, (cudaStream_t)stream>>>((const TI*)vol, C, D, H, W, bl, bd, bh, bw, out, stats)
    switch (raw_dtype) {
        case B200EM_RAW_U8: B2_GATHER(uint8_t); break;
        case B200EM_RAW_I8: B2_GATHER(int8_t); break;
        case B200EM_RAW_U16: B2_GATHER(uint16_t); break;
        case B200EM_RAW_I16: B2_GATHER(int16_t); break;
        case B200EM_RAW_I32: B2_GATHER(int32_t); break;
        case B200EM_RAW_U32: B2_GATHER(uint32_t); break;
        case B200EM_RAW_F16: B2_GATHER(__half); break;
        case B200EM_RAW_F32: B2_GATHER(float); break;
        case B200EM_RAW_F64: B2_GATHER(double); break;
        default: set_error("gather_blocks: unknown raw dtype code %d", raw_dtype); return 1;
    }
#undef B2_GATHER
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_standardize_blocks(float* x, int nblocks, int64_t per_block, const double* stats, float eps, void* stream) {
    B2_CHECK_ARG(x && stats && nblocks > 0 && nblocks <= 65535 && per_block > 0, "standardize_blocks: bad arguments");
    dim3 grid(tile_grid(per_block), (unsigned)nblocks);
    standardize_blocks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, per_block, stats, eps);
    B2_LAUNCH_CHECK();
    return 0;
}

int b200em_scatter_blocks(const float* pred, int Cp, int bd, int bh, int bw, int hd, int hh, int hw, const int* block_begins,
                          const int* block_shapes, int nblocks, float* out, int D, int H, int W, int c0, int nc,
                          const unsigned char* mask, void* stream) {
    B2_CHECK_ARG(pred && out && block_begins && block_shapes && Cp > 0 && nc > 0 && c0 >= 0 && c0 + nc <= Cp, "scatter_blocks: bad arguments");
    BlockList bl;
    if (fill_blocks(bl, block_begins, block_shapes, nblocks)) return 1;
    int64_t big = 0;
    for (int i = 0; i < nblocks; ++i) {
        const int* s = block_shapes + 3 * i;
        const int* b = block_begins + 3 * i;
        B2_CHECK_ARG(s[0] > 0 && s[1] > 0 && s[2] > 0 && hd + s[0] <= bd && hh + s[1] <= bh && hw + s[2] <= bw,
                     "scatter_blocks: inner block %d does not fit into the prediction", i);
        B2_CHECK_ARG(b[0] >= 0 && b[1] >= 0 && b[2] >= 0 && b[0] + s[0] <= D && b[1] + s[1] <= H && b[2] + s[2] <= W,
                     "scatter_blocks: block %d leaves the output volume", i);
        const int64_t t = (int64_t)s[0] * s[1] * nc;               // rows (c, d, h) of the inner block
        if (t > big) big = t;
    }
    dim3 grid((unsigned)((big + TILE_ROWS - 1) / TILE_ROWS), (unsigned)nblocks);
    scatter_blocks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, Cp, bd, bh, bw, hd, hh, hw, bl, out, D, H, W, c0, nc, mask);
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
