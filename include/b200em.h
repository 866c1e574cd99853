/*
 * b200em.h -- C ABI of libb200em.so: the B200 (sm_100a) kernels behind torch-em's 3D U-Net train step.
 *
 * The reference (constantinpape/torch-em) is pure Python over torch.nn; it has no FFI of its own.  Its
 * "plugin API" for this path is nn.Module duck-typing.  This header is the boundary a
 * maintainer would bind instead of the torch.nn calls listed beside each entry point: plain pointers and sizes,
 * a cudaStream_t passed as void*, int status returns (0 = ok, message via b200em_last_error()).  No torch types.
 * No entry point allocates device memory or synchronises the stream; workspaces are caller-provided.
 *
 * Conventions
 *   - Activations are channels-last "NDHWC": element (n,d,h,w,c) at ((n*D+d)*H+h)*W+w)*ld + c, ld >= C being
 *     the per-voxel pitch in elements (so a tensor can be a channel slice of a wider concat buffer).
 *   - dtype codes: B200EM_F32 = 0, B200EM_BF16 = 1.  Statistics / gradients of parameters are always fp32.
 *   - "sums" buffers are ACCUMULATED into with atomics: the caller zeroes them.
 *   - Network input / prediction / targets / labels are NCDHW like the reference's tensors.
 */
#ifndef B200EM_H_
#define B200EM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200EM_ABI_VERSION 1
#define B200EM_F32 0
#define B200EM_BF16 1

#define B200EM_ACT_NONE 0
#define B200EM_ACT_SIGMOID 1
#define B200EM_ACT_RELU 2
#define B200EM_ACT_TANH 3

/* ---- library ------------------------------------------------------------------------------------------ */
int b200em_abi_version(void);
const char* b200em_last_error(void);
/* SM count and compute capability of the current device; umma_ok = 1 iff tcgen05 kernels can run (cc 10.x). */
int b200em_device_info(int* sm_count, int* cc_major, int* cc_minor, int* umma_ok);
/* Number of kernel launches issued through this library since the last reset (bench.py "gpu_launches"). */
int64_t b200em_launch_count(void);
void b200em_reset_launch_count(void);

/* Registers a caller-owned scratch buffer for the CURRENT device (NULL, 0 unregisters).  It must be zero-filled when it is
 * registered and must only be used by one stream at a time; kernels that use it restore it to all-zero before they finish.
 * With a workspace the tcgen05 conv kernels split the reduction (input-channel chunks) of layers that have fewer output tiles
 * than SMs over several CTAs (partial sums added into the workspace, the last CTA of a tile runs the epilogue); without one
 * they run unsplit.  Recommended size: 64 MiB. */
int b200em_set_workspace(void* workspace, int64_t bytes);

/* ---- layout ------------------------------------------------------------------------------------------- */
/* NCDHW fp32 network input -> NDHWC activations (replaces nothing in the reference: it keeps NCDHW). */
int b200em_ncdhw_to_ndhwc(const float* x, void* y, int y_dtype, int64_t y_ld, int N, int C, int64_t S, void* stream);

/* ---- convolution weights -------------------------------------------------------------------------------
 * Re-pack an nn.Conv3d weight (Cout,Cin,kd,kh,kw) fp32 (unet.py:431,435,453,638) into the direct-kernel operand
 * layouts.  Any output may be NULL.  taps = kd*kh*kw, tap index t = (a*kh+b)*kw+c, flipped tap = taps-1-t.
 *   w_fwd_f32   [taps][Cin][Cout]            forward
 *   w_dgrad_f32 [taps][Cout][Cin]  (flipped) data-gradient = forward conv of dY with these ("Cin" <-> "Cout")
 */
int b200em_pack_conv_weights(const float* w, int Cout, int Cin, int kd, int kh, int kw,
                             float* w_fwd_f32, float* w_dgrad_f32, void* stream);

/* ---- convolution: [Norm ->] nn.Conv3d(k, padding=k//2) [-> ReLU]  (unet.py:429-438) and its backward ------- */
/* Direct (CUDA-core, fp32 accumulate) path: any channel counts, f32 or bf16 activations.  w = w_fwd_f32 layout.
 *   x_hat = in_scale_shift ? scale[n,c]*x + shift[n,c] : x   on in-bounds voxels, 0 in the padding
 *   y     = act(conv(x_hat, w) + bias)                       bias may be NULL; relu in {0,1}
 *   sums (nullable) [N][Cout][2] += (sum y, sum y^2) of the STORED (rounded) outputs: the statistics the next
 *   InstanceNorm/GroupNorm needs (unet.py:397,402).
 * The data-gradient is the same call with w = w_dgrad_f32 and Cin/Cout swapped. */
int b200em_conv3d_direct(const void* x, int64_t x_ld, const float* in_scale_shift, const float* w, const float* bias,
                         void* y, int64_t y_ld, float* sums, int dtype, int N, int D, int H, int W, int Cin, int Cout,
                         int kd, int kh, int kw, int relu, void* stream);
/* dW (Cout,Cin,kd,kh,kw) fp32 += sum_{n,vox} dz[n,vox,co] * x_hat[n,vox+tap,ci]  (torch layout, accumulated). */
int b200em_conv3d_wgrad_direct(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld,
                               int dtype, float* dw, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh,
                               int kw, void* stream);

/* Weight + bias gradient of the network's first conv (1 <= Cin <= 4; HBM-bound: one pass over dz and x).
 * dw (Cout,Cin,kd,kh,kw) += ..., db (nullable) (Cout) += sum dz. */
int b200em_conv3d_wgrad_smallcin(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld,
                                 int dtype, float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh,
                                 int kw, void* stream);

/* tcgen05 implicit-GEMM path: bf16 activations and weights, fp32 accumulation in TMEM (csrc/conv_umma.cu).
 * Same fused prologue / epilogue contract as b200em_conv3d_direct.  Takes Cin % 16 == 0, Cout % 16 == 0
 * (Cout <= 256 or Cout % 128 == 0); b200em_conv3d_umma_supported() says so, and the other entry points return
 * 2 ("unsupported shape") otherwise so that the caller can take the direct kernel.
 * Weights are pre-packed once per optimizer step by b200em_conv3d_umma_pack into Cout*Cin*taps bf16:
 *   dgrad = 0: forward operand;  dgrad = 1: transposed, tap-flipped operand -- the data gradient is then
 *   b200em_conv3d_umma(dz, ..., Cin := conv's Cout, Cout := conv's Cin). */
int b200em_conv3d_umma_supported(int Cin, int Cout, int kd, int kh, int kw);
int b200em_conv3d_umma_pack(const float* w, int Cout, int Cin, int kd, int kh, int kw, int dgrad, void* packed,
                            void* stream);
/* dot_x (nullable, NDHWC bf16 with pitch dot_ld, Cout channels): when given, sums += (sum y, sum y*dot_x) instead of
 * (sum y, sum y^2) -- the two reductions of the InstanceNorm/GroupNorm backward, fused into the data-gradient conv. */
int b200em_conv3d_umma(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                       void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W, int Cin,
                       int Cout, int kd, int kh, int kw, int relu, void* stream);

/* TF32 variant of the conv entry point above: fp32 activations and fp32 packed weights (b200em_pack_batch with
 * B200EM_PACK_PLAIN_TF32), fed to tcgen05.mma.kind::tf32, which reads the upper 19 bits of every operand word -- the arithmetic
 * torch / cuDNN use for fp32 convolutions by default (torch.backends.cudnn.allow_tf32; the reference's mixed_precision=False
 * path, default_trainer.py:132-142).  Same contract, Cin % 16 == 0.  (The fp32 weight gradient runs the bf16 weight-gradient
 * kernels three times on split operands, see b200em_split_bf16.) */
int b200em_conv3d_umma_tf32_supported(int Cin, int Cout, int kd, int kh, int kw);
int b200em_conv3d_umma_tf32(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                            void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W, int Cin,
                            int Cout, int kd, int kh, int kw, int relu, void* stream);

/* "h16" variant -- the DEFAULT tensor-core path of fp32 activations when TF32 convolutions are allowed.  IEEE fp16 has the same
 * 11-bit significand as TF32 (and is rounded to nearest where kind::tf32 truncates), so feeding fp16 COPIES of the fp32
 * operands to tcgen05.mma.kind::f16 with fp32 accumulation is TF32-class arithmetic at the bf16 rate (2x kind::tf32); fp16's
 * narrow exponent is handled by an exact power-of-two scale per tensor (b200em_absmax_f32 / b200em_cvt_f16) that the kernel
 * undoes on the fp32 accumulator.  Replaces nn.Conv3d forward / autograd dgrad / wgrad of the reference's mixed_precision=False
 * path (unet.py:429-438, default_trainer.py:132-142).
 *   x_f16: NDHWC fp16 copy of the (normalised) input, pitch x_ld; x_absmax: the DEVICE float its scale was derived from, or NULL
 *   (unscaled);  w_packed: b200em_pack_batch image with B200EM_PACK_PLAIN_F16;  y / dot_x: fp32.  Otherwise as b200em_conv3d_umma
 *   (same supported shapes: b200em_conv3d_umma_supported). */
int b200em_conv3d_umma_h16(const void* x_f16, int64_t x_ld, const float* x_absmax, const void* w_packed, const float* bias, float* y,
                           int64_t y_ld, float* sums, const float* dot_x, int64_t dot_ld, int N, int D, int H, int W, int Cin, int Cout,
                           int kd, int kh, int kw, int relu, void* stream);
/* max |x| over an fp32 tensor of `rows` voxels x C channels with pitch x_ld, as an atomic max into *absmax (device, zeroed by
 * the caller; non-negative floats order like their bit patterns).  colsum (nullable, C floats) += the per-channel sums of the
 * same pass: the bias gradient sum(dz) from the fp32 values, as autograd computes it. */
int b200em_absmax_f32(const float* x, int64_t x_ld, int64_t rows, int C, float* absmax, float* colsum, void* stream);
/* out (N,S,C) fp16, contiguous = fp16(2^k * x_hat), x_hat = scale*x + shift when in_scale_shift is given (the fused norm apply),
 * k derived on the device from *absmax (NULL: k = 0; see common.cuh h16_shift).  colsum (nullable, C floats, C <= 2048) += the
 * per-channel sums of x in the same pass (the bias gradient sum(dz), from the fp32 values). */
int b200em_cvt_f16(const float* x, int64_t x_ld, const float* in_scale_shift, const float* absmax, void* out, float* colsum, int N,
                   int64_t S, int C, void* stream);

/* "depth-stacked" tcgen05 variant for 3 x kh x kw filters with few output channels (Cout <= 80) whose packed filter
 * fits in shared memory: the three depth taps share one operand fetch (N = 3*Cout) and land in the accumulators of
 * three consecutive output slices (a ring of TMEM column blocks); every input slice is staged once per column of
 * output slices (csrc/conv_umma_ds.cu).  Replaces nn.Conv3d forward / autograd dgrad of the shallow wide levels
 * (unet.py:429-438).  Same contract and arguments as b200em_conv3d_umma (incl. dot_x); its own packed weight layout
 * (b200em_conv3d_umma_ds_pack, Cout*Cin*taps bf16). */
int b200em_conv3d_umma_ds_supported(int Cin, int Cout, int kd, int kh, int kw);
int b200em_conv3d_umma_ds_pack(const float* w, int Cout, int Cin, int kd, int kh, int kw, int dgrad, void* packed,
                               void* stream);
int b200em_conv3d_umma_ds(const void* x, int64_t x_ld, const float* in_scale_shift, const void* w_packed, const float* bias,
                          void* y, int64_t y_ld, float* sums, const void* dot_x, int64_t dot_ld, int N, int D, int H, int W,
                          int Cin, int Cout, int kd, int kh, int kw, int relu, void* stream);

/* h16 variant of the depth-stacked kernel (forward convs of fp32 activations whose input is normalised, i.e. needs no range
 * scale): x_f16 from b200em_cvt_f16 with the norm apply in it, w_packed with B200EM_PACK_DEPTH_STACKED_F16, fp32 output and
 * statistics.  No dot_x: the fp32 data gradients take b200em_conv3d_umma_h16. */
int b200em_conv3d_umma_ds_h16(const void* x_f16, int64_t x_ld, const void* w_packed, const float* bias, float* y, int64_t y_ld, float* sums,
                              int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int relu, void* stream);

/* Batched packing: the bf16 operand images of EVERY conv weight of a model in one launch (csrc/pack.cu).  The fp32
 * nn.Parameter (unet.py:431,435,453) stays the master; the images are rebuilt at the start of every forward / backward pass,
 * so no in-place parameter update can leave a stale operand behind.  The caller fills w, packed (Cout*Cin*taps bf16, 16-byte
 * aligned), the filter shape, dgrad and layout of each job; b200em_pack_batch_prepare (host) validates them and fills the
 * derived fields; the table is then copied to the device once and b200em_pack_batch launches over it. */
#define B200EM_PACK_PLAIN 0          /* operand of b200em_conv3d_umma */
#define B200EM_PACK_DEPTH_STACKED 1  /* operand of b200em_conv3d_umma_ds */
#define B200EM_PACK_PLAIN_TF32 2     /* fp32 operand of b200em_conv3d_umma_tf32 (packed: Cout*Cin*taps fp32) */
#define B200EM_PACK_PLAIN_F16 3      /* IEEE fp16 operand of b200em_conv3d_umma_h16 (the plain layout) */
#define B200EM_PACK_DEPTH_STACKED_F16 4   /* IEEE fp16 operand of b200em_conv3d_umma_ds_h16 (the depth-stacked layout) */
typedef struct b200em_pack_job {
    const float* w;    /* torch (Cout, Cin, kd, kh, kw) fp32, device */
    void* packed;      /* bf16 operand image, device */
    int32_t Cout, Cin, kd, kh, kw;
    int32_t dgrad;     /* 1: transposed, tap-flipped operand (data gradient) */
    int32_t layout;    /* B200EM_PACK_* */
    int32_t CC, NPb, block_begin;   /* filled by b200em_pack_batch_prepare */
    int32_t reserved[2];
} b200em_pack_job;     /* 64 bytes */
int b200em_pack_batch_prepare(b200em_pack_job* jobs, int njobs, int* total_blocks);
int b200em_pack_batch(const b200em_pack_job* jobs_device, int njobs, int total_blocks, void* stream);

/* First convolution of the network (Cin = 1, 3x3x3, Cout in {16,32,48,64}; unet.py:412-438 first block): the im2col rows
 * (K = 27 padded to 32) are built on the fly in shared memory and fed to tcgen05.mma; same fused prologue (norm apply on
 * the 1-channel input, x: (N,D,H,W,1) bf16) and epilogue (bias, ReLU, statistics) as b200em_conv3d_umma.  w / dw are the
 * torch-layout fp32 parameter (Cout,1,3,3,3) and its gradient (accumulated); db (nullable) += sum dz. */
int b200em_conv3d_first_supported(int Cin, int Cout, int kd, int kh, int kw);
int b200em_conv3d_first(const void* x, const float* in_scale_shift, const float* w, const float* bias, void* y, int64_t y_ld,
                        float* sums, int N, int D, int H, int W, int Cout, int relu, void* stream);
/* fp32 variant (the h16 path): x (N,D,H,W,1) fp32, y fp32; the im2col image and the filter are rounded to IEEE fp16 (TF32's
 * significand) in shared memory. */
int b200em_conv3d_first_f32(const float* x, const float* in_scale_shift, const float* w, const float* bias, float* y, int64_t y_ld,
                            float* sums, int N, int D, int H, int W, int Cout, int relu, void* stream);
int b200em_conv3d_first_wgrad(const void* x, const float* in_scale_shift, const void* dz, int64_t dz_ld, float* dw, float* db,
                              int N, int D, int H, int W, int Cout, void* stream);

/* Weight gradient on the tensor cores (bf16 operands, fp32 accumulation in TMEM, fp32 atomics into dw).
 * dw (Cout,Cin,kd,kh,kw) fp32 += sum dz * x_hat (same contract as b200em_conv3d_wgrad_direct); db (nullable)
 * (Cout) fp32 += sum dz -- the bias gradient, fused into the dz operand load.  Takes Cin % 32 == 0, Cout % 16 == 0. */
int b200em_conv3d_wgrad_umma_supported(int Cin, int Cout, int kd, int kh, int kw);
int b200em_conv3d_wgrad_umma(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld,
                             float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw,
                             void* stream);

/* "w-stacked", depth-streamed tcgen05 weight gradient for 3 x kh x 3 filters (Cin % 32 == 0, Cout % 32 == 0): the three
 * w-taps share one MMA (N = 3*32 from three shifted shared-memory copies of dz), every x slice is staged once per
 * column of output slices (csrc/conv_wgrad_cs.cu).  Same contract and arguments as b200em_conv3d_wgrad_umma. */
int b200em_conv3d_wgrad_cs_supported(int Cin, int Cout, int kd, int kh, int kw);
int b200em_conv3d_wgrad_cs(const void* x, int64_t x_ld, const float* in_scale_shift, const void* dz, int64_t dz_ld,
                           float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw,
                           void* stream);

/* The two weight-gradient kernels above on IEEE fp16 operand copies of fp32 tensors (the h16 path, see b200em_conv3d_umma_h16):
 * x_f16 = fp16(2^kx * x_hat) and dz_f16 = fp16(2^kz * dz) from b200em_cvt_f16 (the norm apply is already in x_f16), *_absmax
 * the device floats the scales were derived from (NULL: unscaled); dw / db are accumulated in fp32 with the scales undone. */
int b200em_conv3d_wgrad_umma_h16(const void* x_f16, int64_t x_ld, const float* x_absmax, const void* dz_f16, int64_t dz_ld,
                                 const float* dz_absmax, float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd,
                                 int kh, int kw, void* stream);
int b200em_conv3d_wgrad_cs_h16(const void* x_f16, int64_t x_ld, const float* x_absmax, const void* dz_f16, int64_t dz_ld,
                               const float* dz_absmax, float* dw, float* db, int N, int D, int H, int W, int Cin, int Cout, int kd,
                               int kh, int kw, void* stream);

/* im2col of the first conv (thin K = taps*Cin): out (N,D,H,W,Kp) bf16 with out[vox][tap*Cin+ci] = x_hat[vox+tap][ci], zero in
 * the padding and for channels >= taps*Cin.  The conv is then a 1x1x1 conv with Kp input channels on the tcgen05 path. */
int b200em_im2col_taps(const void* x, int64_t x_ld, const float* in_scale_shift, int dtype, void* out, int N, int D, int H,
                       int W, int Cin, int kd, int kh, int kw, int Kp, void* stream);

/* ---- normalisation: nn.InstanceNorm3d(C) / nn.GroupNorm(min(32,C),C)  (unet.py:391-406) ------------------- */
/* sums[N][C][2] += (sum x, sum x^2) over the S voxels of each sample. */
int b200em_channel_sums(const void* x, int64_t x_ld, int dtype, int N, int64_t S, int C, float* sums, void* stream);
/* sums[N][C][2] += (sum g, sum g*x): the two reductions of the norm backward. */
int b200em_channel_dot_sums(const void* g, int64_t g_ld, const void* x, int64_t x_ld, int dtype,
                            int N, int64_t S, int C, float* sums, void* stream);
/* From per-channel sums make x_hat = scale*x + shift.  groups == C: InstanceNorm (gamma/beta may be NULL);
 * groups < C: GroupNorm over C/groups consecutive channels with affine gamma/beta.  Biased variance, eps inside
 * the sqrt.  mean_rstd[N][C][2] is kept for the backward. */
int b200em_norm_finalize(const float* sums, int N, int C, int64_t S, int groups, const float* gamma,
                         const float* beta, float eps, float* scale_shift, float* mean_rstd, void* stream);
/* y = scale[n,c]*x + shift[n,c] */
int b200em_affine_apply(const void* x, int64_t x_ld, const float* scale_shift, void* y, int64_t y_ld, int dtype,
                        int N, int64_t S, int C, void* stream);
/* Backward coefficients: dx = coef0*g + coef1*x + coef2 per (n,c); dgamma/dbeta (nullable) accumulated. */
int b200em_norm_bwd_finalize(const float* dsums, const float* mean_rstd, const float* gamma, int N, int C,
                             int64_t S, int groups, float* coef, float* dgamma, float* dbeta, void* stream);
/* out = (coef0*g + coef1*x + coef2 [+ add]) * (relu_mask ? x > 0 : 1).  coef == NULL means out = g [+ add].
 * absmax (nullable, here and in b200em_maxpool3d_bwd / b200em_head_bwd): device float, zeroed by the caller, raised to max |out|
 * with atomicMax -- the h16 path derives the fp16 operand scale of a gradient tensor from it without another pass over it. */
int b200em_norm_bwd_apply(const void* g, int64_t g_ld, const void* x, int64_t x_ld, const float* coef,
                          const void* add, int64_t add_ld, void* out, int64_t out_ld, int dtype,
                          int N, int64_t S, int C, int relu_mask, float* absmax, void* stream);

/* fp32 -> two bf16 tensors with hi + lo ~= x_hat (x_hat = scale*x + shift when in_scale_shift is given, else x): hi =
 * bf16(x_hat), lo = bf16(x_hat - hi).  Lets the bf16 tensor-core weight-gradient kernels accumulate an fp32-class result as
 * hi*hi + hi*lo + lo*hi (~16 mantissa bits).  x (N,S,C) with pitch x_ld; hi / lo contiguous (pitch C). */
int b200em_split_bf16(const float* x, int64_t x_ld, const float* in_scale_shift, void* hi, void* lo, int N, int64_t S, int C,
                      void* stream);

/* ---- nn.MaxPool3d(factor) (unet.py:645, 316) -------------------------------------------------------------- */
/* (D,H,W) are the INPUT dims; sums (nullable) [N][C][2] += stats of the pooled output. */
int b200em_maxpool3d_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int dtype, int N, int D, int H, int W,
                         int C, int fd, int fh, int fw, float* sums, void* stream);
/* out[hi] = ((hi is the first max of its window ? dp[window] : 0) [+ a[hi]]) * (relu_mask ? x[hi] > 0 : 1), where
 * a = add, or with coef (nullable; (c0,c1,c2) per (n,c), sample stride coef_nstride floats) a = c0*add + c1*x + c2: the norm
 * backward of the consuming decoder block applied on the fly to its raw data gradient, so the skip gradient is never stored. */
int b200em_maxpool3d_bwd(const void* x, int64_t x_ld, const void* dp, int64_t dp_ld, const void* add, int64_t add_ld,
                         const float* coef, int64_t coef_nstride, void* out, int64_t out_ld, int dtype, int N, int D, int H, int W,
                         int C, int fd, int fh, int fw, int relu_mask, float* absmax, void* stream);

/* ---- F.interpolate(mode="trilinear", align_corners=False), integer scale (unet.py:456) --------------------- */
/* (D,H,W) are the LOW-resolution dims. */
int b200em_upsample_trilinear_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int dtype, int N, int D, int H,
                                  int W, int C, int fd, int fh, int fw, float* sums, void* stream);
/* dx = U^T dy (transpose of the interpolation), or with coef (nullable, as above) U^T (c0*dy + c1*up + c2) where up = U zlow is the
 * up-sampled tensor itself (the first half of the decoder block's input): the block's norm backward fused in.  By linearity this
 * is c0 U^T dy + c1 (U^T U) zlow + c2 U^T 1 -- a 3-tap-per-axis stencil on the LOW-resolution zlow and a constant -- so the
 * high-resolution tensor is not read.  zlow: (N, D, H, W, C) with pitch zlow_ld. */
int b200em_upsample_trilinear_bwd(const void* dy, int64_t dy_ld, const void* zlow, int64_t zlow_ld, const float* coef,
                                  int64_t coef_nstride, void* dx, int64_t dx_ld, int dtype, int N, int D, int H, int W, int C,
                                  int fd, int fh, int fw, void* stream);

/* ---- out_conv (1x1x1) + final activation (unet.py:202-205, 638, 162-172) ---------------------------------- */
/* x NDHWC (Cin) -> out NCDHW fp32 (Cout): out = act(W x + b); w is (Cout,Cin) fp32 = the torch weight. */
int b200em_head_fwd(const void* x, int64_t x_ld, int dtype, const float* w, const float* bias, float* out,
                    int N, int64_t S, int Cin, int Cout, int act, void* stream);
/* grad_out, out: NCDHW fp32.  dx NDHWC (nullable), multiplied by [x > 0] when relu_mask (x is then the post-ReLU
 * output of the last conv block, unet.py:437); dw (Cout,Cin) and db (Cout) accumulated. */
int b200em_head_bwd(const float* grad_out, const float* out, const void* x, int64_t x_ld, int dtype, const float* w,
                    void* dx, int64_t dx_ld, float* dw, float* db, int N, int64_t S, int Cin, int Cout, int act,
                    int relu_mask, float* absmax, void* stream);

/* ---- DiceLoss [+ ApplyAndRemoveMask("multiply")]  (loss/dice.py:34-93, loss/wrapper.py:84-87,129-152) ------- */
/* pred (N,C,S) fp32 or bf16; target fp32 with sample stride target_nstride (elements): channel c of sample n at
 * target + n*target_nstride + c*S.  mask (nullable) laid out like target.  sums[C][3] += (sum pm*tm, sum pm^2,
 * sum tm^2) with pm = p*m, tm = t*m. */
int b200em_dice_sums(const void* pred, int pred_dtype, const float* target, const float* mask,
                     int64_t target_nstride, int N, int C, int64_t S, float* sums, void* stream);
/* loss (device scalar) and backward coefficients coef[C][2] (A_c, B_c): dL/dp = A_c*t*m^2 + B_c*p*m^2.
 * reduce: 0 sum, 1 mean, 2 max, 3 min, 4 none(per-channel losses written to loss[0..C)); channelwise in {0,1}. */
int b200em_dice_finalize(const float* sums, int C, float eps, int channelwise, int reduce, float* loss, float* coef,
                         void* stream);
/* grad_pred (N,C,S) of grad_dtype = gout[0..] * (A_c t m^2 + B_c p m^2); gout is a device scalar (or C values,
 * reduce=4). */
int b200em_dice_bwd(const void* pred, int pred_dtype, const float* target, const float* mask, int64_t target_nstride,
                    const float* coef, const float* gout, int gout_per_channel, void* grad_pred, int grad_dtype,
                    int N, int C, int64_t S, void* stream);

/* ---- Dice-family losses beyond DiceLoss (csrc/segloss.cu): DiceLossWithLogits, BCEDiceLoss, BCEDiceLossWithLogits
 * (loss/dice.py:136-256), DistanceLoss / DiceBasedDistanceLoss (loss/distance_based.py:7-69) ---------------------------------
 * chan[C][4] = (w_dice, w_bce, w_mse, use_mask) per channel selects the terms a channel contributes.  With p = logits ?
 * sigmoid(x) : x, m = use_mask ? mask : 1 (mask element of (n, c, i) at mask + n*mask_nstride + c*mask_cstride + i; a channel
 * stride of 0 broadcasts one mask channel), pm = p*m, tm = t*m:
 *   sums[C][5] += (sum pm*tm, sum pm^2, sum tm^2, sum bce, sum (pm-tm)^2)
 *   loss = reduce_c w_dice[c]*(1 - 2 num_c/max(den_c, eps)) + sum_c (w_bce[c]*bce_c + w_mse[c]*mse_c) / numel
 * BCE as ATen: log clamped at -100, gradient (p-t)/max(p(1-p), 1e-12); with logits max(x,0) - x t + log1p(exp(-|x|)).
 * coef[C][4] = (A_c, B_c, w_bce/numel, w_mse/numel); reduce codes as b200em_dice_finalize. */
int b200em_segloss_sums(const void* pred, int pred_dtype, const float* target, const float* mask, int64_t target_nstride,
                        int64_t mask_nstride, int64_t mask_cstride, const float* chan, int logits, int N, int C, int64_t S,
                        float* sums, void* stream);
int b200em_segloss_finalize(const float* sums, const float* chan, int C, float eps, int channelwise, int reduce, float numel,
                            float* loss, float* coef, void* stream);
int b200em_segloss_bwd(const void* pred, int pred_dtype, const float* target, const float* mask, int64_t target_nstride,
                       int64_t mask_nstride, int64_t mask_cstride, const float* chan, const float* coef, const float* gout,
                       int gout_per_channel, int logits, void* grad_pred, int grad_dtype, int N, int C, int64_t S, void* stream);

/* ---- AffinityTransform / BoundaryTransform (transform/label.py:248-327, 100-129) --------------------------- */
/* labels (N,D,H,W) int64; offsets: n_off*3 host ints (dz,dy,dx); out (N, channels, D,H,W) fp32 with channel
 * order [fg?][n_off disaffinities][fg-mask?][n_off masks]  (masks only if add_mask). */
int b200em_affinity_targets(const int64_t* labels, float* out, int N, int D, int H, int W, const int* offsets,
                            int n_off, int has_ignore, int64_t ignore_label, int add_binary_target, int add_mask,
                            int include_ignore_transitions, void* stream);
/* out (N, 1 or 2, D,H,W) fp32: [foreground?][boundary], boundary = some in-bounds 6-neighbour differs. */
int b200em_boundary_targets(const int64_t* labels, float* out, int N, int D, int H, int W, int add_binary_target,
                            void* stream);
/* NoToBackgroundBoundaryTransform (mode 1: aux = mask_label, bg = bg_label; label.py:133-189) and
 * BoundaryTransformWithIgnoreLabel (mode 2: aux = ignore_label; label.py:192-244): the label boundary, with the boundaries of
 * [lab != bg] resp. [lab == aux] set to aux; optional channel 0 = [lab != bg] with lab == aux -> aux.  out fp32. */
int b200em_boundary_targets_masked(const int64_t* labels, float* out, int N, int D, int H, int W, int add_binary_target,
                                   int mode, int64_t aux_label, int64_t bg_label, void* stream);
/* OneHotTransform (label.py:330-353): out (N, n_classes, S) fp32 = [labels == class_ids[k]]; class_ids on the device. */
int b200em_one_hot(const int64_t* labels, const int64_t* class_ids, int n_classes, float* out, int N, int64_t S, void* stream);
/* segmentation_to_affinities (loss/affinity_side_loss.py:70-89): out (N, n_off, D,H,W) fp32 = [lab[p] == lab[clamp(p+off)]]
 * (affinities, replication at the border, no mask). */
int b200em_segmentation_affinities(const int64_t* labels, float* out, int N, int D, int H, int W, const int* offsets, int n_off,
                                   void* stream);
/* Fused target+loss reductions: Dice sums of pred against AffinityTransform(offsets, add_mask=True) targets computed
 * on the fly from labels (no target tensor in HBM).  Same sums/coef contract as b200em_dice_sums / _bwd. */
int b200em_affinity_dice_sums(const void* pred, int pred_dtype, const int64_t* labels, int N, int D, int H, int W,
                              const int* offsets, int n_off, int has_ignore, int64_t ignore_label,
                              int include_ignore_transitions, float* sums, void* stream);
int b200em_affinity_dice_bwd(const void* pred, int pred_dtype, const int64_t* labels, int N, int D, int H, int W,
                             const int* offsets, int n_off, int has_ignore, int64_t ignore_label,
                             int include_ignore_transitions, const float* coef, const float* gout, void* grad_pred,
                             int grad_dtype, void* stream);

/* ---- tiled inference I/O (csrc/tiling.cu): predict_with_halo's per-block host work on a volume resident in HBM ----------
 * (util/prediction.py:98-142 _load_block, transform/raw.py:40-65 standardize, prediction.py:287-309 crop + mask + write) */
#define B200EM_MAX_TILE_BLOCKS 16
#define B200EM_RAW_U8 0
#define B200EM_RAW_I8 1
#define B200EM_RAW_U16 2
#define B200EM_RAW_I16 3
#define B200EM_RAW_I32 4
#define B200EM_RAW_U32 5
#define B200EM_RAW_F16 6
#define B200EM_RAW_F32 7
#define B200EM_RAW_F64 8
/* vol (C,D,H,W) raw dtype -> out (nblocks, C, bd,bh,bw) fp32: block i covers [begin_i, begin_i + (bd,bh,bw)) (begin = offset -
 * halo, may be negative), clipped to the volume and completed by reflecting the clipped data (np.pad "reflect").  block_begins:
 * 3*nblocks host ints.  stats (nullable) [nblocks][2] doubles += (sum, sum of squares) of each block (caller zeroes). */
int b200em_gather_blocks(const void* vol, int raw_dtype, int C, int D, int H, int W, const int* block_begins, int nblocks, int bd,
                         int bh, int bw, float* out, double* stats, void* stream);
/* x (nblocks, per_block) fp32, in place: (x - mean) / (std + eps), population statistics from stats[nblocks][2]. */
int b200em_standardize_blocks(float* x, int nblocks, int64_t per_block, const double* stats, float eps, void* stream);
/* pred (nblocks, Cp, bd,bh,bw) fp32 -> out (nc, D,H,W) fp32: out[c, begin_i + v] = pred[i, c0 + c, halo + v] for v in the inner
 * block shape_i (truncated last blocks), set to 0 where mask (nullable, (D,H,W) bytes) is 0.  begins / shapes: host ints. */
int b200em_scatter_blocks(const float* pred, int Cp, int bd, int bh, int bw, int hd, int hh, int hw, const int* block_begins,
                          const int* block_shapes, int nblocks, float* out, int D, int H, int W, int c0, int nc,
                          const unsigned char* mask, void* stream);

/* ---- utilities ---------------------------------------------------------------------------------------- */
/* Write `bytes` bytes of zeros (L2 flush helper for bench.py and buffer clears without a torch launch). */
int b200em_memset_zero(void* p, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200EM_H_ */
