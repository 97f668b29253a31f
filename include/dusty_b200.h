/*
 * dusty_b200.h -- C ABI of libdusty_b200.so: the B200 (sm_100a) kernels behind the
 * DUSty-v2 generator/discriminator hot path.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; no torch / ATen types;
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (outputs and workspaces included) -- nothing
 *     is allocated here;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous;
 *   - return value: 0 on success, negative DUSTY_E* on error, never throws.
 *     dusty_last_error() gives a thread-local message.  The Python host turns a
 *     non-zero status into RuntimeError, which is what the reference's
 *     TORCH_CHECKs surface as (fused_bias_act.cpp:10-16, upfirdn2d.cpp:10-16);
 *   - dtype is the storage type of activations (DUSTY_F32 / DUSTY_BF16);
 *     accumulation is always fp32.
 *
 * Each entry cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef DUSTY_B200_H_
#define DUSTY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUSTY_ABI_VERSION 1

enum { DUSTY_F32 = 0, DUSTY_BF16 = 1 };
enum { DUSTY_OK = 0, DUSTY_EINVAL = -1, DUSTY_ECUDA = -2, DUSTY_EUNSUPPORTED = -3 };
/* boundary extension used by the FIR family */
enum { DUSTY_PAD_ZERO = 0, DUSTY_PAD_CIRCULAR = 1, DUSTY_PAD_REPLICATE = 2, DUSTY_PAD_REFLECT = 3 };

/* ---- library queries -------------------------------------------------------------- */
int dusty_abi_version(void);
const char *dusty_last_error(void);
/* compute capability major*10+minor of the current device, or <0 */
int dusty_query_sm(void);
/* number of kernels launched by this library in this process (all threads) */
int64_t dusty_launch_count(void);

/* ---- a3: fused bias + activation ------------------------------------------------------
 * Replaces fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *   gans/models/ops/fused_act/fused_bias_act.cpp:18-32, fused_bias_act_kernel.cu:18-105.
 * x viewed as [N, C, inner] contiguous; bias index = (i / inner) % C.
 * bias == NULL / ref == NULL mean "absent" (the reference passes empty tensors).
 * act: 1 linear, 3 leaky-relu.  grad: 0 forward, 1 first derivative gated by `ref`,
 * 2 second derivative (zero).  y = f(x + b) * scale. */
int dusty_bias_act(const void *x, const void *bias, const void *ref, void *y, int64_t n_elem,
                   int C, int64_t inner, int act, int grad, float alpha, float scale, int dtype,
                   void *stream);

/* Fused backward: dx = dy * (out > 0 ? 1 : alpha) * scale and db[c] += sum dx (fp32, the
 * caller zero-fills db).  Replaces FusedLeakyReLUFunctionBackward.forward
 *   gans/models/ops/fused_act/fused_act.py:20-44 (kernel launch + ATen .sum(dim)).
 * db may be NULL. */
int dusty_bias_act_bwd(const void *dy, const void *out, void *dx, float *db, int64_t N, int C,
                       int64_t inner, float alpha, float scale, int dtype, void *stream);

/* ---- a4/a5: FIR family (upfirdn2d, Resample, BlurVH, filter2d, Pad) -------------------
 * out[n,my,mx] = sum_{ty,tx} K[ty,tx] * XU(my*down_y + ty - pad_y0, mx*down_x + tx - pad_x0)
 * XU(qy,qx) = X(qy/up_y, qx/up_x) when both divide exactly, else 0; X() extends x by
 * mode_y / mode_x.  `taps` is a DEVICE fp32 [kh,kw] array; flip != 0 reads it reversed in
 * both axes (true convolution, as upfirdn2d does).  x: [N, in_h, in_w], y: [N, out_h, out_w].
 *
 * With mode ZERO and flip=1 this is upfirdn2d_op.upfirdn2d (minor == 1)
 *   gans/models/ops/upfirdn2d/upfirdn2d.cpp:17-31, upfirdn2d_kernel.cu:44-425
 * with out_h = (in_h*up_y + pad_y0 + pad_y1 - kh + down_y) / down_y.
 * With CIRCULAR (W) / REPLICATE (H) it is Resample.forward
 *   gans/models/ops/common.py:105-135; with a 1x1 unit tap it is Pad.forward (:10-24). */
int dusty_fir2d(const void *x, void *y, const float *taps, int kh, int kw, int flip, int64_t N,
                int in_h, int in_w, int out_h, int out_w, int up_y, int up_x, int down_y,
                int down_x, int pad_y0, int pad_x0, int mode_y, int mode_x, int dtype,
                void *stream);
/* Exact adjoint (transpose) of dusty_fir2d with the same parameters:
 * dy: [N, out_h, out_w] -> dx: [N, in_h, in_w].  Serves backward of every FIR op and, the
 * op being linear, dusty_fir2d itself serves the double backward
 *   (UpFirDn2dBackward, gans/models/ops/upfirdn2d/upfirdn2d.py:20-85). */
int dusty_fir2d_adj(const void *dy, void *dx, const float *taps, int kh, int kw, int flip,
                    int64_t N, int in_h, int in_w, int out_h, int out_w, int up_y, int up_x,
                    int down_y, int down_x, int pad_y0, int pad_x0, int mode_y, int mode_x,
                    int dtype, void *stream);
/* Reference-shaped convenience entry: zero padding, flipped taps, minor == 1. */
int dusty_upfirdn2d(const void *x, const float *kernel, void *y, int64_t major, int in_h,
                    int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                    int pad_x0, int pad_x1, int pad_y0, int pad_y1, int dtype, void *stream);

/* Fast path of Resample.forward (gans/models/ops/common.py:105-135) for 4-tap separable
 * windows, ring=True, direction "hw": up == 1 is the same-size blur (pads 2,1; ResidualBlock
 * dusty_v2.py:331), up == 2 the 2x upsampling (SynthesisBlock.resample dusty_v2.py:85-90).
 * k0..k3 are the per-axis taps exactly as stored in the module's `kernel` buffer.
 * x: [N, H, W] -> y: [N, up*H, up*W]; adjoint != 0 applies the exact transpose
 * (y-shaped input -> x-shaped output).  W must be a multiple of 16 bytes / sizeof(T). */
int dusty_resample4(const void *x, void *y, float k0, float k1, float k2, float k3, int64_t N,
                    int H, int W, int up, int adjoint, int dtype, void *stream);

/* dusty_resample4(up = 2, forward) that also adds sum(y^2) to *sumsq (fp32 device scalar, zeroed by
 * the caller): SynthesisBlock feeds the upsampled tensor straight into conv1, whose training-mode
 * EMA normaliser needs mean(x^2) of its input (gans/models/ops/style.py:99-102) -- the statistic
 * comes out of the producing kernel instead of a second pass over the 4N outputs. */
int dusty_up2_sumsq(const void *x, void *y, float *sumsq, float k0, float k1, float k2, float k3,
                    int64_t N, int H, int W, int dtype, void *stream);

/* Fast path of Pad.forward (gans/models/ops/common.py:10-24) for halos up to 4 pixels:
 * x: [N, H, W] -> y: [N, H+pt+pb, W+pl+pr]; mode_y in {REPLICATE, REFLECT}, mode_x in
 * {CIRCULAR, REPLICATE, REFLECT}.  adjoint != 0: y-shaped gradient in, x-shaped out. */
int dusty_pad2d(const void *x, void *y, int64_t N, int H, int W, int pt, int pb, int pl, int pr,
                int mode_y, int mode_x, int adjoint, int dtype, void *stream);

/* ---- channels-last (NHWC) variants used by the discriminator trunk ---------------------
 * Same semantics as dusty_bias_act / dusty_bias_act_bwd / dusty_pad2d / dusty_resample4(up=1)
 * on tensors stored [B, H, W, C] (torch.channels_last); C must be a multiple of the 16-byte
 * vector width (and C / width a power of two <= 256 for the bias kernels). */
int dusty_bias_act_cl(const void *x, const void *bias, const void *ref, void *y, int64_t n_elem,
                      int C, int act, int grad, float alpha, float scale, int dtype, void *stream);
int dusty_bias_act_bwd_cl(const void *dy, const void *out, void *dx, float *db, int64_t rows, int C,
                          float alpha, float scale, int dtype, void *stream);

/* End of a ResidualBlock (gans/models/dusty_v2.py:387-396: bias_act2, + skip, / sqrt 2) as one
 * NHWC pass:  y = (lrelu(x + bias[c]) * gain + skip) * c.  bias has the activation dtype. */
int dusty_bias_act_add_cl(const void *x, const void *bias, const void *skip, void *y, int64_t n_elem,
                          int C, float alpha, float gain, float c, int dtype, void *stream);
/* g = dy * c; dskip = g (may be NULL); dx = g * gate(x + bias) * gain; db fp32 [C] += sum dx
 * (may be NULL; the caller zeroes it). */
int dusty_bias_act_add_bwd_cl(const void *dy, const void *x, const void *bias, void *dx, void *dskip,
                              float *db, int64_t rows, int C, float alpha, float gain, float c,
                              int dtype, void *stream);
int dusty_pad2d_cl(const void *x, void *y, int B, int H, int W, int C, int pt, int pb, int pl,
                   int pr, int mode_y, int mode_x, int adjoint, int dtype, void *stream);
/* pad == 1 fuses the 1-pixel ring padding that follows the blur in ResidualBlock.residual
 * (dusty_v2.py:337: conv2(resample(h)), conv2 = Pad(1, ring) + conv): forward writes
 * [B, H+2, W+2, C]; the adjoint takes that padded gradient and folds the halo on load. */
int dusty_blur4_cl(const void *x, void *y, float k0, float k1, float k2, float k3, int B, int H,
                   int W, int C, int adjoint, int pad, int dtype, void *stream);

/* Backward of  pre -> yact = lrelu(pre + bias, alpha) * scale -> dusty_blur4_cl(pad = 1)  as ONE
 * kernel: gpad [B, H+2, W+2, C] is the gradient of the padded blurred tensor, yact [B, H, W, C] the
 * activation's output; gpre = blur_pad^T(gpad) * gate(yact) * scale; db: fp32 [64][C] partial sums of
 * gpre over pixels (caller zero-fills and adds the 64 rows: spread over replicas, the CTAs' atomics
 * do not serialise on C addresses).  Replaces UpFirDn2dBackward of the ResidualBlock's Resample + the Pad adjoint +
 * FusedLeakyReLUFunctionBackward over the same tensor (gans/models/dusty_v2.py:300-315,
 * ops/fused_act/fused_act.py:34-66). */
int dusty_blur4_cl_adj_act(const void *gpad, const void *yact, void *gpre, float *db, float k0,
                           float k1, float k2, float k3, int B, int H, int W, int C, float alpha,
                           float scale, int dtype, void *stream);

/* Skip branch of ResidualBlock (dusty_v2.py:387-396): conv1x1_stride2(Resample(x)) only reads
 * the blurred image at even rows / columns.  y[b,i,j,:] = blur4(x)[b,2i,2j,:], y is
 * [B,H/2,W/2,C]; adjoint != 0: x is the [B,H/2,W/2,C] gradient, y the [B,H,W,C] result.
 * H, W even; C a multiple of the 16-byte vector width. */
int dusty_blur4_down2_cl(const void *x, void *y, float k0, float k1, float k2, float k3, int B,
                         int H, int W, int C, int adjoint, int dtype, void *stream);

/* Forward of the ResidualBlock's input fork in one pass over x (NHWC): xp = Pad(1, circular W /
 * replicate H)(x) [B, H+2, W+2, C] for conv1 and xd = blur(x)[::2, ::2] [B, H/2, W/2, C] for the skip
 * branch (gans/models/dusty_v2.py:300-315: conv1's Pad, the skip's Resample + stride-2 conv). */
int dusty_residual_fork_fwd_cl(const void *x, void *xp, void *xd, float k0, float k1, float k2, float k3,
                               int B, int H, int W, int C, int dtype, void *stream);

/* Backward of the ResidualBlock input fork (dusty_v2.py:387-396: x feeds conv1 = Pad(1, ring) +
 * conv AND skip = conv1x1_stride2(Resample(x))): dx = pad_adjoint(g_pad) + blur4_down2_adjoint(
 * g_down) in one pass, replacing two adjoint launches plus the autograd accumulation add.
 * g_pad [B,H+2,W+2,C] (replicate H / circular W, one pixel), g_down [B,H/2,W/2,C], dx [B,H,W,C]. */
int dusty_residual_fork_bwd_cl(const void *g_pad, const void *g_down, void *dx, float k0, float k1,
                               float k2, float k3, int B, int H, int W, int C, int dtype,
                               void *stream);

/* ---- a5/a6: AdaptiveAugment's geometric pipeline ------------------------------------------
 * Single-axis zero-padded polyphase FIR: upfirdn2d with a [1,k] (axis 1 = x) or [k,1]
 * (axis 0 = y) kernel, fp32, up/down in {1,2}:
 *   n_out = (n_in*up + pad0 + pad1 - k + down) / down  along the filtered axis.
 * flip != 0 applies the taps reversed (true convolution, as upfirdn2d does).  The adjoint is
 * the same entry with up<->down, flip toggled and the reference's g_pad
 * (gans/models/ops/upfirdn2d/upfirdn2d.py:108-116). */
int dusty_fir1d(const float *x, float *y, const float *taps, int k, int flip, int64_t N, int in_h,
                int in_w, int axis, int up, int down, int pad0, int pad1, void *stream);
/* F.affine_grid(theta, [N,C,Ho,Wo], align_corners=False) + F.grid_sample(bilinear, zeros,
 * align_corners=False) fused (adaptive_augment.py:523-524).  theta: fp32 [N,2,3].
 * adjoint != 0: src is the [N,C,Ho,Wo] gradient, dst the [N,C,Hi,Wi] image gradient. */
int dusty_affine_warp(const float *src, float *dst, const float *theta, int N, int C, int Hi,
                      int Wi, int Ho, int Wo, int adjoint, void *stream);

/* ---- a2: Fourier features --------------------------------------------------------------
 * Replaces FourierFeature.forward gans/models/ops/fourier.py:77-82.
 * angle: fp32 [Ba, 2, P] (elevation, azimuth); freqs: fp32 [F, 2]; phase: fp32 [F];
 * out: [Ba, 2F, P] (sin block then cos block) in out_dtype. */
int dusty_fourier(const float *angle, const float *freqs, const float *phase, void *out, int Ba,
                  int F, int64_t P, int out_dtype, void *stream);

/* angle pyramid step: cat(sin,cos) -> Resample(down=2, [1,3,3,1], ring) -> atan2
 *   SynthesisBlock.downsample_angle gans/models/dusty_v2.py:135-140.
 * in: fp32 [Ba, 2, H, W] -> out: fp32 [Ba, 2, H/2, W/2]. */
int dusty_angle_down2(const float *angle_in, float *angle_out, int Ba, int H, int W, void *stream);

/* ---- a1: modulated 1x1 convolution as a batched contraction ---------------------------
 * Replaces the grouped conv of ModConv2d.forward gans/models/ops/style.py:106-121 (ksize 1).
 * The per-sample effective weights wb[B, O, K] (modulation, demodulation and the EMA
 * normaliser folded in, style.py:72-103) are produced by the host-side weight prep.
 *
 *   Y[b,o,p] = sum_k wb[b,o,k] * X(b,k,p),   X = x1 for k < C1, x2 (k - C1) otherwise
 *
 * x1: [B, C1, P] activations; x2: [B2, C2, P] Fourier features with B2 == B or 1 (batch
 * shared); K = C1 + C2; either part may be empty (C == 0, pointer ignored).
 * Epilogue: + bias[o] (fp32, may be NULL), then act (1 linear / 3 lrelu(alpha)) * scale.
 * wb has dtype `wdtype`; x1/x2/y have dtype `dtype`.
 * impl: 0 = auto, 1 = SIMT fp32-FMA kernel, 2 = tcgen05 tensor-core kernel (bf16 only; with a
 * batch-shared x2 the Fourier half runs as one dense GEMM over the [(B*O), K] weight matrix,
 * several samples side by side in one accumulator), 3 = tcgen05 with per-sample tiles only,
 * 4 = tcgen05 with fp32 OUTPUT (y / dx1 are fp32 while `dtype` = bf16 names the operands): the
 * fp32 mode, whose operands are split-bf16 with a three times longer K axis
 * (dusty_split_bf16x3).
 * ema_var (optional device scalar, ModConv2d.ema_var): the accumulator is multiplied by
 * 1 / (sqrt(ema_var) + 1e-8) ahead of the bias -- the EMA normaliser of style.py:99-103 applied to
 * the product instead of to the weights, so that wb does not depend on the activation statistics
 * (all layers' weights can then be prepared ahead of the activation chain).  tcgen05 path only.
 * ema_rows (optional HOST array of O device scalars, O <= 4; exclusive with ema_var): the same
 * per OUTPUT ROW for the heads, where every row belongs to its own ModConv2d (dusty_v2.py:48-60);
 * applied by the small-O CUDA-core kernels to the weight rows as they are staged.
 * sumsq (optional device scalar, caller zero-fills; tcgen05 bf16-output path): += sum of the
 * squares of the stored outputs -- the statistic the NEXT ModConv2d takes of its input
 * (style.py:99-102), accumulated in the epilogue instead of by a pass over the tensor. */
int dusty_modconv_fwd(const void *wb, const void *x1, const void *x2, const float *bias, void *y,
                      int B, int O, int C1, int C2, int B2, int64_t P, int act, float alpha,
                      float scale, int dtype, int wdtype, int impl, const float *ema_var,
                      const float *const *ema_rows, float *sumsq, void *stream);
/* dX1[b,k,p] = sum_o wb[b,o,k] * dY[b,o,p]  for k < C1 (Fourier channels carry no grad);
 * ema_var / ema_rows as above (the same factors on the way back). */
int dusty_modconv_bwd_dx(const void *wb, const void *dy, void *dx1, int B, int O, int C1, int K,
                         int64_t P, int dtype, int wdtype, int impl, const float *ema_var,
                         const float *const *ema_rows, void *stream);
/* dwb[b,o,k] = sum_p dY[b,o,p] * X(b,k,p)  (fp32 output, overwritten).
 * dw_ld: row pitch of dwb in floats (0 = C1 + C2); a wider pitch writes the columns of one source
 * into a [B, O, dw_ld] tensor whose other columns come from elsewhere: with a batch-SHARED x2 the
 * Fourier columns of all samples are ONE dense GEMM, dusty_gemm_bf16 over dY viewed [(B*O), P]
 * (the Fourier block is then read once instead of once per sample). */
int dusty_modconv_bwd_dw(const void *dy, const void *x1, const void *x2, float *dwb, int B, int O,
                         int C1, int C2, int B2, int64_t P, int dtype, int impl, long long dw_ld,
                         void *stream);

/* Per-sample effective weights (modulation, pre-normalisation, demodulation, EMA
 * normaliser) -- the part of ModConv2d.forward before the grouped conv, style.py:72-103.
 * slin: fp32 [B, I] = mod(style) (the EqualLR linear stays a library GEMM); weight: fp32
 * [O, I]; ema_var: device scalar or NULL; wb: [B, O, I] in wdtype.
 * stats: fp32 workspace of B + 2 + B*O floats, kept for the backward. */
/* rot (optional, fp32 [B, 2F] = cos | sin of psi[b,f] = f_w[f] * shift_b): per-sample rotation
 * of the Fourier columns [C1, C1+F) (sin block) / [C1+F, C1+2F) (cos block).  The training-time
 * aug-coords azimuth shift (dusty_v2.py:266-274) with integer horizontal frequencies is
 * exactly this rotation of a batch-SHARED Fourier block (angle-addition identity), so the
 * contraction can read one L2-resident Fourier operand for the whole batch. */
int dusty_modprep_fwd(const float *slin, const float *weight, const float *ema_var, void *wb,
                      float *stats, int B, int O, int I, float scale, int demod, int wdtype,
                      const float *rot, int C1, int F, void *stream);
/* Analytic backward: gwb fp32 [B, O, I] -> dslin [B, I], dweight [O, I].
 * work: fp32 workspace of B*O + B*I + O*I + B + 1 floats.
 * ema_late (optional device scalar): the forward ran with ema_var = NULL and the contraction
 * applied the EMA normaliser (dusty_modconv_fwd's ema_var); gwb is then the gradient w.r.t. the
 * NORMALISED weights and the factor is read here, at backward time. */
int dusty_modprep_bwd(const float *gwb, const float *slin, const float *weight, const float *stats,
                      float *dslin, float *dweight, float *work, int B, int O, int I, float scale,
                      int demod, const float *rot, int C1, int F, const float *ema_late,
                      void *stream);

/* ---- a10: Gumbel-sigmoid raydrop -------------------------------------------------------
 * Replaces GumbelSigmoid.forward gans/models/ops/gumbel.py:23-29 (RelaxedBernoulli.rsample
 * closed form, uniform draw supplied by the caller) + RayDropModel.forward
 * gans/models/dusty_v1.py:20-25.  All fp32, n elements.
 * Outputs: mask (hard 0/1), image_out = lerp(image, rconst, 1-mask), dsoft = d soft/d logit
 * (saved for backward); count (int32, caller zero-fills, may be NULL) += sum(mask). */
int dusty_gumbel_raydrop_fwd(const float *logit, const float *image, const float *u, float *mask,
                             float *image_out, float *dsoft, int *count, int64_t n, float rconst,
                             float temperature, void *stream);
/* g_logit = (g_out*(image - rconst) + g_mask) * dsoft ; g_image = g_out * mask.
 * g_mask may be NULL. */
int dusty_gumbel_raydrop_bwd(const float *g_out, const float *g_mask, const float *image,
                             const float *mask, const float *dsoft, float *g_logit,
                             float *g_image, int64_t n, float rconst, void *stream);

/* ---- a14: range image -> point cloud ---------------------------------------------------
 * Replaces CoordBridge.convert(x, "inv_depth_norm", "point_map"|"point_set")
 *   gans/coords.py:139-155,178-185.
 * x: fp32 [B, HW] inverse depth (normalised).  trig: fp32 [4, HW] = cos(el), sin(el),
 * cos(az), sin(az), computed once on the host from CoordBridge.angle.
 * layout 0: out [B, 3, HW] (point_map); 1: out [B, HW, 3] (point_set, index h*W + w).
 * valid_count (int64, caller zero-fills, may be NULL) += number of valid pixels. */
int dusty_point_project(const float *x, const float *trig, float *out, long long *valid_count,
                        int B, int64_t HW, float min_depth, float max_depth, float tol,
                        int layout, void *stream);

/* ---- a12: minibatch standard deviation --------------------------------------------------
 * Replaces MinibatchStdDev.forward gans/models/ops/common.py:237-250 (features == 1).
 * x: [B, C, HW] -> y: [B, C+1, HW]; stat: fp32 [B/G] (overwritten) = the appended value per
 * group slot.  y == NULL: statistic only (x may then be in any per-sample element order, e.g.
 * NHWC -- the statistic is a mean over all C*HW positions). */
int dusty_minibatch_std_fwd(const void *x, void *y, float *stat, int B, int C, int64_t HW,
                            int group, float alpha, int dtype, void *stream);
/* dx = dy[:, :C] + d stat path.  dstat: fp32 [B/G] workspace.  dy == NULL: statistic-only
 * variant, dstat is then the INPUT gradient w.r.t. stat and dx holds only that path. */
int dusty_minibatch_std_bwd(const void *dy, const void *x, void *dx, float *dstat, int B, int C,
                            int64_t HW, int group, float alpha, int dtype, void *stream);

/* ---- a13 / ema_var: row-wise sum of squares ---------------------------------------------
 * out[r] (+)= sum_j x[r, j]^2, fp32.  accumulate == 0 overwrites (out must still be
 * zero-filled by the caller when rows are split across CTAs: see implementation note --
 * the entry zero-fills internally when accumulate == 0).
 * Serves the R1 penalty gans/trainer.py:440 and ModConv2d's EMA statistic style.py:100. */
int dusty_sumsq_rows(const void *x, float *out, int64_t rows, int64_t cols, int accumulate,
                     int dtype, void *stream);

/* ModConv2d's EMA side effect (style.py:99-102) from device-side sums, one tiny launch:
 * ema_var <- lerp(ema_var, (sum_a + rep_b * sum_b) * inv_numel, weight).  Either sum may be NULL. */
int dusty_ema_lerp(float *ema_var, const float *sum_a, const float *sum_b, float rep_b,
                   float inv_numel, float weight, void *stream);

/* ---- a7: aug-coords circular un-shift ---------------------------------------------------
 * Replaces cat([v,v],3) -> affine_grid -> grid_sample -> [..., :W]
 *   gans/models/dusty_v2.py:290-297.
 * out[b,c,h,j] = (1-f_b) v[(j+n_b) % W] + f_b v[(j+n_b+1) % W], n_b + f_b = shift01[b]*W.
 * adjoint != 0 applies the transpose (backward).  scale multiplies the result
 * (output_scale, dusty_v2.py:299-301). fp32. */
int dusty_circular_shift(const float *v, const float *shift01, float *out, int B, int C, int H,
                         int W, float scale, int adjoint, void *stream);

/* ---- a11: EqualLR weight preparation -----------------------------------------------------
 * Replaces the weight side of EqualLR + Conv2d, gans/models/ops/common.py:158-184,187-210: the
 * 1/sqrt(fan_in) scale folded into the filter, cast to the activation dtype, OIHW -> OHWI.
 * w: fp32 [O, C, RS] master weight; out: [O, RS, C] (bf16 or fp32) = w * scale. */
int dusty_weight_prep(const float *w, void *out, void *out_tco, int O, int C, int RS, float scale,
                      int out_dtype, void *stream);
/* out_tco (may be NULL): the same scaled filter as [RS, C, O] (O contiguous), the layout the
 * data-gradient kernels read (dusty_conv2d_tc tap mode / dusty_conv2d_halo_tc with flip). */
/* Adjoint: gw fp32 [O, C, RS] = g * scale, g in [O, RS, C] (g_nhwc = 1) or [O, C, RS] order. */
int dusty_weight_prep_adj(const void *g, float *gw, int O, int C, int RS, float scale, int g_dtype,
                          int g_nhwc, void *stream);

/* Filter gradient from dusty_conv2d_wgrad_tc's fp32 [RS, C, O] order to OHWI [O, RS, C] in
 * dst_dtype (the layout and dtype of the prepared filter it is the gradient of). */
int dusty_filter_rsco_to_ohwi(const float *src, void *dst, int O, int C, int RS, int dst_dtype,
                              void *stream);

/* ---- a11: discriminator stem -------------------------------------------------------------
 * Replaces BlurVH -> Conv2d(2 -> O, 1x1, EqualLR, no bias) -> FusedLeakyReLU(O), the first three
 * layers of Discriminator.layers, gans/models/dusty_v2.py:352-354 (BlurVH common.py:141-155,
 * fused_leaky_relu fused_act.py:93-129), as one pass.
 * x: [B, 1, H, W] (fp32 or bf16); w: fp32 [O, 2] effective weights (EqualLR scale folded in);
 * bias: fp32 [O] or NULL; y: bf16 NHWC [B, H, W, O]; taps k0..k2 of the 3-tap blur (vertical
 * pass clamps rows, horizontal pass wraps columns).  O in {8, 16, 32, 64}. */
int dusty_stem_fwd(const void *x, const float *w, const float *bias, void *y, int B, int H, int W,
                   int O, float k0, float k1, float k2, float alpha, float scale, int x_dtype,
                   void *stream);
/* One pass over (dy, y): dwb fp32 [O, 3] = (dW[o,0], dW[o,1], db[o]) (overwritten); dvh: fp32
 * [B, 2, H, W] gradient w.r.t. the blurred pair, or NULL when the input needs no gradient. */
int dusty_stem_bwd(const void *dy, const void *y, const void *x, const float *w, float *dvh,
                   float *dwb, int B, int H, int W, int O, float k0, float k1, float k2, float alpha,
                   float scale, int x_dtype, void *stream);
/* dx [B, 1, H, W] fp32 = adjoint of the two blurs applied to dvh. */
int dusty_stem_dx(const float *dvh, float *dx, int B, int H, int W, float k0, float k1, float k2,
                  void *stream);

/* ---- a11: dense convolutions of the discriminator trunk (tcgen05, NHWC bf16) ---------------
 * Replaces F.conv2d / cuDNN behind Conv2d + EqualLR, gans/models/ops/common.py:187-210, for
 * the ResidualBlock convolutions gans/models/dusty_v2.py:347-396 (SURVEY 8b dusty_conv2d_*).
 *
 * dusty_conv2d_tc is the implicit-GEMM primitive
 *   y[b,oh,ow,n] = act(sum_g sum_k A_g[b,oh,ow,k] * wpk[g,n,k] + bias[n]) * scale
 * x is [B,H_in,W_in,C] (NHWC), y is written through the element strides (y_off,y_sb,y_sh,y_sw)
 * with n contiguous; wpk is [G][O][K_g] bf16.
 *   mode 1 ("window", valid conv fprop): G filter rows; group g reads the S*C contiguous
 *     elements starting at x[b, oh*stride_h + tap_dh[g], ow*stride_w + tap_dw[g], 0]
 *     (K_g = S*C, wpk[g][n][s*C+c] = w[n][c][g][s]).
 *   mode 0 ("tap", dgrad): G taps; group g reads x[b, oh + tap_dh[g], ow + tap_dw[g], :]
 *     (K_g = C), rows / columns outside x contribute zero; unit stride only -- a strided
 *     dgrad is one call per output parity class with y_sh / y_sw doubled.
 * tap_dh / tap_dw are HOST arrays of G ints.  bias may be NULL; act: 1 linear, 3 leaky-ReLU.
 * Weight addressing (so that no per-call repacking is needed): element (g, n, k) of the filter
 * is read at wpk[wtap[g] * w_sg + n * w_sn + k]; w_sn = 0 / w_sg = 0 mean the dense strides
 * K_g / O * K_g, wtap = NULL means wtap[g] = g; w_taps = number of filter blocks in the tensor
 * (used with wtap).  E.g. window mode straight from an OHWI filter: w_sn = R*S*C, w_sg = S*C;
 * a dgrad parity class from a [R*S][C][O] tensor: wtap[g] = r*S + s of the class's taps.
 * out_dtype: DUSTY_BF16, or DUSTY_F32 (y and its strides / offsets in fp32 elements: the fp32
 * mode's split-bf16 contractions, dusty_split_bf16x3). */
int dusty_conv2d_tc(const void *x, const void *wpk, const float *bias, void *y, int B, int H_in,
                    int W_in, int C, int H_out, int W_out, int O, int mode, int G,
                    const int *tap_dh, const int *tap_dw, int S, int stride_h, int stride_w,
                    long long y_off, long long y_sb, long long y_sh, long long y_sw, int act,
                    float alpha, float scale, long long w_sn, long long w_sg, const int *wtap,
                    int w_taps, int out_dtype, void *stream);

/* Tap mode over `ncls` classes in ONE launch: the data gradient of a strided convolution is one
 * class per output parity (ph, pw), each with its own taps, output origin cls_y_off[c] and
 * extent cls_H_out[c] x cls_W_out[c] (output strides y_sh / y_sw are the doubled ones, shared).
 * tap_dh / tap_dw / wtap are the classes' tap lists concatenated (cls_G[c] entries each); the
 * tiles of the classes interleave so that every CTA gets the same mix of 1-, 2- and 4-tap
 * tiles.  Replaces aten::convolution_backward's input gradient for Conv2d(.., stride=2),
 * gans/models/dusty_v2.py:303,305 (conv2 / skip of ResidualBlock). */
int dusty_conv2d_tc_classes(const void *x, const void *wpk, void *y, int B, int H_in, int W_in,
                            int C, int O, int ncls, const int *cls_G, const int *tap_dh,
                            const int *tap_dw, const int *wtap, const int *cls_H_out,
                            const int *cls_W_out, const long long *cls_y_off, long long y_sb,
                            long long y_sh, long long y_sw, long long w_sn, long long w_sg,
                            int w_taps, int out_dtype, void *stream);

/* The same convolution family on the CUDA cores, fp32 accumulation, for fp32 parity mode and
 * every shape outside the tcgen05 kernels' domain (1- / 2- / 513-channel layers, 4x4 filters in
 * fp32, NCHW tensors): replaces F.conv2d / F.conv_transpose2d / aten::convolution_backward behind
 * gans/models/ops/common.py:158-210 (EqualLR + Conv2d) and gans/models/vanilla.py:7-105.
 *   mode 0: y  [B,O,Ho,Wo] = conv2d(x, w, stride, zero padding) * scale
 *   mode 1: dx [B,C,H,W]   = conv_transpose2d(dy, w, stride, padding) * scale  (data gradient)
 *   mode 2: dw fp32 [O,C,R,S] contiguous = filter gradient * scale
 * x_strides / y_strides / w_strides: HOST arrays of 4 element strides in logical (b,c,h,w) /
 * (b,o,oh,ow) / (o,c,r,s) order -- any memory layout.  dtype: DUSTY_F32 / DUSTY_BF16 of x, dy, w
 * and of the outputs of modes 0 and 1. */
int dusty_conv2d_simt(int mode, const void *x, const void *dy, const void *w, void *out, int B,
                      int C, int H, int W, int O, int Ho, int Wo, int R, int S, int stride_h,
                      int stride_w, int pad_h, int pad_w, const long long *x_strides,
                      const long long *y_strides, const long long *w_strides, float scale, int dtype,
                      void *stream);

/* ---- a11: linears of the discriminator epilogue (gans/models/dusty_v2.py:382-384:
 * EqualLR(Linear(65536 -> 512)), EqualLR(Linear(512 -> 1)); replaces F.linear / cuBLAS) ---------
 * C[M, N] (+)= alpha * A[M, K] * B[N, K]^T, fp32 accumulation, C fp32 row-major (leading
 * dimension ldc); accumulate != 0 adds into C.  Few output tiles -> split-K with TMA reduce-add.
 * dusty_gemm_tf32: tcgen05 kind::tf32 on fp32 operands read in place by TMA (the 134 MB weight
 *   matrix is streamed once, no cast pass); both operands K-major: element (row, k) at
 *   row * ld + k.
 * dusty_gemm_bf16: kind::f16 on bf16 operands; a_mn / b_mn = 1: the operand is MN-major (element
 *   (row, k) at k * ld + row), so data and weight gradient read the same tensors as the forward.
 * Pointers 16-byte aligned, leading dimensions multiples of 16 bytes. */
int dusty_gemm_tf32(const float *a, const float *b, float *c, int M, int N, int K, long long lda,
                    long long ldb, long long ldc, float alpha, int accumulate, void *stream);
int dusty_gemm_bf16(const void *a, const void *b, float *c, int M, int N, int K, int a_mn, int b_mn,
                    long long lda, long long ldb, long long ldc, float alpha, int accumulate,
                    void *stream);
/* The same product for tiny outputs (the 512 -> 1 head and its gradients) on the CUDA cores,
 * fp32, arbitrary element strides: A(m, k) at m * a_sm + k * a_sk, B(n, k) at n * b_sn + k * b_sk,
 * C(m, n) at m * c_sm + n * c_sn. */
int dusty_gemm_simt(const float *a, const float *b, float *c, int M, int N, int K, long long a_sm,
                    long long a_sk, long long b_sn, long long b_sk, long long c_sm, long long c_sn,
                    float alpha, void *stream);

/* ---- f4: KITTI scan -> range image (gans/datasets/kitti.py:216-220 numba `scatter`, 275-279
 * NEAREST resize + mask, 317-370 load_pts_as_img) -------------------------------------------------
 * points: [N, 4] fp32 (x, y, z, reflectance); depth: fp32 [N] norms; cell_h / cell_w: int32 [N]
 * image cell of every point (host-computed: ring index from scan unfolding, column from the azimuth; a ring index of -1
 * wraps to H - 1 as numpy indexing does); keys: caller-owned scratch of H * W 64-bit words.
 * out: fp32 [6, H, W_out] = (x, y, z, reflectance, depth, mask) of the NEAREST point of cell
 * (h, w * W / W_out), times mask = (min_depth <= depth <= max_depth); empty cells are zero. */
int dusty_scan_project(const float *points, const float *depth, const int *cell_h, const int *cell_w,
                       unsigned long long *keys, float *out, int N, int H, int W, int W_out,
                       float min_depth, float max_depth, void *stream);

/* ---- g1: fp32 mode on the tensor cores ------------------------------------------------------
 * fp32 tensor -> three bf16 terms concatenated along the contracted axis, so that an fp32
 * contraction runs on the bf16 tcgen05 kernels with ~2^-16 relative error per product:
 *   sum_k a_k b_k ~= [a_hi | a_hi | a_lo] . [b_hi | b_lo | b_hi],   x = x_hi + x_lo in bf16.
 * src is viewed as [outer][K][inner] through the element strides src_strides[3] (HOST array),
 * dst (bf16) as [outer][3K][inner] through dst_strides[3]; pattern 0 writes hi,hi,lo (the "a"
 * operand), pattern 1 hi,lo,hi (the "b" operand).  Serves Conv2d / ModConv2d in fp32 mode
 * (gans/models/ops/common.py:187-210, ops/style.py:68-126; reference config
 * configs/gans/dusty_v2.yaml:63-65 trains in fp32). */
int dusty_split_bf16x3(const float *src, void *dst, long long outer, long long K, long long inner,
                       const long long *src_strides, const long long *dst_strides, int pattern,
                       void *stream);

/* ---- f1: AdaptiveAugment as one device-side op ------------------------------------------------
 * Replaces AdaptiveAugment.forward for one-channel images and the axis-aligned geometric policies
 * of the shipped configs (gans/augment/adaptive_augment.py:271-291 get_padding, 386-469 sampling,
 * 471-545 pad -> up -> affine grid_sample -> down -> colour).
 * params: fp32 [B, 8] = ax, tx, dy, ty (the INVERSE transform [[ax,0,tx],[0,dy,ty]]), colour gain,
 * colour offset, 2 unused.
 * dusty_ada_apply: img / out fp32 [B, H, W]; one CTA per sample, the image resident in shared
 * memory (dusty_ada_apply_smem(H, W) bytes must fit in 227 KB), rows then columns through the 1-D
 * pad / up / interpolate / down pipeline; fixed maximum padding (W-1 / H-1, index arithmetic).
 *   mode 0: out = gain * A(img) + offset;  mode 1: the adjoint, out = gain * A^T(img);
 *   mode 2: mode 0 without the offset (the backward of the adjoint: R1 differentiates twice).
 * dusty_ada_sample: draws the transforms on the device with probability *p (device scalar,
 * AdaptiveAugment.p) per policy; Philox seeded by `seed`, one subsequence per sample, advanced by
 * the device-resident call counter `counter` (so that a captured graph draws fresh transforms at
 * every replay).  policy: HOST array of 11 floats = the multipliers lr_flip, ud_flip, int_trans,
 * iso_scale, frac_trans, brightness, contrast, luma_flip, hue, saturation, and the vertical
 * translation factor (0 with wonly_trans). */
long long dusty_ada_apply_smem(int H, int W);
int dusty_ada_apply(const float *img, float *out, const float *params, int B, int H, int W, int mode,
                    void *stream);
int dusty_ada_sample(float *params, const float *p, unsigned long long seed,
                     unsigned long long *counter, int B, int H, int W, const float *policy,
                     void *stream);

/* ---- f2: optimiser step and EMA (gans/trainer.py:30-41 ema_inplace, 128-171 Adam) ----------
 * Multi-tensor Adam exactly as torch.optim.Adam (no weight decay / amsgrad) over `count` fp32
 * tensors given as HOST arrays of device pointers: grads are multiplied by grad_scale first
 * (1 / world_size after a summing all-reduce), `step` is the 1-based update count (bias
 * corrections are computed on the host).  ema (array or NULL; entries may be NULL): after the
 * update ema[i] += ema_weight * (param[i] - ema[i]) -- the generator's EMA lerp towards the new
 * weights, folded into the same pass. */
int dusty_multi_adam(void *const *params, const void *const *grads, void *const *exp_avg,
                     void *const *exp_avg_sq, void *const *ema, const long long *numel, int count,
                     float lr, float beta1, float beta2, float eps, int step, float ema_weight,
                     float grad_scale, void *stream);
/* dst[i] = src[i] * scale for `count` fp32 tensors (gradient bucket packing / unpacking, EMA
 * buffer copies) in one launch per 32 tensors. */
int dusty_multi_copy(void *const *dst, const void *const *src, const long long *numel, int count,
                     float scale, void *stream);

/* Tools only: role-cycle counters of the tcgen05 convolution kernels (12 doubles; see
 * conv_tc.cu).  DUSTY_EUNSUPPORTED unless the library was built with -DDUSTY_ROLE_PROF. */
int dusty_conv_role_prof(double *out8, int reset);

/* Filter gradient of the valid convolution above:
 *   dwp[r][s*C+c][n] = sum_{b,oh,ow} x[b, oh*stride_h + r, ow*stride_w + s, c] * dy[b,oh,ow,n]
 * fp32 output [R][S*C][O].  ws: caller-owned fp32 workspace of at least
 * dusty_conv2d_wgrad_tc_workspace(...) elements (split-K partials; may be NULL if that is 0). */
long long dusty_conv2d_wgrad_tc_workspace(int B, int H_out, int W_out, int C, int O, int R, int S);
int dusty_conv2d_wgrad_tc(const void *x, const void *dy, float *dwp, float *ws, long long ws_elems,
                          int B, int H_in, int W_in, int C, int H_out, int W_out, int O, int R,
                          int S, int stride_h, int stride_w, void *stream);

/* Halo-resident variant for the thin layers (C = 32 or 64 input channels, O <= 128, filters up
 * to 3x3, unit stride): every input pixel is landed in shared memory once and the R*S taps are
 * shifted views of that patch (see conv_tc.cu).  y[b,oh,ow,n] = act(sum_{a,b',c}
 * x[b, oh + org_h + a, ow + org_w + b', c] * wpk[a*S + b'][n][c] + bias[n]) * scale, rows /
 * columns outside x read as zero.  fprop of a valid conv: org = (0, 0), wpk[t][n][c] =
 * w[n][c][r][s]; dgrad: x := dY, org = (-(R-1), -(S-1)), wpk[a*S+b'][c][n] = w[n][c][R-1-a][S-1-b'].
 * dusty_conv2d_halo_supported returns non-zero when the shape qualifies.
 * Weight addressing: element (t, n, c) at wpk[t' * w_sg + n * w_sn + c], t' = flip ? T-1-t : t;
 * w_sn = 0 / w_sg = 0 mean dense (C / O*C).  fprop straight from an OHWI filter: w_sn = R*S*C,
 * w_sg = C; dgrad from an un-flipped [R*S][C][O] tensor: flip = 1. */
int dusty_conv2d_halo_supported(int C, int O, int R, int S);
int dusty_conv2d_halo_tc(const void *x, const void *wpk, const float *bias, void *y, int B,
                         int H_in, int W_in, int C, int H_out, int W_out, int O, int R, int S,
                         int org_h, int org_w, long long y_off, long long y_sb, long long y_sh,
                         long long y_sw, int act, float alpha, float scale, long long w_sn,
                         long long w_sg, int flip, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DUSTY_B200_H_ */
