/*
 * kurosiwo_b200.h — C ABI of libkurosiwo_b200.so (sm_100a only).
 *
 * This is the drop-in boundary beneath the reference's Python surface.  The
 * reference (Orion-AI-Lab/KuroSiwo) has no FFI of its own: its training hot
 * path is `model(*inputs)` -> `criterion(output, mask)` -> `backward()` ->
 * `optimizer.step()` (training/change_detection_trainer.py:136-177), lowered
 * to cuDNN/ATen kernels.  Every entry point below replaces one family of those
 * library calls; the reference call site it replaces is cited per function.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers
 *     unless the name says `host`.
 *   - every function returns 0 on success, a negative KS_E* code on a bad
 *     argument, or a positive cudaError_t from the launch.
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous.
 *   - the caller owns every buffer (including workspaces).
 *   - activations are NHWC ("pixels x channels") strided views: channel stride
 *     is 1, the n/h/w strides are explicit so that channel slices of a concat
 *     buffer and the 2x2-strided phases of a transposed conv are views.
 *   - dtype: KS_F32 (parity mode, fp32 storage + fp32 FMA) or KS_BF16 (perf
 *     mode, bf16 storage, fp32 accumulate; tcgen05 tensor cores).
 */
#ifndef KUROSIWO_B200_H
#define KUROSIWO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KS_F32 0
#define KS_BF16 1

#define KS_OK 0
#define KS_EINVAL (-1)      /* bad argument */
#define KS_EUNSUPPORTED (-2) /* shape not handled by the requested implementation */
#define KS_EDRIVER (-3)     /* driver entry point (TMA descriptor encode) unavailable */

#define KS_IMPL_AUTO 0
#define KS_IMPL_SIMT 1  /* CUDA-core implicit GEMM (parity mode / odd shapes) */
#define KS_IMPL_TC 2    /* TMA + tcgen05 implicit GEMM (bf16 only) */

#define KS_MAX_VIEWS 8

/* Strided NHWC view; `ptr` addresses element (n=0,h=0,w=0,c=0) of the view. */
typedef struct ks_view {
  void *ptr;
  int64_t sn, sh, sw; /* strides in ELEMENTS; channel stride is 1 */
  int32_t C;          /* channels of this view */
  int32_t _pad;
} ks_view_t;

/* ---- library ---------------------------------------------------------- */
int ks_version(void);
const char *ks_error_string(int code);
/* Tuning/debug knobs: "tc_mt" (output windows per CTA), "tc_bo_mode", "tc_disable",
 * "wgrad_tc_disable", "tc_sa", "tc_sb" (pipeline depths); A/B switches that select the OLDER kernel of a pair (never needed for
 * correctness, every pair is parity-tested): "loss_variant" = 1 two-pass CE+Dice, "stem_simt" = 1 CUDA-core stem, "ecam_simt" = 1
 * CUDA-core ECAM classifier pass, "tc_stat_mode" = 1 shuffle-butterfly BatchNorm statistics everywhere, "att_no_umma" = 1 mma.sync ViT
 * attention, "dwconv_simple" = 1 / 2 one-output-per-thread / 2x2-register-block depth-wise conv kernels (default 0: shared-memory tiles
 * for bf16), "cf_scalar" = 1 scalar Dropout / DropPath kernels and 64-bit index arithmetic in im2col / col2im / the sigmoid head;
 * "xatt_umma" = 1 turns the tcgen05 forward of the ChangeFormer attention ON.  Returns KS_EINVAL for unknown names. */
int ks_set_option(const char *name, int value);
/* Every knob back to its default (0).  The knobs are process-global: harnesses that toggle them call this when they are done. */
int ks_reset_options(void);

/* ---- layout / precision plumbing --------------------------------------- */
/* dst[i0][i1][i2][i3] (contiguous, dtype dst_dtype) = src[i0*s0+i1*s1+i2*s2+i3*s3]
 * (dtype src_dtype).  Used for NCHW fp32 -> NHWC, OIHW -> [tap][Cout][Cin]
 * weight packing and the inverse gradient unpacking.
 * Replaces: `.to(device)` + cuDNN's internal layout transforms. */
int ks_permute_cast(int src_dtype, const void *src, int dst_dtype, void *dst,
                    int d0, int d1, int d2, int d3,
                    int64_t s0, int64_t s1, int64_t s2, int64_t s3,
                    int accumulate, void *stream);

/* One launch for a table of permute/cast jobs (all weight packs / gradient unpacks of a step).
 * jobs_dev: device array; chunks_dev: device int32 pairs {job index, first dst element}, one per CTA (4096 elements each). */
typedef struct ks_permute_job {
  const void *src;
  void *dst;
  int64_t total, s0, s1, s2, s3;
  int32_t d1, d2, d3, src_dtype, dst_dtype, accumulate;
  /* dst_strided != 0: element (i0,i1,i2,i3) goes to dst[i0*t0+i1*t1+i2*t2+i3*t3] instead of the contiguous position
   * (writes a logical sub-block into a channel-padded buffer whose padding stays zero). */
  int64_t t0, t1, t2, t3;
  int32_t dst_strided;
  float scale;   /* value multiplier; 0 means 1 (ResidualBlock's `* 0.1`, changeformer.py:481, is folded into packed weights) */
} ks_permute_job_t;
int ks_permute_cast_batched(const ks_permute_job_t *jobs_dev, const int32_t *chunks_dev, int n_chunks, void *stream);

/* ---- convolution engine ------------------------------------------------ */
/* out[n,h,w,co] (+)= bias[co] + sum_{tap,src,ci} x_src[n,h+dy,w+dx,ci] * w[tap][co][ci_global]
 * ksize 3: taps (dy,dx) in row-major order over {-1,0,1}^2, zero padding 1.
 * ksize 1: single tap.
 * srcs: K-dimension concat (torch.cat along C never materialised, snunet.py:132-144).
 * dsts: N-dimension split; dst_accumulate[i]!=0 -> read-modify-write (+=).
 * Used as forward conv (snunet.py:15-17 nn.Conv2d), as data-gradient conv with
 * tap-flipped/transposed weights, and for the four phases of
 * ConvTranspose2d(k=2,s=2) (snunet.py:41) and its data gradient.
 * weight: [taps][Cout_total][Cin_total] in `dtype`; bias fp32 [Cout_total] or NULL.
 * stats (optional, fp64 [2][Cout_total], pre-zeroed): per-channel sum and sum
 * of squares of the stored output (BatchNorm batch statistics, snunet.py:23,27). */
int ks_conv2d(int dtype, int N, int H, int W, int ksize,
              const ks_view_t *srcs, int n_src,
              const void *weight, const float *bias,
              const ks_view_t *dsts, int n_dst, const int *dst_accumulate,
              double *stats, int impl, void *stream);

/* dw[tap][co_global][ci_global] (+)= sum_{n,h,w} dy_j[n,h,w,co] * x_i[n,h+dy,w+dx,ci]
 * (fp32 output).  Replaces cuDNN wgrad for nn.Conv2d / nn.ConvTranspose2d. */
int ks_conv2d_wgrad(int dtype, int N, int H, int W, int ksize,
                    const ks_view_t *xs, int n_x,
                    const ks_view_t *dys, int n_dy,
                    float *dw, int accumulate, int impl, void *stream);

/* The same plus the bias gradient of the layer: dbias[(channel offset of the dy view + c) % bias_mod] (+)= sum_pixels dy[p][c]
 * (bias_mod = 0: no folding; the four 1x1 phase views of a ConvTranspose2d(k2, s2), snunet.py:41, share one bias: bias_mod = C).
 * On the tcgen05 path the sum rides along in the weight-gradient UMMAs when the last 128-row tile of Cin has a free channel-group
 * slot (a tile of ones as one more X group); otherwise the call runs ks_channel_sum per view.  Replaces: the bias branch of
 * cuDNN's ConvolutionBackward. */
int ks_conv2d_wgrad_bias(int dtype, int N, int H, int W, int ksize, const ks_view_t *xs, int n_x,
                         const ks_view_t *dys, int n_dy, float *dw, int accumulate, float *dbias, int bias_mod,
                         int accumulate_bias, int impl, void *stream);


/* Stem conv (models/snunet.py:75, conv0_0.conv1: Cin = 2 or 3, Cout = 32): reads the NCHW fp32 network
 * input and the OIHW fp32 master weight directly; NHWC output in `dtype`; optional BN statistics. */
int ks_stem_conv3x3(int dtype, int N, int Cin, int H, int W, const float *x_nchw, const float *w_oihw,
                    const float *bias, const ks_view_t *dst, double *stats, void *stream);
/* dw_oihw[co][ci][3][3] (+)= sum_px dy[px][co] * x[px+tap][ci]  (fp32, OIHW like the parameter). */
int ks_stem_wgrad3x3(int dtype, int N, int Cin, int H, int W, const float *x_nchw, const ks_view_t *dy,
                     float *dw_oihw, int accumulate, void *stream);

/* ---- BatchNorm (training mode) + ReLU + residual + pool ----------------- */
/* sums[0][c] += sum x, sums[1][c] += sum x^2 (fp64, caller zeroes). snunet.py:23,27 */
int ks_bn_stats(int dtype, int N, int H, int W, const ks_view_t *x, double *sums, void *stream);

/* mean/var from sums; scale=gamma*rstd, shift=beta-mean*scale; running stats
 * updated with `momentum` and the unbiased variance (nn.BatchNorm2d training). */
int ks_bn_finalize(int C, double count, const double *sums, const float *gamma,
                   const float *beta, float eps, float momentum,
                   float *running_mean, float *running_var,
                   float *scale, float *shift, float *mean, float *rstd, void *stream);

/* out = act(y*scale+shift (+res)); optional 2x2/s2 max-pooled copy to `pool`
 * (snunet.py:23-28 bn+relu(+identity), :73 MaxPool2d). H,W are y's dims. */
int ks_bn_act(int dtype, int N, int H, int W, const ks_view_t *y,
              const float *scale, const float *shift, const ks_view_t *res,
              int relu, const ks_view_t *out, const ks_view_t *pool, void *stream);

/* BN backward pass 1: g = dout * mask; sums[0][c] += sum g, sums[1][c] += sum g*xhat.
 * out != NULL: mask = (out > 0) and the masked gradient g is written back over `dout` (pass 2 and the
 *              identity path re-use it); out == NULL: mask = (y*scale+shift > 0), nothing is written.
 * dpool != NULL (needs out): the 2x2/s2 max-pool backward is folded in: g = (dout + dpool routed to the first
 *              maximum of each window of `out`) * (out > 0)   (aten max_pool2d_with_indices_backward). */
int ks_bn_bwd_reduce(int dtype, int N, int H, int W, const ks_view_t *dout,
                     const ks_view_t *out, const ks_view_t *y, const ks_view_t *dpool,
                     const float *scale, const float *shift,
                     const float *mean, const float *rstd, double *sums, void *stream);

/* BN backward pass 2: dy = gamma*rstd*(g' - sum_g/M - xhat*sum_gx/M) (+ add), with g' = g if `premasked`
 * else g*(y*scale+shift > 0).  dgamma = sum_gx, dbeta = sum_g, dsum_out = sum_g (the gradient of the
 * preceding conv bias along the identity path), all fp32 (+= if accumulate_param_grads). */
int ks_bn_bwd_apply(int dtype, int N, int H, int W, const ks_view_t *g, int premasked, const ks_view_t *y,
                    const float *scale, const float *shift, const float *mean, const float *rstd, const float *gamma,
                    const double *sums, double count, const ks_view_t *add, const ks_view_t *dy,
                    float *dgamma, float *dbeta, float *dsum_out, int accumulate_param_grads, void *stream);

/* dx[n,2h+i,2w+j,c] (+)= dpool[n,h,w,c] at the first max of each 2x2 window
 * (H,W are the POOLED dims). aten max_pool2d_with_indices_backward. */
int ks_maxpool2x2_bwd(int dtype, int N, int H, int W, const ks_view_t *x,
                      const ks_view_t *dpool, const ks_view_t *dx, int accumulate, void *stream);

/* out[c] (+)= sum_{n,h,w} x[n,h,w,c]  (conv bias gradients). */
int ks_channel_sum(int dtype, int N, int H, int W, const ks_view_t *x, float *out,
                   int accumulate, void *stream);

/* ---- ECAM head (snunet.py:49-62,146-151) ------------------------------- */
/* Global avg+max pool of cat(x_0..x_{J-1}) (J*Cb channels) and of intra=sum_j x_j
 * (Cb channels).  pooled: fp32 [N][2][(J+1)*Cb]  (avg | max; cat channels then intra);
 * argmax: int32 [N][(J+1)*Cb] flat pixel index of the first maximum - reset to INT_MAX here and FILLED IN by
 * ks_ecam_final (which reads the same tensors anyway and compares against the pooled maxima).
 * scratch: N*(J+1)*Cb 8-byte words, contents ignored. */
int ks_ecam_pool(int dtype, int N, int H, int W, const ks_view_t *xs, int J,
                 float *pooled, int *argmax, unsigned long long *scratch, void *stream);

/* gates: ca[N][J*Cb] = sigmoid(fc2(relu(fc1(avg))) + fc2(relu(fc1(max)))), same for ca1[N][Cb].
 * hidden: fp32 [N][2][hid+hid1] pre-ReLU hidden activations kept for backward. */
int ks_ecam_gates(int N, int Cb, int J, int hid, int hid1, const float *pooled,
                  const float *w_fc1, const float *w_fc2, const float *w1_fc1, const float *w1_fc2,
                  float *gates, float *hidden, void *stream);

/* logits[n,k,h,w] = bf[k] + sum_c wf[k][c] * ca[n,c]*(x[c] + ca1[n, c%Cb])   (NCHW fp32 out).
 * pooled/argmax (both or neither): the buffers of ks_ecam_pool; argmax[n][c] = min(argmax, first pixel whose value
 * equals the pooled maximum) - what aten's adaptive_max_pool2d backward routes the gradient to. */
int ks_ecam_final(int dtype, int N, int H, int W, const ks_view_t *xs, int J,
                  const float *gates, const float *wf, const float *bf, int K,
                  float *logits, const float *pooled, int *argmax, void *stream);

/* Pixel reductions of the head backward (callee zeroes `red`):
 * red: fp64 [N][K*J*Cb + K] = { B[k][c] = sum_px dlogits[k]*x[c] , D[k] = sum_px dlogits[k] }. */
int ks_ecam_bwd_reduce(int dtype, int N, int H, int W, const ks_view_t *xs, int J, int K,
                       const float *dlogits, double *red, void *stream);

/* Backward of ks_ecam_gates and of the classifier weights: dpooled fp32 [N][2][(J+1)*Cb];
 * dwf [K][J*Cb], dbf [K] and the fc weight grads are fp32 (+= if accumulate). */
int ks_ecam_gates_bwd(int N, int Cb, int J, int hid, int hid1, int K, const float *pooled,
                      const float *hidden, const float *gates, const double *red, const float *wf,
                      const float *w_fc1, const float *w_fc2, const float *w1_fc1, const float *w1_fc2,
                      float *dpooled, float *dwf, float *dbf, float *dw_fc1, float *dw_fc2,
                      float *dw1_fc1, float *dw1_fc2, int accumulate, void *stream);

/* dx_j = ca*dO + avg/max-pool and intra gradients (assign). */
int ks_ecam_bwd_apply(int dtype, int N, int H, int W, int J, int Cb,
                      const float *gates, const float *wf, int K, const float *dlogits,
                      const float *dpooled, const int *argmax, const ks_view_t *dxs, void *stream);

/* ---- Siamese U-Net passes (models/siam_conc.py, models/siam_diff.py) ------ */
/* nn.Softmax(dim=1) (siam_conc.py:93,177) / nn.LogSoftmax(dim=1) (siam_diff.py:93,173) over the first K channels of the
 * NHWC view z (the conv engine pads the classifier's Cout to 16); out: NCHW fp32 [N][K][H][W]. */
int ks_softmax_head_fwd(int dtype, int N, int H, int W, const ks_view_t *z, int K, int log_mode,
                        float *out, void *stream);
/* dz[.., k] = softmax / log-softmax backward from the saved output `out` and its gradient `dout` (both NCHW fp32);
 * channels K..dz->C-1 are written as 0. */
int ks_softmax_head_bwd(int dtype, int N, int H, int W, const float *out, const float *dout, int K,
                        int log_mode, const ks_view_t *dz, void *stream);
/* nn.Dropout2d masks (siam_conc.py:21...): mask[i] = 1/(1-p) with probability 1-p, else 0, i in [0, n) (n = sum over
 * layers of N*C).  The stream is a pure function of (seed, *step_ptr, i); step_ptr is a DEVICE counter so that a
 * captured CUDA graph draws fresh masks on every replay. */
int ks_dropout_mask(float *mask, int64_t n, float p, uint64_t seed, const int *step_ptr, void *stream);
/* x[n,h,w,c] *= m[n*C + c] in place: Dropout2d forward on activations, backward on gradients. */
int ks_channel_scale(int dtype, int N, int H, int W, const ks_view_t *x, const float *m, void *stream);
/* out = |a - b|  (siam_diff.py:141,150,158,165: torch.abs(x_1 - x_2) skip connections). */
int ks_absdiff_fwd(int dtype, int N, int H, int W, const ks_view_t *a, const ks_view_t *b,
                   const ks_view_t *out, void *stream);
/* da (+)= sign(a-b)*g, db (+)= -sign(a-b)*g  (sign(0) = 0 as in aten). */
int ks_absdiff_bwd(int dtype, int N, int H, int W, const ks_view_t *a, const ks_view_t *b, const ks_view_t *g,
                   const ks_view_t *da, int accumulate_a, const ks_view_t *db, int accumulate_b, void *stream);

/* ---- ViT encoder / FloodViT head passes (models/vision_transformer.py, models/model_utilities.py:36-94) ------------
 * Token matrices are [rows = B*Tp, C] row-major in `dtype` (Tp = tokens per image padded to a multiple of 16; T valid:
 * cls + patches); the Linear layers run on ks_conv2d / ks_conv2d_wgrad as 1x1 convolutions over the same buffers. */
/* nn.LayerNorm over the last dim (vision_transformer.py:22,43,72,123,125): y = (x-mean)*rstd*gamma+beta, biased variance,
 * two-pass fp32 statistics; mean/rstd (fp32 [rows], optional) are kept for the backward; copy_out (optional) receives x
 * unchanged (the next buffer of the residual stream, which the following GEMM accumulates into).  C % 8 == 0, C <= 2048. */
int ks_layernorm_fwd(int dtype, int64_t rows, int C, const void *x, int64_t ldx, const float *gamma, const float *beta,
                     float eps, void *y, int64_t ldy, float *mean, float *rstd, void *copy_out, int64_t ldc, void *stream);
/* dx (+)= rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma; dgamma += sum dy*xhat, dbeta += sum dy (fp32 atomics, caller
 * zeroes).  dx may be NULL (parameter gradients only).  C % 8 == 0, C <= 1024. */
int ks_layernorm_bwd(int dtype, int64_t rows, int C, const void *dy, int64_t lddy, const void *x, int64_t ldx,
                     const float *mean, const float *rstd, const float *gamma, void *dx, int64_t lddx, int accumulate_dx,
                     float *dgamma, float *dbeta, void *stream);
/* Rearrange('b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p=16) + LayerNorm(256*Cc) (vision_transformer.py:122-123) straight
 * from the NCHW fp32 image: patch p of image b -> row b*Tp + 1 + p of out [B*Tp][256*Cc] (row b*Tp is the cls slot; rows the
 * kernel does not write stay as the caller initialised them: zero).  mean/rstd: fp32 [B*Tp]. */
int ks_patchify_ln(int dtype, int B, int Cc, int Hi, int Wi, int Tp, const float *img, const float *gamma, const float *beta,
                   float eps, void *out, float *mean, float *rstd, void *stream);
/* dgamma += sum_patches dy*xhat, dbeta += sum_patches dy (the image itself needs no gradient). */
int ks_patchify_ln_bwd(int dtype, int B, int Cc, int Hi, int Wi, int Tp, const float *img, const float *mean, const float *rstd,
                       const void *dy, float *dgamma, float *dbeta, void *stream);
/* x0 = cat(cls_token, e[:,1:T]) + pos_embedding[:T] (vision_transformer.py:142-146); padding rows t >= T are written as 0. */
int ks_vit_assemble(int dtype, int B, int T, int Tp, int D, const void *e, const float *cls, const float *pos, void *x0, void *stream);
/* dpos[t] = sum_b dx0[b,t], dcls = sum_b dx0[b,0] (assigned), de[b,t] = dx0[b,t] for 1 <= t < T and 0 elsewhere. */
int ks_vit_assemble_bwd(int dtype, int B, int T, int Tp, int D, const void *dx0, void *de, float *dcls, float *dpos, void *stream);
/* Multi-head attention core (vision_transformer.py:56-65): qkv [B*Tp][3*heads*dh] = (q | k | v), head-major inside each third;
 * out [B*Tp][heads*dh] = softmax(q k^T * scale) v over the T valid keys; probs [B][heads][Tp][Tp] keeps the probabilities for
 * the backward (rows/columns >= T are 0).  dh == 64, Tp <= 256. */
int ks_attention_fwd(int dtype, int B, int T, int Tp, int heads, int dh, const void *qkv, float scale, void *out, void *probs,
                     void *stream);
/* dqkv from dout: dV = P^T dO, dS = P*(dO V^T - rowsum(dO V^T * P)), dQ = scale dS K, dK = scale dS^T Q.
 * ds_scratch: [B][heads][Tp][Tp] in `dtype`.  Padding rows of dqkv are written as 0. */
int ks_attention_bwd(int dtype, int B, int T, int Tp, int heads, int dh, const void *qkv, const void *probs, const void *dout,
                     float scale, void *dqkv, void *ds_scratch, void *stream);
/* nn.GELU() (exact erf; vision_transformer.py:25): h = gelu(u); du = dh * gelu'(u).  n % 8 == 0. */
int ks_gelu_fwd(int dtype, int64_t n, const void *u, void *h, void *stream);
int ks_gelu_bwd(int dtype, int64_t n, const void *u, const void *dh, void *du, void *stream);
/* nn.Upsample(size=(Ho,Wo), mode='bilinear') (align_corners=False; model_utilities.py:89-91) of the K-channel G x G token map
 * held in rows b*Tp + row0 + gy*G + gx (channels 0..K-1 of Cs-wide rows) -> NCHW fp32 [B][K][Ho][Wo]; the linear 1x1 head
 * commutes with the interpolation, so it runs on the G x G grid first.  _bwd is the exact adjoint (gather form): every
 * element of dsrc [B*Tp][Cs] is written (0 outside the grid rows / channels >= K). */
int ks_bilinear_up_fwd(int dtype, int B, int G, int Tp, int row0, int Cs, int K, int Ho, int Wo, const void *src, float *dst,
                       void *stream);
int ks_bilinear_up_bwd(int dtype, int B, int G, int Tp, int row0, int Cs, int K, int Ho, int Wo, const float *ddst, void *dsrc,
                       void *stream);

/* nn.AdaptiveAvgPool2d(S) of an NHWC view -> dense [N][S][S][C] (UPerNet pyramid pooling; HF transformers modeling_upernet.py
 * UperNetPyramidPoolingBlock, called through models/upernet.py:80): bin (i,j) = mean over rows [floor(i*H/S), ceil((i+1)*H/S)) x
 * cols [floor(j*W/S), ceil((j+1)*W/S)).  _bwd: dsrc (+)= adjoint(ddst). */
int ks_adaptive_avgpool_fwd(int dtype, int N, int H, int W, int S, const ks_view_t *src, void *dst, void *stream);
int ks_adaptive_avgpool_bwd(int dtype, int N, int H, int W, int S, const void *ddst, const ks_view_t *dsrc, int accumulate, void *stream);

/* ---- ChangeFormerV6 passes (models/changeformer.py) ---------------------------------------------------------------
 * Token matrices [B*N, C] are NHWC images [B, H, W, C]; LayerNorm / GELU / Linear reuse the ViT entry points above, the decoder's
 * 3x3 / 1x1 / transposed convolutions and BatchNorm reuse ks_conv2d / ks_conv2d_wgrad / ks_bn_*. */
/* Generic strided convolution (OverlapPatchEmbed.proj 7x7 s4|s2 p3, changeformer.py:281; Attention.sr k = s = sr, :166):
 * out[n,ho,wo,co] = bias[co] + sum x[n, ho*s-p+ky, wo*s-p+kx, ci] * w[ky*k+kx][co][ci]; weight in `dtype`, k <= 8. */
int ks_conv2d_strided(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *src,
                      const void *weight, const float *bias, const ks_view_t *dst, void *stream);
/* dx (+)= its data gradient (same weight layout); dw[ky*k+kx][co][ci] (+)= its weight gradient (fp32). */
int ks_conv2d_strided_dgrad(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *dy,
                            const void *weight, const ks_view_t *dx, int accumulate, void *stream);
int ks_conv2d_strided_wgrad(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *x,
                            const ks_view_t *dy, float *dw, int accumulate, void *stream);
/* Spatial-reduction attention core (changeformer.py:186-208): q [B*Nq][ldq] (head h at columns h*dh..), kv [B*Nk][ldkv] with k at
 * columns h*dh.. and v at heads*dh + h*dh..; out = softmax(q k^T * scale) v; probs [B][heads][Nq][Nk] kept for the backward.
 * Nk <= 64, dh <= 96.  _bwd writes dq and ACCUMULATES dk|dv into the fp32 buffer dkv [B*Nk][2*heads*dh] (zeroed by the call). */
int ks_xattention_fwd(int dtype, int B, int Nq, int Nk, int heads, int dh, const void *q, int64_t ldq, const void *kv, int64_t ldkv,
                      float scale, void *out, int64_t ldo, void *probs, float pdrop, uint64_t seed, const int *step_ptr, int site,
                      void *stream);
int ks_xattention_bwd(int dtype, int B, int Nq, int Nk, int heads, int dh, const void *q, int64_t ldq, const void *kv, int64_t ldkv,
                      const void *probs, const void *dout, int64_t ldo, float scale, void *dq, int64_t lddq, float *dkv,
                      float pdrop, uint64_t seed, const int *step_ptr, int site, void *stream);
/* Stochastic regularisers of the ChangeFormer encoder (changeformer.py:652-654).  Every mask is a pure function of
 * (seed, *step_ptr, site, element index) - nothing is stored, the backward regenerates it; step_ptr is a DEVICE counter so a
 * captured CUDA graph draws fresh masks per replay.  pdrop of ks_xattention_* is attn_drop (:203) on the probabilities used for P.V.
 * ks_dropout_apply: y = x * keep/(1-p)                                   nn.Dropout after GELU (:129-130) and its backward
 * ks_branch_add:    x += droppath[sample] * keep/(1-p) * t               x = x + drop_path(dropout(branch)) (:246-247; :131-132,206)
 * ks_branch_scale:  dt = droppath[sample] * keep/(1-p) * dx              its backward (gradient of the branch output)
 * droppath: fp32 [samples] factors 0 or 1/(1-p_path) from ks_dropout_mask (may be NULL); per_sample = elements per sample. */
int ks_dropout_apply(int dtype, int64_t n, const void *x, void *y, float p, uint64_t seed, const int *step_ptr, int site, void *stream);
int ks_branch_add(int dtype, int64_t n, int64_t per_sample, void *x, const void *t, float p, const float *droppath, uint64_t seed,
                  const int *step_ptr, int site, void *stream);
int ks_branch_scale(int dtype, int64_t n, int64_t per_sample, const void *dx, void *dt, float p, const float *droppath, uint64_t seed,
                    const int *step_ptr, int site, void *stream);
/* The strided convolutions as GEMMs on the tensor-core conv engine (OverlapPatchEmbed.proj, changeformer.py:281-290; Attention.sr, :166,193):
 * ks_im2col: col[(n,ho,wo)][(u*k+v)*C + c] = x[n, ho*s-p+u, wo*s-p+v, c] (0 outside the image); col is a dense [N*Ho*Wo, Kp] matrix,
 * Kp >= k*k*C, columns >= k*k*C are left untouched (the caller zeroes them once).  ks_col2im is the adjoint in gather form:
 * dx[n,h,w,c] (+)= sum of dcol over the windows containing (h,w).  Then forward = 1x1 ks_conv2d(col, W[Cout][Kp]), weight gradient =
 * 1x1 ks_conv2d_wgrad(col, dy), data gradient = ks_col2im(1x1 ks_conv2d(dy, W^T)). */
int ks_im2col(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const ks_view_t *x, void *col, int Kp,
              void *stream);
int ks_col2im(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int ksize, int stride, int pad, const void *dcol, int Kp, const ks_view_t *dx,
              int accumulate, void *stream);
/* Depth-wise 3x3 conv, padding 1 (DWConv, changeformer.py:84-96) on dense NHWC; w9: fp32 [9][C] (tap-major), bias fp32 [C].
 * _bwd: dx = data gradient; dw9 += weight gradient, dbias += bias gradient (fp32 atomics, caller zeroes).  y / dx are bit-identical
 * across the kernel selections of "dwconv_simple" (bias, then nine fmaf in tap order). */
int ks_dwconv3x3_fwd(int dtype, int N, int H, int W, int C, const void *x, const float *w9, const float *bias, void *y, void *stream);
int ks_dwconv3x3_bwd(int dtype, int N, int H, int W, int C, const void *x, const void *dy, const float *w9, void *dx, float *dw9,
                     float *dbias, void *stream);
/* F.interpolate(mode='bilinear', align_corners=False) of dense NHWC maps (changeformer.py:587,592,600,608): dst (+)= resize(src);
 * _bwd is the exact adjoint in gather form: dsrc (+)= resize^T(ddst). */
int ks_bilinear_nhwc_fwd(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int C, const void *src, void *dst, int accumulate, void *stream);
int ks_bilinear_nhwc_bwd(int dtype, int N, int Hi, int Wi, int Ho, int Wo, int C, const void *ddst, void *dsrc, int accumulate, void *stream);
/* ReLU outside a BatchNorm pass (conv_diff / ResidualBlock, changeformer.py:31-38,471-483): y = max(x, 0) (y may alias x);
 * dx = g * (r > 0) with r the ReLU OUTPUT (dx may alias g). */
int ks_relu_fwd(int dtype, int64_t n, const void *x, void *y, void *stream);
int ks_relu_bwd(int dtype, int64_t n, const void *r, const void *g, void *dx, void *stream);
/* nn.Sigmoid on the K-class map (changeformer.py:635-639): NHWC view z (Cout padded to 16) -> NCHW fp32; _bwd writes every
 * channel of dz (0 for channels >= K). */
int ks_sigmoid_head_fwd(int dtype, int N, int H, int W, const ks_view_t *z, int K, float *out, void *stream);
int ks_sigmoid_head_bwd(int dtype, int N, int H, int W, const float *out, const float *dout, int K, const ks_view_t *dz, void *stream);

/* ---- loss (utilities/bce_and_dice.py:18-24, utilities/dice.py:93-137) --- */
/* Fused softmax -> weighted CE(ignore_index) + Dice, forward + gradient + argmax.
 * logits: NCHW fp32 [N][C][HW] (C==3); labels int64 [N][HW].
 * loss_out: fp32[3] = {total, dice, ce}; dlogits NCHW fp32 scaled by grad_scale
 * (may be NULL: forward only); pred: uint8 [N][HW] argmax (may be NULL).
 * workspace: ks_ce_dice_workspace_bytes(N) bytes, one per batch size N, ZERO-FILLED ONCE by the caller before its first use; every
 * call leaves it ready for the next one (the kernels re-zero what they used: no memset launch per call).
 * Up to 148 x 11 chunks of 2048 pixels (bs = 64 at 224 x 224) with HW % 4 == 0 and dlogits != NULL run as ONE resident pass: one CTA per SM
 * keeps its logits on the chip across a single grid barrier (bounded spin; on a time-out loss_out becomes NaN instead of a hang), so the
 * call must not share the GPU with a kernel that waits on it.  Other shapes / forward-only calls take two streaming passes. */
int64_t ks_ce_dice_workspace_bytes(int N);
int ks_ce_dice_fwd_bwd(const float *logits, const int64_t *labels, int N, int C, int64_t HW,
                       const float *class_weights, int ignore_index, float grad_scale,
                       float *loss_out, float *dlogits, uint8_t *pred,
                       void *workspace, void *stream);

/* Same kernel with the Dice term weighted by dice_weight: loss = dice_weight*Dice + CE.  dice_weight = 0 is exactly
 * nn.CrossEntropyLoss(weight=class_weights, ignore_index) - the reference's DEFAULT criterion (configs/train/train_config.json:11,
 * utilities/utilities.py:308-321); loss_out[1] still reports the Dice value. */
int ks_ce_dice_fwd_bwd_ex(const float *logits, const int64_t *labels, int N, int C, int64_t HW,
                          const float *class_weights, int ignore_index, float grad_scale, float dice_weight,
                          float *loss_out, float *dlogits, uint8_t *pred,
                          void *workspace, void *stream);

/* In-loop metrics: mat[target][pred] += 1 (int64 [4][4], row-major) over the n pixels whose target != ignore_index; replaces
 * the 5 torchmetrics objects of utilities/utilities.py:228-265 updated at change_detection_trainer.py:184-199 (accuracy, F1,
 * precision, recall, IoU per class all derive from the confusion matrix).  pred: the uint8 argmax map of ks_ce_dice_fwd_bwd. */
int ks_confusion_update(const uint8_t *pred, const int64_t *labels, int64_t n, int num_classes_with_ignore, int ignore_index,
                        int64_t *mat, void *stream);

/* Same counts routed additionally by per-sample keys (one launch for the global, the per-activation/AOI and the per-climate-zone
 * metric sets of change_detection_trainer.py:184-199, :445-472 and segmentation_trainer.py:407-512): sample s adds its KxK counts
 * to mat, to mat_a[key_a[s]] and to mat_b[key_b[s]] (keys outside [0, n) are skipped; any of mat / mat_a / mat_b may be NULL).
 * The water-only F-score (evaluate_water, :408-413) is the same matrix with classes 1 and 2 merged. */
int ks_confusion_update_grouped(const uint8_t *pred, const int64_t *labels, int n_samples, int64_t per_sample,
                                int num_classes_with_ignore, int ignore_index, const int32_t *key_a, int n_a,
                                const int32_t *key_b, int n_b, int64_t *mat, int64_t *mat_a, int64_t *mat_b, void *stream);

/* ---- input pipeline (dataset/Dataset.py:162-168 clamp + nan_to_num, :192-198 Normalize) on raw float32 SAR planes [B][C][HW];
 * out may alias raw.  clamp_max > 0: v = NaN ? clamp_max : min(max(v, 0), clamp_max); clamp_max <= 0: NaN -> 200, +-inf -> +-FLT_MAX.
 * Then (v - mean[c]) / std[c] with IEEE subtraction and division: bit-identical to the reference's torch ops. */
int ks_sar_preprocess(const float *raw, float *out, int B, int C, int64_t HW, const float *mean, const float *stdv,
                      float clamp_max, void *stream);

/* ---- optimizer (torch.optim.Adam, change_detection_trainer.py:52-54) ---- */
/* step_ptr: device int32 counter, incremented by the kernel (graph-capturable). */
int ks_adam_step(float *p, const float *g, float *m, float *v, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay,
                 float grad_scale, int *step_ptr, void *stream);

/* torch.optim.AdamW(lr, betas, weight_decay) (change_detection_trainer.py:55-60): decoupled decay p *= 1 - lr*wd, then the Adam
 * update on grad_scale*g. Same state layout as ks_adam_step. */
int ks_adamw_step(float *p, const float *g, float *m, float *v, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay,
                  float grad_scale, int *step_ptr, void *stream);

/* torch.optim.SGD(lr, momentum, weight_decay) (change_detection_trainer.py:61-66; ChangeFormer: momentum 0.99, wd 1e-5,
 * configs/method/changeformer/changeformer.json): g = grad_scale*g + wd*p; buf = momentum*buf + g; p -= lr*buf. */
int ks_sgd_step(float *p, const float *g, float *buf, int64_t n, float lr, float momentum, float weight_decay, float grad_scale,
                void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KUROSIWO_B200_H */
