/*
 * ganb200.h — C ABI of libganb200.so: the B200 (sm_100a) kernels behind the layer-op surface of
 * watsonyanghx/GAN_Lib_Tensorflow (common/ops/*.py, common/resnet_block.py).
 *
 * The reference has no FFI: its arithmetic lives in TensorFlow-1.5 library ops called from Python
 * (SURVEY.md 8(b)).  Each entry point below names the reference call-site whose arithmetic it replaces.
 * A maintainer of the reference binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; nothing is allocated inside the library;
 *   - activations are NHWC, filters HWIO, linear weights [in,out] exactly as in the reference;
 *   - "bf16" buffers are raw __nv_bfloat16, "f32" buffers are float;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant across streams;
 *   - return value: 0 on success, <0 one of GANB_E_*; ganb_last_error() gives a thread-local message;
 *   - no C++ exception crosses this boundary.
 */
#ifndef GANB200_H_
#define GANB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GANB_ABI_VERSION 1

#define GANB_OK 0
#define GANB_E_BADARG (-1)
#define GANB_E_UNSUPPORTED (-2)
#define GANB_E_ARCH (-3)
#define GANB_E_LAUNCH (-4)

/* activation codes shared by several entry points */
#define GANB_ACT_NONE 0
#define GANB_ACT_RELU 1  /* tf.nn.relu, common/resnet_block.py:25-26 */
#define GANB_ACT_LRELU 2 /* tf.maximum(x, 0.2x), common/resnet_block.py:27-29 */
#define GANB_ACT_TANH 3  /* tf.tanh, SNGAN/gan_cifar_resnet.py:261 */

/* output dtype codes */
#define GANB_F32 0
#define GANB_BF16 1

const char* ganb_last_error(void);
int ganb_abi_version(void);
int ganb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of CUDA kernels this library has launched in the calling process (launches recorded into a CUDA graph
 * count once, at capture time) */
int64_t ganb_launch_count(void);
/* Limits the SMs that the persistent tensor-core kernels launched by the CALLING THREAD may occupy (their grids, split
 * counts and workspace sizes follow); 0 = all.  Two independent passes issued on two streams under complementary limits
 * run side by side (the critic step next to the generator's forward pass).  Returns the previous limit. */
int ganb_set_sm_limit(int sms);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core convolution (tcgen05 / TMEM implicit GEMM, TMA-fed, BF16 inputs, FP32 accumulation).
 * Replaces tf.nn.conv2d (common/ops/conv2d.py:181-187, conv2d_.py:138-152), its autodiff
 * (Conv2DBackpropInput / Conv2DBackpropFilter, reached through tf.gradients at
 * SNGAN/gan_cifar_resnet.py:523-524), tf.nn.conv2d_transpose (common/ops/deconv2d.py:102-109) and
 * tf.matmul for linear layers viewed as 1x1 convolutions (common/ops/linear.py:163-173).
 *
 * ganb_conv2d_igemm computes, for every output pixel (n,ho,wo) and channel co,
 *     y = act( alpha * sum_{r,s,ci} x[n, ho*stride + r - pad_t, wo*stride + s - pad_l, ci] * wp[t(r,s)][co][ci]
 *              + bias[co] + residual[n,ho,wo,co] )
 * with zero padding outside the input.  `wp` is a packed bf16 filter [kh*kw][cout][cin] (cin contiguous).
 * flip_taps=0 uses t = r*kw+s (forward convolution); flip_taps=1 uses t = kh*kw-1-(r*kw+s), which turns
 * the same kernel into the data gradient of a stride-1 convolution when x := dy, wp := HWIO filter viewed
 * as [kh*kw][cin_fwd][cout_fwd], pad := k-1-pad.
 * Requirements: cin % 8 == 0, x and wp 16-byte aligned. alpha (device scalar), bias, residual may be NULL.
 * Outputs (and fp32 residuals) whose base address and row length are multiples of 32 bytes are written (read) with
 * 256-bit accesses, one whole sector per instruction; anything else takes the 16- / 8- / 2-byte forms.
 * stride in 1..4 (strided layers gather through TMA element strides; the data gradient of a strided convolution is
 * this same entry applied to the zero-dilated output gradient, see ganb_dilate2d).
 * residual_up2 = 1: `residual` is [n, ho/2, wo/2, cout] and is read through a nearest-neighbour 2x upsample
 * (the shortcut of an 'up' residual block, common/resnet_block.py:123-127, without materialising the upsampled map).
 * ---------------------------------------------------------------------------------------------- */
int ganb_conv2d_igemm(const void* x_bf16, const void* wp_bf16, void* y, int n, int h, int w, int cin, int ho,
                      int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps,
                      const float* alpha, const float* bias, const float* residual, int residual_up2, int act,
                      int out_dtype, void* stream);
/* Data gradient THROUGH an activation that the producer of this layer's input applied in its own epilogue
 * (ganb_conv2d_igemm with act != 0): the critic's residual blocks have no normalisation between Conv1 and Conv2
 * (common/resnet_block.py:129-139 with Normalize = identity, SNGAN/gan_cifar_resnet.py:88-109), so
 * nonlinearity(Conv1(.)) is stored once, by Conv1, and its derivative is applied here instead of by a pass of its own:
 *     y = act'(pre) * alpha * sum x * wp      (same sum as ganb_conv2d_igemm; no bias / residual)
 * gate_bf16 = act(pre), laid out like y ([n, ho, wo, cout] bf16); act'(pre) is read off its sign: relu 1 if gate > 0
 * else 0, leaky relu 1 if gate >= 0 else 0.2.  Requirements of ganb_conv2d_igemm plus cout % 8 == 0. */
int ganb_conv2d_igemm_gated(const void* x_bf16, const void* wp_bf16, void* y, int n, int h, int w, int cin, int ho,
                            int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps,
                            const float* alpha, const void* gate_bf16, int gate_act, int out_dtype, void* stream);
/* Batch statistics fused into the convolution epilogue (the reference computes them with a separate tf.nn.moments over
 * the layer output, common/ops/normalization.py:29,47): the epilogue leaves, per 128-pixel output tile, the column sums
 * of the STORED output y and of y^2 (after alpha / bias / residual / activation and the rounding to out_dtype) in
 *   stats [groups][rows][2][cout]   (fp32; rows = ganb_conv2d_stats_rows(...), tiles are image-major, so the tiles of
 *                                    one statistic tower are contiguous)
 * and ganb_bn_stats_finalize folds the rows in a fixed order (deterministic) into mean / rstd [groups][cout] -- the same
 * quantities ganb_bn_stats returns, without re-reading y.  ganb_conv2d_stats_rows returns 0 when the layer cannot do it
 * (cout % 32 != 0, or a pixel tile would straddle two towers): callers then use ganb_bn_stats. */
int ganb_conv2d_stats_rows(int n, int ho, int wo, int cout, int kh, int kw, int stride, int groups);
int ganb_conv2d_igemm_stats(const void* x_bf16, const void* wp_bf16, void* y, int n, int h, int w, int cin, int ho,
                            int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps,
                            const float* alpha, const float* bias, const float* residual, int residual_up2, int act,
                            int out_dtype, float* stats, int groups, void* stream);
int ganb_bn_stats_finalize(const float* partial, int c, int groups, int chunks, int64_t count, float eps, float* mean,
                           float* rstd, void* stream);

/* Filter gradient: partial[split][t][ci][co] = sum over the split's pixels of
 *     x[n, ho*stride + r - pad_t, wo*stride + s - pad_l, ci] * dy[n, ho, wo, co]
 * `workspace` (32-byte aligned) must hold ganb_conv2d_wgrad_workspace() bytes; the reduction over splits, the optional
 * scale and the accumulation into dw (HWIO f32) are done by the same call (second kernel).
 *     dw = beta * dw + scale * sum_split partial
 * Requirements: cin % 8 == 0, cout % 8 == 0. */
int64_t ganb_conv2d_wgrad_workspace(int n, int h, int w, int cin, int ho, int wo, int cout, int kh, int kw);
int ganb_conv2d_wgrad(const void* x_bf16, const void* dy_bf16, float* dw, void* workspace, int n, int h, int w,
                      int cin, int ho, int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l,
                      const float* scale, float beta, void* stream);


/* CUDA-core convolutions for layers with <= 8 channels on one side (RGB side of D.Block.1.*, G.Output):
 * same arithmetic as ganb_conv2d_igemm / wgrad, bandwidth-bound on the large-channel tensor.
 * Reference shapes: SNGAN/gan_cifar_resnet.py:212-234 (input_dim=3), :260 (output_dim=3).
 *
 * smallcin: y[n,ho,wo,cl] = act(alpha * sum x[n,ho+r-pad_t,wo+s-pad_l,cs] * w(t,cs,cl) + bias[cl]), x fp32 with
 *   cs <= 8 channels; w is fp32 [tap][cs][cl] (w_layout_clcs=0) or [tap][cl][cs] (=1); flip_taps as in igemm.
 * small_wgrad: dw = beta*dw + scale * sum_q xs[q + sign*(tap offset)][cs] * yl[q][cl];
 *   sign=+1: xs = conv input, yl = output gradient; sign=-1: yl = conv input, xs = output gradient;
 *   dw layout [tap][cs][cl] (out_layout_clcs=0) or [tap][cl][cs] (=1). taps*cs <= 27. */
int ganb_conv2d_smallcin(const float* x, const float* w, void* y, int n, int h, int w_in, int cs, int ho, int wo,
                         int cl, int kh, int kw, int pad_t, int pad_l, int flip_taps, int w_layout_clcs,
                         const float* alpha, const float* bias, int act, int out_dtype, void* stream);
int64_t ganb_conv2d_small_wgrad_workspace(int n, int hl, int wl, int cs, int cl, int kh, int kw);
int ganb_conv2d_small_wgrad(const float* xs, const void* yl, int yl_dtype, float* dw, void* workspace, int n, int hs,
                            int ws, int cs, int hl, int wl, int cl, int kh, int kw, int pad_t, int pad_l, int sign,
                            int out_layout_clcs, const float* scale, float beta, void* stream);

/* Tensor-core route for the same <=8-channel layers: bf16 im2col of the small tensor (kh*kw*cs <= kpad columns)
 * makes fprop / wgrad / dgrad 1x1 GEMMs for ganb_conv2d_igemm / ganb_conv2d_wgrad.
 *   im2col_small : out[(n,ho,wo)][(r*kw+s)*cs + c] = xs[n, ho*stride + sign*(r-pad_t), wo*stride + sign*(s-pad_l), c], zero padded
 *   pack_small   : out[l][tap*cs + c] (row stride kpad): small_is_ci ? W[tap][c][l] : W[tap][l][c]
 *   small_wgrad_scatter : dw[tap][cs][cl] (or [tap][cl][cs]) = beta*dw + scale * r[tap*cs + c][l], r = [kpad][cl] */
int ganb_im2col_small(const float* xs, void* out_bf16, int n, int hs, int ws, int cs, int ho, int wo, int kh, int kw,
                      int stride, int pad_t, int pad_l, int sign, int kpad, void* stream);
int ganb_pack_small(const float* w_hwio, void* out_bf16, int taps, int ci, int co, int small_is_ci, int kpad,
                    void* stream);
int ganb_small_wgrad_scatter(const float* r, float* dw, int taps, int cs, int cl, int out_layout_clcs,
                             const float* scale, float beta, void* stream);

/* Tiny dense layers (tf.matmul, common/ops/linear.py:163-173, for shapes the TMA path cannot take):
 * C[m,n] = beta*C + alpha * op(A)[m,k] * op(B)[k,n] + bias[n]; trans_a: A stored [k,m]; trans_b: B stored [n,k]. */
int ganb_sgemm_small(const float* a, const float* b, float* c, int m, int n, int k, int trans_a, int trans_b,
                     const float* alpha, const float* bias, float beta, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral normalisation, common/ops/sn.py:15-69 (power iteration through tf.while_loop + 4 tf.matmul,
 * and its autodiff: the reference has no stop_gradient).  Grouped: one launch handles `count` weights
 * described by an array of ganb_sn_layer that lives in DEVICE memory.
 *   fwd : v = l2n(W u), b = W^T v, u_out = l2n(b), scal = {sigma, 1/sigma, |W u|, |b|}     (W is [k, c])
 *   bwd : dw += g/sigma + v (x) bbar + abar (x) u_used with g = dL/d(W/sigma)
 * ---------------------------------------------------------------------------------------------- */
#define GANB_SN_ROWS 16   /* rows of the [k, c] weight matrix handled by one CTA of the grouped launches */
typedef struct ganb_sn_layer {
  const float* w;   /* [k, c] fp32 weight (HWIO filter flattened to [kh*kw*cin, cout], or linear [in,out]) */
  float* u;         /* [c]  persistent power-iteration vector (sn.py:32); overwritten with u_out when assign=1 */
  float* u_out;     /* [c]  u after one iteration */
  float* u_used;    /* [c]  the u this evaluation started from (needed by the backward pass) */
  float* v;         /* [k]  */
  float* b;         /* [c]  W^T v before normalisation */
  float* scal;      /* [8]  sigma, 1/sigma, |W u|, |b|, (bwd scratch: coef, v.t), -, - */
  const float* g;   /* bwd only: [k, c] gradient w.r.t. W/sigma */
  float* dw;        /* bwd only: [k, c] accumulated gradient w.r.t. W */
  float* t;         /* [k]  bwd scratch (W b) */
  float* work;      /* [ceil(k/GANB_SN_ROWS) * (c + 4)] per-CTA partials */
  int32_t k, c;
  int32_t blk_begin; /* prefix sum of ceil(k/GANB_SN_ROWS) over the preceding layers */
  int32_t pad_;
} ganb_sn_layer;
/* assign=1 reproduces update_collection=None (u.assign(u_final) on every evaluation, sn.py:48-56);
 * assign=0 reproduces update_collection="NO_OPS" (sn.py:62-64): u is left untouched. */
int ganb_sn_power_iter(const ganb_sn_layer* layers_dev, int count, int total_blocks, int max_c, int assign,
                       void* stream);
int ganb_sn_bwd(const ganb_sn_layer* layers_dev, int count, int total_blocks, int max_c, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sub-pixel form of UpsampleConv (common/resnet_block.py:83-97: tf.depth_to_space(concat x4) = nearest 2x, then
 * lib.ops.conv2d.Conv2D 3x3 SAME): replaces tf.concat + tf.depth_to_space + tf.nn.conv2d and their gradients.
 * Output pixel (2a+i, 2b+j) only sees x[a+i-1 .. a+i, b+j-1 .. b+j], so the layer is four 2x2 convolutions over the
 * LOW-resolution tensor with effective filters E_ij[p][q] = sum_{r in R_i[p], s in R_j[q]} W[r][s],
 * R_0 = ({0}, {1,2}), R_1 = ({0,1}, {2}): 4/9 of the MMA work, no upsampled operand in HBM.
 *   ganb_upconv_pack  : W fp32 [3][3][cin][cout] -> E bf16 as we_t [16 = 4*(2i+j) + 2p+q][cout][cin] (fprop operand)
 *                       and we_n [16][cin][cout] (dgrad operand).
 *   ganb_upconv_fprop : x bf16 [n,h,w,cin] -> y in QUAD LAYOUT [n, h, w, 4 = 2i+j, cout] (= NHWC [n,2h,2w,cout] with
 *                       the pixels of each 2x2 cell stored together), y = act(alpha * conv + bias), f32 or bf16.
 *   ganb_upconv_dgrad : dy quad bf16 -> dx [n,h,w,cin] = alpha * sum over the four parities (one launch).
 *   ganb_upconv_wgrad : dW[3][3][cin][cout] = beta*dW + scale * fold(dE), dE_ij[p][q] = x (shifted)^T dy_ij.
 * Supported shapes (ganb_upconv_supported): h % 16 == 0, w % 8 == 0, cin, cout multiples of 64 and >= 128.
 * Quad-layout tensors are consumed by ganb_bn_stats (layout-agnostic), ganb_colsum, and by ganb_norm_act_fwd / _bwd
 * with upsample = 2. */
int ganb_upconv_supported(int n, int h, int w, int cin, int cout);
int ganb_upconv_pack(const float* w_hwio, void* we_t_bf16, void* we_n_bf16, int cin, int cout, void* stream);
int ganb_upconv_fprop(const void* x_bf16, const void* we_t_bf16, void* y_quad, int n, int h, int w, int cin, int cout,
                      const float* alpha, const float* bias, int act, int out_dtype, void* stream);
/* ganb_upconv_fprop with the fused statistics of ganb_conv2d_igemm_stats: one row per (low-resolution pixel tile,
 * output parity), rows = ganb_upconv_stats_rows(...) per tower (0: unsupported). */
int ganb_upconv_stats_rows(int n, int h, int w, int cin, int cout, int groups);
int ganb_upconv_fprop_stats(const void* x_bf16, const void* we_t_bf16, void* y_quad, int n, int h, int w, int cin,
                            int cout, const float* alpha, const float* bias, int act, int out_dtype, float* stats,
                            int groups, void* stream);
int ganb_upconv_dgrad(const void* dy_quad_bf16, const void* we_n_bf16, void* dx, int n, int h, int w, int cin, int cout,
                      const float* alpha, int out_dtype, void* stream);
/* workspace: ganb_upconv_wgrad_workspace() bytes, 32-byte aligned (as for ganb_conv2d_wgrad) */
int64_t ganb_upconv_wgrad_workspace(int n, int h, int w, int cin, int cout);
int ganb_upconv_wgrad(const void* x_bf16, const void* dy_quad_bf16, float* dw_hwio, void* workspace, int n, int h, int w,
                      int cin, int cout, const float* scale, float beta, void* stream);

/* fp32 HWIO filters -> bf16 operands of the tensor-core kernels; wn = [tap][ci_pad][co], wt = [tap][co][ci_pad]
 * (either may be NULL).  ci_pad >= ci (0 = ci) is the input-channel count of the operand copies: rows / columns
 * ci..ci_pad-1 are left untouched (the caller zeroes them once), so that a filter with ci % 8 != 0 (the 513-channel
 * convolution behind minibatch_std, PGGAN/model_nvidia.py:223-229) meets the 16-byte rows of the TMA path.
 * tile_begin = prefix sum of taps*ceil(ci/64)*ceil(co/64) over the layers (one block per 64 x 64 tile of a tap). */
typedef struct ganb_pack_layer {
  const float* w;
  void* wn;
  void* wt;
  int32_t taps, ci, co, tile_begin;
  int32_t ci_pad, pad_;
} ganb_pack_layer;
int ganb_pack_weights(const ganb_pack_layer* layers_dev, int count, int total_tiles, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation + activation + resampling (HBM-bound, float4 / bf16x4 vectorised; c % 4 == 0).
 * Replaces tf.nn.moments + tf.nn.embedding_lookup x2 + tf.nn.batch_normalization
 * (common/ops/normalization.py:47-57), tf.contrib.layers.batch_norm / instance_norm in training mode
 * (:8-24, :105-122), tf.nn.relu / tf.maximum(x,0.2x) (common/resnet_block.py:24-29) and the nearest
 * upsample tf.depth_to_space(concat x4) (:87-88) that follows them in every generator block.
 *
 * `groups` splits the batch into contiguous chunks with separate statistics: 1 = BatchNorm over the call,
 * 2 = the reference's two per-device towers batched into one call, n = instance norm.
 * gamma/beta are tables [n_labels, c] indexed by labels[n] (labels NULL -> row 0; gamma NULL -> 1/0).
 * mean == NULL skips normalisation (pure activation + cast + resample).
 * upsample: 0 = none, 1 = the output is written through a nearest 2x upsample (backward: dz is [n,2h,2w,c]),
 * 2 = x (and dx) are stored in the QUAD LAYOUT of ganb_upconv_fprop, out / dz are plain NHWC at the same resolution
 *     (all-bf16 path only).
 * ---------------------------------------------------------------------------------------------- */
int64_t ganb_bn_stats_workspace(int n, int hw, int c, int groups);
/* x is fp32 or bf16 (x_dtype); statistics are accumulated in fp32 per chunk and fp64 across chunks. One launch:
 * the last block of each group finalises mean / rstd. */
int ganb_bn_stats(const void* x, int x_dtype, int n, int hw, int c, int groups, float eps, float* mean, float* rstd,
                  void* workspace, void* stream);
/* out[n,(2)h,(2)w, out_cstride] (f32/bf16) = act(norm(x)), optionally replicated 2x2; out_raw (bf16, input
 * resolution, optional) = x.  cstride arguments of 0 mean "c". */
int ganb_norm_act_fwd(const void* x, int x_dtype, int n, int h, int w, int c, const float* mean, const float* rstd, int groups,
                      const float* gamma, const float* beta, const int* labels, int act, int upsample, void* out,
                      int out_dtype, int out_cstride, void* out_raw_bf16, int raw_cstride, void* stream);
/* dx = d(loss)/dx given dz = d(loss)/d(out); dgamma/dbeta (tables of n_rows rows, may be NULL) are accumulated;
 * `add` (fp32 or bf16, optional) is added to dx. */
int64_t ganb_norm_act_bwd_workspace(int n, int hw, int c, int groups);
int ganb_norm_act_bwd(const void* x, int x_dtype, const void* dz, int dz_dtype, int dz_cstride, int n, int h, int w, int c,
                      const float* mean, const float* rstd, int groups, const float* gamma, const float* beta,
                      const int* labels, int n_rows, int act, int upsample, float* dgamma, float* dbeta,
                      const void* add, int add_dtype, void* dx, int dx_dtype, void* workspace, void* stream);

/* Cross-GPU batch statistics (the BatchNorm statistic reduction of the multi-GPU path; the reference itself keeps
 * statistics per tower, common/ops/normalization.py:47).  Forward: ganb_bn_stats on the local share, then
 * ganb_bn_moments_pack -> [mean | E[x^2]] (2*count floats, count = groups*c) -> NCCL all-reduce(sum) ->
 * ganb_bn_moments_unpack (inv_world = 1/ranks; equal shares per rank) -> global mean / rstd.
 * Backward: ganb_norm_act_bwd_phase(phase 1) leaves [sum(dy) | sum(dy*xhat)] (2*groups*c floats) at byte offset
 * ganb_norm_act_bwd_sums_offset() of the workspace -> all-reduce(sum) -> phase 2 applies with count_scale = 1/ranks. */
int ganb_bn_moments_pack(const float* mean, const float* rstd, int count, float eps, float* out, void* stream);
int ganb_bn_moments_unpack(const float* sums, int count, float inv_world, float eps, float* mean, float* rstd,
                           void* stream);
int64_t ganb_norm_act_bwd_sums_offset(int n, int hw, int c, int groups);
int ganb_norm_act_bwd_phase(const void* x, int x_dtype, const void* dz, int dz_dtype, int dz_cstride, int n, int h, int w,
                            int c, const float* mean, const float* rstd, int groups, const float* gamma, const float* beta,
                            const int* labels, int n_rows, int act, int upsample, float* dgamma, float* dbeta,
                            const void* add, int add_dtype, void* dx, int dx_dtype, void* workspace, int phase,
                            float count_scale, void* stream);

/* 2x2 mean-pool written as in common/resnet_block.py:62-63 (add_n of four strided slices / 4), + optional add */
int ganb_meanpool2_fwd(const void* x, int x_dtype, const float* add, void* out, int out_dtype, int n, int h, int w,
                       int c, void* stream);
/* out[n,2h,2w,c] = scale * x replicated 2x2 (nearest upsample: scale 1; mean-pool backward: scale 0.25) */
int ganb_expand2(const void* x, int x_dtype, void* out, int out_dtype, int n, int h, int w, int c, float scale,
                 void* stream);
/* out[n,h/2,w/2,c] = scale * sum of each 2x2 block (nearest-upsample backward: scale 1) */
int ganb_sum2x2(const void* x, int x_dtype, void* out, int out_dtype, int n, int h, int w, int c, float scale,
                void* stream);
/* out[n, i*stride, j*stride, c] = x[n, i, j, c], zero elsewhere; out is [n, (h-1)*stride+1+extra_h, (w-1)*stride+1+extra_w, c].
 * The data gradient of a stride-s convolution (Conv2DBackpropInput) and tf.nn.conv2d_transpose (deconv2d.py:102) are
 * stride-1 convolutions of this dilated tensor. c % 4 == 0. */
int ganb_dilate2d(const void* x, int x_dtype, void* out, int out_dtype, int n, int h, int w, int c, int stride,
                  int out_h, int out_w, void* stream);
int ganb_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t count, float scale, void* stream);
int ganb_axpby(const float* x, float* y, int64_t count, float a, float b, void* stream); /* y = a*x + b*y */
/* out[c] = beta*out[c] + sum_rows x[row][c]: gradient of tf.nn.bias_add (conv2d.py:216, linear.py:180) */
int64_t ganb_colsum_workspace(int64_t rows, int c);
int ganb_colsum(const void* x, int x_dtype, int64_t rows, int c, float beta, float* out, void* workspace,
                void* stream);

/* Label conditioning of D: tf.tile + tf.concat of the embedded label (SNGAN/gan_cifar_resnet.py:282-284).
 * fwd writes e[n, 0:c2] into channels [coff, coff+c2) of every pixel of bf16 tensors with pixel stride cstride
 * (raw and/or activated copy); bwd reduces the two wide gradients back to de[n, c2];
 * concat_bwd_x gathers the gradient of the first c1 channels: dx = d_raw + act'(x) * d_act. */
int ganb_bcast_channels_fwd(const float* e, int n, int hw, int c2, int coff, int cstride, int act,
                            void* out_raw_bf16, void* out_act_bf16, void* stream);
int ganb_bcast_channels_bwd(const float* e, int n, int hw, int c2, int coff, int cstride, int act,
                            const void* d_raw, const void* d_act, int d_dtype, float* de, void* stream);
int ganb_concat_bwd_x(const float* x, int64_t pixels, int c1, int cstride, int act, const void* d_raw,
                      const void* d_act, int d_dtype, void* dx, int dx_dtype, void* stream);

/* out[n,c] = mean_hw act(x) : nonlinearity + tf.reduce_mean(axis=[1,2]) (gan_cifar_resnet.py:299-301) */
int ganb_act_mean_hw_fwd(const float* x, int n, int hw, int c, int act, float* out, void* stream);
int ganb_act_mean_hw_bwd(const float* x, const float* dout, int n, int hw, int c, int act, void* dx, int dx_dtype,
                         void* stream);

/* mode 0: hinge D loss mean(relu(1-d[:n_real])) + mean(relu(1+d[n_real:])) (gan_cifar_resnet.py:376-378)
 * mode 1: G loss -mean(d) (:492).  loss_out[0] (+)= scale*loss; dlogits = scale * dloss/dd.
 * General form (lib.misc.get_loss, common/misc.py:310-394): mode = 2 * loss_type + side, side 0 = d_loss over
 * logits = [disc_real (n_real) | disc_fake], side 1 = g_loss over disc_fake; loss_type 0 HINGE, 1 WGAN / WGAN-GP
 * (without the penalty term), 2 LSGAN, 3 CGAN, 4 Modified_MiniMax, 5 MiniMax. */
int ganb_gan_loss(const float* logits, int n, int n_real, int mode, float scale, int accumulate, float* loss_out,
                  float* dlogits, void* stream);

/* tf.reduce_mean(tf.nn.sparse_softmax_cross_entropy_with_logits(logits [n, c], labels [n])) -- the auxiliary-classifier
 * losses of ACGAN (ACGAN/train.py:110-121).  loss_out[0] (+)= scale*loss; dlogits [n, c] = scale * dloss/dlogits. */
int ganb_softmax_xent(const float* logits, const int* labels, int n, int c, float scale, int accumulate,
                      float* loss_out, float* dlogits, void* stream);

/* tf.train.AdamOptimizer update (gan_cifar_resnet.py:521-526) over one flat parameter buffer:
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t * m / (sqrt(v) + eps); lr_t is a DEVICE scalar that
 * already includes sqrt(1-b2^t)/(1-b1^t).  g is read as grad_scale * grads. */
int ganb_adam(float* params, const float* grads, float* m, float* v, int64_t count, const float* lr_t, float beta1,
              float beta2, float eps, float grad_scale, void* stream);

/* int32 [B, 3*hw] CHW pixels -> 2*(x/256 - .5) + noise -> NHWC fp32 (gan_cifar_resnet.py:334-337) */
int ganb_preprocess_real(const int* data, const float* noise, int b, int hw, float* out, void* stream);

/* tf.nn.embedding_lookup (common/ops/embedding.py:51) and its IndexedSlices gradient summed by index */
int ganb_embedding_fwd(const float* table, const int* labels, int n, int dim, float* out, void* stream);
int ganb_embedding_bwd(const float* dout, const int* labels, int n, int dim, int vocab, float* dtable, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PGGAN / Pix2Pix sides of the hot path (bandwidth-bound; c % 4 == 0).
 * pixel_norm: y = act(x * rsqrt(mean_c(x^2) + eps)) -- common/ops/normalization.py:125-140 followed by lrelu
 *   (PGGAN/model_nvidia.py:15-17, 63-68); bwd: dz = dy*act'(x r), dx = r dz - x r^3 mean_c(dz x).
 * minibatch_std: PGGAN/model_nvidia.py:20-28; out = concat(x, mean_{h,w,c} sqrt(var_b(x) + 1e-8)), channel stride cs.
 * copy_channels: dst[pix, dst_off + j] = scale * mask[pix, j] * src[pix, src_off + j]: tf.concat of U-Net skips and
 *   its backward slice (Pix2Pix/networks.py:268-270), tf.nn.dropout with a given keep mask (:263-264).
 * ---------------------------------------------------------------------------------------------- */
int ganb_pixel_norm_fwd(const void* x, int x_dtype, void* y, int y_dtype, int64_t pixels, int c, float eps, int act,
                        void* stream);
int ganb_pixel_norm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype, int64_t pixels,
                        int c, float eps, int act, void* stream);
int64_t ganb_minibatch_std_workspace(int b, int h, int w, int c);
int ganb_minibatch_std_fwd(const float* x, int b, int h, int w, int c, int cs, float* out, void* workspace, void* stream);
int ganb_minibatch_std_bwd(const float* x, const float* dout, int b, int h, int w, int c, int cs, float* dx,
                           void* workspace, void* stream);
int ganb_copy_channels(const void* src, int src_dtype, int src_cstride, int src_off, void* dst, int dst_dtype,
                       int dst_cstride, int dst_off, int64_t pixels, int c, const float* mask, float scale, void* stream);

/* scale * tf.reduce_mean(tf.abs(targets - outputs)) (Pix2Pix/train.py:511) and its gradient w.r.t. outputs
 * (scale * sign(outputs - targets) / count, 0 at equality like tf.abs); loss_out[0] (+)= the value. */
int64_t ganb_l1_loss_workspace(int64_t count);
int ganb_l1_loss(const float* targets, const float* outputs, int64_t count, float scale, int accumulate, float* loss_out,
                 float* doutputs, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * WGAN-GP gradient penalty (ACGAN/train.py:97-105; tf.random_uniform interpolation, tf.gradients of D w.r.t. the
 * interpolates, 10 * mean((slopes - 1)^2)), differentiated w.r.t. D's parameters through the backward pass.
 *   ganb_interpolate : out[n, :] = real[n, :] + alpha[n] * (fake[n, :] - real[n, :])
 *   ganb_gp_loss     : loss_out[0] (+)= scale * mean_n (sqrt(sum g[n,:]^2 + 1e-10) - 1)^2, dg = d/dg; workspace n floats
 *   ganb_bn_bwd_vjp  : vector-Jacobian product of the BACKWARD of act(batch_norm(x)) (training mode, one statistic
 *                      group): for gx = BNbwd(x, gy) and a cotangent `cot` of gx it returns d/dx, d/dgy and accumulates
 *                      d/dgamma (the "grad-grad" of fused batch norm that tf.gradients builds for the penalty). */
int ganb_interpolate(const float* real, const float* fake, const float* alpha, int n, int64_t per_sample, float* out,
                     void* stream);
int ganb_gp_loss(const float* g, int n, int64_t per_sample, float scale, int accumulate, float* loss_out, float* dg,
                 float* workspace_n, void* stream);
int64_t ganb_bn_bwd_vjp_workspace(int64_t pixels, int c);
int ganb_bn_bwd_vjp(const float* x, const float* gy, const float* cot, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, int64_t pixels, int c, int act, float* dx, float* dgy,
                    float* dgamma, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Secondary layer variants, all bandwidth-bound fp32 arithmetic.
 *
 * Effective filters: W_eff = W * (g / ||W||) * mask with the norm over every axis but the output-channel one
 *   (weight-norm: common/ops/conv2d.py:153-163, linear.py:143-155, deconv2d.py:87-96; PixelCNN mask: conv2d.py:63-81,
 *   165-167).  Geometry [a][c][b], element (i, ch, j) at (i*c + ch)*b + j, the norm runs over i and j:
 *   Conv2D Filters [k,k,Cin,Cout] -> (k*k*Cin, Cout, 1); Linear W [in,out] -> (in, out, 1);
 *   Deconv2D Filters [k,k,Cout,Cin] -> (k*k, Cout, Cin).  g == NULL: mask only; mask == NULL: weight-norm only.
 *   fwd writes w_eff and norms[c]; bwd ADDS into dw and dg:  dV = dw_eff*mask, dot = sum dV*W,
 *   dw += (g/||W||) * (dV - W*dot/||W||^2), dg += dot/||W||.  workspace: ganb_weight_transform_workspace(a, c) bytes.
 *
 * Layer norm of the critic: tf.contrib.layers.layer_norm(begin_norm_axis=1, begin_params_axis=-1), i.e. moments over
 *   (h, w, c) per sample with tf.nn.moments, then tf.nn.batch_normalization(x, mean, var, beta[c], gamma[c], 1e-12)
 *   (common/ops/normalization.py:62-82, reached through SNGAN/gan_cifar_resnet.py:99-100), followed by `act`
 *   (common/resnet_block.py:24-29) in the same pass.  per_sample = h*w*c, 1024 % c == 0.  fwd writes y and
 *   mean_rstd[n][2]; bwd writes dx (may be NULL) and chan_partials[2][rows][c] (rows = ganb_layer_norm_rows), whose
 *   column sums (ganb_colsum) are dgamma and dbeta.  workspace: ganb_layer_norm_workspace bytes for either direction.
 *
 * Fade-in blend of PGGAN ((1 - alpha)*a + alpha*b, PGGAN/model_nvidia.py:118, :200) with alpha in DEVICE memory (the
 *   reference feeds it per step, PGGAN/train.py:83, 184), so that captured CUDA graphs follow it.
 * ---------------------------------------------------------------------------------------------- */
int64_t ganb_weight_transform_workspace(int a, int c);
int ganb_weight_transform_fwd(const float* w, const float* g, const float* mask, float* w_eff, float* norms,
                              void* workspace, int a, int c, int b, void* stream);
int ganb_weight_transform_bwd(const float* w, const float* dw_eff, const float* g, const float* mask, const float* norms,
                              float* dw, float* dg, void* workspace, int a, int c, int b, void* stream);
int64_t ganb_layer_norm_workspace(int n, int64_t per_sample, int c);
int64_t ganb_layer_norm_rows(int n, int64_t per_sample);
int ganb_layer_norm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype,
                        float* mean_rstd, void* workspace, int n, int64_t per_sample, int c, float eps, int act,
                        void* stream);
int ganb_layer_norm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* mean_rstd,
                        const float* gamma, const float* beta, void* dx, int dx_dtype, float* chan_partials,
                        void* workspace, int n, int64_t per_sample, int c, int act, void* stream);
/* y = alpha * x, alpha a device scalar: W / sigma for the `depthwise_filters` of a spectrally-normalised
 * depthwise / separable convolution (common/ops/conv2d.py:173-175); the other layers carry 1/sigma in the GEMM epilogue. */
int ganb_scale_dev(const float* x, const float* alpha, float* y, int64_t count, void* stream);
int ganb_lerp_fwd(const float* a, const float* b, float* y, int64_t count, const float* alpha, void* stream);
int ganb_lerp_bwd(const float* dy, void* da, int da_dtype, void* db, int db_dtype, int64_t count, const float* alpha,
                  void* stream);
/* tf.image.resize_nearest_neighbor to 1/stride of the size (align_corners=False: source index = i*stride), the skip path
 * of the ResNet PGGAN critic (common/resnet_block.py:286-287).  (h, w) is always the LARGE size.
 *   scatter = 0: x [n,h,w,c] -> y [n,ceil(h/s),ceil(w/s),c], y[i,j] = x[i*s, j*s];
 *   scatter = 1 (its gradient): x [n,ceil(h/s),ceil(w/s),c] -> y [n,h,w,c], zero off the sampled grid.  Any c. */
int ganb_subsample2d(const void* x, int x_dtype, void* y, int y_dtype, int n, int h, int w, int c, int stride,
                     int scatter, void* stream);

/* Sample grid of generate_image / save_images (SNGAN/gan_cifar_resnet.py:536-539, common/misc.py:215-244): samples
 * [n,h,w,c] in (-1, 1) -> ((s + 1) * 127.5) truncated toward zero (astype('int32')), clamped to [0, 255], image k
 * written to tile (k / nw, k % nw) of the uint8 grid [ceil(n/nw)*h, nw*w, c]; unused tiles are zero. */
int ganb_sample_grid(const void* samples, int dtype, int n, int h, int w, int c, int nw, unsigned char* grid,
                     void* stream);

/* tf.nn.depthwise_conv2d, and the depthwise half of tf.nn.separable_conv2d (common/ops/conv2d.py:188-208; the
 * pointwise half is ganb_conv2d_igemm with a 1x1 filter):
 *   y[n,ho,wo,ci*cm+m] = sum_{r,s} x[n, ho*stride+r-pad_t, wo*stride+s-pad_l, ci] * filter[r,s,ci,m],
 * filter = `depthwise_filters` [kh,kw,c,cm] fp32 (conv2d.py:146-148), + bias[c*cm] when not NULL.  bwd_input writes dx [n,h,w,c]; bwd_filter writes
 * partials[chunks][kh*kw][c*cm] (chunks = ganb_depthwise_conv2d_chunks(n, ho, wo)) whose column sums over the chunk
 * axis (ganb_colsum, rows = chunks, c = kh*kw*c*cm) are the filter gradient in filter layout.  Any channel count. */
int ganb_depthwise_conv2d_fwd(const void* x, int x_dtype, const float* filter, const float* bias, void* y, int y_dtype,
                              int n, int h, int w, int c, int cm, int ho, int wo, int kh, int kw, int stride, int pad_t,
                              int pad_l, void* stream);
int ganb_depthwise_conv2d_bwd_input(const void* dy, int dy_dtype, const float* filter, void* dx, int dx_dtype, int n,
                                    int h, int w, int c, int cm, int ho, int wo, int kh, int kw, int stride, int pad_t,
                                    int pad_l, void* stream);
int64_t ganb_depthwise_conv2d_chunks(int n, int ho, int wo);
int ganb_depthwise_conv2d_bwd_filter(const void* x, int x_dtype, const void* dy, int dy_dtype, float* partials, int n,
                                     int h, int w, int c, int cm, int ho, int wo, int kh, int kw, int stride, int pad_t,
                                     int pad_l, void* stream);

/* TF32 operand mode (north star: "BF16 or TF32 inputs and FP32 accumulation"; tolerance <= 1e-3 relative per layer): the
 * implicit-GEMM convolution and its filter gradient with fp32 NHWC operands read by tcgen05.mma kind::tf32 -- the
 * TensorFlow calls replaced are the same tf.nn.conv2d / Conv2DBackpropInput / Conv2DBackpropFilter
 * (common/ops/conv2d.py:181-187).  Arguments as ganb_conv2d_igemm / ganb_conv2d_wgrad with fp32 x / dy and the filter
 * copies in fp32: wp = [taps][cout][cin] for fprop (ganb_transpose_tf32 of the HWIO filter), the HWIO filter itself
 * [taps][cin][cout] with flip_taps = 1 and the roles of cin / cout exchanged for the data gradient.  cin % 4 == 0
 * (and cout % 4 == 0 for the filter gradient).  The MMA truncates the 13 low mantissa bits: callers round operands to
 * nearest with ganb_round_tf32 / ganb_transpose_tf32 first. */
int ganb_conv2d_igemm_tf32(const float* x, const float* wp, void* y, int n, int h, int w, int cin, int ho, int wo,
                           int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps, const float* alpha,
                           const float* bias, const float* residual, int residual_up2, int act, int out_dtype,
                           void* stream);
/* workspace: ganb_conv2d_wgrad_tf32_workspace() bytes, 32-byte aligned */
int64_t ganb_conv2d_wgrad_tf32_workspace(int n, int ho, int wo, int cin, int cout, int kh, int kw);
int ganb_conv2d_wgrad_tf32(const float* x, const float* dy, float* dw, void* workspace, int n, int h, int w, int cin,
                           int ho, int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l, const float* scale,
                           float beta, void* stream);
int ganb_round_tf32(const float* x, float* y, int64_t count, void* stream);
int ganb_transpose_tf32(const float* w_hwio, float* wt, int taps, int cin, int cout, void* stream);

/* Cross-GPU minibatch-stddev (PGGAN/model_nvidia.py:20-28 with the batch statistics taken over the GLOBAL batch of a
 * data-parallel run; SURVEY 8(e) collective 3).  Two phases around the caller's all-reduce, like ganb_norm_act_bwd_phase:
 *   fwd phase 1: workspace[0, 2m) = [sum_b x | sum_b x^2] per (h, w, c) position (m = h*w*c)  -> all-reduce (sum)
 *   fwd phase 2: mean / sd from the global sums (batch b * world), s = mean(sd), out = concat(x, s)
 *   bwd phase 1: g = sum of dout[..., c] at workspace float offset ganb_minibatch_std_sync_g_offset() -> all-reduce
 *   bwd phase 2: dx = dout[..., :c] + g * (x - mean) / (b * world * m * sd) */
int64_t ganb_minibatch_std_sync_workspace(int h, int w, int c);
int64_t ganb_minibatch_std_sync_g_offset(int h, int w, int c);
int ganb_minibatch_std_sync_fwd(const float* x, int b, int h, int w, int c, int cs, float* out, void* workspace,
                                int phase, int world, void* stream);
int ganb_minibatch_std_sync_bwd(const float* x, const float* dout, int b, int h, int w, int c, int cs, float* dx,
                                void* workspace, int phase, int world, void* stream);

/* Small all-reduces over NVLink / NVSwitch peer memory (csrc/peer.cu) for the statistic exchanges of the data-parallel
 * path: cross-GPU BatchNorm moments (reference coupling point common/ops/normalization.py:47) and the
 * [sum dy | sum dy*xhat] pair of its backward pass.  peer_bufs: HOST array of `world` device pointers = this process'
 * mappings of every rank's symmetric buffer (same size and layout on every rank, zero-initialised, all ranks past a
 * barrier before the first call); site_offset: byte offset of the call site's region (ganb_peer_site_bytes(count) bytes,
 * 128-byte aligned) -- one region per call site, the same offsets on every rank.  One single-CTA kernel per call: the
 * contribution is written into our region, a system-scope release store raises our flag in every peer's region, the peers'
 * flags are awaited (bounded spin), and the contributions are summed in rank order (bit-identical result on every
 * rank).  The call epoch lives in device memory, so the kernels can be captured in CUDA graphs.
 *   ganb_peer_allreduce : out[i] = scale * sum_ranks in[i], count <= ganb_peer_max_count() (out may alias in)
 *   ganb_peer_bn_moments: (mean, rstd)[count] of the local share -> statistics over all ranks (equal shares), in place:
 *                         pack [mean | E x^2], exchange, unpack -- one launch */
int64_t ganb_peer_site_bytes(int count);
int ganb_peer_max_count(void);
int ganb_peer_allreduce(const float* in, float* out, int count, float scale, void* const* peer_bufs, int rank, int world,
                        int64_t site_offset, void* stream);
int ganb_peer_bn_moments(float* mean, float* rstd, int count, float eps, void* const* peer_bufs, int rank, int world,
                         int64_t site_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANB200_H_ */
