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
 * ---------------------------------------------------------------------------------------------- */
int ganb_conv2d_igemm(const void* x_bf16, const void* wp_bf16, void* y, int n, int h, int w, int cin, int ho,
                      int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps,
                      const float* alpha, const float* bias, const float* residual, int act, int out_dtype,
                      void* stream);

/* Filter gradient: partial[split][t][ci][co] = sum over the split's pixels of
 *     x[n, ho + r - pad_t, wo + s - pad_l, ci] * dy[n, ho, wo, co]          (stride 1)
 * `workspace` must hold ganb_conv2d_wgrad_workspace() bytes; the reduction over splits, the optional
 * scale and the accumulation into dw (HWIO f32) are done by the same call (second kernel).
 *     dw = beta * dw + scale * sum_split partial
 * Requirements: cin % 8 == 0, cout % 8 == 0. */
int64_t ganb_conv2d_wgrad_workspace(int n, int h, int w, int cin, int ho, int wo, int cout, int kh, int kw);
int ganb_conv2d_wgrad(const void* x_bf16, const void* dy_bf16, float* dw, void* workspace, int n, int h, int w,
                      int cin, int ho, int wo, int cout, int kh, int kw, int pad_t, int pad_l,
                      const float* scale, float beta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANB200_H_ */
