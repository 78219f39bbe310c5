// Secondary layer variants of the hot path -- all bandwidth-bound, fp32 arithmetic:
//
//   * effective filters: weight-norm  W * (g / ||W||)  over every axis but the output-channel one, followed by the
//     constant PixelCNN mask (common/ops/conv2d.py:63-81, 153-167; linear.py:143-155; deconv2d.py:87-96), and the
//     backward map from dL/dW_eff to dL/dW and dL/dg;
//   * layer norm of the critic (tf.contrib.layers.layer_norm, begin_norm_axis=1, begin_params_axis=-1,
//     common/ops/normalization.py:62-82; reached with NORMALIZATION_D, SNGAN/gan_cifar_resnet.py:99-100) with the
//     activation that follows it fused, forward and backward;
//   * the fade-in blend of PGGAN (PGGAN/model_nvidia.py:118, :200) with alpha read from device memory, so a captured
//     CUDA graph follows alpha = step / max_iter without being re-captured;
//   * tf.image.resize_nearest_neighbor to HALF the size (the skip path of the ResNet PGGAN critic,
//     common/resnet_block.py:286-287): y[i, j] = x[2i, 2j], and its gradient (zeros off the sampled grid), for any
//     channel count (the critic's input is RGB).
#include "host_common.h"

#include <cuda_bf16.h>

#define STREAM static_cast<cudaStream_t>(stream)

namespace ganb {

struct alignas(8) bf16x4v {
  __nv_bfloat162 lo, hi;
};
__device__ __forceinline__ float4 vld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 vld4(const __nv_bfloat16* p) {
  const bf16x4v v = *reinterpret_cast<const bf16x4v*>(p);
  const float2 a = __bfloat1622float2(v.lo), b = __bfloat1622float2(v.hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void vst4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void vst4(__nv_bfloat16* p, float4 v) {
  bf16x4v o;
  o.lo = __floats2bfloat162_rn(v.x, v.y);
  o.hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<bf16x4v*>(p) = o;
}
__device__ __forceinline__ float vact(float v, int act) {
  if (act == GANB_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == GANB_ACT_LRELU) return v >= 0.f ? v : 0.2f * v;
  return v;
}
__device__ __forceinline__ float vdact(float z, int act) {
  if (act == GANB_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == GANB_ACT_LRELU) return z >= 0.f ? 1.f : 0.2f;
  return 1.f;
}
__device__ __forceinline__ float vwarp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum over the 256 threads of a block, returned to every thread.  `red` holds >= 8 floats.
__device__ __forceinline__ float vblock_sum(float v, float* red) {
  v = vwarp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  return s;
}

// ================================================================================================ effective filters
// Geometry [A][C][B] (element (a, c, b) at (a*C + c)*B + b), the norm runs over a and b for every c:
//   Conv2D Filters [k, k, Cin, Cout]  -> A = k*k*Cin, C = Cout, B = 1     (conv2d.py:154, axis (0, 1, 2))
//   Linear W [in, out]                -> A = in,      C = out,  B = 1     (linear.py:146, axis 0)
//   Deconv2D Filters [k, k, Cout, Cin]-> A = k*k,     C = Cout, B = Cin   (deconv2d.py:90, axes (0, 1, 3))
// B == 1: a block owns 32 neighbouring channels x WT_ROWS rows (128-byte row segments); B > 1: a block owns one
// channel x WT_ROWS slabs of B contiguous floats.  Two launches per direction: per-chunk partial sums, then the apply
// kernel folds the partials of its channels (deterministic order) and streams the tile.
constexpr int WT_ROWS = 64;

template <bool kDot>
__global__ void __launch_bounds__(256)
wt_partial_kernel(const float* __restrict__ W, const float* __restrict__ dWe, const float* __restrict__ mask,
                  float* __restrict__ partial, int A, int C, int B) {
  pdl_wait();
  __shared__ float red[256];
  const int chunk = blockIdx.y;
  const int a0 = chunk * WT_ROWS;
  const int a1 = min(A, a0 + WT_ROWS);
  float acc = 0.f;
  if (B == 1) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (c < C) {
      for (int a = a0 + (threadIdx.x >> 5); a < a1; a += 8) {
        const int64_t i = static_cast<int64_t>(a) * C + c;
        const float w = W[i];
        if (kDot) {
          float d = dWe[i];
          if (mask) d *= mask[i];
          acc += d * w;
        } else {
          acc += w * w;
        }
      }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < 32 && c < C) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) s += red[r * 32 + threadIdx.x];
      partial[static_cast<int64_t>(chunk) * C + c] = s;
    }
  } else {
    const int c = blockIdx.x;
    const int64_t items = static_cast<int64_t>(a1 - a0) * B;
    for (int64_t j = threadIdx.x; j < items; j += 256) {
      const int a = a0 + static_cast<int>(j / B), b = static_cast<int>(j % B);
      const int64_t i = (static_cast<int64_t>(a) * C + c) * B + b;
      const float w = W[i];
      if (kDot) {
        float d = dWe[i];
        if (mask) d *= mask[i];
        acc += d * w;
      } else {
        acc += w * w;
      }
    }
    const float s = vblock_sum(acc, red);
    if (threadIdx.x == 0) partial[static_cast<int64_t>(chunk) * C + c] = s;
  }
}

// kBwd = false: We = W * (g / ||W||) * mask, norms out.      kBwd = true: dW += s*(dV - W*dot/||W||^2), dg += dot/||W||
// with dV = dWe * mask, s = g / ||W||, dot = sum dV*W (the partials).  g == nullptr: mask only (scale 1, no norm term).
template <bool kBwd>
__global__ void __launch_bounds__(256)
wt_apply_kernel(const float* __restrict__ W, const float* __restrict__ src, const float* __restrict__ g,
                const float* __restrict__ mask, const float* __restrict__ partial, int chunks,
                float* __restrict__ norms, float* __restrict__ dst, float* __restrict__ dg, int A, int C, int B) {
  pdl_wait();
  __shared__ float s_scale[32], s_coef[32];
  const int chunk = blockIdx.y;
  const int a0 = chunk * WT_ROWS;
  const int a1 = min(A, a0 + WT_ROWS);
  const int width = (B == 1) ? 32 : 1;
  if (threadIdx.x < width) {
    const int c = (B == 1) ? blockIdx.x * 32 + threadIdx.x : blockIdx.x;
    float scale = 1.f, coef = 0.f;
    if (c < C && g != nullptr) {
      float tot = 0.f;
      for (int k = 0; k < chunks; ++k) tot += partial[static_cast<int64_t>(k) * C + c];
      if (!kBwd) {
        const float nrm = sqrtf(tot);
        scale = g[c] / nrm;
        if (chunk == 0) norms[c] = nrm;
      } else {
        const float nrm = norms[c];
        scale = g[c] / nrm;
        coef = tot / (nrm * nrm);
        if (chunk == 0 && dg != nullptr) dg[c] += tot / nrm;
      }
    }
    s_scale[threadIdx.x] = scale;
    s_coef[threadIdx.x] = coef;
  }
  __syncthreads();
  if (B == 1) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + lane;
    if (c >= C) return;
    const float scale = s_scale[lane], coef = s_coef[lane];
    for (int a = a0 + (threadIdx.x >> 5); a < a1; a += 8) {
      const int64_t i = static_cast<int64_t>(a) * C + c;
      const float m = mask ? mask[i] : 1.f;
      if (!kBwd) {
        dst[i] = W[i] * scale * m;
      } else {
        dst[i] += scale * (src[i] * m - W[i] * coef);
      }
    }
  } else {
    const int c = blockIdx.x;
    const float scale = s_scale[0], coef = s_coef[0];
    const int64_t items = static_cast<int64_t>(a1 - a0) * B;
    for (int64_t j = threadIdx.x; j < items; j += 256) {
      const int a = a0 + static_cast<int>(j / B), b = static_cast<int>(j % B);
      const int64_t i = (static_cast<int64_t>(a) * C + c) * B + b;
      const float m = mask ? mask[i] : 1.f;
      if (!kBwd) {
        dst[i] = W[i] * scale * m;
      } else {
        dst[i] += scale * (src[i] * m - W[i] * coef);
      }
    }
  }
}

static inline dim3 wt_grid(int A, int C, int B) {
  return dim3(B == 1 ? ceil_div(C, 32) : C, ceil_div(A, WT_ROWS));
}

// ================================================================================================ layer norm
// One sample = per_sample = H*W*C contiguous elements; a block owns LN_CHUNK of them (16 float4 per thread, held in
// registers so that the chunk mean and the squared deviations around it come from ONE pass over memory).  The apply
// kernels fold the per-chunk (count, mean, M2) triples with Chan's formula -- the two-pass accuracy of tf.nn.moments
// without the second read.
constexpr int LN_VEC = 16;
constexpr int LN_CHUNK = 256 * 4 * LN_VEC;   // 16384 elements

template <typename T>
__global__ void __launch_bounds__(256)
ln_stats_partial_kernel(const T* __restrict__ x, float* __restrict__ partial /*[n][splits][2]: mean, M2*/,
                        int64_t per_sample, int splits) {
  pdl_wait();
  __shared__ float red[8];
  const int s = blockIdx.x, n = blockIdx.y;
  const int64_t begin = static_cast<int64_t>(s) * LN_CHUNK;
  const int64_t len = min(static_cast<int64_t>(LN_CHUNK), per_sample - begin);
  const T* xp = x + n * per_sample + begin;
  float4 v[LN_VEC];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_VEC; ++i) {
    const int64_t e = (static_cast<int64_t>(i) * 256 + threadIdx.x) * 4;
    if (e < len) {
      v[i] = vld4(xp + e);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = vblock_sum(sum, red) / static_cast<float>(len);
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < LN_VEC; ++i) {
    const int64_t e = (static_cast<int64_t>(i) * 256 + threadIdx.x) * 4;
    if (e < len) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      m2 += (a * a + b * b) + (c * c + d * d);
    }
  }
  m2 = vblock_sum(m2, red);
  if (threadIdx.x == 0) {
    float* p = partial + (static_cast<int64_t>(n) * splits + s) * 2;
    p[0] = mean;
    p[1] = m2;
  }
}

// Folds the chunk statistics of sample n: returns (mean, rstd).  Every thread of the block computes the same value
// (splits <= a few dozen, the partials sit in L2).
__device__ __forceinline__ float2 ln_fold(const float* __restrict__ partial, int n, int splits, int64_t per_sample,
                                          float eps) {
  const float* p = partial + static_cast<int64_t>(n) * splits * 2;
  double tot = 0.0;
  for (int k = 0; k < splits; ++k) {
    const double cnt = static_cast<double>(min(static_cast<int64_t>(LN_CHUNK), per_sample - static_cast<int64_t>(k) * LN_CHUNK));
    tot += cnt * p[2 * k];
  }
  const double mean = tot / static_cast<double>(per_sample);
  double m2 = 0.0;
  for (int k = 0; k < splits; ++k) {
    const double cnt = static_cast<double>(min(static_cast<int64_t>(LN_CHUNK), per_sample - static_cast<int64_t>(k) * LN_CHUNK));
    const double d = p[2 * k] - mean;
    m2 += p[2 * k + 1] + cnt * d * d;
  }
  const float var = static_cast<float>(m2 / static_cast<double>(per_sample));
  return make_float2(static_cast<float>(mean), rsqrtf(var + eps));
}

// y = act(x*inv + (beta - mean*inv)), inv = rstd*gamma  (tf.nn.batch_normalization's arithmetic)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
ln_fwd_apply_kernel(const TIn* __restrict__ x, const float* __restrict__ partial, const float* __restrict__ gamma,
                    const float* __restrict__ beta, TOut* __restrict__ y, float* __restrict__ mean_rstd /*[n][2]*/,
                    int64_t per_sample, int splits, int c, float eps, int act) {
  pdl_wait();
  const int s = blockIdx.x, n = blockIdx.y;
  const float2 mr = ln_fold(partial, n, splits, per_sample, eps);
  if (s == 0 && threadIdx.x == 0) {
    mean_rstd[2 * n] = mr.x;
    mean_rstd[2 * n + 1] = mr.y;
  }
  const int64_t begin = static_cast<int64_t>(s) * LN_CHUNK;
  const int64_t len = min(static_cast<int64_t>(LN_CHUNK), per_sample - begin);
  const TIn* xp = x + n * per_sample + begin;
  TOut* yp = y + n * per_sample + begin;
#pragma unroll 4
  for (int i = 0; i < LN_VEC; ++i) {
    const int64_t e = (static_cast<int64_t>(i) * 256 + threadIdx.x) * 4;
    if (e >= len) break;
    const int ch = static_cast<int>((begin + e) % c);
    const float4 xv = vld4(xp + e), gv = vld4(gamma + ch), bv = vld4(beta + ch);
    float4 o;
    float inv = mr.y * gv.x; o.x = vact(xv.x * inv + (bv.x - mr.x * inv), act);
    inv = mr.y * gv.y;       o.y = vact(xv.y * inv + (bv.y - mr.x * inv), act);
    inv = mr.y * gv.z;       o.z = vact(xv.z * inv + (bv.z - mr.x * inv), act);
    inv = mr.y * gv.w;       o.w = vact(xv.w * inv + (bv.w - mr.x * inv), act);
    vst4(yp + e, o);
  }
}

// Backward, pass 1.  With xh = (x - mean)*rstd, z = xh*gamma + beta, dz = dy*act'(z), dyh = dz*gamma:
//   per sample:  s1 = sum dyh, s2 = sum dyh*xh           -> sums[n][splits][2]
//   per channel: dgamma += dz*xh, dbeta += dz            -> chan[2][n*splits][c]  (rows folded by ganb_colsum)
// 1024 % c == 0, so a thread meets the same four channels in every one of its LN_VEC float4s.
template <typename TX, typename TDY>
__global__ void __launch_bounds__(256)
ln_bwd_partial_kernel(const TX* __restrict__ x, const TDY* __restrict__ dy, const float* __restrict__ mean_rstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ sums,
                      float* __restrict__ chan, int64_t per_sample, int splits, int c, int act, int64_t rows) {
  pdl_wait();
  __shared__ float red[8];
  __shared__ float4 sg[256], sb[256];
  const int s = blockIdx.x, n = blockIdx.y;
  const float mean = mean_rstd[2 * n], rstd = mean_rstd[2 * n + 1];
  const int64_t begin = static_cast<int64_t>(s) * LN_CHUNK;
  const int64_t len = min(static_cast<int64_t>(LN_CHUNK), per_sample - begin);
  const TX* xp = x + n * per_sample + begin;
  const TDY* dp = dy + n * per_sample + begin;
  const int ch = static_cast<int>((begin + threadIdx.x * 4) % c);
  const float4 gv = vld4(gamma + ch), bv = vld4(beta + ch);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
  for (int i = 0; i < LN_VEC; ++i) {
    const int64_t e = (static_cast<int64_t>(i) * 256 + threadIdx.x) * 4;
    if (e >= len) break;
    const float4 xv = vld4(xp + e), dv = vld4(dp + e);
    float xh, dz, dyh;
    xh = (xv.x - mean) * rstd; dz = dv.x * vdact(xh * gv.x + bv.x, act); dyh = dz * gv.x;
    ag.x += dz * xh; ab.x += dz; s1 += dyh; s2 += dyh * xh;
    xh = (xv.y - mean) * rstd; dz = dv.y * vdact(xh * gv.y + bv.y, act); dyh = dz * gv.y;
    ag.y += dz * xh; ab.y += dz; s1 += dyh; s2 += dyh * xh;
    xh = (xv.z - mean) * rstd; dz = dv.z * vdact(xh * gv.z + bv.z, act); dyh = dz * gv.z;
    ag.z += dz * xh; ab.z += dz; s1 += dyh; s2 += dyh * xh;
    xh = (xv.w - mean) * rstd; dz = dv.w * vdact(xh * gv.w + bv.w, act); dyh = dz * gv.w;
    ag.w += dz * xh; ab.w += dz; s1 += dyh; s2 += dyh * xh;
  }
  s1 = vblock_sum(s1, red);
  s2 = vblock_sum(s2, red);
  sg[threadIdx.x] = ag;
  sb[threadIdx.x] = ab;
  __syncthreads();
  const int64_t row = static_cast<int64_t>(n) * splits + s;
  if (threadIdx.x == 0) {
    sums[row * 2] = s1;
    sums[row * 2 + 1] = s2;
  }
  // channel group j (4 channels) is met by the threads t with (g0 + t) % q == j, q = c / 4, g0 = (begin / 4) % q
  const int q = c >> 2;
  if (threadIdx.x < q) {
    const int j = threadIdx.x;
    const int g0 = static_cast<int>((begin >> 2) % q);
    const int first = ((j - g0) % q + q) % q;
    float4 tg = make_float4(0.f, 0.f, 0.f, 0.f), tb = tg;
    for (int t = first; t < 256; t += q) {
      const float4 a = sg[t], b = sb[t];
      tg.x += a.x; tg.y += a.y; tg.z += a.z; tg.w += a.w;
      tb.x += b.x; tb.y += b.y; tb.z += b.z; tb.w += b.w;
    }
    vst4(chan + row * c + j * 4, tg);
    vst4(chan + (rows + row) * c + j * 4, tb);
  }
}

// Backward, pass 2: dx = rstd * (dyh - mean(dyh) - xh * mean(dyh*xh))
template <typename TX, typename TDY, typename TDX>
__global__ void __launch_bounds__(256)
ln_bwd_apply_kernel(const TX* __restrict__ x, const TDY* __restrict__ dy, const float* __restrict__ mean_rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ sums,
                    TDX* __restrict__ dx, int64_t per_sample, int splits, int c, int act) {
  pdl_wait();
  const int s = blockIdx.x, n = blockIdx.y;
  const float mean = mean_rstd[2 * n], rstd = mean_rstd[2 * n + 1];
  double t1 = 0.0, t2 = 0.0;
  for (int k = 0; k < splits; ++k) {
    t1 += sums[(static_cast<int64_t>(n) * splits + k) * 2];
    t2 += sums[(static_cast<int64_t>(n) * splits + k) * 2 + 1];
  }
  const float m1 = static_cast<float>(t1 / static_cast<double>(per_sample));
  const float m2 = static_cast<float>(t2 / static_cast<double>(per_sample));
  const int64_t begin = static_cast<int64_t>(s) * LN_CHUNK;
  const int64_t len = min(static_cast<int64_t>(LN_CHUNK), per_sample - begin);
  const TX* xp = x + n * per_sample + begin;
  const TDY* dp = dy + n * per_sample + begin;
  TDX* op = dx + n * per_sample + begin;
  const int ch = static_cast<int>((begin + threadIdx.x * 4) % c);
  const float4 gv = vld4(gamma + ch), bv = vld4(beta + ch);
#pragma unroll 4
  for (int i = 0; i < LN_VEC; ++i) {
    const int64_t e = (static_cast<int64_t>(i) * 256 + threadIdx.x) * 4;
    if (e >= len) break;
    const float4 xv = vld4(xp + e), dv = vld4(dp + e);
    float4 o;
    float xh;
    xh = (xv.x - mean) * rstd; o.x = rstd * (dv.x * vdact(xh * gv.x + bv.x, act) * gv.x - m1 - xh * m2);
    xh = (xv.y - mean) * rstd; o.y = rstd * (dv.y * vdact(xh * gv.y + bv.y, act) * gv.y - m1 - xh * m2);
    xh = (xv.z - mean) * rstd; o.z = rstd * (dv.z * vdact(xh * gv.z + bv.z, act) * gv.z - m1 - xh * m2);
    xh = (xv.w - mean) * rstd; o.w = rstd * (dv.w * vdact(xh * gv.w + bv.w, act) * gv.w - m1 - xh * m2);
    vst4(op + e, o);
  }
}

static inline int ln_splits(int64_t per_sample) { return static_cast<int>(ceil_div64(per_sample, LN_CHUNK)); }
static inline bool ln_shape_ok(int64_t per_sample, int c) {
  return c >= 4 && c <= 1024 && (1024 % c) == 0 && per_sample % c == 0;
}

// ================================================================================================ fade-in blend
__global__ void __launch_bounds__(256)
lerp_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n,
                const float* __restrict__ alpha) {
  pdl_wait();
  const float t = *alpha, u = 1.f - t;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    const float4 p = vld4(a + i * 4), q = vld4(b + i * 4);
    vst4(y + i * 4, make_float4(u * p.x + t * q.x, u * p.y + t * q.y, u * p.z + t * q.z, u * p.w + t * q.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    y[i] = u * a[i] + t * b[i];
  }
}

// round-to-nearest (ties away) of fp32 values to TF32 (10-bit mantissa), kept in fp32 storage: operands of the kind::tf32
// tensor-core kernels (the MMA itself would truncate)
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_wait();
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) y[i] = round_tf32(x[i]);
}
// [taps][ci][co] (HWIO) -> [taps][co][ci], rounded to TF32: the fprop operand copy of a filter
__global__ void __launch_bounds__(256) transpose_tf32_kernel(const float* __restrict__ w, float* __restrict__ wt, int ci_n,
                                                             int co_n) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int tiles_co = (co_n + 31) / 32;
  const int tco = blockIdx.x % tiles_co, tci = blockIdx.x / tiles_co;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t plane = static_cast<int64_t>(ci_n) * co_n;
  const float* src = w + blockIdx.y * plane;
  float* dst = wt + blockIdx.y * plane;
  for (int j = ty; j < 32; j += 8) {
    const int ci = tci * 32 + j, co = tco * 32 + tx;
    tile[j][tx] = (ci < ci_n && co < co_n) ? src[static_cast<int64_t>(ci) * co_n + co] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int co = tco * 32 + j, ci = tci * 32 + tx;
    if (ci < ci_n && co < co_n) dst[static_cast<int64_t>(co) * ci_n + ci] = round_tf32(tile[tx][j]);
  }
}

// y = alpha * x with alpha read from device memory: the spectrally-normalised depthwise filter W_d / sigma (the
// depthwise kernels have no GEMM epilogue to carry 1/sigma; the filter is k*k*c*cm floats, i.e. tiny)
__global__ void __launch_bounds__(256)
scale_dev_kernel(const float* __restrict__ x, const float* __restrict__ alpha, float* __restrict__ y, int64_t n) {
  pdl_wait();
  const float t = *alpha;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) y[i] = t * x[i];
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
lerp_bwd_kernel(const float* __restrict__ dy, TA* __restrict__ da, TB* __restrict__ db, int64_t n,
                const float* __restrict__ alpha) {
  pdl_wait();
  const float t = *alpha, u = 1.f - t;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    const float4 g = vld4(dy + i * 4);
    if (da) vst4(da + i * 4, make_float4(u * g.x, u * g.y, u * g.z, u * g.w));
    if (db) vst4(db + i * 4, make_float4(t * g.x, t * g.y, t * g.z, t * g.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    if (da) da[i] = static_cast<TA>(u * dy[i]);
    if (db) db[i] = static_cast<TB>(t * dy[i]);
  }
}

// ================================================================================================ nearest 1/2 resize
// Forward (kScatter = false): y[n, i, j, :] = x[n, i*s, j*s, :], y is [n, oh, ow, c], x is [n, h, w, c].
// Backward (kScatter = true): y[n, I, J, :] = (I % s == 0 && J % s == 0) ? x[n, I/s, J/s, :] : 0, y is [n, oh, ow, c].
// One thread per output element (gather on both sides, coalesced writes); tensors on this path are RGB-sized.
template <typename TIn, typename TOut, bool kScatter>
__global__ void __launch_bounds__(256)
subsample2d_kernel(const TIn* __restrict__ x, TOut* __restrict__ y, int h, int w, int c, int oh, int ow, int stride,
                   int64_t total) {
  pdl_wait();
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ch = static_cast<int>(i % c);
    int64_t p = i / c;
    const int oj = static_cast<int>(p % ow);
    p /= ow;
    const int oi = static_cast<int>(p % oh);
    const int64_t img = p / oh;
    float v = 0.f;
    if (kScatter) {
      if (oi % stride == 0 && oj % stride == 0)
        v = static_cast<float>(x[((img * h + oi / stride) * w + oj / stride) * c + ch]);
    } else {
      v = static_cast<float>(x[((img * h + static_cast<int64_t>(oi) * stride) * w + static_cast<int64_t>(oj) * stride) * c + ch]);
    }
    y[i] = static_cast<TOut>(v);
  }
}

// ---- 8-channel fast paths (channel_multiplier 1, c % 8 == 0, bf16 activations): one 16-byte load per tap and thread
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

template <typename TOut>
__global__ void __launch_bounds__(256)
depthwise_fwd_v8_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ f, const float* __restrict__ bias,
                        TOut* __restrict__ y, int h, int w, int c, int ho, int wo, int kh, int kw, int stride, int pad_t,
                        int pad_l, int64_t total8) {
  pdl_wait();
  const int cg = c >> 3;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total8; i += gridDim.x * 256LL) {
    const int c8 = static_cast<int>(i % cg) * 8;
    int64_t p = i / cg;
    const int ow = static_cast<int>(p % wo);
    p /= wo;
    const int oh = static_cast<int>(p % ho);
    const int64_t img = p / ho;
    float acc[8];
    if (bias) ld8(bias + c8, acc);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
    for (int r = 0; r < kh; ++r) {
      const int hi = oh * stride + r - pad_t;
      if (hi < 0 || hi >= h) continue;
      for (int q = 0; q < kw; ++q) {
        const int wi = ow * stride + q - pad_l;
        if (wi < 0 || wi >= w) continue;
        float a[8], ff[8];
        ld8(x + ((img * h + hi) * w + wi) * c + c8, a);
        ld8(f + (r * kw + q) * c + c8, ff);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += a[j] * ff[j];
      }
    }
    st8(y + i * 8, acc);
  }
}

template <typename TDy, typename TOut>
__global__ void __launch_bounds__(256)
depthwise_bwd_input_v8_kernel(const TDy* __restrict__ dy, const float* __restrict__ f, TOut* __restrict__ dx, int h, int w,
                              int c, int ho, int wo, int kh, int kw, int stride, int pad_t, int pad_l, int64_t total8) {
  pdl_wait();
  const int cg = c >> 3;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total8; i += gridDim.x * 256LL) {
    const int c8 = static_cast<int>(i % cg) * 8;
    int64_t p = i / cg;
    const int wi = static_cast<int>(p % w);
    p /= w;
    const int hi = static_cast<int>(p % h);
    const int64_t img = p / h;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = 0; r < kh; ++r) {
      const int th = hi + pad_t - r;
      if (th < 0 || th % stride) continue;
      const int oh = th / stride;
      if (oh >= ho) continue;
      for (int q = 0; q < kw; ++q) {
        const int tw = wi + pad_l - q;
        if (tw < 0 || tw % stride) continue;
        const int ow = tw / stride;
        if (ow >= wo) continue;
        float g[8], ff[8];
        ld8(dy + ((img * ho + oh) * wo + ow) * c + c8, g);
        ld8(f + (r * kw + q) * c + c8, ff);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j] * ff[j];
      }
    }
    st8(dx + i * 8, acc);
  }
}

// Filter gradient, fast path: one block per pixel chunk handles ALL KS*KS taps (dy is read once per pixel); thread =
// (pixel lane, 8-channel group) with KS*KS*8 accumulators, lanes folded through shared memory tap by tap.
template <int KS, typename TDy>
__global__ void __launch_bounds__(256)
depthwise_bwd_filter_v8_kernel(const __nv_bfloat16* __restrict__ x, const TDy* __restrict__ dy,
                               float* __restrict__ partial, int h, int w, int c, int ho, int wo, int stride, int pad_t,
                               int pad_l, int64_t pixels, int64_t per_chunk) {
  pdl_wait();
  __shared__ float sm[2048];
  const int cg = c >> 3;
  const int lanes = 256 / cg;
  const int cgi = threadIdx.x % cg, lane = threadIdx.x / cg;
  const bool active = lane < lanes;
  const int c8 = cgi * 8;
  float acc[KS * KS][8];
#pragma unroll
  for (int t = 0; t < KS * KS; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int64_t p0 = blockIdx.x * per_chunk;
  const int64_t p1 = p0 + per_chunk < pixels ? p0 + per_chunk : pixels;
  if (active) {
    for (int64_t p = p0 + lane; p < p1; p += lanes) {
      const int ow = static_cast<int>(p % wo);
      const int64_t t2 = p / wo;
      const int oh = static_cast<int>(t2 % ho);
      const int64_t img = t2 / ho;
      float g[8];
      ld8(dy + p * c + c8, g);
#pragma unroll
      for (int r = 0; r < KS; ++r) {
        const int hi = oh * stride + r - pad_t;
        if (hi < 0 || hi >= h) continue;
#pragma unroll
        for (int q = 0; q < KS; ++q) {
          const int wi = ow * stride + q - pad_l;
          if (wi < 0 || wi >= w) continue;
          float a[8];
          ld8(x + ((img * h + hi) * w + wi) * c + c8, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r * KS + q][j] += a[j] * g[j];
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < KS * KS; ++t) {
    __syncthreads();
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sm[lane * c + c8 + j] = acc[t][j];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += 256) {
      float s2 = 0.f;
      for (int l = 0; l < lanes; ++l) s2 += sm[l * c + ch];
      partial[(static_cast<int64_t>(blockIdx.x) * (KS * KS) + t) * c + ch] = s2;
    }
  }
}

// ================================================================================================ sample grid
// generate_image + save_images (SNGAN/gan_cifar_resnet.py:536-539, common/misc.py:215-244): samples in (-1, 1) ->
// ((s + 1) * 127.5) truncated like astype('int32') -> image k at tile (k / nw, k % nw) of an [nh*h, nw*w, c] uint8 grid.
// One thread per grid byte (coalesced 1-byte stores, the reads of a tile row are contiguous h*w*c floats per image).
template <typename TIn>
__global__ void __launch_bounds__(256)
sample_grid_kernel(const TIn* __restrict__ x, unsigned char* __restrict__ grid, int n, int h, int w, int c, int nw,
                   int64_t total) {
  pdl_wait();
  const int64_t row_elems = static_cast<int64_t>(nw) * w * c;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int64_t gy = i / row_elems;
    const int64_t r = i - gy * row_elems;
    const int gx = static_cast<int>(r / c), ch = static_cast<int>(r - static_cast<int64_t>(gx) * c);
    const int tj = static_cast<int>(gy / h), y = static_cast<int>(gy - static_cast<int64_t>(tj) * h);
    const int ti = gx / w, xx = gx - ti * w;
    const int img = tj * nw + ti;
    int v = 0;
    if (img < n) {
      const float s = static_cast<float>(x[((static_cast<int64_t>(img) * h + y) * w + xx) * c + ch]);
      v = static_cast<int>((s + 1.f) * 127.5f);   // float -> int truncates toward zero, as NumPy's astype does
      v = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
    grid[i] = static_cast<unsigned char>(v);
  }
}

// ================================================================================================ depthwise conv
// tf.nn.depthwise_conv2d / the first half of tf.nn.separable_conv2d (common/ops/conv2d.py:188-208, Pix2Pix --conv_type):
//   y[n, ho, wo, ci*cm + m] = sum_{r,s} x[n, ho*stride + r - pad_t, wo*stride + s - pad_l, ci] * f[r, s, ci, m]
// No reuse across output channels: bandwidth-bound, one thread per output element with the channel fastest (coalesced
// stores; for cm = 1 the k*k tap reads are coalesced too and hit L1/L2 k*k/stride^2 times each).  fp32 arithmetic.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
depthwise_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ f, const float* __restrict__ bias,
                     TOut* __restrict__ y, int h, int w, int c, int cm, int ho, int wo, int kh, int kw, int stride,
                     int pad_t, int pad_l, int64_t total) {
  pdl_wait();
  const int oc_n = c * cm;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int oc = static_cast<int>(i % oc_n);
    int64_t p = i / oc_n;
    const int ow = static_cast<int>(p % wo);
    p /= wo;
    const int oh = static_cast<int>(p % ho);
    const int64_t img = p / ho;
    const int ci = oc / cm, m = oc - ci * cm;
    float acc = bias ? bias[oc] : 0.f;
    for (int r = 0; r < kh; ++r) {
      const int hi = oh * stride + r - pad_t;
      if (hi < 0 || hi >= h) continue;
      for (int q = 0; q < kw; ++q) {
        const int wi = ow * stride + q - pad_l;
        if (wi < 0 || wi >= w) continue;
        acc += static_cast<float>(x[((img * h + hi) * w + wi) * c + ci]) * f[((r * kw + q) * c + ci) * cm + m];
      }
    }
    y[i] = static_cast<TOut>(acc);
  }
}

// dx[n, hi, wi, ci] = sum_{r,s,m} dy[n, (hi + pad_t - r)/stride, (wi + pad_l - s)/stride, ci*cm + m] * f[r, s, ci, m]
template <typename TDy, typename TOut>
__global__ void __launch_bounds__(256)
depthwise_bwd_input_kernel(const TDy* __restrict__ dy, const float* __restrict__ f, TOut* __restrict__ dx, int h, int w,
                           int c, int cm, int ho, int wo, int kh, int kw, int stride, int pad_t, int pad_l,
                           int64_t total) {
  pdl_wait();
  const int oc_n = c * cm;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ci = static_cast<int>(i % c);
    int64_t p = i / c;
    const int wi = static_cast<int>(p % w);
    p /= w;
    const int hi = static_cast<int>(p % h);
    const int64_t img = p / h;
    float acc = 0.f;
    for (int r = 0; r < kh; ++r) {
      const int th = hi + pad_t - r;
      if (th < 0 || th % stride) continue;
      const int oh = th / stride;
      if (oh >= ho) continue;
      for (int q = 0; q < kw; ++q) {
        const int tw = wi + pad_l - q;
        if (tw < 0 || tw % stride) continue;
        const int ow = tw / stride;
        if (ow >= wo) continue;
        const TDy* g = dy + ((img * ho + oh) * wo + ow) * oc_n + ci * cm;
        const float* ff = f + ((r * kw + q) * c + ci) * cm;
        for (int m = 0; m < cm; ++m) acc += static_cast<float>(g[m]) * ff[m];
      }
    }
    dx[i] = static_cast<TOut>(acc);
  }
}

// Filter gradient, deterministic two-level reduction: block (chunk, tap) sums its pixel chunk for every output channel
// (threads stride over ci*cm + m: coalesced dy reads) into partial[chunk][tap][c*cm]; the chunks are folded by
// ganb_colsum (rows = chunks, columns = taps*c*cm).
constexpr int DW_CHUNKS_MAX = 512;
template <typename TIn, typename TDy>
__global__ void __launch_bounds__(256)
depthwise_bwd_filter_kernel(const TIn* __restrict__ x, const TDy* __restrict__ dy, float* __restrict__ partial, int h,
                            int w, int c, int cm, int ho, int wo, int kw, int stride, int pad_t, int pad_l,
                            int64_t pixels, int64_t per_chunk) {
  pdl_wait();
  const int oc_n = c * cm;
  const int tap = blockIdx.y, r = tap / kw, q = tap - r * kw;
  const int64_t p0 = blockIdx.x * per_chunk;
  const int64_t p1 = p0 + per_chunk < pixels ? p0 + per_chunk : pixels;
  for (int oc = threadIdx.x; oc < oc_n; oc += 256) {
    const int ci = oc / cm;
    float acc = 0.f;
    for (int64_t p = p0; p < p1; ++p) {
      const int ow = static_cast<int>(p % wo);
      const int64_t t = p / wo;
      const int oh = static_cast<int>(t % ho);
      const int64_t img = t / ho;
      const int hi = oh * stride + r - pad_t, wi = ow * stride + q - pad_l;
      if (hi < 0 || hi >= h || wi < 0 || wi >= w) continue;
      acc += static_cast<float>(x[((img * h + hi) * w + wi) * c + ci]) * static_cast<float>(dy[p * oc_n + oc]);
    }
    partial[(static_cast<int64_t>(blockIdx.x) * gridDim.y + tap) * oc_n + oc] = acc;
  }
}

static inline int flat_grid(int64_t items) {
  int64_t b = ceil_div64(items, 256);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace ganb

using namespace ganb;

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int64_t ganb_weight_transform_workspace(int a, int c) {
  return static_cast<int64_t>(ceil_div(a, WT_ROWS)) * c * 4;
}

extern "C" int ganb_weight_transform_fwd(const float* w, const float* g, const float* mask, float* w_eff, float* norms,
                                         void* workspace, int a, int c, int b, void* stream) {
  if (!w || !w_eff) return fail(GANB_E_BADARG, "weight_transform_fwd: null buffer");
  if (a < 1 || c < 1 || b < 1) return fail(GANB_E_BADARG, "weight_transform_fwd: bad geometry %d x %d x %d", a, c, b);
  if (g && (!norms || !workspace)) return fail(GANB_E_BADARG, "weight_transform_fwd: weight-norm needs norms + workspace");
  const dim3 grid = wt_grid(a, c, b);
  float* partial = static_cast<float*>(workspace);
  if (g) {
    launch_k(wt_partial_kernel<false>, grid, 256, 0, STREAM, w, static_cast<const float*>(nullptr), mask, partial, a, c, b);
    GANB_CHECK_LAUNCH("wt_partial_kernel<fwd>");
  }
  launch_k(wt_apply_kernel<false>, grid, 256, 0, STREAM, w, static_cast<const float*>(nullptr), g, mask,
           static_cast<const float*>(partial), static_cast<int>(grid.y), norms, w_eff, static_cast<float*>(nullptr), a, c, b);
  GANB_CHECK_LAUNCH("wt_apply_kernel<fwd>");
  return 0;
}

extern "C" int ganb_weight_transform_bwd(const float* w, const float* dw_eff, const float* g, const float* mask,
                                         const float* norms, float* dw, float* dg, void* workspace, int a, int c, int b,
                                         void* stream) {
  if (!w || !dw_eff || !dw) return fail(GANB_E_BADARG, "weight_transform_bwd: null buffer");
  if (a < 1 || c < 1 || b < 1) return fail(GANB_E_BADARG, "weight_transform_bwd: bad geometry %d x %d x %d", a, c, b);
  if (g && (!norms || !workspace)) return fail(GANB_E_BADARG, "weight_transform_bwd: weight-norm needs norms + workspace");
  const dim3 grid = wt_grid(a, c, b);
  float* partial = static_cast<float*>(workspace);
  if (g) {
    launch_k(wt_partial_kernel<true>, grid, 256, 0, STREAM, w, dw_eff, mask, partial, a, c, b);
    GANB_CHECK_LAUNCH("wt_partial_kernel<bwd>");
  }
  launch_k(wt_apply_kernel<true>, grid, 256, 0, STREAM, w, dw_eff, g, mask, static_cast<const float*>(partial),
           static_cast<int>(grid.y), const_cast<float*>(norms), dw, dg, a, c, b);
  GANB_CHECK_LAUNCH("wt_apply_kernel<bwd>");
  return 0;
}

extern "C" int64_t ganb_layer_norm_workspace(int n, int64_t per_sample, int c) {
  // forward: [n][splits][2] chunk statistics; backward: [n][splits][2] sums + [2][n*splits][c] channel partials
  const int64_t rows = static_cast<int64_t>(n) * ln_splits(per_sample);
  return rows * 2 * 4 + 2 * rows * c * 4 + 64;
}

extern "C" int64_t ganb_layer_norm_rows(int n, int64_t per_sample) {
  return static_cast<int64_t>(n) * ln_splits(per_sample);
}

extern "C" int ganb_layer_norm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype,
                                   float* mean_rstd, void* workspace, int n, int64_t per_sample, int c, float eps, int act,
                                   void* stream) {
  if (!x || !gamma || !beta || !y || !mean_rstd || !workspace) return fail(GANB_E_BADARG, "layer_norm_fwd: null buffer");
  if (n < 1 || !ln_shape_ok(per_sample, c))
    return fail(GANB_E_UNSUPPORTED, "layer_norm_fwd: channels must divide 1024 (got c=%d, %lld per sample)", c,
                static_cast<long long>(per_sample));
  const int splits = ln_splits(per_sample);
  const dim3 grid(splits, n);
  float* partial = static_cast<float*>(workspace);
#define LN_FWD(TI, TO)                                                                                                 \
  do {                                                                                                                 \
    launch_k(ln_stats_partial_kernel<TI>, grid, 256, 0, STREAM, static_cast<const TI*>(x), partial, per_sample, splits); \
    GANB_CHECK_LAUNCH("ln_stats_partial_kernel");                                                                      \
    launch_k(ln_fwd_apply_kernel<TI, TO>, grid, 256, 0, STREAM, static_cast<const TI*>(x),                            \
             static_cast<const float*>(partial), gamma, beta, static_cast<TO*>(y), mean_rstd, per_sample, splits, c,  \
             eps, act);                                                                                                \
    GANB_CHECK_LAUNCH("ln_fwd_apply_kernel");                                                                          \
    return 0;                                                                                                          \
  } while (0)
  if (x_dtype == GANB_F32 && y_dtype == GANB_F32) LN_FWD(float, float);
  if (x_dtype == GANB_F32 && y_dtype == GANB_BF16) LN_FWD(float, __nv_bfloat16);
  if (x_dtype == GANB_BF16 && y_dtype == GANB_F32) LN_FWD(__nv_bfloat16, float);
  LN_FWD(__nv_bfloat16, __nv_bfloat16);
#undef LN_FWD
}

namespace ganb {
template <typename TX, typename TDY>
static int ln_bwd_launch(const void* x, const void* dy, const float* mean_rstd, const float* gamma, const float* beta,
                         void* dx, int dx_dtype, float* sums, float* chan, int n, int64_t per_sample, int c, int act,
                         cudaStream_t s) {
  const int splits = ln_splits(per_sample);
  const dim3 grid(splits, n);
  const int64_t rows = static_cast<int64_t>(n) * splits;
  launch_k(ln_bwd_partial_kernel<TX, TDY>, grid, 256, 0, s, static_cast<const TX*>(x), static_cast<const TDY*>(dy),
           mean_rstd, gamma, beta, sums, chan, per_sample, splits, c, act, rows);
  GANB_CHECK_LAUNCH("ln_bwd_partial_kernel");
  if (dx == nullptr) return 0;
  if (dx_dtype == GANB_F32) {
    launch_k(ln_bwd_apply_kernel<TX, TDY, float>, grid, 256, 0, s, static_cast<const TX*>(x),
             static_cast<const TDY*>(dy), mean_rstd, gamma, beta, static_cast<const float*>(sums),
             static_cast<float*>(dx), per_sample, splits, c, act);
  } else {
    launch_k(ln_bwd_apply_kernel<TX, TDY, __nv_bfloat16>, grid, 256, 0, s, static_cast<const TX*>(x),
             static_cast<const TDY*>(dy), mean_rstd, gamma, beta, static_cast<const float*>(sums),
             static_cast<__nv_bfloat16*>(dx), per_sample, splits, c, act);
  }
  GANB_CHECK_LAUNCH("ln_bwd_apply_kernel");
  return 0;
}
}  // namespace ganb

extern "C" int ganb_layer_norm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* mean_rstd,
                                   const float* gamma, const float* beta, void* dx, int dx_dtype, float* chan_partials,
                                   void* workspace, int n, int64_t per_sample, int c, int act, void* stream) {
  if (!x || !dy || !mean_rstd || !gamma || !beta || !chan_partials || !workspace)
    return fail(GANB_E_BADARG, "layer_norm_bwd: null buffer");
  if (n < 1 || !ln_shape_ok(per_sample, c))
    return fail(GANB_E_UNSUPPORTED, "layer_norm_bwd: channels must divide 1024 (got c=%d)", c);
  float* sums = static_cast<float*>(workspace);
  if (x_dtype == GANB_F32 && dy_dtype == GANB_F32)
    return ln_bwd_launch<float, float>(x, dy, mean_rstd, gamma, beta, dx, dx_dtype, sums, chan_partials, n, per_sample, c, act, STREAM);
  if (x_dtype == GANB_F32 && dy_dtype == GANB_BF16)
    return ln_bwd_launch<float, __nv_bfloat16>(x, dy, mean_rstd, gamma, beta, dx, dx_dtype, sums, chan_partials, n, per_sample, c, act, STREAM);
  if (x_dtype == GANB_BF16 && dy_dtype == GANB_F32)
    return ln_bwd_launch<__nv_bfloat16, float>(x, dy, mean_rstd, gamma, beta, dx, dx_dtype, sums, chan_partials, n, per_sample, c, act, STREAM);
  return ln_bwd_launch<__nv_bfloat16, __nv_bfloat16>(x, dy, mean_rstd, gamma, beta, dx, dx_dtype, sums, chan_partials, n, per_sample, c, act, STREAM);
}

extern "C" int ganb_lerp_fwd(const float* a, const float* b, float* y, int64_t count, const float* alpha, void* stream) {
  if (!a || !b || !y || !alpha) return fail(GANB_E_BADARG, "lerp_fwd: null buffer");
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15)
    return fail(GANB_E_BADARG, "lerp_fwd: buffers must be 16-byte aligned");
  launch_k(lerp_fwd_kernel, flat_grid(count / 4 + 1), 256, 0, STREAM, a, b, y, count, alpha);
  GANB_CHECK_LAUNCH("lerp_fwd_kernel");
  return 0;
}

extern "C" int ganb_round_tf32(const float* x, float* y, int64_t count, void* stream) {
  if (!x || !y) return fail(GANB_E_BADARG, "round_tf32: null buffer");
  launch_k(round_tf32_kernel, flat_grid(count), 256, 0, STREAM, x, y, count);
  GANB_CHECK_LAUNCH("round_tf32_kernel");
  return 0;
}

extern "C" int ganb_transpose_tf32(const float* w_hwio, float* wt, int taps, int cin, int cout, void* stream) {
  if (!w_hwio || !wt || taps <= 0 || cin <= 0 || cout <= 0) return fail(GANB_E_BADARG, "transpose_tf32: bad arguments");
  launch_k(transpose_tf32_kernel, dim3(((cin + 31) / 32) * ((cout + 31) / 32), taps), 256, 0, STREAM, w_hwio, wt, cin, cout);
  GANB_CHECK_LAUNCH("transpose_tf32_kernel");
  return 0;
}

extern "C" int ganb_scale_dev(const float* x, const float* alpha, float* y, int64_t count, void* stream) {
  if (!x || !alpha || !y) return fail(GANB_E_BADARG, "scale_dev: null buffer");
  launch_k(scale_dev_kernel, flat_grid(count), 256, 0, STREAM, x, alpha, y, count);
  GANB_CHECK_LAUNCH("scale_dev_kernel");
  return 0;
}

extern "C" int ganb_lerp_bwd(const float* dy, void* da, int da_dtype, void* db, int db_dtype, int64_t count,
                             const float* alpha, void* stream) {
  if (!dy || !alpha || (!da && !db)) return fail(GANB_E_BADARG, "lerp_bwd: null buffer");
  const int grid = flat_grid(count / 4 + 1);
#define LERP_BWD(TA, TB)                                                                                             \
  launch_k(lerp_bwd_kernel<TA, TB>, grid, 256, 0, STREAM, dy, static_cast<TA*>(da), static_cast<TB*>(db), count, alpha)
  if (da_dtype == GANB_F32 && db_dtype == GANB_F32) LERP_BWD(float, float);
  else if (da_dtype == GANB_F32) LERP_BWD(float, __nv_bfloat16);
  else if (db_dtype == GANB_F32) LERP_BWD(__nv_bfloat16, float);
  else LERP_BWD(__nv_bfloat16, __nv_bfloat16);
#undef LERP_BWD
  GANB_CHECK_LAUNCH("lerp_bwd_kernel");
  return 0;
}

extern "C" int ganb_subsample2d(const void* x, int x_dtype, void* y, int y_dtype, int n, int h, int w, int c, int stride,
                                int scatter, void* stream) {
  if (!x || !y) return fail(GANB_E_BADARG, "subsample2d: null buffer");
  if (n < 1 || h < 1 || w < 1 || c < 1 || stride < 1) return fail(GANB_E_BADARG, "subsample2d: bad shape");
  // forward: x [n,h,w,c] -> y [n,ceil(h/s),ceil(w/s),c]; scatter: x [n,ceil(h/s),ceil(w/s),c] -> y [n,h,w,c]
  const int sh = (h + stride - 1) / stride, sw = (w + stride - 1) / stride;
  const int64_t total = static_cast<int64_t>(n) * (scatter ? h : sh) * (scatter ? w : sw) * c;
  const int grid = flat_grid(total);
#define SUBS(TI, TO)                                                                                                  \
  do {                                                                                                                \
    if (scatter)                                                                                                      \
      launch_k(subsample2d_kernel<TI, TO, true>, grid, 256, 0, STREAM, static_cast<const TI*>(x), static_cast<TO*>(y), \
               sh, sw, c, h, w, stride, total);                                                                       \
    else                                                                                                              \
      launch_k(subsample2d_kernel<TI, TO, false>, grid, 256, 0, STREAM, static_cast<const TI*>(x), static_cast<TO*>(y), \
               h, w, c, sh, sw, stride, total);                                                                       \
  } while (0)
  if (x_dtype == GANB_F32 && y_dtype == GANB_F32) SUBS(float, float);
  else if (x_dtype == GANB_F32) SUBS(float, __nv_bfloat16);
  else if (y_dtype == GANB_F32) SUBS(__nv_bfloat16, float);
  else SUBS(__nv_bfloat16, __nv_bfloat16);
#undef SUBS
  GANB_CHECK_LAUNCH("subsample2d_kernel");
  return 0;
}

extern "C" int ganb_sample_grid(const void* samples, int dtype, int n, int h, int w, int c, int nw, unsigned char* grid,
                                void* stream) {
  if (!samples || !grid) return fail(GANB_E_BADARG, "sample_grid: null buffer");
  if (n < 1 || h < 1 || w < 1 || c < 1 || nw < 1) return fail(GANB_E_BADARG, "sample_grid: bad shape");
  const int nh = (n + nw - 1) / nw;
  const int64_t total = static_cast<int64_t>(nh) * h * nw * w * c;
  if (dtype == GANB_F32)
    launch_k(sample_grid_kernel<float>, flat_grid(total), 256, 0, STREAM, static_cast<const float*>(samples), grid, n, h,
             w, c, nw, total);
  else
    launch_k(sample_grid_kernel<__nv_bfloat16>, flat_grid(total), 256, 0, STREAM,
             static_cast<const __nv_bfloat16*>(samples), grid, n, h, w, c, nw, total);
  GANB_CHECK_LAUNCH("sample_grid_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------- depthwise conv entries
static bool dw_args_ok(int n, int h, int w, int c, int cm, int ho, int wo, int kh, int kw, int stride) {
  return n >= 1 && h >= 1 && w >= 1 && c >= 1 && cm >= 1 && ho >= 1 && wo >= 1 && kh >= 1 && kw >= 1 && stride >= 1;
}

extern "C" int ganb_depthwise_conv2d_fwd(const void* x, int x_dtype, const float* filter, const float* bias, void* y,
                                         int y_dtype, int n, int h, int w, int c, int cm, int ho, int wo, int kh, int kw,
                                         int stride, int pad_t, int pad_l, void* stream) {
  if (!x || !filter || !y) return fail(GANB_E_BADARG, "depthwise_conv2d_fwd: null buffer");
  if (!dw_args_ok(n, h, w, c, cm, ho, wo, kh, kw, stride)) return fail(GANB_E_BADARG, "depthwise_conv2d_fwd: bad shape");
  const int64_t total = static_cast<int64_t>(n) * ho * wo * c * cm;
  if (cm == 1 && c % 8 == 0 && x_dtype == GANB_BF16) {   // 8 channels per thread
    const int64_t total8 = total / 8;
    if (y_dtype == GANB_F32)
      launch_k(depthwise_fwd_v8_kernel<float>, flat_grid(total8), 256, 0, STREAM, static_cast<const __nv_bfloat16*>(x),
               filter, bias, static_cast<float*>(y), h, w, c, ho, wo, kh, kw, stride, pad_t, pad_l, total8);
    else
      launch_k(depthwise_fwd_v8_kernel<__nv_bfloat16>, flat_grid(total8), 256, 0, STREAM,
               static_cast<const __nv_bfloat16*>(x), filter, bias, static_cast<__nv_bfloat16*>(y), h, w, c, ho, wo, kh,
               kw, stride, pad_t, pad_l, total8);
    GANB_CHECK_LAUNCH("depthwise_fwd_v8_kernel");
    return 0;
  }
#define DW_FWD(TI, TO)                                                                                                \
  launch_k(depthwise_fwd_kernel<TI, TO>, flat_grid(total), 256, 0, STREAM, static_cast<const TI*>(x), filter, bias,   \
           static_cast<TO*>(y), h, w, c, cm, ho, wo, kh, kw, stride, pad_t, pad_l, total)
  if (x_dtype == GANB_F32 && y_dtype == GANB_F32) DW_FWD(float, float);
  else if (x_dtype == GANB_F32) DW_FWD(float, __nv_bfloat16);
  else if (y_dtype == GANB_F32) DW_FWD(__nv_bfloat16, float);
  else DW_FWD(__nv_bfloat16, __nv_bfloat16);
#undef DW_FWD
  GANB_CHECK_LAUNCH("depthwise_fwd_kernel");
  return 0;
}

extern "C" int ganb_depthwise_conv2d_bwd_input(const void* dy, int dy_dtype, const float* filter, void* dx, int dx_dtype,
                                               int n, int h, int w, int c, int cm, int ho, int wo, int kh, int kw,
                                               int stride, int pad_t, int pad_l, void* stream) {
  if (!dy || !filter || !dx) return fail(GANB_E_BADARG, "depthwise_conv2d_bwd_input: null buffer");
  if (!dw_args_ok(n, h, w, c, cm, ho, wo, kh, kw, stride))
    return fail(GANB_E_BADARG, "depthwise_conv2d_bwd_input: bad shape");
  const int64_t total = static_cast<int64_t>(n) * h * w * c;
  if (cm == 1 && c % 8 == 0) {
    const int64_t total8 = total / 8;
#define DW_BI8(TD, TO)                                                                                                \
  launch_k(depthwise_bwd_input_v8_kernel<TD, TO>, flat_grid(total8), 256, 0, STREAM, static_cast<const TD*>(dy), filter, \
           static_cast<TO*>(dx), h, w, c, ho, wo, kh, kw, stride, pad_t, pad_l, total8)
    if (dy_dtype == GANB_F32 && dx_dtype == GANB_F32) DW_BI8(float, float);
    else if (dy_dtype == GANB_F32) DW_BI8(float, __nv_bfloat16);
    else if (dx_dtype == GANB_F32) DW_BI8(__nv_bfloat16, float);
    else DW_BI8(__nv_bfloat16, __nv_bfloat16);
#undef DW_BI8
    GANB_CHECK_LAUNCH("depthwise_bwd_input_v8_kernel");
    return 0;
  }
#define DW_BI(TD, TO)                                                                                                 \
  launch_k(depthwise_bwd_input_kernel<TD, TO>, flat_grid(total), 256, 0, STREAM, static_cast<const TD*>(dy), filter,  \
           static_cast<TO*>(dx), h, w, c, cm, ho, wo, kh, kw, stride, pad_t, pad_l, total)
  if (dy_dtype == GANB_F32 && dx_dtype == GANB_F32) DW_BI(float, float);
  else if (dy_dtype == GANB_F32) DW_BI(float, __nv_bfloat16);
  else if (dx_dtype == GANB_F32) DW_BI(__nv_bfloat16, float);
  else DW_BI(__nv_bfloat16, __nv_bfloat16);
#undef DW_BI
  GANB_CHECK_LAUNCH("depthwise_bwd_input_kernel");
  return 0;
}

static int dw_chunks(int64_t pixels) {
  int64_t ch = ceil_div64(pixels, 64);       // >= 64 pixels per block
  if (ch > DW_CHUNKS_MAX) ch = DW_CHUNKS_MAX;
  return static_cast<int>(ch < 1 ? 1 : ch);
}

extern "C" int64_t ganb_depthwise_conv2d_chunks(int n, int ho, int wo) {
  return dw_chunks(static_cast<int64_t>(n) * ho * wo);
}

extern "C" int ganb_depthwise_conv2d_bwd_filter(const void* x, int x_dtype, const void* dy, int dy_dtype,
                                                float* partials, int n, int h, int w, int c, int cm, int ho, int wo,
                                                int kh, int kw, int stride, int pad_t, int pad_l, void* stream) {
  if (!x || !dy || !partials) return fail(GANB_E_BADARG, "depthwise_conv2d_bwd_filter: null buffer");
  if (!dw_args_ok(n, h, w, c, cm, ho, wo, kh, kw, stride))
    return fail(GANB_E_BADARG, "depthwise_conv2d_bwd_filter: bad shape");
  const int64_t pixels = static_cast<int64_t>(n) * ho * wo;
  const int chunks = dw_chunks(pixels);
  const int64_t per_chunk = ceil_div64(pixels, chunks);
  if (cm == 1 && c % 8 == 0 && c <= 2048 && x_dtype == GANB_BF16 && kh == kw && (kh == 3 || kh == 4)) {
#define DW_BF8(KS, TD)                                                                                                \
  launch_k(depthwise_bwd_filter_v8_kernel<KS, TD>, chunks, 256, 0, STREAM, static_cast<const __nv_bfloat16*>(x),      \
           static_cast<const TD*>(dy), partials, h, w, c, ho, wo, stride, pad_t, pad_l, pixels, per_chunk)
    if (kh == 3 && dy_dtype == GANB_F32) DW_BF8(3, float);
    else if (kh == 3) DW_BF8(3, __nv_bfloat16);
    else if (dy_dtype == GANB_F32) DW_BF8(4, float);
    else DW_BF8(4, __nv_bfloat16);
#undef DW_BF8
    GANB_CHECK_LAUNCH("depthwise_bwd_filter_v8_kernel");
    return 0;
  }
  const dim3 grid(chunks, kh * kw);
#define DW_BF(TI, TD)                                                                                                 \
  launch_k(depthwise_bwd_filter_kernel<TI, TD>, grid, 256, 0, STREAM, static_cast<const TI*>(x),                      \
           static_cast<const TD*>(dy), partials, h, w, c, cm, ho, wo, kw, stride, pad_t, pad_l, pixels, per_chunk)
  if (x_dtype == GANB_F32 && dy_dtype == GANB_F32) DW_BF(float, float);
  else if (x_dtype == GANB_F32) DW_BF(float, __nv_bfloat16);
  else if (dy_dtype == GANB_F32) DW_BF(__nv_bfloat16, float);
  else DW_BF(__nv_bfloat16, __nv_bfloat16);
#undef DW_BF
  GANB_CHECK_LAUNCH("depthwise_bwd_filter_kernel");
  return 0;
}
