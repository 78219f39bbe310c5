// Bandwidth-bound kernels of the PGGAN / Pix2Pix sides of the hot path:
// pixel-norm (+ leaky-ReLU), minibatch standard deviation, channel concatenation (U-Net skips), dropout masks.
// NHWC, fp32 or bf16 storage, fp32 arithmetic, channels % 4 == 0.
//
// Reference call-sites replaced: common/ops/normalization.py:125-140 (pixel_norm), PGGAN/model_nvidia.py:15-28
// (lrelu, minibatch_std), Pix2Pix/networks.py:263-264, 268-270 (tf.nn.dropout, tf.concat of skip connections).
#include "host_common.h"

#include <cuda_bf16.h>

namespace ganb {

struct alignas(8) bf16x4e {
  __nv_bfloat162 lo, hi;
};
__device__ __forceinline__ float4 eld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 eld4(const __nv_bfloat16* p) {
  const bf16x4e v = *reinterpret_cast<const bf16x4e*>(p);
  const float2 a = __bfloat1622float2(v.lo), b = __bfloat1622float2(v.hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void est4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void est4(__nv_bfloat16* p, float4 v) {
  bf16x4e o;
  o.lo = __floats2bfloat162_rn(v.x, v.y);
  o.hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<bf16x4e*>(p) = o;
}
__device__ __forceinline__ float eact(float v, int act) {
  if (act == GANB_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == GANB_ACT_LRELU) return v >= 0.f ? v : 0.2f * v;
  return v;
}
__device__ __forceinline__ float edact(float z, int act) {
  if (act == GANB_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == GANB_ACT_LRELU) return z >= 0.f ? 1.f : 0.2f;
  return 1.f;
}
__device__ __forceinline__ float ewarp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static inline int egrid(int64_t items, int threads, int per_sm = 8) {
  int64_t b = ceil_div64(items, threads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------------ pixel norm
// y = act(x * rsqrt(mean_c(x^2) + eps)); one warp per pixel, lanes stride over float4 channel groups.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
pixel_norm_fwd_kernel(const TIn* __restrict__ x, TOut* __restrict__ y, int64_t pixels, int c, float eps, int act) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int v = c >> 2;
  for (int64_t pix = blockIdx.x * 8LL + (threadIdx.x >> 5); pix < pixels; pix += gridDim.x * 8LL) {
    const TIn* xp = x + pix * c;
    float ss = 0.f;
    for (int j = lane; j < v; j += 32) {
      const float4 a = eld4(xp + j * 4);
      ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    ss = ewarp_sum(ss);
    const float r = 1.0f / sqrtf(ss / c + eps);
    for (int j = lane; j < v; j += 32) {
      const float4 a = eld4(xp + j * 4);
      est4(y + pix * c + j * 4, make_float4(eact(a.x * r, act), eact(a.y * r, act), eact(a.z * r, act), eact(a.w * r, act)));
    }
  }
}
// dz = dy * act'(x r);  dx = r dz - x r^3 mean_c(dz x)
template <typename TIn, typename TG, typename TOut>
__global__ void __launch_bounds__(256)
pixel_norm_bwd_kernel(const TIn* __restrict__ x, const TG* __restrict__ dy, TOut* __restrict__ dx, int64_t pixels,
                      int c, float eps, int act) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int v = c >> 2;
  for (int64_t pix = blockIdx.x * 8LL + (threadIdx.x >> 5); pix < pixels; pix += gridDim.x * 8LL) {
    const TIn* xp = x + pix * c;
    const TG* gp = dy + pix * c;
    float ss = 0.f;
    for (int j = lane; j < v; j += 32) {
      const float4 a = eld4(xp + j * 4);
      ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    ss = ewarp_sum(ss);
    const float r = 1.0f / sqrtf(ss / c + eps);
    float dot = 0.f;
    for (int j = lane; j < v; j += 32) {
      const float4 a = eld4(xp + j * 4), g = eld4(gp + j * 4);
      dot += g.x * edact(a.x * r, act) * a.x + g.y * edact(a.y * r, act) * a.y + g.z * edact(a.z * r, act) * a.z +
             g.w * edact(a.w * r, act) * a.w;
    }
    dot = ewarp_sum(dot);
    const float k = r * r * r * dot / c;
    for (int j = lane; j < v; j += 32) {
      const float4 a = eld4(xp + j * 4), g = eld4(gp + j * 4);
      est4(dx + pix * c + j * 4,
           make_float4(r * g.x * edact(a.x * r, act) - a.x * k, r * g.y * edact(a.y * r, act) - a.y * k,
                       r * g.z * edact(a.z * r, act) - a.z * k, r * g.w * edact(a.w * r, act) - a.w * k));
    }
  }
}

// ------------------------------------------------------------------------------------------------ minibatch std
// PGGAN/model_nvidia.py:20-28.  m = mean_b x, v = mean_b (x - m)^2 per (h, w, c); s = mean_{h,w,c} sqrt(v + 1e-8);
// out[b, h, w, 0:c] = x, out[b, h, w, c] = s  (out channel stride = cs >= c + 1).
// stage 1: one thread per (h, w, c) position -> sd[pos] = sqrt(v + eps) and block partial sums
__global__ void __launch_bounds__(256)
mbstd_stats_kernel(const float* __restrict__ x, int b, int64_t m, float eps, float* __restrict__ sd,
                   float* __restrict__ partial) {
  pdl_wait();
  __shared__ float sh[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < m; i += gridDim.x * 256LL) {
    float mean = 0.f;
    for (int k = 0; k < b; ++k) mean += x[k * m + i];
    mean /= b;
    float var = 0.f;
    for (int k = 0; k < b; ++k) {
      const float d = x[k * m + i] - mean;
      var += d * d;
    }
    const float s = sqrtf(var / b + eps);
    sd[i] = s;
    acc += s;
  }
  acc = ewarp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}
// stage 2: s = sum(partial) / m (fixed order); out = concat(x, s)
__global__ void __launch_bounds__(256)
mbstd_concat_kernel(const float* __restrict__ x, const float* __restrict__ partial, int nparts, int64_t pixels, int c,
                    int cs, int64_t m, float* __restrict__ out, float* __restrict__ s_out) {
  pdl_wait();
  float s = 0.f;
  for (int k = 0; k < nparts; ++k) s += partial[k];
  s /= static_cast<float>(m);
  if (blockIdx.x == 0 && threadIdx.x == 0) s_out[0] = s;
  const int64_t total = pixels * cs;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ch = static_cast<int>(i % cs);
    const int64_t pix = i / cs;
    out[i] = ch < c ? x[pix * c + ch] : (ch == c ? s : 0.f);
  }
}
// backward: dx = dout[..., 0:c] + g * (x - mean_b x) / (b * m * sd), g = sum over (b,h,w) of dout[..., c]
__global__ void __launch_bounds__(256)
mbstd_gsum_kernel(const float* __restrict__ dout, int64_t pixels, int c, int cs, float* __restrict__ g) {
  pdl_wait();
  __shared__ float sh[8];
  float acc = 0.f;   // single block: the map is tiny (PGGAN applies it at 4x4)
  for (int64_t p = threadIdx.x; p < pixels; p += 256) acc += dout[p * cs + c];
  acc = ewarp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    g[0] = t;
  }
}
__global__ void __launch_bounds__(256)
mbstd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, const float* __restrict__ sd,
                 const float* __restrict__ g, int b, int64_t m, int c, int cs, float* __restrict__ dx) {
  pdl_wait();
  const float gs = g[0] / (static_cast<float>(b) * static_cast<float>(m));
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < m; i += gridDim.x * 256LL) {
    float mean = 0.f;
    for (int k = 0; k < b; ++k) mean += x[k * m + i];
    mean /= b;
    const float inv = gs / sd[i];
    const int64_t pos = i / c;
    const int ch = static_cast<int>(i % c);
    for (int k = 0; k < b; ++k) {
      const int64_t pix = k * (m / c) + pos;
      dx[k * m + i] = dout[pix * cs + ch] + (x[k * m + i] - mean) * inv;
    }
  }
}

// ---- cross-GPU variant (the statistic then spans the GLOBAL batch, like a synced batch norm; SURVEY 8(e) collective 3):
// phase 1 leaves the local [sum_b x | sum_b x^2] per position for an all-reduce, phase 2 turns the global sums into
// mean / sd; the backward pass all-reduces the scalar g and uses the global batch size and the global mean.
__global__ void __launch_bounds__(256)
mbstd_moments_kernel(const float* __restrict__ x, int b, int64_t m, float* __restrict__ sums) {
  pdl_wait();
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < m; i += gridDim.x * 256LL) {
    float s = 0.f, q = 0.f;
    for (int k = 0; k < b; ++k) {
      const float v = x[k * m + i];
      s += v;
      q += v * v;
    }
    sums[i] = s;
    sums[m + i] = q;
  }
}
__global__ void __launch_bounds__(256)
mbstd_global_sd_kernel(const float* __restrict__ sums, int64_t m, float inv_batch, float eps, float* __restrict__ sd,
                       float* __restrict__ mean, float* __restrict__ partial) {
  pdl_wait();
  __shared__ float sh[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < m; i += gridDim.x * 256LL) {
    const float mu = sums[i] * inv_batch;
    float var = sums[m + i] * inv_batch - mu * mu;
    if (var < 0.f) var = 0.f;
    const float s = sqrtf(var + eps);
    mean[i] = mu;
    sd[i] = s;
    acc += s;
  }
  acc = ewarp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256)
mbstd_bwd_global_kernel(const float* __restrict__ x, const float* __restrict__ dout, const float* __restrict__ sd,
                        const float* __restrict__ mean, const float* __restrict__ g, int b, float batch_total, int64_t m,
                        int c, int cs, float* __restrict__ dx) {
  pdl_wait();
  const float gs = g[0] / (batch_total * static_cast<float>(m));
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < m; i += gridDim.x * 256LL) {
    const float inv = gs / sd[i], mu = mean[i];
    const int64_t pos = i / c;
    const int ch = static_cast<int>(i % c);
    for (int k = 0; k < b; ++k) {
      const int64_t pix = k * (m / c) + pos;
      dx[k * m + i] = dout[pix * cs + ch] + (x[k * m + i] - mu) * inv;
    }
  }
}

// ------------------------------------------------------------------------------------------------ concat / slice
// dst[pix, off + j] = scale * mask[pix, j] * src[pix, j]   (mask optional): channel concatenation, its backward
// (a strided slice), and dropout (mask in {0, 1}, scale = 1 / keep_prob) share one kernel.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
copy_channels_kernel(const TIn* __restrict__ src, int src_cs, int src_off, TOut* __restrict__ dst, int dst_cs, int dst_off,
                     int64_t pixels, int c, const float* __restrict__ mask, float scale) {
  pdl_wait();
  const int v = c >> 2;
  const int64_t total = pixels * v;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    float4 a = eld4(src + pix * src_cs + src_off + c4);
    if (mask) {
      const float4 mk = eld4(mask + pix * c + c4);
      a.x *= mk.x; a.y *= mk.y; a.z *= mk.z; a.w *= mk.w;
    }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    est4(dst + pix * dst_cs + dst_off + c4, a);
  }
}

// channel counts / offsets that are not multiples of 4 (RGB tensors)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
copy_channels_scalar_kernel(const TIn* __restrict__ src, int src_cs, int src_off, TOut* __restrict__ dst, int dst_cs,
                            int dst_off, int64_t pixels, int c, const float* __restrict__ mask, float scale) {
  pdl_wait();
  const int64_t total = pixels * c;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ch = static_cast<int>(i % c);
    const int64_t pix = i / c;
    float a = static_cast<float>(src[pix * src_cs + src_off + ch]);
    if (mask) a *= mask[i];
    dst[pix * dst_cs + dst_off + ch] = static_cast<TOut>(a * scale);
  }
}

}  // namespace ganb

using namespace ganb;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ganb_pixel_norm_fwd(const void* x, int x_dtype, void* y, int y_dtype, int64_t pixels, int c, float eps,
                                   int act, void* stream) {
  if (!x || !y) return fail(GANB_E_BADARG, "pixel_norm_fwd: null buffer");
  if (c % 4) return fail(GANB_E_UNSUPPORTED, "pixel_norm_fwd: c=%d must be a multiple of 4", c);
  const int grid = egrid(pixels * 32, 256);
  const bool xi = x_dtype == GANB_BF16, yo = y_dtype == GANB_BF16;
  if (!xi && !yo) launch_k(pixel_norm_fwd_kernel<float, float>, grid, 256, 0, STREAM, static_cast<const float*>(x), static_cast<float*>(y), pixels, c, eps, act);
  else if (!xi) launch_k(pixel_norm_fwd_kernel<float, __nv_bfloat16>, grid, 256, 0, STREAM, static_cast<const float*>(x), static_cast<__nv_bfloat16*>(y), pixels, c, eps, act);
  else if (!yo) launch_k(pixel_norm_fwd_kernel<__nv_bfloat16, float>, grid, 256, 0, STREAM, static_cast<const __nv_bfloat16*>(x), static_cast<float*>(y), pixels, c, eps, act);
  else launch_k(pixel_norm_fwd_kernel<__nv_bfloat16, __nv_bfloat16>, grid, 256, 0, STREAM, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), pixels, c, eps, act);
  GANB_CHECK_LAUNCH("pixel_norm_fwd_kernel");
  return 0;
}

namespace ganb {
template <typename TIn, typename TG>
static void launch_pn_bwd(const void* x, const void* dy, void* dx, int dx_dtype, int64_t pixels, int c, float eps, int act,
                          int grid, cudaStream_t s) {
  if (dx_dtype == GANB_BF16)
    launch_k(pixel_norm_bwd_kernel<TIn, TG, __nv_bfloat16>, grid, 256, 0, s, static_cast<const TIn*>(x), static_cast<const TG*>(dy),
             static_cast<__nv_bfloat16*>(dx), pixels, c, eps, act);
  else
    launch_k(pixel_norm_bwd_kernel<TIn, TG, float>, grid, 256, 0, s, static_cast<const TIn*>(x), static_cast<const TG*>(dy),
             static_cast<float*>(dx), pixels, c, eps, act);
}
}  // namespace ganb

extern "C" int ganb_pixel_norm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype,
                                   int64_t pixels, int c, float eps, int act, void* stream) {
  if (!x || !dy || !dx) return fail(GANB_E_BADARG, "pixel_norm_bwd: null buffer");
  if (c % 4) return fail(GANB_E_UNSUPPORTED, "pixel_norm_bwd: c=%d must be a multiple of 4", c);
  const int grid = egrid(pixels * 32, 256);
  const bool xi = x_dtype == GANB_BF16, gi = dy_dtype == GANB_BF16;
  if (!xi && !gi) launch_pn_bwd<float, float>(x, dy, dx, dx_dtype, pixels, c, eps, act, grid, STREAM);
  else if (!xi) launch_pn_bwd<float, __nv_bfloat16>(x, dy, dx, dx_dtype, pixels, c, eps, act, grid, STREAM);
  else if (!gi) launch_pn_bwd<__nv_bfloat16, float>(x, dy, dx, dx_dtype, pixels, c, eps, act, grid, STREAM);
  else launch_pn_bwd<__nv_bfloat16, __nv_bfloat16>(x, dy, dx, dx_dtype, pixels, c, eps, act, grid, STREAM);
  GANB_CHECK_LAUNCH("pixel_norm_bwd_kernel");
  return 0;
}

extern "C" int64_t ganb_minibatch_std_workspace(int b, int h, int w, int c) {
  (void)b;
  return (static_cast<int64_t>(h) * w * c + 1024 + 8) * 4;
}

// out [b,h,w,cs] (cs >= c+1; channels above c are zero-filled), sd = workspace (kept for the backward pass)
extern "C" int ganb_minibatch_std_fwd(const float* x, int b, int h, int w, int c, int cs, float* out, void* workspace,
                                      void* stream) {
  if (!x || !out || !workspace) return fail(GANB_E_BADARG, "minibatch_std_fwd: null buffer");
  if (cs < c + 1) return fail(GANB_E_BADARG, "minibatch_std_fwd: output stride %d < c+1", cs);
  const int64_t m = static_cast<int64_t>(h) * w * c;
  float* sd = static_cast<float*>(workspace);
  float* partial = sd + m;
  float* s_out = partial + 1024;
  int nparts = static_cast<int>(ceil_div64(m, 256));
  if (nparts > 1024) nparts = 1024;
  launch_k(mbstd_stats_kernel, nparts, 256, 0, STREAM, x, b, m, 1e-8f, sd, partial);
  GANB_CHECK_LAUNCH("mbstd_stats_kernel");
  const int64_t pixels = static_cast<int64_t>(b) * h * w;
  launch_k(mbstd_concat_kernel, egrid(pixels * cs, 256), 256, 0, STREAM, x, static_cast<const float*>(partial), nparts, pixels, c,
           cs, m, out, s_out);
  GANB_CHECK_LAUNCH("mbstd_concat_kernel");
  return 0;
}

extern "C" int ganb_minibatch_std_bwd(const float* x, const float* dout, int b, int h, int w, int c, int cs, float* dx,
                                      void* workspace, void* stream) {
  if (!x || !dout || !dx || !workspace) return fail(GANB_E_BADARG, "minibatch_std_bwd: null buffer");
  const int64_t m = static_cast<int64_t>(h) * w * c;
  float* sd = static_cast<float*>(workspace);
  float* g = sd + m + 1024 + 1;
  const int64_t pixels = static_cast<int64_t>(b) * h * w;
  launch_k(mbstd_gsum_kernel, 1, 256, 0, STREAM, dout, pixels, c, cs, g);
  GANB_CHECK_LAUNCH("mbstd_gsum_kernel");
  launch_k(mbstd_bwd_kernel, egrid(m, 256), 256, 0, STREAM, x, dout, static_cast<const float*>(sd), static_cast<const float*>(g), b,
           m, c, cs, dx);
  GANB_CHECK_LAUNCH("mbstd_bwd_kernel");
  return 0;
}

// Cross-GPU minibatch-stddev.  workspace (fp32, ganb_minibatch_std_sync_workspace bytes):
//   [0, 2m) sums = [sum_b x | sum_b x^2] (all-reduced in place by the caller between the phases), [2m, 3m) sd,
//   [3m, 4m) mean, then 1024 block partials, s, g.
extern "C" int64_t ganb_minibatch_std_sync_workspace(int h, int w, int c) {
  return (4LL * h * w * c + 1024 + 8) * 4;
}

extern "C" int64_t ganb_minibatch_std_sync_g_offset(int h, int w, int c) {
  return 4LL * h * w * c + 1024 + 1;    // float index of g inside the workspace
}

extern "C" int ganb_minibatch_std_sync_fwd(const float* x, int b, int h, int w, int c, int cs, float* out, void* workspace,
                                           int phase, int world, void* stream) {
  if (!x || !workspace || (phase == 2 && !out)) return fail(GANB_E_BADARG, "minibatch_std_sync_fwd: null buffer");
  if (phase != 1 && phase != 2) return fail(GANB_E_BADARG, "minibatch_std_sync_fwd: phase must be 1 or 2");
  if (cs < c + 1 || world < 1) return fail(GANB_E_BADARG, "minibatch_std_sync_fwd: bad cs / world");
  const int64_t m = static_cast<int64_t>(h) * w * c;
  float* sums = static_cast<float*>(workspace);
  float* sd = sums + 2 * m;
  float* mean = sd + m;
  float* partial = mean + m;
  float* s_out = partial + 1024;
  int nparts = static_cast<int>(ceil_div64(m, 256));
  if (nparts > 1024) nparts = 1024;
  if (phase == 1) {
    launch_k(mbstd_moments_kernel, nparts, 256, 0, STREAM, x, b, m, sums);
    GANB_CHECK_LAUNCH("mbstd_moments_kernel");
    return 0;
  }
  launch_k(mbstd_global_sd_kernel, nparts, 256, 0, STREAM, static_cast<const float*>(sums), m,
           1.0f / (static_cast<float>(b) * world), 1e-8f, sd, mean, partial);
  GANB_CHECK_LAUNCH("mbstd_global_sd_kernel");
  const int64_t pixels = static_cast<int64_t>(b) * h * w;
  launch_k(mbstd_concat_kernel, egrid(pixels * cs, 256), 256, 0, STREAM, x, static_cast<const float*>(partial), nparts, pixels, c,
           cs, m, out, s_out);
  GANB_CHECK_LAUNCH("mbstd_concat_kernel");
  return 0;
}

extern "C" int ganb_minibatch_std_sync_bwd(const float* x, const float* dout, int b, int h, int w, int c, int cs, float* dx,
                                           void* workspace, int phase, int world, void* stream) {
  if (!x || !dout || !workspace || (phase == 2 && !dx)) return fail(GANB_E_BADARG, "minibatch_std_sync_bwd: null buffer");
  if (phase != 1 && phase != 2) return fail(GANB_E_BADARG, "minibatch_std_sync_bwd: phase must be 1 or 2");
  const int64_t m = static_cast<int64_t>(h) * w * c;
  float* sums = static_cast<float*>(workspace);
  float* sd = sums + 2 * m;
  float* mean = sd + m;
  float* g = mean + m + 1024 + 1;       // all-reduced by the caller between the phases (1 float)
  const int64_t pixels = static_cast<int64_t>(b) * h * w;
  if (phase == 1) {
    launch_k(mbstd_gsum_kernel, 1, 256, 0, STREAM, dout, pixels, c, cs, g);
    GANB_CHECK_LAUNCH("mbstd_gsum_kernel");
    return 0;
  }
  launch_k(mbstd_bwd_global_kernel, egrid(m, 256), 256, 0, STREAM, x, dout, static_cast<const float*>(sd),
           static_cast<const float*>(mean), static_cast<const float*>(g), b, static_cast<float>(b) * world, m, c, cs, dx);
  GANB_CHECK_LAUNCH("mbstd_bwd_global_kernel");
  return 0;
}

namespace ganb {
template <typename TIn, typename TOut>
static void launch_copy_channels(const void* src, int src_cs, int src_off, void* dst, int dst_cs, int dst_off,
                                 int64_t pixels, int c, const float* mask, float scale, cudaStream_t s) {
  if (c % 4 || src_cs % 4 || dst_cs % 4 || src_off % 4 || dst_off % 4) {
    launch_k(copy_channels_scalar_kernel<TIn, TOut>, egrid(pixels * c, 256), 256, 0, s, static_cast<const TIn*>(src), src_cs,
             src_off, static_cast<TOut*>(dst), dst_cs, dst_off, pixels, c, mask, scale);
    return;
  }
  launch_k(copy_channels_kernel<TIn, TOut>, egrid(pixels * (c / 4), 256), 256, 0, s, static_cast<const TIn*>(src), src_cs, src_off,
           static_cast<TOut*>(dst), dst_cs, dst_off, pixels, c, mask, scale);
}
}  // namespace ganb

extern "C" int ganb_copy_channels(const void* src, int src_dtype, int src_cstride, int src_off, void* dst, int dst_dtype,
                                  int dst_cstride, int dst_off, int64_t pixels, int c, const float* mask, float scale,
                                  void* stream) {
  if (!src || !dst) return fail(GANB_E_BADARG, "copy_channels: null buffer");
  const bool si = src_dtype == GANB_BF16, di = dst_dtype == GANB_BF16;
  if (!si && !di) launch_copy_channels<float, float>(src, src_cstride, src_off, dst, dst_cstride, dst_off, pixels, c, mask, scale, STREAM);
  else if (!si) launch_copy_channels<float, __nv_bfloat16>(src, src_cstride, src_off, dst, dst_cstride, dst_off, pixels, c, mask, scale, STREAM);
  else if (!di) launch_copy_channels<__nv_bfloat16, float>(src, src_cstride, src_off, dst, dst_cstride, dst_off, pixels, c, mask, scale, STREAM);
  else launch_copy_channels<__nv_bfloat16, __nv_bfloat16>(src, src_cstride, src_off, dst, dst_cstride, dst_off, pixels, c, mask, scale, STREAM);
  GANB_CHECK_LAUNCH("copy_channels_kernel");
  return 0;
}

namespace ganb {
// L1 reconstruction loss of Pix2Pix (Pix2Pix/train.py:511: tf.reduce_mean(tf.abs(targets - outputs))):
// stage 1 writes d(loss)/d(outputs) = scale * sign(outputs - targets) / count and one partial sum per block,
// stage 2 adds the partials in block order (deterministic).
__global__ void __launch_bounds__(256)
l1_loss_partial_kernel(const float* __restrict__ targets, const float* __restrict__ outputs, int64_t count, float gscale,
                       float* __restrict__ dout, float* __restrict__ partial) {
  pdl_wait();
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < count; i += gridDim.x * 256LL) {
    const float d = outputs[i] - targets[i];
    acc += fabsf(d);
    dout[i] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
  }
  __shared__ float sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256)
l1_loss_finish_kernel(const float* __restrict__ partial, int blocks, float scale_over_count, int accumulate,
                      float* __restrict__ loss_out) {
  pdl_wait();
  __shared__ float sh[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < blocks; i += 256) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (accumulate ? loss_out[0] : 0.f) + sh[0] * scale_over_count;
}
}  // namespace ganb

extern "C" int64_t ganb_l1_loss_workspace(int64_t count) {
  (void)count;
  return 1024 * 4;
}

extern "C" int ganb_l1_loss(const float* targets, const float* outputs, int64_t count, float scale, int accumulate,
                            float* loss_out, float* doutputs, void* workspace, void* stream) {
  if (!targets || !outputs || !loss_out || !doutputs || !workspace || count <= 0)
    return fail(GANB_E_BADARG, "l1_loss: bad arguments");
  int blocks = egrid(count, 256);
  if (blocks > 1024) blocks = 1024;
  launch_k(l1_loss_partial_kernel, blocks, 256, 0, STREAM, targets, outputs, count, scale / static_cast<float>(count),
           doutputs, static_cast<float*>(workspace));
  GANB_CHECK_LAUNCH("l1_loss_partial_kernel");
  launch_k(l1_loss_finish_kernel, 1, 256, 0, STREAM, static_cast<const float*>(workspace), blocks,
           scale / static_cast<float>(count), accumulate, loss_out);
  GANB_CHECK_LAUNCH("l1_loss_finish_kernel");
  return 0;
}
