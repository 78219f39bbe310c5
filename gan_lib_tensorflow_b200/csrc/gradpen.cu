// Kernels of the WGAN-GP gradient penalty (ACGAN/train.py:97-105, also Pix2Pix/train.py:489-507):
//   x_hat = real + alpha * (fake - real);  g = d sum(D(x_hat)) / d x_hat;
//   penalty = lambda * mean_n (sqrt(sum_hwc g^2 + 1e-10) - 1)^2
// The penalty is differentiated w.r.t. D's parameters, i.e. THROUGH the backward pass that produced g.  Every step of
// that backward pass is a differentiable op on the host side (functional.py, "second order"); the only step whose
// vector-Jacobian product is not one of the existing kernels is the backward of a training-mode batch norm:
//   gx = gamma*r * (gy' - mean(gy') - xh * mean(gy' xh)),   gy' = gy * act'(gamma*xh + beta),  xh = (x - mu) * r
// whose VJP for a cotangent c of gx is (per channel, M = pixels of the batch, bars = means over the batch):
//   d/dgy  = act' * gamma*r * (c - cbar - xh * tbar)                         t = c * xh      (self-adjoint form)
//   d/dx   = gamma*r^2 * ( (3 s t - mean(c gy') + gbar cbar) * xh - tbar' ... )  -- see bn_vjp_apply_kernel
//   d/dgamma = r * (sum(c gy') - M gbar cbar - M sbar tbar)                  s = gy' * xh
// fp32 throughout; the tensors involved are the small activations of a CIFAR-size critic (not a hot path).
#include "host_common.h"

#include <cuda_bf16.h>

namespace ganb {

__device__ __forceinline__ float gp_dact(float y, int act) {
  if (act == GANB_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == GANB_ACT_LRELU) return y >= 0.f ? 1.f : 0.2f;
  return 1.f;
}

// out[n, :] = real[n, :] + alpha[n] * (fake[n, :] - real[n, :])
__global__ void __launch_bounds__(256)
interpolate_kernel(const float* __restrict__ real, const float* __restrict__ fake, const float* __restrict__ alpha,
                   int64_t per_sample, int64_t total, float* __restrict__ out) {
  pdl_wait();
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const float a = alpha[i / per_sample];
    const float r = real[i];
    out[i] = r + a * (fake[i] - r);
  }
}

// one block per sample: slope = sqrt(sum g^2 + 1e-10); term[n] = (slope - 1)^2; dg = scale * 2 (slope - 1) / (n * slope) * g
__global__ void __launch_bounds__(256)
gp_loss_sample_kernel(const float* __restrict__ g, int n, int64_t m, float scale, float* __restrict__ dg,
                      float* __restrict__ term) {
  pdl_wait();
  const int ni = blockIdx.x;
  const float* gs = g + static_cast<int64_t>(ni) * m;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < m; i += 256) { const float v = gs[i]; acc += v * v; }
  __shared__ float sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const float slope = sqrtf(sh[0] + 1e-10f);
  const float coef = scale * 2.f * (slope - 1.f) / (static_cast<float>(n) * slope);
  float* ds = dg + static_cast<int64_t>(ni) * m;
  for (int64_t i = threadIdx.x; i < m; i += 256) ds[i] = coef * gs[i];
  if (threadIdx.x == 0) term[ni] = (slope - 1.f) * (slope - 1.f);
}
__global__ void __launch_bounds__(256)
gp_loss_finish_kernel(const float* __restrict__ term, int n, float scale, int accumulate, float* __restrict__ loss_out) {
  pdl_wait();
  __shared__ float sh[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += term[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (accumulate ? loss_out[0] : 0.f) + scale * sh[0] / static_cast<float>(n);
}

// ---- batch-norm backward VJP.  partial[chunk][5][C]: sum c, sum c*xh, sum c*gy', sum gy', sum gy'*xh
constexpr int BNV_LANES = 8;   // pixel lanes per block (x 32 channels)
__global__ void __launch_bounds__(256)
bn_vjp_reduce_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ cot,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int64_t pixels, int C, int act, int pix_per_chunk,
                     float* __restrict__ partial) {
  pdl_wait();
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cx;
  const int chunk = blockIdx.y;
  const int64_t p0 = static_cast<int64_t>(chunk) * pix_per_chunk;
  const int64_t p1 = min(pixels, p0 + pix_per_chunk);
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (ch < C) {
    const float mu = mean[ch], r = rstd[ch];
    const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
    for (int64_t p = p0 + py; p < p1; p += BNV_LANES) {
      const int64_t o = p * C + ch;
      const float xh = (x[o] - mu) * r;
      const float gp = gy[o] * gp_dact(ga * xh + be, act);
      const float c = cot[o];
      s[0] += c; s[1] += c * xh; s[2] += c * gp; s[3] += gp; s[4] += gp * xh;
    }
  }
  __shared__ float sh[BNV_LANES][5][32];
#pragma unroll
  for (int k = 0; k < 5; ++k) sh[py][k][cx] = s[k];
  __syncthreads();
  if (py == 0 && ch < C) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float t = 0.f;
      for (int l = 0; l < BNV_LANES; ++l) t += sh[l][k][cx];
      partial[(static_cast<int64_t>(chunk) * 5 + k) * C + ch] = t;
    }
  }
}
// sums[5][C] = sum over chunks (fixed order); dgamma[ch] += r * (S_cg - M gbar cbar - M sbar tbar)
__global__ void __launch_bounds__(256)
bn_vjp_finalize_kernel(const float* __restrict__ partial, int chunks, int C, float inv_m, const float* __restrict__ rstd,
                       float* __restrict__ sums, float* __restrict__ dgamma) {
  pdl_wait();
  const int ch = blockIdx.x * 256 + threadIdx.x;
  if (ch >= C) return;
  float s[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float t = 0.f;
    for (int q = 0; q < chunks; ++q) t += partial[(static_cast<int64_t>(q) * 5 + k) * C + ch];
    s[k] = t;
    sums[k * C + ch] = t;
  }
  if (dgamma) {
    const float m = 1.f / inv_m;
    const float cbar = s[0] * inv_m, tbar = s[1] * inv_m, gbar = s[3] * inv_m, sbar = s[4] * inv_m;
    dgamma[ch] += rstd[ch] * (s[2] - m * gbar * cbar - m * sbar * tbar);
  }
}
__global__ void __launch_bounds__(256)
bn_vjp_apply_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ cot,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ sums, int64_t total, int C, float inv_m,
                    int act, float* __restrict__ dx, float* __restrict__ dgy) {
  pdl_wait();
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ch = static_cast<int>(i % C);
    const float mu = mean[ch], r = rstd[ch];
    const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
    const float cbar = sums[ch] * inv_m, tbar = sums[C + ch] * inv_m, cg = sums[2 * C + ch] * inv_m;
    const float gbar = sums[3 * C + ch] * inv_m, sbar = sums[4 * C + ch] * inv_m;
    const float xh = (x[i] - mu) * r;
    const float da = gp_dact(ga * xh + be, act);
    const float gp = gy[i] * da;
    const float c = cot[i];
    // d/dx of sum(c * gx):  gamma r^2 [ (3 sbar tbar - mean(c gy') + gbar cbar) xh - tbar (gy' - gbar) - sbar (c - cbar) ]
    dx[i] = ga * r * r * ((3.f * sbar * tbar - cg + gbar * cbar) * xh - tbar * (gp - gbar) - sbar * (c - cbar));
    dgy[i] = da * ga * r * (c - cbar - xh * tbar);
  }
}

}  // namespace ganb

using namespace ganb;
#define STREAM static_cast<cudaStream_t>(stream)

static int gp_grid(int64_t items) {
  int64_t b = ceil_div64(items, 256);
  const int64_t cap = 8LL * sm_count();
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

extern "C" int ganb_interpolate(const float* real, const float* fake, const float* alpha, int n, int64_t per_sample,
                                float* out, void* stream) {
  if (!real || !fake || !alpha || !out || n <= 0 || per_sample <= 0) return fail(GANB_E_BADARG, "interpolate: bad arguments");
  const int64_t total = static_cast<int64_t>(n) * per_sample;
  launch_k(interpolate_kernel, gp_grid(total), 256, 0, STREAM, real, fake, alpha, per_sample, total, out);
  GANB_CHECK_LAUNCH("interpolate_kernel");
  return 0;
}

extern "C" int ganb_gp_loss(const float* g, int n, int64_t per_sample, float scale, int accumulate, float* loss_out,
                            float* dg, float* workspace_n, void* stream) {
  if (!g || !loss_out || !dg || !workspace_n || n <= 0 || per_sample <= 0) return fail(GANB_E_BADARG, "gp_loss: bad arguments");
  launch_k(gp_loss_sample_kernel, n, 256, 0, STREAM, g, n, per_sample, scale, dg, workspace_n);
  GANB_CHECK_LAUNCH("gp_loss_sample_kernel");
  launch_k(gp_loss_finish_kernel, 1, 256, 0, STREAM, static_cast<const float*>(workspace_n), n, scale, accumulate, loss_out);
  GANB_CHECK_LAUNCH("gp_loss_finish_kernel");
  return 0;
}

static int bn_vjp_chunks(int64_t pixels) {
  int64_t c = ceil_div64(pixels, 512);
  if (c > 64) c = 64;
  return static_cast<int>(c < 1 ? 1 : c);
}

extern "C" int64_t ganb_bn_bwd_vjp_workspace(int64_t pixels, int c) {
  return (static_cast<int64_t>(bn_vjp_chunks(pixels)) * 5 * c + 5LL * c) * 4;
}

extern "C" int ganb_bn_bwd_vjp(const float* x, const float* gy, const float* cot, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, int64_t pixels, int c, int act, float* dx,
                               float* dgy, float* dgamma, void* workspace, void* stream) {
  if (!x || !gy || !cot || !mean || !rstd || !dx || !dgy || !workspace || pixels <= 0 || c <= 0)
    return fail(GANB_E_BADARG, "bn_bwd_vjp: bad arguments");
  const int chunks = bn_vjp_chunks(pixels);
  const int ppc = static_cast<int>(ceil_div64(pixels, chunks));
  float* partial = static_cast<float*>(workspace);
  float* sums = partial + static_cast<int64_t>(chunks) * 5 * c;
  const float inv_m = 1.0f / static_cast<float>(pixels);
  launch_k(bn_vjp_reduce_kernel, dim3(ceil_div(c, 32), chunks), 256, 0, STREAM, x, gy, cot, mean, rstd, gamma, beta, pixels,
           c, act, ppc, partial);
  GANB_CHECK_LAUNCH("bn_vjp_reduce_kernel");
  launch_k(bn_vjp_finalize_kernel, ceil_div(c, 256), 256, 0, STREAM, static_cast<const float*>(partial), chunks, c, inv_m,
           rstd, sums, dgamma);
  GANB_CHECK_LAUNCH("bn_vjp_finalize_kernel");
  launch_k(bn_vjp_apply_kernel, gp_grid(pixels * c), 256, 0, STREAM, x, gy, cot, mean, rstd, gamma, beta,
           static_cast<const float*>(sums), pixels * c, c, inv_m, act, dx, dgy);
  GANB_CHECK_LAUNCH("bn_vjp_apply_kernel");
  return 0;
}
