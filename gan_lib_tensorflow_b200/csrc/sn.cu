// Spectral normalisation (reference common/ops/sn.py:15-69) as grouped, bandwidth-bound warp-shuffle kernels:
// one launch processes every spectrally-normalised weight of a network (one CTA per weight).
//
//   forward : a = W u ; v = a/(|a|+eps) ; b = W^T v ; u' = b/(|b|+eps) ; sigma = b.u' = |b|^2/(|b|+eps)
//             (one power iteration, sn.py:34-47; sigma as at sn.py:52/58).  W/sigma is never materialised:
//             1/sigma is consumed as the `alpha` of the convolution epilogue.
//   backward: the reference has no stop_gradient, so dL/dW carries three terms:
//             G/sigma  +  v (x) bbar  +  abar (x) u
//             with gs = -<G,W>/sigma^2, bbar = gs * dsigma/db, vbar = W bbar, abar = vbar/(|a|+eps) - v (v.vbar)/|a|.
//   pack    : fp32 HWIO filters -> bf16 in the two layouts the tensor-core kernels consume
//             ([tap][ci][co] for dgrad, [tap][co][ci] for fprop).
#include "host_common.h"

#include <cuda_bf16.h>

namespace ganb {

constexpr float SN_EPS = 1e-12f;  // sn.py:11
constexpr int SN_THREADS = 1024;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; every thread receives the result. `red` holds 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  const float r = red[0];
  __syncthreads();
  return r;
}

// out_k = sum_c W[k,c] * x_c for all k: one warp per row, lanes stride the contiguous c dimension.
__device__ __forceinline__ void rows_dot(const float* __restrict__ W, int K, int C, const float* __restrict__ xs,
                                         float* __restrict__ out_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarps) {
    const float* row = W + static_cast<int64_t>(k) * C;
    float acc = 0.f;
    if ((C & 3) == 0) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 w = *reinterpret_cast<const float4*>(row + c);
        acc += w.x * xs[c] + w.y * xs[c + 1] + w.z * xs[c + 2] + w.w * xs[c + 3];
      }
    } else {
      for (int c = lane; c < C; c += 32) acc += row[c] * xs[c];
    }
    acc = warp_sum(acc);
    if (lane == 0) out_s[k] = acc;
  }
}

// out_c = sum_k W[k,c] * y_k for all c: threads own columns (coalesced along c), k split over row lanes.
// `scratch` needs blockDim.x floats; result left in out_s[0..C).
__device__ __forceinline__ void cols_dot(const float* __restrict__ W, int K, int C, const float* __restrict__ ys,
                                         float* __restrict__ out_s, float* __restrict__ scratch) {
  const int cols = min(C, static_cast<int>(blockDim.x));
  const int lanes = blockDim.x / cols;
  const int cx = threadIdx.x % cols, kl = threadIdx.x / cols;
  for (int cb = 0; cb < C; cb += cols) {
    const int c = cb + cx;
    float acc = 0.f;
    if (c < C && kl < lanes)
      for (int k = kl; k < K; k += lanes) acc += W[static_cast<int64_t>(k) * C + c] * ys[k];
    scratch[threadIdx.x] = acc;
    __syncthreads();
    if (kl == 0 && c < C) {
      for (int l = 1; l < lanes; ++l) acc += scratch[l * cols + cx];
      out_s[c] = acc;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SN_THREADS) sn_fwd_kernel(const ganb_sn_layer* __restrict__ layers, int assign) {
  const ganb_sn_layer L = layers[blockIdx.x];
  const int K = L.k, C = L.c;
  extern __shared__ float sm[];
  float* a_s = sm;              // [K]  a, then v
  float* u_s = a_s + K;         // [C]  u, later b
  float* b_s = u_s + C;         // [C]
  float* scratch = b_s + C;     // [blockDim]
  float* red = scratch + blockDim.x;  // [32]

  for (int c = threadIdx.x; c < C; c += blockDim.x) u_s[c] = L.u[c];
  __syncthreads();
  rows_dot(L.w, K, C, u_s, a_s);
  __syncthreads();
  float part = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) part += a_s[k] * a_s[k];
  const float na = sqrtf(block_sum(part, red));
  const float inv_a = 1.f / (na + SN_EPS);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = a_s[k] * inv_a;
    a_s[k] = v;
    L.v[k] = v;
  }
  __syncthreads();
  cols_dot(L.w, K, C, a_s, b_s, scratch);
  part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) part += b_s[c] * b_s[c];
  const float nb2 = block_sum(part, red);
  const float nb = sqrtf(nb2);
  const float inv_b = 1.f / (nb + SN_EPS);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float un = b_s[c] * inv_b;
    L.b[c] = b_s[c];
    L.u_out[c] = un;
    L.u_used[c] = u_s[c];
    if (assign) L.u[c] = un;  // every read of u happened before the first block-wide barrier
  }
  if (threadIdx.x == 0) {
    const float sigma = nb2 * inv_b;
    L.scal[0] = sigma;
    L.scal[1] = 1.f / sigma;
    L.scal[2] = na;
    L.scal[3] = nb;
  }
}

__global__ void __launch_bounds__(SN_THREADS) sn_bwd_kernel(const ganb_sn_layer* __restrict__ layers) {
  const ganb_sn_layer L = layers[blockIdx.x];
  const int K = L.k, C = L.c;
  const int64_t total = static_cast<int64_t>(K) * C;
  extern __shared__ float sm[];
  float* v_s = sm;              // [K]
  float* ab_s = v_s + K;        // [K]  vbar then abar
  float* bb_s = ab_s + K;       // [C]  bbar
  float* u_s = bb_s + C;        // [C]
  float* scratch = u_s + C;
  float* red = scratch + blockDim.x;

  const float sigma = L.scal[0], inv_sigma = L.scal[1], na = L.scal[2], nb = L.scal[3];
  for (int k = threadIdx.x; k < K; k += blockDim.x) v_s[k] = L.v[k];
  for (int c = threadIdx.x; c < C; c += blockDim.x) u_s[c] = L.u_used[c];

  // <G, W>
  float part = 0.f;
  if ((total & 3) == 0) {
    const float4* g4 = reinterpret_cast<const float4*>(L.g);
    const float4* w4 = reinterpret_cast<const float4*>(L.w);
    for (int64_t i = threadIdx.x; i < (total >> 2); i += blockDim.x) {
      const float4 g = g4[i], w = w4[i];
      part += g.x * w.x + g.y * w.y + g.z * w.z + g.w * w.w;
    }
  } else {
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) part += L.g[i] * L.w[i];
  }
  const float gw = block_sum(part, red);
  const float gs = -gw * inv_sigma * inv_sigma;                   // dL/dsigma
  const float den = nb + SN_EPS;
  const float dsig_dnb = (nb * nb + 2.f * nb * SN_EPS) / (den * den);
  const float coef = gs * dsig_dnb / nb;                          // bbar = coef * b
  for (int c = threadIdx.x; c < C; c += blockDim.x) bb_s[c] = coef * L.b[c];
  __syncthreads();
  rows_dot(L.w, K, C, bb_s, ab_s);                                // vbar = W bbar
  __syncthreads();
  part = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) part += v_s[k] * ab_s[k];
  const float vv = block_sum(part, red);
  const float inv_a = 1.f / (na + SN_EPS);
  for (int k = threadIdx.x; k < K; k += blockDim.x) ab_s[k] = ab_s[k] * inv_a - v_s[k] * vv / na;
  __syncthreads();
  (void)sigma;
  // dW += G/sigma + v (x) bbar + abar (x) u
  if ((C & 3) == 0) {
    const int c4n = C >> 2;
    for (int64_t i = threadIdx.x; i < (total >> 2); i += blockDim.x) {
      const int k = static_cast<int>(i / c4n), c = static_cast<int>(i % c4n) * 4;
      const float4 g = reinterpret_cast<const float4*>(L.g)[i];
      float4 d = reinterpret_cast<float4*>(L.dw)[i];
      const float vk = v_s[k], ak = ab_s[k];
      d.x += g.x * inv_sigma + vk * bb_s[c] + ak * u_s[c];
      d.y += g.y * inv_sigma + vk * bb_s[c + 1] + ak * u_s[c + 1];
      d.z += g.z * inv_sigma + vk * bb_s[c + 2] + ak * u_s[c + 2];
      d.w += g.w * inv_sigma + vk * bb_s[c + 3] + ak * u_s[c + 3];
      reinterpret_cast<float4*>(L.dw)[i] = d;
    }
  } else {
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
      const int k = static_cast<int>(i / C), c = static_cast<int>(i % C);
      L.dw[i] += L.g[i] * inv_sigma + v_s[k] * bb_s[c] + ab_s[k] * u_s[c];
    }
  }
}

// ------------------------------------------------------------------------------------------------ pack
__global__ void __launch_bounds__(256) pack_weights_kernel(const ganb_pack_layer* __restrict__ layers, int nlayers) {
  __shared__ float tile[32][33];
  int l = 0;
  while (l + 1 < nlayers && static_cast<int>(blockIdx.x) >= layers[l + 1].tile_begin) ++l;
  const ganb_pack_layer L = layers[l];
  int local = blockIdx.x - L.tile_begin;
  const int tiles_co = (L.co + 31) / 32, tiles_ci = (L.ci + 31) / 32;
  const int tco = local % tiles_co; local /= tiles_co;
  const int tci = local % tiles_ci; local /= tiles_ci;
  const int tap = local;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* w = L.w + static_cast<int64_t>(tap) * L.ci * L.co;
  __nv_bfloat16* wn = L.wn ? static_cast<__nv_bfloat16*>(L.wn) + static_cast<int64_t>(tap) * L.ci * L.co : nullptr;
  __nv_bfloat16* wt = L.wt ? static_cast<__nv_bfloat16*>(L.wt) + static_cast<int64_t>(tap) * L.ci * L.co : nullptr;
  for (int j = ty; j < 32; j += 8) {
    const int ci = tci * 32 + j, co = tco * 32 + tx;
    float v = 0.f;
    if (ci < L.ci && co < L.co) {
      v = w[static_cast<int64_t>(ci) * L.co + co];
      if (wn) wn[static_cast<int64_t>(ci) * L.co + co] = __float2bfloat16_rn(v);
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  if (wt) {
    for (int j = ty; j < 32; j += 8) {
      const int co = tco * 32 + j, ci = tci * 32 + tx;
      if (ci < L.ci && co < L.co) wt[static_cast<int64_t>(co) * L.ci + ci] = __float2bfloat16_rn(tile[tx][j]);
    }
  }
}

}  // namespace ganb

using namespace ganb;
#define STREAM static_cast<cudaStream_t>(stream)

static int sn_smem_bytes(int max_k, int max_c, bool bwd) {
  const int floats = bwd ? (2 * max_k + 2 * max_c) : (max_k + 2 * max_c);
  return (floats + SN_THREADS + 32) * 4;
}

extern "C" int ganb_sn_power_iter(const ganb_sn_layer* layers_dev, int count, int max_k, int max_c, int assign,
                                  void* stream) {
  if (!layers_dev || count <= 0) return fail(GANB_E_BADARG, "sn_power_iter: no layers");
  const int smem = sn_smem_bytes(max_k, max_c, false);
  if (smem > 200 * 1024) return fail(GANB_E_UNSUPPORTED, "sn_power_iter: K=%d too large for one CTA", max_k);
  cudaError_t e = cudaFuncSetAttribute(sn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "sn_power_iter: %s", cudaGetErrorString(e));
  sn_fwd_kernel<<<count, SN_THREADS, smem, STREAM>>>(layers_dev, assign);
  GANB_CHECK_LAUNCH("sn_fwd_kernel");
  return 0;
}

extern "C" int ganb_sn_bwd(const ganb_sn_layer* layers_dev, int count, int max_k, int max_c, void* stream) {
  if (!layers_dev || count <= 0) return fail(GANB_E_BADARG, "sn_bwd: no layers");
  const int smem = sn_smem_bytes(max_k, max_c, true);
  if (smem > 200 * 1024) return fail(GANB_E_UNSUPPORTED, "sn_bwd: K=%d too large for one CTA", max_k);
  cudaError_t e = cudaFuncSetAttribute(sn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "sn_bwd: %s", cudaGetErrorString(e));
  sn_bwd_kernel<<<count, SN_THREADS, smem, STREAM>>>(layers_dev);
  GANB_CHECK_LAUNCH("sn_bwd_kernel");
  return 0;
}

extern "C" int ganb_pack_weights(const ganb_pack_layer* layers_dev, int count, int total_tiles, void* stream) {
  if (!layers_dev || count <= 0 || total_tiles <= 0) return fail(GANB_E_BADARG, "pack_weights: no layers");
  pack_weights_kernel<<<total_tiles, 256, 0, STREAM>>>(layers_dev, count);
  GANB_CHECK_LAUNCH("pack_weights_kernel");
  return 0;
}
