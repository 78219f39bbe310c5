// Spectral normalisation (reference common/ops/sn.py:15-69) as grouped, bandwidth-bound warp-shuffle kernels:
// one launch sequence processes every spectrally-normalised weight of a network; each weight is split into
// CTAs of 64 rows so that all SMs stream W (a single CTA per weight is DRAM-latency bound).
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
constexpr int SN_THREADS = 256;
constexpr int SN_ROWS = GANB_SN_ROWS;   // rows of W per CTA (2 per warp: the per-row dot -> axpy chain is latency-bound)

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

__device__ __forceinline__ int find_layer(const ganb_sn_layer* __restrict__ layers, int count, int blk) {
  int l = 0;
  while (l + 1 < count && blk >= layers[l + 1].blk_begin) ++l;
  return l;
}

__device__ __forceinline__ float row_dot(const float* __restrict__ row, const float* __restrict__ xs, int C, int lane) {
  float acc = 0.f;
  if ((C & 3) == 0) {
    for (int c = lane * 4; c < C; c += 128) {
      const float4 w = *reinterpret_cast<const float4*>(row + c);
      acc += w.x * xs[c] + w.y * xs[c + 1] + w.z * xs[c + 2] + w.w * xs[c + 3];
    }
  } else {
    for (int c = lane; c < C; c += 32) acc += row[c] * xs[c];
  }
  return warp_sum(acc);
}

// Forward, stage 1 (one pass over W, rows independent): a_k = row_k . u ; bt += a_k * row_k ; sa += a_k^2.
// Since b = W^T v = W^T (a / (|a|+eps)), the un-normalised bt = W^T a is all that the second stage needs.
// work layout per CTA of the layer: [C] partial bt, then 1 float partial sa (stride C + 4).
__global__ void __launch_bounds__(SN_THREADS) sn_fwd_rows_kernel(const ganb_sn_layer* __restrict__ layers, int count) {
  pdl_wait();
  const int li = find_layer(layers, count, blockIdx.x);
  const ganb_sn_layer L = layers[li];
  const int K = L.k, C = L.c;
  const int blk = blockIdx.x - L.blk_begin;
  extern __shared__ float sm[];
  float* u_s = sm;                 // [C]
  float* bt_s = u_s + C;           // [8][C] per-warp partials
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) u_s[c] = L.u[c];
  for (int i = threadIdx.x; i < 8 * C; i += blockDim.x) bt_s[i] = 0.f;
  __syncthreads();
  float* bt = bt_s + warp * C;
  float sa = 0.f;
  const int r0 = blk * SN_ROWS, r1 = min(K, r0 + SN_ROWS);
  for (int k = r0 + warp; k < r1; k += 8) {
    const float* row = L.w + static_cast<int64_t>(k) * C;
    const float a = row_dot(row, u_s, C, lane);
    if (lane == 0) L.v[k] = a;  // normalised in place by the finish kernel
    sa += a * a;
    if ((C & 3) == 0) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 w = *reinterpret_cast<const float4*>(row + c);
        bt[c] += a * w.x; bt[c + 1] += a * w.y; bt[c + 2] += a * w.z; bt[c + 3] += a * w.w;
      }
    } else {
      for (int c = lane; c < C; c += 32) bt[c] += a * row[c];
    }
  }
  __syncthreads();
  float* out = L.work + static_cast<int64_t>(blk) * (C + 4);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += bt_s[w8 * C + c];
    out[c] = t;
  }
  // sa is identical across the lanes of a warp; count it once per warp
  const float tot = block_sum(lane == 0 ? sa : 0.f, red);
  if (threadIdx.x == 0) out[C] = tot;
}

// Forward, stage 2 (one CTA per weight): reduce the partials, normalise, emit u', sigma.
constexpr int SN_FINISH_THREADS = 1024;
__global__ void __launch_bounds__(SN_FINISH_THREADS) sn_fwd_finish_kernel(const ganb_sn_layer* __restrict__ layers, int assign) {
  pdl_wait();
  const ganb_sn_layer L = layers[blockIdx.x];
  const int K = L.k, C = L.c;
  const int nblk = (K + SN_ROWS - 1) / SN_ROWS;
  extern __shared__ float sm[];
  float* b_s = sm;          // [C]
  float* part_s = sm + C;   // [slices][C]: the per-CTA partials are summed by `slices` thread groups side by side
  __shared__ float red[32];
  float sa = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) sa += L.work[static_cast<int64_t>(i) * (C + 4) + C];
  const float na = sqrtf(block_sum(sa, red));
  const float inv_a = 1.f / (na + SN_EPS);
  const int cols = min(C, static_cast<int>(blockDim.x));
  const int slices = blockDim.x / cols;
  const int cg = threadIdx.x % cols, sl = threadIdx.x / cols;
  if (sl < slices) {
    for (int c = cg; c < C; c += cols) {
      float t = 0.f;
#pragma unroll 4
      for (int i = sl; i < nblk; i += slices) t += __ldg(L.work + static_cast<int64_t>(i) * (C + 4) + c);
      part_s[sl * C + c] = t;
    }
  }
  __syncthreads();
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int g = 0; g < slices; ++g) t += part_s[g * C + c];   // fixed order: deterministic
    t *= inv_a;  // b = W^T v
    b_s[c] = t;
    part += t * t;
  }
  const float nb2 = block_sum(part, red);
  const float nb = sqrtf(nb2);
  const float inv_b = 1.f / (nb + SN_EPS);
  for (int k = threadIdx.x; k < K; k += blockDim.x) L.v[k] *= inv_a;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float un = b_s[c] * inv_b;
    const float uo = L.u[c];
    L.b[c] = b_s[c];
    L.u_out[c] = un;
    L.u_used[c] = uo;
    if (assign) L.u[c] = un;
  }
  if (threadIdx.x == 0) {
    const float sigma = nb2 * inv_b;
    L.scal[0] = sigma;
    L.scal[1] = 1.f / sigma;
    L.scal[2] = na;
    L.scal[3] = nb;
  }
}

// Backward, stage 1 (rows independent): partial <G,W>, t_k = row_k . b, partial sum v_k t_k.
__global__ void __launch_bounds__(SN_THREADS) sn_bwd_rows_kernel(const ganb_sn_layer* __restrict__ layers, int count) {
  pdl_wait();
  const int li = find_layer(layers, count, blockIdx.x);
  const ganb_sn_layer L = layers[li];
  const int K = L.k, C = L.c;
  const int blk = blockIdx.x - L.blk_begin;
  extern __shared__ float sm[];
  float* b_s = sm;  // [C]
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) b_s[c] = L.b[c];
  __syncthreads();
  float gw = 0.f, vt = 0.f;
  const int r0 = blk * SN_ROWS, r1 = min(K, r0 + SN_ROWS);
  for (int k = r0 + warp; k < r1; k += 8) {
    const float* row = L.w + static_cast<int64_t>(k) * C;
    const float* grow = L.g + static_cast<int64_t>(k) * C;
    const float t = row_dot(row, b_s, C, lane);
    float acc = 0.f;
    if ((C & 3) == 0) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 w = *reinterpret_cast<const float4*>(row + c);
        const float4 g = *reinterpret_cast<const float4*>(grow + c);
        acc += w.x * g.x + w.y * g.y + w.z * g.z + w.w * g.w;
      }
    } else {
      for (int c = lane; c < C; c += 32) acc += row[c] * grow[c];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      L.t[k] = t;
      gw += acc;
      vt += L.v[k] * t;
    }
  }
  const float gw_tot = block_sum(gw, red);
  const float vt_tot = block_sum(vt, red);
  if (threadIdx.x == 0) {
    float* out = L.work + static_cast<int64_t>(blk) * (C + 4);
    out[C + 1] = gw_tot;
    out[C + 2] = vt_tot;
  }
}

// Backward, stage 2 (one small CTA per weight): scal[4] = coef (bbar = coef*b), scal[5] = sum v_k t_k.
__global__ void __launch_bounds__(SN_THREADS) sn_bwd_finish_kernel(const ganb_sn_layer* __restrict__ layers) {
  pdl_wait();
  const ganb_sn_layer L = layers[blockIdx.x];
  const int K = L.k, C = L.c;
  const int nblk = (K + SN_ROWS - 1) / SN_ROWS;
  __shared__ float red[32];
  float gw = 0.f, vt = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    gw += L.work[static_cast<int64_t>(i) * (C + 4) + C + 1];
    vt += L.work[static_cast<int64_t>(i) * (C + 4) + C + 2];
  }
  gw = block_sum(gw, red);
  vt = block_sum(vt, red);
  if (threadIdx.x == 0) {
    const float inv_sigma = L.scal[1], nb = L.scal[3];
    const float gs = -gw * inv_sigma * inv_sigma;  // dL/dsigma
    const float den = nb + SN_EPS;
    const float dsig_dnb = (nb * nb + 2.f * nb * SN_EPS) / (den * den);
    L.scal[4] = gs * dsig_dnb / nb;
    L.scal[5] = vt;
  }
}

// Backward, stage 3 (rows independent): dW += G/sigma + v (x) bbar + abar (x) u_used.
__global__ void __launch_bounds__(SN_THREADS) sn_bwd_apply_kernel(const ganb_sn_layer* __restrict__ layers, int count) {
  pdl_wait();
  const int li = find_layer(layers, count, blockIdx.x);
  const ganb_sn_layer L = layers[li];
  const int K = L.k, C = L.c;
  const int blk = blockIdx.x - L.blk_begin;
  extern __shared__ float sm[];
  float* bb_s = sm;        // [C] bbar
  float* u_s = bb_s + C;   // [C]
  const float inv_sigma = L.scal[1], na = L.scal[2], coef = L.scal[4], vt = L.scal[5];
  const float inv_a = 1.f / (na + SN_EPS);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    bb_s[c] = coef * L.b[c];
    u_s[c] = L.u_used[c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blk * SN_ROWS, r1 = min(K, r0 + SN_ROWS);
  for (int k = r0 + warp; k < r1; k += 8) {
    const float vk = L.v[k];
    // vbar_k = coef * t_k ; abar_k = vbar_k/(|a|+eps) - v_k (v.vbar)/|a|
    const float ak = coef * (L.t[k] * inv_a - vk * vt / na);
    const float* grow = L.g + static_cast<int64_t>(k) * C;
    float* drow = L.dw + static_cast<int64_t>(k) * C;
    if ((C & 3) == 0) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 g = *reinterpret_cast<const float4*>(grow + c);
        float4 d = *reinterpret_cast<float4*>(drow + c);
        d.x += g.x * inv_sigma + vk * bb_s[c] + ak * u_s[c];
        d.y += g.y * inv_sigma + vk * bb_s[c + 1] + ak * u_s[c + 1];
        d.z += g.z * inv_sigma + vk * bb_s[c + 2] + ak * u_s[c + 2];
        d.w += g.w * inv_sigma + vk * bb_s[c + 3] + ak * u_s[c + 3];
        *reinterpret_cast<float4*>(drow + c) = d;
      }
    } else {
      for (int c = lane; c < C; c += 32) drow[c] += grow[c] * inv_sigma + vk * bb_s[c] + ak * u_s[c];
    }
  }
}

// ------------------------------------------------------------------------------------------------ pack
// One block per 64 x 64 (ci x co) tile of one tap: fp32 rows in (256 bytes per warp), bf16 rows out -- 128 bytes per warp for
// both copies (the 32 x 32 tiles of the first version wrote 64-byte rows: 28 us for the generator's 6 M parameters, a serial
// tail of every step).
constexpr int PACK_TILE = 64;
__global__ void __launch_bounds__(256) pack_weights_kernel(const ganb_pack_layer* __restrict__ layers, int nlayers) {
  pdl_wait();
  __shared__ float tile[PACK_TILE][PACK_TILE + 1];
  int l = 0;
  while (l + 1 < nlayers && static_cast<int>(blockIdx.x) >= layers[l + 1].tile_begin) ++l;
  const ganb_pack_layer L = layers[l];
  int local = blockIdx.x - L.tile_begin;
  const int tiles_co = (L.co + PACK_TILE - 1) / PACK_TILE, tiles_ci = (L.ci + PACK_TILE - 1) / PACK_TILE;
  const int tco = local % tiles_co; local /= tiles_co;
  const int tci = local % tiles_ci; local /= tiles_ci;
  const int tap = local;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8: a thread owns two adjacent columns
  const int cip = L.ci_pad > L.ci ? L.ci_pad : L.ci;   // channel count of the operand copies (zero rows beyond ci)
  const float* w = L.w + static_cast<int64_t>(tap) * L.ci * L.co;
  __nv_bfloat16* wn = L.wn ? static_cast<__nv_bfloat16*>(L.wn) + static_cast<int64_t>(tap) * cip * L.co : nullptr;
  __nv_bfloat16* wt = L.wt ? static_cast<__nv_bfloat16*>(L.wt) + static_cast<int64_t>(tap) * cip * L.co : nullptr;
  const bool co_even = (L.co & 1) == 0, ci_even = (cip & 1) == 0;
  const bool w_al8 = (reinterpret_cast<uintptr_t>(L.w) & 7) == 0;
  for (int j = ty; j < PACK_TILE; j += 8) {
    const int ci = tci * PACK_TILE + j, co = tco * PACK_TILE + 2 * tx;
    float v0 = 0.f, v1 = 0.f;
    if (ci < L.ci) {
      const float* src = w + static_cast<int64_t>(ci) * L.co + co;
      if (co_even && co + 1 < L.co) {
        if (w_al8) {      // (a filter is a slice of its network's flat parameter buffer: 4-byte alignment only is guaranteed)
          const float2 v = *reinterpret_cast<const float2*>(src);
          v0 = v.x; v1 = v.y;
        } else {
          v0 = src[0]; v1 = src[1];
        }
        if (wn) *reinterpret_cast<__nv_bfloat162*>(wn + static_cast<int64_t>(ci) * L.co + co) = __floats2bfloat162_rn(v0, v1);
      } else {
        if (co < L.co) {
          v0 = src[0];
          if (wn) wn[static_cast<int64_t>(ci) * L.co + co] = __float2bfloat16_rn(v0);
        }
        if (co + 1 < L.co) {
          v1 = src[1];
          if (wn) wn[static_cast<int64_t>(ci) * L.co + co + 1] = __float2bfloat16_rn(v1);
        }
      }
    }
    tile[j][2 * tx] = v0;
    tile[j][2 * tx + 1] = v1;
  }
  __syncthreads();
  if (wt) {
    for (int j = ty; j < PACK_TILE; j += 8) {
      const int co = tco * PACK_TILE + j, ci = tci * PACK_TILE + 2 * tx;
      if (co >= L.co) continue;
      __nv_bfloat16* dst = wt + static_cast<int64_t>(co) * cip + ci;
      const float v0 = tile[2 * tx][j], v1 = tile[2 * tx + 1][j];
      if (ci_even && ci + 1 < L.ci) {
        *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(v0, v1);
      } else {
        if (ci < L.ci) dst[0] = __float2bfloat16_rn(v0);
        if (ci + 1 < L.ci) dst[1] = __float2bfloat16_rn(v1);
      }
    }
  }
}

}  // namespace ganb

using namespace ganb;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ganb_sn_power_iter(const ganb_sn_layer* layers_dev, int count, int total_blocks, int max_c, int assign,
                                  void* stream) {
  if (!layers_dev || count <= 0 || total_blocks <= 0) return fail(GANB_E_BADARG, "sn_power_iter: no layers");
  const int smem1 = 9 * max_c * 4;
  if (smem1 > 200 * 1024) return fail(GANB_E_UNSUPPORTED, "sn_power_iter: c=%d too large", max_c);
  if (smem1 > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sn_fwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "sn_power_iter: %s", cudaGetErrorString(e));
  }
  launch_k(sn_fwd_rows_kernel, total_blocks, SN_THREADS, smem1, STREAM, layers_dev, count);
  GANB_CHECK_LAUNCH("sn_fwd_rows_kernel");
  // b_s [C] + part_s [slices][C], slices * C <= max(C, threads)
  const int smem2 = (max_c + (max_c > SN_FINISH_THREADS ? max_c : SN_FINISH_THREADS)) * 4;
  launch_k(sn_fwd_finish_kernel, count, SN_FINISH_THREADS, smem2, STREAM, layers_dev, assign);
  GANB_CHECK_LAUNCH("sn_fwd_finish_kernel");
  return 0;
}

extern "C" int ganb_sn_bwd(const ganb_sn_layer* layers_dev, int count, int total_blocks, int max_c, void* stream) {
  if (!layers_dev || count <= 0 || total_blocks <= 0) return fail(GANB_E_BADARG, "sn_bwd: no layers");
  launch_k(sn_bwd_rows_kernel, total_blocks, SN_THREADS, max_c * 4, STREAM, layers_dev, count);
  GANB_CHECK_LAUNCH("sn_bwd_rows_kernel");
  launch_k(sn_bwd_finish_kernel, count, SN_THREADS, 0, STREAM, layers_dev);
  GANB_CHECK_LAUNCH("sn_bwd_finish_kernel");
  launch_k(sn_bwd_apply_kernel, total_blocks, SN_THREADS, 2 * max_c * 4, STREAM, layers_dev, count);
  GANB_CHECK_LAUNCH("sn_bwd_apply_kernel");
  return 0;
}

extern "C" int ganb_pack_weights(const ganb_pack_layer* layers_dev, int count, int total_tiles, void* stream) {
  if (!layers_dev || count <= 0 || total_tiles <= 0) return fail(GANB_E_BADARG, "pack_weights: no layers");
  launch_k(pack_weights_kernel, total_tiles, 256, 0, STREAM, layers_dev, count);
  GANB_CHECK_LAUNCH("pack_weights_kernel");
  return 0;
}
