// CUDA-core kernels for the tensor-core-hostile corners of the path (SURVEY.md 7 "hard parts" #5): convolutions
// whose input or output has <= 8 channels (RGB image side of D.Block.1.*, G.Output) and tiny dense layers
// (D.Embedding_y 300->128, D.Output 128->1).  They are bandwidth-bound on their large-channel side.
//
// Replaces the same TF call-sites as conv_tc.cu (tf.nn.conv2d and gradients, tf.matmul) for those shapes.
#include "host_common.h"

#include <cuda_bf16.h>

namespace ganb {

struct alignas(8) sc_bf16x4 {
  __nv_bfloat162 lo, hi;
};
__device__ __forceinline__ float4 sc_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 sc_ld4(const __nv_bfloat16* p) {
  const sc_bf16x4 v = *reinterpret_cast<const sc_bf16x4*>(p);
  const float2 a = __bfloat1622float2(v.lo), b = __bfloat1622float2(v.hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void sc_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void sc_st4(__nv_bfloat16* p, float4 v) {
  sc_bf16x4 o;
  o.lo = __floats2bfloat162_rn(v.x, v.y);
  o.hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<sc_bf16x4*>(p) = o;
}
__device__ __forceinline__ float sc_act(float v, int act) {
  if (act == GANB_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == GANB_ACT_LRELU) return v >= 0.f ? v : 0.2f * v;
  if (act == GANB_ACT_TANH) return tanhf(v);
  return v;
}

// ------------------------------------------------------------------------------------------------
// y[n,ho,wo,cl] = act(alpha * sum_{r,s,cs} x[n, ho+r-pad_t, wo+s-pad_l, cs] * w(tap, cs, cl) + bias[cl])
// x has CS <= 8 channels (fp32); cl % 4 == 0.  w layout: [tap][cs][cl] or, when w_clcs, [tap][cl][cs];
// flip uses tap' = taps-1-tap (data gradient of a stride-1 convolution).
struct SmallCinParams {
  const float* x; const float* w; void* y;
  int n, h, w_in, cs, ho, wo, cl, kh, kw, pad_t, pad_l;
  int flip, w_clcs;
  const float* alpha; const float* bias;
  int act, out_bf16;
};

__global__ void __launch_bounds__(256) conv_smallcin_kernel(const SmallCinParams p) {
  pdl_wait();
  extern __shared__ float wsm[];  // [taps][cs][cl]
  const int taps = p.kh * p.kw;
  const int wcount = taps * p.cs * p.cl;
  for (int i = threadIdx.x; i < wcount; i += blockDim.x) {
    const int cl = i % p.cl, cs = (i / p.cl) % p.cs, tap = i / (p.cl * p.cs);
    const int tsrc = p.flip ? taps - 1 - tap : tap;
    wsm[i] = p.w_clcs ? p.w[(static_cast<int64_t>(tsrc) * p.cl + cl) * p.cs + cs]
                      : p.w[(static_cast<int64_t>(tsrc) * p.cs + cs) * p.cl + cl];
  }
  __syncthreads();
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.f;
  const int v = p.cl >> 2;
  const int64_t total = static_cast<int64_t>(p.n) * p.ho * p.wo * v;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const int wo = static_cast<int>(pix % p.wo);
    const int ho = static_cast<int>((pix / p.wo) % p.ho);
    const int ni = static_cast<int>(pix / (static_cast<int64_t>(p.wo) * p.ho));
    float4 acc = make_float4(0, 0, 0, 0);
    for (int r = 0; r < p.kh; ++r) {
      const int hi = ho + r - p.pad_t;
      if (hi < 0 || hi >= p.h) continue;
      for (int s = 0; s < p.kw; ++s) {
        const int wi = wo + s - p.pad_l;
        if (wi < 0 || wi >= p.w_in) continue;
        const float* xp = p.x + ((static_cast<int64_t>(ni) * p.h + hi) * p.w_in + wi) * p.cs;
        const float* wp = wsm + (r * p.kw + s) * p.cs * p.cl + c4;
        for (int cs = 0; cs < p.cs; ++cs) {
          const float xv = __ldg(xp + cs);
          const float4 w4 = *reinterpret_cast<const float4*>(wp + cs * p.cl);
          acc.x += xv * w4.x; acc.y += xv * w4.y; acc.z += xv * w4.z; acc.w += xv * w4.w;
        }
      }
    }
    float4 o = make_float4(acc.x * alpha, acc.y * alpha, acc.z * alpha, acc.w * alpha);
    if (p.bias) {
      const float4 b = sc_ld4(p.bias + c4);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
    o = make_float4(sc_act(o.x, p.act), sc_act(o.y, p.act), sc_act(o.z, p.act), sc_act(o.w, p.act));
    if (p.out_bf16) sc_st4(reinterpret_cast<__nv_bfloat16*>(p.y) + pix * p.cl + c4, o);
    else sc_st4(reinterpret_cast<float*>(p.y) + pix * p.cl + c4, o);
  }
}

// ------------------------------------------------------------------------------------------------
// partial[chunk][tap][cs][cl] = sum over the chunk's pixels q of  xs[q + sign*(tap offset)][cs] * yl[q][cl]
// xs: small-channel fp32 tensor [n,hs,ws,cs]; yl: large-channel tensor [n,hl,wl,cl] (fp32 or bf16).
// sign=+1: xs is the convolution input, yl the output gradient  (D.Block.1.Conv1 wgrad)
// sign=-1: yl is the convolution input, xs the output gradient  (G.Output wgrad)
constexpr int SW_MAX_ACC = 27;  // taps*cs accumulators (float4 each) kept in registers

struct SmallWgradParams {
  const float* xs; const void* yl; int yl_bf16;
  int n, hs, ws, cs, hl, wl, cl, kh, kw, pad_t, pad_l, sign;
  int pix_per_chunk;
  float* partial;
};

template <typename TL>
__global__ void __launch_bounds__(256) conv_small_wgrad_kernel(const SmallWgradParams p) {
  pdl_wait();
  const int v = p.cl >> 2;
  const int cols = min(v, 256);
  const int lanes = 256 / cols;
  const int cx = threadIdx.x % cols, ly = threadIdx.x / cols;
  const int taps = p.kh * p.kw;
  const int nacc = taps * p.cs;
  const int64_t total_pix = static_cast<int64_t>(p.n) * p.hl * p.wl;
  const int64_t q0 = static_cast<int64_t>(blockIdx.x) * p.pix_per_chunk;
  const int64_t q1 = min(total_pix, q0 + p.pix_per_chunk);
  const TL* yl = static_cast<const TL*>(p.yl);
  extern __shared__ float4 red[];  // [lanes][cols] reused per accumulator
  for (int cb = 0; cb < v; cb += cols) {
    const int col = cb + cx;
    float4 acc[SW_MAX_ACC];
#pragma unroll
    for (int a = 0; a < SW_MAX_ACC; ++a) acc[a] = make_float4(0, 0, 0, 0);
    if (col < v && ly < lanes) {
      for (int64_t q = q0 + ly; q < q1; q += lanes) {
        const int wq = static_cast<int>(q % p.wl);
        const int hq = static_cast<int>((q / p.wl) % p.hl);
        const int ni = static_cast<int>(q / (static_cast<int64_t>(p.wl) * p.hl));
        const float4 y4 = sc_ld4(yl + q * p.cl + col * 4);
#pragma unroll
        for (int a = 0; a < SW_MAX_ACC; ++a) {
          if (a < nacc) {
            const int tap = a / p.cs, cs = a - tap * p.cs;
            const int r = tap / p.kw, s = tap - r * p.kw;
            const int hx = hq + p.sign * (r - p.pad_t), wx = wq + p.sign * (s - p.pad_l);
            if (hx >= 0 && hx < p.hs && wx >= 0 && wx < p.ws) {
              const float xv = __ldg(p.xs + ((static_cast<int64_t>(ni) * p.hs + hx) * p.ws + wx) * p.cs + cs);
              acc[a].x += xv * y4.x; acc[a].y += xv * y4.y; acc[a].z += xv * y4.z; acc[a].w += xv * y4.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < SW_MAX_ACC; ++a) {
      if (a < nacc) {
        red[threadIdx.x] = acc[a];
        __syncthreads();
        if (ly == 0 && col < v) {
          float4 s4 = acc[a];
          for (int l = 1; l < lanes; ++l) {
            const float4 t = red[l * cols + cx];
            s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w;
          }
          sc_st4(p.partial + (static_cast<int64_t>(blockIdx.x) * nacc + a) * p.cl + col * 4, s4);
        }
        __syncthreads();
      }
    }
  }
}

// dw = beta*dw + scale * sum_chunks partial ; partial index [chunk][tap][cs][cl]; dw layout [tap][cs][cl] or
// (out_clcs) [tap][cl][cs]
__global__ void small_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int taps, int cs, int cl,
                                          int out_clcs, const float* __restrict__ scale, float beta,
                                          float* __restrict__ dw) {
  pdl_wait();
  const int total = taps * cs * cl;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[static_cast<int64_t>(k) * total + i];
  if (scale) s *= __ldg(scale);
  const int c = i % cl, j = (i / cl) % cs, tap = i / (cl * cs);
  const int64_t o = out_clcs ? (static_cast<int64_t>(tap) * cl + c) * cs + j : i;
  dw[o] = (beta != 0.f ? beta * dw[o] : 0.f) + s;
}

// ------------------------------------------------------------------------------------------------
// C[m,n] = beta*C[m,n] + alpha * sum_k A(m,k)*B(k,n) + bias[n]     generic strides, fp32, 16x16 tiles.
struct SgemmParams {
  const float* a; const float* b; float* c;
  int m, n, k;
  int64_t a_sm, a_sk, b_sk, b_sn;
  const float* alpha; const float* bias; float beta;
};

// 16 x 16 outputs per block, 64-deep k tiles: every thread has 8 independent global loads in flight per round trip
// (the 16-deep version spent one full memory latency per 16 k: 13 us for k = 300, all of it on D's critical chain).
__global__ void __launch_bounds__(256) sgemm_small_kernel(const SgemmParams p) {
  pdl_wait();
  constexpr int KT = 64;
  __shared__ float As[16][KT + 1], Bs[KT][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < p.k; k0 += KT) {
    float av[KT / 16], bv[KT / 16];
#pragma unroll
    for (int i = 0; i < KT / 16; ++i) {
      const int ka = k0 + tx + 16 * i, kb = k0 + ty + 16 * i;
      av[i] = (row < p.m && ka < p.k) ? p.a[row * p.a_sm + ka * p.a_sk] : 0.f;
      bv[i] = (kb < p.k && col < p.n) ? p.b[kb * p.b_sk + col * p.b_sn] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < KT / 16; ++i) {
      As[ty][tx + 16 * i] = av[i];
      Bs[ty + 16 * i][tx] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) acc += As[ty][kk] * Bs[kk][tx];
    __syncthreads();
  }
  if (row < p.m && col < p.n) {
    float v = acc * (p.alpha ? __ldg(p.alpha) : 1.f);
    if (p.bias) v += p.bias[col];
    float* o = p.c + static_cast<int64_t>(row) * p.n + col;
    *o = (p.beta != 0.f ? p.beta * *o : 0.f) + v;
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core route for the <=8-channel layers: an explicit bf16 im2col of the SMALL tensor (kh*kw*cs <= kpad
// columns, 64 bytes per pixel for kpad = 32) turns fprop / wgrad / dgrad into 1x1 GEMMs for conv_tc.cu.
//   out[(n,ho,wo)][(r*kw+s)*cs + c] = xs[n, ho + sign*(r-pad_t), wo + sign*(s-pad_l), c]   (0 outside / padding)
__global__ void __launch_bounds__(256)
im2col_small_kernel(const float* __restrict__ xs, __nv_bfloat16* __restrict__ out, int n, int hs, int ws, int cs,
                    int ho, int wo, int kh, int kw, int stride, int pad_t, int pad_l, int sign, int kpad) {
  // per-column offsets (dh, dw, channel) once per block instead of two integer divisions per element
  __shared__ int lut[128];
  const int kvalid = kh * kw * cs;
  for (int j = threadIdx.x; j < kpad; j += blockDim.x) {
    int packed = -1;
    if (j < kvalid) {
      const int tap = j / cs, c = j - tap * cs;
      const int r = tap / kw, sx = tap - r * kw;
      packed = ((sign * (r - pad_t) + 64) << 16) | ((sign * (sx - pad_l) + 64) << 8) | c;
    }
    lut[j] = packed;
  }
  pdl_wait();
  __syncthreads();
  const int groups = kpad >> 3;  // 8 bf16 = 16 bytes per thread
  const int64_t total = static_cast<int64_t>(n) * ho * wo * groups;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int gq = static_cast<int>(i % groups);
    const int64_t pix = i / groups;
    const int pq = static_cast<int>(pix);          // pixel counts fit 32 bits (checked by the launcher)
    const int w0 = pq % wo;
    const int t2 = pq / wo;
    const int h0 = t2 % ho;
    const int ni = t2 / ho;
    const float* xn = xs + static_cast<int64_t>(ni) * hs * ws * cs;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int packed = lut[gq * 8 + e];
      float val = 0.f;
      if (packed >= 0) {
        const int hx = h0 * stride + (packed >> 16) - 64, wx = w0 * stride + ((packed >> 8) & 0xff) - 64;
        if (hx >= 0 && hx < hs && wx >= 0 && wx < ws) val = __ldg(xn + (hx * ws + wx) * cs + (packed & 0xff));
      }
      v[e] = __float2bfloat16_rn(val);
    }
    *reinterpret_cast<uint4*>(out + pix * kpad + gq * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// out[l][tap*cs + c] (row stride kpad, zero padded) from the HWIO fp32 filter:
//   small_is_ci: l = co, c = ci -> W[tap][c][l]   (fprop operand of a small-cin layer)
//   otherwise  : l = ci, c = co -> W[tap][l][c]   (dgrad operand of a small-cout layer)
__global__ void pack_small_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int taps, int ci,
                                  int co, int small_is_ci, int kpad) {
  pdl_wait();
  const int nl = small_is_ci ? co : ci, cs = small_is_ci ? ci : co;
  const int total = nl * kpad;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int l = i / kpad, j = i - l * kpad;
  float v = 0.f;
  if (j < taps * cs) {
    const int tap = j / cs, c = j - tap * cs;
    v = small_is_ci ? w[(static_cast<int64_t>(tap) * ci + c) * co + l] : w[(static_cast<int64_t>(tap) * ci + l) * co + c];
  }
  out[i] = __float2bfloat16_rn(v);
}

// dw[tap][cs][cl] (or [tap][cl][cs] when out_clcs) = beta*dw + scale * r[tap*cs + c][l],  r is [kpad][cl] fp32
__global__ void small_wgrad_scatter_kernel(const float* __restrict__ r, float* __restrict__ dw, int taps, int cs,
                                           int cl, int out_clcs, const float* __restrict__ scale, float beta) {
  pdl_wait();
  const int total = taps * cs * cl;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int l = i % cl, j = i / cl;  // j = tap*cs + c
  float v = r[static_cast<int64_t>(j) * cl + l];
  if (scale) v *= __ldg(scale);
  const int tap = j / cs, c = j - tap * cs;
  const int64_t o = out_clcs ? (static_cast<int64_t>(tap) * cl + l) * cs + c : i;
  dw[o] = (beta != 0.f ? beta * dw[o] : 0.f) + v;
}

}  // namespace ganb

using namespace ganb;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ganb_conv2d_smallcin(const float* x, const float* w, void* y, int n, int h, int w_in, int cs, int ho,
                                    int wo, int cl, int kh, int kw, int pad_t, int pad_l, int flip_taps,
                                    int w_layout_clcs, const float* alpha, const float* bias, int act,
                                    int out_dtype, void* stream) {
  if (!x || !w || !y) return fail(GANB_E_BADARG, "conv2d_smallcin: null buffer");
  if (cl % 4 != 0) return fail(GANB_E_UNSUPPORTED, "conv2d_smallcin: cl=%d must be a multiple of 4", cl);
  const int smem = kh * kw * cs * cl * 4;
  if (cs > 8 || smem > 96 * 1024) return fail(GANB_E_UNSUPPORTED, "conv2d_smallcin: cs=%d / filter too large", cs);
  SmallCinParams p;
  p.x = x; p.w = w; p.y = y;
  p.n = n; p.h = h; p.w_in = w_in; p.cs = cs; p.ho = ho; p.wo = wo; p.cl = cl; p.kh = kh; p.kw = kw;
  p.pad_t = pad_t; p.pad_l = pad_l; p.flip = flip_taps; p.w_clcs = w_layout_clcs;
  p.alpha = alpha; p.bias = bias; p.act = act; p.out_bf16 = (out_dtype == GANB_BF16);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv_smallcin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "conv2d_smallcin: %s", cudaGetErrorString(e));
  }
  const int64_t items = static_cast<int64_t>(n) * ho * wo * (cl / 4);
  int64_t blocks = ceil_div64(items, 256 * 4);
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  if (blocks < 1) blocks = 1;
  launch_k(conv_smallcin_kernel, static_cast<int>(blocks), 256, smem, STREAM, p);
  GANB_CHECK_LAUNCH("conv_smallcin_kernel");
  return 0;
}

static int small_wgrad_chunks(int64_t total_pix) {
  int chunks = 2 * sm_count();
  const int64_t max_chunks = ceil_div64(total_pix, 64);
  if (chunks > max_chunks) chunks = static_cast<int>(max_chunks);
  if (chunks < 1) chunks = 1;
  return chunks;
}

extern "C" int64_t ganb_conv2d_small_wgrad_workspace(int n, int hl, int wl, int cs, int cl, int kh, int kw) {
  const int chunks = small_wgrad_chunks(static_cast<int64_t>(n) * hl * wl);
  return static_cast<int64_t>(chunks) * kh * kw * cs * cl * 4;
}

extern "C" int ganb_conv2d_small_wgrad(const float* xs, const void* yl, int yl_dtype, float* dw, void* workspace, int n,
                                       int hs, int ws, int cs, int hl, int wl, int cl, int kh, int kw, int pad_t,
                                       int pad_l, int sign, int out_layout_clcs, const float* scale, float beta,
                                       void* stream) {
  if (!xs || !yl || !dw || !workspace) return fail(GANB_E_BADARG, "conv2d_small_wgrad: null buffer");
  if (cl % 4 != 0) return fail(GANB_E_UNSUPPORTED, "conv2d_small_wgrad: cl=%d must be a multiple of 4", cl);
  if (kh * kw * cs > SW_MAX_ACC) return fail(GANB_E_UNSUPPORTED, "conv2d_small_wgrad: taps*cs=%d > %d", kh * kw * cs, SW_MAX_ACC);
  if (sign != 1 && sign != -1) return fail(GANB_E_BADARG, "conv2d_small_wgrad: sign must be +-1");
  const int64_t total_pix = static_cast<int64_t>(n) * hl * wl;
  const int chunks = small_wgrad_chunks(total_pix);
  SmallWgradParams p;
  p.xs = xs; p.yl = yl; p.yl_bf16 = (yl_dtype == GANB_BF16);
  p.n = n; p.hs = hs; p.ws = ws; p.cs = cs; p.hl = hl; p.wl = wl; p.cl = cl; p.kh = kh; p.kw = kw;
  p.pad_t = pad_t; p.pad_l = pad_l; p.sign = sign;
  p.pix_per_chunk = static_cast<int>(ceil_div64(total_pix, chunks));
  p.partial = static_cast<float*>(workspace);
  const int used = static_cast<int>(ceil_div64(total_pix, p.pix_per_chunk));
  const int smem = 256 * 16;
  if (p.yl_bf16) launch_k(conv_small_wgrad_kernel<__nv_bfloat16>, used, 256, smem, STREAM, p);
  else launch_k(conv_small_wgrad_kernel<float>, used, 256, smem, STREAM, p);
  GANB_CHECK_LAUNCH("conv_small_wgrad_kernel");
  const int total = kh * kw * cs * cl;
  launch_k(small_wgrad_reduce_kernel, ceil_div(total, 256), 256, 0, STREAM, p.partial, used, kh * kw, cs, cl, out_layout_clcs, scale, beta, dw);
  GANB_CHECK_LAUNCH("small_wgrad_reduce_kernel");
  return 0;
}

// C[m,n] = beta*C + alpha * op(A)[m,k] * op(B)[k,n] + bias[n];  trans_a: A stored [k,m]; trans_b: B stored [n,k]
extern "C" int ganb_sgemm_small(const float* a, const float* b, float* c, int m, int n, int k, int trans_a, int trans_b,
                                const float* alpha, const float* bias, float beta, void* stream) {
  if (!a || !b || !c) return fail(GANB_E_BADARG, "sgemm_small: null buffer");
  SgemmParams p;
  p.a = a; p.b = b; p.c = c; p.m = m; p.n = n; p.k = k;
  p.a_sm = trans_a ? 1 : k; p.a_sk = trans_a ? m : 1;
  p.b_sk = trans_b ? 1 : n; p.b_sn = trans_b ? k : 1;
  p.alpha = alpha; p.bias = bias; p.beta = beta;
  launch_k(sgemm_small_kernel, dim3(ceil_div(n, 16), ceil_div(m, 16)), 256, 0, STREAM, p);
  GANB_CHECK_LAUNCH("sgemm_small_kernel");
  return 0;
}

extern "C" int ganb_im2col_small(const float* xs, void* out_bf16, int n, int hs, int ws, int cs, int ho, int wo, int kh,
                                 int kw, int stride, int pad_t, int pad_l, int sign, int kpad, void* stream) {
  if (!xs || !out_bf16) return fail(GANB_E_BADARG, "im2col_small: null buffer");
  if (kpad % 8 != 0 || kh * kw * cs > kpad) return fail(GANB_E_BADARG, "im2col_small: kh*kw*cs=%d does not fit kpad=%d", kh * kw * cs, kpad);
  if (sign != 1 && sign != -1) return fail(GANB_E_BADARG, "im2col_small: sign must be +-1");
  if (kpad > 128 || kh > 32 || kw > 32 || static_cast<int64_t>(n) * ho * wo >= (1LL << 31))
    return fail(GANB_E_UNSUPPORTED, "im2col_small: kpad=%d (<= 128), kernel %dx%d (<= 32) or pixel count out of range", kpad, kh, kw);
  const int64_t items = static_cast<int64_t>(n) * ho * wo * (kpad / 8);
  int64_t blocks = ceil_div64(items, 256);
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  launch_k(im2col_small_kernel, static_cast<int>(blocks), 256, 0, STREAM, xs, static_cast<__nv_bfloat16*>(out_bf16), n, hs, ws, cs,
                                                                   ho, wo, kh, kw, stride, pad_t, pad_l, sign, kpad);
  GANB_CHECK_LAUNCH("im2col_small_kernel");
  return 0;
}

extern "C" int ganb_pack_small(const float* w_hwio, void* out_bf16, int taps, int ci, int co, int small_is_ci, int kpad,
                               void* stream) {
  if (!w_hwio || !out_bf16) return fail(GANB_E_BADARG, "pack_small: null buffer");
  const int nl = small_is_ci ? co : ci;
  launch_k(pack_small_kernel, ceil_div(nl * kpad, 256), 256, 0, STREAM, w_hwio, static_cast<__nv_bfloat16*>(out_bf16), taps, ci, co,
                                                                 small_is_ci, kpad);
  GANB_CHECK_LAUNCH("pack_small_kernel");
  return 0;
}

extern "C" int ganb_small_wgrad_scatter(const float* r, float* dw, int taps, int cs, int cl, int out_layout_clcs,
                                        const float* scale, float beta, void* stream) {
  if (!r || !dw) return fail(GANB_E_BADARG, "small_wgrad_scatter: null buffer");
  launch_k(small_wgrad_scatter_kernel, ceil_div(taps * cs * cl, 256), 256, 0, STREAM, r, dw, taps, cs, cl, out_layout_clcs, scale, beta);
  GANB_CHECK_LAUNCH("small_wgrad_scatter_kernel");
  return 0;
}
