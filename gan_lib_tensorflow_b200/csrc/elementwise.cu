// Bandwidth-bound kernels of the hot path: batch statistics, (conditional) normalise + activation + resample
// forward/backward, pooling, casts, bias gradients, label-map concat, global pooling, losses, Adam.
// All activations are NHWC; float4 / bf16x4 vectorised along the channel dimension (C % 4 == 0 fast path).
//
// Reference call-sites replaced: common/ops/normalization.py:8-59,105-140 (moments, batch_normalization,
// embedding_lookup of gamma/beta), common/resnet_block.py:24-29,62-63,71-72,87-88 (relu / leaky relu,
// mean-pool, nearest upsample), SNGAN/gan_cifar_resnet.py:282-284,301,334-337,376-378,492,521-526.
#include "host_common.h"

#include <stdlib.h>

#include <cuda_bf16.h>

namespace ganb {

// ------------------------------------------------------------------------------------------------ helpers
struct alignas(8) bf16x4 {
  __nv_bfloat162 lo, hi;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const bf16x4 v = *reinterpret_cast<const bf16x4*>(p);
  const float2 a = __bfloat1622float2(v.lo), b = __bfloat1622float2(v.hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  bf16x4 o;
  o.lo = __floats2bfloat162_rn(v.x, v.y);
  o.hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<bf16x4*>(p) = o;
}
// Branch-free forms (none / relu / leaky relu; the host sends tanh to the generic kernels): a per-element switch over the
// activation codes compiles to a branch per element with the tanh code inline (190 branches in a 2100-instruction
// kernel body), which -- not the memory system -- paced these kernels at ~0.5 of the HBM roofline.
struct ActLin {
  float slope;     // value multiplier on the negative side: none 1, relu 0, leaky relu 0.2
  bool relu;       // exact +0 on the negative side
  bool zero_on;    // derivative at y == 0: relu 0, otherwise 1
  __device__ __forceinline__ explicit ActLin(int act)
      : slope(act == GANB_ACT_RELU ? 0.f : act == GANB_ACT_LRELU ? 0.2f : 1.f), relu(act == GANB_ACT_RELU),
        zero_on(act != GANB_ACT_RELU) {}
  __device__ __forceinline__ float f(float v) const { return fmaxf(v, relu ? 0.f : slope * v); }
  __device__ __forceinline__ float d(float y) const { return (zero_on ? y >= 0.f : y > 0.f) ? 1.f : slope; }
};

// tanh out of line: inlined, its code sits behind a branch in every unrolled element of every caller
__device__ __noinline__ float tanh_ool(float v) { return tanhf(v); }
__device__ __forceinline__ float act_f(float v, int act) {
  if (act == GANB_ACT_TANH) return tanh_ool(v);
  return ActLin(act).f(v);
}
// derivative of the activation expressed with the PRE-activation value y
__device__ __forceinline__ float dact_f(float y, int act) {
  if (act == GANB_ACT_TANH) { const float t = tanh_ool(y); return 1.f - t * t; }
  return ActLin(act).d(y);
}
#define F4_OP(r, a, expr_x, expr_y, expr_z, expr_w) \
  float4 r = make_float4(expr_x, expr_y, expr_z, expr_w)

static inline int grid_for(int64_t work_items, int threads, int max_blocks_per_sm = 8) {
  int64_t b = ceil_div64(work_items, threads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * max_blocks_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------------------------------------ bn stats
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[(g*chunks + chunk)*2*C + {0,1}*C + c] = sum / sum of squares over the chunk's rows.
template <typename TX>
__global__ void __launch_bounds__(256)
bn_stats_partial_kernel(const TX* __restrict__ x, int rows_per_group, int c, int chunks, int rows_per_chunk,
                float* __restrict__ partial) {
  pdl_wait();
  const int g = blockIdx.y, chunk = blockIdx.x;
  const int v = c >> 2;                          // float4 columns
  const int lanes = max(1, 256 / min(v, 256));   // row lanes per column block
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(rows_per_group, r0 + rows_per_chunk);
  const TX* xg = x + static_cast<int64_t>(g) * rows_per_group * c;
  __shared__ float4 sh_s[256], sh_q[256];
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    if (col < v && ry < lanes) {
      constexpr int U = 4;
      for (int r = r0 + ry; r < r1; r += lanes * U) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          a[u] = (r + u * lanes < r1) ? ld4(xg + static_cast<int64_t>(r + u * lanes) * c + col * 4)
                                      : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          s.x += a[u].x; s.y += a[u].y; s.z += a[u].z; s.w += a[u].w;
          q.x += a[u].x * a[u].x; q.y += a[u].y * a[u].y; q.z += a[u].z * a[u].z; q.w += a[u].w * a[u].w;
        }
      }
    }
    sh_s[threadIdx.x] = s;
    sh_q[threadIdx.x] = q;
    __syncthreads();
    if (ry == 0 && col < v) {
      for (int l = 1; l < lanes; ++l) {
        const float4 a = sh_s[l * cols_per_pass + cx], b = sh_q[l * cols_per_pass + cx];
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
      }
      float* out = partial + (static_cast<int64_t>(g) * chunks + chunk) * 2 * c;
      st4(out + col * 4, s);
      st4(out + c + col * 4, q);
    }
    __syncthreads();
  }
}

// One block per (32 channels, group): 32 thread rows split the chunk partials -- every load is a coalesced 128-byte row of
// 32 channels, all of a thread's loads are independent (one round trip for <= 256 chunks) -- and meet in shared memory in a
// fixed order (deterministic).  (One warp per channel with the lanes over the chunks read 4 bytes per 32-byte sector.)
__global__ void __launch_bounds__(1024)
bn_stats_finalize_kernel(const float* __restrict__ partial, int c, int groups, int chunks, float inv_count, float eps,
                         float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_wait();
  __shared__ double ssum[32][33], ssq[32][33];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cx;
  const int g = blockIdx.y;
  double s = 0.0, q = 0.0;
  if (ch < c) {
#pragma unroll 8
    for (int k = ly; k < chunks; k += 32) {
      const float* p = partial + (static_cast<int64_t>(g) * chunks + k) * 2 * c;
      s += __ldg(p + ch);
      q += __ldg(p + c + ch);
    }
  }
  ssum[ly][cx] = s;
  ssq[ly][cx] = q;
  __syncthreads();
  // warp `ly` folds channel `ly`: lanes read the 32 row sums (transposed access, padded: conflict-free), fixed shuffle tree
  const int chw = blockIdx.x * 32 + ly;
  double a = ssum[cx][ly], b = ssq[cx][ly];
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  if (cx == 0 && chw < c) {
    const double m = a * inv_count;
    double var = b * inv_count - m * m;
    if (var < 0.0) var = 0.0;
    mean[g * c + chw] = static_cast<float>(m);
    rstd[g * c + chw] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

// ------------------------------------------------------------------------------------------------ norm+act fwd
struct NormActFwd {
  const void* x; int x_bf16;
  int n, h, w, c;
  const float* mean; const float* rstd; int groups;       // mean == nullptr: no normalisation
  const float* gamma; const float* beta; const int* labels;  // tables [n_labels, c]; labels == nullptr: row 0
  int act, upsample;
  int quad;   // x is stored in quad layout [n, h/2, w/2, 4 = 2i+j, c] (ganb_upconv_fprop); out is plain NHWC [n,h,w,c]
  void* out; int out_bf16; int out_cstride;               // channel stride of the output pixel (>= c)
  __nv_bfloat16* out_raw; int raw_cstride;                 // optional bf16 copy of x at input resolution
};

// grid (pixel chunks, n): a thread owns one float4 of channels, so scale/shift (and the label lookup) are
// computed once and the loop streams x with several independent 16-byte loads in flight.
template <typename TX>
__global__ void __launch_bounds__(256) norm_act_fwd_kernel(const NormActFwd p, int pix_per_chunk) {
  pdl_wait();
  const int ni = blockIdx.y;
  const int v = p.c >> 2;
  const int cols = min(v, 256);
  const int lanes = 256 / cols;
  const int cx = threadIdx.x % cols, ly = threadIdx.x / cols;
  if (ly >= lanes) return;
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(hw, p0 + pix_per_chunk);
  const int g = ni / (p.n / p.groups);
  for (int cb = 0; cb < v; cb += cols) {
    const int col = cb + cx;
    if (col >= v) continue;
    const int c4 = col * 4;
    float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
    if (p.mean) {
      const float4 m = ld4(p.mean + g * p.c + c4), r = ld4(p.rstd + g * p.c + c4);
      float4 ga = make_float4(1, 1, 1, 1), be = make_float4(0, 0, 0, 0);
      if (p.gamma) {
        const int row = p.labels ? __ldg(p.labels + ni) : 0;
        ga = ld4(p.gamma + static_cast<int64_t>(row) * p.c + c4);
        be = ld4(p.beta + static_cast<int64_t>(row) * p.c + c4);
      }
      // tf.nn.batch_normalization: inv = rsqrt(var+eps)*gamma; y = x*inv + (beta - mean*inv)
      sc = make_float4(r.x * ga.x, r.y * ga.y, r.z * ga.z, r.w * ga.w);
      sh = make_float4(be.x - m.x * sc.x, be.y - m.y * sc.y, be.z - m.z * sc.z, be.w - m.w * sc.w);
    }
    const TX* xb = static_cast<const TX*>(p.x) + static_cast<int64_t>(ni) * hw * p.c + c4;
    constexpr int U = 4;
    for (int px = p0 + ly; px < p1; px += lanes * U) {
      float4 a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = px + u * lanes;
        if (q < p1) a[u] = ld4(xb + static_cast<int64_t>(q) * p.c);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = px + u * lanes;
        if (q >= p1) break;
        const float4 y = make_float4(act_f(a[u].x * sc.x + sh.x, p.act), act_f(a[u].y * sc.y + sh.y, p.act),
                                     act_f(a[u].z * sc.z + sh.z, p.act), act_f(a[u].w * sc.w + sh.w, p.act));
        const int64_t pix = static_cast<int64_t>(ni) * hw + q;
        if (p.out_raw) st4(p.out_raw + pix * p.raw_cstride + c4, a[u]);
        if (!p.upsample) {
          const int64_t o = pix * p.out_cstride + c4;
          if (p.out_bf16) st4(reinterpret_cast<__nv_bfloat16*>(p.out) + o, y);
          else st4(reinterpret_cast<float*>(p.out) + o, y);
        } else {
          const int hi = q / p.w, wi = q - hi * p.w;
          const int ow = 2 * p.w;
          const int64_t base = (static_cast<int64_t>(ni) * 2 * p.h + 2 * hi) * ow + 2 * wi;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int64_t o = (base + dy * ow + dx) * p.out_cstride + c4;
              if (p.out_bf16) st4(reinterpret_cast<__nv_bfloat16*>(p.out) + o, y);
              else st4(reinterpret_cast<float*>(p.out) + o, y);
            }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ norm+act bwd
struct NormActBwd {
  const void* x; int x_bf16;      // forward input (pre-normalisation), [n,h,w,c], fp32 or bf16
  const void* dz; int dz_bf16; int dz_cstride;   // gradient of the forward OUTPUT ([n,(2)h,(2)w,*])
  int n, h, w, c;
  const float* mean; const float* rstd; int groups;
  const float* gamma; const float* beta; const int* labels;
  int act, upsample;
  int quad;   // x and dx are stored in quad layout [n, h/2, w/2, 4, c]; dz is plain NHWC [n,h,w,c]
  // reduce outputs: per-sample sums
  float* part;    // [n][chunks][2][c]
  int chunks, pix_per_chunk;
  // apply inputs/outputs
  const float* s1; const float* s2;   // [groups, c]  sum(dxhat), sum(dxhat*xhat)
  float inv_count;
  const void* add; int add_bf16;      // optional fp32 / bf16 tensor added to dx (second gradient path)
  void* dx; int dx_bf16;
};

// upstream gradient of one float4 of channels (summed over the 2x2 replicas when the forward upsampled);
// `px` is the pixel index inside the sample
template <bool UPS, bool DZ16>
__device__ __forceinline__ float4 norm_act_bwd_load_dz(const NormActBwd& p, int ni, int px, int c4) {
  if (!UPS) {
    const int64_t o = (static_cast<int64_t>(ni) * p.h * p.w + px) * p.dz_cstride + c4;
    return DZ16 ? ld4(reinterpret_cast<const __nv_bfloat16*>(p.dz) + o) : ld4(reinterpret_cast<const float*>(p.dz) + o);
  }
  const int hi = px / p.w, wi = px - hi * p.w;
  const int ow = 2 * p.w;
  const int64_t base = (static_cast<int64_t>(ni) * 2 * p.h + 2 * hi) * ow + 2 * wi;
  float4 t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t o = (base + (k >> 1) * ow + (k & 1)) * p.dz_cstride + c4;
    t[k] = DZ16 ? ld4(reinterpret_cast<const __nv_bfloat16*>(p.dz) + o) : ld4(reinterpret_cast<const float*>(p.dz) + o);
  }
  return make_float4(t[0].x + t[1].x + t[2].x + t[3].x, t[0].y + t[1].y + t[2].y + t[3].y,
                     t[0].z + t[1].z + t[2].z + t[3].z, t[0].w + t[1].w + t[2].w + t[3].w);
}

// per-thread constants of the backward pass for one float4 of channels of one sample
struct NormActBwdConst {
  float4 m, r, ga, be;
};
template <bool NORM>
__device__ __forceinline__ NormActBwdConst norm_act_bwd_const(const NormActBwd& p, int ni, int g, int c4) {
  NormActBwdConst k;
  k.m = make_float4(0, 0, 0, 0); k.r = make_float4(1, 1, 1, 1);
  k.ga = make_float4(1, 1, 1, 1); k.be = make_float4(0, 0, 0, 0);
  if (NORM) {
    k.m = ld4(p.mean + g * p.c + c4);
    k.r = ld4(p.rstd + g * p.c + c4);
    if (p.gamma) {
      const int row = p.labels ? __ldg(p.labels + ni) : 0;
      k.ga = ld4(p.gamma + static_cast<int64_t>(row) * p.c + c4);
      k.be = ld4(p.beta + static_cast<int64_t>(row) * p.c + c4);
    }
  }
  return k;
}
// dy = dz * act'(y) and xhat, from the forward input a and the upstream gradient dz
template <bool NORM>
__device__ __forceinline__ void norm_act_bwd_math(const NormActBwd& p, const NormActBwdConst& k, float4 a, float4 dz,
                                                  float4& dy, float4& xhat) {
  float4 y = a;
  xhat = a;
  if (NORM) {
    xhat = make_float4((a.x - k.m.x) * k.r.x, (a.y - k.m.y) * k.r.y, (a.z - k.m.z) * k.r.z, (a.w - k.m.w) * k.r.w);
    y = make_float4(xhat.x * k.ga.x + k.be.x, xhat.y * k.ga.y + k.be.y, xhat.z * k.ga.z + k.be.z, xhat.w * k.ga.w + k.be.w);
  }
  dy = make_float4(dz.x * dact_f(y.x, p.act), dz.y * dact_f(y.y, p.act), dz.z * dact_f(y.z, p.act),
                   dz.w * dact_f(y.w, p.act));
}

// grid (chunks, n); per-sample partial sums A = sum dy, B = sum dy*xhat
template <bool UPS, bool DZ16, typename TX>
__global__ void __launch_bounds__(256, 3) norm_act_bwd_reduce_kernel(const NormActBwd p) {
  pdl_wait();
  const int ni = blockIdx.y, chunk = blockIdx.x;
  const int v = p.c >> 2;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  const int hw = p.h * p.w;
  const int p0 = chunk * p.pix_per_chunk, p1 = min(hw, p0 + p.pix_per_chunk);
  const int n_per_group = p.n / p.groups;
  const int g = ni / n_per_group;
  const TX* xp = static_cast<const TX*>(p.x);
  __shared__ float4 sh_a[256], sh_b[256];
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float4 sa = make_float4(0, 0, 0, 0), sb = make_float4(0, 0, 0, 0);
    if (col < v && ry < lanes) {
      const NormActBwdConst kc = norm_act_bwd_const<true>(p, ni, g, col * 4);
      constexpr int U = 4;
      for (int px = p0 + ry; px < p1; px += lanes * U) {
        float4 a[U], dz[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int qx = px + u * lanes;
          if (qx < p1) {
            a[u] = ld4(xp + (static_cast<int64_t>(ni) * hw + qx) * p.c + col * 4);
            dz[u] = norm_act_bwd_load_dz<UPS, DZ16>(p, ni, qx, col * 4);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (px + u * lanes >= p1) break;
          float4 dy, xh;
          norm_act_bwd_math<true>(p, kc, a[u], dz[u], dy, xh);
          sa.x += dy.x; sa.y += dy.y; sa.z += dy.z; sa.w += dy.w;
          sb.x += dy.x * xh.x; sb.y += dy.y * xh.y; sb.z += dy.z * xh.z; sb.w += dy.w * xh.w;
        }
      }
    }
    sh_a[threadIdx.x] = sa;
    sh_b[threadIdx.x] = sb;
    __syncthreads();
    if (ry == 0 && col < v) {
      for (int l = 1; l < lanes; ++l) {
        const float4 a = sh_a[l * cols_per_pass + cx], b = sh_b[l * cols_per_pass + cx];
        sa.x += a.x; sa.y += a.y; sa.z += a.z; sa.w += a.w;
        sb.x += b.x; sb.y += b.y; sb.z += b.z; sb.w += b.w;
      }
      float* out = p.part + (static_cast<int64_t>(ni) * p.chunks + chunk) * 2 * p.c;
      st4(out + col * 4, sa);
      st4(out + p.c + col * 4, sb);
    }
    __syncthreads();
  }
}

// Stage 1: per-sample sums A_n = sum_chunks, B_n, and the group sums S1 = sum gamma*A, S2 = sum gamma*B.
// One warp per (group, channel): lanes split the samples, fixed shuffle tree (deterministic). grid (c/8, groups).
__global__ void __launch_bounds__(256)
norm_act_bwd_finalize_kernel(const float* __restrict__ part, int n, int c, int chunks, int groups,
                             const float* __restrict__ gamma, const int* __restrict__ labels,
                             float* __restrict__ sums /*[n][2][c]*/, float* __restrict__ s1, float* __restrict__ s2) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int g = blockIdx.y;
  if (ch >= c) return;
  const int n_per_group = n / groups;
  float acc1 = 0.f, acc2 = 0.f;
  for (int ni = g * n_per_group + lane; ni < (g + 1) * n_per_group; ni += 32) {
    // the label -> gamma lookup is issued before the partial sums so that its two dependent round trips overlap them
    float ga = 1.f;
    if (gamma) ga = __ldg(gamma + static_cast<int64_t>(labels ? __ldg(labels + ni) : 0) * c + ch);
    float a = 0.f, b = 0.f;
#pragma unroll 4
    for (int k = 0; k < chunks; ++k) {
      const float* q = part + (static_cast<int64_t>(ni) * chunks + k) * 2 * c;
      a += __ldg(q + ch);
      b += __ldg(q + c + ch);
    }
    sums[(static_cast<int64_t>(ni) * 2) * c + ch] = a;
    sums[(static_cast<int64_t>(ni) * 2 + 1) * c + ch] = b;
    acc1 += ga * a;
    acc2 += ga * b;
  }
  acc1 = warp_sum_f(acc1);
  acc2 = warp_sum_f(acc2);
  if (lane == 0) {
    s1[g * c + ch] = acc1;
    s2[g * c + ch] = acc2;
  }
}

// Stage 2: dgamma[row, ch] += sum_{n: label_n == row} B_n ; dbeta likewise with A_n. One warp per (row, ch): lanes
// split the samples and combine with a fixed shuffle tree (deterministic scatter of the lookup's gradient).
__global__ void __launch_bounds__(256)
norm_act_bwd_scatter_kernel(const float* __restrict__ sums, int n, int c, int n_rows, const int* __restrict__ labels,
                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_rows * c) return;
  const int row = i / c, ch = i - row * c;
  float a = 0.f, b = 0.f;
  for (int ni = lane; ni < n; ni += 32) {
    if ((labels ? labels[ni] : 0) == row) {
      a += sums[(static_cast<int64_t>(ni) * 2) * c + ch];
      b += sums[(static_cast<int64_t>(ni) * 2 + 1) * c + ch];
    }
  }
  a = warp_sum_f(a);
  b = warp_sum_f(b);
  if (lane == 0) {
    dbeta[i] += a;
    dgamma[i] += b;
  }
}

// Stages 1 + 2 in one launch for the training shapes (n <= 256 samples, label tables of <= 32 rows): one block per 16
// channels keeps the per-sample sums of ALL samples in shared memory, so the group sums and the label scatter follow
// without a second pass through global memory, and every load of the chunk partials is a coalesced float4 (the two-kernel
// form reads them 4 bytes per lane with a sample stride: 11 + 4 us per normalisation on the critical chain of the
// generator's backward pass, 7 of them per step).  Fixed summation orders (deterministic).
// threads: 4 float4 columns x 64 sample lanes.  dynamic shared memory: A[n][16] | B[n][16] | red1[64][4] f4 | red2 | labels[n]
__global__ void __launch_bounds__(256)
norm_act_bwd_finalize_fused_kernel(const float* __restrict__ part, int n, int c, int chunks, int groups,
                                   const float* __restrict__ gamma, const int* __restrict__ labels, int rows,
                                   float* __restrict__ sums, float* __restrict__ s1, float* __restrict__ s2,
                                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_wait();
  extern __shared__ __align__(16) float fsm[];
  float* smA = fsm;
  float* smB = smA + static_cast<size_t>(n) * 16;
  float4* red1 = reinterpret_cast<float4*>(smB + static_cast<size_t>(n) * 16);
  float4* red2 = red1 + 256;
  int* lab = reinterpret_cast<int*>(red2 + 256);
  const int col = threadIdx.x & 3, lane = threadIdx.x >> 2;
  const int ch = blockIdx.x * 16 + col * 4;
  const bool ch_ok = ch < c;                       // c % 4 == 0: a float4 column is inside or outside as a whole
  for (int i = threadIdx.x; i < n; i += 256) lab[i] = labels ? __ldg(labels + i) : 0;
  __syncthreads();
  const int npg = n / groups;
  for (int g = 0; g < groups; ++g) {
    float4 acc1 = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc1;
    for (int ni = g * npg + lane; ni < (g + 1) * npg; ni += 64) {
      float4 ga = make_float4(1.f, 1.f, 1.f, 1.f);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (ch_ok) {
        if (gamma) ga = __ldg(reinterpret_cast<const float4*>(gamma + static_cast<int64_t>(lab[ni]) * c + ch));
#pragma unroll 4
        for (int k = 0; k < chunks; ++k) {
          const float* q = part + (static_cast<int64_t>(ni) * chunks + k) * 2 * c + ch;
          const float4 va = __ldg(reinterpret_cast<const float4*>(q));
          const float4 vb = __ldg(reinterpret_cast<const float4*>(q + c));
          a.x += va.x; a.y += va.y; a.z += va.z; a.w += va.w;
          b.x += vb.x; b.y += vb.y; b.z += vb.z; b.w += vb.w;
        }
        st4(sums + (static_cast<int64_t>(ni) * 2) * c + ch, a);
        st4(sums + (static_cast<int64_t>(ni) * 2 + 1) * c + ch, b);
      }
      st4(smA + ni * 16 + col * 4, a);
      st4(smB + ni * 16 + col * 4, b);
      acc1.x += ga.x * a.x; acc1.y += ga.y * a.y; acc1.z += ga.z * a.z; acc1.w += ga.w * a.w;
      acc2.x += ga.x * b.x; acc2.y += ga.y * b.y; acc2.z += ga.z * b.z; acc2.w += ga.w * b.w;
    }
    red1[threadIdx.x] = acc1;     // index = lane * 4 + col
    red2[threadIdx.x] = acc2;
    __syncthreads();
#pragma unroll
    for (int s = 32; s > 0; s >>= 1) {
      if (lane < s) {
        const float4 u = red1[threadIdx.x + s * 4], v = red2[threadIdx.x + s * 4];
        float4 x = red1[threadIdx.x], y = red2[threadIdx.x];
        x.x += u.x; x.y += u.y; x.z += u.z; x.w += u.w;
        y.x += v.x; y.y += v.y; y.z += v.z; y.w += v.w;
        red1[threadIdx.x] = x;
        red2[threadIdx.x] = y;
      }
      __syncthreads();
    }
    if (lane == 0 && ch_ok) {
      st4(s1 + static_cast<int64_t>(g) * c + ch, red1[col]);
      st4(s2 + static_cast<int64_t>(g) * c + ch, red2[col]);
    }
    __syncthreads();
  }
  if (dgamma) {
    // dgamma[row, ch] += sum_{n: label_n == row} B_n, dbeta with A_n: samples in ascending order
    for (int i = threadIdx.x; i < rows * 16; i += 256) {
      const int row = i >> 4, cc = i & 15;
      const int ch2 = blockIdx.x * 16 + cc;
      if (ch2 >= c) continue;
      float a = 0.f, b = 0.f;
      for (int ni = 0; ni < n; ++ni) {
        if (lab[ni] == row) {
          a += smA[ni * 16 + cc];
          b += smB[ni * 16 + cc];
        }
      }
      dbeta[static_cast<int64_t>(row) * c + ch2] += a;
      dgamma[static_cast<int64_t>(row) * c + ch2] += b;
    }
  }
}

template <bool NORM, bool UPS, bool DZ16, typename TX>
__global__ void __launch_bounds__(256, 3) norm_act_bwd_apply_kernel(const NormActBwd p, int pix_per_chunk) {
  pdl_wait();
  const int ni = blockIdx.y;
  const TX* xp = static_cast<const TX*>(p.x);
  const int v = p.c >> 2;
  const int cols = min(v, 256);
  const int lanes = 256 / cols;
  const int cx = threadIdx.x % cols, ly = threadIdx.x / cols;
  if (ly >= lanes) return;
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(hw, p0 + pix_per_chunk);
  const int g = ni / (p.n / p.groups);
  const float k = p.inv_count;
  for (int cb = 0; cb < v; cb += cols) {
    const int col = cb + cx;
    if (col >= v) continue;
    const int c4 = col * 4;
    float4 t1 = make_float4(0, 0, 0, 0), t2 = make_float4(0, 0, 0, 0);
    if (NORM) {
      t1 = ld4(p.s1 + g * p.c + c4);
      t2 = ld4(p.s2 + g * p.c + c4);
      t1 = make_float4(k * t1.x, k * t1.y, k * t1.z, k * t1.w);
      t2 = make_float4(k * t2.x, k * t2.y, k * t2.z, k * t2.w);
    }
    const NormActBwdConst kc = norm_act_bwd_const<NORM>(p, ni, g, c4);
    // loads of a whole batch are issued before its stores (the output may alias the inputs for the compiler)
    constexpr int U = 4;
    for (int px = p0 + ly; px < p1; px += lanes * U) {
      float4 a[U], dz[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int qx = px + u * lanes;
        if (qx < p1) {
          a[u] = ld4(xp + (static_cast<int64_t>(ni) * hw + qx) * p.c + c4);
          dz[u] = norm_act_bwd_load_dz<UPS, DZ16>(p, ni, qx, c4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int qx = px + u * lanes;
        if (qx >= p1) break;
        const int64_t pix = static_cast<int64_t>(ni) * hw + qx;
        float4 dy, xh;
        norm_act_bwd_math<NORM>(p, kc, a[u], dz[u], dy, xh);
        float4 dx = dy;
        if (NORM)
          dx = make_float4(kc.r.x * (kc.ga.x * dy.x - t1.x - xh.x * t2.x), kc.r.y * (kc.ga.y * dy.y - t1.y - xh.y * t2.y),
                           kc.r.z * (kc.ga.z * dy.z - t1.z - xh.z * t2.z), kc.r.w * (kc.ga.w * dy.w - t1.w - xh.w * t2.w));
        if (p.add) {
          const float4 q = p.add_bf16 ? ld4(static_cast<const __nv_bfloat16*>(p.add) + pix * p.c + c4)
                                      : ld4(static_cast<const float*>(p.add) + pix * p.c + c4);
          dx.x += q.x; dx.y += q.y; dx.z += q.z; dx.w += q.w;
        }
        if (p.dx_bf16) st4(reinterpret_cast<__nv_bfloat16*>(p.dx) + pix * p.c + c4, dx);
        else st4(reinterpret_cast<float*>(p.dx) + pix * p.c + c4, dx);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ bf16 x 8 paths
// All-bf16 variants of the statistics / normalise kernels (input, upstream gradient and output in bf16, c % 8 == 0):
// a thread owns 8 channels, i.e. one 16-byte load per tensor per pixel.  The generic kernels above move 8 bytes per
// load on bf16 data and were bound by loads in flight, not by HBM (same time for bf16 as for fp32 inputs).
__device__ __forceinline__ void cvt8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}
__device__ __forceinline__ void ldf8(const float* p, float (&f)[8]) {
  const float4 a = ld4(p), b = ld4(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// block-level sum over the `lanes` row lanes of 16 per-thread values; result valid in lane 0 threads (ry == 0)
__device__ __forceinline__ void lanes_sum16(float (&a)[8], float (&b)[8], int lanes, int cols_per_pass, int cx, int ry,
                                            float (*sh)[17]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh[threadIdx.x][j] = a[j]; sh[threadIdx.x][8 + j] = b[j]; }
  __syncthreads();
  if (ry == 0) {
    for (int l = 1; l < lanes; ++l) {
      const float* o = sh[l * cols_per_pass + cx];
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] += o[j]; b[j] += o[8 + j]; }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ cp.async rings
// The streaming reductions below are bound by HBM latency x bytes in flight, not by instruction issue: with plain
// loads a thread holds 4 x 16 B per tensor in registers and nothing is in flight while it computes (ncu: 3.1-3.7 TB/s,
// profiles/r01_ncu_norm_kernels.txt).  The *_v8p kernels issue the same 16-byte accesses as cp.async into a per-thread
// shared-memory ring (STAGES iterations deep, no registers held, no block-level synchronisation: every thread reads
// back only what it copied), which keeps 2-3x more bytes in flight per SM.
__device__ __forceinline__ void cp_async16(uint4* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = pred ? 16 : 0;                 // src-size 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int PIPE_U = 4;        // 16-byte accesses per tensor per iteration and thread
constexpr int STATS_STAGES = 3;  // bn_stats: 3 x 4 x 256 x 16 B = 48 KB per block
constexpr int BWD_STAGES = 2;    // backward kernels: up to 3 tensors x 2 x 4 x 256 x 16 B = 96 KB per block

__global__ void __launch_bounds__(256)
bn_stats_partial_v8p_kernel(const __nv_bfloat16* __restrict__ x, int rows_per_group, int c, int chunks,
                            int rows_per_chunk, float* __restrict__ partial) {
  pdl_wait();
  extern __shared__ uint4 ring_raw[];
  uint4* ring = ring_raw + threadIdx.x;        // [stage][u][256 threads]
  const int g = blockIdx.y, chunk = blockIdx.x;
  const int v = c >> 3;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(rows_per_group, r0 + rows_per_chunk);
  const __nv_bfloat16* xg = x + static_cast<int64_t>(g) * rows_per_group * c;
  __shared__ float sh[256][17];
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    if (col < v && ry < lanes) {
      const __nv_bfloat16* xc = xg + col * 8;
      const int step = lanes * PIPE_U;
      const int iters = (r1 - r0 - ry + step - 1) / step;
      auto issue = [&](int it) {
        uint4* dst = ring + (it % STATS_STAGES) * PIPE_U * 256;
#pragma unroll
        for (int u = 0; u < PIPE_U; ++u) {
          const int r = r0 + ry + (it * PIPE_U + u) * lanes;
          const bool ok = r < r1;
          cp_async16(dst + u * 256, xc + static_cast<int64_t>(ok ? r : r0) * c, ok);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int it = 0; it < STATS_STAGES - 1; ++it) issue(it);
      for (int it = 0; it < iters; ++it) {
        issue(it + STATS_STAGES - 1);
        cp_async_wait<STATS_STAGES - 1>();
        const uint4* src = ring + (it % STATS_STAGES) * PIPE_U * 256;
#pragma unroll
        for (int u = 0; u < PIPE_U; ++u) {
          float a[8];
          cvt8(src[u * 256], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) { s[j] += a[j]; q[j] += a[j] * a[j]; }
        }
      }
      cp_async_wait<0>();
    }
    lanes_sum16(s, q, lanes, cols_per_pass, cx, ry, sh);
    if (ry == 0 && col < v) {
      float* out = partial + (static_cast<int64_t>(g) * chunks + chunk) * 2 * c;
      st4(out + col * 8, make_float4(s[0], s[1], s[2], s[3]));
      st4(out + col * 8 + 4, make_float4(s[4], s[5], s[6], s[7]));
      st4(out + c + col * 8, make_float4(q[0], q[1], q[2], q[3]));
      st4(out + c + col * 8 + 4, make_float4(q[4], q[5], q[6], q[7]));
    }
  }
}

// pixel index inside one sample: quad layout [h/2][w/2][2i+j] -> NHWC row-major [h][w]
__device__ __forceinline__ int quad_to_nhwc(int q, int w) {
  const int g = q & 3, cell = q >> 2;
  const int w2 = w >> 1;
  const int a = cell / w2, b = cell - a * w2;
  return (2 * a + (g >> 1)) * w + 2 * b + (g & 1);
}

template <bool UPS>
__global__ void __launch_bounds__(256) norm_act_fwd_v8_kernel(const NormActFwd p, int pix_per_chunk) {
  pdl_wait();
  const ActLin al(p.act);
  const int ni = blockIdx.y;
  const int v = p.c >> 3;
  const int cols = min(v, 256);
  const int lanes = 256 / cols;
  const int cx = threadIdx.x % cols, ly = threadIdx.x / cols;
  if (ly >= lanes) return;
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(hw, p0 + pix_per_chunk);
  const int g = ni / (p.n / p.groups);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(p.x);
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out);
  for (int cb = 0; cb < v; cb += cols) {
    const int col = cb + cx;
    if (col >= v) continue;
    const int c8 = col * 8;
    float sc[8], sf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sf[j] = 0.f; }
    if (p.mean) {
      float m[8], r[8];
      ldf8(p.mean + g * p.c + c8, m);
      ldf8(p.rstd + g * p.c + c8, r);
      float ga[8], be[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { ga[j] = 1.f; be[j] = 0.f; }
      if (p.gamma) {
        const int row = p.labels ? __ldg(p.labels + ni) : 0;
        ldf8(p.gamma + static_cast<int64_t>(row) * p.c + c8, ga);
        ldf8(p.beta + static_cast<int64_t>(row) * p.c + c8, be);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = r[j] * ga[j]; sf[j] = be[j] - m[j] * sc[j]; }
    }
    const __nv_bfloat16* xb = xp + static_cast<int64_t>(ni) * hw * p.c + c8;
    constexpr int U = 4;
    for (int px = p0 + ly; px < p1; px += lanes * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = px + u * lanes;
        raw[u] = (q < p1) ? ldg16(xb + static_cast<int64_t>(q) * p.c) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = px + u * lanes;
        if (q >= p1) break;
        float a[8];
        cvt8(raw[u], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = al.f(a[j] * sc[j] + sf[j]);
        const uint4 y = pack8(a);
        const int64_t pix = static_cast<int64_t>(ni) * hw + (p.quad ? quad_to_nhwc(q, p.w) : q);
        if (!UPS) {
          stg16(op + pix * p.out_cstride + c8, y);
        } else {
          const int hi = q / p.w, wi = q - hi * p.w;
          const int ow = 2 * p.w;
          const int64_t base = (static_cast<int64_t>(ni) * 2 * p.h + 2 * hi) * ow + 2 * wi;
          stg16(op + base * p.out_cstride + c8, y);
          stg16(op + (base + 1) * p.out_cstride + c8, y);
          stg16(op + (base + ow) * p.out_cstride + c8, y);
          stg16(op + (base + ow + 1) * p.out_cstride + c8, y);
        }
      }
    }
  }
}

// per-thread constants of the backward pass for 8 channels of one sample: xhat = a*r - mr ; y = xhat*ga + be
struct BwdConst8 {
  float r[8], mr[8], ga[8], be[8];
};
template <bool NORM>
__device__ __forceinline__ void bwd_const8(const NormActBwd& p, int ni, int g, int c8, BwdConst8& k) {
#pragma unroll
  for (int j = 0; j < 8; ++j) { k.r[j] = 1.f; k.mr[j] = 0.f; k.ga[j] = 1.f; k.be[j] = 0.f; }
  if (NORM) {
    float m[8];
    ldf8(p.mean + g * p.c + c8, m);
    ldf8(p.rstd + g * p.c + c8, k.r);
#pragma unroll
    for (int j = 0; j < 8; ++j) k.mr[j] = m[j] * k.r[j];
    if (p.gamma) {
      const int row = p.labels ? __ldg(p.labels + ni) : 0;
      ldf8(p.gamma + static_cast<int64_t>(row) * p.c + c8, k.ga);
      ldf8(p.beta + static_cast<int64_t>(row) * p.c + c8, k.be);
    }
  }
}
// raw upstream gradient of 8 channels at pixel px of sample ni (UPS: the four replicas, summed by the caller)
template <bool UPS>
__device__ __forceinline__ void bwd_load_dz8(const NormActBwd& p, int ni, int px, int c8, uint4 (&raw)[UPS ? 4 : 1]) {
  const __nv_bfloat16* dz = static_cast<const __nv_bfloat16*>(p.dz);
  if (!UPS) {
    const int qz = p.quad ? quad_to_nhwc(px, p.w) : px;
    raw[0] = ldg16(dz + (static_cast<int64_t>(ni) * p.h * p.w + qz) * p.dz_cstride + c8);
  } else {
    const int hi = px / p.w, wi = px - hi * p.w;
    const int ow = 2 * p.w;
    const int64_t base = (static_cast<int64_t>(ni) * 2 * p.h + 2 * hi) * ow + 2 * wi;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      raw[k] = ldg16(dz + (base + (k >> 1) * ow + (k & 1)) * p.dz_cstride + c8);
  }
}
template <bool UPS>
__device__ __forceinline__ void bwd_sum_dz8(const uint4 (&raw)[UPS ? 4 : 1], float (&dz)[8]) {
  cvt8(raw[0], dz);
  if (UPS) {
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      float t[8];
      cvt8(raw[k], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) dz[j] += t[j];
    }
  }
}

template <bool UPS>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_reduce_v8_kernel(const NormActBwd p) {
  pdl_wait();
  const ActLin al(p.act);
  const int ni = blockIdx.y, chunk = blockIdx.x;
  const int v = p.c >> 3;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  const int hw = p.h * p.w;
  const int p0 = chunk * p.pix_per_chunk, p1 = min(hw, p0 + p.pix_per_chunk);
  const int g = ni / (p.n / p.groups);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(p.x);
  __shared__ float sh[256][17];
  constexpr int U = UPS ? 2 : 4;
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float sa[8], sb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sa[j] = 0.f; sb[j] = 0.f; }
    if (col < v && ry < lanes) {
      BwdConst8 k;
      bwd_const8<true>(p, ni, g, col * 8, k);
      for (int px = p0 + ry; px < p1; px += lanes * U) {
        uint4 xa[U];
        uint4 dzr[U][UPS ? 4 : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int qx = px + u * lanes;
          if (qx < p1) {
            xa[u] = ldg16(xp + (static_cast<int64_t>(ni) * hw + qx) * p.c + col * 8);
            bwd_load_dz8<UPS>(p, ni, qx, col * 8, dzr[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (px + u * lanes >= p1) break;
          float a[8], dz[8];
          cvt8(xa[u], a);
          bwd_sum_dz8<UPS>(dzr[u], dz);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = a[j] * k.r[j] - k.mr[j];
            const float dy = dz[j] * al.d(xh * k.ga[j] + k.be[j]);
            sa[j] += dy;
            sb[j] += dy * xh;
          }
        }
      }
    }
    lanes_sum16(sa, sb, lanes, cols_per_pass, cx, ry, sh);
    if (ry == 0 && col < v) {
      float* out = p.part + (static_cast<int64_t>(ni) * p.chunks + chunk) * 2 * p.c;
      st4(out + col * 8, make_float4(sa[0], sa[1], sa[2], sa[3]));
      st4(out + col * 8 + 4, make_float4(sa[4], sa[5], sa[6], sa[7]));
      st4(out + p.c + col * 8, make_float4(sb[0], sb[1], sb[2], sb[3]));
      st4(out + p.c + col * 8 + 4, make_float4(sb[4], sb[5], sb[6], sb[7]));
    }
  }
}

template <bool NORM, bool UPS>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_apply_v8_kernel(const NormActBwd p, int pix_per_chunk) {
  pdl_wait();
  const ActLin al(p.act);
  const int ni = blockIdx.y;
  const int v = p.c >> 3;
  const int cols = min(v, 256);
  const int lanes = 256 / cols;
  const int cx = threadIdx.x % cols, ly = threadIdx.x / cols;
  if (ly >= lanes) return;
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(hw, p0 + pix_per_chunk);
  const int g = ni / (p.n / p.groups);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* addp = static_cast<const __nv_bfloat16*>(p.add);
  __nv_bfloat16* dxp = static_cast<__nv_bfloat16*>(p.dx);
  constexpr int U = UPS ? 2 : 4;
  for (int cb = 0; cb < v; cb += cols) {
    const int col = cb + cx;
    if (col >= v) continue;
    const int c8 = col * 8;
    float t1[8], t2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { t1[j] = 0.f; t2[j] = 0.f; }
    if (NORM) {
      ldf8(p.s1 + g * p.c + c8, t1);
      ldf8(p.s2 + g * p.c + c8, t2);
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1[j] *= p.inv_count; t2[j] *= p.inv_count; }
    }
    BwdConst8 k;
    bwd_const8<NORM>(p, ni, g, c8, k);
    for (int px = p0 + ly; px < p1; px += lanes * U) {
      uint4 xa[U], ad[U];
      uint4 dzr[U][UPS ? 4 : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int qx = px + u * lanes;
        if (qx < p1) {
          const int64_t pix = static_cast<int64_t>(ni) * hw + qx;
          xa[u] = ldg16(xp + pix * p.c + c8);
          bwd_load_dz8<UPS>(p, ni, qx, c8, dzr[u]);
          if (addp) ad[u] = ldg16(addp + pix * p.c + c8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int qx = px + u * lanes;
        if (qx >= p1) break;
        const int64_t pix = static_cast<int64_t>(ni) * hw + qx;
        float a[8], dz[8], dx[8];
        cvt8(xa[u], a);
        bwd_sum_dz8<UPS>(dzr[u], dz);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = a[j] * k.r[j] - k.mr[j];
          const float dy = dz[j] * al.d(xh * k.ga[j] + k.be[j]);
          dx[j] = NORM ? k.r[j] * (k.ga[j] * dy - t1[j] - xh * t2[j]) : dy;
        }
        if (addp) {
          float q[8];
          cvt8(ad[u], q);
#pragma unroll
          for (int j = 0; j < 8; ++j) dx[j] += q[j];
        }
        stg16(dxp + pix * p.c + c8, pack8(dx));
      }
    }
  }
}


// cp.async-ring variant of the backward apply kernel for the common case (no upsample in the forward pass).
// element offset of the upstream gradient of pixel px (sample ni)
__device__ __forceinline__ int64_t bwd_dz_off(const NormActBwd& p, int ni, int px, int c8) {
  const int qz = p.quad ? quad_to_nhwc(px, p.w) : px;
  return (static_cast<int64_t>(ni) * p.h * p.w + qz) * p.dz_cstride + c8;
}

template <bool NORM>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_apply_v8p_kernel(const NormActBwd p, int pix_per_chunk) {
  pdl_wait();
  const ActLin al(p.act);
  extern __shared__ uint4 ring_raw[];
  uint4* ring_x = ring_raw + threadIdx.x;
  uint4* ring_z = ring_x + BWD_STAGES * PIPE_U * 256;
  uint4* ring_a = ring_z + BWD_STAGES * PIPE_U * 256;
  const int ni = blockIdx.y;
  const int v = p.c >> 3;
  const int cols = min(v, 256);
  const int lanes = 256 / cols;
  const int cx = threadIdx.x % cols, ly = threadIdx.x / cols;
  if (ly >= lanes) return;
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(hw, p0 + pix_per_chunk);
  const int g = ni / (p.n / p.groups);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* dzp = static_cast<const __nv_bfloat16*>(p.dz);
  const __nv_bfloat16* addp = static_cast<const __nv_bfloat16*>(p.add);
  __nv_bfloat16* dxp = static_cast<__nv_bfloat16*>(p.dx);
  for (int cb = 0; cb < v; cb += cols) {
    const int col = cb + cx;
    if (col >= v) continue;
    const int c8 = col * 8;
    float t1[8], t2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { t1[j] = 0.f; t2[j] = 0.f; }
    if (NORM) {
      ldf8(p.s1 + g * p.c + c8, t1);
      ldf8(p.s2 + g * p.c + c8, t2);
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1[j] *= p.inv_count; t2[j] *= p.inv_count; }
    }
    BwdConst8 k;
    bwd_const8<NORM>(p, ni, g, c8, k);
    const int step = lanes * PIPE_U;
    const int iters = (p1 - p0 - ly + step - 1) / step;
    auto issue = [&](int it) {
      const int so = (it % BWD_STAGES) * PIPE_U * 256;
#pragma unroll
      for (int u = 0; u < PIPE_U; ++u) {
        const int qx = p0 + ly + (it * PIPE_U + u) * lanes;
        const bool ok = qx < p1;
        const int qq = ok ? qx : p0;
        const int64_t pix = static_cast<int64_t>(ni) * hw + qq;
        cp_async16(ring_x + so + u * 256, xp + pix * p.c + c8, ok);
        cp_async16(ring_z + so + u * 256, dzp + bwd_dz_off(p, ni, qq, c8), ok);
        if (addp) cp_async16(ring_a + so + u * 256, addp + pix * p.c + c8, ok);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int it = 0; it < BWD_STAGES - 1; ++it) issue(it);
    for (int it = 0; it < iters; ++it) {
      issue(it + BWD_STAGES - 1);
      cp_async_wait<BWD_STAGES - 1>();
      const int so = (it % BWD_STAGES) * PIPE_U * 256;
#pragma unroll
      for (int u = 0; u < PIPE_U; ++u) {
        const int qx = p0 + ly + (it * PIPE_U + u) * lanes;
        if (qx >= p1) break;
        const int64_t pix = static_cast<int64_t>(ni) * hw + qx;
        float a[8], dz[8], dx[8];
        cvt8(ring_x[so + u * 256], a);
        cvt8(ring_z[so + u * 256], dz);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = a[j] * k.r[j] - k.mr[j];
          const float dy = dz[j] * al.d(xh * k.ga[j] + k.be[j]);
          dx[j] = NORM ? k.r[j] * (k.ga[j] * dy - t1[j] - xh * t2[j]) : dy;
        }
        if (addp) {
          float q[8];
          cvt8(ring_a[so + u * 256], q);
#pragma unroll
          for (int j = 0; j < 8; ++j) dx[j] += q[j];
        }
        stg16(dxp + pix * p.c + c8, pack8(dx));
      }
    }
    cp_async_wait<0>();
  }
}

// cp.async-ring variant of norm_act_bwd_reduce_v8_kernel<false> (RED_STAGES iterations of x and dz in flight per thread, no
// registers held while they fly): the plain-load kernel computes ~320 instructions per thread between two batches of
// loads with nothing in flight meanwhile (0.49 of the copy bandwidth standalone; reduce + finalize + apply of a 67 MB tensor 85.4 -> 79.6 us with the ring).
// GANB_BWD_REDUCE_RING=0 selects the plain-load kernel.
constexpr int RED_STAGES = 3;    // 2 tensors x 3 x 4 x 256 x 16 B = 96 KB per block
__global__ void __launch_bounds__(256, 2) norm_act_bwd_reduce_v8p_kernel(const NormActBwd p) {
  pdl_wait();
  const ActLin al(p.act);
  extern __shared__ uint4 ring_raw[];
  uint4* ring_x = ring_raw + threadIdx.x;
  uint4* ring_z = ring_x + RED_STAGES * PIPE_U * 256;
  const int ni = blockIdx.y, chunk = blockIdx.x;
  const int v = p.c >> 3;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  const int hw = p.h * p.w;
  const int p0 = chunk * p.pix_per_chunk, p1 = min(hw, p0 + p.pix_per_chunk);
  const int g = ni / (p.n / p.groups);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* dzp = static_cast<const __nv_bfloat16*>(p.dz);
  __shared__ float sh[256][17];
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float sa[8], sb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sa[j] = 0.f; sb[j] = 0.f; }
    if (col < v && ry < lanes) {
      const int c8 = col * 8;
      BwdConst8 k;
      bwd_const8<true>(p, ni, g, c8, k);
      const int step = lanes * PIPE_U;
      const int iters = (p1 - p0 - ry + step - 1) / step;
      auto issue = [&](int it) {
        const int so = (it % RED_STAGES) * PIPE_U * 256;
#pragma unroll
        for (int u = 0; u < PIPE_U; ++u) {
          const int qx = p0 + ry + (it * PIPE_U + u) * lanes;
          const bool ok = qx < p1;
          const int qq = ok ? qx : p0;
          cp_async16(ring_x + so + u * 256, xp + (static_cast<int64_t>(ni) * hw + qq) * p.c + c8, ok);
          cp_async16(ring_z + so + u * 256, dzp + bwd_dz_off(p, ni, qq, c8), ok);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int it = 0; it < RED_STAGES - 1; ++it) issue(it);
      for (int it = 0; it < iters; ++it) {
        issue(it + RED_STAGES - 1);
        cp_async_wait<RED_STAGES - 1>();
        const int so = (it % RED_STAGES) * PIPE_U * 256;
#pragma unroll
        for (int u = 0; u < PIPE_U; ++u) {
          // rows beyond p1 were zero-filled: dz = 0 contributes nothing to either sum
          float a[8], dz[8];
          cvt8(ring_x[so + u * 256], a);
          cvt8(ring_z[so + u * 256], dz);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = a[j] * k.r[j] - k.mr[j];
            const float dy = dz[j] * al.d(xh * k.ga[j] + k.be[j]);
            sa[j] += dy;
            sb[j] += dy * xh;
          }
        }
      }
      cp_async_wait<0>();
    }
    lanes_sum16(sa, sb, lanes, cols_per_pass, cx, ry, sh);
    if (ry == 0 && col < v) {
      float* out = p.part + (static_cast<int64_t>(ni) * p.chunks + chunk) * 2 * p.c;
      st4(out + col * 8, make_float4(sa[0], sa[1], sa[2], sa[3]));
      st4(out + col * 8 + 4, make_float4(sa[4], sa[5], sa[6], sa[7]));
      st4(out + p.c + col * 8, make_float4(sb[0], sb[1], sb[2], sb[3]));
      st4(out + p.c + col * 8 + 4, make_float4(sb[4], sb[5], sb[6], sb[7]));
    }
  }
}

// cp.async-ring variant (see bn_stats_partial_v8p_kernel): same access pattern, 3 iterations in flight
__global__ void __launch_bounds__(256)
colsum_partial_v8p_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int c, int rows_per_chunk,
                          float* __restrict__ partial) {
  pdl_wait();
  extern __shared__ uint4 ring_raw[];
  uint4* ring = ring_raw + threadIdx.x;
  const int chunk = blockIdx.x;
  const int v = c >> 3;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  const int64_t r0 = static_cast<int64_t>(chunk) * rows_per_chunk;
  const int64_t r1 = min(rows, r0 + rows_per_chunk);
  __shared__ float sh[256][17];
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float s[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; z[j] = 0.f; }
    if (col < v && ry < lanes) {
      const __nv_bfloat16* xc = x + col * 8;
      const int step = lanes * PIPE_U;
      const int iters = static_cast<int>((r1 - r0 - ry + step - 1) / step);
      auto issue = [&](int it) {
        uint4* dst = ring + (it % STATS_STAGES) * PIPE_U * 256;
#pragma unroll
        for (int u = 0; u < PIPE_U; ++u) {
          const int64_t r = r0 + ry + static_cast<int64_t>(it * PIPE_U + u) * lanes;
          const bool ok = r < r1;
          cp_async16(dst + u * 256, xc + (ok ? r : r0) * c, ok);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int it = 0; it < STATS_STAGES - 1; ++it) issue(it);
      for (int it = 0; it < iters; ++it) {
        issue(it + STATS_STAGES - 1);
        cp_async_wait<STATS_STAGES - 1>();
        const uint4* src = ring + (it % STATS_STAGES) * PIPE_U * 256;
#pragma unroll
        for (int u = 0; u < PIPE_U; ++u) {
          float a[8];
          cvt8(src[u * 256], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) s[j] += a[j];
        }
      }
      cp_async_wait<0>();
    }
    lanes_sum16(s, z, lanes, cols_per_pass, cx, ry, sh);
    if (ry == 0 && col < v) {
      st4(partial + static_cast<int64_t>(chunk) * c + col * 8, make_float4(s[0], s[1], s[2], s[3]));
      st4(partial + static_cast<int64_t>(chunk) * c + col * 8 + 4, make_float4(s[4], s[5], s[6], s[7]));
    }
  }
}

// ------------------------------------------------------------------------------------------------ pooling
// out[n,h/2,w/2,c] = mean of the 2x2 block (+ add[n,h/2,w/2,c])
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
meanpool2_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ add, TOut* __restrict__ out, int n, int ho,
                     int wo, int c) {
  pdl_wait();
  const int v = c >> 2;
  const int64_t total = static_cast<int64_t>(n) * ho * wo * v;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const int wi = static_cast<int>(pix % wo);
    const int hi = static_cast<int>((pix / wo) % ho);
    const int ni = static_cast<int>(pix / (static_cast<int64_t>(wo) * ho));
    const int64_t base = ((static_cast<int64_t>(ni) * 2 * ho + 2 * hi) * (2 * wo) + 2 * wi) * c + c4;
    const int64_t rowstride = static_cast<int64_t>(2 * wo) * c;
    // same summation order as tf.add_n([x[::2,::2], x[1::2,::2], x[::2,1::2], x[1::2,1::2]]) / 4
    const float4 a = ld4(x + base), b = ld4(x + base + rowstride), d = ld4(x + base + c), e = ld4(x + base + rowstride + c);
    float4 r = make_float4((a.x + b.x + d.x + e.x) * 0.25f, (a.y + b.y + d.y + e.y) * 0.25f,
                           (a.z + b.z + d.z + e.z) * 0.25f, (a.w + b.w + d.w + e.w) * 0.25f);
    if (add) {
      const float4 q = ld4(add + pix * c + c4);
      r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
    }
    st4(out + pix * c + c4, r);
  }
}

// dx[n,2h,2w,c] = scale * dz[n,h,w,c]   (scale = 0.25: mean-pool backward; scale = 1: nearest-upsample forward)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
expand2_kernel(const TIn* __restrict__ dz, TOut* __restrict__ dx, int n, int h, int w, int c, float scale) {
  pdl_wait();
  const int v = c >> 2;
  const int64_t total = static_cast<int64_t>(n) * h * w * v;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const int wi = static_cast<int>(pix % w);
    const int hi = static_cast<int>((pix / w) % h);
    const int ni = static_cast<int>(pix / (static_cast<int64_t>(w) * h));
    float4 g = ld4(dz + pix * c + c4);
    g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
    const int64_t base = ((static_cast<int64_t>(ni) * 2 * h + 2 * hi) * (2 * w) + 2 * wi) * c + c4;
    const int64_t rowstride = static_cast<int64_t>(2 * w) * c;
    st4(dx + base, g);
    st4(dx + base + c, g);
    st4(dx + base + rowstride, g);
    st4(dx + base + rowstride + c, g);
  }
}

// out[n,h,w,c] = scale * sum of the 2x2 block of x[n,2h,2w,c]  (nearest-upsample backward when scale = 1)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
sum2x2_kernel(const TIn* __restrict__ x, TOut* __restrict__ out, int n, int ho, int wo, int c, float scale) {
  pdl_wait();
  const int v = c >> 2;
  const int64_t total = static_cast<int64_t>(n) * ho * wo * v;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const int wi = static_cast<int>(pix % wo);
    const int hi = static_cast<int>((pix / wo) % ho);
    const int ni = static_cast<int>(pix / (static_cast<int64_t>(wo) * ho));
    const int64_t base = ((static_cast<int64_t>(ni) * 2 * ho + 2 * hi) * (2 * wo) + 2 * wi) * c + c4;
    const int64_t rowstride = static_cast<int64_t>(2 * wo) * c;
    const float4 a = ld4(x + base), b = ld4(x + base + rowstride), d = ld4(x + base + c), e = ld4(x + base + rowstride + c);
    st4(out + pix * c + c4, make_float4((a.x + b.x + d.x + e.x) * scale, (a.y + b.y + d.y + e.y) * scale,
                                         (a.z + b.z + d.z + e.z) * scale, (a.w + b.w + d.w + e.w) * scale));
  }
}

// out[n, oh, ow, c]: x[n, i, j, c] at (i*stride, j*stride), zero elsewhere (one pass writes the whole output)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
dilate2d_kernel(const TIn* __restrict__ x, TOut* __restrict__ out, int n, int h, int w, int c, int stride, int oh, int ow) {
  pdl_wait();
  const int v = c >> 2;
  const int64_t total = static_cast<int64_t>(n) * oh * ow * v;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const int wi = static_cast<int>(pix % ow);
    const int hi = static_cast<int>((pix / ow) % oh);
    const int ni = static_cast<int>(pix / (static_cast<int64_t>(ow) * oh));
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hi % stride == 0 && wi % stride == 0 && hi / stride < h && wi / stride < w)
      g = ld4(x + ((static_cast<int64_t>(ni) * h + hi / stride) * w + wi / stride) * c + c4);
    st4(out + pix * c + c4, g);
  }
}

// ---- scalar variants for channel counts that are not a multiple of 4 (RGB images)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
pool2_scalar_kernel(const TIn* __restrict__ x, const float* __restrict__ add, TOut* __restrict__ out, int n, int ho,
                    int wo, int c, float scale) {
  pdl_wait();
  const int64_t total = static_cast<int64_t>(n) * ho * wo * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const int64_t pix = i / c;
    const int wi = static_cast<int>(pix % wo);
    const int hi = static_cast<int>((pix / wo) % ho);
    const int ni = static_cast<int>(pix / (static_cast<int64_t>(wo) * ho));
    const int64_t base = ((static_cast<int64_t>(ni) * 2 * ho + 2 * hi) * (2 * wo) + 2 * wi) * c + ch;
    const int64_t rowstride = static_cast<int64_t>(2 * wo) * c;
    float r = (static_cast<float>(x[base]) + static_cast<float>(x[base + rowstride]) + static_cast<float>(x[base + c]) +
               static_cast<float>(x[base + rowstride + c])) * scale;
    if (add) r += add[i];
    out[i] = static_cast<TOut>(r);
  }
}
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
expand2_scalar_kernel(const TIn* __restrict__ dz, TOut* __restrict__ dx, int n, int h, int w, int c, float scale) {
  pdl_wait();
  const int64_t total = static_cast<int64_t>(n) * h * w * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const int64_t pix = i / c;
    const int wi = static_cast<int>(pix % w);
    const int hi = static_cast<int>((pix / w) % h);
    const int ni = static_cast<int>(pix / (static_cast<int64_t>(w) * h));
    const TOut g = static_cast<TOut>(static_cast<float>(dz[i]) * scale);
    const int64_t base = ((static_cast<int64_t>(ni) * 2 * h + 2 * hi) * (2 * w) + 2 * wi) * c + ch;
    const int64_t rowstride = static_cast<int64_t>(2 * w) * c;
    dx[base] = g; dx[base + c] = g; dx[base + rowstride] = g; dx[base + rowstride + c] = g;
  }
}

// ------------------------------------------------------------------------------------------------ casts / axpy
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) cast_kernel(const TIn* __restrict__ x, TOut* __restrict__ y, int64_t n4, float scale) {
  pdl_wait();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 a = ld4(x + i * 4);
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    st4(y + i * 4, a);
  }
}
template <typename TIn, typename TOut>
__global__ void cast_tail_kernel(const TIn* __restrict__ x, TOut* __restrict__ y, int64_t begin, int64_t n, float scale) {
  pdl_wait();
  const int64_t i = begin + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) y[i] = static_cast<TOut>(static_cast<float>(x[i]) * scale);
}

// y = a*x + b*y  (fp32, flat)
__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float a, float b) {
  pdl_wait();
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 u = ld4(x + i * 4);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b != 0.f) w = ld4(y + i * 4);     // b == 0: y may be uninitialised (NaN * 0 would poison the result)
    w.x = a * u.x + b * w.x; w.y = a * u.y + b * w.y; w.z = a * u.z + b * w.z; w.w = a * u.w + b * w.w;
    st4(y + i * 4, w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    y[i] = a * x[i] + (b != 0.f ? b * y[i] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[c] = beta*out[c] + sum_rows x[row][c].  Every block sums a row chunk into partial[chunk][c]; the last block to
// finish adds the chunk partials in chunk order (deterministic) -- one launch instead of partial + finalize.
template <typename TIn>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const TIn* __restrict__ x, int64_t rows, int c, int rows_per_chunk, float* __restrict__ partial) {
  pdl_wait();
  const int chunk = blockIdx.x;
  const int v = c >> 2;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  const int64_t r0 = static_cast<int64_t>(chunk) * rows_per_chunk;
  const int64_t r1 = min(rows, r0 + rows_per_chunk);
  __shared__ float4 sh[256];
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float4 s = make_float4(0, 0, 0, 0);
    if (col < v && ry < lanes) {
      constexpr int U = 4;
      for (int64_t r = r0 + ry; r < r1; r += lanes * U) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          a[u] = (r + u * lanes < r1) ? ld4(x + (r + u * lanes) * c + col * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) { s.x += a[u].x; s.y += a[u].y; s.z += a[u].z; s.w += a[u].w; }
      }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    if (ry == 0 && col < v) {
      for (int l = 1; l < lanes; ++l) {
        const float4 a = sh[l * cols_per_pass + cx];
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      }
      st4(partial + static_cast<int64_t>(chunk) * c + col * 4, s);
    }
    __syncthreads();
  }
}
// One block per 32 channels: 32 thread rows split the chunk partials (coalesced 128-byte rows, independent loads) and meet
// in shared memory in a fixed order (deterministic); see bn_stats_finalize_kernel.
__global__ void __launch_bounds__(1024)
colsum_finalize_kernel(const float* __restrict__ partial, int c, int chunks, float beta, float* __restrict__ out) {
  pdl_wait();
  __shared__ float ssum[32][33];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (ch < c) {
#pragma unroll 8
    for (int k = ly; k < chunks; k += 32) s += __ldg(partial + static_cast<int64_t>(k) * c + ch);
  }
  ssum[ly][cx] = s;
  __syncthreads();
  const int chw = blockIdx.x * 32 + ly;
  const float a = warp_sum_f(ssum[cx][ly]);
  if (cx == 0 && chw < c) out[chw] = (beta != 0.f ? beta * out[chw] : 0.f) + a;
}


// Narrow tensors (c <= 8, e.g. the RGB bias or a scalar head): a thread walks rows and keeps all c sums in registers.
template <typename TIn>
__global__ void __launch_bounds__(256)
colsum_narrow_kernel(const TIn* __restrict__ x, int64_t rows, int c, int rows_per_chunk, float* __restrict__ partial) {
  pdl_wait();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_chunk;
  const int64_t r1 = min(rows, r0 + rows_per_chunk);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < c) s[j] += static_cast<float>(x[r * c + j]);
  }
  __shared__ float shn[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t = warp_sum_f(s[j]);
    if (lane == 0) shn[warp][j] = t;
  }
  __syncthreads();
  if (threadIdx.x < c) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += shn[w][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.x) * c + threadIdx.x] = t;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(256)
colsum_wide_kernel(const TIn* __restrict__ x, int64_t rows, int c, float beta, float* __restrict__ out) {
  pdl_wait();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int64_t r = 0; r < rows; ++r) s += static_cast<float>(x[r * c + ch]);
  out[ch] = (beta != 0.f ? beta * out[ch] : 0.f) + s;
}

// any other channel count: one block per channel
template <typename TIn>
__global__ void __launch_bounds__(256)
colsum_scalar_kernel(const TIn* __restrict__ x, int64_t rows, int c, float beta, float* __restrict__ out) {
  pdl_wait();
  const int ch = blockIdx.x;
  __shared__ float sh[256];
  float s = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) s += static_cast<float>(x[r * c + ch]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[ch] = (beta != 0.f ? beta * out[ch] : 0.f) + sh[0];
}

// ------------------------------------------------------------------------------------------------ label map
// out[n, hw, coff + j] = e[n, j] (raw and, optionally, activated) : tf.tile + tf.concat of the embedded label
__global__ void __launch_bounds__(256)
bcast_channels_kernel(const float* __restrict__ e, int n, int hw, int c2, int coff, int cstride, int act,
                      __nv_bfloat16* __restrict__ out_raw, __nv_bfloat16* __restrict__ out_act) {
  pdl_wait();
  const int v = c2 >> 2;
  const int64_t total = static_cast<int64_t>(n) * hw * v;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const int ni = static_cast<int>(pix / hw);
    const float4 a = ld4(e + static_cast<int64_t>(ni) * c2 + c4);
    if (out_raw) st4(out_raw + pix * cstride + coff + c4, a);
    if (out_act)
      st4(out_act + pix * cstride + coff + c4,
          make_float4(act_f(a.x, act), act_f(a.y, act), act_f(a.z, act), act_f(a.w, act)));
  }
}
// de[n, j] = sum_hw ( d_raw[n,hw,coff+j] + dact(e[n,j]) * d_act[n,hw,coff+j] );  grid (n), block = c2/4 x lanes
template <typename TG>
__global__ void __launch_bounds__(256)
bcast_channels_bwd_kernel(const float* __restrict__ e, int hw, int c2, int coff, int cstride, int act,
                          const TG* __restrict__ d_raw, const TG* __restrict__ d_act,
                          float* __restrict__ de) {
  pdl_wait();
  const int ni = blockIdx.x;
  const int v = c2 >> 2;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  __shared__ float4 sh[256];
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float4 s = make_float4(0, 0, 0, 0);
    if (col < v && ry < lanes) {
      const float4 ev = ld4(e + static_cast<int64_t>(ni) * c2 + col * 4);
      const float4 m = make_float4(dact_f(ev.x, act), dact_f(ev.y, act), dact_f(ev.z, act), dact_f(ev.w, act));
      for (int px = ry; px < hw; px += lanes) {
        const int64_t o = (static_cast<int64_t>(ni) * hw + px) * cstride + coff + col * 4;
        if (d_raw) { const float4 a = ld4(d_raw + o); s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
        if (d_act) { const float4 a = ld4(d_act + o); s.x += m.x * a.x; s.y += m.y * a.y; s.z += m.z * a.z; s.w += m.w * a.w; }
      }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    if (ry == 0 && col < v) {
      for (int l = 1; l < lanes; ++l) {
        const float4 a = sh[l * cols_per_pass + cx];
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      }
      st4(de + static_cast<int64_t>(ni) * c2 + col * 4, s);
    }
    __syncthreads();
  }
}

// dx[n,hw,0:c1] = d_raw[..., 0:c1] + dact(x) * d_act[..., 0:c1]   (wide bf16 gradients -> fp32 dx)
template <typename TG, typename TOut>
__global__ void __launch_bounds__(256)
concat_bwd_x_kernel(const float* __restrict__ x, int64_t pixels, int c1, int cstride, int act,
                    const TG* __restrict__ d_raw, const TG* __restrict__ d_act,
                    TOut* __restrict__ dx) {
  pdl_wait();
  const int v = c1 >> 2;
  const int64_t total = pixels * v;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const float4 a = ld4(x + pix * c1 + c4);
    float4 s = make_float4(0, 0, 0, 0);
    if (d_raw) s = ld4(d_raw + pix * cstride + c4);
    if (d_act) {
      const float4 g = ld4(d_act + pix * cstride + c4);
      s.x += dact_f(a.x, act) * g.x; s.y += dact_f(a.y, act) * g.y;
      s.z += dact_f(a.z, act) * g.z; s.w += dact_f(a.w, act) * g.w;
    }
    st4(dx + pix * c1 + c4, s);
  }
}

// ------------------------------------------------------------------------------------------------ global pool
// out[n,c] = mean_hw act(x[n,hw,c]) ; grid (n)
__global__ void __launch_bounds__(256)
act_mean_hw_fwd_kernel(const float* __restrict__ x, int hw, int c, int act, float* __restrict__ out) {
  pdl_wait();
  const int ni = blockIdx.x;
  const int v = c >> 2;
  const int lanes = max(1, 256 / min(v, 256));
  const int cols_per_pass = 256 / lanes;
  const int cx = threadIdx.x % cols_per_pass, ry = threadIdx.x / cols_per_pass;
  __shared__ float4 sh[256];
  const float inv = 1.0f / hw;
  for (int cb = 0; cb < v; cb += cols_per_pass) {
    const int col = cb + cx;
    float4 s = make_float4(0, 0, 0, 0);
    if (col < v && ry < lanes)
      for (int px = ry; px < hw; px += lanes) {
        const float4 a = ld4(x + (static_cast<int64_t>(ni) * hw + px) * c + col * 4);
        s.x += act_f(a.x, act); s.y += act_f(a.y, act); s.z += act_f(a.z, act); s.w += act_f(a.w, act);
      }
    sh[threadIdx.x] = s;
    __syncthreads();
    if (ry == 0 && col < v) {
      for (int l = 1; l < lanes; ++l) {
        const float4 a = sh[l * cols_per_pass + cx];
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      }
      st4(out + static_cast<int64_t>(ni) * c + col * 4, make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv));
    }
    __syncthreads();
  }
}
// dx[n,hw,c] = dact(x) * dout[n,c] / hw
template <typename TOut>
__global__ void __launch_bounds__(256)
act_mean_hw_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, int n, int hw, int c, int act,
                       TOut* __restrict__ dx) {
  pdl_wait();
  const int v = c >> 2;
  const int64_t total = static_cast<int64_t>(n) * hw * v;
  const float inv = 1.0f / hw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % v) * 4;
    const int64_t pix = i / v;
    const int ni = static_cast<int>(pix / hw);
    const float4 a = ld4(x + pix * c + c4);
    const float4 g = ld4(dout + static_cast<int64_t>(ni) * c + c4);
    st4(dx + pix * c + c4, make_float4(dact_f(a.x, act) * g.x * inv, dact_f(a.y, act) * g.y * inv,
                                        dact_f(a.z, act) * g.z * inv, dact_f(a.w, act) * g.w * inv));
  }
}

// ------------------------------------------------------------------------------------------------ losses
// mode 0: hinge D loss = mean(relu(1 - d[0:n_real])) + mean(relu(1 + d[n_real:n]))   (gan_cifar_resnet.py:376-378)
// mode 1: G loss = -mean(d[0:n])                                                      (gan_cifar_resnet.py:492)
// loss_out[0] (+)= scale * loss ; dlogits = scale * dloss/dd
// mode = 2 * loss_type + side; side 0 = discriminator loss over d = [real | fake], side 1 = generator loss over the
// fake logits; loss_type 0 HINGE, 1 WGAN / WGAN-GP, 2 LSGAN, 3 CGAN, 4 Modified_MiniMax, 5 MiniMax
// (common/misc.py:310-394).  l = this element's contribution before the 1/count of its mean, g = dl/dv.
__device__ __forceinline__ void gan_loss_elem(int type, int side, bool real, float v, float& l, float& g) {
  const float sg = 1.f / (1.f + __expf(-v));                      // sigmoid(v)
  if (side == 0) {
    switch (type) {
      case 0: { const float t = real ? 1.f - v : 1.f + v; l = t > 0.f ? t : 0.f; g = t > 0.f ? (real ? -1.f : 1.f) : 0.f; } break;
      case 1: l = real ? -v : v; g = real ? -1.f : 1.f; break;
      case 2: if (real) { l = 0.5f * (1.f - v) * (1.f - v); g = -(1.f - v); } else { l = 0.5f * v * v; g = v; } break;
      case 3: {   // tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))
        const float sp = log1pf(__expf(-fabsf(v)));
        if (real) { l = fmaxf(-v, 0.f) + sp; g = sg - 1.f; } else { l = fmaxf(v, 0.f) + sp; g = sg; }
      } break;
      default: if (real) { l = -__logf(sg); g = sg - 1.f; } else { l = -__logf(1.f - sg); g = sg; } break;
    }
  } else {
    switch (type) {
      case 0: case 1: l = -v; g = -1.f; break;
      case 2: l = 0.5f * (1.f - v) * (1.f - v); g = -(1.f - v); break;
      case 3: l = fmaxf(-v, 0.f) + log1pf(__expf(-fabsf(v))); g = sg - 1.f; break;
      case 4: l = -__logf(sg); g = sg - 1.f; break;
      default: l = __logf(1.f - sg); g = -sg; break;
    }
  }
}

__global__ void __launch_bounds__(256)
gan_loss_kernel(const float* __restrict__ d, int n, int n_real, int mode, float scale, int accumulate,
                float* __restrict__ loss_out, float* __restrict__ dlogits) {
  pdl_wait();
  __shared__ float sh[256];
  float acc = 0.f;
  const int type = mode >> 1, side = mode & 1;
  const int n_fake = n - n_real;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const bool real = side == 0 && i < n_real;
    float l, g;
    gan_loss_elem(type, side, real, d[i], l, g);
    const float w = 1.f / (side == 1 ? n : (real ? n_real : n_fake));   // LSGAN's 1/2 is already in l
    acc += l * w;
    dlogits[i] = g * w * scale;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (accumulate ? loss_out[0] : 0.f) + scale * sh[0];
}

// tf.reduce_mean(tf.nn.sparse_softmax_cross_entropy_with_logits(logits, labels)) (ACGAN/train.py:110-121):
// loss_out[0] (+)= scale * mean_i(logsumexp(z_i) - z_i[label_i]); dlogits = scale * (softmax(z_i) - onehot) / n.
__global__ void __launch_bounds__(256)
softmax_xent_kernel(const float* __restrict__ z, const int* __restrict__ labels, int n, int c, float scale,
                    int accumulate, float* __restrict__ loss_out, float* __restrict__ dlogits) {
  pdl_wait();
  __shared__ float sh[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* row = z + static_cast<int64_t>(i) * c;
    float m = row[0];
    for (int j = 1; j < c; ++j) m = fmaxf(m, row[j]);
    float se = 0.f;
    for (int j = 0; j < c; ++j) se += expf(row[j] - m);
    const float lse = m + logf(se);
    const int lab = labels[i];
    acc += (lse - row[lab]) / n;
    for (int j = 0; j < c; ++j)
      dlogits[static_cast<int64_t>(i) * c + j] = scale * (expf(row[j] - lse) - (j == lab ? 1.f : 0.f)) / n;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (accumulate ? loss_out[0] : 0.f) + scale * sh[0];
}

// ------------------------------------------------------------------------------------------------ Adam
// tf.train.AdamOptimizer: theta -= lr_t * m / (sqrt(v) + eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t) given as a device
// scalar.  One flat buffer per network = "multi-tensor".
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, const float* __restrict__ lr_t, float b1, float b2, float eps, float grad_scale) {
  pdl_wait();
  const float lr = __ldg(lr_t);
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pp = ld4(p + i * 4), gg = ld4(g + i * 4), mm = ld4(m + i * 4), vv = ld4(v + i * 4);
    gg.x *= grad_scale; gg.y *= grad_scale; gg.z *= grad_scale; gg.w *= grad_scale;
    mm.x = b1 * mm.x + (1.f - b1) * gg.x; mm.y = b1 * mm.y + (1.f - b1) * gg.y;
    mm.z = b1 * mm.z + (1.f - b1) * gg.z; mm.w = b1 * mm.w + (1.f - b1) * gg.w;
    vv.x = b2 * vv.x + (1.f - b2) * gg.x * gg.x; vv.y = b2 * vv.y + (1.f - b2) * gg.y * gg.y;
    vv.z = b2 * vv.z + (1.f - b2) * gg.z * gg.z; vv.w = b2 * vv.w + (1.f - b2) * gg.w * gg.w;
    pp.x -= lr * mm.x / (sqrtf(vv.x) + eps); pp.y -= lr * mm.y / (sqrtf(vv.y) + eps);
    pp.z -= lr * mm.z / (sqrtf(vv.z) + eps); pp.w -= lr * mm.w / (sqrtf(vv.w) + eps);
    st4(p + i * 4, pp); st4(m + i * 4, mm); st4(v + i * 4, vv);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float gg = g[i] * grad_scale;
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm; v[i] = vv;
    p[i] -= lr * mm / (sqrtf(vv) + eps);
  }
}

// ------------------------------------------------------------------------------------------------ input edge
// gan_cifar_resnet.py:334-337: int32 [B, 3*H*W] CHW -> 2*(x/256 - .5) + noise -> NHWC fp32
__global__ void __launch_bounds__(256)
preprocess_real_kernel(const int* __restrict__ data, const float* __restrict__ noise, int b, int hw, float* __restrict__ out) {
  pdl_wait();
  const int64_t total = static_cast<int64_t>(b) * hw * 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % 3);
    const int64_t pix = i / 3;
    const int px = static_cast<int>(pix % hw);
    const int64_t ni = pix / hw;
    const int64_t src = (ni * 3 + ch) * hw + px;
    float v = 2.f * (static_cast<float>(data[src]) / 256.f - 0.5f);
    if (noise) v += noise[src];
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ embedding
__global__ void embedding_fwd_kernel(const float* __restrict__ table, const int* __restrict__ labels, int n, int dim,
                                     float* __restrict__ out) {
  pdl_wait();
  const int64_t total = static_cast<int64_t>(n) * dim;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ni = static_cast<int>(i / dim), j = static_cast<int>(i % dim);
    out[i] = table[static_cast<int64_t>(labels[ni]) * dim + j];
  }
}
// dtable[row, j] += sum_{n: labels[n]==row} dout[n, j]   (IndexedSlices summed by index; deterministic)
__global__ void embedding_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ labels, int n, int dim,
                                     int vocab, float* __restrict__ dtable) {
  pdl_wait();
  const int64_t total = static_cast<int64_t>(vocab) * dim;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / dim), j = static_cast<int>(i % dim);
    float s = 0.f;
    for (int ni = 0; ni < n; ++ni)
      if (labels[ni] == row) s += dout[static_cast<int64_t>(ni) * dim + j];
    dtable[i] += s;
  }
}

}  // namespace ganb

using namespace ganb;
#define STREAM static_cast<cudaStream_t>(stream)

// ================================================================================================ C ABI
extern "C" int64_t ganb_bn_stats_workspace(int n, int hw, int c, int groups) {
  (void)n; (void)hw;
  return static_cast<int64_t>(groups) * 1024 * 2 * c * 4;
}

// Four blocks per SM are needed to keep HBM busy (two measured 80 % slower); small tensors get at least 128 rows per
// chunk (the finalize kernels are latency-bound on the number of partials).
// blocks per SM of the streaming kernels (tuning hooks for tests/probe_bw_kernels.py; read once)
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
// 2 long-running blocks per SM: 23.2 -> 16.4 us for the statistics of a 67 MB tensor, 21.4 -> 16.6 us for its column sums
// (4 / 8 blocks per SM spend their time in the pipeline prologue and the shared-memory reduction; profiles/r02_bandwidth_kernels.txt)
static int stats_bps() { static const int v = env_int("GANB_STATS_BPS", 2); return v; }
static int fwd_bps() { static const int v = env_int("GANB_V8_FWD_BPS", 3); return v; }
static int bwd_bps() { static const int v = env_int("GANB_V8_BWD_BPS", 2); return v; }
static bool reduce_ring() { static const int v = env_int("GANB_BWD_REDUCE_RING", 1); return v != 0; }
static int colsum_bps() { static const int v = env_int("GANB_COLSUM_BPS", 2); return v; }

static int stats_chunks(int rows_per_group, int groups) {
  int chunks = ceil_div(stats_bps() * sm_count(), groups);
  const int max_chunks = ceil_div(rows_per_group, 128);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 1024) chunks = 1024;
  if (chunks < 1) chunks = 1;
  return chunks;
}

extern "C" int ganb_bn_stats(const void* x, int x_dtype, int n, int hw, int c, int groups, float eps, float* mean,
                             float* rstd, void* workspace, void* stream) {
  if (!x || !mean || !rstd || !workspace) return fail(GANB_E_BADARG, "bn_stats: null buffer");
  if (c % 4 != 0) return fail(GANB_E_UNSUPPORTED, "bn_stats: c=%d must be a multiple of 4", c);
  if (groups <= 0 || n % groups != 0) return fail(GANB_E_BADARG, "bn_stats: n=%d not divisible by groups=%d", n, groups);
  const int rows_per_group = (n / groups) * hw;
  const int chunks = stats_chunks(rows_per_group, groups);
  const int rows_per_chunk = ceil_div(rows_per_group, chunks);
  const int used = ceil_div(rows_per_group, rows_per_chunk);
  const dim3 grid(used, groups);
  if (x_dtype == GANB_BF16 && c % 8 == 0) {
    constexpr int smem = STATS_STAGES * PIPE_U * 256 * 16;
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(bn_stats_partial_v8p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      configured = true;
    }
    launch_k(bn_stats_partial_v8p_kernel, grid, 256, smem, STREAM, static_cast<const __nv_bfloat16*>(x), rows_per_group, c,
             used, rows_per_chunk, static_cast<float*>(workspace));
  }
  else if (x_dtype == GANB_BF16)
    launch_k(bn_stats_partial_kernel<__nv_bfloat16>, grid, 256, 0, STREAM, static_cast<const __nv_bfloat16*>(x), rows_per_group, c,
                                                                      used, rows_per_chunk, static_cast<float*>(workspace));
  else
    launch_k(bn_stats_partial_kernel<float>, grid, 256, 0, STREAM, static_cast<const float*>(x), rows_per_group, c, used,
                                                              rows_per_chunk, static_cast<float*>(workspace));
  GANB_CHECK_LAUNCH("bn_stats_partial_kernel");
  launch_k(bn_stats_finalize_kernel, dim3(ceil_div(c, 32), groups), 1024, 0, STREAM, static_cast<float*>(workspace), c, groups, used, 1.0f / rows_per_group, eps, mean, rstd);
  GANB_CHECK_LAUNCH("bn_stats_finalize_kernel");
  return 0;
}

// mean / rstd from per-tile column sums produced by a convolution epilogue (ganb_conv2d_igemm_stats,
// ganb_upconv_fprop_stats): partial [groups][chunks][{sum, sumsq}][c], count = elements per channel and tower.
extern "C" int ganb_bn_stats_finalize(const float* partial, int c, int groups, int chunks, int64_t count, float eps,
                                      float* mean, float* rstd, void* stream) {
  if (!partial || !mean || !rstd) return fail(GANB_E_BADARG, "bn_stats_finalize: null buffer");
  if (c <= 0 || groups <= 0 || chunks <= 0 || count <= 0) return fail(GANB_E_BADARG, "bn_stats_finalize: bad shape");
  launch_k(bn_stats_finalize_kernel, dim3(ceil_div(c, 32), groups), 1024, 0, STREAM, partial, c, groups, chunks,
           1.0f / static_cast<float>(count), eps, mean, rstd);
  GANB_CHECK_LAUNCH("bn_stats_finalize_kernel");
  return 0;
}

static int bwd_chunks(int n, int hw) {
  int chunks = ceil_div(8 * sm_count(), n);
  const int max_chunks = ceil_div(hw, 16);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  return chunks;
}

// The 8-channel kernels are launched as ONE wave of long-running blocks: floor(resident blocks / n) pixel chunks per
// sample (>= 64 pixels each).  Many short blocks spent their time in the per-block prologue (constants, label lookup)
// and epilogue (shared-memory reduction) -- the reduce kernel ran at 44 % of its HBM time.  Always <= bwd_chunks().
static int v8_chunks(int n, int hw, int blocks_per_sm) {
  int chunks = (sm_count() * blocks_per_sm) / (n > 0 ? n : 1);
  const int max_chunks = hw / 64;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int cap = bwd_chunks(n, hw);
  return chunks < cap ? chunks : cap;
}

extern "C" int ganb_norm_act_fwd(const void* x, int x_dtype, int n, int h, int w, int c, const float* mean, const float* rstd,
                                 int groups, const float* gamma, const float* beta, const int* labels, int act,
                                 int upsample, void* out, int out_dtype, int out_cstride, void* out_raw_bf16,
                                 int raw_cstride, void* stream) {
  if (!x || !out) return fail(GANB_E_BADARG, "norm_act_fwd: null buffer");
  if (c % 4 != 0 || out_cstride % 4 != 0) return fail(GANB_E_UNSUPPORTED, "norm_act_fwd: c=%d, stride=%d must be multiples of 4", c, out_cstride);
  if (groups <= 0 || n % groups != 0) return fail(GANB_E_BADARG, "norm_act_fwd: bad groups");
  NormActFwd p;
  p.x = x; p.x_bf16 = (x_dtype == GANB_BF16); p.n = n; p.h = h; p.w = w; p.c = c;
  p.mean = mean; p.rstd = rstd; p.groups = groups;
  p.gamma = gamma; p.beta = beta; p.labels = labels;
  p.act = act; p.quad = (upsample == 2); p.upsample = p.quad ? 0 : upsample;
  upsample = p.upsample;
  p.out = out; p.out_bf16 = (out_dtype == GANB_BF16); p.out_cstride = out_cstride > 0 ? out_cstride : c;
  p.out_raw = static_cast<__nv_bfloat16*>(out_raw_bf16); p.raw_cstride = raw_cstride > 0 ? raw_cstride : c;
  const bool v8 = p.x_bf16 && p.out_bf16 && !p.out_raw && c % 8 == 0 && p.out_cstride % 8 == 0 && act != GANB_ACT_TANH;
  if (p.quad && (!v8 || (h & 1) || (w & 1)))
    return fail(GANB_E_UNSUPPORTED, "norm_act_fwd: quad-layout input needs bf16 in/out, c %% 8 == 0, even h, w and no raw copy");
  const int chunks = v8 ? v8_chunks(n, h * w, fwd_bps()) : bwd_chunks(n, h * w);
  const int ppc = ceil_div(h * w, chunks);
  if (v8 && upsample) launch_k(norm_act_fwd_v8_kernel<true>, dim3(ceil_div(h * w, ppc), n), 256, 0, STREAM, p, ppc);
  else if (v8) launch_k(norm_act_fwd_v8_kernel<false>, dim3(ceil_div(h * w, ppc), n), 256, 0, STREAM, p, ppc);
  else if (p.x_bf16) launch_k(norm_act_fwd_kernel<__nv_bfloat16>, dim3(ceil_div(h * w, ppc), n), 256, 0, STREAM, p, ppc);
  else launch_k(norm_act_fwd_kernel<float>, dim3(ceil_div(h * w, ppc), n), 256, 0, STREAM, p, ppc);
  GANB_CHECK_LAUNCH("norm_act_fwd_kernel");
  return 0;
}

extern "C" int64_t ganb_norm_act_bwd_workspace(int n, int hw, int c, int groups) {
  const int chunks = bwd_chunks(n, hw);
  return (static_cast<int64_t>(n) * chunks * 2 * c + 2LL * n * c + 2LL * groups * c) * 4;
}

namespace ganb {
template <typename TX>
static void launch_bwd_reduce(const NormActBwd& p, dim3 grid, cudaStream_t s) {
  if (p.upsample && p.dz_bf16) launch_k(norm_act_bwd_reduce_kernel<true, true, TX>, grid, 256, 0, s, p);
  else if (p.upsample) launch_k(norm_act_bwd_reduce_kernel<true, false, TX>, grid, 256, 0, s, p);
  else if (p.dz_bf16) launch_k(norm_act_bwd_reduce_kernel<false, true, TX>, grid, 256, 0, s, p);
  else launch_k(norm_act_bwd_reduce_kernel<false, false, TX>, grid, 256, 0, s, p);
}
template <typename TX>
static void launch_bwd_apply(const NormActBwd& p, bool norm, dim3 grid, int ppc, cudaStream_t s) {
  const int idx = (norm ? 4 : 0) | (p.upsample ? 2 : 0) | (p.dz_bf16 ? 1 : 0);
  switch (idx) {
    case 0: launch_k(norm_act_bwd_apply_kernel<false, false, false, TX>, grid, 256, 0, s, p, ppc); break;
    case 1: launch_k(norm_act_bwd_apply_kernel<false, false, true, TX>, grid, 256, 0, s, p, ppc); break;
    case 2: launch_k(norm_act_bwd_apply_kernel<false, true, false, TX>, grid, 256, 0, s, p, ppc); break;
    case 3: launch_k(norm_act_bwd_apply_kernel<false, true, true, TX>, grid, 256, 0, s, p, ppc); break;
    case 4: launch_k(norm_act_bwd_apply_kernel<true, false, false, TX>, grid, 256, 0, s, p, ppc); break;
    case 5: launch_k(norm_act_bwd_apply_kernel<true, false, true, TX>, grid, 256, 0, s, p, ppc); break;
    case 6: launch_k(norm_act_bwd_apply_kernel<true, true, false, TX>, grid, 256, 0, s, p, ppc); break;
    default: launch_k(norm_act_bwd_apply_kernel<true, true, true, TX>, grid, 256, 0, s, p, ppc); break;
  }
}
}  // namespace ganb

// phase 0: everything; phase 1: per-sample / per-group sums only (reduce + finalize + dgamma/dbeta scatter), leaving
// [s1 | s2] in the workspace at ganb_norm_act_bwd_sums_offset() for a cross-GPU all-reduce; phase 2: apply only, reading
// the (all-reduced) sums back, with 1/count scaled by count_scale = 1/world.
static int norm_act_bwd_impl(const void* x, int x_dtype, const void* dz, int dz_dtype, int dz_cstride, int n, int h,
                             int w, int c, const float* mean, const float* rstd, int groups, const float* gamma,
                             const float* beta, const int* labels, int n_rows, int act, int upsample,
                             float* dgamma, float* dbeta, const void* add, int add_dtype, void* dx, int dx_dtype,
                             void* workspace, int phase, float count_scale, void* stream) {
  if (!x || !dz || !dx) return fail(GANB_E_BADARG, "norm_act_bwd: null buffer");
  if (c % 4 != 0) return fail(GANB_E_UNSUPPORTED, "norm_act_bwd: c=%d must be a multiple of 4", c);
  if (groups <= 0 || n % groups != 0) return fail(GANB_E_BADARG, "norm_act_bwd: bad groups");
  NormActBwd p;
  p.x = x; p.x_bf16 = (x_dtype == GANB_BF16);
  p.dz = dz; p.dz_bf16 = (dz_dtype == GANB_BF16); p.dz_cstride = dz_cstride > 0 ? dz_cstride : c;
  p.n = n; p.h = h; p.w = w; p.c = c;
  p.mean = mean; p.rstd = rstd; p.groups = groups;
  p.gamma = gamma; p.beta = beta; p.labels = labels;
  p.act = act; p.quad = (upsample == 2); p.upsample = p.quad ? 0 : upsample;
  upsample = p.upsample;
  p.add = add; p.add_bf16 = (add_dtype == GANB_BF16); p.dx = dx; p.dx_bf16 = (dx_dtype == GANB_BF16);
  p.part = nullptr; p.s1 = nullptr; p.s2 = nullptr; p.chunks = 0; p.pix_per_chunk = 0; p.inv_count = 0.f;
  // all-bf16 fast path: 8 channels per thread
  const bool v8 = p.x_bf16 && p.dz_bf16 && p.dx_bf16 && (!add || p.add_bf16) && c % 8 == 0 && p.dz_cstride % 8 == 0 &&
                  act != GANB_ACT_TANH;
  if (p.quad && (!v8 || add || (h & 1) || (w & 1)))
    return fail(GANB_E_UNSUPPORTED, "norm_act_bwd: quad-layout input needs the all-bf16 path, even h, w and no added gradient");
  if (mean) {
    if (!workspace) return fail(GANB_E_BADARG, "norm_act_bwd: workspace required with normalisation");
    const int hw = h * w;
    const int chunks = v8 ? v8_chunks(n, hw, bwd_bps()) : bwd_chunks(n, hw);
    p.pix_per_chunk = ceil_div(hw, chunks);
    p.chunks = ceil_div(hw, p.pix_per_chunk);
    p.part = static_cast<float*>(workspace);
    // fixed layout (independent of the chunk count actually used): [n][bwd_chunks][2][c] | [n][2][c] | s1 | s2
    float* sums = p.part + static_cast<int64_t>(n) * bwd_chunks(n, hw) * 2 * c;
    float* s1 = sums + 2LL * n * c;
    float* s2 = s1 + static_cast<int64_t>(groups) * c;
    if (phase != 2) {
      const dim3 grid(p.chunks, n);
      if (v8 && upsample) launch_k(norm_act_bwd_reduce_v8_kernel<true>, grid, 256, 0, STREAM, p);
      // (round 1's cp.async ring for this kernel measured slower inside the step, profiles/r01_elementwise_pipe_ab.txt; the
      // round-2 ring below -- three stages, one wave of long blocks -- is faster alone and no slower in the step)
      else if (v8 && reduce_ring()) {
        constexpr int smem = 2 * RED_STAGES * PIPE_U * 256 * 16;
        static bool configured = false;
        if (!configured) {
          cudaFuncSetAttribute(norm_act_bwd_reduce_v8p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          configured = true;
        }
        launch_k(norm_act_bwd_reduce_v8p_kernel, grid, 256, smem, STREAM, p);
      }
      else if (v8) launch_k(norm_act_bwd_reduce_v8_kernel<false>, grid, 256, 0, STREAM, p);
      else if (p.x_bf16) launch_bwd_reduce<__nv_bfloat16>(p, grid, STREAM);
      else launch_bwd_reduce<float>(p, grid, STREAM);
      GANB_CHECK_LAUNCH("norm_act_bwd_reduce_kernel");
      const bool scatter = gamma && dgamma && dbeta;
      const int rows = (labels && n_rows > 0) ? n_rows : 1;
      static const bool fused_ok = !(getenv("GANB_BN_BWD_FUSED") && getenv("GANB_BN_BWD_FUSED")[0] == '0');   // A/B switch
      // (few statistic groups only: the block walks the groups one after the other -- with instance norm, groups = n, that
      // loop made the Pix2Pix generator step 25 % slower, 9.5 -> 11.9 ms)
      if (fused_ok && n <= 256 && rows <= 32 && groups <= 4) {
        const size_t smem = static_cast<size_t>(n) * 128 + 2 * 256 * 16 + static_cast<size_t>(n) * 4;
        launch_k(norm_act_bwd_finalize_fused_kernel, ceil_div(c, 16), 256, smem, STREAM, p.part, n, c, p.chunks, groups,
                 gamma, labels, rows, sums, s1, s2, scatter ? dgamma : nullptr, scatter ? dbeta : nullptr);
        GANB_CHECK_LAUNCH("norm_act_bwd_finalize_fused_kernel");
      } else {
        launch_k(norm_act_bwd_finalize_kernel, dim3(ceil_div(c, 8), groups), 256, 0, STREAM, p.part, n, c, p.chunks, groups,
                 gamma, labels, sums, s1, s2);
        GANB_CHECK_LAUNCH("norm_act_bwd_finalize_kernel");
        if (scatter) {
          launch_k(norm_act_bwd_scatter_kernel, ceil_div(rows * c, 8), 256, 0, STREAM, sums, n, c, rows, labels, dgamma, dbeta);
          GANB_CHECK_LAUNCH("norm_act_bwd_scatter_kernel");
        }
      }
    }
    p.s1 = s1; p.s2 = s2;
    p.inv_count = count_scale / (static_cast<float>(n / groups) * hw);
  }
  if (phase == 1) return mean ? 0 : fail(GANB_E_BADARG, "norm_act_bwd: phase 1 needs normalisation statistics");
  {
    const int chunks2 = v8 ? v8_chunks(n, h * w, bwd_bps()) : bwd_chunks(n, h * w);
    const int ppc = ceil_div(h * w, chunks2);
    const dim3 grid(ceil_div(h * w, ppc), n);
    if (v8) {
      constexpr int smem = 3 * BWD_STAGES * PIPE_U * 256 * 16;
      static bool configured = false;
      if (!configured) {
        cudaFuncSetAttribute(norm_act_bwd_apply_v8p_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
      }
      const int idx = (mean ? 2 : 0) | (upsample ? 1 : 0);
      switch (idx) {
        case 0: launch_k(norm_act_bwd_apply_v8_kernel<false, false>, grid, 256, 0, STREAM, p, ppc); break;   // neutral in A/B
        case 1: launch_k(norm_act_bwd_apply_v8_kernel<false, true>, grid, 256, 0, STREAM, p, ppc); break;
        case 2: launch_k(norm_act_bwd_apply_v8p_kernel<true>, grid, 256, smem, STREAM, p, ppc); break;
        default: launch_k(norm_act_bwd_apply_v8_kernel<true, true>, grid, 256, 0, STREAM, p, ppc); break;
      }
    } else if (p.x_bf16) launch_bwd_apply<__nv_bfloat16>(p, mean != nullptr, grid, ppc, STREAM);
    else launch_bwd_apply<float>(p, mean != nullptr, grid, ppc, STREAM);
  }
  GANB_CHECK_LAUNCH("norm_act_bwd_apply_kernel");
  return 0;
}

extern "C" int ganb_norm_act_bwd(const void* x, int x_dtype, const void* dz, int dz_dtype, int dz_cstride, int n, int h,
                                 int w, int c, const float* mean, const float* rstd, int groups, const float* gamma,
                                 const float* beta, const int* labels, int n_rows, int act, int upsample,
                                 float* dgamma, float* dbeta, const void* add, int add_dtype, void* dx, int dx_dtype,
                                 void* workspace, void* stream) {
  return norm_act_bwd_impl(x, x_dtype, dz, dz_dtype, dz_cstride, n, h, w, c, mean, rstd, groups, gamma, beta, labels, n_rows,
                           act, upsample, dgamma, dbeta, add, add_dtype, dx, dx_dtype, workspace, 0, 1.0f, stream);
}

extern "C" int ganb_norm_act_bwd_phase(const void* x, int x_dtype, const void* dz, int dz_dtype, int dz_cstride, int n,
                                       int h, int w, int c, const float* mean, const float* rstd, int groups,
                                       const float* gamma, const float* beta, const int* labels, int n_rows, int act,
                                       int upsample, float* dgamma, float* dbeta, const void* add, int add_dtype, void* dx,
                                       int dx_dtype, void* workspace, int phase, float count_scale, void* stream) {
  if (phase != 1 && phase != 2) return fail(GANB_E_BADARG, "norm_act_bwd_phase: phase must be 1 or 2");
  return norm_act_bwd_impl(x, x_dtype, dz, dz_dtype, dz_cstride, n, h, w, c, mean, rstd, groups, gamma, beta, labels, n_rows,
                           act, upsample, dgamma, dbeta, add, add_dtype, dx, dx_dtype, workspace, phase, count_scale, stream);
}

extern "C" int64_t ganb_norm_act_bwd_sums_offset(int n, int hw, int c, int groups) {
  (void)groups;
  return (static_cast<int64_t>(n) * bwd_chunks(n, hw) * 2 * c + 2LL * n * c) * 4;
}

namespace ganb {
// cross-GPU batch statistics: every rank contributes [mean, E[x^2]] of its (equally sized) share of a statistic group
__global__ void __launch_bounds__(256) bn_moments_pack_kernel(const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              int count, float eps, float* __restrict__ out) {
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= count) return;
  const float m = mean[i], r = rstd[i];
  out[i] = m;
  out[count + i] = 1.0f / (r * r) - eps + m * m;
}
__global__ void __launch_bounds__(256) bn_moments_unpack_kernel(const float* __restrict__ sums, int count, float inv_world,
                                                                float eps, float* __restrict__ mean,
                                                                float* __restrict__ rstd) {
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= count) return;
  const float m = sums[i] * inv_world;
  float var = sums[count + i] * inv_world - m * m;
  if (var < 0.f) var = 0.f;
  mean[i] = m;
  rstd[i] = rsqrtf(var + eps);
}
}  // namespace ganb

extern "C" int ganb_bn_moments_pack(const float* mean, const float* rstd, int count, float eps, float* out, void* stream) {
  if (!mean || !rstd || !out || count <= 0) return fail(GANB_E_BADARG, "bn_moments_pack: bad arguments");
  launch_k(bn_moments_pack_kernel, ceil_div(count, 256), 256, 0, STREAM, mean, rstd, count, eps, out);
  GANB_CHECK_LAUNCH("bn_moments_pack_kernel");
  return 0;
}

extern "C" int ganb_bn_moments_unpack(const float* sums, int count, float inv_world, float eps, float* mean, float* rstd,
                                      void* stream) {
  if (!sums || !mean || !rstd || count <= 0) return fail(GANB_E_BADARG, "bn_moments_unpack: bad arguments");
  launch_k(bn_moments_unpack_kernel, ceil_div(count, 256), 256, 0, STREAM, sums, count, inv_world, eps, mean, rstd);
  GANB_CHECK_LAUNCH("bn_moments_unpack_kernel");
  return 0;
}

template <typename TIn, typename TOut>
static int launch_meanpool(const void* x, const float* add, void* out, int n, int ho, int wo, int c, cudaStream_t s) {
  if (c % 4 != 0) {
    launch_k(pool2_scalar_kernel<TIn, TOut>, grid_for(static_cast<int64_t>(n) * ho * wo * c, 256), 256, 0, s, static_cast<const TIn*>(x), add, static_cast<TOut*>(out), n, ho, wo, c, 0.25f);
    GANB_CHECK_LAUNCH("pool2_scalar_kernel");
    return 0;
  }
  const int64_t items = static_cast<int64_t>(n) * ho * wo * (c / 4);
  launch_k(meanpool2_fwd_kernel<TIn, TOut>, grid_for(items, 256), 256, 0, s, static_cast<const TIn*>(x), add,
                                                                       static_cast<TOut*>(out), n, ho, wo, c);
  GANB_CHECK_LAUNCH("meanpool2_fwd_kernel");
  return 0;
}

extern "C" int ganb_meanpool2_fwd(const void* x, int x_dtype, const float* add, void* out, int out_dtype, int n,
                                  int h, int w, int c, void* stream) {
  if (!x || !out) return fail(GANB_E_BADARG, "meanpool2_fwd: null buffer");
  if ((h & 1) || (w & 1)) return fail(GANB_E_UNSUPPORTED, "meanpool2_fwd: even h,w required");
  const int ho = h / 2, wo = w / 2;
  if (x_dtype == GANB_F32 && out_dtype == GANB_F32) return launch_meanpool<float, float>(x, add, out, n, ho, wo, c, STREAM);
  if (x_dtype == GANB_F32 && out_dtype == GANB_BF16) return launch_meanpool<float, __nv_bfloat16>(x, add, out, n, ho, wo, c, STREAM);
  if (x_dtype == GANB_BF16 && out_dtype == GANB_BF16) return launch_meanpool<__nv_bfloat16, __nv_bfloat16>(x, add, out, n, ho, wo, c, STREAM);
  return launch_meanpool<__nv_bfloat16, float>(x, add, out, n, ho, wo, c, STREAM);
}

template <typename TIn, typename TOut>
static int launch_expand(const void* x, void* out, int n, int h, int w, int c, float scale, cudaStream_t s) {
  if (c % 4 != 0) {
    launch_k(expand2_scalar_kernel<TIn, TOut>, grid_for(static_cast<int64_t>(n) * h * w * c, 256), 256, 0, s, static_cast<const TIn*>(x), static_cast<TOut*>(out), n, h, w, c, scale);
    GANB_CHECK_LAUNCH("expand2_scalar_kernel");
    return 0;
  }
  const int64_t items = static_cast<int64_t>(n) * h * w * (c / 4);
  launch_k(expand2_kernel<TIn, TOut>, grid_for(items, 256), 256, 0, s, static_cast<const TIn*>(x), static_cast<TOut*>(out), n, h, w, c, scale);
  GANB_CHECK_LAUNCH("expand2_kernel");
  return 0;
}

// out[n,2h,2w,c] = scale * x[n,h,w,c] replicated 2x2 (upsample fwd: scale 1; mean-pool bwd: scale 0.25)
extern "C" int ganb_expand2(const void* x, int x_dtype, void* out, int out_dtype, int n, int h, int w, int c,
                            float scale, void* stream) {
  if (!x || !out) return fail(GANB_E_BADARG, "expand2: null buffer");
  if (x_dtype == GANB_F32 && out_dtype == GANB_F32) return launch_expand<float, float>(x, out, n, h, w, c, scale, STREAM);
  if (x_dtype == GANB_F32 && out_dtype == GANB_BF16) return launch_expand<float, __nv_bfloat16>(x, out, n, h, w, c, scale, STREAM);
  if (x_dtype == GANB_BF16 && out_dtype == GANB_BF16) return launch_expand<__nv_bfloat16, __nv_bfloat16>(x, out, n, h, w, c, scale, STREAM);
  return launch_expand<__nv_bfloat16, float>(x, out, n, h, w, c, scale, STREAM);
}

template <typename TIn, typename TOut>
static int launch_sum2x2(const void* x, void* out, int n, int ho, int wo, int c, float scale, cudaStream_t s) {
  if (c % 4 != 0) {
    launch_k(pool2_scalar_kernel<TIn, TOut>, grid_for(static_cast<int64_t>(n) * ho * wo * c, 256), 256, 0, s, static_cast<const TIn*>(x), nullptr, static_cast<TOut*>(out), n, ho, wo, c, scale);
    GANB_CHECK_LAUNCH("pool2_scalar_kernel");
    return 0;
  }
  const int64_t items = static_cast<int64_t>(n) * ho * wo * (c / 4);
  launch_k(sum2x2_kernel<TIn, TOut>, grid_for(items, 256), 256, 0, s, static_cast<const TIn*>(x), static_cast<TOut*>(out), n, ho, wo, c, scale);
  GANB_CHECK_LAUNCH("sum2x2_kernel");
  return 0;
}

// out[n,h/2,w/2,c] = scale * (sum of each 2x2 block of x[n,h,w,c])   (upsample bwd: scale 1)
extern "C" int ganb_sum2x2(const void* x, int x_dtype, void* out, int out_dtype, int n, int h, int w, int c,
                           float scale, void* stream) {
  if (!x || !out) return fail(GANB_E_BADARG, "sum2x2: null buffer");
  if ((h & 1) || (w & 1)) return fail(GANB_E_UNSUPPORTED, "sum2x2: even h,w required");
  const int ho = h / 2, wo = w / 2;
  if (x_dtype == GANB_F32 && out_dtype == GANB_F32) return launch_sum2x2<float, float>(x, out, n, ho, wo, c, scale, STREAM);
  if (x_dtype == GANB_F32 && out_dtype == GANB_BF16) return launch_sum2x2<float, __nv_bfloat16>(x, out, n, ho, wo, c, scale, STREAM);
  if (x_dtype == GANB_BF16 && out_dtype == GANB_BF16) return launch_sum2x2<__nv_bfloat16, __nv_bfloat16>(x, out, n, ho, wo, c, scale, STREAM);
  return launch_sum2x2<__nv_bfloat16, float>(x, out, n, ho, wo, c, scale, STREAM);
}

template <typename TIn, typename TOut>
static int launch_dilate(const void* x, void* out, int n, int h, int w, int c, int stride, int oh, int ow, cudaStream_t s) {
  const int64_t items = static_cast<int64_t>(n) * oh * ow * (c / 4);
  launch_k(dilate2d_kernel<TIn, TOut>, grid_for(items, 256), 256, 0, s, static_cast<const TIn*>(x), static_cast<TOut*>(out),
           n, h, w, c, stride, oh, ow);
  GANB_CHECK_LAUNCH("dilate2d_kernel");
  return 0;
}

extern "C" int ganb_dilate2d(const void* x, int x_dtype, void* out, int out_dtype, int n, int h, int w, int c, int stride,
                             int out_h, int out_w, void* stream) {
  if (!x || !out) return fail(GANB_E_BADARG, "dilate2d: null buffer");
  if (c % 4 != 0) return fail(GANB_E_UNSUPPORTED, "dilate2d: c=%d must be a multiple of 4", c);
  if (stride < 1 || out_h < (h - 1) * stride + 1 || out_w < (w - 1) * stride + 1)
    return fail(GANB_E_BADARG, "dilate2d: output %dx%d too small for %dx%d at stride %d", out_h, out_w, h, w, stride);
  if (x_dtype == GANB_F32 && out_dtype == GANB_F32) return launch_dilate<float, float>(x, out, n, h, w, c, stride, out_h, out_w, STREAM);
  if (x_dtype == GANB_F32 && out_dtype == GANB_BF16) return launch_dilate<float, __nv_bfloat16>(x, out, n, h, w, c, stride, out_h, out_w, STREAM);
  if (x_dtype == GANB_BF16 && out_dtype == GANB_BF16) return launch_dilate<__nv_bfloat16, __nv_bfloat16>(x, out, n, h, w, c, stride, out_h, out_w, STREAM);
  return launch_dilate<__nv_bfloat16, float>(x, out, n, h, w, c, stride, out_h, out_w, STREAM);
}

template <typename TIn, typename TOut>
static int launch_cast(const void* x, void* y, int64_t n, float scale, cudaStream_t s) {
  const int64_t n4 = n / 4;
  if (n4 > 0) {
    launch_k(cast_kernel<TIn, TOut>, grid_for(n4, 256), 256, 0, s, static_cast<const TIn*>(x), static_cast<TOut*>(y), n4, scale);
    GANB_CHECK_LAUNCH("cast_kernel");
  }
  if (n % 4) {
    launch_k(cast_tail_kernel<TIn, TOut>, 1, 32, 0, s, static_cast<const TIn*>(x), static_cast<TOut*>(y), n4 * 4, n, scale);
    GANB_CHECK_LAUNCH("cast_tail_kernel");
  }
  return 0;
}

extern "C" int ganb_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t count, float scale, void* stream) {
  if (!x || !y) return fail(GANB_E_BADARG, "cast: null buffer");
  if (x_dtype == GANB_F32 && y_dtype == GANB_BF16) return launch_cast<float, __nv_bfloat16>(x, y, count, scale, STREAM);
  if (x_dtype == GANB_BF16 && y_dtype == GANB_F32) return launch_cast<__nv_bfloat16, float>(x, y, count, scale, STREAM);
  if (x_dtype == GANB_F32 && y_dtype == GANB_F32) return launch_cast<float, float>(x, y, count, scale, STREAM);
  return launch_cast<__nv_bfloat16, __nv_bfloat16>(x, y, count, scale, STREAM);
}

extern "C" int ganb_axpby(const float* x, float* y, int64_t count, float a, float b, void* stream) {
  if (!x || !y) return fail(GANB_E_BADARG, "axpby: null buffer");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return fail(GANB_E_BADARG, "axpby: buffers must be 16-byte aligned");
  launch_k(axpby_kernel, grid_for(count / 4 + 1, 256), 256, 0, STREAM, x, y, count, a, b);
  GANB_CHECK_LAUNCH("axpby_kernel");
  return 0;
}

extern "C" int64_t ganb_colsum_workspace(int64_t rows, int c) {
  (void)rows;
  return static_cast<int64_t>(1024) * c * 4;
}

namespace ganb {
template <typename TIn>
static int launch_colsum(const void* xv, int64_t rows, int c, float beta, float* out, void* workspace, cudaStream_t s) {
  const TIn* x = static_cast<const TIn*>(xv);
  if (rows <= 512 && c >= 1024) {  // dense-layer bias gradients: coalesced over columns, short serial loop over rows
    launch_k(colsum_wide_kernel<TIn>, ceil_div(c, 256), 256, 0, s, x, rows, c, beta, out);
    GANB_CHECK_LAUNCH("colsum_wide_kernel");
    return 0;
  }
  if (c % 4 != 0 && c > 8) {
    launch_k(colsum_scalar_kernel<TIn>, c, 256, 0, s, x, rows, c, beta, out);
    GANB_CHECK_LAUNCH("colsum_scalar_kernel");
    return 0;
  }
  const bool narrow = (c % 4 != 0);
  int chunks = narrow ? sm_count() : colsum_bps() * sm_count();
  const int64_t max_chunks = ceil_div64(rows, narrow ? 256 : 128);
  if (chunks > max_chunks) chunks = static_cast<int>(max_chunks);
  if (chunks > 1024) chunks = 1024;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = static_cast<int>(ceil_div64(rows, chunks));
  const int used = static_cast<int>(ceil_div64(rows, rows_per_chunk));
  if (narrow) {
    launch_k(colsum_narrow_kernel<TIn>, used, 256, 0, s, x, rows, c, rows_per_chunk, static_cast<float*>(workspace));
    GANB_CHECK_LAUNCH("colsum_narrow_kernel");
  } else {
    if (sizeof(TIn) == 2 && c % 8 == 0) {
      constexpr int smem = STATS_STAGES * PIPE_U * 256 * 16;
      static bool configured = false;
      if (!configured) {
        cudaFuncSetAttribute(colsum_partial_v8p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
      }
      launch_k(colsum_partial_v8p_kernel, used, 256, smem, s, reinterpret_cast<const __nv_bfloat16*>(x), rows, c,
               rows_per_chunk, static_cast<float*>(workspace));
    }
    else
      launch_k(colsum_partial_kernel<TIn>, used, 256, 0, s, x, rows, c, rows_per_chunk, static_cast<float*>(workspace));
    GANB_CHECK_LAUNCH("colsum_partial_kernel");
  }
  launch_k(colsum_finalize_kernel, ceil_div(c, 32), 1024, 0, s, static_cast<float*>(workspace), c, used, beta, out);
  GANB_CHECK_LAUNCH("colsum_finalize_kernel");
  return 0;
}
}  // namespace ganb

// out[c] = beta*out[c] + sum_rows x[row][c]    (bias gradient: tf.nn.bias_add backward)
extern "C" int ganb_colsum(const void* x, int x_dtype, int64_t rows, int c, float beta, float* out, void* workspace,
                           void* stream) {
  if (!x || !out || !workspace) return fail(GANB_E_BADARG, "colsum: null buffer");
  if (x_dtype == GANB_BF16) return launch_colsum<__nv_bfloat16>(x, rows, c, beta, out, workspace, STREAM);
  return launch_colsum<float>(x, rows, c, beta, out, workspace, STREAM);
}

extern "C" int ganb_bcast_channels_fwd(const float* e, int n, int hw, int c2, int coff, int cstride, int act,
                                       void* out_raw_bf16, void* out_act_bf16, void* stream) {
  if (!e) return fail(GANB_E_BADARG, "bcast_channels_fwd: null buffer");
  if (c2 % 4 || coff % 4 || cstride % 4) return fail(GANB_E_UNSUPPORTED, "bcast_channels_fwd: channel counts must be multiples of 4");
  const int64_t items = static_cast<int64_t>(n) * hw * (c2 / 4);
  launch_k(bcast_channels_kernel, grid_for(items, 256), 256, 0, STREAM, e, n, hw, c2, coff, cstride, act,
                                                                  static_cast<__nv_bfloat16*>(out_raw_bf16),
                                                                  static_cast<__nv_bfloat16*>(out_act_bf16));
  GANB_CHECK_LAUNCH("bcast_channels_kernel");
  return 0;
}

extern "C" int ganb_bcast_channels_bwd(const float* e, int n, int hw, int c2, int coff, int cstride, int act,
                                       const void* d_raw, const void* d_act, int d_dtype, float* de, void* stream) {
  if (!e || !de) return fail(GANB_E_BADARG, "bcast_channels_bwd: null buffer");
  if (c2 % 4 || coff % 4 || cstride % 4) return fail(GANB_E_UNSUPPORTED, "bcast_channels_bwd: channel counts must be multiples of 4");
  if (d_dtype == GANB_BF16)
    launch_k(bcast_channels_bwd_kernel<__nv_bfloat16>, n, 256, 0, STREAM, e, hw, c2, coff, cstride, act,
                                                                    static_cast<const __nv_bfloat16*>(d_raw),
                                                                    static_cast<const __nv_bfloat16*>(d_act), de);
  else
    launch_k(bcast_channels_bwd_kernel<float>, n, 256, 0, STREAM, e, hw, c2, coff, cstride, act,
                                                            static_cast<const float*>(d_raw),
                                                            static_cast<const float*>(d_act), de);
  GANB_CHECK_LAUNCH("bcast_channels_bwd_kernel");
  return 0;
}

namespace ganb {
template <typename TG, typename TOut>
static void launch_concat_bwd_x(const float* x, int64_t pixels, int c1, int cstride, int act, const void* d_raw,
                                const void* d_act, void* dx, cudaStream_t s) {
  launch_k(concat_bwd_x_kernel<TG, TOut>, grid_for(pixels * (c1 / 4), 256), 256, 0, s, x, pixels, c1, cstride, act, static_cast<const TG*>(d_raw), static_cast<const TG*>(d_act), static_cast<TOut*>(dx));
}
}  // namespace ganb

extern "C" int ganb_concat_bwd_x(const float* x, int64_t pixels, int c1, int cstride, int act, const void* d_raw,
                                 const void* d_act, int d_dtype, void* dx, int dx_dtype, void* stream) {
  if (!x || !dx) return fail(GANB_E_BADARG, "concat_bwd_x: null buffer");
  if (c1 % 4 || cstride % 4) return fail(GANB_E_UNSUPPORTED, "concat_bwd_x: channel counts must be multiples of 4");
  const bool g16 = d_dtype == GANB_BF16, o16 = dx_dtype == GANB_BF16;
  if (g16 && o16) launch_concat_bwd_x<__nv_bfloat16, __nv_bfloat16>(x, pixels, c1, cstride, act, d_raw, d_act, dx, STREAM);
  else if (g16) launch_concat_bwd_x<__nv_bfloat16, float>(x, pixels, c1, cstride, act, d_raw, d_act, dx, STREAM);
  else if (o16) launch_concat_bwd_x<float, __nv_bfloat16>(x, pixels, c1, cstride, act, d_raw, d_act, dx, STREAM);
  else launch_concat_bwd_x<float, float>(x, pixels, c1, cstride, act, d_raw, d_act, dx, STREAM);
  GANB_CHECK_LAUNCH("concat_bwd_x_kernel");
  return 0;
}

extern "C" int ganb_act_mean_hw_fwd(const float* x, int n, int hw, int c, int act, float* out, void* stream) {
  if (!x || !out) return fail(GANB_E_BADARG, "act_mean_hw_fwd: null buffer");
  if (c % 4) return fail(GANB_E_UNSUPPORTED, "act_mean_hw_fwd: c=%d must be a multiple of 4", c);
  launch_k(act_mean_hw_fwd_kernel, n, 256, 0, STREAM, x, hw, c, act, out);
  GANB_CHECK_LAUNCH("act_mean_hw_fwd_kernel");
  return 0;
}

extern "C" int ganb_act_mean_hw_bwd(const float* x, const float* dout, int n, int hw, int c, int act, void* dx,
                                    int dx_dtype, void* stream) {
  if (!x || !dout || !dx) return fail(GANB_E_BADARG, "act_mean_hw_bwd: null buffer");
  if (c % 4) return fail(GANB_E_UNSUPPORTED, "act_mean_hw_bwd: c=%d must be a multiple of 4", c);
  const int grid = grid_for(static_cast<int64_t>(n) * hw * (c / 4), 256);
  if (dx_dtype == GANB_BF16)
    launch_k(act_mean_hw_bwd_kernel<__nv_bfloat16>, grid, 256, 0, STREAM, x, dout, n, hw, c, act, static_cast<__nv_bfloat16*>(dx));
  else
    launch_k(act_mean_hw_bwd_kernel<float>, grid, 256, 0, STREAM, x, dout, n, hw, c, act, static_cast<float*>(dx));
  GANB_CHECK_LAUNCH("act_mean_hw_bwd_kernel");
  return 0;
}

extern "C" int ganb_gan_loss(const float* logits, int n, int n_real, int mode, float scale, int accumulate,
                             float* loss_out, float* dlogits, void* stream) {
  if (!logits || !loss_out || !dlogits) return fail(GANB_E_BADARG, "gan_loss: null buffer");
  if (mode < 0 || mode > 11) return fail(GANB_E_UNSUPPORTED, "gan_loss: mode %d", mode);
  if ((mode & 1) == 0 && (n_real <= 0 || n_real >= n)) return fail(GANB_E_BADARG, "gan_loss: the discriminator loss needs 0 < n_real < n");
  launch_k(gan_loss_kernel, 1, 256, 0, STREAM, logits, n, n_real, mode, scale, accumulate, loss_out, dlogits);
  GANB_CHECK_LAUNCH("gan_loss_kernel");
  return 0;
}

extern "C" int ganb_softmax_xent(const float* logits, const int* labels, int n, int c, float scale, int accumulate,
                                 float* loss_out, float* dlogits, void* stream) {
  if (!logits || !labels || !loss_out || !dlogits) return fail(GANB_E_BADARG, "softmax_xent: null buffer");
  if (n <= 0 || c <= 0) return fail(GANB_E_BADARG, "softmax_xent: empty input");
  launch_k(softmax_xent_kernel, 1, 256, 0, STREAM, logits, labels, n, c, scale, accumulate, loss_out, dlogits);
  GANB_CHECK_LAUNCH("softmax_xent_kernel");
  return 0;
}

extern "C" int ganb_adam(float* params, const float* grads, float* m, float* v, int64_t count, const float* lr_t,
                         float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!params || !grads || !m || !v || !lr_t) return fail(GANB_E_BADARG, "adam: null buffer");
  launch_k(adam_kernel, grid_for(count / 4 + 1, 256), 256, 0, STREAM, params, grads, m, v, count, lr_t, beta1, beta2, eps, grad_scale);
  GANB_CHECK_LAUNCH("adam_kernel");
  return 0;
}

extern "C" int ganb_preprocess_real(const int* data, const float* noise, int b, int hw, float* out, void* stream) {
  if (!data || !out) return fail(GANB_E_BADARG, "preprocess_real: null buffer");
  launch_k(preprocess_real_kernel, grid_for(static_cast<int64_t>(b) * hw * 3, 256), 256, 0, STREAM, data, noise, b, hw, out);
  GANB_CHECK_LAUNCH("preprocess_real_kernel");
  return 0;
}

extern "C" int ganb_embedding_fwd(const float* table, const int* labels, int n, int dim, float* out, void* stream) {
  if (!table || !labels || !out) return fail(GANB_E_BADARG, "embedding_fwd: null buffer");
  launch_k(embedding_fwd_kernel, grid_for(static_cast<int64_t>(n) * dim, 256), 256, 0, STREAM, table, labels, n, dim, out);
  GANB_CHECK_LAUNCH("embedding_fwd_kernel");
  return 0;
}

extern "C" int ganb_embedding_bwd(const float* dout, const int* labels, int n, int dim, int vocab, float* dtable,
                                  void* stream) {
  if (!dout || !labels || !dtable) return fail(GANB_E_BADARG, "embedding_bwd: null buffer");
  launch_k(embedding_bwd_kernel, grid_for(static_cast<int64_t>(vocab) * dim, 256), 256, 0, STREAM, dout, labels, n, dim, vocab, dtable);
  GANB_CHECK_LAUNCH("embedding_bwd_kernel");
  return 0;
}
