// Tensor-core convolution kernels for sm_100a: implicit GEMM on tcgen05.mma with TMEM accumulators,
// operands staged in shared memory by TMA (4-D tiled boxes give the im2col view without materialising it).
//
//   conv_igemm_kernel : fprop / dgrad / 1x1 / linear.  A = activation box (K-major, 128B swizzle),
//                       B = packed filter [tap][cout][cin] (K-major).  Persistent, warp-specialised:
//                       warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-9 epilogue.
//   conv_wgrad_kernel : filter gradient.  Both operands are MN-major (channels contiguous, pixels = K),
//                       split over pixel ranges; partials reduced by splitk_reduce_kernel.
//
// Replaces tf.nn.conv2d and its two gradients (reference common/ops/conv2d.py:181-187).
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace ganb {

// Pipeline depths (shared-memory stages).  Every kernel below is one CTA per SM; what it leaves of the 227 KB decides
// which bandwidth-bound blocks of the other streams can run next to it (the cp.async-ring kernels of elementwise.cu
// need 48-96 KB).  GANB_LEAN_SMEM builds the shallower set for A/B runs (profiles/r02_smem_stages_ab.txt).
#ifdef GANB_LEAN_SMEM
constexpr int ST_PAIR256_SB = 6, ST_PAIR128_SB = 8, ST_HALO64_SB = 6, ST_HALO128_SB = 6, ST_HALO256_SB = 4;
constexpr int ST_IG64 = 6, ST_IG128 = 5, ST_IG256 = 3, ST_WG64 = 4, ST_WG128 = 4, ST_WG256 = 3;
#else
constexpr int ST_PAIR256_SB = 8, ST_PAIR128_SB = 12, ST_HALO64_SB = 9, ST_HALO128_SB = 8, ST_HALO256_SB = 5;
constexpr int ST_IG64 = 8, ST_IG128 = 6, ST_IG256 = 4, ST_WG64 = 6, ST_WG128 = 6, ST_WG256 = 4;
#endif

constexpr int BM = 128;                      // UMMA M: output pixels (igemm) / input channels (wgrad)
constexpr int BK = 64;                       // bf16 elements per 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
// Warp roles of the fprop / dgrad kernels: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2.. epilogue.
// A warp may only read the TMEM lane quadrant (warp & 3), so the 128 accumulator rows need four warps.
// GANB_EPI_SPLIT = 2 drains tiles of >= 64 channels with EIGHT warps (two per quadrant, half of the columns each).
// Measured and NOT used: the kernel time of the K = 2304 layers is unchanged (they are MMA-paced), while 352 threads x
// ~165 registers take 88 % of the SM's register file and evict the bandwidth-bound kernels of the other streams that
// otherwise run next to the convolution: +0.15 ms per D+G pair (profiles/r02_fused_stats_ab.txt).
#ifndef GANB_EPI_SPLIT
#define GANB_EPI_SPLIT 1
#endif
constexpr int NUM_THREADS = 64 + 128 * GANB_EPI_SPLIT;
constexpr int WG_THREADS = 192;              // filter-gradient kernel: warps 2..5 drain its single accumulator
__host__ __device__ constexpr int epi_split(int bn) { return (GANB_EPI_SPLIT == 2 && bn >= 64) ? 2 : 1; }
__host__ __device__ constexpr int epi_threads(int bn) { return 128 * epi_split(bn); }

struct IgemmParams {
  int N, Ho, Wo, Cout;
  int taps, kw;
  int stride, pad_t, pad_l;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n, tiles_co, num_tiles;
  int kchunks, flip;
  const float* alpha;
  const float* bias;
  const float* residual;
  int res_up2;   // residual is stored at half resolution [N, Ho/2, Wo/2, Cout] and read through a nearest-2x upsample
  void* out;
  int out_bf16, act;
  // Convolution groups over ONE input (sub-pixel form of UpsampleConv, see ganb_conv2d_up2_*): `og` output groups =
  // every pixel tile is computed og times with its own filter taps, padding and output channel offset g*Cout (output
  // pixel stride out_cstride); `rg` reduction groups = the K loop runs over rg channel slices of the input, each with its
  // own taps and padding.  At most one of og / rg exceeds 1.
  int og, rg, out_cstride;
  signed char gpad_t[4], gpad_l[4];
  // Batch statistics of the STORED output (after alpha / bias / residual / activation and the rounding to the output
  // dtype), produced by the epilogue for the batch norm that follows the layer: stats[(pixel_tile * og + g) * 2 + {0, 1}]
  // [Cout] = per-tile column sums of y and y^2 (rows outside the tensor contribute 0).  Pixel tiles are image-major,
  // so the rows of one statistic tower are contiguous and bn_stats_finalize_kernel folds them in a fixed order
  // (deterministic).  nullptr: off.
  float* stats;
  // Derivative of an activation that the PRODUCER of this layer's input applied in its own epilogue (data-gradient
  // launches): gate = act(pre-activation), bf16, laid out like the output; y *= act'(pre), read off the sign of gate
  // (relu: gate > 0; leaky relu: gate >= 0 ? 1 : 0.2).  nullptr: off.
  const __nv_bfloat16* gate;
  int gate_act;
  // 32-byte (256-bit) stores: the output pointer, the pixel stride and the row length are multiples of 32 bytes, so every
  // store instruction of a thread fills one whole 32-byte sector (16-byte stores leave the sector to a second instruction:
  // 2x the L2 write sectors on the store-bound shallow-K layers, profiles/r02_ncu_igemm_k32.txt)
  int st32;
  int res32;   // the same for the fp32 residual rows (32-byte loads)
};

template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[NC]);

template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&r)[32]) {
  tmem_ld_32x32(taddr, r);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void stg32(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                      uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
__device__ __forceinline__ void ldg32(const void* ptr, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(ptr));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 q = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&q);
}

__device__ __noinline__ float4 tanh4(float4 a) {
  return make_float4(tanhf(a.x), tanhf(a.y), tanhf(a.z), tanhf(a.w));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == GANB_ACT_TANH) return tanhf(v);
  if (act == GANB_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == GANB_ACT_LRELU) return v >= 0.f ? v : 0.2f * v;
  return v;
}

// Fused batch statistics: the 4 epilogue warps of a CTA leave their 32-pixel column sums in shared memory
// ([warp][{sum, sumsq}][BN]); epilogue_stats_flush folds them and writes the tile's row of IgemmParams::stats.
// The sums over the 32 pixels of a warp go through a padded shared-memory transpose (t): 32 conflict-free stores and
// 32 conflict-free loads per 32-channel chunk, all independent.  (A shuffle butterfly costs the same instruction
// count but its five dependent levels stall the single epilogue warp of each scheduler: +35 % on the 256-channel
// CTA-pair kernel, profiles/r02_fused_stats_ab.txt.)
template <int BN>
struct EpiStats {
  float s[4][2][BN];
  float t[8][8][33];
};

// `ew`: index of the epilogue warp (0..7), q its TMEM lane quadrant, c0 the first of its 32 channels inside the tile
template <int NC, int BN>
__device__ __forceinline__ void epilogue_stats_chunk(EpiStats<BN>& st, const float (&val)[NC], bool valid, int ew, int q,
                                                     int lane, int c0) {
  static_assert(NC == 32, "fused statistics need 32-column chunks");
  const int ch = lane & 7, k0 = (lane >> 3) * 8;
#pragma unroll
  for (int part = 0; part < 4; ++part) {
#pragma unroll
    for (int j = 0; j < 8; ++j) st.t[ew][j][lane] = valid ? val[part * 8 + j] : 0.f;
    __syncwarp();
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {          // fixed order: deterministic
      const float v = st.t[ew][ch][k0 + k];
      sum += v;
      sq += v * v;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
    sq += __shfl_xor_sync(0xffffffffu, sq, 8);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    sq += __shfl_xor_sync(0xffffffffu, sq, 16);
    if (lane < 8) {
      st.s[q][0][c0 + part * 8 + ch] = sum;
      st.s[q][1][c0 + part * 8 + ch] = sq;
    }
    __syncwarp();
  }
}

// called by all epilogue threads of the CTA after the chunk loop of a tile (contains a named barrier)
template <int BN>
__device__ __forceinline__ void epilogue_stats_flush(const IgemmParams& p, const EpiStats<BN>& st, int64_t row,
                                                     bool row_ok, int co_tile, int et) {
  asm volatile("bar.sync 1, %0;" ::"n"(epi_threads(BN)) : "memory");
  if (row_ok) {
    float* out = p.stats + row * 2 * p.Cout;
    for (int i = et; i < 2 * BN; i += epi_threads(BN)) {
      const int k = i / BN, c = i - k * BN;
      const int co = co_tile + c;
      if (co < p.Cout) out[k * p.Cout + co] = (st.s[0][k][c] + st.s[1][k][c]) + (st.s[2][k][c] + st.s[3][k][c]);
    }
  }
}

// One thread owns one output pixel (TMEM lane) and NC consecutive channels held in registers.
// All loads (bias from shared memory, residual from global) are issued BEFORE the first store: the output pointer
// may alias the inputs as far as the compiler knows, so interleaving them would serialise one memory round trip
// per float4 (this was the limiter of the first version of this kernel, profiles/r01_*).
// `val` (optional): receives the values as stored (rounded to the output dtype), for the fused statistics.
template <int NC>
__device__ __forceinline__ void epilogue_row(const IgemmParams& p, const uint32_t (&r)[NC], float alpha,
                                             int64_t out_off, int64_t res_pix, int co_base,
                                             const float* __restrict__ bias_s, float* __restrict__ val = nullptr) {
  const int cout = p.Cout;
  const int64_t off = out_off + co_base;           // out_off = pixel * out_cstride + group channel offset
  const int64_t roff = res_pix * cout + co_base;
  if (cout % 4 == 0) {
    float4 v[NC / 4];
#pragma unroll
    for (int j = 0; j < NC; j += 4)
      v[j / 4] = make_float4(__uint_as_float(r[j]) * alpha, __uint_as_float(r[j + 1]) * alpha,
                             __uint_as_float(r[j + 2]) * alpha, __uint_as_float(r[j + 3]) * alpha);
    if (p.residual) {
      float4 q[NC / 4];
      if (p.res32 && NC % 8 == 0) {
#pragma unroll
        for (int j = 0; j < NC; j += 8) {
          uint32_t w8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (co_base + j < cout) ldg32(p.residual + roff + j, w8);
          q[j / 4] = make_float4(__uint_as_float(w8[0]), __uint_as_float(w8[1]), __uint_as_float(w8[2]), __uint_as_float(w8[3]));
          q[j / 4 + 1] = make_float4(__uint_as_float(w8[4]), __uint_as_float(w8[5]), __uint_as_float(w8[6]), __uint_as_float(w8[7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < NC; j += 4)
          q[j / 4] = (co_base + j < cout) ? __ldg(reinterpret_cast<const float4*>(p.residual + roff + j))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) { v[j].x += q[j].x; v[j].y += q[j].y; v[j].z += q[j].z; v[j].w += q[j].w; }
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < NC; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + j);
        v[j / 4].x += b.x; v[j / 4].y += b.y; v[j / 4].z += b.z; v[j / 4].w += b.w;
      }
    }
    // one uniform branch per activation AROUND the unrolled loops (a per-element switch costs a branch per element and
    // doubled the time of the shallow-K layers)
    if (p.act == GANB_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) {
        v[j].x = v[j].x > 0.f ? v[j].x : 0.f; v[j].y = v[j].y > 0.f ? v[j].y : 0.f;
        v[j].z = v[j].z > 0.f ? v[j].z : 0.f; v[j].w = v[j].w > 0.f ? v[j].w : 0.f;
      }
    } else if (p.act == GANB_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) {
        v[j].x = v[j].x >= 0.f ? v[j].x : 0.2f * v[j].x; v[j].y = v[j].y >= 0.f ? v[j].y : 0.2f * v[j].y;
        v[j].z = v[j].z >= 0.f ? v[j].z : 0.2f * v[j].z; v[j].w = v[j].w >= 0.f ? v[j].w : 0.2f * v[j].w;
      }
    } else if (p.act == GANB_ACT_TANH) {
      // unrolled (a rolled loop would index v[] dynamically and push the array into local memory) around an out-of-line
      // tanh (inlined, its 32 interleaved copies cost every instance of the kernel ~35 registers)
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) v[j] = tanh4(v[j]);
    }
    if (p.gate) {
      if constexpr (NC % 8 == 0) {
        uint4 g[NC / 8];
#pragma unroll
        for (int j = 0; j < NC; j += 8)
          g[j / 8] = (co_base + j < cout) ? __ldg(reinterpret_cast<const uint4*>(p.gate + off + j)) : make_uint4(0, 0, 0, 0);
        const float neg = p.gate_act == GANB_ACT_LRELU ? 0.2f : 0.f;
        const uint32_t zero_is_on = p.gate_act == GANB_ACT_LRELU ? 1u : 0u;   // lrelu'(0) = 1, relu'(0) = 0
        // bf16 pair: element 2k in the low half, 2k+1 in the high half; "on" = positive (or +-0 for leaky relu)
        auto gate2 = [&](float& a, float& b, uint32_t w2) {
          const uint32_t lo = w2 & 0xffffu, hi = w2 >> 16;
          const bool on_lo = (lo & 0x7fffu) ? !(lo & 0x8000u) : zero_is_on;
          const bool on_hi = (hi & 0x7fffu) ? !(hi & 0x8000u) : zero_is_on;
          if (!on_lo) a *= neg;
          if (!on_hi) b *= neg;
        };
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
          gate2(v[2 * j].x, v[2 * j].y, g[j].x);
          gate2(v[2 * j].z, v[2 * j].w, g[j].y);
          gate2(v[2 * j + 1].x, v[2 * j + 1].y, g[j].z);
          gate2(v[2 * j + 1].z, v[2 * j + 1].w, g[j].w);
        }
      }
    }
    if (val) {
#pragma unroll
      for (int j = 0; j < NC; j += 4) {
        float4 t = v[j / 4];
        if (p.out_bf16) {
          t.x = __bfloat162float(__float2bfloat16_rn(t.x)); t.y = __bfloat162float(__float2bfloat16_rn(t.y));
          t.z = __bfloat162float(__float2bfloat16_rn(t.z)); t.w = __bfloat162float(__float2bfloat16_rn(t.w));
        }
        val[j] = t.x; val[j + 1] = t.y; val[j + 2] = t.z; val[j + 3] = t.w;
      }
    }
    if (p.st32 && p.out_bf16 && (NC % 16 == 0)) {
      // 32-byte stores: 16 bf16 channels = one whole sector per instruction (st32 implies cout % 16 == 0)
#pragma unroll
      for (int j = 0; j < NC; j += 16) {
        if (co_base + j >= cout) break;
        const float4 a = v[j / 4], b = v[j / 4 + 1], c = v[j / 4 + 2], d = v[j / 4 + 3];
        stg32(reinterpret_cast<__nv_bfloat16*>(p.out) + off + j, pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w),
              pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w), pack_bf16x2(c.x, c.y), pack_bf16x2(c.z, c.w),
              pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
      }
    } else if (p.st32 && !p.out_bf16 && (NC % 8 == 0)) {
      // 32-byte stores: 8 fp32 channels per instruction (st32 implies cout % 8 == 0)
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        if (co_base + j >= cout) break;
        const float4 a = v[j / 4], b = v[j / 4 + 1];
        stg32(reinterpret_cast<float*>(p.out) + off + j, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z),
              __float_as_uint(a.w), __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
      }
    } else if (p.out_bf16 && (cout % 8 == 0) && (NC % 8 == 0)) {
      // 16-byte stores: 8 bf16 channels per instruction (half the store instructions of the 8-byte form)
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        if (co_base + j >= cout) break;
        const float4 a = v[j / 4], b = v[j / 4 + 1];
        __nv_bfloat162 q0 = __floats2bfloat162_rn(a.x, a.y), q1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 q2 = __floats2bfloat162_rn(b.x, b.y), q3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
        pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off + j) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; j += 4) {
        if (co_base + j >= cout) break;
        if (p.out_bf16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(v[j / 4].x, v[j / 4].y);
          __nv_bfloat162 hi = __floats2bfloat162_rn(v[j / 4].z, v[j / 4].w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off + j) = pk;
        } else {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off + j) = v[j / 4];
        }
      }
    }
  } else {
    float v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      v[j] = __uint_as_float(r[j]) * alpha;
      if (co_base + j < cout) {
        if (p.residual) v[j] += __ldg(p.residual + roff + j);
        if (p.bias) v[j] += bias_s[j];
        v[j] = apply_act(v[j], p.act);
      }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (co_base + j < cout) {
        if (p.out_bf16)
          reinterpret_cast<__nv_bfloat16*>(p.out)[off + j] = __float2bfloat16_rn(v[j]);
        else
          reinterpret_cast<float*>(p.out)[off + j] = v[j];
      }
    }
  }
}

// Epilogue warps: TMEM -> registers -> (alpha, bias, residual, activation) -> global, one output pixel per thread and
// (for BN >= 64) one half of the tile's channels per warp.
// `stat_mem`: 16-byte aligned dynamic shared memory behind the barriers; holds an EpiStats<BN> when p.stats is set (the
// host adds its size to the launch only then: resident bandwidth-bound blocks of other streams need the space otherwise)
// STATS is a template parameter: the statistics code needs ~65 more registers per thread, and the register footprint of
// a convolution CTA decides how many bandwidth-bound blocks of the other streams run next to it.
template <int BN, bool STATS>
__device__ __forceinline__ void epilogue_loop(const IgemmParams& p, uint32_t tmem_base, uint64_t* tfull,
                                              uint64_t* tempty, int warp, int lane, void* stat_mem) {
  constexpr int ACC_STAGES = 2;
  constexpr int SPLIT = epi_split(BN);
  constexpr int EPI_T = epi_threads(BN);
  constexpr int CB = BN / SPLIT;          // channels drained by one warp
  constexpr int NC = CB < 32 ? CB : 32;   // columns per tcgen05.ld
  __shared__ __align__(16) float bias_s[ACC_STAGES][BN];
  EpiStats<(STATS ? BN : 1)>& stat_s = *static_cast<EpiStats<(STATS ? BN : 1)>*>(stat_mem);
  const int ew = warp - 2;                // epilogue warp 0..7
  const int half = ew >> 2;
  if (half >= SPLIT) return;              // narrow tiles: warps 6..9 have nothing to drain
  const int q = warp & 3;  // TMEM lane quadrant this warp may access
  const int m = q * 32 + lane;
  const int et = threadIdx.x - 64;        // 0..EPI_T-1 among the epilogue threads
  const int iw = m % p.bw;
  const int ih = (m / p.bw) % p.bh;
  const int in_ = m / (p.bw * p.bh);
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
  int as = 0;
  uint32_t aphase = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    int t = tile;
    const int tco = t % p.tiles_co; t /= p.tiles_co;
    const int tw = t % p.tiles_w;   t /= p.tiles_w;
    const int th = t % p.tiles_h;   t /= p.tiles_h;
    const int tn = t;
    const int wo = tw * p.bw + iw, ho = th * p.bh + ih, n = tn * p.bn + in_;
    const bool valid = (wo < p.Wo) && (ho < p.Ho) && (n < p.N);
    const int64_t pix = (static_cast<int64_t>(n) * p.Ho + ho) * p.Wo + wo;
    const int64_t res_pix =
        p.res_up2 ? (static_cast<int64_t>(n) * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1) : pix;
    if (p.bias) {  // stage this tile's bias slice; the two buffers alternate with the accumulator stage
      for (int c = et; c < BN; c += EPI_T) {
        const int co = tco * BN + c;
        bias_s[as][c] = co < p.Cout ? __ldg(p.bias + co) : 0.f;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_T) : "memory");
    mbar_wait(&tfull[as], aphase);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * CB;
#pragma unroll 1
    for (int c = 0; c < CB / NC; ++c) {
      const int col = half * CB + c * NC;   // first channel of this chunk inside the tile
      uint32_t r[NC];
      tmem_ld_cols<NC>(trow + c * NC, r);
      tmem_ld_wait();
      if constexpr (STATS) {
        float val[NC];
        if (valid)
          epilogue_row<NC>(p, r, alpha, pix * p.out_cstride, res_pix, tco * BN + col, &bias_s[as][col], val);
        epilogue_stats_chunk<NC, BN>(stat_s, val, valid, ew, q, lane, col);
      } else {
        if (valid) epilogue_row<NC>(p, r, alpha, pix * p.out_cstride, res_pix, tco * BN + col, &bias_s[as][col]);
      }
    }
    tc_fence_before();
    mbar_arrive(&tempty[as]);
    // one statistics row per pixel tile: tile / tiles_co is the image-major pixel-tile index
    if constexpr (STATS) epilogue_stats_flush<BN>(p, stat_s, tile / p.tiles_co, true, tco * BN, et);
    if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
  }
}

// TMA-store epilogue (shallow-K launches: K = 32 im2col routes, 1x1 shortcuts).  These layers are bound by their output
// stores: one thread per pixel writes 16 bytes per instruction to its own row, i.e. 32 partial sectors per warp store
// (2x the ideal L2 sector count, profiles/r02_ncu_igemm_k32.txt).  Here the epilogue warps write their rows into a
// shared-memory tile in the tensor map's 128-byte swizzle -- conflict-free: 4 wavefronts for 512 bytes -- and ONE bulk
// tensor store per 128-byte channel group and tile moves full rows; rows / channels outside the tensor are clipped by
// the TMA unit.  Two 16 KB tiles alternate; a tile is rewritten once the store that read it has finished reading
// (cp.async.bulk.wait_group.read).  No residual / statistics on this route.
constexpr int TS_TILE_BYTES = 128 * 128;

template <int BN>
__device__ __forceinline__ void epilogue_loop_tma(const IgemmParams& p, const CUtensorMap* tmO, uint32_t tmem_base,
                                                  uint64_t* tfull, uint64_t* tempty, int warp, int lane,
                                                  uint8_t* stage) {
  constexpr int ACC_STAGES = 2;
  constexpr int EPI_T = 128;
  constexpr int NC = 32;
  static_assert(BN % 64 == 0, "TMA-store epilogue: 64-channel groups");
  __shared__ __align__(16) float bias_s[ACC_STAGES][BN];
  const int q = warp & 3;
  const int m = q * 32 + lane;
  const int et = threadIdx.x - 64;
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
  const int group_ch = p.out_bf16 ? 64 : 32;            // channels of one 128-byte row
  const int groups = BN / group_ch;
  uint8_t* my_row[2] = {stage + m * 128, stage + TS_TILE_BYTES + m * 128};
  const uint32_t sw = static_cast<uint32_t>(m & 7);
  int as = 0, buf = 0;
  uint32_t aphase = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    int t = tile;
    const int tco = t % p.tiles_co; t /= p.tiles_co;
    const int tw = t % p.tiles_w;   t /= p.tiles_w;
    const int th = t % p.tiles_h;   t /= p.tiles_h;
    const int tn = t;
    if (p.bias) {
      for (int c = et; c < BN; c += EPI_T) {
        const int co = tco * BN + c;
        bias_s[as][c] = co < p.Cout ? __ldg(p.bias + co) : 0.f;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_T) : "memory");
    mbar_wait(&tfull[as], aphase);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
    for (int g = 0; g < groups; ++g) {
      if (tco * BN + g * group_ch >= p.Cout) break;       // whole group outside the tensor (uniform over the CTA)
      // the tile `buf` was read by the store issued two groups ago: wait until that read is done
      if (et == 0) bulk_wait_group_read<1>();
      asm volatile("bar.sync 2, %0;" ::"n"(EPI_T) : "memory");
      uint8_t* row = my_row[buf];
      const int chunks = p.out_bf16 ? 2 : 1;
#pragma unroll 1
      for (int cc = 0; cc < chunks; ++cc) {
        const int col = g * group_ch + cc * NC;
        uint32_t r[NC];
        tmem_ld_cols<NC>(trow + col, r);
        tmem_ld_wait();
        float v[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          float x = __uint_as_float(r[j]) * alpha;
          if (p.bias) x += bias_s[as][col + j];
          v[j] = apply_act(x, p.act);
        }
        if (p.out_bf16) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {       // 4 x 16 bytes = 32 bf16 channels: units cc*4 .. cc*4+3 of the row
            __nv_bfloat162 q0 = __floats2bfloat162_rn(v[8 * u], v[8 * u + 1]), q1 = __floats2bfloat162_rn(v[8 * u + 2], v[8 * u + 3]);
            __nv_bfloat162 q2 = __floats2bfloat162_rn(v[8 * u + 4], v[8 * u + 5]), q3 = __floats2bfloat162_rn(v[8 * u + 6], v[8 * u + 7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
            pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
            *reinterpret_cast<uint4*>(row + (((cc * 4 + u) ^ sw) << 4)) = pk;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u)         // 8 x 16 bytes = 32 fp32 channels
            *reinterpret_cast<float4*>(row + ((u ^ sw) << 4)) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
        }
      }
      fence_proxy_async();                    // generic-proxy writes visible to the bulk-copy engine
      asm volatile("bar.sync 2, %0;" ::"n"(EPI_T) : "memory");
      if (et == 0) {
        tma_store_4d(tmO, stage + buf * TS_TILE_BYTES, tco * BN + g * group_ch, tw * p.bw, th * p.bh, tn * p.bn);
        bulk_commit_group();
      }
      buf ^= 1;
    }
    tc_fence_before();
    mbar_arrive(&tempty[as]);
    if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
  }
  if (et == 0) bulk_wait_group<0>();          // every store has completed before the CTA exits
}

// TF32 = true: fp32 operands read by kind::tf32 MMAs (north star: "BF16 or TF32 inputs").  A 128-byte swizzle row then
// holds 32 channels instead of 64; a stage is still 128 rows x 128 bytes and four K-steps of 32 bytes, so only the
// channel coordinates of the TMA boxes, the instruction descriptor and the MMA kind differ.
template <int BN, int STAGES, bool STATS, bool TF32 = false, bool TSTORE = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const IgemmParams p, const __grid_constant__ CUtensorMap tmO) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int ACC_STAGES = 2;
  constexpr uint32_t TMEM_COLS = (ACC_STAGES * BN) < 32 ? 32 : (ACC_STAGES * BN);
  constexpr int NC = BN < 32 ? BN : 32;  // columns per tcgen05.ld
  constexpr uint32_t IDESC = TF32 ? umma_idesc_tf32(BM, BN, 0, 0) : umma_idesc_bf16(BM, BN, 0, 0);
  constexpr int BKE = TF32 ? 32 : BK;    // channels per 128-byte row

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], epi_threads(BN));
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue above overlaps the previous kernel's tail; global memory is touched only below

  const int kiters = p.taps * p.kchunks;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    const int kh = p.taps / p.kw;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int t = tile;
      const int tco = t % p.tiles_co; t /= p.tiles_co;
      const int tw = t % p.tiles_w;   t /= p.tiles_w;
      const int th = t % p.tiles_h;   t /= p.tiles_h;
      const int tn = t;
      const int w0 = tw * p.bw * p.stride - p.pad_l;
      const int h0 = th * p.bh * p.stride - p.pad_t;
      const int n0 = tn * p.bn;
      const int co0 = tco * BN;
      int tap = 0;
      for (int r = 0; r < kh; ++r) {
        for (int s2 = 0; s2 < p.kw; ++s2, ++tap) {
          const int tap_b = p.flip ? (p.taps - 1 - tap) : tap;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
              tma_load_4d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], kc * BKE, w0 + s2, h0 + r, n0);
              tma_load_3d(sB + stage * B_STAGE_BYTES, &tmB, &full[stage], kc * BKE, co0, tap_b);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (warp-uniform control flow keeps the descriptors in uniform registers and lets
    // the compiler emit one predicated UTCHMMA per MMA); one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    const uint64_t desc_base = umma_desc_base_sw128(16, 1024);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      uint32_t acc = 0;
      for (int kit = 0; kit < kiters; ++kit) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_desc_at(desc_base, sA_addr + stage * A_STAGE_BYTES);
          const uint64_t bdesc = umma_desc_at(desc_base, sB_addr + stage * B_STAGE_BYTES);
          // advance 16 bf16 (8 tf32) = 32 bytes along K inside the swizzle row: +2 in 16-byte units
          if constexpr (TF32) {
            umma_tf32(tmem_d, adesc, bdesc, IDESC, acc);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_tf32(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, 1u);
          } else {
            umma_bf16(tmem_d, adesc, bdesc, IDESC, acc);
#pragma unroll
            for (int k = 1; k < BK / 16; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, 1u);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        acc = 1u;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull[as]);
      __syncwarp();
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    if constexpr (TSTORE) {
      if (threadIdx.x == 64) tma_prefetch_desc(&tmO);
      // staging tiles: 1024-byte aligned, behind the barrier block
      const uint32_t a = smem_u32(tmem_slot + 4);
      uint8_t* stage = reinterpret_cast<uint8_t*>(tmem_slot + 4) + ((1024u - (a & 1023u)) & 1023u);
      epilogue_loop_tma<BN>(p, &tmO, tmem_base, tfull, tempty, warp, lane, stage);
    } else {
      epilogue_loop<BN, STATS>(p, tmem_base, tfull, tempty, warp, lane, tmem_slot + 4);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Halo variant for kh x kw > 1x1, stride 1: ONE TMA box of (bh+kh-1) x (bw+kw-1) input pixels x 64 channels is
// loaded per channel chunk and all kh*kw taps read it through shifted UMMA descriptors
//   start = halo + (r*(bw+kw-1) + s) * 128 B,  SBO = (bw+kw-1) * 128 B,  base_offset = 0
// which is valid because the 128B swizzle is a function of absolute shared-memory address bits
// (profiles/r01_halo_descriptor_experiment.log).  Per 128-pixel tile this cuts the activation bytes that cross
// L2->SMEM from taps x 16 KiB to ~23 KiB per 64 channels; the filter tiles travel through their own ring.
// Tile = 16 rows x 8 columns of one image (8-pixel row groups are what makes the descriptor regular).
constexpr int HALO_BW = 8, HALO_BH = 16;

constexpr int HALO_A_WARP = NUM_THREADS / 32;     // extra warp after the epilogue warps
constexpr int HALO_THREADS = NUM_THREADS + 32;

template <int BN, int SA, int SB, bool STATS>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const IgemmParams p, int a_stage_bytes, int halo_w, int halo_bytes) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int ACC_STAGES = 2;
  constexpr uint32_t TMEM_COLS = (ACC_STAGES * BN) < 32 ? 32 : (ACC_STAGES * BN);
  constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * a_stage_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + SB * B_STAGE_BYTES);
  uint64_t* emptyA = fullA + SA;
  uint64_t* fullB = emptyA + SA;
  uint64_t* emptyB = fullB + SB;
  uint64_t* tfull = emptyB + SB;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], epi_threads(BN)); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue above overlaps the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer: filter tiles (warp-uniform loop, elected lane issues) =====================
    int sb = 0;
    uint32_t pb = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int co0 = (tile % p.tiles_co) * BN;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        for (int tap = 0; tap < p.taps; ++tap) {
          const int tap_b = p.flip ? (p.taps - 1 - tap) : tap;
          mbar_wait(&emptyB[sb], pb ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&fullB[sb], B_STAGE_BYTES);
            tma_load_3d(sB + sb * B_STAGE_BYTES, &tmB, &fullB[sb], kc * BK, co0, tap_b);
          }
          __syncwarp();
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == HALO_A_WARP) {
    // ===================== TMA producer: activation halos =====================
    // Own warp so that halo requests run as far ahead as the SA-deep ring allows, independent of the filter ring.
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int t = tile / p.tiles_co;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&emptyA[sa], pa ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&fullA[sa], halo_bytes);
          tma_load_4d(sA + sa * a_stage_bytes, &tmA, &fullA[sa], kc * BK, tw * p.bw - p.pad_l, th * p.bh - p.pad_t, t);
        }
        __syncwarp();
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    int sa = 0, sb = 0, as = 0;
    uint32_t pa = 0, pb = 0, aphase = 0;
    const uint32_t sbo = static_cast<uint32_t>(halo_w) * 128u;
    const uint64_t adesc_base = umma_desc_base_sw128(16, sbo);
    const uint64_t bdesc_base = umma_desc_base_sw128(16, 1024);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    const int kw = p.kw, kh = p.taps / p.kw;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      uint32_t acc = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&fullA[sa], pa);
        uint32_t a_row = sA_addr + sa * a_stage_bytes;   // halo row r of this stage
        for (int r = 0; r < kh; ++r, a_row += sbo) {
          for (int s2 = 0; s2 < kw; ++s2) {
            mbar_wait(&fullB[sb], pb);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t adesc = umma_desc_at(adesc_base, a_row + s2 * 128);
              const uint64_t bdesc = umma_desc_at(bdesc_base, sB_addr + sb * B_STAGE_BYTES);
              umma_bf16(tmem_d, adesc, bdesc, IDESC, acc);
#pragma unroll
              for (int k = 1; k < BK / 16; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, 1u);
              umma_commit(&emptyB[sb]);
            }
            __syncwarp();
            acc = 1u;
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
        if (elect_one()) umma_commit(&emptyA[sa]);
        __syncwarp();
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull[as]);
      __syncwarp();
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  } else {
    epilogue_loop<BN, STATS>(p, tmem_base, tfull, tempty, warp, lane, tmem_slot + 4);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Narrow-output variant of the halo kernel (Cout <= BN <= 32, e.g. G.Output 256 -> 3): the whole packed filter
// (taps x kchunks tiles of BN x 64) is loaded into shared memory ONCE per CTA.  With a per-tap filter ring the 2 KiB
// filter loads queue behind the 23 KiB halo loads in the SM's TMA FIFO and the MMA thread waits a ring revolution per
// channel chunk (62 us for 128x32x32x256 -> 3, four times its activation-read time); resident, the loop has no
// filter barriers at all.
template <int BN, int SA>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_narrow_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const IgemmParams p, int a_stage_bytes, int halo_w, int halo_bytes) {
  constexpr int B_TILE_BYTES = BN * BK * 2;
  constexpr int ACC_STAGES = 2;
  constexpr uint32_t TMEM_COLS = (ACC_STAGES * BN) < 32 ? 32 : (ACC_STAGES * BN);
  constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int b_tiles = p.taps * p.kchunks;
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * a_stage_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + b_tiles * B_TILE_BYTES);
  uint64_t* emptyA = fullA + SA;
  uint64_t* fullB = emptyA + SA;
  uint64_t* tfull = fullB + 1;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    mbar_init(fullB, 1);
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], epi_threads(BN)); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== filter: one-time load of every (channel chunk, tap) tile =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(fullB, b_tiles * B_TILE_BYTES);
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int tap = 0; tap < p.taps; ++tap) {
          const int tap_b = p.flip ? (p.taps - 1 - tap) : tap;
          tma_load_3d(sB + (kc * p.taps + tap) * B_TILE_BYTES, &tmB, fullB, kc * BK, 0, tap_b);
        }
    }
    __syncwarp();
  } else if (warp == HALO_A_WARP) {
    // ===================== TMA producer: activation halos =====================
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int t = tile;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&emptyA[sa], pa ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&fullA[sa], halo_bytes);
          tma_load_4d(sA + sa * a_stage_bytes, &tmA, &fullA[sa], kc * BK, tw * p.bw - p.pad_l, th * p.bh - p.pad_t, t);
        }
        __syncwarp();
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int sa = 0, as = 0;
    uint32_t pa = 0, aphase = 0;
    const uint32_t sbo = static_cast<uint32_t>(halo_w) * 128u;
    const uint64_t adesc_base = umma_desc_base_sw128(16, sbo);
    const uint64_t bdesc_base = umma_desc_base_sw128(16, 1024);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    const int kw = p.kw, kh = p.taps / p.kw;
    mbar_wait(fullB, 0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      uint32_t acc = 0;
      uint32_t b_addr = sB_addr;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&fullA[sa], pa);
        tc_fence_after();
        if (elect_one()) {
          uint32_t a_row = sA_addr + sa * a_stage_bytes;
          for (int r = 0; r < kh; ++r, a_row += sbo) {
            for (int s2 = 0; s2 < kw; ++s2, b_addr += B_TILE_BYTES) {
              const uint64_t adesc = umma_desc_at(adesc_base, a_row + s2 * 128);
              const uint64_t bdesc = umma_desc_at(bdesc_base, b_addr);
              umma_bf16(tmem_d, adesc, bdesc, IDESC, acc);
#pragma unroll
              for (int k = 1; k < BK / 16; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, 1u);
              acc = 1u;
            }
          }
          umma_commit(&emptyA[sa]);
        }
        __syncwarp();
        b_addr = sB_addr + (kc + 1) * p.taps * B_TILE_BYTES;
        acc = 1u;
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull[as]);
      __syncwarp();
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  } else {
    epilogue_loop<BN, false>(p, tmem_base, tfull, tempty, warp, lane, tmem_slot + 4);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the halo kernel (tcgen05 cta_group::2, UMMA M = 256): a cluster of two CTAs owns two adjacent
// 16x8 pixel tiles and ONE filter tile of BN output channels.  Each CTA loads its own activation halo and the half of
// the filter tile with rows [rank*BN/2, (rank+1)*BN/2); the leader (rank 0) issues the MMAs for both SMs.  Filter
// bytes crossing L2->SMEM per SM are halved, which is what bounded the single-CTA kernel at ~60 % tensor-pipe
// utilisation (10 TB/s of TMA traffic = the chip's L2 throughput, profiles/r01_halo_ncu_summary.txt).
// Barriers: fullA / fullB live in the leader and count both CTAs' bytes; emptyA / emptyB / tfull are signalled in both
// CTAs by a multicast commit; tempty lives in the leader and collects one arrival per epilogue warp of both CTAs.
template <int BN, int SA, int SB, bool STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO_THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const IgemmParams p, int a_stage_bytes, int halo_w, int halo_bytes) {
  constexpr int BH = BN / 2;                    // filter rows held by one CTA
  constexpr int B_STAGE_BYTES = BH * BK * 2;
  constexpr int ACC_STAGES = 2;
  constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;
  constexpr int NC = 32;
  constexpr uint32_t IDESC = umma_idesc_bf16(2 * BM, BN, 0, 0);
  static_assert(BN == 128 || BN == 256, "pair kernel tiles 128 or 256 output channels");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * a_stage_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + SB * B_STAGE_BYTES);
  uint64_t* emptyA = fullA + SA;
  uint64_t* fullB = emptyA + SA;
  uint64_t* emptyB = fullB + SB;
  uint64_t* tfull = emptyB + SB;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int pair_tiles = ((m_tiles + 1) >> 1) * p.tiles_co * p.og;   // work items of a cluster
  const int tiles_cg = p.tiles_co * p.og;                            // (filter tile, output group) per pixel-tile pair

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * (epi_threads(BN) / 32)); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();   // the peer's barriers are initialised and its TMEM is allocated before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue above overlaps the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer: this CTA's half of the filter tile =====================
    int sb = 0;
    uint32_t pb = 0;
    for (int tile = pair; tile < pair_tiles; tile += num_pairs) {
      const int co0 = (tile % p.tiles_co) * BN + static_cast<int>(rank) * BH;
      const int go = (tile / p.tiles_co) % p.og;
      for (int gr = 0; gr < p.rg; ++gr) {
        const int tap_base = (go + gr) * p.taps;   // at most one of og / rg exceeds 1
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int tap = 0; tap < p.taps; ++tap) {
            const int tap_b = tap_base + (p.flip ? (p.taps - 1 - tap) : tap);
            mbar_wait(&emptyB[sb], pb ^ 1);
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(&fullB[sb], 2 * B_STAGE_BYTES);
              tma_load_3d_pair(sB + sb * B_STAGE_BYTES, &tmB, &fullB[sb], kc * BK, co0, tap_b);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == HALO_A_WARP) {
    // ===================== TMA producer: this CTA's activation halo =====================
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = pair; tile < pair_tiles; tile += num_pairs) {
      int t = 2 * (tile / tiles_cg) + static_cast<int>(rank);     // pixel tile of this CTA
      const int go = (tile / p.tiles_co) % p.og;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;                // t = image; past the batch -> TMA zero fill
      for (int gr = 0; gr < p.rg; ++gr) {
        const int g = go + gr;
        const int pad_t = (p.og * p.rg > 1) ? p.gpad_t[g] : p.pad_t, pad_l = (p.og * p.rg > 1) ? p.gpad_l[g] : p.pad_l;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&emptyA[sa], pa ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&fullA[sa], 2 * halo_bytes);
            tma_load_4d_pair(sA + sa * a_stage_bytes, &tmA, &fullA[sa], (gr * p.kchunks + kc) * BK, tw * p.bw - pad_l,
                             th * p.bh - pad_t, t);
          }
          __syncwarp();
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) =====================
      int sa = 0, sb = 0, as = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      const uint32_t sbo = static_cast<uint32_t>(halo_w) * 128u;
      const uint64_t adesc_base = umma_desc_base_sw128(16, sbo);
      const uint64_t bdesc_base = umma_desc_base_sw128(16, 1024);
      const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
      const int kw = p.kw, kh = p.taps / p.kw;
      for (int tile = pair; tile < pair_tiles; tile += num_pairs) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        uint32_t acc = 0;
        const int a_stages = p.rg * p.kchunks;   // one halo per (reduction group, channel chunk)
        for (int kc = 0; kc < a_stages; ++kc) {
          mbar_wait(&fullA[sa], pa);
          uint32_t a_row = sA_addr + sa * a_stage_bytes;
          for (int r = 0; r < kh; ++r, a_row += sbo) {
            for (int s2 = 0; s2 < kw; ++s2) {
              mbar_wait(&fullB[sb], pb);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t adesc = umma_desc_at(adesc_base, a_row + s2 * 128);
                const uint64_t bdesc = umma_desc_at(bdesc_base, sB_addr + sb * B_STAGE_BYTES);
                umma_bf16_pair(tmem_d, adesc, bdesc, IDESC, acc);
#pragma unroll
                for (int k = 1; k < BK / 16; ++k) umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, 1u);
                umma_commit_pair(&emptyB[sb]);
              }
              __syncwarp();
              acc = 1u;
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
          if (elect_one()) umma_commit_pair(&emptyA[sa]);
          __syncwarp();
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&tfull[as]);
        __syncwarp();
        if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs): own 128 pixels x BN channels =====================
    constexpr int EPI_T = epi_threads(BN);
    constexpr int CB = BN / epi_split(BN);  // channels drained by one warp
    __shared__ __align__(16) float bias_s[ACC_STAGES][BN];
    EpiStats<BN>& stat_s = *reinterpret_cast<EpiStats<BN>*>(tmem_slot + 4);   // dynamic, present when p.stats is set
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int iw = m % p.bw;
    const int ih = (m / p.bw) % p.bh;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const uint32_t tempty_leader = map_to_cta(smem_u32(&tempty[0]), 0);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = pair; tile < pair_tiles; tile += num_pairs) {
      const int tco = tile % p.tiles_co;
      const int go = (tile / p.tiles_co) % p.og;
      int t = 2 * (tile / tiles_cg) + static_cast<int>(rank);
      const int pix_tile = t;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int wo = tw * p.bw + iw, ho = th * p.bh + ih, n = t;
      const bool valid = (wo < p.Wo) && (ho < p.Ho) && (n < p.N);
      const int64_t pix = (static_cast<int64_t>(n) * p.Ho + ho) * p.Wo + wo;
      const int64_t res_pix =
          p.res_up2 ? (static_cast<int64_t>(n) * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1) : pix;
      if (p.bias) {
        for (int c = et; c < BN; c += EPI_T) {
          const int co = tco * BN + c;
          bias_s[as][c] = co < p.Cout ? __ldg(p.bias + co) : 0.f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_T) : "memory");
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * CB;
#pragma unroll 1
      for (int c = 0; c < CB / NC; ++c) {
        const int col = half * CB + c * NC;
        uint32_t r[NC];
        tmem_ld_cols<NC>(trow + c * NC, r);
        tmem_ld_wait();
        if constexpr (STATS) {
          float val[NC];
          if (valid)
            epilogue_row<NC>(p, r, alpha, pix * p.out_cstride + go * p.Cout, res_pix, tco * BN + col,
                             &bias_s[as][col], val);
          epilogue_stats_chunk<NC, BN>(stat_s, val, valid, ew, q, lane, col);
        } else {
          if (valid)
            epilogue_row<NC>(p, r, alpha, pix * p.out_cstride + go * p.Cout, res_pix, tco * BN + col, &bias_s[as][col]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + as * 8);
      if constexpr (STATS)   // the odd tile of the last pair lies past the batch (n >= N): it has no statistics row
        epilogue_stats_flush<BN>(p, stat_s, static_cast<int64_t>(pix_tile) * p.og + go, n < p.N, tco * BN, et);
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // both CTAs are done with both TMEMs and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[ci, co] = sum_pixels X[pixel + tap, ci] * dY[pixel, co]; channels are contiguous in NHWC so both
// operands are MN-major.  One CTA = (tap, ci tile of 128, co tile of BN, pixel split).
// ------------------------------------------------------------------------------------------------
constexpr int PB = 64;  // pixels per pipeline stage (4 UMMA K-steps of 16)
#ifndef GANB_TF32_SBO
#define GANB_TF32_SBO 512
#endif

struct WgradParams {
  int N, Ho, Wo;  // dy spatial extent
  int Cin, Cout;
  int kw, taps, pad_t, pad_l, stride;
  int bw, bh, bn;          // pixel box, bw*bh*bn == PB
  int pb_w, pb_h, pb_n;    // pixel blocks per dim
  int num_pb, splits, pb_per_split;
  int ci_tiles, co_tiles;
  float* partial;  // [splits][taps][Cin][Cout]
  // sub-pixel UpsampleConv (ganb_upconv_wgrad): taps = 4 parities x (2 x 2); parity g = 2i + j reads the channel slice
  // [g*Cout, (g+1)*Cout) of the quad-layout gradient [N, Ho, Wo, 4*Cout] and pads its 2x2 window by (1 - i, 1 - j)
  int quad;
};

// TF32 = true: fp32 operands (32 channels per 128-byte row: BM / 32 boxes per operand tile), K-steps of 8 pixels.
template <int BN, int STAGES, bool TF32 = false>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const WgradParams p) {
  constexpr int ESZ = TF32 ? 4 : 2;
  constexpr int BOXC = 128 / ESZ;       // channels of one 128-byte-wide box
  constexpr int A_BYTES = PB * BM * ESZ;  // BM / BOXC boxes of PB pixel rows
  constexpr int B_BYTES = PB * BN * ESZ;
  constexpr int BOX_BYTES = PB * 128;   // one box
  constexpr int KSTEP = TF32 ? 8 : 16;  // pixels per MMA
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = TF32 ? umma_idesc_tf32(BM, BN, 1, 1) : umma_idesc_bf16(BM, BN, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue above overlaps the previous kernel's tail; global memory is touched only below

  int u = blockIdx.x;
  const int tco = u % p.co_tiles; u /= p.co_tiles;
  const int tci = u % p.ci_tiles; u /= p.ci_tiles;
  const int tap = u % p.taps;     u /= p.taps;
  const int split = u;
  const int ltap = p.quad ? (tap & 3) : tap;                 // tap inside its parity group
  const int grp = p.quad ? (tap >> 2) : 0;
  const int r = ltap / p.kw, s = ltap - r * p.kw;
  const int pad_t = p.quad ? 1 - (grp >> 1) : p.pad_t, pad_l = p.quad ? 1 - (grp & 1) : p.pad_l;
  const int pb_begin = split * p.pb_per_split;
  const int pb_end = min(p.num_pb, pb_begin + p.pb_per_split);
  const int ci0 = tci * BM, co0 = tco * BN;
  const int dy_c0 = grp * p.Cout + co0;                     // channel coordinate inside the dy tensor map

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    int bwi = pb_begin % p.pb_w, bhi = (pb_begin / p.pb_w) % p.pb_h, bni = pb_begin / (p.pb_w * p.pb_h);
    for (int pb = pb_begin; pb < pb_end; ++pb) {
      const int w0 = bwi * p.bw, h0 = bhi * p.bh, n0 = bni * p.bn;
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
        uint8_t* a = sA + stage * A_BYTES;
        uint8_t* b = sB + stage * B_BYTES;
#pragma unroll
        for (int j = 0; j < BM / BOXC; ++j)
          tma_load_4d(a + j * BOX_BYTES, &tmX, &full[stage], ci0 + BOXC * j, w0 * p.stride + s - pad_l,
                      h0 * p.stride + r - pad_t, n0);
#pragma unroll
        for (int j = 0; j < BN / BOXC; ++j)
          tma_load_4d(b + j * BOX_BYTES, &tmDY, &full[stage], dy_c0 + BOXC * j, w0, h0, n0);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
      if (++bwi == p.pb_w) { bwi = 0; if (++bhi == p.pb_h) { bhi = 0; ++bni; } }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc = 0;
    // MN-major, 128B swizzle: 64 channels per row, 8-pixel groups 1024 B apart (SBO),
    // next 64-channel box BOX_BYTES away (LBO).
    // TF32: MN-major fp32 operands exist only in the "128-byte swizzle on a 32-byte base" layout (4-row pattern: SBO = 512)
    const uint64_t desc_base = TF32 ? umma_desc_base_sw128_base32(BOX_BYTES, GANB_TF32_SBO) : umma_desc_base_sw128(BOX_BYTES, 1024);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    for (int pb = pb_begin; pb < pb_end; ++pb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = umma_desc_at(desc_base, sA_addr + stage * A_BYTES);
        const uint64_t bdesc = umma_desc_at(desc_base, sB_addr + stage * B_BYTES);
        // 16 (tf32: 8) pixels = rows of 128 B: +128 (+64) in 16-byte units per K-step
        if constexpr (TF32) {
          umma_tf32(tmem_base, adesc, bdesc, IDESC, acc);
#pragma unroll
          for (int k = 1; k < PB / KSTEP; ++k) umma_tf32(tmem_base, adesc + 64 * k, bdesc + 64 * k, IDESC, 1u);
        } else {
          umma_bf16(tmem_base, adesc, bdesc, IDESC, acc);
#pragma unroll
          for (int k = 1; k < PB / KSTEP; ++k) umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, IDESC, 1u);
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      acc = 1u;
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int ci = ci0 + q * 32 + lane;
    float* out = p.partial + ((static_cast<int64_t>(split) * p.taps + tap) * p.Cin + ci) * p.Cout;
    const bool has_work = pb_end > pb_begin;
    if (has_work) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      if (has_work) {
        tmem_ld_32x32(trow + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (ci < p.Cin) {
        if ((p.Cout & 7) == 0) {
          // groups of 8 never straddle the edge and every row of the partials starts on a 32-byte boundary: one whole
          // sector per store instruction (the rows of a warp are Cout * 4 bytes apart)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int co = co0 + c * 32 + j;
            if (co < p.Cout) stg32(out + co, v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
          }
        } else {      // TF32 operand mode admits Cout % 4 == 0
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int co = co0 + c * 32 + j;
            if (co < p.Cout)
              *reinterpret_cast<float4*>(out + co) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dw = beta*dw + scale * sum_s partial[s].  A block owns 256 / G consecutive float4 outputs; its G thread groups take the
// splits g, g + G, ... and their sums meet in shared memory in a fixed order (deterministic).  Small filters with many
// splits (the RGB-side layers: 4096 outputs x 148 splits) were latency chains of `splits` dependent-issue loads per thread.
template <int G>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                            int64_t n4, int splits, const float* __restrict__ scale,
                                                            float beta) {
  pdl_wait();
  constexpr int OUT = 256 / G;
  __shared__ float4 sm[G > 1 ? G : 1][OUT];
  const int o = threadIdx.x % OUT, g = threadIdx.x / OUT;
  const float sc = scale ? __ldg(scale) : 1.0f;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * OUT; base < n4; base += static_cast<int64_t>(gridDim.x) * OUT) {
    const int64_t i = base + o;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
#pragma unroll 4
      for (int s = g; s < splits; s += G) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(partial) + s * n4 + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (G > 1) {
      sm[g][o] = acc;
      __syncthreads();
      if (g == 0) {
#pragma unroll
        for (int k = 1; k < G; ++k) {
          const float4 v = sm[k][o];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    if (g == 0 && i < n4) {
      float4 r = make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc);
      if (beta != 0.f) {
        const float4 d = reinterpret_cast<const float4*>(dw)[i];
        r.x += beta * d.x; r.y += beta * d.y; r.z += beta * d.z; r.w += beta * d.w;
      }
      reinterpret_cast<float4*>(dw)[i] = r;
    }
    if (G > 1) __syncthreads();
  }
}

static cudaError_t launch_splitk_reduce(const float* partial, float* dw, int64_t n4, int splits, const float* scale,
                                        float beta, cudaStream_t stream) {
  // enough thread groups per output to put ~2 blocks on every SM, never more groups than splits
  const int64_t want = 2LL * tc_sm_count() * 256;
  int g = 1;
  while (g < 32 && n4 * g * 2 <= want && g * 2 <= splits) g *= 2;
  if (g == 2) g = 1;
  if (g == 16) g = 8;
  const int out = 256 / g;
  int blocks = static_cast<int>(ceil_div64(n4, out));
  if (blocks > 4 * tc_sm_count()) blocks = 4 * tc_sm_count();
  switch (g) {
    case 32: return launch_k(splitk_reduce_kernel<32>, blocks, 256, 0, stream, partial, dw, n4, splits, scale, beta);
    case 8: return launch_k(splitk_reduce_kernel<8>, blocks, 256, 0, stream, partial, dw, n4, splits, scale, beta);
    case 4: return launch_k(splitk_reduce_kernel<4>, blocks, 256, 0, stream, partial, dw, n4, splits, scale, beta);
    default: return launch_k(splitk_reduce_kernel<1>, blocks, 256, 0, stream, partial, dw, n4, splits, scale, beta);
  }
}

// ------------------------------------------------------------------------------------------------ host
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Splits `total` (a power of two) pixels into a bn x bh x bw box that tiles an image of size H x W.
static void pick_box(int total, int H, int W, int* bw, int* bh, int* bn) {
  int w = pow2_ceil(W);
  if (w > total) w = total;
  int h = pow2_ceil(H);
  if (h > total / w) h = total / w;
  *bw = w;
  *bh = h;
  *bn = total / (w * h);
}

// dynamic shared memory of the fused-statistics buffers (only when the launch produces them)
template <int BN>
static int stats_smem(const IgemmParams& p) {
  return (p.stats && BN >= 64) ? static_cast<int>(sizeof(EpiStats<BN>)) + 16 : 0;
}

// `ctas_per_sm`: shallow-K launches (<= 4 k-iterations per tile: K = 32 im2col routes, 1x1 shortcuts) are bound by the
// epilogue -- one warp per scheduler, ~10 cycles per issued instruction (profiles/r02_ncu_igemm_k32.txt) -- not by the
// tensor pipe.  A 2-stage instance needs 64 KB of shared memory and 2 x BN <= 256 TMEM columns, so TWO CTAs fit on an
// SM and their epilogues run side by side.
template <int BN, int STAGES>
static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, IgemmParams& p, cudaStream_t stream,
                        int ctas_per_sm = 1, const CUtensorMap* tmO = nullptr) {
  int smem = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 + 256 + stats_smem<BN>(p);
  auto kern = conv_igemm_kernel<BN, STAGES, false>;
  int variant = 0;
  if constexpr (BN >= 64) {
    if (p.stats) { kern = conv_igemm_kernel<BN, STAGES, true>; variant = 1; }
  }
  if constexpr (STAGES == 2 && BN % 64 == 0) {     // TMA-store epilogue: shallow-K instances only
    if (tmO && !p.stats) {
      kern = conv_igemm_kernel<BN, STAGES, false, false, true>;
      variant = 2;
      smem += 1024 + 2 * TS_TILE_BYTES;
    }
  }
  static int configured[3] = {0, 0, 0};
  if (configured[variant] < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "igemm smem attribute: %s", cudaGetErrorString(e));
    configured[variant] = smem;
  }
  p.tiles_co = ceil_div(p.Cout, BN);
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_co;
  const int slots = tc_sm_count() * ctas_per_sm;
  int grid = p.num_tiles < slots ? p.num_tiles : slots;
  launch_k(kern, grid, NUM_THREADS, smem, stream, tmA, tmB, p, variant == 2 ? *tmO : tmA);
  GANB_CHECK_LAUNCH("conv_igemm_kernel");
  return 0;
}

template <int BN, int SA, int SB>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, IgemmParams& p, int a_stage_bytes, int halo_w,
                       int halo_bytes, cudaStream_t stream) {
  const int smem = SA * a_stage_bytes + SB * BN * BK * 2 + 1024 + 512 + stats_smem<BN>(p);
  if (smem + 8 * BN + 64 > 232448) return fail(GANB_E_UNSUPPORTED, "conv halo kernel: %d bytes of shared memory needed", smem);
  auto kern = conv_halo_kernel<BN, SA, SB, false>;
  bool with_stats = false;
  if constexpr (BN >= 64) {
    if (p.stats) { kern = conv_halo_kernel<BN, SA, SB, true>; with_stats = true; }
  }
  static int configured[2] = {0, 0};
  if (configured[with_stats] < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "halo smem attribute: %s", cudaGetErrorString(e));
    configured[with_stats] = smem;
  }
  p.tiles_co = ceil_div(p.Cout, BN);
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_co;
  int grid = p.num_tiles < tc_sm_count() ? p.num_tiles : tc_sm_count();
  launch_k(kern, grid, HALO_THREADS, smem, stream, tmA, tmB, p, a_stage_bytes, halo_w, halo_bytes);
  GANB_CHECK_LAUNCH("conv_halo_kernel");
  return 0;
}

template <int BN, int SA>
static int launch_halo_narrow(const CUtensorMap& tmA, const CUtensorMap& tmB, IgemmParams& p, int a_stage_bytes,
                              int halo_w, int halo_bytes, cudaStream_t stream) {
  const int smem = SA * a_stage_bytes + p.taps * p.kchunks * BN * BK * 2 + 1024 + 512;
  auto kern = conv_halo_narrow_kernel<BN, SA>;
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "narrow halo smem attribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  p.tiles_co = 1;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  int grid = p.num_tiles < tc_sm_count() ? p.num_tiles : tc_sm_count();
  launch_k(kern, grid, HALO_THREADS, smem, stream, tmA, tmB, p, a_stage_bytes, halo_w, halo_bytes);
  GANB_CHECK_LAUNCH("conv_halo_narrow_kernel");
  return 0;
}

// GANB_PAIR=0 disables the cta_group::2 kernels (A/B comparison of the two code paths)
static bool pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("GANB_PAIR");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode == 1;
}

template <int BN, int SA, int SB>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, IgemmParams& p, int a_stage_bytes, int halo_w,
                       int halo_bytes, cudaStream_t stream) {
  const int smem = SA * a_stage_bytes + SB * (BN / 2) * BK * 2 + 1024 + 512 + stats_smem<BN>(p);
  if (smem + 8 * BN + 64 > 232448) return fail(GANB_E_UNSUPPORTED, "conv pair kernel: %d bytes of shared memory needed", smem);
  auto kern = p.stats ? conv_pair_kernel<BN, SA, SB, true> : conv_pair_kernel<BN, SA, SB, false>;
  const bool with_stats = p.stats != nullptr;
  static int configured[2] = {0, 0};
  if (configured[with_stats] < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "pair smem attribute: %s", cudaGetErrorString(e));
    configured[with_stats] = smem;
  }
  p.tiles_co = ceil_div(p.Cout, BN);
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int pair_tiles = ((m_tiles + 1) / 2) * p.tiles_co * p.og;
  p.num_tiles = pair_tiles;
  int clusters = tc_sm_count() / 2;
  if (clusters > pair_tiles) clusters = pair_tiles;
  launch_k(kern, 2 * clusters, HALO_THREADS, smem, stream, tmA, tmB, p, a_stage_bytes, halo_w, halo_bytes);
  GANB_CHECK_LAUNCH("conv_pair_kernel");
  return 0;
}

}  // namespace ganb

using namespace ganb;

namespace ganb {
// pixel tiling of ganb_conv2d_igemm: halo tiles (16 x 8 pixels of one image) or a 128-pixel box over (n, h, w)
static bool igemm_tiling(int ho, int wo, int kh, int kw, int stride, int* bw, int* bh, int* bn) {
  const bool halo = (kh * kw > 1) && stride == 1 && ho >= HALO_BH && wo >= HALO_BW && (kh + HALO_BH - 1) <= 256;
  if (halo) {
    *bw = HALO_BW; *bh = HALO_BH; *bn = 1;
  } else {
    pick_box(BM, ho, wo, bw, bh, bn);
  }
  return halo;
}
}  // namespace ganb

static int conv2d_igemm_impl(const void* x, const void* wp, void* y, int n, int h, int w, int cin, int ho, int wo,
                             int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps,
                             const float* alpha, const float* bias, const float* residual, int residual_up2, int act,
                             int out_dtype, float* stats, const void* gate, int gate_act, void* stream_);

extern "C" int ganb_conv2d_igemm(const void* x, const void* wp, void* y, int n, int h, int w, int cin, int ho,
                                 int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l,
                                 int flip_taps, const float* alpha, const float* bias, const float* residual,
                                 int residual_up2, int act, int out_dtype, void* stream_) {
  return conv2d_igemm_impl(x, wp, y, n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l, flip_taps, alpha, bias,
                           residual, residual_up2, act, out_dtype, nullptr, nullptr, 0, stream_);
}

extern "C" int ganb_conv2d_igemm_gated(const void* x, const void* wp, void* y, int n, int h, int w, int cin, int ho,
                                       int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l,
                                       int flip_taps, const float* alpha, const void* gate_bf16, int gate_act,
                                       int out_dtype, void* stream_) {
  if (!gate_bf16 || (gate_act != GANB_ACT_RELU && gate_act != GANB_ACT_LRELU))
    return fail(GANB_E_BADARG, "conv2d_igemm_gated: gate=%p gate_act=%d (relu / leaky relu)", gate_bf16, gate_act);
  if (cout % 8 != 0 || (reinterpret_cast<uintptr_t>(gate_bf16) & 15))
    return fail(GANB_E_UNSUPPORTED, "conv2d_igemm_gated: cout=%d must be a multiple of 8 and the gate 16-byte aligned", cout);
  return conv2d_igemm_impl(x, wp, y, n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l, flip_taps, alpha, nullptr,
                           nullptr, 0, GANB_ACT_NONE, out_dtype, nullptr, gate_bf16, gate_act, stream_);
}

extern "C" int ganb_conv2d_stats_rows(int n, int ho, int wo, int cout, int kh, int kw, int stride, int groups) {
  if (n <= 0 || ho <= 0 || wo <= 0 || groups <= 0 || cout < 64 || cout % 32 != 0 || n % groups != 0) return 0;
  int bw, bh, bn;
  igemm_tiling(ho, wo, kh, kw, stride, &bw, &bh, &bn);
  if ((n / groups) % bn != 0) return 0;          // a pixel tile would straddle two statistic towers
  return ceil_div(wo, bw) * ceil_div(ho, bh) * ((n / groups) / bn);
}

extern "C" int ganb_conv2d_igemm_stats(const void* x, const void* wp, void* y, int n, int h, int w, int cin, int ho,
                                       int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l,
                                       int flip_taps, const float* alpha, const float* bias, const float* residual,
                                       int residual_up2, int act, int out_dtype, float* stats, int groups,
                                       void* stream_) {
  if (!stats) return fail(GANB_E_BADARG, "conv2d_igemm_stats: null statistics buffer");
  if (!ganb_conv2d_stats_rows(n, ho, wo, cout, kh, kw, stride, groups))
    return fail(GANB_E_UNSUPPORTED, "conv2d_igemm_stats: n=%d ho=%d wo=%d cout=%d groups=%d (see ganb_conv2d_stats_rows)",
                n, ho, wo, cout, groups);
  return conv2d_igemm_impl(x, wp, y, n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l, flip_taps, alpha, bias,
                           residual, residual_up2, act, out_dtype, stats, nullptr, 0, stream_);
}

static int conv2d_igemm_impl(const void* x, const void* wp, void* y, int n, int h, int w, int cin, int ho, int wo,
                             int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps,
                             const float* alpha, const float* bias, const float* residual, int residual_up2, int act,
                             int out_dtype, float* stats, const void* gate, int gate_act, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !wp || !y) return fail(GANB_E_BADARG, "conv2d_igemm: null buffer");
  if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || ho <= 0 || wo <= 0 || cout <= 0 || kh <= 0 || kw <= 0)
    return fail(GANB_E_BADARG, "conv2d_igemm: non-positive dimension");
  if (cin % 8 != 0) return fail(GANB_E_UNSUPPORTED, "conv2d_igemm: cin=%d must be a multiple of 8", cin);
  if (stride < 1 || stride > 4) return fail(GANB_E_UNSUPPORTED, "conv2d_igemm: stride=%d (1..4 supported)", stride);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wp)) & 15)
    return fail(GANB_E_BADARG, "conv2d_igemm: x / wp must be 16-byte aligned");

  IgemmParams p;
  p.N = n; p.Ho = ho; p.Wo = wo; p.Cout = cout;
  p.taps = kh * kw; p.kw = kw;
  p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
  // halo path: every tap reads one shared-memory halo tile (needs 8-pixel row groups: 16 x 8 tiles per image)
  // strided convolutions gather every stride-th pixel through the TMA element strides of the per-tap box
  const bool halo = igemm_tiling(ho, wo, kh, kw, stride, &p.bw, &p.bh, &p.bn);
  p.stats = stats;
  p.gate = static_cast<const __nv_bfloat16*>(gate); p.gate_act = gate_act;
  p.tiles_w = ceil_div(wo, p.bw);
  p.tiles_h = ceil_div(ho, p.bh);
  p.tiles_n = ceil_div(n, p.bn);
  p.kchunks = ceil_div(cin, BK);
  p.flip = flip_taps;
  p.alpha = alpha; p.bias = bias; p.residual = residual;
  p.res_up2 = (residual && residual_up2) ? 1 : 0;
  if (p.res_up2 && ((ho | wo) & 1)) return fail(GANB_E_BADARG, "conv2d_igemm: upsampled residual needs even ho, wo");
  p.out = y; p.out_bf16 = (out_dtype == GANB_BF16); p.act = act;
  p.og = 1; p.rg = 1; p.out_cstride = cout;
  {
    const int64_t es_ = p.out_bf16 ? 2 : 4;
    static const bool st32_ok = !(getenv("GANB_ST32") && getenv("GANB_ST32")[0] == '0');   // A/B switch
    p.st32 = st32_ok && (reinterpret_cast<uintptr_t>(p.out) % 32 == 0) && ((p.Cout * es_) % 32 == 0) &&
             ((static_cast<int64_t>(p.out_cstride) * es_) % 32 == 0);
    p.res32 = st32_ok && p.residual && (reinterpret_cast<uintptr_t>(p.residual) % 32 == 0) && (p.Cout % 8 == 0);
  }
  for (int i = 0; i < 4; ++i) { p.gpad_t[i] = 0; p.gpad_l[i] = 0; }

  // choose the N tile with a two-term cost model (cycles): tensor time = waves x k-iterations x MMA cycles of a tile,
  // operand time = bytes crossing L2->SMEM / what the chip (~6500 B/clk measured) or the active SMs (~64 B/clk
  // each) can ingest.  Small-M, deep-K layers (8x8 images, K = 9216) want FEW LARGE tiles even when that leaves SMs
  // idle: halving BN doubles the activation re-reads.
  int bn_tile = cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : cout <= 128 ? 128 : 256;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  if (bn_tile > 64) {
    const double kit = static_cast<double>(kh) * kw * p.kchunks;
    const double a_bytes = halo ? 16384.0 * 1.5 / (kh * kw) : 16384.0;   // per k-iteration of one tile
    double best = 0;
    int best_bn = bn_tile;
    for (int bn = bn_tile; bn >= 64; bn >>= 1) {
      const double tiles = static_cast<double>(m_tiles) * ceil_div(cout, bn);
      const double active = tiles < tc_sm_count() ? tiles : tc_sm_count();
      const double waves = ceil_div(static_cast<int>(tiles), tc_sm_count());
      const double t_mma = waves * kit * 4.0 * (bn / 2.0);
      const double rate = 6500.0 < 64.0 * active ? 6500.0 : 64.0 * active;
      const double t_l2 = tiles * kit * (a_bytes + bn * 128.0) / rate;
      const double t = (t_mma > t_l2 ? t_mma : t_l2) + 3000.0 * waves;     // + per-tile epilogue / ramp
      if (best == 0 || t < best) { best = t; best_bn = bn; }
    }
    bn_tile = best_bn;
  }

  // CTA pairs (cta_group::2) halve the filter traffic per SM: used when there is at least one wave of pair tiles
  const int pair_bn = cout > 128 ? 256 : 128;
  const bool pair = halo && cout >= 128 && pair_mode() &&
                    ((m_tiles + 1) / 2) * ceil_div(cout, pair_bn) >= tc_sm_count() / 2;
  if (pair) bn_tile = pair_bn;

  // shallow-K launches (see launch_igemm): two CTAs per SM; BN = 256 would take all 512 TMEM columns, so the channels
  // are split over 128-wide tiles.  GANB_SHALLOW=0 disables the route (A/B).
  static const bool shallow_ok = !(getenv("GANB_SHALLOW") && getenv("GANB_SHALLOW")[0] == '0');
  const bool shallow = shallow_ok && !halo && !stats && p.taps * p.kchunks <= 4 && bn_tile >= 64;
  if (shallow && bn_tile == 256) bn_tile = 128;

  CUtensorMap tmA, tmB;
  const int halo_w = p.bw + kw - 1, halo_h = p.bh + kh - 1;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cin * 2, (uint64_t)w * cin * 2, (uint64_t)h * w * cin * 2};
    // with element strides the box is given as the traversed extent: (b - 1) * stride + 1 input pixels -> b loaded
    const uint32_t box[4] = {BK, (uint32_t)(halo ? halo_w : (p.bw - 1) * stride + 1),
                             (uint32_t)(halo ? halo_h : (p.bh - 1) * stride + 1), (uint32_t)p.bn};
    const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (box[1] > 256 || box[2] > 256) return fail(GANB_E_UNSUPPORTED, "conv2d_igemm: tile extent %u x %u exceeds the TMA box limit", box[1], box[2]);
    int rc = encode_tmap_bf16(&tmA, x, 4, dims, strides, box, stride > 1 ? estr : nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cout, (uint64_t)(kh * kw)};
    const uint64_t strides[2] = {(uint64_t)cin * 2, (uint64_t)cin * cout * 2};
    const uint32_t box[3] = {BK, (uint32_t)(pair ? bn_tile / 2 : bn_tile), 1};
    int rc = encode_tmap_bf16(&tmB, wp, 3, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  if (halo) {
    const int halo_bytes = halo_w * halo_h * 128;
    const int a_stage = (halo_bytes + 1023) / 1024 * 1024;
    if (pair) {
      if (pair_bn == 256) return launch_pair<256, 3, ST_PAIR256_SB>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
      return launch_pair<128, 4, ST_PAIR128_SB>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
    }
    // narrow outputs: resident filter when it fits next to four halo stages
    if (bn_tile == 16 && cout <= 16 && 4 * a_stage + p.taps * p.kchunks * 16 * BK * 2 + 2048 <= 232448)
      return launch_halo_narrow<16, 4>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
    switch (bn_tile) {
      case 16: return launch_halo<16, 7, 9>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
      case 32: return launch_halo<32, 6, 9>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
      case 64: return launch_halo<64, 5, ST_HALO64_SB>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
      case 128: return launch_halo<128, 3, ST_HALO128_SB>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
      default: return launch_halo<256, 2, ST_HALO256_SB>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
    }
  }
  if (shallow) {
    // TMA-store epilogue where the output is a plain tensor of full 16-byte rows.  OPT-IN (GANB_TMA_STORE=1): correct
    // (the whole GPU suite passes with it on) but measured 1.4-2x SLOWER than the per-thread 16-byte stores on every
    // shallow-K layer (21.0 -> 39.9 us, 32.8 -> 68.8 us, 17.3 -> 24.3 us; step 3.14 -> 3.30 ms,
    // profiles/r02_tma_store_ab.txt) -- like round 1's shared-memory-transposed epilogue: two named barriers, a proxy fence
    // and a bulk-group wait per 128-byte channel group serialise the four epilogue warps.
    static const bool tstore_ok = getenv("GANB_TMA_STORE") && getenv("GANB_TMA_STORE")[0] == '1';
    const int es = p.out_bf16 ? 2 : 4, group_ch = p.out_bf16 ? 64 : 32;
    CUtensorMap tmO;
    const CUtensorMap* tmo = nullptr;
    if (tstore_ok && !residual && !gate && (static_cast<int64_t>(cout) * es) % 16 == 0 && cout >= group_ch &&
        (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
      const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)wo, (uint64_t)ho, (uint64_t)n};
      const uint64_t strides[3] = {(uint64_t)cout * es, (uint64_t)wo * cout * es, (uint64_t)ho * wo * cout * es};
      const uint32_t box[4] = {(uint32_t)group_ch, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
      int rc = p.out_bf16 ? encode_tmap_bf16(&tmO, y, 4, dims, strides, box, nullptr)
                          : encode_tmap_f32(&tmO, y, 4, dims, strides, box, nullptr);
      if (rc) return rc;
      tmo = &tmO;
    }
    if (bn_tile == 64) return launch_igemm<64, 2>(tmA, tmB, p, stream, 2, tmo);
    return launch_igemm<128, 2>(tmA, tmB, p, stream, 2, tmo);
  }
  switch (bn_tile) {
    case 16: return launch_igemm<16, 8>(tmA, tmB, p, stream);
    case 32: return launch_igemm<32, 8>(tmA, tmB, p, stream);
    case 64: return launch_igemm<64, ST_IG64>(tmA, tmB, p, stream);
    case 128: return launch_igemm<128, ST_IG128>(tmA, tmB, p, stream);
    default: return launch_igemm<256, ST_IG256>(tmA, tmB, p, stream);
  }
}

namespace ganb {

struct WgradPlan {
  WgradParams p;
  int bn_tile;
  int grid;
};

static void plan_wgrad(int n, int ho, int wo, int cin, int cout, int kh, int kw, WgradPlan* plan, int max_bn = 256) {
  WgradParams& p = plan->p;
  p.N = n; p.Ho = ho; p.Wo = wo; p.Cin = cin; p.Cout = cout;
  p.kw = kw; p.taps = kh * kw;
  pick_box(PB, ho, wo, &p.bw, &p.bh, &p.bn);
  p.pb_w = ceil_div(wo, p.bw);
  p.pb_h = ceil_div(ho, p.bh);
  p.pb_n = ceil_div(n, p.bn);
  p.num_pb = p.pb_w * p.pb_h * p.pb_n;
  plan->bn_tile = cout <= 64 ? 64 : cout <= 128 ? 128 : 256;
  if (plan->bn_tile > max_bn) plan->bn_tile = max_bn;
  p.ci_tiles = ceil_div(cin, BM);
  p.co_tiles = ceil_div(cout, plan->bn_tile);
  const int units = p.taps * p.ci_tiles * p.co_tiles;
  // one CTA per SM and a single wave: a second, nearly empty wave would idle most of the machine
  int splits = tc_sm_count() / units;
  if (splits < 1) splits = 1;
  // keep at least 4 pixel blocks per split so the pipeline has something to overlap
  const int max_splits = p.num_pb / 4 > 0 ? p.num_pb / 4 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits > 64) splits = 64;
  p.pb_per_split = ceil_div(p.num_pb, splits);
  p.splits = ceil_div(p.num_pb, p.pb_per_split);
  plan->grid = units * p.splits;
}

template <int BN, int STAGES>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradParams& p, int grid,
                        cudaStream_t stream) {
  constexpr int smem = STAGES * (PB * BM * 2 + PB * BN * 2) + 1024 + 256;
  static bool configured = false;
  auto kern = conv_wgrad_kernel<BN, STAGES>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "wgrad smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  launch_k(kern, grid, WG_THREADS, smem, stream, tmX, tmDY, p);
  GANB_CHECK_LAUNCH("conv_wgrad_kernel");
  return 0;
}

}  // namespace ganb

extern "C" int64_t ganb_conv2d_wgrad_workspace(int n, int h, int w, int cin, int ho, int wo, int cout, int kh,
                                               int kw) {
  (void)h; (void)w;
  WgradPlan plan;
  plan_wgrad(n, ho, wo, cin, cout, kh, kw, &plan);
  return static_cast<int64_t>(plan.p.splits) * kh * kw * cin * cout * 4;
}

extern "C" int ganb_conv2d_wgrad(const void* x, const void* dy, float* dw, void* workspace, int n, int h, int w,
                                 int cin, int ho, int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l,
                                 const float* scale, float beta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !dy || !dw || !workspace) return fail(GANB_E_BADARG, "conv2d_wgrad: null buffer");
  if (cin % 8 != 0 || cout % 8 != 0)
    return fail(GANB_E_UNSUPPORTED, "conv2d_wgrad: cin=%d and cout=%d must be multiples of 8", cin, cout);
  WgradPlan plan;
  plan_wgrad(n, ho, wo, cin, cout, kh, kw, &plan);
  if (stride < 1 || stride > 4) return fail(GANB_E_UNSUPPORTED, "conv2d_wgrad: stride=%d (1..4 supported)", stride);
  WgradParams& p = plan.p;
  p.pad_t = pad_t; p.pad_l = pad_l; p.stride = stride; p.quad = 0;
  if (reinterpret_cast<uintptr_t>(workspace) & 31) return fail(GANB_E_BADARG, "filter gradient: workspace must be 32-byte aligned");
  p.partial = static_cast<float*>(workspace);

  CUtensorMap tmX, tmDY;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cin * 2, (uint64_t)w * cin * 2, (uint64_t)h * w * cin * 2};
    const uint32_t box[4] = {64, (uint32_t)((p.bw - 1) * stride + 1), (uint32_t)((p.bh - 1) * stride + 1), (uint32_t)p.bn};
    const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = encode_tmap_bf16(&tmX, x, 4, dims, strides, box, stride > 1 ? estr : nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)wo, (uint64_t)ho, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cout * 2, (uint64_t)wo * cout * 2, (uint64_t)ho * wo * cout * 2};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = encode_tmap_bf16(&tmDY, dy, 4, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  int rc;
  switch (plan.bn_tile) {
    case 64: rc = launch_wgrad<64, ST_WG64>(tmX, tmDY, p, plan.grid, stream); break;
    case 128: rc = launch_wgrad<128, ST_WG128>(tmX, tmDY, p, plan.grid, stream); break;
    default: rc = launch_wgrad<256, ST_WG256>(tmX, tmDY, p, plan.grid, stream); break;
  }
  if (rc) return rc;
  const int64_t total = static_cast<int64_t>(kh) * kw * cin * cout;
  const int64_t n4 = total / 4;  // cout % 8 == 0
  launch_splitk_reduce(p.partial, dw, n4, p.splits, scale, beta, stream);
  GANB_CHECK_LAUNCH("splitk_reduce_kernel");
  return 0;
}


// ================================================================================================
// Sub-pixel form of UpsampleConv (common/resnet_block.py:83-97: nearest-2x upsample, then a 3x3 SAME convolution).
// Output pixel (2a+i, 2b+j) only sees the 2x2 low-resolution neighbourhood rows {a+i-1, a+i}, columns {b+j-1, b+j}:
//     y[2a+i, 2b+j] = sum_{p,q in {0,1}} E_ij[p][q] . x[a+i-1+p, b+j-1+q],   E_ij[p][q] = sum_{r in R_i[p], s in R_j[q]} W[r][s]
// with R_0 = ({0}, {1,2}), R_1 = ({0,1}, {2}).  Four 2x2 convolutions over the LOW-resolution tensor replace one 3x3
// convolution over the 4x larger one: 16/36 = 4/9 of the MMA work, and the upsampled operand is never written.
// The result is kept in "quad layout" [N, h, w, 4 = 2i+j, C]: a pixel permutation of NHWC [N, 2h, 2w, C] that the
// batch-statistics normalisation behind it (the only consumer) reads directly (ganb_norm_act_* x_quad = 1).
// ================================================================================================
namespace ganb {

// E (bf16) in both GEMM layouts: we_t [16 = 4*(2i+j) + 2p+q][co][ci] (fprop), we_n [16][ci][co] (dgrad)
__global__ void __launch_bounds__(256) upconv_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ we_t,
                                                          __nv_bfloat16* __restrict__ we_n, int ci_n, int co_n) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int tiles_co = (co_n + 31) / 32;
  const int tco = blockIdx.x % tiles_co, tci = blockIdx.x / tiles_co;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t plane = static_cast<int64_t>(ci_n) * co_n;
  {
    const int t16 = blockIdx.y;                              // one (parity, tap) per block row
    const int i = t16 >> 3, j = (t16 >> 2) & 1, pp = (t16 >> 1) & 1, qq = t16 & 1;
    // rows r with (r + 1 - i) >> 1 == pp, columns s with (s + 1 - j) >> 1 == qq
    for (int jj = ty; jj < 32; jj += 8) {
      const int ci = tci * 32 + jj, co = tco * 32 + tx;
      float v = 0.f;
      if (ci < ci_n && co < co_n) {
        for (int r = 0; r < 3; ++r) {
          if (((r + 1 - i) >> 1) != pp) continue;
          for (int s2 = 0; s2 < 3; ++s2) {
            if (((s2 + 1 - j) >> 1) != qq) continue;
            v += w[(r * 3 + s2) * plane + static_cast<int64_t>(ci) * co_n + co];
          }
        }
        we_n[t16 * plane + static_cast<int64_t>(ci) * co_n + co] = __float2bfloat16_rn(v);
      }
      tile[jj][tx] = v;
    }
    __syncthreads();
    for (int jj = ty; jj < 32; jj += 8) {
      const int co = tco * 32 + jj, ci = tci * 32 + tx;
      if (ci < ci_n && co < co_n) we_t[t16 * plane + static_cast<int64_t>(co) * ci_n + ci] = __float2bfloat16_rn(tile[tx][jj]);
    }
  }
}

// dW[r][s] = beta*dW[r][s] + scale * sum_splits sum_{i,j} dE_ij[p(i,r)][q(j,s)]   (transpose of the map W -> E)
__global__ void upconv_fold_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int64_t plane4,
                                          int splits, const float* __restrict__ scale, float beta) {
  pdl_wait();
  const float sc = scale ? __ldg(scale) : 1.0f;
  const int64_t total = 9 * plane4;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int rs = static_cast<int>(idx / plane4);
    const int64_t e = idx - rs * plane4;
    const int r = rs / 3, s2 = rs - 3 * r;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int i = g >> 1, j = g & 1;
        const int t16 = g * 4 + ((r + 1 - i) >> 1) * 2 + ((s2 + 1 - j) >> 1);
        const float4 v = __ldg(reinterpret_cast<const float4*>(partial) + (static_cast<int64_t>(sp) * 16 + t16) * plane4 + e);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    float4 o = make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc);
    if (beta != 0.f) {
      const float4 d = reinterpret_cast<const float4*>(dw)[idx];
      o.x += beta * d.x; o.y += beta * d.y; o.z += beta * d.z; o.w += beta * d.w;
    }
    reinterpret_cast<float4*>(dw)[idx] = o;
  }
}

// shapes the CTA-pair kernel covers with 16 x 8 pixel tiles and whole 64-channel chunks on both sides
static bool upconv_shape_ok(int n, int h, int w, int cin, int cout) {
  return n > 0 && h >= HALO_BH && w >= HALO_BW && h % HALO_BH == 0 && w % HALO_BW == 0 && cin % 64 == 0 &&
         cout % 64 == 0 && cin >= 128 && cout >= 128;
}

// One launch of the pair kernel over the low-resolution tile grid.  fprop: og = 4 output groups; dgrad: rg = 4
// reduction groups over the channel slices of the quad-layout gradient.
static int launch_upconv(bool dgrad, const void* a, const void* wp, void* out, int n, int h, int w, int ca, int cn,
                         const float* alpha, const float* bias, int act, int out_dtype, cudaStream_t stream,
                         float* stats = nullptr) {
  IgemmParams p;
  p.stats = stats;
  p.gate = nullptr; p.gate_act = 0;
  p.N = n; p.Ho = h; p.Wo = w; p.Cout = cn;
  p.taps = 4; p.kw = 2; p.stride = 1; p.pad_t = 0; p.pad_l = 0;
  p.bw = HALO_BW; p.bh = HALO_BH; p.bn = 1;
  p.tiles_w = w / HALO_BW; p.tiles_h = h / HALO_BH; p.tiles_n = n;
  p.kchunks = ca / BK;
  p.flip = dgrad ? 1 : 0;
  p.alpha = alpha; p.bias = bias; p.residual = nullptr; p.res_up2 = 0;
  p.out = out; p.out_bf16 = (out_dtype == GANB_BF16); p.act = act;
  p.og = dgrad ? 1 : 4; p.rg = dgrad ? 4 : 1;
  p.out_cstride = dgrad ? cn : 4 * cn;
  {
    const int64_t es_ = p.out_bf16 ? 2 : 4;
    static const bool st32_ok = !(getenv("GANB_ST32") && getenv("GANB_ST32")[0] == '0');   // A/B switch
    p.st32 = st32_ok && (reinterpret_cast<uintptr_t>(p.out) % 32 == 0) && ((p.Cout * es_) % 32 == 0) &&
             ((static_cast<int64_t>(p.out_cstride) * es_) % 32 == 0);
    p.res32 = st32_ok && p.residual && (reinterpret_cast<uintptr_t>(p.residual) % 32 == 0) && (p.Cout % 8 == 0);
  }
  for (int g = 0; g < 4; ++g) {
    const int i = g >> 1, j = g & 1;
    p.gpad_t[g] = static_cast<signed char>(dgrad ? i : 1 - i);
    p.gpad_l[g] = static_cast<signed char>(dgrad ? j : 1 - j);
  }
  const int pair_bn = cn > 128 ? 256 : 128;
  const int a_channels = dgrad ? 4 * ca : ca;      // the gradient tensor holds the four parity slices side by side
  CUtensorMap tmA, tmB;
  const int halo_w = HALO_BW + 1, halo_h = HALO_BH + 1;
  {
    const uint64_t dims[4] = {(uint64_t)a_channels, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)a_channels * 2, (uint64_t)w * a_channels * 2, (uint64_t)h * w * a_channels * 2};
    const uint32_t box[4] = {BK, (uint32_t)halo_w, (uint32_t)halo_h, 1};
    int rc = encode_tmap_bf16(&tmA, a, 4, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)ca, (uint64_t)cn, 16};
    const uint64_t strides[2] = {(uint64_t)ca * 2, (uint64_t)ca * cn * 2};
    const uint32_t box[3] = {BK, (uint32_t)(pair_bn / 2), 1};
    int rc = encode_tmap_bf16(&tmB, wp, 3, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  const int halo_bytes = halo_w * halo_h * 128;
  const int a_stage = (halo_bytes + 1023) / 1024 * 1024;
  if (pair_bn == 256) return launch_pair<256, 3, ST_PAIR256_SB>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
  return launch_pair<128, 4, ST_PAIR128_SB>(tmA, tmB, p, a_stage, halo_w, halo_bytes, stream);
}

}  // namespace ganb

extern "C" int ganb_upconv_supported(int n, int h, int w, int cin, int cout) {
  return upconv_shape_ok(n, h, w, cin, cout) ? 1 : 0;
}

extern "C" int ganb_upconv_pack(const float* w_hwio, void* we_t_bf16, void* we_n_bf16, int cin, int cout, void* stream) {
  if (!w_hwio || !we_t_bf16 || !we_n_bf16) return fail(GANB_E_BADARG, "upconv_pack: null buffer");
  const int blocks = ceil_div(cin, 32) * ceil_div(cout, 32);
  launch_k(upconv_pack_kernel, dim3(blocks, 16), 256, 0, static_cast<cudaStream_t>(stream), w_hwio,
           static_cast<__nv_bfloat16*>(we_t_bf16), static_cast<__nv_bfloat16*>(we_n_bf16), cin, cout);
  GANB_CHECK_LAUNCH("upconv_pack_kernel");
  return 0;
}

extern "C" int ganb_upconv_fprop(const void* x_bf16, const void* we_t_bf16, void* y_quad, int n, int h, int w, int cin,
                                 int cout, const float* alpha, const float* bias, int act, int out_dtype, void* stream) {
  if (!x_bf16 || !we_t_bf16 || !y_quad) return fail(GANB_E_BADARG, "upconv_fprop: null buffer");
  if (!upconv_shape_ok(n, h, w, cin, cout))
    return fail(GANB_E_UNSUPPORTED, "upconv_fprop: n=%d h=%d w=%d cin=%d cout=%d (see ganb_upconv_supported)", n, h, w, cin, cout);
  return launch_upconv(false, x_bf16, we_t_bf16, y_quad, n, h, w, cin, cout, alpha, bias, act, out_dtype,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int ganb_upconv_stats_rows(int n, int h, int w, int cin, int cout, int groups) {
  if (!upconv_shape_ok(n, h, w, cin, cout) || groups <= 0 || n % groups != 0 || cout % 32 != 0) return 0;
  return (w / HALO_BW) * (h / HALO_BH) * (n / groups) * 4;      // four output parities per low-resolution pixel tile
}

extern "C" int ganb_upconv_fprop_stats(const void* x_bf16, const void* we_t_bf16, void* y_quad, int n, int h, int w,
                                       int cin, int cout, const float* alpha, const float* bias, int act, int out_dtype,
                                       float* stats, int groups, void* stream) {
  if (!x_bf16 || !we_t_bf16 || !y_quad || !stats) return fail(GANB_E_BADARG, "upconv_fprop_stats: null buffer");
  if (!ganb_upconv_stats_rows(n, h, w, cin, cout, groups))
    return fail(GANB_E_UNSUPPORTED, "upconv_fprop_stats: n=%d h=%d w=%d cin=%d cout=%d groups=%d", n, h, w, cin, cout, groups);
  return launch_upconv(false, x_bf16, we_t_bf16, y_quad, n, h, w, cin, cout, alpha, bias, act, out_dtype,
                       static_cast<cudaStream_t>(stream), stats);
}

extern "C" int ganb_upconv_dgrad(const void* dy_quad_bf16, const void* we_n_bf16, void* dx, int n, int h, int w, int cin,
                                 int cout, const float* alpha, int out_dtype, void* stream) {
  if (!dy_quad_bf16 || !we_n_bf16 || !dx) return fail(GANB_E_BADARG, "upconv_dgrad: null buffer");
  if (!upconv_shape_ok(n, h, w, cin, cout))
    return fail(GANB_E_UNSUPPORTED, "upconv_dgrad: n=%d h=%d w=%d cin=%d cout=%d (see ganb_upconv_supported)", n, h, w, cin, cout);
  return launch_upconv(true, dy_quad_bf16, we_n_bf16, dx, n, h, w, cout, cin, alpha, nullptr, 0, out_dtype,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int64_t ganb_upconv_wgrad_workspace(int n, int h, int w, int cin, int cout) {
  WgradPlan plan;
  plan_wgrad(n, h, w, cin, cout, 4, 4, &plan);     // 16 (parity, tap) units
  return static_cast<int64_t>(plan.p.splits) * 16 * cin * cout * 4;
}

extern "C" int ganb_upconv_wgrad(const void* x_bf16, const void* dy_quad_bf16, float* dw_hwio, void* workspace, int n,
                                 int h, int w, int cin, int cout, const float* scale, float beta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x_bf16 || !dy_quad_bf16 || !dw_hwio || !workspace) return fail(GANB_E_BADARG, "upconv_wgrad: null buffer");
  if (!upconv_shape_ok(n, h, w, cin, cout))
    return fail(GANB_E_UNSUPPORTED, "upconv_wgrad: n=%d h=%d w=%d cin=%d cout=%d (see ganb_upconv_supported)", n, h, w, cin, cout);
  WgradPlan plan;
  plan_wgrad(n, h, w, cin, cout, 4, 4, &plan);
  WgradParams& p = plan.p;
  p.kw = 2; p.pad_t = 0; p.pad_l = 0; p.stride = 1; p.quad = 1;
  if (reinterpret_cast<uintptr_t>(workspace) & 31) return fail(GANB_E_BADARG, "filter gradient: workspace must be 32-byte aligned");
  p.partial = static_cast<float*>(workspace);
  CUtensorMap tmX, tmDY;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cin * 2, (uint64_t)w * cin * 2, (uint64_t)h * w * cin * 2};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = encode_tmap_bf16(&tmX, x_bf16, 4, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t c4 = 4ull * cout;
    const uint64_t dims[4] = {c4, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {c4 * 2, (uint64_t)w * c4 * 2, (uint64_t)h * w * c4 * 2};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = encode_tmap_bf16(&tmDY, dy_quad_bf16, 4, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  int rc;
  switch (plan.bn_tile) {
    case 64: rc = launch_wgrad<64, ST_WG64>(tmX, tmDY, p, plan.grid, stream); break;
    case 128: rc = launch_wgrad<128, ST_WG128>(tmX, tmDY, p, plan.grid, stream); break;
    default: rc = launch_wgrad<256, ST_WG256>(tmX, tmDY, p, plan.grid, stream); break;
  }
  if (rc) return rc;
  const int64_t plane4 = static_cast<int64_t>(cin) * cout / 4;
  int blocks = static_cast<int>(ceil_div64(9 * plane4, 256));
  if (blocks > 4 * tc_sm_count()) blocks = 4 * tc_sm_count();
  launch_k(upconv_fold_reduce_kernel, blocks, 256, 0, stream, p.partial, dw_hwio, plane4, p.splits, scale, beta);
  GANB_CHECK_LAUNCH("upconv_fold_reduce_kernel");
  return 0;
}


// ================================================================================================
// TF32 operand mode (north star: "BF16 or TF32 inputs and FP32 accumulation"): the same implicit-GEMM kernels with fp32
// operands read by kind::tf32 MMAs -- per-layer results within 1e-3 of fp32 (tests/test_gpu_tf32.py).  Every shape takes
// the per-tap TMA-box route (conv_igemm_kernel<..., TF32 = true>) and the pixel-split filter-gradient kernel
// (conv_wgrad_kernel<..., TF32 = true>); operands should be rounded to TF32 beforehand (ganb_round_tf32 /
// ganb_transpose_tf32: round-to-nearest) because the MMA itself truncates the 13 low mantissa bits.
// ================================================================================================
namespace ganb {

template <int BN, int STAGES>
static int launch_igemm_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, IgemmParams& p, cudaStream_t stream) {
  const int smem = STAGES * (A_STAGE_BYTES + BN * 128) + 1024 + 256;
  auto kern = conv_igemm_kernel<BN, STAGES, false, true>;
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "igemm tf32 smem attribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  p.tiles_co = ceil_div(p.Cout, BN);
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_co;
  const int grid = p.num_tiles < tc_sm_count() ? p.num_tiles : tc_sm_count();
  launch_k(kern, grid, NUM_THREADS, smem, stream, tmA, tmB, p, tmA);
  GANB_CHECK_LAUNCH("conv_igemm_kernel<tf32>");
  return 0;
}

template <int BN, int STAGES>
static int launch_wgrad_tf32(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradParams& p, int grid,
                             cudaStream_t stream) {
  constexpr int smem = STAGES * (PB * BM * 4 + PB * BN * 4) + 1024 + 256;
  auto kern = conv_wgrad_kernel<BN, STAGES, true>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "wgrad tf32 smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  launch_k(kern, grid, WG_THREADS, smem, stream, tmX, tmDY, p);
  GANB_CHECK_LAUNCH("conv_wgrad_kernel<tf32>");
  return 0;
}

}  // namespace ganb

extern "C" int ganb_conv2d_igemm_tf32(const float* x, const float* wp, void* y, int n, int h, int w, int cin, int ho,
                                      int wo, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int flip_taps,
                                      const float* alpha, const float* bias, const float* residual, int residual_up2,
                                      int act, int out_dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !wp || !y) return fail(GANB_E_BADARG, "conv2d_igemm_tf32: null buffer");
  if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || ho <= 0 || wo <= 0 || cout <= 0 || kh <= 0 || kw <= 0)
    return fail(GANB_E_BADARG, "conv2d_igemm_tf32: non-positive dimension");
  if (cin % 4 != 0) return fail(GANB_E_UNSUPPORTED, "conv2d_igemm_tf32: cin=%d must be a multiple of 4", cin);
  if (stride < 1 || stride > 4) return fail(GANB_E_UNSUPPORTED, "conv2d_igemm_tf32: stride=%d (1..4 supported)", stride);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wp)) & 15)
    return fail(GANB_E_BADARG, "conv2d_igemm_tf32: x / wp must be 16-byte aligned");
  IgemmParams p;
  p.N = n; p.Ho = ho; p.Wo = wo; p.Cout = cout;
  p.taps = kh * kw; p.kw = kw;
  p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
  pick_box(BM, ho, wo, &p.bw, &p.bh, &p.bn);
  p.stats = nullptr;
  p.gate = nullptr; p.gate_act = 0;
  p.tiles_w = ceil_div(wo, p.bw);
  p.tiles_h = ceil_div(ho, p.bh);
  p.tiles_n = ceil_div(n, p.bn);
  p.kchunks = ceil_div(cin, 32);
  p.flip = flip_taps;
  p.alpha = alpha; p.bias = bias; p.residual = residual;
  p.res_up2 = (residual && residual_up2) ? 1 : 0;
  if (p.res_up2 && ((ho | wo) & 1)) return fail(GANB_E_BADARG, "conv2d_igemm_tf32: upsampled residual needs even ho, wo");
  p.out = y; p.out_bf16 = (out_dtype == GANB_BF16); p.act = act;
  p.og = 1; p.rg = 1; p.out_cstride = cout;
  {
    const int64_t es_ = p.out_bf16 ? 2 : 4;
    static const bool st32_ok = !(getenv("GANB_ST32") && getenv("GANB_ST32")[0] == '0');   // A/B switch
    p.st32 = st32_ok && (reinterpret_cast<uintptr_t>(p.out) % 32 == 0) && ((p.Cout * es_) % 32 == 0) &&
             ((static_cast<int64_t>(p.out_cstride) * es_) % 32 == 0);
    p.res32 = st32_ok && p.residual && (reinterpret_cast<uintptr_t>(p.residual) % 32 == 0) && (p.Cout % 8 == 0);
  }
  for (int i = 0; i < 4; ++i) { p.gpad_t[i] = 0; p.gpad_l[i] = 0; }
  const int bn_tile = cout <= 16 ? 16 : cout <= 64 ? 64 : 128;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cin * 4, (uint64_t)w * cin * 4, (uint64_t)h * w * cin * 4};
    const uint32_t box[4] = {32, (uint32_t)((p.bw - 1) * stride + 1), (uint32_t)((p.bh - 1) * stride + 1), (uint32_t)p.bn};
    const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (box[1] > 256 || box[2] > 256) return fail(GANB_E_UNSUPPORTED, "conv2d_igemm_tf32: tile extent exceeds the TMA box limit");
    int rc = encode_tmap_f32(&tmA, x, 4, dims, strides, box, stride > 1 ? estr : nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cout, (uint64_t)(kh * kw)};
    const uint64_t strides[2] = {(uint64_t)cin * 4, (uint64_t)cin * cout * 4};
    const uint32_t box[3] = {32, (uint32_t)bn_tile, 1};
    int rc = encode_tmap_f32(&tmB, wp, 3, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  switch (bn_tile) {
    case 16: return launch_igemm_tf32<16, 8>(tmA, tmB, p, stream);
    case 64: return launch_igemm_tf32<64, 8>(tmA, tmB, p, stream);
    default: return launch_igemm_tf32<128, 6>(tmA, tmB, p, stream);
  }
}

extern "C" int64_t ganb_conv2d_wgrad_tf32_workspace(int n, int ho, int wo, int cin, int cout, int kh, int kw) {
  WgradPlan plan;
  plan_wgrad(n, ho, wo, cin, cout, kh, kw, &plan, 128);
  return static_cast<int64_t>(plan.p.splits) * kh * kw * cin * cout * 4;
}

extern "C" int ganb_conv2d_wgrad_tf32(const float* x, const float* dy, float* dw, void* workspace, int n, int h, int w,
                                      int cin, int ho, int wo, int cout, int kh, int kw, int stride, int pad_t,
                                      int pad_l, const float* scale, float beta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !dy || !dw || !workspace) return fail(GANB_E_BADARG, "conv2d_wgrad_tf32: null buffer");
  if (cin % 4 != 0 || cout % 4 != 0)
    return fail(GANB_E_UNSUPPORTED, "conv2d_wgrad_tf32: cin=%d and cout=%d must be multiples of 4", cin, cout);
  if (stride < 1 || stride > 4) return fail(GANB_E_UNSUPPORTED, "conv2d_wgrad_tf32: stride=%d (1..4 supported)", stride);
  WgradPlan plan;
  plan_wgrad(n, ho, wo, cin, cout, kh, kw, &plan, 128);
  WgradParams& p = plan.p;
  p.pad_t = pad_t; p.pad_l = pad_l; p.stride = stride; p.quad = 0;
  if (reinterpret_cast<uintptr_t>(workspace) & 31) return fail(GANB_E_BADARG, "filter gradient: workspace must be 32-byte aligned");
  p.partial = static_cast<float*>(workspace);
  CUtensorMap tmX, tmDY;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cin * 4, (uint64_t)w * cin * 4, (uint64_t)h * w * cin * 4};
    const uint32_t box[4] = {32, (uint32_t)((p.bw - 1) * stride + 1), (uint32_t)((p.bh - 1) * stride + 1), (uint32_t)p.bn};
    const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = encode_tmap_f32_base32(&tmX, x, 4, dims, strides, box, stride > 1 ? estr : nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)wo, (uint64_t)ho, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cout * 4, (uint64_t)wo * cout * 4, (uint64_t)ho * wo * cout * 4};
    const uint32_t box[4] = {32, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = encode_tmap_f32_base32(&tmDY, dy, 4, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  int rc;
  if (plan.bn_tile == 64) rc = launch_wgrad_tf32<64, 4>(tmX, tmDY, p, plan.grid, stream);
  else rc = launch_wgrad_tf32<128, 3>(tmX, tmDY, p, plan.grid, stream);
  if (rc) return rc;
  const int64_t n4 = static_cast<int64_t>(kh) * kw * cin * cout / 4;
  launch_splitk_reduce(p.partial, dw, n4, p.splits, scale, beta, stream);
  GANB_CHECK_LAUNCH("splitk_reduce_kernel");
  return 0;
}
