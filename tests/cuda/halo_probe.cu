// Experiment (not part of the library): can ONE TMA-loaded halo tile [(bh+2) x (bw+2) pixels][64 ch] in shared
// memory serve all nine taps of a 3x3 convolution through shifted UMMA descriptors?
//   rows of tap (r,s):  smem row (g + r) * (bw+2) + s + i   for pixel row-group g (8 pixels wide), i = 0..7
//   => descriptor start = base + (r*(bw+2)+s)*128 B, SBO = (bw+2)*128 B (NOT a multiple of 1024), 128B swizzle.
// This only works if the hardware applies the 128B swizzle XOR on absolute shared-memory address bits.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../gan_lib_tensorflow_b200/csrc/host_common.h"
#include "../../gan_lib_tensorflow_b200/csrc/ptx.cuh"

using namespace ganb;

constexpr int BW = 8, BH = 16, HW2 = BW + 2, HH2 = BH + 2, ROWS = HW2 * HH2;  // 180 halo rows
constexpr int N = 64;

__global__ void __launch_bounds__(128, 1)
halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
            int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 180 * 128 = 23040 -> pad to 23552 (1024 multiple)
  uint8_t* sB = smem + 23552;         // 64 * 128 = 8192
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, N, 0, 0);
  for (int tap = 0; tap < 9; ++tap) {
    const int r = tap / 3, s = tap % 3;
    if (threadIdx.x == 0) {
      if (tap == 0) {
        mbar_arrive_expect_tx(bar, ROWS * 128 + N * 128);
        tma_load_2d(sA, &tmA, bar, 0, 0);
        tma_load_2d(sB, &tmB, bar, 0, 0);
        mbar_wait(bar, 0);
      }
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA) + (r * HW2 + s) * 128;
      uint64_t adesc = umma_smem_desc_sw128(a_addr, 16, HW2 * 128);
      if (use_base_offset) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sB), 16, 1024);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, k > 0);
      umma_commit(done);
      mbar_wait(done, tap & 1);
    }
    __syncthreads();
    tc_fence_after();
    uint32_t v[32];
    for (int c = 0; c < N / 32; ++c) {
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(tap * 128 + warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<__nv_bfloat16> hA(ROWS * 64), hB(N * 64);
  std::vector<float> fA(ROWS * 64), fB(N * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = bf((rand() % 2001 - 1000) / 1000.f); hA[i] = __float2bfloat16(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = bf((rand() % 2001 - 1000) / 1000.f); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB; float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 9 * 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  { uint64_t d[2] = {64, (uint64_t)ROWS}, st[1] = {128}; uint32_t b[2] = {64, (uint32_t)ROWS};
    if (encode_tmap_bf16(&tmA, dA, 2, d, st, b, nullptr)) { printf("encode A: %s\n", ganb_last_error()); return 1; } }
  { uint64_t d[2] = {64, (uint64_t)N}, st[1] = {128}; uint32_t b[2] = {64, (uint32_t)N};
    if (encode_tmap_bf16(&tmB, dB, 2, d, st, b, nullptr)) { printf("encode B: %s\n", ganb_last_error()); return 1; } }
  const int smem = 23552 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ref(9 * 128 * N), got(9 * 128 * N);
  for (int tap = 0; tap < 9; ++tap) for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    const int r = tap / 3, s = tap % 3, row = (m / 8 + r) * HW2 + s + (m % 8);
    float acc = 0; for (int k = 0; k < 64; ++k) acc += fA[row * 64 + k] * fB[n * 64 + k];
    ref[(tap * 128 + m) * N + n] = acc;
  }
  for (int ubo = 0; ubo < 2; ++ubo) {
    cudaMemset(dOut, 0, got.size() * 4);
    halo_kernel<<<1, 128, smem>>>(tmA, tmB, dOut, ubo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("base_offset=%d: CUDA error %s\n", ubo, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(got.data(), dOut, got.size() * 4, cudaMemcpyDeviceToHost);
    for (int tap = 0; tap < 9; ++tap) {
      double worst = 0; for (int i = 0; i < 128 * N; ++i) worst = fmax(worst, fabs(got[tap * 128 * N + i] - ref[tap * 128 * N + i]));
      printf("base_offset_field=%d tap(%d,%d) max_abs_err=%.3e %s\n", ubo, tap / 3, tap % 3, worst, worst < 1e-3 ? "OK" : "MISMATCH");
    }
  }
  return 0;
}
