// Host-side helpers shared by every translation unit of libganb200: error reporting, launch checks,
// and TMA descriptor encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ganb200.h"

namespace ganb {

// Records a thread-local message retrievable through ganb_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);

// Number of SMs of the current device (cached).
int sm_count();
int tc_sm_count();   // sm_count() under the calling thread's ganb_set_sm_limit

// Encodes a tiled TMA descriptor over a bf16 tensor. dims/strides innermost-first; strides in bytes
// for dims 1..rank-1. Returns 0 or a negative GANB_E_* code.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);
// The same over an fp32 tensor (operands of the kind::tf32 tensor-core kernels).
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);
// fp32 with the 128-byte swizzle on a 32-byte base (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): the only shared-memory layout
// tcgen05 accepts for MN-major (channels-contiguous) TF32 operands -- the filter-gradient kernel.
int encode_tmap_f32_base32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

// Counts every kernel launch issued by the library (exported as ganb_launch_count()).
void count_launch();

#define GANB_CHECK_LAUNCH(name)                                                          \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) return ::ganb::fail(GANB_E_LAUNCH, "%s: %s", name, cudaGetErrorString(e__)); \
    ::ganb::count_launch();                                                              \
  } while (0)

// Every kernel of the library is launched through launch_k with programmatic dependent launch (PDL) allowed: the
// kernel may be scheduled while its predecessor in the stream is still draining, runs its prologue (barrier init,
// TMEM allocation, descriptor prefetch) and blocks in pdl_wait() until the predecessor's results are visible.  This
// hides the ~1-2 us kernel-to-kernel gap of the ~300 launches of a training step (also inside captured CUDA graphs).
// Opt-in with GANB_PDL=1 (measured neutral for the captured training step; plain stream-ordered launches otherwise).
bool pdl_enabled();

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(args)...);
}

#ifdef __CUDACC__
// Device side of PDL: wait until the grids this launch depends on have completed and their writes are visible, then
// allow the next kernel of the stream to be scheduled.  Must precede the first access to global memory.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ganb
