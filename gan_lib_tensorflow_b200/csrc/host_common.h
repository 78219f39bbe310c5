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

// Encodes a tiled TMA descriptor over a bf16 tensor. dims/strides innermost-first; strides in bytes
// for dims 1..rank-1. Returns 0 or a negative GANB_E_* code.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

// Counts every kernel launch issued by the library (exported as ganb_launch_count()).
void count_launch();

#define GANB_CHECK_LAUNCH(name)                                                          \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) return ::ganb::fail(GANB_E_LAUNCH, "%s: %s", name, cudaGetErrorString(e__)); \
    ::ganb::count_launch();                                                              \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ganb
