#include "host_common.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

namespace ganb {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("GANB_PDL");
    // opt-in, and NOT recommended: -0.4 ... -0.5 % inside captured graphs, but one of two 300-pair bench runs on the final
    // round-2 tree (Tape branches on several streams) did not finish with it on (profiles/r02_schedule_ab.txt, item 9)
    mode = (e && e[0] == '1') ? 1 : 0;
  }
  return mode == 1;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

// Tensor-core kernels are persistent (one CTA per SM, grid = min(tiles, SMs)).  A per-thread limit on the SMs they may
// take lets two independent passes issued on two streams share the GPU side by side instead of kernel by kernel
// (ganb_set_sm_limit; used by the D+G pair schedule).  0 = no limit.
static thread_local int tl_sm_limit = 0;
int tc_sm_count() {
  const int n = sm_count();
  return (tl_sm_limit > 0 && tl_sm_limit < n) ? tl_sm_limit : n;
}
extern "C" int ganb_set_sm_limit(int sms) {
  const int prev = tl_sm_limit;
  tl_sm_limit = sms > 0 ? sms : 0;
  return prev;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static encode_tiled_fn g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode = reinterpret_cast<encode_tiled_fn>(fn);
}

static int encode_tmap(CUtensorMapDataType dtype, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                       CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B);

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return encode_tmap(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, out, base, rank, dims, strides_bytes, box, elem_strides);
}

int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return encode_tmap(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, base, rank, dims, strides_bytes, box, elem_strides);
}

int encode_tmap_f32_base32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return encode_tmap(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, base, rank, dims, strides_bytes, box, elem_strides,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

static int encode_tmap(CUtensorMapDataType dtype, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                       CUtensorMapSwizzle swizzle) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return fail(GANB_E_ARCH, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(out, dtype, static_cast<cuuint32_t>(rank),
                        const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(GANB_E_BADARG, "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu %llu %llu %llu)",
                static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0));
  return 0;
}

}  // namespace ganb

extern "C" const char* ganb_last_error(void) { return ganb::g_err; }

extern "C" int ganb_abi_version(void) { return GANB_ABI_VERSION; }

extern "C" int64_t ganb_launch_count(void) { return static_cast<int64_t>(ganb::g_launches.load()); }

extern "C" int ganb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return ganb::fail(GANB_E_ARCH, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return ganb::fail(GANB_E_ARCH, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}
