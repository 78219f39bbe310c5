// Small all-reduces over NVLink / NVSwitch peer memory, written for the statistic exchanges of the data-parallel path:
// cross-GPU BatchNorm moments (north star; reference coupling point common/ops/normalization.py:47), the
// [sum dy | sum dy*xhat] pair of its backward pass, and PGGAN's minibatch-stddev scalar (PGGAN/model_nvidia.py:20-28).
// These messages are a few KB: a library collective costs its launch + protocol latency (~20-30 us each, dozens per
// step), so the exchange is ONE single-CTA kernel per call over buffers that every GPU of the node has mapped
// (symmetric memory, one allocation per rank, same layout everywhere):
//
//   site region of rank r :  [ flags: world x u32 | epoch: u32 | pad ][ data parity 0: count floats ][ data parity 1 ]
//
//   1. epoch = ++site.epoch (device memory: the counter advances under CUDA-graph replay as well)
//   2. every thread copies its part of the local contribution into OUR data[epoch & 1]
//   3. system-scope release, then one thread per peer stores `epoch` into the peer's flags[our rank]
//   4. one thread per peer spins (acquire, system scope) on OUR flags[peer] >= epoch
//   5. every thread sums data[epoch & 1] of all ranks in RANK ORDER (remote loads): the result is bit-identical on
//      every GPU, which keeps replicated parameters identical without a broadcast
//
// A rank can overwrite data[p] only two calls later, after it has seen every peer's flag of the call in between, which
// a peer raises only once its previous call (all of whose remote reads are done) has completed: no further barrier.
// Every spin is bounded and traps (a peer that died must not wedge the box).
#include "host_common.h"

namespace ganb {

constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_HEADER_BYTES = 128;
constexpr unsigned PEER_SPIN_LIMIT = 1u << 28;

struct PeerBufs {
  char* base[PEER_MAX_WORLD];   // this process' mapping of every rank's symmetric buffer
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// exchange of `count` floats staged in shared memory; returns with sm[] holding the sums over ranks
__device__ __forceinline__ void peer_exchange(const PeerBufs& pb, int rank, int world, int64_t site_off, int count,
                                              float* sm) {
  char* mine = pb.base[rank] + site_off;
  unsigned* my_flags = reinterpret_cast<unsigned*>(mine);
  unsigned* my_epoch = my_flags + PEER_MAX_WORLD;
  __shared__ unsigned epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = *my_epoch + 1;
    *my_epoch = epoch_s;
  }
  __syncthreads();
  const unsigned epoch = epoch_s;
  const int64_t data_off = site_off + PEER_HEADER_BYTES + static_cast<int64_t>(epoch & 1u) * count * 4;
  float* my_data = reinterpret_cast<float*>(pb.base[rank] + data_off);
  for (int i = threadIdx.x; i < count; i += blockDim.x) my_data[i] = sm[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    st_release_sys(reinterpret_cast<unsigned*>(pb.base[r] + site_off) + rank, epoch);
    unsigned spins = 0;
    // epochs only grow; the signed difference tolerates wrap-around
    while (static_cast<int>(ld_acquire_sys(my_flags + r) - epoch) < 0) {
      if (++spins > PEER_SPIN_LIMIT) __trap();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < world; ++r) acc += ld_relaxed_sys(reinterpret_cast<const float*>(pb.base[r] + data_off) + i);
    sm[i] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) peer_allreduce_kernel(const PeerBufs pb, int rank, int world, int64_t site_off,
                                                              const float* __restrict__ in, float* __restrict__ out,
                                                              int count, float scale) {
  pdl_wait();
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < count; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  peer_exchange(pb, rank, world, site_off, count, sm);
  for (int i = threadIdx.x; i < count; i += blockDim.x) out[i] = sm[i] * scale;
}

// mean / rstd [count] of the local share -> mean / rstd over all ranks (equal shares), in place: pack [mean | E x^2],
// exchange, unpack -- the whole forward statistic reduction of a synced batch norm in one launch
__global__ void __launch_bounds__(1024) peer_bn_moments_kernel(const PeerBufs pb, int rank, int world, int64_t site_off,
                                                               float* __restrict__ mean, float* __restrict__ rstd,
                                                               int count, float eps) {
  pdl_wait();
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const float m = mean[i], r = rstd[i];
    sm[i] = m;
    sm[count + i] = 1.0f / (r * r) - eps + m * m;
  }
  __syncthreads();
  peer_exchange(pb, rank, world, site_off, 2 * count, sm);
  const float inv_world = 1.0f / world;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const float m = sm[i] * inv_world;
    float var = sm[count + i] * inv_world - m * m;
    if (var < 0.f) var = 0.f;
    mean[i] = m;
    rstd[i] = rsqrtf(var + eps);
  }
}

static int make_bufs(void* const* peer_bufs, int rank, int world, PeerBufs* pb) {
  if (!peer_bufs || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
    return fail(GANB_E_BADARG, "peer all-reduce: world=%d (1..%d), rank=%d", world, PEER_MAX_WORLD, rank);
  for (int r = 0; r < PEER_MAX_WORLD; ++r) pb->base[r] = r < world ? static_cast<char*>(peer_bufs[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (!pb->base[r]) return fail(GANB_E_BADARG, "peer all-reduce: null buffer of rank %d", r);
  return 0;
}

}  // namespace ganb

using namespace ganb;

extern "C" int64_t ganb_peer_site_bytes(int count) {
  // header + two parities, rounded to 128 bytes
  return (PEER_HEADER_BYTES + 2LL * count * 4 + 127) / 128 * 128;
}

extern "C" int ganb_peer_max_count(void) { return 12032; }   // floats per exchange (47 KB of the default 48 KB of shared memory)

extern "C" int ganb_peer_allreduce(const float* in, float* out, int count, float scale, void* const* peer_bufs, int rank,
                                   int world, int64_t site_offset, void* stream) {
  if (!in || !out || count <= 0 || count > ganb_peer_max_count())
    return fail(GANB_E_BADARG, "peer_allreduce: count=%d (1..%d)", count, ganb_peer_max_count());
  PeerBufs pb;
  if (int rc = make_bufs(peer_bufs, rank, world, &pb)) return rc;
  const int threads = count >= 1024 ? 1024 : (count + 31) / 32 * 32;   // >= 32 >= world
  (void)cudaGetLastError();   // a stale (non-sticky) error of the symmetric-memory set-up must not be blamed on this launch
  const cudaError_t e = launch_k(peer_allreduce_kernel, 1, threads, static_cast<size_t>(count) * 4,
                                 static_cast<cudaStream_t>(stream), pb, rank, world, site_offset, in, out, count, scale);
  if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "peer_allreduce_kernel: %s (threads %d)", cudaGetErrorString(e), threads);
  count_launch();
  return 0;
}

extern "C" int ganb_peer_bn_moments(float* mean, float* rstd, int count, float eps, void* const* peer_bufs, int rank,
                                    int world, int64_t site_offset, void* stream) {
  if (!mean || !rstd || count <= 0 || 2 * count > ganb_peer_max_count())
    return fail(GANB_E_BADARG, "peer_bn_moments: count=%d (2*count <= %d)", count, ganb_peer_max_count());
  PeerBufs pb;
  if (int rc = make_bufs(peer_bufs, rank, world, &pb)) return rc;
  const int threads = count >= 1024 ? 1024 : (count + 31) / 32 * 32;
  (void)cudaGetLastError();
  const cudaError_t e = launch_k(peer_bn_moments_kernel, 1, threads, static_cast<size_t>(count) * 8,
                                 static_cast<cudaStream_t>(stream), pb, rank, world, site_offset, mean, rstd, count, eps);
  if (e != cudaSuccess) return fail(GANB_E_LAUNCH, "peer_bn_moments_kernel: %s (threads %d)", cudaGetErrorString(e), threads);
  count_launch();
  return 0;
}
