// Multi-GPU exchange over peer memory (one process per GPU, CUDA IPC mappings of the other ranks'
// tables, NVLink / NVSwitch underneath): no collective library call on the data path.
//
//   * the scattering-density kernel stores every texel it computes to all ranks (kernel_density.cu):
//     the all-gather of SURVEY.md section 8e is fused into the pass that produces the data;
//   * peer_push: a rank's irradiance partial sums (60 KiB) and, at the end of Init, its slab of the
//     final scattering table(s) are copied to the same offset of every rank's buffer;
//   * peer_barrier: every rank writes the epoch number into its slot of every other rank's flag
//     array (release, system scope) and waits until all the slots of its own array have reached it.
//     Stream order + the fence make the stores of the kernels enqueued before the barrier visible
//     to the kernels every rank enqueues after it. A rank that waits longer than the timeout fails
//     its Init and poisons the flag arrays of all ranks, so that every rank fails the same Init.
//   * sum_partials: irradiance = sum over ranks of the partial sums, in rank order on every rank
//     (bit-identical results everywhere).
#include <cstdlib>

#include "pas_kernels.h"

namespace pas {
namespace {

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_barrier_kernel(PeerFlags f, unsigned long long timeout_ns) {
  const int p = threadIdx.x;
  // the channel's epoch counter lives beside the flags and is advanced here, by the one warp of the
  // one barrier kernel that can run at a time on the channel's stream
  unsigned epoch = 0;
  if (p == 0) {
    unsigned* counter = f.flags[f.rank] + PAS_FLAG_EPOCH;
    epoch = *counter + 1;
    *counter = epoch;
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  if (p >= f.world || p == f.rank) return;
  __threadfence_system();
  unsigned* dst = f.flags[p] + f.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
  const unsigned* src = f.flags[f.rank] + p;
  const unsigned* poison = f.poison[f.rank];
  const unsigned long long t0 = global_timer_ns();
  for (;;) {
    unsigned v, bad;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(bad) : "l"(poison) : "memory");
    if (bad != 0) {
      // another rank gave up waiting: its tables and ours are out of step; fail this Init too
      *f.error = 2;
      break;
    }
    if ((int)(v - epoch) >= 0) break;
    if (global_timer_ns() - t0 > timeout_ns) {
      // a rank died or fell out of step: report it (mapped host memory) and tell every rank, so that
      // a late rank does not sail through barriers whose flags it finds already set
      *f.error = 1;
      for (int r = 0; r < f.world; ++r) {
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.poison[r]), "r"(1u) : "memory");
      }
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

// `chunks` ranges of n elements each, the first at `offset`, `stride` elements apart.
template <typename V>
__global__ void peer_push_kernel(const V* __restrict__ src, size_t n, size_t offset, size_t stride,
                                 int chunks, PeerTargets t) {
  const size_t total = n * (size_t)chunks;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t at = offset + (i / n) * stride + i % n;
    const V v = src[at];
    for (int p = 0; p < t.n; ++p) reinterpret_cast<V*>(t.dst[p])[at] = v;
  }
}

__global__ void sum_partials_kernel(const float* __restrict__ parts, int world, size_t stride, int n,
                                    float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int r = 0; r < world; ++r) s += parts[(size_t)r * stride + i];
  out[i] = s;
}

}  // namespace

cudaError_t launch_peer_barrier(const PeerFlags& f, cudaStream_t stream) {
  // PAS_PEER_TIMEOUT_MS: how long a rank waits for the others before it fails the Init (default 5 s;
  // raise it when ranks may start their Init calls further apart than that)
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("PAS_PEER_TIMEOUT_MS");
    const long long ms = e != nullptr ? atoll(e) : 0;
    return (unsigned long long)(ms > 0 ? ms : 5000) * 1000000ull;
  }();
  peer_barrier_kernel<<<1, 32, 0, stream>>>(f, timeout_ns);
  return cudaGetLastError();
}

cudaError_t launch_peer_push(const void* src, size_t bytes, size_t offset_bytes, const PeerTargets& t,
                             cudaStream_t stream, int chunks, size_t stride_bytes) {
  if (bytes == 0 || t.n == 0 || chunks <= 0) return cudaSuccess;
  if (bytes % 16 == 0 && offset_bytes % 16 == 0 && stride_bytes % 16 == 0) {
    const size_t n = bytes / 16, total = n * chunks;
    const int blocks = (int)((total + 255) / 256 < 592 ? (total + 255) / 256 : 592);
    peer_push_kernel<uint4><<<blocks, 256, 0, stream>>>(static_cast<const uint4*>(src), n, offset_bytes / 16,
                                                        stride_bytes / 16, chunks, t);
  } else {
    if (bytes % 4 != 0 || offset_bytes % 4 != 0 || stride_bytes % 4 != 0) return cudaErrorInvalidValue;
    const size_t n = bytes / 4, total = n * chunks;
    const int blocks = (int)((total + 255) / 256 < 592 ? (total + 255) / 256 : 592);
    peer_push_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(src), n, offset_bytes / 4,
                                                        stride_bytes / 4, chunks, t);
  }
  return cudaGetLastError();
}

cudaError_t launch_sum_partials(const float* parts, int world, size_t stride, int n, float* out,
                                cudaStream_t stream) {
  sum_partials_kernel<<<(n + 255) / 256, 256, 0, stream>>>(parts, world, stride, n, out);
  return cudaGetLastError();
}

}  // namespace pas
