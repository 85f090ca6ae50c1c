// Peer-memory halo transport of the axis-0 slab decomposition (one process per GPU, NVLink 5 / NVSwitch).
//
// The reference has no multi-device data path (kernex/_src/utils.py:25-30 offers jax.pmap with one WINDOW per device),
// so there is no reference interface to mirror here; the north star asks for "per-step halo exchange ... over NVLink,
// overlapped with interior compute on a side stream".  A two-sided NCCL send/recv of a 64 KB halo row costs a
// rendezvous between the ranks' proxy threads and an NCCL kernel per step (+17 us at 2 GPUs, +138 us at 8, BENCH r01).
// Here every rank keeps its slab in a cudaMalloc'ed buffer whose CUDA-IPC handle the neighbours have opened, and ONE
// kernel per step does the whole exchange one-sidedly:
//
//   1. store `epoch` into the neighbours' READY flags        (my own rows are final: stream order + system fence)
//   2. spin until the neighbours have stored `epoch` into mine (ld.acquire.sys on local memory, written over NVLink)
//   3. pull the neighbours' edge rows into my halo rows        (16-byte ld.relaxed.sys over NVLink -> local stores;
//      `accumulate`: ADD them to my rows instead -- the adjoint of the exchange, for the sharded VJP)
//   4. the last block stores `epoch` into the neighbours' DONE flags (they may overwrite those rows again)
//
// No host thread of another rank is involved and nothing but this kernel sits between "own rows final" and the two
// O(halo) strip launches.  Flags are monotonically increasing 64-bit epochs, so a flag is never reset.
#include "kx_internal.cuh"

namespace kx {
namespace {

constexpr int kPullThreads = 512;

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// peer memory is not kept coherent in this SM's L1: bypass it
__device__ __forceinline__ uint4 ld_peer16(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// spin until *flag >= value; a neighbour that never arrives (crashed rank) traps after `timeout_ns` instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const uint64_t* flag, uint64_t value, uint64_t timeout_ns) {
  const uint64_t t0 = global_ns();
  unsigned backoff = 32;
  while (ld_acquire_sys(flag) < value) {
    __nanosleep(backoff);
    if (backoff < 1024) backoff <<= 1;
    if (global_ns() - t0 > timeout_ns) {
      printf("kx_halo_pull: timed out waiting for flag %p >= %llu (neighbouring rank missing?)\n", (const void*)flag,
             (unsigned long long)value);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(kPullThreads) halo_pull_kernel(const __grid_constant__ KxHaloPull p) {
  const int tid = threadIdx.x;
  const uint64_t timeout_ns = 20ull * 1000000000ull;
  if (blockIdx.x == 0 && tid < p.n_sides && p.side[tid].peer_ready) {
    __threadfence_system();
    st_release_sys(p.side[tid].peer_ready, p.epoch);
  }
  if (tid < p.n_sides && p.side[tid].bytes > 0) wait_flag(p.side[tid].my_ready, p.epoch, timeout_ns);
  __syncthreads();
#pragma unroll 1
  for (int s = 0; s < p.n_sides; ++s) {
    const int64_t n16 = p.side[s].bytes >> 4;
    const uint4* src = (const uint4*)p.side[s].src;
    uint4* dst = (uint4*)p.side[s].dst;
    // 4 independent 16-byte peer loads in flight per thread: an NVLink round trip is ~2 us
    int64_t i = (int64_t)blockIdx.x * kPullThreads * 4 + tid;
    const int64_t step = (int64_t)gridDim.x * kPullThreads * 4;
    for (; i < n16; i += step) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * kPullThreads < n16) v[u] = ld_peer16(src + i + u * kPullThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * kPullThreads < n16) {
          if (p.accumulate) {   // adjoint exchange: the pulled cotangent rows are ADDED to my edge rows (fp32)
            const uint4 o = dst[i + u * kPullThreads];
            v[u].x = __float_as_uint(__uint_as_float(o.x) + __uint_as_float(v[u].x));
            v[u].y = __float_as_uint(__uint_as_float(o.y) + __uint_as_float(v[u].y));
            v[u].z = __float_as_uint(__uint_as_float(o.z) + __uint_as_float(v[u].z));
            v[u].w = __float_as_uint(__uint_as_float(o.w) + __uint_as_float(v[u].w));
          }
          dst[i + u * kPullThreads] = v[u];
        }
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned done = atomicAdd(p.counter, 1u);
    if (done == gridDim.x - 1) {
      *p.counter = 0;   // next launch on this stream starts from zero (launches of one rank are stream-ordered)
      __threadfence_system();
      for (int s = 0; s < p.n_sides; ++s)
        if (p.side[s].peer_done) st_release_sys(p.side[s].peer_done, p.epoch);
    }
  }
}

__global__ void flag_wait_kernel(const uint64_t* flag, uint64_t value) { wait_flag(flag, value, 20ull * 1000000000ull); }

}  // namespace
}  // namespace kx

using namespace kx;

extern "C" {

int kx_peer_alloc(int64_t bytes, void** ptr) {
  if (!ptr || bytes <= 0) return fail(KX_ERR_INVALID, "kx_peer_alloc: bytes=%lld", (long long)bytes);
  KX_CUDA_CHECK(cudaMalloc(ptr, (size_t)bytes));
  KX_CUDA_CHECK(cudaMemset(*ptr, 0, (size_t)bytes));
  KX_CUDA_CHECK(cudaDeviceSynchronize());
  return KX_OK;
}

int kx_peer_free(void* ptr) {
  if (ptr) KX_CUDA_CHECK(cudaFree(ptr));
  return KX_OK;
}

int kx_peer_export(const void* ptr, void* handle64) {
  if (!ptr || !handle64) return fail(KX_ERR_INVALID, "kx_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == KX_PEER_HANDLE_BYTES, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  KX_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, sizeof(h));
  return KX_OK;
}

int kx_peer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(KX_ERR_INVALID, "kx_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  KX_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KX_OK;
}

int kx_peer_close(void* ptr) {
  if (ptr) KX_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return KX_OK;
}

int kx_halo_pull(const KxHaloPull* p, void* stream) {
  if (!p || p->n_sides < 0 || p->n_sides > 2 || !p->counter) return fail(KX_ERR_INVALID, "kx_halo_pull: malformed request");
  int64_t total = 0;
  for (int s = 0; s < p->n_sides; ++s) {
    const KxHaloSide& sd = p->side[s];
    if (sd.bytes < 0 || (sd.bytes & 15) || (sd.bytes > 0 && (!sd.dst || !sd.src || !sd.my_ready)) ||
        (((uintptr_t)sd.dst | (uintptr_t)sd.src) & 15))
      return fail(KX_ERR_INVALID, "kx_halo_pull: side %d needs 16-byte aligned rows and a multiple of 16 bytes (got %lld)", s,
                  (long long)sd.bytes);
    total += sd.bytes;
  }
  if (p->n_sides == 0) return KX_OK;
  DeviceGuard guard(p->counter);
  // enough blocks to keep ~1 MB of NVLink loads in flight, few enough to leave the SMs to the interior launch
  int64_t blocks = (total / 16 + kPullThreads * 4 - 1) / (kPullThreads * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  halo_pull_kernel<<<(unsigned)blocks, kPullThreads, 0, (cudaStream_t)stream>>>(*p);
  return after_launch("halo_pull_kernel");
}

int kx_flag_wait(const uint64_t* flag, uint64_t value, void* stream) {
  if (!flag) return fail(KX_ERR_INVALID, "kx_flag_wait: null flag");
  DeviceGuard guard(flag);
  flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
  return after_launch("flag_wait_kernel");
}

}  // extern "C"
