// Shared host/device helpers of libkx_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/kx_b200.h"

namespace kx {

// ---- thread-local diagnostics (no shared mutable state across host threads) ----
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
// process-wide (not thread-local): autograd runs backward passes on its own thread, and the tests / bench ask
// "which kernel ran last" from the main thread
inline std::atomic<const char*>& last_kernel_name() {
  static std::atomic<const char*> name{""};
  return name;
}
inline std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> n{0};
  return n;
}

inline int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return status;
}

#define KX_CUDA_CHECK(expr)                                                                 \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::kx::fail(KX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                \
  } while (0)

inline int after_launch(const char* name) {
  last_kernel_name().store(name, std::memory_order_relaxed);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(KX_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
  }
  return KX_OK;
}

// The launchers enqueue on the caller's stream; CUDA launches go to the calling thread's CURRENT device, so
// a caller that hands over buffers of another GPU (several GPUs in one process) would fault.  Every entry point
// switches to the device that owns the buffer for the duration of the call and restores the previous one.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const void* device_ptr) {
    cudaPointerAttributes at;
    int cur = -1;
    if (device_ptr && cudaPointerGetAttributes(&at, device_ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice &&
        cudaGetDevice(&cur) == cudaSuccess && cur != at.device) {
      if (cudaSetDevice(at.device) == cudaSuccess) prev = cur;
    }
    cudaGetLastError();  // a host / unregistered pointer is the caller's problem, reported by the launch itself
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

inline int64_t prod(const int64_t* v, int n) {
  int64_t p = 1;
  for (int i = 0; i < n; ++i) p *= v[i];
  return p;
}

// ---- device helpers ----
template <typename T>
struct Cast;
template <>
struct Cast<float> {
  __device__ __forceinline__ static float from(double v) { return (float)v; }
};
template <>
struct Cast<int32_t> {
  __device__ __forceinline__ static int32_t from(double v) { return (int32_t)(long long)v; }
};

// wrap-around int32 multiply-add (signed overflow is UB in C++, defined here)
__device__ __forceinline__ int32_t mad_wrap(int32_t c, int32_t v, int32_t acc) {
  return (int32_t)((uint32_t)c * (uint32_t)v + (uint32_t)acc);
}
__device__ __forceinline__ float mad_wrap(float c, float v, float acc) { return fmaf(c, v, acc); }

// window max / min: NaN propagates like np.max / jnp.max (a comparison-select would drop a NaN that is not first)
__device__ __forceinline__ float red_max(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float red_min(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ int32_t red_max(int32_t a, int32_t b) { return a > b ? a : b; }
__device__ __forceinline__ int32_t red_min(int32_t a, int32_t b) { return a < b ? a : b; }
__device__ __forceinline__ double red_max(double a, double b) { return (a != a || b != b) ? (a + b) : (a > b ? a : b); }
__device__ __forceinline__ double red_min(double a, double b) { return (a != a || b != b) ? (a + b) : (a < b ? a : b); }

// Packed fp32 FMA (sm_100 FFMA2): (d0, d1) = (a0 * b0 + d0, a1 * b1 + d1), each half rounded exactly like
// fmaf.  One issue slot for two FMAs: what the issue-bound kernels (27 FMA per conv output) need.  ptxas
// broadcasts a scalar operand (a0 == a1) without building a register pair.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}

// streaming (read-once / write-once) vector accesses: keep L1 for the halo re-reads
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace kx
