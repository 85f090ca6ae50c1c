// TMA bulk-copy (cp.async.bulk = UBLKCP, no tensor map) + mbarrier primitives shared by the
// row-streaming kernels (kx_conv2d.cu, kx_pool2d.cu).
#pragma once

#include <stdint.h>

namespace kx {

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ void lds_v4(unsigned addr, T (&d)[4]) {
  int4 raw;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(addr));
  d[0] = reinterpret_cast<const T&>(raw.x);
  d[1] = reinterpret_cast<const T&>(raw.y);
  d[2] = reinterpret_cast<const T&>(raw.z);
  d[3] = reinterpret_cast<const T&>(raw.w);
}
template <typename T>
__device__ __forceinline__ T lds_1(unsigned addr) {
  int raw;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(addr));
  return reinterpret_cast<const T&>(raw);
}
__device__ __forceinline__ void sts_zero16(unsigned addr) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0) : "memory");
}


}  // namespace kx
