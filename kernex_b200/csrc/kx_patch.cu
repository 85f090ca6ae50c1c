// Fast path: GATHER programs (patch extraction, kmap(lambda x: x); the reference's
// equivalent of lax.conv_general_dilated_patches, README.md:157-186,
// tests_and_benchmarks/test_benchmark_patch.py).
//
// The output [n_win, n_taps] is 9x (3x3) larger than the input, so the kernel is
// WRITE-bound: 4 B read + 4*n_taps B written per window.  Each thread produces four
// consecutive output floats (one 128-bit streaming store; the flat output offset of a
// thread is a multiple of 4 so stores are always 16-byte aligned and a warp writes 512
// contiguous bytes), and gathers its four sources through L1 -- neighbouring lanes read
// neighbouring input cells, every input line is fetched from L2 once per block row and
// then served n_taps times by L1.  Index decoding uses 32-bit arithmetic with a
// compile-time n_taps for the common windows.
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPatchTaps = 128;

struct PatchArgs {
  const void* in;
  void* out;
  int32_t ndim, n_taps;
  uint32_t out_shape[KX_MAX_DIMS];
  int32_t in_shape[KX_MAX_DIMS];
  int32_t stride[KX_MAX_DIMS], lo[KX_MAX_DIMS];
  int64_t in_stride[KX_MAX_DIMS];
  uint64_t total;                       // output elements
  int8_t tap[kMaxPatchTaps][KX_MAX_DIMS];
};

template <typename T, int NT, int ND>
__device__ __forceinline__ T gather_one(const PatchArgs& a, const T* __restrict__ in, uint32_t w, int t) {
  int64_t off = 0;
  bool inb = true;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    const uint32_t j = (d == 0) ? w : (w % a.out_shape[d]);
    if (d != 0) w /= a.out_shape[d];
    const int c = (int)j * a.stride[d] - a.lo[d] + a.tap[t][d];
    inb = inb && (c >= 0) && (c < a.in_shape[d]);
    off += (int64_t)c * a.in_stride[d];
  }
  return inb ? __ldg(in + off) : T(0);
}

template <typename T, int NT, int ND>
__global__ void __launch_bounds__(kThreads)
patch_kernel(const __grid_constant__ PatchArgs a) {
  const T* __restrict__ in = (const T*)a.in;
  T* __restrict__ out = (T*)a.out;
  const uint32_t nt = NT > 0 ? NT : (uint32_t)a.n_taps;
  const uint64_t i0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
  if (i0 >= a.total) return;
  uint32_t w = (uint32_t)(i0 / nt);
  int t = (int)(i0 - (uint64_t)w * nt);
  T v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[e] = (i0 + e < a.total) ? gather_one<T, NT, ND>(a, in, w, t) : T(0);
    if (++t == (int)nt) {
      t = 0;
      ++w;
    }
  }
  if (i0 + 3 < a.total) {
    int4 raw;
    raw.x = reinterpret_cast<const int&>(v[0]);
    raw.y = reinterpret_cast<const int&>(v[1]);
    raw.z = reinterpret_cast<const int&>(v[2]);
    raw.w = reinterpret_cast<const int&>(v[3]);
    __stcs(reinterpret_cast<int4*>(out + i0), raw);
  } else {
    for (int e = 0; e < 4 && i0 + e < a.total; ++e) out[i0 + e] = v[e];
  }
}

template <typename T, int ND>
int launch_nd(const PatchArgs& a, cudaStream_t s) {
  const uint64_t threads = (a.total + 3) / 4;
  const uint64_t blocks = (threads + kThreads - 1) / kThreads;
  if (blocks > 0x7fffffffull) return KX_ERR_UNSUPPORTED;
  switch (a.n_taps) {
    case 3: patch_kernel<T, 3, ND><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 4: patch_kernel<T, 4, ND><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 9: patch_kernel<T, 9, ND><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 25: patch_kernel<T, 25, ND><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 27: patch_kernel<T, 27, ND><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    default: patch_kernel<T, 0, ND><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
  }
  return after_launch("patch_kernel");
}

}  // namespace

int try_patch(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (p.n_funcs != 1 || p.func[0].kind != KX_GATHER) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype || m.batch != 1) return KX_ERR_UNSUPPORTED;
  const int nt = p.func[0].tap_end - p.func[0].tap_begin;
  if (nt < 1 || nt > kMaxPatchTaps) return KX_ERR_UNSUPPORTED;
  if (m.n_win > 0xfffffff0ll) return KX_ERR_UNSUPPORTED;            // 32-bit window index
  if ((reinterpret_cast<uintptr_t>(m.out) & 15) != 0) return KX_ERR_UNSUPPORTED;
  PatchArgs a;
  memset(&a, 0, sizeof(a));
  a.in = m.in;
  a.out = m.out;
  a.ndim = g.ndim;
  a.n_taps = nt;
  a.total = (uint64_t)m.n_win * nt;
  int64_t st = 1;
  for (int d = g.ndim - 1; d >= 0; --d) {
    if (g.in_shape[d] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
    a.out_shape[d] = (uint32_t)g.out_shape[d];
    a.in_shape[d] = (int32_t)g.in_shape[d];
    a.stride[d] = g.stride[d];
    a.lo[d] = g.lo[d];
    a.in_stride[d] = st;
    st *= g.in_shape[d];
  }
  for (int t = 0; t < nt; ++t)
    for (int d = 0; d < g.ndim; ++d) {
      const int o = p.tap_off[p.func[0].tap_begin + t][d];
      if (o > 127) return KX_ERR_UNSUPPORTED;
      a.tap[t][d] = (int8_t)o;
    }
  const bool f32 = g.in_dtype == KX_F32;
  switch (g.ndim) {
    case 1: return f32 ? launch_nd<float, 1>(a, s) : launch_nd<int32_t, 1>(a, s);
    case 2: return f32 ? launch_nd<float, 2>(a, s) : launch_nd<int32_t, 2>(a, s);
    case 3: return f32 ? launch_nd<float, 3>(a, s) : launch_nd<int32_t, 3>(a, s);
    case 4: return f32 ? launch_nd<float, 4>(a, s) : launch_nd<int32_t, 4>(a, s);
  }
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
