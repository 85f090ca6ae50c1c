// Fast path: conv-as-kmap with MANY channels -- a (C, 3, 3) window over a channel-first image whose window spans the
// whole channel axis (`kmap(kernel_size=(C,3,3), padding=("valid","same","same"))(lambda x, w: sum(x*w))`, the
// reference's own benchmark tests_and_benchmarks/test_benchmark_conv2d.py:21-44 with C in {4, 8, 16, 32, 64}), C >= 5.
// C <= 4 stays on the per-warp row ring of kx_conv2d.cu (BASELINE config 3a); before this kernel C >= 5 ran on
// generic_map_kernel, and C = 64 (576 taps) did not fit KxProgram at all.
//
// The channel axis is the MARCHING axis of a 2.5-D blocking (the geometry of kx_stencil3d_star.cu): a block (8 warps)
// owns an 8-row x 256-column tile of the single output plane and walks through the C input planes; every plane tile
// (10 rows x 264 columns) is fetched once with 16-byte cp.async.cg (zero-fill = the reference's zero padding) into a
// 3-stage shared-memory ring two planes ahead of the math; a thread (4 columns x 2 rows) reads its 4 rows x 6 columns
// of the plane once and accumulates the plane's 9 taps into its 8 outputs, which stay in registers until the last
// plane: 4 B read per input element + 4 B written per output, (C * 4 + 4) bytes per lattice update.  fp32 uses packed
// FFMA2 (two outputs per instruction, the weight is a broadcast scalar).  The C * 9 weights (host table or device
// array) are staged in shared memory once per block and read back as broadcasts, 9 per plane.
// Accumulation order of one output: bias, then the window in row-major (c, ky, kx) order.
#include <type_traits>

#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kCWX = 2, kCWarps = 8, kCRY = 2;
constexpr int kCX = 128 * kCWX, kCY = kCWarps / kCWX * kCRY;   // tile 8 x 256
constexpr int kCThreads = 32 * kCWarps;
constexpr int kCStages = 3;
constexpr int kCMaxC = KX_MAX_TAPS / 9;

template <typename T>
struct ConvCArgs {
  const T* in;
  T* out;
  const T* coef_dev;        // optional [C * 9] device weights (row-major c, ky, kx)
  int64_t in_batch, out_batch;
  int C, hin, win, hout, wout;
  int ly, lx;
  int tiles_x, ntiles;
  T bias;
  T coef[kCMaxC * 9];       // host weights (row-major c, ky, kx), absent taps = 0 with `mask` bit clear
  uint32_t full;            // 1: every (c, ky, kx) tap is present -> no per-tap mask tests
  uint32_t mask[kCMaxC];    // 9 bits per channel
};

__device__ __forceinline__ void c_cp_async16(unsigned smem_addr, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void c_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void c_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
__device__ __forceinline__ void c_lds4(const T* p, T (&d)[4]) {
  const int4 raw = *reinterpret_cast<const int4*>(p);
  d[0] = reinterpret_cast<const T&>(raw.x);
  d[1] = reinterpret_cast<const T&>(raw.y);
  d[2] = reinterpret_cast<const T&>(raw.z);
  d[3] = reinterpret_cast<const T&>(raw.w);
}

template <typename T, int SHIFT, bool FULL>
__global__ void __launch_bounds__(kCThreads)
convc_kernel(const __grid_constant__ ConvCArgs<T> a) {
  constexpr int NV = (SHIFT + 6 + 3) / 4;          // aligned 4-vectors covering the 6 needed columns
  constexpr int SW = kCX + 4 * NV;                  // tile width (columns), multiple of 4
  constexpr int SR = kCY + 2;
  constexpr int CH = SW / 4;
  constexpr int NCH = (SR * CH + kCThreads - 1) / kCThreads;   // 16-byte chunks per thread per plane
  __shared__ __align__(16) T smem[kCStages * SR * SW];
  __shared__ T wsm[kCMaxC * 9];

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int b = blockIdx.x / a.ntiles;
  const int tile = blockIdx.x - b * a.ntiles;
  const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
  const int x0 = tile_x * kCX, y0 = tile_y * kCY;
  const T* __restrict__ in = a.in + (int64_t)b * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)b * a.out_batch;
  const int nplanes = a.C;
  const int64_t plane_elems = (int64_t)a.hin * a.win;

  // weights -> shared memory (read back as broadcasts)
  for (int i = tid; i < a.C * 9; i += kCThreads) wsm[i] = a.coef_dev ? a.coef_dev[i] : a.coef[i];

  int src_off[NCH];        // element offset inside a plane, or -1 = outside the array (zero fill)
  unsigned dst_off[NCH];   // byte offset inside a stage
  {
    const int gx0 = x0 - a.lx - SHIFT;             // global column of tile column 0 (multiple of 4)
    const int gy0 = y0 - a.ly;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int i = tid + k * kCThreads;
      const int r = i / CH, c = (i - r * CH) * 4;
      const int gy = gy0 + r, gx = gx0 + c;
      const bool ok = (i < SR * CH) && (gy >= 0) && (gy < a.hin) && (gx >= 0) && (gx + 3 < a.win);
      src_off[k] = ok ? gy * a.win + gx : -1;
      dst_off[k] = (i < SR * CH) ? (unsigned)((r * SW + c) * sizeof(T)) : 0xffffffffu;
    }
  }
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);

  auto issue = [&](int pl) {
    const T* src = in + (int64_t)pl * plane_elems;
    const unsigned dst = smem_base + (unsigned)((pl % kCStages) * SR * SW * sizeof(T));
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      if (dst_off[k] != 0xffffffffu) {
        const bool ok = src_off[k] >= 0;
        c_cp_async16(dst + dst_off[k], ok ? (const void*)(src + src_off[k]) : (const void*)in, ok ? 16 : 0);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < kCStages - 1; ++s) {
    if (s < nplanes) issue(s);
    c_cp_commit();
  }

  const int jx = 4 * tx + 128 * (ty % kCWX);   // thread's first output column inside the tile
  const int iy = (ty / kCWX) * kCRY;            // thread's first output row inside the tile
  const int ox = x0 + jx;
  const int n_valid = min(4, a.wout - ox);
  const bool col_ok = ox < a.wout;

  T acc[kCRY][4];
#pragma unroll
  for (int r = 0; r < kCRY; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = a.bias;

  auto step = [&](auto phase, int pl) {
    constexpr int PH = decltype(phase)::value;       // pl % kCStages, compile time
    c_cp_wait<kCStages - 2>();
    __syncthreads();            // plane pl landed for all threads (and, first trip, the weights); the stage read last step is free
    if (pl + kCStages - 1 < nplanes) issue(pl + kCStages - 1);
    c_cp_commit();
    // rows iy .. iy+RY+1 of the plane tile: the 6 window columns SHIFT+jx .. SHIFT+jx+5 of each
    T P[kCRY + 2][6];
    {
      const T* t = smem + PH * SR * SW + iy * SW + jx;
#pragma unroll
      for (int r = 0; r < kCRY + 2; ++r) {
        T w[4 * NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          T d[4];
          c_lds4(t + r * SW + 4 * q, d);
#pragma unroll
          for (int e = 0; e < 4; ++e) w[4 * q + e] = d[e];
        }
#pragma unroll
        for (int e = 0; e < 6; ++e) P[r][e] = w[SHIFT + e];
      }
    }
    T wk[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = wsm[pl * 9 + t];
    const uint32_t m = FULL ? 0x1ffu : a.mask[pl];
#pragma unroll
    for (int r = 0; r < kCRY; ++r) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if (FULL || (m & (1u << (ky * 3 + kx)))) {
            const T w1 = wk[ky * 3 + kx];
            if constexpr (std::is_same<T, float>::value) {   // two packed FMAs instead of four scalar ones
              ffma2(acc[r][0], acc[r][1], w1, w1, P[r + ky][kx], P[r + ky][kx + 1]);
              ffma2(acc[r][2], acc[r][3], w1, w1, P[r + ky][kx + 2], P[r + ky][kx + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[r][i] = mad_wrap(w1, P[r + ky][kx + i], acc[r][i]);
            }
          }
        }
      }
    }
  };
  int pl = 0;
#pragma unroll 1
  for (; pl + 2 < nplanes; pl += 3) {   // unrolled by the ring depth: stage addresses are compile-time
    step(std::integral_constant<int, 0>{}, pl);
    step(std::integral_constant<int, 1>{}, pl + 1);
    step(std::integral_constant<int, 2>{}, pl + 2);
  }
  if (pl < nplanes) step(std::integral_constant<int, 0>{}, pl);
  if (pl + 1 < nplanes) step(std::integral_constant<int, 1>{}, pl + 1);

#pragma unroll
  for (int r = 0; r < kCRY; ++r) {
    if (col_ok && (y0 + iy + r < a.hout)) {
      T* o = out + (int64_t)(y0 + iy + r) * a.wout + ox;
      const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
      if (n_valid == 4 && (addr & 15) == 0) {
        int4 raw;
        raw.x = reinterpret_cast<const int&>(acc[r][0]);
        raw.y = reinterpret_cast<const int&>(acc[r][1]);
        raw.z = reinterpret_cast<const int&>(acc[r][2]);
        raw.w = reinterpret_cast<const int&>(acc[r][3]);
        __stcs(reinterpret_cast<int4*>(o), raw);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < n_valid) o[i] = acc[r][i];
      }
    }
  }
}

template <typename T>
int run_convc(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int ac = g.ndim - 3, ay = g.ndim - 2, ax = g.ndim - 1;
  const int C = g.ksize[ac];
  static thread_local ConvCArgs<T> a;
  memset(&a, 0, sizeof(a));
  int n_present = 0;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int c = p.tap_off[t][ac], slot = p.tap_off[t][ay] * 3 + p.tap_off[t][ax];
    if (a.mask[c] & (1u << slot)) return KX_ERR_UNSUPPORTED;          // duplicate tap: generic kernel
    a.mask[c] |= 1u << slot;
    a.coef[c * 9 + slot] = (T)p.coef[t];
    // device weights are indexed by TAP number: only the identity layout (tap t <-> c*9 + slot) is taken
    if (m.coef_dev && (t - f.tap_begin) != c * 9 + slot) return KX_ERR_UNSUPPORTED;
    ++n_present;
  }
  a.full = n_present == C * 9 ? 1u : 0u;
  if (m.coef_dev && !a.full) return KX_ERR_UNSUPPORTED;
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.coef_dev = m.coef_dev ? (const T*)m.coef_dev + f.tap_begin : nullptr;
  a.C = C;
  a.hin = (int)g.in_shape[ay];
  a.win = (int)g.in_shape[ax];
  a.hout = (int)g.out_shape[ay];
  a.wout = (int)g.out_shape[ax];
  a.ly = g.lo[ay];
  a.lx = g.lo[ax];
  if ((int64_t)a.hin * a.win > 0x7ffffff0ll) return KX_ERR_UNSUPPORTED;  // 32-bit in-plane offsets
  a.in_batch = (int64_t)C * a.hin * a.win;
  a.out_batch = (int64_t)a.hout * a.wout;
  a.bias = (T)f.bias;
  a.tiles_x = (a.wout + kCX - 1) / kCX;
  const int tiles_y = (a.hout + kCY - 1) / kCY;
  a.ntiles = a.tiles_x * tiles_y;
  const int64_t nblocks = (int64_t)a.ntiles * batch;
  if (nblocks > 0x7fffffffll || nblocks < 1) return KX_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nblocks), block(32, kCWarps);
  const int shift = (((-a.lx) % 4) + 4) % 4;
#define KX_CONVC(S)                                                   \
  if (a.full) convc_kernel<T, S, true><<<grid, block, 0, s>>>(a);     \
  else convc_kernel<T, S, false><<<grid, block, 0, s>>>(a);
  switch (shift) {
    case 0: KX_CONVC(0) break;
    case 1: KX_CONVC(1) break;
    case 2: KX_CONVC(2) break;
    default: KX_CONVC(3) break;
  }
#undef KX_CONVC
  return after_launch("convc_kernel");
}

}  // namespace

// (C, 3, 3) windows spanning the whole channel axis (output depth 1), C >= 5, stride 1, rows 16-byte aligned.
int try_convc(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  if (nd < 3 || p.n_funcs != 1 || p.func[0].kind != KX_AFFINE || p.func[0].post_div != 0.0) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype || m.fn_elems != 1 || m.use_box) return KX_ERR_UNSUPPORTED;
  const int ac = nd - 3;
  const int C = g.ksize[ac];
  if (C < 5 || C > kCMaxC || C != g.in_shape[ac] || g.lo[ac] != 0 || g.hi[ac] != 0 || g.stride[ac] != 1) return KX_ERR_UNSUPPORTED;
  if (g.ksize[nd - 2] != 3 || g.ksize[nd - 1] != 3 || g.stride[nd - 2] != 1 || g.stride[nd - 1] != 1) return KX_ERR_UNSUPPORTED;
  int64_t lead = 1;
  for (int d = 0; d < nd - 3; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    lead *= g.in_shape[d];
  }
  const int64_t batch = m.batch * lead;
  if (m.coef_dev && batch > 1 && m.coef_batch_stride != 0) return KX_ERR_UNSUPPORTED;
  if (g.in_shape[nd - 1] % 4 != 0 || (reinterpret_cast<uintptr_t>(m.in) & 15) != 0) return KX_ERR_UNSUPPORTED;
  if (g.in_shape[nd - 1] > 0x3fffffff || g.in_shape[nd - 2] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
  if (getenv("KX_NO_CONVC")) return KX_ERR_UNSUPPORTED;   // cross-check: the generic kernel
  if (g.in_dtype == KX_F32) return run_convc<float>(m, p, batch, s);
  if (g.in_dtype == KX_I32 && p.func[0].coef_is_int) return run_convc<int32_t>(m, p, batch, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
