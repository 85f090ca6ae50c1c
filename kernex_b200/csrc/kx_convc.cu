// Fast path: conv-as-kmap with MANY channels -- a (C, 3, 3) window over a channel-first image whose window spans the
// whole channel axis (`kmap(kernel_size=(C,3,3), padding=("valid","same","same"))(lambda x, w: sum(x*w))`, the
// reference's own benchmark tests_and_benchmarks/test_benchmark_conv2d.py:21-44 with C in {4, 8, 16, 32, 64}), C >= 5.
// C <= 4 stays on the per-warp row ring of kx_conv2d.cu (BASELINE config 3a); before this kernel C >= 5 ran on
// generic_map_kernel, and C = 64 (576 taps) did not fit KxProgram at all.
//
// The channel axis is the MARCHING axis of a 2.5-D blocking (the geometry of kx_stencil3d_star.cu): a block (8 warps)
// owns an 8-row x 256-column tile of the single output plane and walks through the C input planes; every plane tile
// (10 rows x 264 columns) is fetched once -- one TMA bulk copy per row segment, clipped to the row; the ring starts out
// as zeros, which is the reference's zero padding -- into a 3-stage (small grids: 6-stage) shared-memory ring ahead of
// the math; a thread (4 columns x 2 rows) reads its 4 rows x 6 columns of the plane once and accumulates the plane's
// 9 taps into its 8 outputs, which stay in registers until the last plane: 4 B read per input element + 4 B written per output, (C * 4 + 4) bytes per lattice update.  fp32 uses packed
// FFMA2 (two outputs per instruction, the weight is a broadcast scalar).  The C * 9 weights (host table or device
// array) are staged in shared memory once per block and read back as broadcasts, 9 per plane.
// Accumulation order of one output: bias, then the window in row-major (c, ky, kx) order.
// ncu at (16, 4096^2), 0.77 of the HBM roofline: the shared-memory pipe is the limiter (l1tex 92 % of its peak, short
// scoreboard + MIO throttle the top stalls: 12 LDS.128 of plane data + 3 of weights per plane and thread).  Measured and
// rejected: one own vector per row + the neighbours' columns by __shfl_up / __shfl_down (edge lanes read the adjacent
// vector): 4 MIO instructions per row instead of 3 -> 0.585; 8 contiguous columns per lane would halve the LDS count
// but puts lanes 32 B apart (2-way bank conflicts on LDS.128).
#include <type_traits>

#include "kx_internal.cuh"
#include "kx_tma.cuh"

namespace kx {

namespace {

constexpr int kCWX = 2, kCWarps = 8, kCRY = 2;
constexpr int kCX = 128 * kCWX, kCY = kCWarps / kCWX * kCRY;   // tile 8 x 256
constexpr int kCThreads = 32 * kCWarps;
constexpr int kCMaxC = KX_MAX_TAPS / 9;      // channels at 3x3; K x K windows: C * K * K <= KX_MAX_TAPS
constexpr int kCWsm = 1536;                    // floats of staged weights: C * (K * K rounded up to 4) for every admissible (C, K)

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
  T coef[KX_MAX_TAPS];      // host weights (row-major c, ky, kx), absent taps = 0 with `mask` bit clear
  uint32_t full;            // 1: every (c, ky, kx) tap is present -> no per-tap mask tests
  uint64_t mask[kCMaxC];    // K * K bits per channel
};

template <int I, int N, typename F>
__device__ __forceinline__ void c_static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    c_static_for<I + 1, N>(f);
  }
}

template <typename T>
__device__ __forceinline__ void c_lds4(const T* p, T (&d)[4]) {
  const int4 raw = *reinterpret_cast<const int4*>(p);
  d[0] = reinterpret_cast<const T&>(raw.x);
  d[1] = reinterpret_cast<const T&>(raw.y);
  d[2] = reinterpret_cast<const T&>(raw.z);
  d[3] = reinterpret_cast<const T&>(raw.w);
}

template <typename T, int SHIFT, bool FULL, int kCStages, int K>
__global__ void __launch_bounds__(kCThreads)
convc_kernel(const __grid_constant__ ConvCArgs<T> a) {
  constexpr int NC = 4 + K - 1;                    // window columns a thread's 4 outputs touch
  constexpr int KK = K * K, KP = (KK + 3) / 4 * 4;  // weights per plane, padded to whole LDS.128
  constexpr int NV = (SHIFT + NC + 3) / 4;         // aligned 4-vectors covering them
  constexpr int SW = kCX + 4 * NV;                  // tile width (columns), multiple of 4
  constexpr int SR = kCY + K - 1;
  extern __shared__ __align__(128) unsigned char dyn_smem[];   // the ring: kCStages * SR * SW elements (dynamic: 6 stages exceed 48 KB)
  T* smem = reinterpret_cast<T*>(dyn_smem);
  __shared__ __align__(16) T wsm[kCWsm];             // K * K weights per plane, padded to KP: KP / 4 LDS.128 per plane
  __shared__ __align__(8) uint64_t bars[kCStages];

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int b = blockIdx.x / a.ntiles;
  const int tile = blockIdx.x - b * a.ntiles;
  const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
  const int x0 = tile_x * kCX, y0 = tile_y * kCY;
  const T* __restrict__ in = a.in + (int64_t)b * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)b * a.out_batch;
  const int nplanes = a.C;
  const int64_t plane_elems = (int64_t)a.hin * a.win;

  // weights -> shared memory (read back as broadcasts); the ring starts out as zeros: rows above / below the array and
  // the columns clipped off the row copies are the same in every plane, so they are never written again = zero padding
  for (int i = tid; i < a.C * KP; i += kCThreads) {
    const int c = i / KP, t = i - c * KP;
    wsm[i] = t < KK ? (a.coef_dev ? a.coef_dev[c * KK + t] : a.coef[c * KK + t]) : T(0);
  }
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
  for (int i = tid; i < kCStages * SR * SW / 4; i += kCThreads) sts_zero16(smem_base + (unsigned)(i * 16));
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);
  if (tid < kCStages) mbar_init(bar0 + tid * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncthreads();

  // A plane tile = SR row segments of SW columns, each ONE TMA bulk copy (clipped to the row) issued by thread 0 and
  // completing on the stage's mbarrier: no LDGSTS (ptxas pads every LDGSTS with three dummy LDS; ncu of the cp.async
  // version: l1tex at 88 % of its peak, MIO throttle the top stall), no per-thread chunk bookkeeping.
  const int gx0 = x0 - a.lx - SHIFT;               // global column of tile column 0 (multiple of 4)
  const int gy0 = y0 - a.ly;
  const int c_lo = max(gx0, 0), c_hi = min(gx0 + SW, a.win);
  const unsigned row_bytes = c_hi > c_lo ? (unsigned)((c_hi - c_lo) * 4) : 0u;
  const int r_lo = max(0, -gy0), r_hi = min(SR, a.hin - gy0);      // tile rows inside the array
  const unsigned dst_col = (unsigned)((c_lo - gx0) * 4);
  auto issue = [&](int pl) {                        // thread 0 only
    const unsigned bar = bar0 + (unsigned)((pl % kCStages) * 8);
    const unsigned st = smem_base + (unsigned)((pl % kCStages) * SR * SW * 4);
    if (row_bytes == 0 || r_hi <= r_lo) {
      mbar_arrive(bar);
      return;
    }
    mbar_expect_tx(bar, row_bytes * (unsigned)(r_hi - r_lo));
    const T* src = in + (int64_t)pl * plane_elems + (int64_t)(gy0 + r_lo) * a.win + c_lo;
    for (int r = r_lo; r < r_hi; ++r, src += a.win) bulk_g2s(st + (unsigned)(r * SW * 4) + dst_col, src, row_bytes, bar);
  };
  if (tid == 0) {
#pragma unroll
    for (int s_ = 0; s_ < kCStages - 1; ++s_)
      if (s_ < nplanes) issue(s_);
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
    __syncthreads();            // everybody is done reading the stage that is refilled now
    if (tid == 0 && pl + kCStages - 1 < nplanes) issue(pl + kCStages - 1);
    mbar_wait(bar0 + PH * 8, (unsigned)((pl / kCStages) & 1));      // plane pl has landed
    // rows iy .. iy+RY+K-2 of the plane tile: the NC window columns SHIFT+jx .. SHIFT+jx+NC-1 of each
    T P[kCRY + K - 1][NC];
    {
      const T* t = smem + PH * SR * SW + iy * SW + jx;
#pragma unroll
      for (int r = 0; r < kCRY + K - 1; ++r) {
        T w[4 * NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          T d[4];
          c_lds4(t + r * SW + 4 * q, d);
#pragma unroll
          for (int e = 0; e < 4; ++e) w[4 * q + e] = d[e];
        }
#pragma unroll
        for (int e = 0; e < NC; ++e) P[r][e] = w[SHIFT + e];
      }
    }
    T wk[KP];
#pragma unroll
    for (int q = 0; q < KP / 4; ++q) {
      T d[4];
      c_lds4(wsm + pl * KP + 4 * q, d);
#pragma unroll
      for (int e = 0; e < 4; ++e) wk[4 * q + e] = d[e];
    }
    const uint64_t m = FULL ? ~0ull : a.mask[pl];
#pragma unroll
    for (int r = 0; r < kCRY; ++r) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          if (FULL || (m & (1ull << (ky * K + kx)))) {
            const T w1 = wk[ky * K + kx];
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
  for (; pl + kCStages <= nplanes; pl += kCStages)   // unrolled by the ring depth: stage addresses are compile-time
    c_static_for<0, kCStages>([&](auto ph) { step(ph, pl + decltype(ph)::value); });
  c_static_for<0, kCStages>([&](auto ph) {
    if (pl + decltype(ph)::value < nplanes) step(ph, pl + decltype(ph)::value);
  });

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

template <typename T, int K>
int run_convc(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  constexpr int KK = K * K;
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int ac = g.ndim - 3, ay = g.ndim - 2, ax = g.ndim - 1;
  const int C = g.ksize[ac];
  static thread_local ConvCArgs<T> a;
  memset(&a, 0, sizeof(a));
  int n_present = 0;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int c = p.tap_off[t][ac], slot = p.tap_off[t][ay] * K + p.tap_off[t][ax];
    if (a.mask[c] & (1ull << slot)) return KX_ERR_UNSUPPORTED;        // duplicate tap: generic kernel
    a.mask[c] |= 1ull << slot;
    a.coef[c * KK + slot] = (T)p.coef[t];
    // device weights are indexed by TAP number: only the identity layout (tap t <-> c*9 + slot) is taken
    if (m.coef_dev && (t - f.tap_begin) != c * KK + slot) return KX_ERR_UNSUPPORTED;
    ++n_present;
  }
  a.full = n_present == C * KK ? 1u : 0u;
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
  // ring depth: 3 stages (5 blocks / SM); 6 stages (3 blocks / SM) only when the whole grid is resident anyway -- there
  // a block's C-plane march is one dependent chain of DRAM round trips and depth is all that hides it (measured,
  // benchmarks/convc_probe.py: 4-10 % at <= 512^2, but 1.6x SLOWER at 1024^2 where it costs a second wave)
  int stages = (K == 3 && nblocks <= (int64_t)sm_count() * 3) ? 6 : 3;
  if (const char* e = getenv("KX_CONVC_STAGES")) stages = (atoi(e) == 6 && K == 3) ? 6 : 3;   // tuning knob
  const int sw = kCX + 4 * ((shift + 4 + K - 1 + 3) / 4);
  const size_t ring = (size_t)stages * (kCY + K - 1) * sw * sizeof(T);
#define KX_CONVC2(S, FULLV, ST)                                                                                  \
  {                                                                                                              \
    auto kern = convc_kernel<T, S, FULLV, ST, K>;                                                                   \
    if (ring + 8192 > 48 * 1024) KX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring)); \
    kern<<<grid, block, ring, s>>>(a);                                                                           \
  }
#define KX_CONVC(S)                                                    \
  if (a.full && stages == 6) KX_CONVC2(S, true, 6)                      \
  else if (a.full) KX_CONVC2(S, true, 3)                               \
  else if (stages == 6) KX_CONVC2(S, false, 6)                          \
  else KX_CONVC2(S, false, 3)
  switch (shift) {
    case 0: KX_CONVC(0) break;
    case 1: KX_CONVC(1) break;
    case 2: KX_CONVC(2) break;
    default: KX_CONVC(3) break;
  }
#undef KX_CONVC
#undef KX_CONVC2
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
  const int K = g.ksize[nd - 1];
  if (g.ksize[nd - 2] != K || (K != 3 && K != 5 && K != 7) || g.stride[nd - 2] != 1 || g.stride[nd - 1] != 1) return KX_ERR_UNSUPPORTED;
  // 3x3 windows with C <= 4 belong to the per-warp row ring of kx_conv2d.cu; wider windows take every C >= 2
  if (C < (K == 3 ? 5 : 2) || C * K * K > KX_MAX_TAPS || C > kCMaxC || C * ((K * K + 3) / 4 * 4) > kCWsm) return KX_ERR_UNSUPPORTED;
  if (C != g.in_shape[ac] || g.lo[ac] != 0 || g.hi[ac] != 0 || g.stride[ac] != 1) return KX_ERR_UNSUPPORTED;
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
  if (g.in_dtype == KX_F32) {
    if (K == 3) return run_convc<float, 3>(m, p, batch, s);
    if (K == 5) return run_convc<float, 5>(m, p, batch, s);
    return run_convc<float, 7>(m, p, batch, s);
  }
  if (g.in_dtype == KX_I32 && p.func[0].coef_is_int) {
    if (K == 3) return run_convc<int32_t, 3>(m, p, batch, s);
    if (K == 5) return run_convc<int32_t, 5>(m, p, batch, s);
    return run_convc<int32_t, 7>(m, p, batch, s);
  }
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
