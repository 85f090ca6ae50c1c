// Fast path: STRIDED windows and window REDUCTIONS over the last two axes -- pooling (README.md:
// 327-337 average pooling, max / min pooling), strided sums and strided weighted stencils, and the
// stride-1 max / min filters that kx_stencil2d.cu (AFFINE only) does not cover.  Class (b) of the
// three kernel templates: "window reductions (sum / max / mean)".
//
// HBM-bound design (4 B read per input element + 4 B written per window):
//   * every WARP streams the input rows of its 128 (stride 1: 256) output columns through a private
//     shared-memory ring filled by TMA bulk copies (one elected lane, two half-row copies, clipped
//     to the row; the clipped-away part of the ring is zeroed once = the reference's zero padding,
//     kernex/_src/utils.py:43-53; rows outside the array are zero-filled by the lanes); copies
//     run RING-1 rows ahead and complete on one mbarrier per stage, so the only synchronisation
//     is __syncwarp and the bytes in flight cost no registers;
//   * the kernel is input-row stationary: input row iy is read from shared memory exactly once
//     (NV aligned LDS.128 per lane; the window misalignment (-lo_x) mod 4 is a template parameter)
//     and folded into the ceil(KY/SY) output rows it belongs to (output row r takes window row
//     ky = iy + lo_y - r*SY), so there is no input row queue; the row loop is unrolled by the
//     period of that pattern so ring stages and accumulator slots are compile-time constants;
//   * lanes own 4 CONSECUTIVE output columns for stride 1 (aligned LDS.128, 128-bit stores) and 4
//     INTERLEAVED columns (lane, lane+32, ...) for stride > 1: consecutive columns would put the
//     lanes 32 bytes apart in shared memory (ncu: 8.7-way bank conflicts, the shared-memory port
//     starved the TMA writes and the kernel ran at 0.35 of peak); interleaved, a warp's LDS.64 /
//     LDS.32 are conflict-free and its 32-bit stores still cover whole 128-byte lines;
//   * accumulation order of one output = bias, then the window in row-major order (ky, kx).
#include <type_traits>

#include "kx_internal.cuh"
#include "kx_tma.cuh"

namespace kx {

namespace {

constexpr int kPT = 128;   // threads per block (4 warps), 4 output columns per thread

enum PoolOp { kOpAffine = 0, kOpMax = 1, kOpMin = 2 };

template <typename T>
struct PoolArgs {
  const T* in;
  T* out;
  int64_t in_batch, out_batch;
  int hin, win, hout, wout;
  int ly, lx, th;
  float post_div;
  T bias;
  T coef[32];
};

// A warp covers kWarpOut<SX> output columns: 256 for stride 1 (8 per lane, two groups of 4 consecutive
// columns 128 apart), 128 otherwise, so that an input row of the warp is ~1 KB = two ~512-byte bulk
// copies, the size the TMA unit streams fastest (benchmarks/exp_tma.cu).
template <int SX>
struct WarpOut {
  static constexpr int groups = SX == 1 ? 2 : 1;
  static constexpr int value = 128 * groups;
};
template <int SX>
struct RowW {
  static constexpr int value = (((WarpOut<SX>::value - 1) * SX + 3 + 3) + 3) / 4 * 4 + 4;   // span + shift, rounded, + slack vector
};

template <int I, int N, typename F>
__device__ __forceinline__ void p_static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    p_static_for<I + 1, N>(f);
  }
}

template <typename T, int OP>
__device__ __forceinline__ T fold(T acc, T c, T v) {
  if (OP == kOpAffine) return mad_wrap(c, v, acc);
  if (OP == kOpMax) return red_max(acc, v);
  return red_min(acc, v);
}

template <typename T, int OP, int KY, int KX, int SY, int SX, int SHIFT>
__global__ void __launch_bounds__(kPT)
pool2d_kernel(const __grid_constant__ PoolArgs<T> a) {
  constexpr int NACC = (KY + SY - 1) / SY;               // output rows an input row can belong to
  constexpr int PERIOD = SY * NACC;
  constexpr int U = (PERIOD == 1 || PERIOD == 3) ? 3 : (PERIOD == 6 ? 6 : 4);  // unroll = ring depth, a multiple of PERIOD
  constexpr int RW = RowW<SX>::value;                    // floats per ring row
  constexpr int NE = SHIFT + 3 * SX + KX;                // ring columns a lane needs, from its aligned base
  constexpr int NV = (NE + 3) / 4;
  static_assert(U % PERIOD == 0, "unroll must cover the accumulator rotation");
  constexpr int NG = WarpOut<SX>::groups;                // groups of 4 outputs per lane
  constexpr int WO = WarpOut<SX>::value;                 // output columns per warp
  static_assert((4 * 31 + 128 * (NG - 1)) * SX + NV * 4 <= RW && (WO - 1) * SX + SHIFT + KX + 1 <= RW, "ring row too narrow");
  __shared__ __align__(128) T smem[(kPT / 32) * U * RW];
  __shared__ __align__(8) uint64_t bars[(kPT / 32) * U];

  constexpr bool kInter = SX > 1;                        // interleaved output columns (see header)
  constexpr int NL = SHIFT + KX;                         // floats a lane needs per output (interleaved)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wo0 = (blockIdx.x * (kPT / 32) + warp) * WO;           // first output column of the warp
  const int o0 = kInter ? wo0 + lane : wo0 + 4 * lane;             // this lane's first output column
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const T* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int iy0 = r0 * SY - a.ly, iy1 = (r1 - 1) * SY - a.ly + KY - 1;   // first / last input row of the band

  // ring column 0 <-> input column c_start (a multiple of 4; wo0 * SX is one, so the first window
  // column of output wo0 + j sits at ring column j * SX + SHIFT, SHIFT = (-lx) mod 4)
  const int c_start = wo0 * SX - a.lx - SHIFT;
  const int c_lo = max(c_start, 0), c_hi = min(c_start + RW, a.win);
  const unsigned bytes = (wo0 < a.wout && c_hi > c_lo) ? (unsigned)((c_hi - c_lo) * 4) : 0u;
  const unsigned row0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)(warp * U * RW * 4);
  const unsigned dst_off = (unsigned)((c_lo - c_start) * 4);
  const unsigned own = row0 + (unsigned)((kInter ? lane * SX : 4 * lane) * 4);
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(bars) + (unsigned)(warp * U * 8);
  const T* src = in + c_lo;
  const unsigned half = bytes >= 512 ? ((bytes / 2 + 15) & ~15u) : bytes;

  for (int i = lane; i < U * RW / 4; i += 32) sts_zero16(row0 + (unsigned)(i * 16));
  if (lane < U) mbar_init(mbar + lane * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncwarp();

  auto issue = [&](auto stage, int iy) {
    constexpr int S = decltype(stage)::value;
    const bool ok = (iy >= 0) && (iy < a.hin) && (iy <= iy1) && bytes != 0;
    if (lane == 0) {
      if (ok) {
        // two half-row copies: benchmarks/exp_tma.cu measures 2.5 TB/s for one 1 KB bulk copy per row
        // and 7.0 TB/s for the same bytes as two copies (per-copy pipelining inside the TMA unit)
        mbar_expect_tx(mbar + S * 8, bytes);
        const T* rp = src + (int64_t)iy * a.win;
        const unsigned dst = row0 + (unsigned)(S * RW * 4) + dst_off;
        bulk_g2s(dst, rp, half, mbar + S * 8);
        if (bytes > half) bulk_g2s(dst + half, rp + half / 4, bytes - half, mbar + S * 8);
      } else {
        mbar_arrive(mbar + S * 8);
      }
    }
    if (!ok) {  // a padding row: the warp clears the whole stage row
      for (int i = lane; i < RW / 4; i += 32) sts_zero16(row0 + (unsigned)(S * RW * 4 + i * 16));
      fence_proxy_async();
      __syncwarp();
    }
  };

  T acc[NACC][4 * NG];
#pragma unroll
  for (int s = 0; s < NACC; ++s)
#pragma unroll
    for (int i = 0; i < 4 * NG; ++i) acc[s][i] = T(0);
  T* orow = out + (int64_t)r0 * a.wout + o0;
  unsigned parity = 0;

  // input row iy = iy0 + n with n = U * trip + PH
  auto step = [&](auto phase, int n) {
    constexpr int PH = decltype(phase)::value;
    __syncwarp();                                        // all lanes are done with the stage refilled now
    issue(std::integral_constant<int, (PH + U - 1) % U>{}, iy0 + n + U - 1);
    mbar_wait(mbar + PH * 8, parity);
    // x(i, kx): window column kx of this lane's output i
    T e[kInter ? 4 * (NL + 1) : NG * NV * 4];
    if (kInter) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned p = own + (unsigned)(PH * RW * 4 + i * 32 * SX * 4);
        if (SX % 2 == 0) {
#pragma unroll
          for (int v = 0; v < (NL + 1) / 2; ++v) {
            int2 raw;
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "r"(p + v * 8));
            e[i * (NL + 1) + 2 * v] = reinterpret_cast<const T&>(raw.x);
            e[i * (NL + 1) + 2 * v + 1] = reinterpret_cast<const T&>(raw.y);
          }
        } else {
#pragma unroll
          for (int v = 0; v < NL; ++v) e[i * (NL + 1) + v] = lds_1<T>(p + v * 4);
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          T t4[4];
          lds_v4(own + (unsigned)(PH * RW * 4 + g * 128 * SX * 4 + v * 16), t4);
#pragma unroll
          for (int k = 0; k < 4; ++k) e[(g * NV + v) * 4 + k] = t4[k];
        }
    }
#pragma unroll
    for (int ky = 0; ky < KY; ++ky) {
      if (((PH - ky) % SY + SY) % SY != 0) continue;     // window row ky of no output row
      constexpr int kBig = 64 * SY * NACC;               // keeps the compile-time division non-negative
      const int slot = ((PH - ky + kBig) / SY) % NACC;
#pragma unroll
      for (int i = 0; i < 4 * NG; ++i) {
        T v = acc[slot][i];
#pragma unroll
        for (int kx = 0; kx < KX; ++kx) {
          const T x = kInter ? e[i * (NL + 1) + SHIFT + kx] : e[(i / 4) * NV * 4 + SHIFT + (i % 4) * SX + kx];
          if (ky == 0 && kx == 0) v = (OP == kOpAffine) ? mad_wrap(a.coef[0], x, a.bias) : x;   // a new output row starts
          else v = fold<T, OP>(v, a.coef[ky * KX + kx], x);
        }
        acc[slot][i] = v;
      }
      if (ky == KY - 1 && n >= KY - 1) {                 // output row (n - ky) / SY of the band is complete
        T res[4 * NG];
#pragma unroll
        for (int i = 0; i < 4 * NG; ++i) {
          res[i] = acc[slot][i];
          if (OP == kOpAffine && std::is_same<T, float>::value && a.post_div != 0.f)
            res[i] = (T)((float)res[i] / a.post_div);
        }
        if (kInter) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (o0 + 32 * i < a.wout) __stcs(orow + 32 * i, res[i]);
        } else {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const int n_valid = min(4, a.wout - (o0 + 128 * g));
            T* o = orow + 128 * g;
            const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
            if (n_valid == 4 && (addr & 15) == 0) {
              int4 raw;
              raw.x = reinterpret_cast<const int&>(res[4 * g + 0]);
              raw.y = reinterpret_cast<const int&>(res[4 * g + 1]);
              raw.z = reinterpret_cast<const int&>(res[4 * g + 2]);
              raw.w = reinterpret_cast<const int&>(res[4 * g + 3]);
              __stcs(reinterpret_cast<int4*>(o), raw);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < n_valid) o[i] = res[4 * g + i];
            }
          }
        }
        orow += a.wout;
      }
    }
  };

  // prologue: rows iy0 .. iy0 + U - 2 into stages 0 .. U - 2
  p_static_for<0, U - 1>([&](auto k) { issue(k, iy0 + decltype(k)::value); });

  const int nrows = iy1 - iy0 + 1;
#pragma unroll 1
  for (int n = 0; n < nrows; n += U) {
    bool done = false;
    p_static_for<0, U>([&](auto ph) {
      constexpr int PHv = decltype(ph)::value;
      if (!done && n + PHv < nrows) step(ph, n + PHv);
      else done = true;
    });
    if (done) break;
    parity ^= 1u;
  }
}

// ---- dispatch -------------------------------------------------------------------------------
template <typename K>
int pool_band_rows(K kernel, int hout, int64_t tiles_per_band, int sy, int ky) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPT, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 4;
  }
  const int64_t resident = (int64_t)sm_count() * per_sm;
  const int lo = 8, hi = 96;
  if (tiles_per_band * ((hout + lo - 1) / lo) <= resident) return hout >= 64 ? 16 : lo;   // small problem
  int best = 32;
  double best_eff = 0.0;
  for (int th = lo; th <= hi; ++th) {
    if ((hout + th - 1) / th > 65535) continue;
    const int64_t tiles = tiles_per_band * ((hout + th - 1) / th);
    const int64_t waves = (tiles + resident - 1) / resident;
    const double halo = (double)(th * sy) / (double)(th * sy + (ky > sy ? ky - sy : 0));
    const double eff = (double)tiles / (double)(waves * resident) * halo;
    if (eff >= best_eff) {
      best_eff = eff;
      best = th;
    }
  }
  return best;
}

template <typename T, int OP, int KY, int KX, int SY, int SX>
int launch_pool_shift(PoolArgs<T>& a, dim3 grid, int shift, cudaStream_t s) {
#define KX_POOL_LAUNCH(SH)                                                                        \
  {                                                                                                \
    auto kern = pool2d_kernel<T, OP, KY, KX, SY, SX, SH>;                                          \
    a.th = pool_band_rows(kern, a.hout, (int64_t)grid.x * grid.z, SY, KY);                        \
    grid.y = (a.hout + a.th - 1) / a.th;                                                           \
    kern<<<grid, kPT, 0, s>>>(a);                                                                  \
  }
  switch (shift) {
    case 0: KX_POOL_LAUNCH(0) break;
    case 1: KX_POOL_LAUNCH(1) break;
    case 2: KX_POOL_LAUNCH(2) break;
    default: KX_POOL_LAUNCH(3) break;
  }
#undef KX_POOL_LAUNCH
  return after_launch("pool2d_kernel");
}

template <typename T, int OP>
int launch_pool_geom(PoolArgs<T>& a, dim3 grid, int ky, int kx, int sy, int sx, int shift, cudaStream_t s) {
#define KX_POOL_CASE(KY_, KX_, SY_, SX_) \
  if (ky == KY_ && kx == KX_ && sy == SY_ && sx == SX_) return launch_pool_shift<T, OP, KY_, KX_, SY_, SX_>(a, grid, shift, s);
  KX_POOL_CASE(2, 2, 2, 2)
  KX_POOL_CASE(3, 3, 2, 2)
  KX_POOL_CASE(3, 3, 3, 3)
  KX_POOL_CASE(4, 4, 4, 4)
  KX_POOL_CASE(4, 4, 2, 2)
  KX_POOL_CASE(1, 2, 1, 2)
  KX_POOL_CASE(1, 3, 1, 2)
  KX_POOL_CASE(1, 3, 1, 3)
  KX_POOL_CASE(1, 4, 1, 4)
  KX_POOL_CASE(5, 5, 2, 2)
  KX_POOL_CASE(1, 5, 1, 2)
  if (OP != kOpAffine) {  // stride-1 AFFINE windows are kx_stencil2d.cu's job
    KX_POOL_CASE(3, 3, 1, 1)
    KX_POOL_CASE(2, 2, 1, 1)
    KX_POOL_CASE(1, 3, 1, 1)
    KX_POOL_CASE(1, 2, 1, 1)
  }
#undef KX_POOL_CASE
  return KX_ERR_UNSUPPORTED;
}

template <typename T>
int run_pool(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  const KxFunc& f = p.func[0];
  const int ay = nd >= 2 ? nd - 2 : -1, ax = nd - 1;
  const int ky = ay >= 0 ? g.ksize[ay] : 1, kx = g.ksize[ax];
  const int sy = ay >= 0 ? g.stride[ay] : 1, sx = g.stride[ax];
  int64_t lead = 1;
  for (int d = 0; d < nd - 2; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    lead *= g.in_shape[d];
  }
  const int64_t batch = m.batch * lead;
  if (batch > 65535 || ky * kx > 32) return KX_ERR_UNSUPPORTED;
  if (f.tap_end - f.tap_begin != ky * kx) return KX_ERR_UNSUPPORTED;          // dense windows only
  if (m.coef_dev) return KX_ERR_UNSUPPORTED;
  PoolArgs<T> a;
  memset(&a, 0, sizeof(a));
  uint32_t seen = 0;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int slot = (ay >= 0 ? p.tap_off[t][ay] : 0) * kx + p.tap_off[t][ax];
    if (seen & (1u << slot)) return KX_ERR_UNSUPPORTED;
    seen |= 1u << slot;
    a.coef[slot] = (T)p.coef[t];
  }
  a.hin = ay >= 0 ? (int)g.in_shape[ay] : 1;
  a.win = (int)g.in_shape[ax];
  a.hout = ay >= 0 ? (int)g.out_shape[ay] : 1;
  a.wout = (int)g.out_shape[ax];
  a.ly = ay >= 0 ? g.lo[ay] : 0;
  a.lx = g.lo[ax];
  if (g.in_shape[ax] > 0x3fffffff || (ay >= 0 && g.in_shape[ay] > 0x3fffffff)) return KX_ERR_UNSUPPORTED;
  // TMA bulk copies need 16-byte aligned rows
  if (a.win % 4 != 0 || (reinterpret_cast<uintptr_t>(m.in) & 15) != 0) return KX_ERR_UNSUPPORTED;
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.in_batch = (int64_t)a.hin * a.win;
  a.out_batch = (int64_t)a.hout * a.wout;
  a.bias = (T)f.bias;
  a.post_div = (float)f.post_div;
  const int shift = (((-a.lx) % 4) + 4) % 4;
  const int block_out = (kPT / 32) * (sx == 1 ? 256 : 128);   // WarpOut<SX>::value per warp
  dim3 grid((a.wout + block_out - 1) / block_out, 1, (unsigned)batch);
  if (f.kind == KX_AFFINE) return launch_pool_geom<T, kOpAffine>(a, grid, ky, kx, sy, sx, shift, s);
  if (f.kind == KX_REDUCE_MAX) return launch_pool_geom<T, kOpMax>(a, grid, ky, kx, sy, sx, shift, s);
  if (f.kind == KX_REDUCE_MIN) return launch_pool_geom<T, kOpMin>(a, grid, ky, kx, sy, sx, shift, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace

int try_pool2d(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (p.n_funcs != 1 || m.fn_elems != 1 || p.func[0].kind == KX_GATHER) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype) return KX_ERR_UNSUPPORTED;
  if (p.func[0].kind == KX_AFFINE && p.func[0].post_div != 0.0 && g.in_dtype != KX_F32) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype == KX_F32) return run_pool<float>(m, p, s);
  if (g.in_dtype == KX_I32 && (p.func[0].kind != KX_AFFINE || p.func[0].coef_is_int)) return run_pool<int32_t>(m, p, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
