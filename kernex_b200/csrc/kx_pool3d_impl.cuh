// Fast path: 3-D POOLING -- cubic windows K x K x K with stride S x S x S (S >= 2) over the last three axes:
// average / sum pooling (README.md:327-337 generalised to volumes), max / min pooling, strided weighted windows.
// Class (b) of the kernel templates.  Before this kernel these programs ran on generic_map_kernel at 0.06 of the HBM
// roofline (profiles/r1c_misc_paths.jsonl).
//
// Same building blocks as kx_pool2d.cu (per-warp shared-memory ring filled by TMA bulk copies, one mbarrier per
// stage, __syncwarp as the only synchronisation, interleaved output columns for stride > 1), with the march turned
// by 90 degrees:
//   * a WARP owns ONE output row (y', 128 output columns) and marches along Z.  One ring stage = the K input rows
//     y'*S - lo_y .. + K-1 of one input plane (two ~512-byte bulk copies per row), fetched U-1 planes ahead;
//   * the kernel is input-plane stationary: a plane's K rows are read from shared memory once and folded into the
//     ceil(K/S) output planes they belong to (output plane z' takes window plane kz = z + lo_z - z'*S); the plane
//     loop is unrolled by the period of that pattern, so ring stages and accumulator slots are compile-time;
//   * accumulation order of one output = bias, then the window in row-major (kz, ky, kx) order -- the chain runs
//     straight through the accumulator, no per-plane partial sums;
//   * rows / planes outside the array are zero-filled stages (the reference's zero padding, utils.py:43-53);
//   * neighbouring output rows share K - S input rows: the warps of consecutive rows are dispatched together and
//     march in step, so the shared rows are L2 hits (DRAM traffic stays at 4 B per input element);
//   * long volumes are cut into chunks of output planes when the grid would not fill the GPU.
#pragma once

#include <type_traits>

#include "kx_internal.cuh"
#include "kx_tma.cuh"

namespace kx {
namespace p3 {

constexpr int kPT3 = 64;   // threads per block: 2 warps, one output row each

enum Pool3Op { kOpAffine = 0, kOpMax = 1, kOpMin = 2 };

template <typename T>
struct Pool3Args {
  const T* in;
  T* out;
  int64_t in_batch, out_batch;
  int din, hin, win, dout, hout, wout;
  int lz, ly, lx;
  int zc, nzc;            // output planes per chunk, chunks per volume
  float post_div;
  T bias;
  T coef[27];
};

template <int S>
struct RowW3 {
  static constexpr int value = (((128 - 1) * S + 3 + 3) + 3) / 4 * 4 + 4;   // span of 128 outputs + shift, rounded, + slack vector
};

template <typename T, int OP>
__device__ __forceinline__ T fold3(T acc, T c, T v) {
  if (OP == kOpAffine) return mad_wrap(c, v, acc);
  if (OP == kOpMax) return red_max(acc, v);
  return red_min(acc, v);
}

template <typename T, int OP, int K, int S, int SHIFT>
__global__ void __launch_bounds__(kPT3)
pool3d_kernel(const __grid_constant__ Pool3Args<T> a) {
  static_assert(S >= 2, "stride-1 windows are stencils");
  constexpr int NACC = (K + S - 1) / S;                   // output planes an input plane can belong to
  constexpr int PERIOD = S * NACC;
  constexpr int U = (PERIOD == 3) ? 3 : 4;                // unroll = ring depth (planes), a multiple of PERIOD
  static_assert(U % PERIOD == 0, "unroll must cover the accumulator rotation");
  constexpr int RW = RowW3<S>::value;                     // floats per ring row
  constexpr int NL = SHIFT + K;                           // floats a lane needs per output
  constexpr int STAGE = K * RW;                           // floats per stage (K rows of one plane)
  __shared__ __align__(128) T smem[(kPT3 / 32) * U * STAGE];
  __shared__ __align__(8) uint64_t bars[(kPT3 / 32) * U];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wo0 = (blockIdx.x * (kPT3 / 32) + warp) * 128;         // first output column of the warp
  const int o0 = wo0 + lane;                                        // lane's outputs: o0, o0+32, o0+64, o0+96
  const int yo = blockIdx.y;                                        // output row
  const int b = blockIdx.z / a.nzc, chunk = blockIdx.z - b * a.nzc;
  const int z0 = chunk * a.zc, z1 = min(z0 + a.zc, a.dout);         // output planes [z0, z1)
  const T* __restrict__ in = a.in + (int64_t)b * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)b * a.out_batch;
  const int iz0 = z0 * S - a.lz, iz1 = (z1 - 1) * S - a.lz + K - 1;  // first / last input plane of the chunk
  const int iy0 = yo * S - a.ly;                                     // first input row of the window rows

  const int c_start = wo0 * S - a.lx - SHIFT;                        // ring column 0 <-> this input column (multiple of 4)
  const int c_lo = max(c_start, 0), c_hi = min(c_start + RW, a.win);
  const unsigned bytes = (wo0 < a.wout && c_hi > c_lo) ? (unsigned)((c_hi - c_lo) * 4) : 0u;
  const unsigned ring0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)(warp * U * STAGE * 4);
  const unsigned dst_off = (unsigned)((c_lo - c_start) * 4);
  const unsigned own = ring0 + (unsigned)(lane * S * 4);
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(bars) + (unsigned)(warp * U * 8);
  const T* src = in + c_lo;
  const unsigned half = bytes >= 512 ? ((bytes / 2 + 15) & ~15u) : bytes;
  const int64_t plane_elems = (int64_t)a.hin * a.win;
  // which of the K window rows exist (the others are zero padding): fixed for the whole march
  unsigned row_ok = 0;
#pragma unroll
  for (int r = 0; r < K; ++r)
    if (iy0 + r >= 0 && iy0 + r < a.hin) row_ok |= 1u << r;
  const int n_ok = __popc(row_ok);

  for (int i = lane; i < U * STAGE / 4; i += 32) sts_zero16(ring0 + (unsigned)(i * 16));
  if (lane < U) mbar_init(mbar + lane * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncwarp();

  // stage S_ <- input plane iz: its K window rows (padding rows of a real plane stay zero from the clear below)
  auto issue = [&](auto stage, int iz) {
    constexpr int S_ = decltype(stage)::value;
    const bool plane_ok = (iz >= 0) && (iz < a.din) && (iz <= iz1) && bytes != 0 && n_ok > 0;
    const unsigned st = ring0 + (unsigned)(S_ * STAGE * 4);
    if (!plane_ok || n_ok < K) {   // some (or all) rows are padding: clear them before the copies of the others land
#pragma unroll
      for (int r = 0; r < K; ++r)
        if (!plane_ok || !(row_ok & (1u << r)))
          for (int i = lane; i < RW / 4; i += 32) sts_zero16(st + (unsigned)(r * RW * 4 + i * 16));
      fence_proxy_async();
      __syncwarp();
    }
    if (lane == 0) {
      if (plane_ok) {
        mbar_expect_tx(mbar + S_ * 8, bytes * (unsigned)n_ok);
        const T* pp = src + (int64_t)iz * plane_elems + (int64_t)iy0 * a.win;
#pragma unroll
        for (int r = 0; r < K; ++r) {
          if (row_ok & (1u << r)) {
            const T* rp = pp + (int64_t)r * a.win;
            const unsigned dst = st + (unsigned)(r * RW * 4) + dst_off;
            bulk_g2s(dst, rp, half, mbar + S_ * 8);
            if (bytes > half) bulk_g2s(dst + half, rp + half / 4, bytes - half, mbar + S_ * 8);
          }
        }
      } else {
        mbar_arrive(mbar + S_ * 8);
      }
    }
  };

  T acc[NACC][4];
#pragma unroll
  for (int s = 0; s < NACC; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[s][i] = T(0);
  T* orow = out + ((int64_t)z0 * a.hout + yo) * a.wout + o0;
  const int64_t out_plane = (int64_t)a.hout * a.wout;
  unsigned parity = 0;

  // input plane iz = iz0 + n with n = U * trip + PH
  auto step = [&](auto phase, int n) {
    constexpr int PH = decltype(phase)::value;
    __syncwarp();                                        // all lanes are done with the stage refilled now
    issue(std::integral_constant<int, (PH + U - 1) % U>{}, iz0 + n + U - 1);
    mbar_wait(mbar + PH * 8, parity);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      // window columns of this lane's 4 outputs in row ky of the plane
      T e[4][NL + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned p = own + (unsigned)((PH * STAGE + ky * RW + i * 32 * S) * 4);
        if (S % 2 == 0) {
#pragma unroll
          for (int v = 0; v < (NL + 1) / 2; ++v) {
            int2 raw;
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "r"(p + v * 8));
            e[i][2 * v] = reinterpret_cast<const T&>(raw.x);
            e[i][2 * v + 1] = reinterpret_cast<const T&>(raw.y);
          }
        } else {
#pragma unroll
          for (int v = 0; v < NL; ++v) e[i][v] = lds_1<T>(p + v * 4);
        }
      }
#pragma unroll
      for (int kz = 0; kz < K; ++kz) {
        if (((PH - kz) % S + S) % S != 0) continue;      // this plane is window plane kz of no output plane
        constexpr int kBig = 64 * S * NACC;              // keeps the compile-time division non-negative
        const int slot = ((PH - kz + kBig) / S) % NACC;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          T v = acc[slot][i];
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const T x = e[i][SHIFT + kx];
            if (kz == 0 && ky == 0 && kx == 0) v = (OP == kOpAffine) ? mad_wrap(a.coef[0], x, a.bias) : x;  // a new output starts
            else v = fold3<T, OP>(v, a.coef[(kz * K + ky) * K + kx], x);
          }
          acc[slot][i] = v;
        }
      }
    }
    // the output plane whose LAST window plane this was is complete
    if (((PH - (K - 1)) % S + S) % S == 0 && n >= K - 1) {
      constexpr int kBig = 64 * S * NACC;
      const int slot = ((PH - (K - 1) + kBig) / S) % NACC;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        T res = acc[slot][i];
        if (OP == kOpAffine && std::is_same<T, float>::value && a.post_div != 0.f) res = (T)((float)res / a.post_div);
        if (o0 + 32 * i < a.wout) __stcs(orow + 32 * i, res);
      }
      orow += out_plane;
    }
  };

#pragma unroll
  for (int k = 0; k < U - 1; ++k) {
    // prologue: planes iz0 .. iz0 + U - 2 into stages 0 .. U - 2
    if (k == 0) issue(std::integral_constant<int, 0>{}, iz0);
    if (k == 1) issue(std::integral_constant<int, 1 % U>{}, iz0 + 1);
    if (k == 2) issue(std::integral_constant<int, 2 % U>{}, iz0 + 2);
  }

  const int nplanes = iz1 - iz0 + 1;
#pragma unroll 1
  for (int n = 0; n < nplanes; n += U) {
    step(std::integral_constant<int, 0>{}, n);
    if (n + 1 >= nplanes) break;
    step(std::integral_constant<int, 1>{}, n + 1);
    if (n + 2 >= nplanes) break;
    step(std::integral_constant<int, 2>{}, n + 2);
    if (U == 4) {
      if (n + 3 >= nplanes) break;
      step(std::integral_constant<int, 3 % U>{}, n + 3);
    }
    parity ^= 1u;
  }
}

template <typename T, int OP, int K, int S>
int launch_pool3(Pool3Args<T>& a, dim3 grid, int shift, cudaStream_t s) {
  switch (shift) {
    case 0: pool3d_kernel<T, OP, K, S, 0><<<grid, kPT3, 0, s>>>(a); break;
    case 1: pool3d_kernel<T, OP, K, S, 1><<<grid, kPT3, 0, s>>>(a); break;
    case 2: pool3d_kernel<T, OP, K, S, 2><<<grid, kPT3, 0, s>>>(a); break;
    default: pool3d_kernel<T, OP, K, S, 3><<<grid, kPT3, 0, s>>>(a); break;
  }
  return after_launch("pool3d_kernel");
}

// one translation unit per (K, S): kx_pool3d_k2s2.cu, kx_pool3d_k3s2.cu, kx_pool3d_k3s3.cu
template <int K, int S>
int run_pool3_ks(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s);

template <typename T, int K, int S>
int run_pool3_t(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  const KxFunc& f = p.func[0];
  const int az = nd - 3, ay = nd - 2, ax = nd - 1;
  Pool3Args<T> a;
  memset(&a, 0, sizeof(a));
  uint32_t seen = 0;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int slot = (p.tap_off[t][az] * K + p.tap_off[t][ay]) * K + p.tap_off[t][ax];
    if (seen & (1u << slot)) return KX_ERR_UNSUPPORTED;
    seen |= 1u << slot;
    a.coef[slot] = (T)p.coef[t];
  }
  a.din = (int)g.in_shape[az];
  a.hin = (int)g.in_shape[ay];
  a.win = (int)g.in_shape[ax];
  a.dout = (int)g.out_shape[az];
  a.hout = (int)g.out_shape[ay];
  a.wout = (int)g.out_shape[ax];
  a.lz = g.lo[az];
  a.ly = g.lo[ay];
  a.lx = g.lo[ax];
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.in_batch = (int64_t)a.din * a.hin * a.win;
  a.out_batch = (int64_t)a.dout * a.hout * a.wout;
  a.bias = (T)f.bias;
  a.post_div = (float)f.post_div;
  // z chunks: enough blocks for ~3 waves of 8 resident blocks per SM (measured at 1024^3, 3x3x3 s2 max pooling:
  // 1 chunk 0.89 of the HBM roofline, 2 chunks 0.98, 4 chunks 0.996; each chunk re-reads K - S planes)
  const int64_t rows = (int64_t)((a.wout + 255) / 256) * a.hout * batch;     // blocks per chunk
  int nzc = 1;
  const int64_t want = (int64_t)sm_count() * 24;
  if (rows < want) nzc = (int)((want + rows - 1) / rows);
  if (nzc > a.dout) nzc = a.dout;
  if (const char* e = getenv("KX_POOL3_NZC")) nzc = atoi(e) > 0 ? (atoi(e) < a.dout ? atoi(e) : a.dout) : nzc;   // tuning knob
  a.zc = (a.dout + nzc - 1) / nzc;
  a.nzc = (a.dout + a.zc - 1) / a.zc;
  if ((int64_t)a.nzc * batch > 65535 || a.hout > 65535) return KX_ERR_UNSUPPORTED;
  dim3 grid((a.wout + 255) / 256, a.hout, (unsigned)(a.nzc * batch));
  const int shift = (((-a.lx) % 4) + 4) % 4;
  if (f.kind == KX_AFFINE) return launch_pool3<T, kOpAffine, K, S>(a, grid, shift, s);
  if (f.kind == KX_REDUCE_MAX) return launch_pool3<T, kOpMax, K, S>(a, grid, shift, s);
  if (f.kind == KX_REDUCE_MIN) return launch_pool3<T, kOpMin, K, S>(a, grid, shift, s);
  return KX_ERR_UNSUPPORTED;
}

#define KX_POOL3_INSTANTIATE(K, S)                                                                         \
  template <>                                                                                               \
  int run_pool3_ks<K, S>(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {             \
    const KxGeom& g = m.g;                                                                                  \
    if (g.in_dtype == KX_F32) return run_pool3_t<float, K, S>(m, p, batch, s);                              \
    if (g.in_dtype == KX_I32 && (p.func[0].kind != KX_AFFINE || p.func[0].coef_is_int))                     \
      return run_pool3_t<int32_t, K, S>(m, p, batch, s);                                                    \
    return KX_ERR_UNSUPPORTED;                                                                              \
  }

}  // namespace p3
}  // namespace kx
