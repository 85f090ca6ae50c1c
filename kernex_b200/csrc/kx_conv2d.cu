// Fast paths for conv-as-kmap (README.md:96-111, tests/test_interface.py:315-348, BASELINE
// config 3a): a (C, KY, KX) window over a channel-first (C, H, W) image with padding
// ('valid', *, *), i.e. out[0, y, x] = sum_c sum_ky sum_kx w[c,ky,kx] * in[c, y+ky-ly, x+kx-lx],
// and the fused weight-gradient reduction of the same geometry.
//
// Both kernels use the row-marching skeleton of kx_stencil2d.cu with one register queue per
// channel: a thread owns 4 output columns, walks down a band of rows and keeps the last KY
// rows of each of the C input planes in registers, so every input row is fetched once per
// thread column (16 B per lattice update for C = 3: three planes read, one written) and the
// 27 taps are pure register FMAs.  Marching along axis 0 (kx_stencil3d.cu) has nothing to
// reuse here because the output is a single plane.
//
//   conv_fwd_kernel    : forward; weights from kernel parameters or a device table
//   conv_wgrad_kernel  : dW[c,ky,kx] = sum_{y,x} g[y,x] * in[c, y+ky-ly, x+kx-lx]; each thread
//                        accumulates C*KY*KX partial sums in registers, blocks reduce through
//                        shuffles + shared memory into partial[block][tap]; a second kernel sums
//                        the partials in a fixed order (deterministic, no atomics).
#include <algorithm>
#include <type_traits>

#include "kx_internal.cuh"
#include "kx_tma.cuh"

namespace kx {

namespace {

constexpr int kTX = 128;
constexpr int kCols = 4;
constexpr int kMaxC = 4;

template <typename T>
struct ConvArgs {
  const T* in;
  const T* g;                  // wgrad: cotangent [hout, wout]
  T* out;                      // fwd: [hout, wout]; wgrad: partial[blocks][C*KY*KX]
  const T* coef_dev;
  int64_t in_batch, out_batch, plane;
  int hin, win, hout, wout;
  int ly, lx, th;
  int coef_rev;                // device table is in reversed tap order (transposed program)
  T bias;
  T coef[kMaxC * 9];
};

template <typename T>
__device__ __forceinline__ void ldv4(const T* p, T (&d)[4]) {
  const int4 raw = __ldg(reinterpret_cast<const int4*>(p));
  d[0] = reinterpret_cast<const T&>(raw.x);
  d[1] = reinterpret_cast<const T&>(raw.y);
  d[2] = reinterpret_cast<const T&>(raw.z);
  d[3] = reinterpret_cast<const T&>(raw.w);
}

template <typename T, int KX, int SHIFT, bool VEC, int NE>
__device__ __forceinline__ void load_row(const T* __restrict__ plane, int iy, int xa, int hin, int win, T (&dst)[NE]) {
  const bool row_ok = (iy >= 0) && (iy < hin);
  const T* rp = plane + (int64_t)iy * win;
  if (VEC) {
#pragma unroll
    for (int v = 0; v < NE / 4; ++v) {
      const int c = xa + 4 * v;
      T t4[4] = {T(0), T(0), T(0), T(0)};
      if (row_ok && c >= 0 && c + 3 < win) ldv4(rp + c, t4);
#pragma unroll
      for (int e = 0; e < 4; ++e) dst[4 * v + e] = t4[e];
    }
  } else {
#pragma unroll
    for (int e = 0; e < KX + kCols - 1; ++e) {
      const int c = xa + e;
      dst[e] = (row_ok && c >= 0 && c < win) ? __ldg(rp + c) : T(0);
    }
  }
}

template <typename T, int C, int KY, int KX, int SHIFT, bool VEC>
__global__ void __launch_bounds__(kTX)
conv_fwd_kernel(const __grid_constant__ ConvArgs<T> a) {
  constexpr int NV = (SHIFT + KX + kCols - 1 + 3) / 4;
  constexpr int NE = NV * 4;
  constexpr int NT = C * KY * KX;
  const int j0 = (blockIdx.x * kTX + threadIdx.x) * kCols;
  if (j0 >= a.wout) return;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const T* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int xa = j0 - a.lx - SHIFT;
  const int n_valid = min(kCols, a.wout - j0);

  T w[NT];
  if (a.coef_dev) {
#pragma unroll
    for (int i = 0; i < NT; ++i) w[i] = a.coef_dev[a.coef_rev ? NT - 1 - i : i];
  } else {
#pragma unroll
    for (int i = 0; i < NT; ++i) w[i] = a.coef[i];
  }

  T q[C][KY][NE];
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky)
      load_row<T, KX, SHIFT, VEC, NE>(in + c * a.plane, r0 - a.ly + ky, xa, a.hin, a.win, q[c][ky]);
  }
  for (int r = r0; r < r1; ++r) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      load_row<T, KX, SHIFT, VEC, NE>(in + c * a.plane, r - a.ly + KY - 1, xa, a.hin, a.win, q[c][KY - 1]);
    T acc[kCols];
#pragma unroll
    for (int i = 0; i < kCols; ++i) acc[i] = a.bias;
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int ky = 0; ky < KY; ++ky) {
#pragma unroll
        for (int kx = 0; kx < KX; ++kx) {
#pragma unroll
          for (int i = 0; i < kCols; ++i)
            acc[i] = mad_wrap(w[(c * KY + ky) * KX + kx], q[c][ky][SHIFT + kx + i], acc[i]);
        }
      }
    }
    T* o = out + (int64_t)r * a.wout + j0;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
    if (n_valid == 4 && (addr & 15) == 0) {
      int4 raw;
      raw.x = reinterpret_cast<const int&>(acc[0]);
      raw.y = reinterpret_cast<const int&>(acc[1]);
      raw.z = reinterpret_cast<const int&>(acc[2]);
      raw.w = reinterpret_cast<const int&>(acc[3]);
      __stcs(reinterpret_cast<int4*>(o), raw);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < n_valid) o[i] = acc[i];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
        for (int e = 0; e < NE; ++e) q[c][ky][e] = q[c][ky + 1][e];
      }
    }
  }
}

template <int C, int KY, int KX, int SHIFT, bool VEC>
__global__ void __launch_bounds__(kTX)
conv_wgrad_kernel(const __grid_constant__ ConvArgs<float> a) {
  constexpr int NV = (SHIFT + KX + kCols - 1 + 3) / 4;
  constexpr int NE = NV * 4;
  constexpr int NT = C * KY * KX;
  __shared__ float red[kTX / 32][NT];
  const int j0 = (blockIdx.x * kTX + threadIdx.x) * kCols;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const float* __restrict__ in = a.in;
  const int xa = j0 - a.lx - SHIFT;
  float acc[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t] = 0.f;

  if (j0 < a.wout) {
    float q[C][KY][NE];
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int ky = 0; ky < KY - 1; ++ky)
        load_row<float, KX, SHIFT, VEC, NE>(in + c * a.plane, r0 - a.ly + ky, xa, a.hin, a.win, q[c][ky]);
    }
    for (int r = r0; r < r1; ++r) {
#pragma unroll
      for (int c = 0; c < C; ++c)
        load_row<float, KX, SHIFT, VEC, NE>(in + c * a.plane, r - a.ly + KY - 1, xa, a.hin, a.win, q[c][KY - 1]);
      float gv[kCols];
      const float* gp = a.g + (int64_t)r * a.wout + j0;
      if (j0 + 3 < a.wout && ((reinterpret_cast<uintptr_t>(gp) & 15) == 0)) {
        ldv4(gp, gv);
      } else {
#pragma unroll
        for (int i = 0; i < kCols; ++i) gv[i] = (j0 + i < a.wout) ? __ldg(gp + i) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int ky = 0; ky < KY; ++ky) {
#pragma unroll
          for (int kx = 0; kx < KX; ++kx) {
#pragma unroll
            for (int i = 0; i < kCols; ++i)
              acc[(c * KY + ky) * KX + kx] = fmaf(gv[i], q[c][ky][SHIFT + kx + i], acc[(c * KY + ky) * KX + kx]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
          for (int e = 0; e < NE; ++e) q[c][ky][e] = q[c][ky + 1][e];
        }
      }
    }
  }
  // block reduction: shuffles inside a warp, shared memory across the 4 warps
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    float v = acc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < NT) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < kTX / 32; ++wv) v += red[wv][threadIdx.x];
    const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    a.out[blk * NT + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Aligned-row variants (rows 16-byte aligned, the usual case).  The register-queue kernels above
// fetch 3 vectors per channel row for 'same' padding, keep C*KY rows live (96-115 registers) and
// consume every load right after issuing it: ncu shows them latency-bound at ~4.7 TB/s with 20
// warps per SM.  Here
//   * every WARP owns a private shared-memory ring of kRing input rows (C channels [+ the
//     cotangent row for the weight gradient] x (128 + 2*4) columns).  Lane l copies its own 16
//     bytes of each channel row with cp.async.cg (LDGSTS, zero-fill = the reference's zero
//     padding), lanes 0 / 31 also copy the left / right halo vector; rows are requested
//     kRing-1 = 2 iterations before they are used, so the in-flight bytes cost no registers and the
//     only synchronisation is __syncwarp (no block barrier, warps drift freely);
//   * a thread reads back only its own vector (one conflict-free LDS.128 per channel row); the
//     x halo comes from the neighbouring lanes by shuffle, the warp-edge lanes read the halo
//     vector copied above;
//   * the kernels are input-row stationary: the forward pass keeps 3 rows of 4 output
//     accumulators (input row iy feeds output rows iy+ly-ky), the weight gradient keeps the last 3
//     cotangent rows, so there is no input row queue at all;
//   * the band height is chosen so that the grid is close to a whole number of resident waves.
// Accumulation order of one output: bias, then input rows as they stream in (ky = 2, 1, 0),
// channels, kx; integer results are order-independent.
constexpr int kRing = 3;         // the row loop is unrolled by kRing: stages and queues rotate by renaming
constexpr int kRowW = 128 + 8;   // floats per channel row in the ring: [4 left halo | 128 own | 4 right halo]

// Per-warp streamer of the C input planes.  One elected lane issues one bulk copy per channel row:
// the warp's 128 columns plus the 4-column halo vectors on either side, clipped to the row (the
// clipped-away part of the ring is zeroed once and never overwritten = zero padding in x).  Rows
// outside the image are zero-filled by the lanes themselves.  Completion is tracked by one
// mbarrier per stage (expect_tx = C * bytes).
template <typename T, int C, int LX>
struct RowRing {
  unsigned own;        // shared address of this lane's vector in channel row 0 of stage 0
  unsigned row0;       // shared address of the warp's ring
  unsigned mbar;       // shared address of the warp's kRing mbarriers
  unsigned dst_off;    // byte offset inside a channel row where the copy lands
  unsigned bytes;      // bytes per channel-row copy (0 = the warp lies outside the image)
  const T* src;        // first copied column in row 0 of channel 0
  int64_t plane;
  int pitch, nrow, last;   // rows outside [0, nrow) or after `last` are zero-filled, never read
  bool lane_first, lane_last;

  __device__ __forceinline__ void init(T* smem, uint64_t* bars, int warp, int lane, const T* in, int64_t plane_,
                                       int wj0, int win, int hin, int last_) {
    row0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)(warp * kRing * C * kRowW * 4);
    own = row0 + (unsigned)((4 + 4 * lane) * 4);
    mbar = (unsigned)__cvta_generic_to_shared(bars) + (unsigned)(warp * kRing * 8);
    const int c_lo = max(wj0 - 4, 0), c_hi = min(wj0 + 132, win);
    bytes = c_hi > c_lo ? (unsigned)((c_hi - c_lo) * 4) : 0u;
    dst_off = (unsigned)((c_lo - (wj0 - 4)) * 4);
    src = in + c_lo;
    plane = plane_;
    pitch = win;
    nrow = hin;
    last = last_;
    lane_first = lane == 0;
    lane_last = lane == 31;
    // zero the whole ring once (x padding + rows never copied), initialise the barriers
    for (int i = lane; i < kRing * C * kRowW / 4; i += 32) sts_zero16(row0 + (unsigned)(i * 16));
    if (lane < kRing) mbar_init(mbar + lane * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    __syncwarp();
  }

  __device__ __forceinline__ bool row_ok(int iy) const { return (iy >= 0) && (iy < nrow) && (iy <= last); }

  // request row iy into STAGE (`extra` more bytes will be added to the same barrier by the caller)
  template <int STAGE>
  __device__ __forceinline__ void issue(int iy, unsigned extra = 0) const {
    const bool ok = row_ok(iy) && bytes != 0;
    if (lane_first) {
      const unsigned bar = mbar + STAGE * 8;
      if (ok) {
        mbar_expect_tx(bar, bytes * C + extra);
        const T* rp = src + (int64_t)iy * pitch;
#pragma unroll
        for (int ch = 0; ch < C; ++ch)
          bulk_g2s(row0 + (unsigned)((STAGE * C + ch) * kRowW * 4) + dst_off, rp + ch * plane, bytes, bar);
      } else if (extra) {
        mbar_expect_tx(bar, extra);
      } else {
        mbar_arrive(bar);
      }
    }
    if (!ok) {  // a padding row: every lane clears what it will read back (own vector, edge lanes the halo)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        const unsigned off = (unsigned)((STAGE * C + ch) * kRowW * 4);
        sts_zero16(own + off);
        if (lane_first) sts_zero16(own + off - 16);
        if (lane_last) sts_zero16(own + off + 16);
      }
      fence_proxy_async();
    }
  }

  template <int STAGE>
  __device__ __forceinline__ void wait(unsigned parity) const {
    mbar_wait(mbar + STAGE * 8, parity);
  }

  // e[0..5] = columns j0-LX .. j0-LX+5 of channel ch in STAGE
  template <int STAGE>
  __device__ __forceinline__ void window(int ch, T (&e)[6]) const {
    const unsigned off = (unsigned)((STAGE * C + ch) * kRowW * 4);
    T v[4];
    lds_v4(own + off, v);
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      const T up = __shfl_up_sync(0xffffffffu, v[4 - LX + k], 1);
      T edge = T(0);
      if (lane_first) edge = lds_1<T>(own + off - (LX - k) * 4);
      e[k] = lane_first ? edge : up;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) e[LX + k] = v[k];
#pragma unroll
    for (int k = 0; k < 2 - LX; ++k) {
      const T dn = __shfl_down_sync(0xffffffffu, v[k], 1);
      T edge = T(0);
      if (lane_last) edge = lds_1<T>(own + off + 16 + k * 4);
      e[LX + 4 + k] = lane_last ? edge : dn;
    }
  }
};

template <typename T>
__device__ __forceinline__ void st_out4(T* o, const T (&v)[4], int n_valid) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
  if (n_valid == 4 && (addr & 15) == 0) {
    int4 raw;
    raw.x = reinterpret_cast<const int&>(v[0]);
    raw.y = reinterpret_cast<const int&>(v[1]);
    raw.z = reinterpret_cast<const int&>(v[2]);
    raw.w = reinterpret_cast<const int&>(v[3]);
    __stcs(reinterpret_cast<int4*>(o), raw);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n_valid) o[i] = v[i];
  }
}

template <typename T, int C, int LX>
__global__ void __launch_bounds__(kTX)
conv_fwd_rows_kernel(const __grid_constant__ ConvArgs<T> a) {
  constexpr int NT = C * 9;
  __shared__ __align__(128) T smem[(kTX / 32) * kRing * C * kRowW];
  __shared__ __align__(8) uint64_t bars[(kTX / 32) * kRing];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = (blockIdx.x * kTX + threadIdx.x) * kCols;   // whole warps stay alive for the shuffles
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const T* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int n_valid = min(kCols, a.wout - j0);
  const int iy0 = r0 - a.ly, iy1 = r1 - 1 - a.ly + 2;  // first / last input row of the band

  RowRing<T, C, LX> ring;
  ring.init(smem, bars, warp, lane, in, a.plane, j0 - 4 * lane, a.win, a.hin, iy1);

  T w[NT];
  if (a.coef_dev) {
#pragma unroll
    for (int i = 0; i < NT; ++i) w[i] = a.coef_dev[a.coef_rev ? NT - 1 - i : i];
  } else {
#pragma unroll
    for (int i = 0; i < NT; ++i) w[i] = a.coef[i];
  }

  ring.template issue<0>(iy0);
  ring.template issue<1>(iy0 + 1);

  // acc[(P + s) % 3] is output row iy + ly - 2 + s while input row iy is processed in phase P
  T acc[3][kCols];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int i = 0; i < kCols; ++i) acc[s][i] = a.bias;
  T* orow = out + (int64_t)(iy0 + a.ly - 2) * a.wout + j0;

  unsigned parity = 0;
  auto step = [&](auto phase, int iy) {
    constexpr int P = decltype(phase)::value;
    __syncwarp();  // every lane is done reading the stage that is refilled now
    ring.template issue<(P + 2) % 3>(iy + 2);
    ring.template wait<P>(parity);  // row iy has landed
#pragma unroll
    for (int c = 0; c < C; ++c) {
      T e[6];
      ring.template window<P>(c, e);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if constexpr (std::is_same<T, float>::value) {  // two packed FMAs instead of four scalar ones
            const float wk = w[(c * 3 + ky) * 3 + kx];
            ffma2(acc[(P + 2 - ky) % 3][0], acc[(P + 2 - ky) % 3][1], wk, wk, e[kx], e[kx + 1]);
            ffma2(acc[(P + 2 - ky) % 3][2], acc[(P + 2 - ky) % 3][3], wk, wk, e[kx + 2], e[kx + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < kCols; ++i)
              acc[(P + 2 - ky) % 3][i] = mad_wrap(w[(c * 3 + ky) * 3 + kx], e[kx + i], acc[(P + 2 - ky) % 3][i]);
          }
        }
    }
    if (iy >= iy0 + 2 && n_valid > 0) st_out4(orow, acc[P % 3], n_valid);  // output row iy + ly - 2 is complete
    orow += a.wout;
#pragma unroll
    for (int i = 0; i < kCols; ++i) acc[P % 3][i] = a.bias;
  };

#pragma unroll 1
  for (int iy = iy0; iy <= iy1; iy += 3) {
    step(std::integral_constant<int, 0>{}, iy);
    if (iy + 1 > iy1) break;
    step(std::integral_constant<int, 1>{}, iy + 1);
    if (iy + 2 > iy1) break;
    step(std::integral_constant<int, 2>{}, iy + 2);
    parity ^= 1u;
  }
}

template <int C, int LX>
__global__ void __launch_bounds__(kTX)
conv_wgrad_rows_kernel(const __grid_constant__ ConvArgs<float> a) {
  constexpr int NT = C * 9;
  __shared__ __align__(128) float smem[(kTX / 32) * kRing * C * kRowW];
  __shared__ __align__(128) float gsm[(kTX / 32) * kRing * 128];   // cotangent rows, own vectors only
  __shared__ __align__(8) uint64_t bars[(kTX / 32) * kRing];
  __shared__ float red[kTX / 32][NT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = (blockIdx.x * kTX + threadIdx.x) * kCols;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const int iy0 = r0 - a.ly, iy1 = r1 - 1 - a.ly + 2;
  float acc[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t] = 0.f;

  RowRing<float, C, LX> ring;
  ring.init(smem, bars, warp, lane, a.in, a.plane, j0 - 4 * lane, a.win, a.hin, iy1);
  // cotangent row iy + ly travels with input row iy on the same barrier; rows outside this band
  // read as zero.  The launcher guarantees wout % 4 == 0 and 16-byte aligned rows.
  const int wj0 = j0 - 4 * lane;
  const unsigned grow0 = (unsigned)__cvta_generic_to_shared(gsm) + (unsigned)(warp * kRing * 128 * 4);
  const unsigned gown = grow0 + (unsigned)(4 * lane * 4);
  const unsigned gbytes = wj0 < a.wout ? (unsigned)(min(128, a.wout - wj0) * 4) : 0u;
  const float* gsrc = a.g + wj0;
  for (int i = lane; i < kRing * 128 / 4; i += 32) sts_zero16(grow0 + (unsigned)(i * 16));
  fence_proxy_async();
  __syncwarp();
  auto issue_row = [&](auto stage, int iy) {
    constexpr int S = decltype(stage)::value;
    const int r = iy + a.ly;
    const bool gok = r >= r0 && r < r1 && gbytes != 0;
    ring.template issue<S>(iy, gok ? gbytes : 0u);
    if (gok) {
      if (lane == 0) bulk_g2s(grow0 + (unsigned)(S * 128 * 4), gsrc + (int64_t)r * a.wout, gbytes, ring.mbar + S * 8);
    } else {
      sts_zero16(gown + (unsigned)(S * 128 * 4));
      fence_proxy_async();
    }
  };

  issue_row(std::integral_constant<int, 0>{}, iy0);
  issue_row(std::integral_constant<int, 1>{}, iy0 + 1);

  // g[(P + 3 - ky) % 3]: cotangent row iy + ly - ky while input row iy is processed in phase P
  float g[3][kCols];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int i = 0; i < kCols; ++i) g[s][i] = 0.f;

  unsigned parity = 0;
  auto step = [&](auto phase, int iy) {
    constexpr int P = decltype(phase)::value;
    __syncwarp();
    issue_row(std::integral_constant<int, (P + 2) % 3>{}, iy + 2);
    ring.template wait<P>(parity);
    lds_v4(gown + (unsigned)(P * 128 * 4), g[P % 3]);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float e[6];
      ring.template window<P>(c, e);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        // taps kx = 0, 1 share the cotangent factor: one packed FMA per column; kx = 2 stays scalar
#pragma unroll
        for (int i = 0; i < kCols; ++i) {
          const float gi = g[(P + 3 - ky) % 3][i];
          ffma2(acc[(c * 3 + ky) * 3 + 0], acc[(c * 3 + ky) * 3 + 1], gi, gi, e[i], e[i + 1]);
          acc[(c * 3 + ky) * 3 + 2] = fmaf(gi, e[2 + i], acc[(c * 3 + ky) * 3 + 2]);
        }
      }
    }
  };

#pragma unroll 1
  for (int iy = iy0; iy <= iy1; iy += 3) {
    step(std::integral_constant<int, 0>{}, iy);
    if (iy + 1 > iy1) break;
    step(std::integral_constant<int, 1>{}, iy + 1);
    if (iy + 2 > iy1) break;
    step(std::integral_constant<int, 2>{}, iy + 2);
    parity ^= 1u;
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    float v = acc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < NT) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < kTX / 32; ++wv) v += red[wv][threadIdx.x];
    const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    a.out[blk * NT + threadIdx.x] = v;
  }
}

// fixed-order sum of the per-block partials; slot order -> program tap order through `perm`
struct PermArgs {
  int n_taps;
  int perm[kMaxC * 9];   // dcoef[t] = sum over blocks of partial[block][perm[t]]
};
__global__ void conv_wgrad_finish(const float* __restrict__ partial, float* __restrict__ dcoef, int64_t nblocks,
                                  int nslots, const __grid_constant__ PermArgs pa) {
  const int t = blockIdx.x;
  const int slot = pa.perm[t];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partial[i * nslots + slot];
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    dcoef[t] = s;
  }
}

struct ConvGeom {
  int C, KY, KX, ac, ay, ax;
  int64_t batch;
};

// (C, KY, KX) window over (C, H, W) with the channel axis fully consumed; leading axes batch
bool conv_geometry(const KxGeom& g, ConvGeom& cg) {
  const int nd = g.ndim;
  if (nd < 3) return false;
  cg.ac = nd - 3;
  cg.ay = nd - 2;
  cg.ax = nd - 1;
  cg.C = g.ksize[cg.ac];
  cg.KY = g.ksize[cg.ay];
  cg.KX = g.ksize[cg.ax];
  if (cg.C < 1 || cg.C > kMaxC || cg.KY != 3 || cg.KX != 3) return false;
  if (g.in_shape[cg.ac] != cg.C || g.lo[cg.ac] != 0 || g.hi[cg.ac] != 0 || g.out_shape[cg.ac] != 1) return false;
  cg.batch = 1;
  for (int d = 0; d < nd - 3; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return false;
    cg.batch *= g.in_shape[d];
  }
  for (int d = nd - 3; d < nd; ++d) {
    if (g.stride[d] != 1 || g.in_shape[d] > 0x3fffffff || g.out_shape[d] > 0x3fffffff) return false;
  }
  return true;
}

template <typename T>
void fill_common(ConvArgs<T>& a, const KxGeom& g, const ConvGeom& cg) {
  memset(&a, 0, sizeof(a));
  a.hin = (int)g.in_shape[cg.ay];
  a.win = (int)g.in_shape[cg.ax];
  a.hout = (int)g.out_shape[cg.ay];
  a.wout = (int)g.out_shape[cg.ax];
  a.ly = g.lo[cg.ay];
  a.lx = g.lo[cg.ax];
  a.plane = (int64_t)a.hin * a.win;
  a.in_batch = a.plane * cg.C;
  a.out_batch = (int64_t)a.hout * a.wout;
  a.th = a.hout >= 512 ? 32 : (a.hout >= 64 ? 16 : 8);
  if (const char* e = getenv("KX_CONV_TH")) a.th = atoi(e) > 0 ? atoi(e) : a.th;  // tuning knob
  while ((a.hout + a.th - 1) / a.th > 65535) a.th *= 2;
}

inline bool conv_rows_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("KX_CONV_LEGACY") ? 0 : 1;  // A/B switch for the register-queue kernels
  return on == 1;
}

constexpr int kMinTH = 16;  // the weight-gradient scratch is sized for bands of at least this many rows

// Band height for the row-streaming kernels: every band re-reads 2 halo rows (th/(th+2)) and the
// grid should fill a whole number of resident waves (148 SMs x blocks per SM).
template <typename K>
int pick_band_rows(K kernel, int hout, int64_t tiles_per_band, int fallback) {
  if (getenv("KX_CONV_TH")) return fallback;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTX, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return fallback;
  }
  const int64_t resident = (int64_t)sm_count() * per_sm;
  if (tiles_per_band * ((hout + kMinTH - 1) / kMinTH) <= resident) return fallback;  // small problem
  int best = fallback;
  double best_eff = 0.0;
  for (int th = kMinTH; th <= 96; ++th) {
    const int64_t tiles = tiles_per_band * ((hout + th - 1) / th);
    if ((hout + th - 1) / th > 65535) continue;
    const int64_t waves = (tiles + resident - 1) / resident;
    const double eff = (double)tiles / (double)(waves * resident) * th / (th + 2.0);
    if (eff >= best_eff) {
      best_eff = eff;
      best = th;
    }
  }
  return best;
}

template <typename T, int C, int LX>
int launch_fwd_rows(ConvArgs<T> a, dim3 grid, cudaStream_t s) {
  a.th = pick_band_rows(conv_fwd_rows_kernel<T, C, LX>, a.hout, (int64_t)grid.x * grid.z, a.th);
  grid.y = (a.hout + a.th - 1) / a.th;
  conv_fwd_rows_kernel<T, C, LX><<<grid, kTX, 0, s>>>(a);
  return after_launch("conv_fwd_rows_kernel");
}

template <int C, int LX>
int launch_wgrad_rows(ConvArgs<float> a, dim3& grid, cudaStream_t s) {
  const int th = pick_band_rows(conv_wgrad_rows_kernel<C, LX>, a.hout, (int64_t)grid.x * grid.z, a.th);
  if (th >= kMinTH) a.th = th;
  grid.y = (a.hout + a.th - 1) / a.th;
  conv_wgrad_rows_kernel<C, LX><<<grid, kTX, 0, s>>>(a);
  return after_launch("conv_wgrad_rows_kernel");
}

template <typename T, int C>
int launch_fwd(const ConvArgs<T>& a, dim3 grid, bool vec, cudaStream_t s) {
  if (vec && conv_rows_enabled() && a.lx >= 0 && a.lx <= 2) {
    if (a.lx == 0) return launch_fwd_rows<T, C, 0>(a, grid, s);
    if (a.lx == 1) return launch_fwd_rows<T, C, 1>(a, grid, s);
    return launch_fwd_rows<T, C, 2>(a, grid, s);
  }
  const int shift = vec ? (((-a.lx) % 4) + 4) % 4 : 0;
  if (!vec) conv_fwd_kernel<T, C, 3, 3, 0, false><<<grid, kTX, 0, s>>>(a);
  else if (shift == 0) conv_fwd_kernel<T, C, 3, 3, 0, true><<<grid, kTX, 0, s>>>(a);
  else if (shift == 1) conv_fwd_kernel<T, C, 3, 3, 1, true><<<grid, kTX, 0, s>>>(a);
  else if (shift == 2) conv_fwd_kernel<T, C, 3, 3, 2, true><<<grid, kTX, 0, s>>>(a);
  else conv_fwd_kernel<T, C, 3, 3, 3, true><<<grid, kTX, 0, s>>>(a);
  return after_launch("conv_fwd_kernel");
}

template <int C>
int launch_wgrad(const ConvArgs<float>& a, dim3& grid, bool vec, cudaStream_t s) {
  if (vec && conv_rows_enabled() && a.lx >= 0 && a.lx <= 2 && a.wout % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(a.g) & 15) == 0) {
    if (a.lx == 0) return launch_wgrad_rows<C, 0>(a, grid, s);
    if (a.lx == 1) return launch_wgrad_rows<C, 1>(a, grid, s);
    return launch_wgrad_rows<C, 2>(a, grid, s);
  }
  const int shift = vec ? (((-a.lx) % 4) + 4) % 4 : 0;
  if (!vec) conv_wgrad_kernel<C, 3, 3, 0, false><<<grid, kTX, 0, s>>>(a);
  else if (shift == 0) conv_wgrad_kernel<C, 3, 3, 0, true><<<grid, kTX, 0, s>>>(a);
  else if (shift == 1) conv_wgrad_kernel<C, 3, 3, 1, true><<<grid, kTX, 0, s>>>(a);
  else if (shift == 2) conv_wgrad_kernel<C, 3, 3, 2, true><<<grid, kTX, 0, s>>>(a);
  else conv_wgrad_kernel<C, 3, 3, 3, true><<<grid, kTX, 0, s>>>(a);
  return after_launch("conv_wgrad_kernel");
}

template <typename T>
int run_fwd(const MapArgs& m, const KxProgram& p, const ConvGeom& cg, cudaStream_t s) {
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int NT = cg.C * 9;
  if (f.tap_end - f.tap_begin != NT) return KX_ERR_UNSUPPORTED;  // dense windows only (no tap mask here)
  ConvArgs<T> a;
  fill_common(a, g, cg);
  uint64_t seen = 0;
  bool fwd_order = true, rev_order = true;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int slot = (p.tap_off[t][cg.ac] * 3 + p.tap_off[t][cg.ay]) * 3 + p.tap_off[t][cg.ax];
    if (seen & (1ull << slot)) return KX_ERR_UNSUPPORTED;
    seen |= 1ull << slot;
    a.coef[slot] = (T)p.coef[t];
    if (slot != t - f.tap_begin) fwd_order = false;
    if (slot != NT - 1 - (t - f.tap_begin)) rev_order = false;
  }
  if (m.coef_dev) {
    if (!(fwd_order || rev_order) || (m.coef_batch_stride != 0 && m.batch * cg.batch > 1)) return KX_ERR_UNSUPPORTED;
    a.coef_dev = (const T*)m.coef_dev;
    a.coef_rev = fwd_order ? 0 : 1;
  }
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.bias = (T)f.bias;
  const int64_t batch = m.batch * cg.batch;
  if (batch > 65535) return KX_ERR_UNSUPPORTED;
  const bool vec = (a.win % 4 == 0) && ((reinterpret_cast<uintptr_t>(m.in) & 15) == 0);
  dim3 grid((a.wout + kTX * kCols - 1) / (kTX * kCols), (a.hout + a.th - 1) / a.th, (unsigned)batch);
  switch (cg.C) {
    case 1: return launch_fwd<T, 1>(a, grid, vec, s);
    case 2: return launch_fwd<T, 2>(a, grid, vec, s);
    case 3: return launch_fwd<T, 3>(a, grid, vec, s);
    case 4: return launch_fwd<T, 4>(a, grid, vec, s);
  }
  return KX_ERR_UNSUPPORTED;
}

// ---- dense 3x3x3 stencils over a volume ------------------------------------------------
// out[z, y, x] = sum_kz sum_ky sum_kx w[kz,ky,kx] * in[z+kz-lz, y+ky-ly, x+kx-lx] is the conv above with the
// three input PLANES z-lz .. z-lz+2 as channels and one launch batch item per output plane (batch items
// overlap: in_batch = one plane).  Consecutive output planes run back to back, so two of the three planes
// an output plane reads are still in the 126 MB L2 from its predecessor: DRAM traffic stays 8 B per
// lattice update, the 27 FMAs per output run as packed FFMA2.  Output planes next to a padded z face see
// only one or two real planes: they are separate launches of the C = 1 / C = 2 instantiations on the
// matching slice of the coefficient table (zero planes contribute nothing).
template <typename T>
int run_dense3d(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int nd = g.ndim, az = nd - 3, ay = nd - 2, ax = nd - 1;
  if (f.tap_end - f.tap_begin != 27) return KX_ERR_UNSUPPORTED;
  T coef[27];
  uint32_t seen = 0;
  bool fwd_order = true, rev_order = true;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int slot = (p.tap_off[t][az] * 3 + p.tap_off[t][ay]) * 3 + p.tap_off[t][ax];
    if (slot < 0 || slot >= 27 || (seen & (1u << slot))) return KX_ERR_UNSUPPORTED;
    seen |= 1u << slot;
    coef[slot] = (T)p.coef[t];
    if (slot != t - f.tap_begin) fwd_order = false;
    if (slot != 26 - (t - f.tap_begin)) rev_order = false;
  }
  if (m.coef_dev && !(fwd_order || rev_order)) return KX_ERR_UNSUPPORTED;
  const int din = (int)g.in_shape[az], dout = (int)g.out_shape[az], lz = g.lo[az];
  // interior planes [i0, i1) see three real input planes; the (at most four) others at least one
  const int i0 = std::min(dout, std::max(0, lz)), i1 = std::max(i0, std::min(dout, din - 2 + lz));
  if (i0 + (dout - i1) > 4) return KX_ERR_UNSUPPORTED;
  for (int oz = 0; oz < dout; ++oz) {
    if (oz == i0) oz = i1;
    if (oz < dout && std::max(0, lz - oz) > std::min(2, din - 1 - oz + lz)) return KX_ERR_UNSUPPORTED;
  }
  ConvArgs<T> base;
  memset(&base, 0, sizeof(base));
  base.hin = (int)g.in_shape[ay];
  base.win = (int)g.in_shape[ax];
  base.hout = (int)g.out_shape[ay];
  base.wout = (int)g.out_shape[ax];
  base.ly = g.lo[ay];
  base.lx = g.lo[ax];
  base.plane = (int64_t)base.hin * base.win;
  base.in_batch = base.plane;
  base.out_batch = (int64_t)base.hout * base.wout;
  base.th = base.hout >= 512 ? 32 : (base.hout >= 64 ? 16 : 8);
  while ((base.hout + base.th - 1) / base.th > 65535) base.th *= 2;
  base.bias = (T)f.bias;
  const int tiles_x = (base.wout + kTX * kCols - 1) / (kTX * kCols);

  // planes [oz0, oz1) all use window planes kz0 .. kz0 + C - 1
  auto launch = [&](const T* in, T* out, int oz0, int oz1, int kz0, int C) -> int {
    ConvArgs<T> a = base;
    a.in = in + (int64_t)(oz0 - lz + kz0) * base.plane;
    a.out = out + (int64_t)oz0 * base.out_batch;
    for (int i = 0; i < C * 9; ++i) a.coef[i] = coef[kz0 * 9 + i];
    if (m.coef_dev) {
      a.coef_rev = fwd_order ? 0 : 1;
      a.coef_dev = (const T*)m.coef_dev + (fwd_order ? kz0 * 9 : (3 - kz0 - C) * 9);
    }
    dim3 grid(tiles_x, (base.hout + base.th - 1) / base.th, (unsigned)(oz1 - oz0));
    switch (C) {
      case 1: return launch_fwd<T, 1>(a, grid, true, s);
      case 2: return launch_fwd<T, 2>(a, grid, true, s);
      default: return launch_fwd<T, 3>(a, grid, true, s);
    }
  };

  for (int64_t b = 0; b < batch; ++b) {
    const T* in = (const T*)m.in + b * (int64_t)din * base.plane;          // leading 1-wide axes are the batch
    T* out = (T*)m.out + b * (int64_t)dout * base.out_batch;
    for (int oz = 0; oz < dout; ++oz) {
      if (oz == i0 && i1 > i0) {
        for (int z = i0; z < i1; z += 65535) {
          const int rc = launch(in, out, z, std::min(i1, z + 65535), 0, 3);
          if (rc != KX_OK) return rc;
        }
        oz = i1 - 1;
        continue;
      }
      const int kz0 = std::max(0, lz - oz), kz1 = std::min(2, din - 1 - oz + lz);
      const int rc = launch(in, out, oz, oz + 1, kz0, kz1 - kz0 + 1);
      if (rc != KX_OK) return rc;
    }
  }
  return KX_OK;
}

}  // namespace

// Called by try_stencil3d after the star kernel declined: dense 3x3x3 windows, unit strides, 16-byte aligned rows.
int try_conv_dense3d(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  if (!conv_rows_enabled() || getenv("KX_NO_DENSE3D")) return KX_ERR_UNSUPPORTED;
  if (g.ksize[nd - 3] != 3 || g.ksize[nd - 2] != 3 || g.ksize[nd - 1] != 3) return KX_ERR_UNSUPPORTED;
  if (g.lo[nd - 1] < 0 || g.lo[nd - 1] > 2 || batch < 1 || batch > 8) return KX_ERR_UNSUPPORTED;
  if (g.in_shape[nd - 1] % 4 != 0 || (reinterpret_cast<uintptr_t>(m.in) & 15) != 0) return KX_ERR_UNSUPPORTED;
  if (g.out_shape[nd - 3] < 1 || g.out_shape[nd - 2] < 1 || g.out_shape[nd - 1] < 1) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype == KX_F32) return run_dense3d<float>(m, p, batch, s);
  if (g.in_dtype == KX_I32 && p.func[0].coef_is_int) return run_dense3d<int32_t>(m, p, batch, s);
  return KX_ERR_UNSUPPORTED;
}

int try_conv2d(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (p.n_funcs != 1 || p.func[0].kind != KX_AFFINE || p.func[0].post_div != 0.0) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype || m.fn_elems != 1) return KX_ERR_UNSUPPORTED;
  ConvGeom cg;
  if (!conv_geometry(g, cg) || cg.C < 2) return KX_ERR_UNSUPPORTED;  // C == 1 is stencil2d's job
  if (g.in_dtype == KX_F32) return run_fwd<float>(m, p, cg, s);
  if (g.in_dtype == KX_I32 && p.func[0].coef_is_int) return run_fwd<int32_t>(m, p, cg, s);
  return KX_ERR_UNSUPPORTED;
}

// ---- weight gradient -------------------------------------------------------------------
// geometry accepted: the conv geometry above (C in 1..4, 3x3), or a plain 2-D 3x3 stencil (C = 1)
static bool wgrad_geometry(const KxGeom& g, ConvGeom& cg) {
  if (conv_geometry(g, cg)) return cg.batch == 1;
  if (g.ndim == 2 && g.ksize[0] == 3 && g.ksize[1] == 3 && g.stride[0] == 1 && g.stride[1] == 1 &&
      g.in_shape[0] <= 0x3fffffff && g.in_shape[1] <= 0x3fffffff) {
    cg.C = 1;
    cg.KY = cg.KX = 3;
    cg.ac = -1;
    cg.ay = 0;
    cg.ax = 1;
    cg.batch = 1;
    return true;
  }
  return false;
}

static dim3 wgrad_grid(const ConvArgs<float>& a) {
  return dim3((a.wout + kTX * kCols - 1) / (kTX * kCols), (a.hout + a.th - 1) / a.th, 1);
}

int64_t conv_wgrad_scratch_bytes(const KxGeom& g, const KxProgram& p) {
  ConvGeom cg;
  if (p.n_funcs != 1 || !wgrad_geometry(g, cg)) return 0;
  ConvArgs<float> a;
  fill_common(a, g, cg);
  const dim3 grid = wgrad_grid(a);
  const int64_t bands = a.th < kMinTH ? (int64_t)grid.y : (int64_t)(a.hout + kMinTH - 1) / kMinTH;  // upper bound
  return (int64_t)grid.x * bands * cg.C * 9 * (int64_t)sizeof(float);
}

int try_conv_wgrad(const KxGeom& g, const KxProgram& p, const void* x, const void* gout, void* scratch, void* dcoef,
                   cudaStream_t s) {
  ConvGeom cg;
  if (p.n_funcs != 1 || p.func[0].kind != KX_AFFINE || p.func[0].post_div != 0.0) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != KX_F32 || g.out_dtype != KX_F32 || !wgrad_geometry(g, cg)) return KX_ERR_UNSUPPORTED;
  const KxFunc& f = p.func[0];
  const int NT = cg.C * 9, ntaps = f.tap_end - f.tap_begin;
  if (ntaps < 1 || ntaps > NT) return KX_ERR_UNSUPPORTED;
  ConvArgs<float> a;
  fill_common(a, g, cg);
  PermArgs pa;
  pa.n_taps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    const int oc = cg.ac >= 0 ? p.tap_off[f.tap_begin + t][cg.ac] : 0;
    pa.perm[t] = (oc * 3 + p.tap_off[f.tap_begin + t][cg.ay]) * 3 + p.tap_off[f.tap_begin + t][cg.ax];
  }
  a.in = (const float*)x;
  a.g = (const float*)gout;
  a.out = (float*)scratch;
  const bool vec = (a.win % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  dim3 grid = wgrad_grid(a);
  int rc = KX_ERR_UNSUPPORTED;
  switch (cg.C) {
    case 1: rc = launch_wgrad<1>(a, grid, vec, s); break;
    case 2: rc = launch_wgrad<2>(a, grid, vec, s); break;
    case 3: rc = launch_wgrad<3>(a, grid, vec, s); break;
    case 4: rc = launch_wgrad<4>(a, grid, vec, s); break;
  }
  if (rc) return rc;
  conv_wgrad_finish<<<ntaps, 256, 0, s>>>((const float*)scratch, (float*)dcoef, (int64_t)grid.x * grid.y, NT, pa);
  return after_launch("conv_wgrad_finish");
}

}  // namespace kx
