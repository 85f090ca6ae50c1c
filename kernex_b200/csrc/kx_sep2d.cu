// Fast path: WIDE separable windows -- fp32 coefficient tables that are an outer product cy (x) cx (Gaussian blur of
// README.md:277-294 with kernel sizes 7 .. 15), stride 1, rows 16-byte aligned.  kx_stencil2d.cu recognises the
// outer product and, for square odd windows >= 7x7, hands the program to this kernel; smaller and non-square
// windows stay on stencil2d_kernel's register queue (OP = kSep), which these sizes outgrow: its row buffers and
// queues fill the register file (80-94 registers, 2 rows in flight), ncu shows it latency-bound at 7x7 and
// issue-bound at 15x15 (0.63 / 0.54 / 0.29 of the HBM roofline at 7x7 / 11x11 / 15x15).
//
// Same building blocks as the conv / pooling row rings (kx_conv2d.cu, kx_pool2d.cu):
//   * every WARP streams the input rows of its 128 output columns through a private shared-memory ring filled by TMA
//     bulk copies (one elected lane, clipped to the row; the clipped-away part of the ring is zeroed once and rows
//     outside the array are zero-filled stages = the reference's zero padding), U - 1 = 7 rows ahead, one mbarrier per
//     stage, __syncwarp as the only synchronisation: the bytes in flight cost no registers;
//   * HORIZONTAL fold: a lane reads the KX + 3 ring columns of its 4 outputs with conflict-free LDS.128 and folds them
//     with packed FFMA2 (two outputs per instruction);
//   * VERTICAL fold, input-row stationary: the folded row h feeds the KY output rows it belongs to,
//     acc[(n - ky) mod KY] += cy[ky] * h, so there is no queue to shift; the row loop is unrolled by KY so the
//     accumulator slots are compile-time; window row 0 starts a slot from the bias, the output row whose last
//     window row this was is stored.
// Accumulation order of one output: bias, then ky = 0 .. KY-1 of cy[ky] * (cx[0] x_0 + cx[1] x_1 + ...): the order of
// stencil2d_kernel's kSep path, so the two kernels agree bit for bit (KX_NO_SEP2D=1 selects the old one).
#include <type_traits>
#include <utility>

#include "kx_internal.cuh"
#include "kx_tma.cuh"

namespace kx {

namespace {

constexpr int kST = 128;   // threads per block (4 warps), 4 output columns per lane
constexpr int kSU = 8;     // ring stages

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

template <int KX>
struct SepRowW {
  static constexpr int value = 128 + ((3 + KX - 1 + 3) / 4) * 4 + 4;   // 128 outputs + shift + halo, rounded, + slack vector
};

template <int KY, int KX, int SHIFT>
__global__ void __launch_bounds__(kST)
sep2d_kernel(const __grid_constant__ Sep2dReq a) {
  constexpr int RW = SepRowW<KX>::value;            // floats per ring row
  constexpr int NE = SHIFT + KX + 3;                // ring columns a lane needs, from its aligned base
  constexpr int NV = (NE + 3) / 4;
  static_assert(4 * 31 + NV * 4 <= RW, "ring row too narrow");
  __shared__ __align__(128) float smem[(kST / 32) * kSU * RW];
  __shared__ __align__(8) uint64_t bars[(kST / 32) * kSU];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wo0 = (blockIdx.x * (kST / 32) + warp) * 128;          // first output column of the warp
  if (wo0 >= a.wout) return;
  const int o0 = wo0 + 4 * lane;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const float* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  float* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int iy0 = r0 - a.ly;                                        // first input row of the band
  const int nrows = (r1 - r0) + KY - 1;                             // input rows of the band

  const int c_start = wo0 - a.lx - SHIFT;                           // ring column 0 <-> this input column (multiple of 4)
  const int c_lo = max(c_start, 0), c_hi = min(c_start + RW, a.win);
  const unsigned bytes = c_hi > c_lo ? (unsigned)((c_hi - c_lo) * 4) : 0u;
  const unsigned ring0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)(warp * kSU * RW * 4);
  const unsigned dst_off = (unsigned)((c_lo - c_start) * 4);
  const unsigned own = ring0 + (unsigned)(lane * 16);
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(bars) + (unsigned)(warp * kSU * 8);
  const float* src = in + c_lo;
  const unsigned half = bytes >= 512 ? ((bytes / 2 + 15) & ~15u) : bytes;

  for (int i = lane; i < kSU * RW / 4; i += 32) sts_zero16(ring0 + (unsigned)(i * 16));
  if (lane < kSU) mbar_init(mbar + lane * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncwarp();

  // input row n of the band (array row iy0 + n) -> stage n % kSU
  auto issue = [&](int n) {
    const int iy = iy0 + n, st = n & (kSU - 1);
    const bool ok = (n < nrows) && (iy >= 0) && (iy < a.hin) && bytes != 0;
    const unsigned row = ring0 + (unsigned)(st * RW * 4);
    if (lane == 0) {
      if (ok) {
        mbar_expect_tx(mbar + st * 8, bytes);
        const float* rp = src + (int64_t)iy * a.win;
        bulk_g2s(row + dst_off, rp, half, mbar + st * 8);
        if (bytes > half) bulk_g2s(row + dst_off + half, rp + half / 4, bytes - half, mbar + st * 8);
      } else {
        mbar_arrive(mbar + st * 8);
      }
    }
    if (!ok && n < nrows) {   // a padding row: the warp clears the whole stage row
      for (int i = lane; i < RW / 4; i += 32) sts_zero16(row + (unsigned)(i * 16));
      fence_proxy_async();
      __syncwarp();
    }
  };

  float acc[KY][4];
#pragma unroll
  for (int s = 0; s < KY; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[s][i] = a.bias;
  float* orow = out + (int64_t)r0 * a.wout + o0;
  const int n_valid = min(4, a.wout - o0);

#pragma unroll 1
  for (int k = 0; k < kSU - 1; ++k) issue(k);

  auto step = [&](auto phase, int n) {
    constexpr int PH = decltype(phase)::value;       // n % KY, compile time
    __syncwarp();                                        // all lanes are done with the stage refilled now
    issue(n + kSU - 1);
    const int st = n & (kSU - 1);
    mbar_wait(mbar + st * 8, (unsigned)((n / kSU) & 1));
    // horizontal fold of this input row for the lane's 4 outputs
    float e[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float t4[4];
      lds_v4(own + (unsigned)(st * RW * 4 + v * 16), t4);
#pragma unroll
      for (int q = 0; q < 4; ++q) e[4 * v + q] = t4[q];
    }
    float h[4];
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      float v0 = a.cx[0] * e[SHIFT + i], v1 = a.cx[0] * e[SHIFT + i + 1];
#pragma unroll
      for (int kx = 1; kx < KX; ++kx) ffma2(v0, v1, a.cx[kx], a.cx[kx], e[SHIFT + kx + i], e[SHIFT + kx + i + 1]);
      h[i] = v0;
      h[i + 1] = v1;
    }
    // vertical fold: this row is window row ky of output row n - ky
#pragma unroll
    for (int ky = 0; ky < KY; ++ky) {
      constexpr int kBig = 64 * KY;
      const int slot = (PH - ky + kBig) % KY;
      if (ky == 0) {                                     // first window row of output row n: the chain starts from the bias
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[slot][i] = a.bias;
      }
      ffma2(acc[slot][0], acc[slot][1], a.cy[ky], a.cy[ky], h[0], h[1]);
      ffma2(acc[slot][2], acc[slot][3], a.cy[ky], a.cy[ky], h[2], h[3]);
    }
    if (n >= KY - 1) {                                   // output row n - (KY-1) of the band is complete
      constexpr int slot = (PH + 1) % KY;
      float res[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        res[i] = a.post_mul != 0.f ? acc[slot][i] * a.post_mul : acc[slot][i];
      }
      const uintptr_t addr = reinterpret_cast<uintptr_t>(orow);
      if (n_valid == 4 && (addr & 15) == 0) {
        __stcs(reinterpret_cast<float4*>(orow), make_float4(res[0], res[1], res[2], res[3]));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < n_valid) orow[i] = res[i];
      }
      orow += a.wout;
    }
  };

  // the row loop, unrolled by KY (accumulator slots rotate by renaming)
  int n = 0;
#pragma unroll 1
  for (; n + KY <= nrows; n += KY) static_for<0, KY>([&](auto ph) { step(ph, n + decltype(ph)::value); });
  // tail: fewer than KY rows left
  static_for<0, KY>([&](auto ph) {
    if (n + decltype(ph)::value < nrows) step(ph, n + decltype(ph)::value);
  });
}

template <int K>
int launch_sep(Sep2dReq& r, dim3 grid, int shift, cudaStream_t s) {
  switch (shift) {
    case 0: sep2d_kernel<K, K, 0><<<grid, kST, 0, s>>>(r); break;
    case 1: sep2d_kernel<K, K, 1><<<grid, kST, 0, s>>>(r); break;
    case 2: sep2d_kernel<K, K, 2><<<grid, kST, 0, s>>>(r); break;
    default: sep2d_kernel<K, K, 3><<<grid, kST, 0, s>>>(r); break;
  }
  return after_launch("sep2d_kernel");
}

}  // namespace

int try_sep2d(const Sep2dReq& req, cudaStream_t s) {
  if (getenv("KX_NO_SEP2D")) return KX_ERR_UNSUPPORTED;   // cross-check: stencil2d_kernel's kSep path
  if (req.KY != req.KX || req.KY < 7 || req.KY > 15 || (req.KY & 1) == 0) return KX_ERR_UNSUPPORTED;
  if (req.win % 4 != 0 || (reinterpret_cast<uintptr_t>(req.in) & 15) != 0 || req.in_batch % 4 != 0) return KX_ERR_UNSUPPORTED;
  if (req.batch > 65535 || req.hout < 1 || req.wout < 1) return KX_ERR_UNSUPPORTED;
  Sep2dReq r = req;
  // band height: the K - 1 halo rows of a band are re-read by its neighbour (L2 hits when the bands run together)
  r.th = req.hout >= 1024 ? 128 : (req.hout >= 256 ? 64 : 32);
  if (const char* e = getenv("KX_SEP2D_TH")) r.th = atoi(e) > 0 ? atoi(e) : r.th;
  while ((r.hout + r.th - 1) / r.th > 65535) r.th *= 2;
  dim3 grid((r.wout + 511) / 512, (r.hout + r.th - 1) / r.th, (unsigned)r.batch);
  const int shift = (((-r.lx) % 4) + 4) % 4;
  switch (r.KY) {
    case 7: return launch_sep<7>(r, grid, shift, s);
    case 9: return launch_sep<9>(r, grid, shift, s);
    case 11: return launch_sep<11>(r, grid, shift, s);
    case 13: return launch_sep<13>(r, grid, shift, s);
    case 15: return launch_sep<15>(r, grid, shift, s);
  }
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
