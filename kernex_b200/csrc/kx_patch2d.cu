// Fast path: dense patch extraction over the last two axes (1-D arrays as one row), stride 1
// -- kmap(lambda x: x), README.md:157-186, tests_and_benchmarks/test_benchmark_patch.py,
// BASELINE config 3b (3x3 'same' patches of an 8192^2 image: 4 B read + 36 B written per
// window, i.e. a WRITE-bound kernel).
//
// A block of 128 threads owns 512 consecutive windows of one output row and marches down a
// band of rows:
//   * the KY input row segments a window row needs live in a shared-memory ring; each new
//     output row fetches ONE new input row segment (coalesced), so the input is read once;
//   * every thread builds its 4 windows x (KY*KX) outputs from (KY rows) x (KX+3 columns)
//     it reads with 128/64-bit shared loads -- the patch layout (row-major, or rolled for
//     relative=True, kernex/_src/utils.py:163-181) is a compile-time register permutation;
//   * the 4*KY*KX values are staged in shared memory with conflict-free 128-bit stores and
//     the block then writes the (512*KY*KX)-float output segment with fully coalesced
//     128-bit streaming stores (a thread's own 36 floats are 144 B apart from its
//     neighbour's, which would otherwise cost 32 transactions per store instruction).
// Values are moved as 32-bit words, so one kernel serves float32 and int32 bit-exactly.
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kPT = 128;            // threads per block
constexpr int kWPT = 4;             // windows per thread
constexpr int kTW = kPT * kWPT;     // windows per block row

struct P2Args {
  const uint32_t* in;
  uint32_t* out;
  int64_t in_batch, out_batch;
  int hin, win, hout, wout;
  int ly, lx, th;
  int aligned;         // input rows 16-byte aligned: the warp-independent kernel applies
};

template <int KY, int KX, bool ROLLED>
__global__ void __launch_bounds__(kPT)
patch2d_kernel(const __grid_constant__ P2Args a) {
  constexpr int NT = KY * KX;
  constexpr int PW = kTW + 8;                     // input row pitch in shared memory (>= kTW + KX - 1, multiple of 4)
  constexpr int NC = KX + kWPT - 1;               // input columns a thread touches per row
  __shared__ __align__(16) uint32_t rows[KY][PW];
  __shared__ __align__(16) uint32_t stage[kTW * NT];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTW;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const uint32_t* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  uint32_t* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int nwin = min(kTW, a.wout - x0);          // windows of this block row that exist
  const int count = nwin * NT;                     // output words per row segment

  auto load_row = [&](int iy, int slot) {          // input row iy -> ring slot
    const bool row_ok = (iy >= 0) && (iy < a.hin);
    const uint32_t* rp = in + (int64_t)iy * a.win;
    for (int c = tid; c < kTW + KX - 1; c += kPT) {
      const int gx = x0 - a.lx + c;
      rows[slot][c] = (row_ok && gx >= 0 && gx < a.win) ? __ldg(rp + gx) : 0u;
    }
  };

  // software pipeline: the next input row travels global -> registers while this row's
  // patches are built and written out, and only then registers -> ring (its slot still holds
  // the oldest row of the current window until every thread has finished building)
  constexpr int NLD = (kTW + KX - 1 + kPT - 1) / kPT;
  uint32_t nextv[NLD];
  auto fetch_row = [&](int iy) {
    const bool row_ok = (iy >= 0) && (iy < a.hin);
    const uint32_t* rp = in + (int64_t)iy * a.win;
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int c = tid + k * kPT;
      const int gx = x0 - a.lx + c;
      nextv[k] = (row_ok && c < kTW + KX - 1 && gx >= 0 && gx < a.win) ? __ldg(rp + gx) : 0u;
    }
  };
  auto commit_row = [&](int slot) {
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int c = tid + k * kPT;
      if (c < kTW + KX - 1) rows[slot][c] = nextv[k];
    }
  };

#pragma unroll
  for (int ky = 0; ky < KY; ++ky) load_row(r0 - a.ly + ky, (r0 + ky) % KY);

  for (int r = r0; r < r1; ++r) {
    if (r + 1 < r1) fetch_row(r + 1 - a.ly + KY - 1);
    __syncthreads();                               // ring complete; previous row's staging drained
    // ---- build this thread's 4 windows --------------------------------------------------
    uint32_t v[KY][NC];
#pragma unroll
    for (int ky = 0; ky < KY; ++ky) {
      const uint32_t* rp = rows[(r + ky) % KY] + kWPT * tid;
      const uint4 q = *reinterpret_cast<const uint4*>(rp);
      v[ky][0] = q.x; v[ky][1] = q.y; v[ky][2] = q.z; v[ky][3] = q.w;
#pragma unroll
      for (int e = 4; e < NC; ++e) v[ky][e] = rp[e];
    }
    uint32_t o[kWPT * NT];
#pragma unroll
    for (int w = 0; w < kWPT; ++w) {
#pragma unroll
      for (int ty = 0; ty < KY; ++ty) {
#pragma unroll
        for (int tx = 0; tx < KX; ++tx) {
          // rolled layout: patch index (ty,tx) holds window element ((ty+cy)%KY, (tx+cx)%KX)
          const int sy = ROLLED ? (ty + (KY - 1) / 2) % KY : ty;
          const int sx = ROLLED ? (tx + (KX - 1) / 2) % KX : tx;
          o[w * NT + ty * KX + tx] = v[sy][w + sx];
        }
      }
    }
    static_assert((kWPT * NT) % 4 == 0, "a thread's outputs must be whole 128-bit words");
    uint32_t* sp = stage + tid * (kWPT * NT);
#pragma unroll
    for (int q = 0; q < kWPT * NT / 4; ++q)
      *reinterpret_cast<uint4*>(sp + 4 * q) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    __syncthreads();
    if (r + 1 < r1) commit_row((r + KY) % KY);     // every thread has finished reading the ring
    // ---- coalesced write-out of the row segment ------------------------------------------
    uint32_t* dst = out + ((int64_t)r * a.wout + x0) * NT;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      for (int i = tid; 4 * i + 3 < count; i += kPT)
        __stcs(reinterpret_cast<uint4*>(dst) + i, *reinterpret_cast<const uint4*>(stage + 4 * i));
      for (int i = (count & ~3) + tid; i < count; i += kPT) dst[i] = stage[i];
    } else {
      for (int i = tid; i < count; i += kPT) dst[i] = stage[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-independent variant for 16-byte aligned input rows (KX <= 3): no block barriers at all.
// A warp owns 128 consecutive windows of a row and marches down the band.  Input: every lane
// loads only the aligned vector of its own 4 columns per new row (prefetched one row ahead), the x
// halo comes from the neighbouring lanes by shuffle (edge lanes: scalar loads); the last KY rows
// live in registers.  Output: the lane's 4 * KY*KX words go to a warp-private staging buffer
// (conflict-free 128-bit stores; double buffered, so ONE __syncwarp per row) and the warp writes
// its 128 * KY*KX contiguous words with fully coalesced 128-bit streaming stores.
template <int KY, int KX, bool ROLLED, int LX>
__global__ void __launch_bounds__(kPT)
patch2d_warp_kernel(const __grid_constant__ P2Args a) {
  constexpr int NT = KY * KX;
  constexpr int NC = KX + kWPT - 1;                 // input columns a lane needs per row (<= 6)
  constexpr int RX = KX - 1 - LX;                   // columns needed right of the lane's vector
  __shared__ __align__(16) uint32_t stage[kPT / 32][2][128 * NT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = (blockIdx.x * (kPT / 32) + warp) * 128;     // first window of the warp
  const int j0 = x0 + 4 * lane;                               // first window (= aligned input column) of the lane
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const uint32_t* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  uint32_t* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  if (x0 >= a.wout) return;                                    // whole warp outside (warp-uniform)
  const int nwin = min(128, a.wout - x0);
  const int count = nwin * NT;                                 // output words of the warp per row
  const bool first = lane == 0, last = lane == 31;

  // input columns j0-LX .. j0-LX+NC-1 of row iy, as raw 32-bit words (zero outside the array)
  struct Raw {
    uint32_t v[4], el[LX > 0 ? LX : 1], er[RX > 0 ? RX : 1];
  };
  auto fetch = [&](int iy, Raw& r) {
    const bool row_ok = (iy >= 0) && (iy < a.hin);
    const uint32_t* rp = in + (int64_t)iy * a.win;
    if (row_ok && j0 + 3 < a.win) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp + j0));
      r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) r.v[e] = (row_ok && j0 + e < a.win) ? __ldg(rp + j0 + e) : 0u;
    }
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      const int c = j0 - LX + k;
      r.el[k] = (first && row_ok && c >= 0 && c < a.win) ? __ldg(rp + c) : 0u;
    }
#pragma unroll
    for (int k = 0; k < RX; ++k) {
      const int c = j0 + 4 + k;
      r.er[k] = (last && row_ok && c < a.win) ? __ldg(rp + c) : 0u;
    }
  };
  auto assemble = [&](const Raw& r, uint32_t (&e)[NC]) {
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, r.v[4 - LX + k], 1);
      e[k] = first ? r.el[k] : up;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) e[LX + k] = r.v[k];
#pragma unroll
    for (int k = 0; k < RX; ++k) {
      const uint32_t dn = __shfl_down_sync(0xffffffffu, r.v[k], 1);
      e[LX + 4 + k] = last ? r.er[k] : dn;
    }
  };

  uint32_t q[KY][NC];
#pragma unroll
  for (int ky = 0; ky < KY - 1; ++ky) {
    Raw t;
    fetch(r0 - a.ly + ky, t);
    assemble(t, q[ky]);
  }
  Raw nxt;
  fetch(r0 - a.ly + KY - 1, nxt);
  uint32_t* dst = out + ((int64_t)r0 * a.wout + x0) * NT;
  const int64_t dst_pitch = (int64_t)a.wout * NT;
  for (int r = r0; r < r1; ++r) {
    assemble(nxt, q[KY - 1]);
    if (r + 1 < r1) fetch(r + 1 - a.ly + KY - 1, nxt);          // in flight while this row is built and written
    uint32_t o[kWPT * NT];
#pragma unroll
    for (int w = 0; w < kWPT; ++w)
#pragma unroll
      for (int ty = 0; ty < KY; ++ty)
#pragma unroll
        for (int tx = 0; tx < KX; ++tx) {
          const int sy = ROLLED ? (ty + (KY - 1) / 2) % KY : ty;
          const int sx = ROLLED ? (tx + (KX - 1) / 2) % KX : tx;
          o[w * NT + ty * KX + tx] = q[sy][w + sx];
        }
    uint32_t* sb = stage[warp][(r - r0) & 1];
    uint32_t* sp = sb + lane * (kWPT * NT);
#pragma unroll
    for (int k = 0; k < kWPT * NT / 4; ++k)
      *reinterpret_cast<uint4*>(sp + 4 * k) = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    __syncwarp();
    if (count == 128 * NT && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int k = 0; k < NT; ++k)
        __stcs(reinterpret_cast<uint4*>(dst) + lane + 32 * k, *reinterpret_cast<const uint4*>(sb + 4 * (lane + 32 * k)));
    } else {
      for (int i = lane; i < count; i += 32) dst[i] = sb[i];
    }
    dst += dst_pitch;
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky)
#pragma unroll
      for (int e = 0; e < NC; ++e) q[ky][e] = q[ky + 1][e];
  }
}

template <int KY, int KX, bool ROLLED>
int launch_p2_warp(const P2Args& a, int64_t batch, cudaStream_t s) {
  dim3 grid((a.wout + kTW - 1) / kTW, (a.hout + a.th - 1) / a.th, (unsigned)batch);
  if (a.lx == 0) patch2d_warp_kernel<KY, KX, ROLLED, 0><<<grid, kPT, 0, s>>>(a);
  else if (KX >= 2 && a.lx == 1) patch2d_warp_kernel<KY, KX, ROLLED, (KX >= 2 ? 1 : 0)><<<grid, kPT, 0, s>>>(a);
  else if (KX >= 3 && a.lx == 2) patch2d_warp_kernel<KY, KX, ROLLED, (KX >= 3 ? 2 : 0)><<<grid, kPT, 0, s>>>(a);
  else return KX_ERR_UNSUPPORTED;
  return after_launch("patch2d_kernel");
}

// ---------------------------------------------------------------------------------------------
// Larger windows (5x5, 7x7, 4x4, 1x7, ...): 4 windows x KY*KX outputs no longer fit a thread's registers (5x5: 100), so
// nothing is staged -- the block walks the row segment's output words in ORDER (fully coalesced 128-bit streaming
// stores) and picks every word out of the shared-memory row ring: word i of the segment is tap i % NT of window i / NT
// (compile-time divisors).  Write-bound like the small windows: 4 B read + 4 KY KX B written per window; before this
// kernel these shapes ran on patch_kernel (flat gather through L1: 0.15 of the HBM roofline).
template <int KY, int KX, bool ROLLED>
__global__ void __launch_bounds__(kPT)
patch2d_direct_kernel(const __grid_constant__ P2Args a) {
  constexpr int NT = KY * KX;
  constexpr int TW = 256;                           // windows per block row
  constexpr int PW = TW + ((KX - 1 + 3) / 4) * 4;
  __shared__ __align__(16) uint32_t rows[KY][PW];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TW;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const uint32_t* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  uint32_t* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int nwin = min(TW, a.wout - x0);
  const int count = nwin * NT;
  auto load_row = [&](int iy, int slot) {
    const bool row_ok = (iy >= 0) && (iy < a.hin);
    const uint32_t* rp = in + (int64_t)iy * a.win;
    for (int c = tid; c < TW + KX - 1; c += kPT) {
      const int gx = x0 - a.lx + c;
      rows[slot][c] = (row_ok && gx >= 0 && gx < a.win) ? __ldg(rp + gx) : 0u;
    }
  };
  auto word = [&](int i, int rb) -> uint32_t {      // output word i of the row segment; rb = r % KY
    const int w = i / NT, t = i - w * NT;
    const int ty = t / KX, tx = t - ty * KX;
    // rolled layout: patch index (ty,tx) holds window element ((ty+cy)%KY, (tx+cx)%KX)
    const int sy = ROLLED ? (ty + (KY - 1) / 2) % KY : ty;
    const int sx = ROLLED ? (tx + (KX - 1) / 2) % KX : tx;
    int slot = rb + sy;
    if (slot >= KY) slot -= KY;
    return rows[slot][w + sx];
  };
#pragma unroll 1
  for (int ky = 0; ky < KY - 1; ++ky) load_row(r0 - a.ly + ky, (r0 + ky) % KY);
  for (int r = r0; r < r1; ++r) {
    load_row(r - a.ly + KY - 1, (r + KY - 1) % KY);
    __syncthreads();
    const int rb = r % KY;
    uint32_t* dst = out + ((int64_t)r * a.wout + x0) * NT;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      for (int v = tid; 4 * v + 3 < count; v += kPT)
        __stcs(reinterpret_cast<uint4*>(dst) + v, make_uint4(word(4 * v, rb), word(4 * v + 1, rb), word(4 * v + 2, rb), word(4 * v + 3, rb)));
      for (int i = (count & ~3) + tid; i < count; i += kPT) dst[i] = word(i, rb);
    } else {
      for (int i = tid; i < count; i += kPT) dst[i] = word(i, rb);
    }
    __syncthreads();                                 // the oldest row's slot is refilled next
  }
}

template <int KY, int KX>
int launch_p2_direct(const P2Args& a, bool rolled, int64_t batch, cudaStream_t s) {
  dim3 grid((a.wout + 255) / 256, (a.hout + a.th - 1) / a.th, (unsigned)batch);
  if (rolled) patch2d_direct_kernel<KY, KX, true><<<grid, kPT, 0, s>>>(a);
  else patch2d_direct_kernel<KY, KX, false><<<grid, kPT, 0, s>>>(a);
  return after_launch("patch2d_direct_kernel");
}

template <int KY, int KX>
int launch_p2(const P2Args& a, bool rolled, int64_t batch, cudaStream_t s) {
  if (KX <= 3 && a.aligned && a.lx >= 0 && a.lx <= KX - 1 && !getenv("KX_PATCH_BLOCK")) {
    // short bands: measured on C3b (8192^2): th = 3..4 164 GLUP/s, 8 160, 16 155, 32 151, 128 141 -- the input is
    // only 10 % of the traffic, so re-reading two halo rows per band costs less than the coarser write pattern
    P2Args w = a;
    if (!getenv("KX_PATCH_TH") && w.hout >= 64) w.th = 4;
    while ((w.hout + w.th - 1) / w.th > 65535) w.th *= 2;
    const int rc = rolled ? launch_p2_warp<KY, KX, true>(w, batch, s) : launch_p2_warp<KY, KX, false>(w, batch, s);
    if (rc != KX_ERR_UNSUPPORTED) return rc;
  }
  dim3 grid((a.wout + kTW - 1) / kTW, (a.hout + a.th - 1) / a.th, (unsigned)batch);
  if (rolled) patch2d_kernel<KY, KX, true><<<grid, kPT, 0, s>>>(a);
  else patch2d_kernel<KY, KX, false><<<grid, kPT, 0, s>>>(a);
  return after_launch("patch2d_kernel");
}

// ---------------------------------------------------------------------------------------------
// 3-D windows (3x3x3, 2x2x2 patches of a volume: 4 B read + 4 KZ KY KX B written per window): patch2d_direct_kernel with
// KZ planes in the ring -- a block owns 256 windows of an output row, marches down a band of rows of ONE output plane
// and fetches KZ new input row segments (one per window plane) per output row; the rows of neighbouring output planes
// are re-read by their own blocks (L2 hits, and 3 x 4 B against 108 B written per window for 3x3x3).
template <int KZ, int KY, int KX, bool ROLLED>
__global__ void __launch_bounds__(kPT)
patch3d_direct_kernel(const __grid_constant__ P2Args a, int din, int dout, int lz) {
  constexpr int NT = KZ * KY * KX;
  constexpr int TW = 256;
  constexpr int PW = TW + ((KX - 1 + 3) / 4) * 4;
  __shared__ __align__(16) uint32_t rows[KZ][KY][PW];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TW;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const int b = blockIdx.z / dout, z = blockIdx.z - b * dout;
  const uint32_t* __restrict__ in = a.in + (int64_t)b * a.in_batch;
  uint32_t* __restrict__ out = a.out + (int64_t)b * a.out_batch + (int64_t)z * a.hout * a.wout * NT;
  const int nwin = min(TW, a.wout - x0);
  const int count = nwin * NT;
  const int64_t plane = (int64_t)a.hin * a.win;
  auto load_rows = [&](int iy, int slot) {           // input row iy of the KZ window planes -> ring slot
    const bool row_ok = (iy >= 0) && (iy < a.hin);
#pragma unroll
    for (int kz = 0; kz < KZ; ++kz) {
      const int iz = z - lz + kz;
      const bool ok = row_ok && iz >= 0 && iz < din;
      const uint32_t* rp = in + (int64_t)iz * plane + (int64_t)iy * a.win;
      for (int c = tid; c < TW + KX - 1; c += kPT) {
        const int gx = x0 - a.lx + c;
        rows[kz][slot][c] = (ok && gx >= 0 && gx < a.win) ? __ldg(rp + gx) : 0u;
      }
    }
  };
  auto word = [&](int i, int rb) -> uint32_t {
    const int w = i / NT, t = i - w * NT;
    const int tz = t / (KY * KX), t2 = t - tz * (KY * KX);
    const int ty = t2 / KX, tx = t2 - ty * KX;
    const int sz = ROLLED ? (tz + (KZ - 1) / 2) % KZ : tz;
    const int sy = ROLLED ? (ty + (KY - 1) / 2) % KY : ty;
    const int sx = ROLLED ? (tx + (KX - 1) / 2) % KX : tx;
    int slot = rb + sy;
    if (slot >= KY) slot -= KY;
    return rows[sz][slot][w + sx];
  };
#pragma unroll 1
  for (int ky = 0; ky < KY - 1; ++ky) load_rows(r0 - a.ly + ky, (r0 + ky) % KY);
  for (int r = r0; r < r1; ++r) {
    load_rows(r - a.ly + KY - 1, (r + KY - 1) % KY);
    __syncthreads();
    const int rb = r % KY;
    uint32_t* dst = out + ((int64_t)r * a.wout + x0) * NT;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      for (int v = tid; 4 * v + 3 < count; v += kPT)
        __stcs(reinterpret_cast<uint4*>(dst) + v, make_uint4(word(4 * v, rb), word(4 * v + 1, rb), word(4 * v + 2, rb), word(4 * v + 3, rb)));
      for (int i = (count & ~3) + tid; i < count; i += kPT) dst[i] = word(i, rb);
    } else {
      for (int i = tid; i < count; i += kPT) dst[i] = word(i, rb);
    }
    __syncthreads();
  }
}

template <int KZ, int KY, int KX>
int launch_p3_direct(const P2Args& a, int din, int dout, int lz, bool rolled, int64_t batch, cudaStream_t s) {
  dim3 grid((a.wout + 255) / 256, (a.hout + a.th - 1) / a.th, (unsigned)(batch * dout));
  if (rolled) patch3d_direct_kernel<KZ, KY, KX, true><<<grid, kPT, 0, s>>>(a, din, dout, lz);
  else patch3d_direct_kernel<KZ, KY, KX, false><<<grid, kPT, 0, s>>>(a, din, dout, lz);
  return after_launch("patch3d_direct_kernel");
}

}  // namespace

// dense patch extraction over the last THREE axes (stride 1), plain or rolled tap order
int try_patch3d(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  if (nd < 3 || p.n_funcs != 1 || p.func[0].kind != KX_GATHER || g.in_dtype != g.out_dtype) return KX_ERR_UNSUPPORTED;
  const int az = nd - 3, ay = nd - 2, ax = nd - 1;
  int64_t lead = 1;
  for (int d = 0; d < az; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    lead *= g.in_shape[d];
  }
  if (g.stride[az] != 1 || g.stride[ay] != 1 || g.stride[ax] != 1) return KX_ERR_UNSUPPORTED;
  const int KZ = g.ksize[az], KY = g.ksize[ay], KX = g.ksize[ax];
  if (KZ < 2) return KX_ERR_UNSUPPORTED;                       // 2-D windows: try_patch2d
  const KxFunc& f = p.func[0];
  if (f.tap_end - f.tap_begin != KZ * KY * KX) return KX_ERR_UNSUPPORTED;
  bool plain = true, rolled = true;
  for (int t = 0; t < KZ * KY * KX; ++t) {
    const int tz = t / (KY * KX), ty = (t / KX) % KY, tx = t % KX;
    const int oz = p.tap_off[f.tap_begin + t][az], oy = p.tap_off[f.tap_begin + t][ay], ox = p.tap_off[f.tap_begin + t][ax];
    if (oz != tz || oy != ty || ox != tx) plain = false;
    if (oz != (tz + (KZ - 1) / 2) % KZ || oy != (ty + (KY - 1) / 2) % KY || ox != (tx + (KX - 1) / 2) % KX) rolled = false;
    for (int d = 0; d < az; ++d)
      if (p.tap_off[f.tap_begin + t][d] != 0) return KX_ERR_UNSUPPORTED;
  }
  if (!plain && !rolled) return KX_ERR_UNSUPPORTED;
  const int64_t batch = m.batch * lead;
  if (g.in_shape[ax] > 0x3fffffff || g.in_shape[ay] > 0x3fffffff || g.in_shape[az] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
  if (batch * g.out_shape[az] > 65535) return KX_ERR_UNSUPPORTED;
  P2Args a;
  memset(&a, 0, sizeof(a));
  a.in = (const uint32_t*)m.in;
  a.out = (uint32_t*)m.out;
  a.hin = (int)g.in_shape[ay];
  a.win = (int)g.in_shape[ax];
  a.hout = (int)g.out_shape[ay];
  a.wout = (int)g.out_shape[ax];
  a.ly = g.lo[ay];
  a.lx = g.lo[ax];
  a.in_batch = (int64_t)g.in_shape[az] * a.hin * a.win;
  a.out_batch = (int64_t)g.out_shape[az] * a.hout * a.wout * KZ * KY * KX;
  a.th = a.hout >= 64 ? 16 : 4;
  while ((a.hout + a.th - 1) / a.th > 65535) a.th *= 2;
  const int din = (int)g.in_shape[az], dout = (int)g.out_shape[az], lz = g.lo[az];
  if (KZ == 3 && KY == 3 && KX == 3) return launch_p3_direct<3, 3, 3>(a, din, dout, lz, !plain, batch, s);
  if (KZ == 2 && KY == 2 && KX == 2) return launch_p3_direct<2, 2, 2>(a, din, dout, lz, !plain, batch, s);
  return KX_ERR_UNSUPPORTED;
}

namespace {
}  // namespace

int try_patch2d(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (p.n_funcs != 1 || p.func[0].kind != KX_GATHER || g.in_dtype != g.out_dtype) return KX_ERR_UNSUPPORTED;
  const int nd = g.ndim, ax = nd - 1, ay = nd - 2;
  int64_t lead = 1;
  for (int d = 0; d < nd - 2; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    lead *= g.in_shape[d];
  }
  if (g.stride[ax] != 1 || (ay >= 0 && g.stride[ay] != 1)) return KX_ERR_UNSUPPORTED;
  const int KY = ay >= 0 ? g.ksize[ay] : 1, KX = g.ksize[ax];
  const KxFunc& f = p.func[0];
  if (f.tap_end - f.tap_begin != KY * KX) return KX_ERR_UNSUPPORTED;
  // tap order must be the plain row-major window or its roll_view rotation
  bool plain = true, rolled = true;
  for (int t = 0; t < KY * KX; ++t) {
    const int ty = t / KX, tx = t % KX;
    const int oy = ay >= 0 ? p.tap_off[f.tap_begin + t][ay] : 0, ox = p.tap_off[f.tap_begin + t][ax];
    if (oy != ty || ox != tx) plain = false;
    if (oy != (ty + (KY - 1) / 2) % KY || ox != (tx + (KX - 1) / 2) % KX) rolled = false;
    for (int d = 0; d < nd - 2; ++d)
      if (p.tap_off[f.tap_begin + t][d] != 0) return KX_ERR_UNSUPPORTED;
  }
  if (!plain && !rolled) return KX_ERR_UNSUPPORTED;
  const int64_t batch = m.batch * lead;
  if (batch > 65535) return KX_ERR_UNSUPPORTED;
  P2Args a;
  memset(&a, 0, sizeof(a));
  a.in = (const uint32_t*)m.in;
  a.out = (uint32_t*)m.out;
  a.hin = ay >= 0 ? (int)g.in_shape[ay] : 1;
  a.win = (int)g.in_shape[ax];
  a.hout = ay >= 0 ? (int)g.out_shape[ay] : 1;
  a.wout = (int)g.out_shape[ax];
  if (g.in_shape[ax] > 0x3fffffff || (ay >= 0 && g.in_shape[ay] > 0x3fffffff)) return KX_ERR_UNSUPPORTED;
  a.ly = ay >= 0 ? g.lo[ay] : 0;
  a.lx = g.lo[ax];
  a.in_batch = (int64_t)a.hin * a.win;
  a.out_batch = (int64_t)a.hout * a.wout * KY * KX;
  a.th = a.hout >= 512 ? 16 : (a.hout >= 64 ? 8 : 4);
  if (const char* e = getenv("KX_PATCH_TH")) a.th = atoi(e) > 0 ? atoi(e) : a.th;   // tuning knob
  while ((a.hout + a.th - 1) / a.th > 65535) a.th *= 2;
  const bool roll = !plain;
  a.aligned = (a.win % 4 == 0) && ((reinterpret_cast<uintptr_t>(m.in) & 15) == 0) && (a.in_batch % 4 == 0);
  if (KY == 3 && KX == 3) return launch_p2<3, 3>(a, roll, batch, s);
  if (KY == 1 && KX == 3) return launch_p2<1, 3>(a, roll, batch, s);
  if (KY == 2 && KX == 2) return launch_p2<2, 2>(a, roll, batch, s);
  if (KY == 1 && KX == 5) return launch_p2<1, 5>(a, roll, batch, s);
  if (KY == 5 && KX == 5) return launch_p2_direct<5, 5>(a, roll, batch, s);
  if (KY == 7 && KX == 7) return launch_p2_direct<7, 7>(a, roll, batch, s);
  if (KY == 4 && KX == 4) return launch_p2_direct<4, 4>(a, roll, batch, s);
  if (KY == 3 && KX == 5) return launch_p2_direct<3, 5>(a, roll, batch, s);
  if (KY == 5 && KX == 3) return launch_p2_direct<5, 3>(a, roll, batch, s);
  if (KY == 2 && KX == 3) return launch_p2_direct<2, 3>(a, roll, batch, s);
  if (KY == 3 && KX == 2) return launch_p2_direct<3, 2>(a, roll, batch, s);
  if (KY == 1 && KX == 2) return launch_p2_direct<1, 2>(a, roll, batch, s);
  if (KY == 1 && KX == 4) return launch_p2_direct<1, 4>(a, roll, batch, s);
  if (KY == 1 && KX == 7) return launch_p2_direct<1, 7>(a, roll, batch, s);
  if (KY == 1 && KX == 9) return launch_p2_direct<1, 9>(a, roll, batch, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
