// kscan as a persistent hyperplane-wavefront kernel.
//
// kernel_scan (kernex/_src/scan.py:84-113) sweeps all windows in row-major order, carrying the
// padded array and writing each result at the window centre before the next window is read
// (scan.py:46-62).  Every cell is written at most once (window j owns the cell j*s - lo + (k-1)/2),
// so what a window reads from cell x is fully determined by WHO writes x:
//   * nobody, or a window later in the sweep  -> the original (zero padded) input   [old]
//   * a window earlier in the sweep           -> that window's result              [new]
// Reading old values from `in` and new values from `out` removes every write-after-read hazard and
// leaves only true dependencies.  A tap at centre-relative cell offset `off` depends on window
// j + delta, delta = off / stride (when divisible in every axis); it is a true dependency when
// delta is lexicographically negative (positive for lax.scan(reverse=True)).
//
// Schedule: t(j) = sum_d a_d * j_d with non-negative integers a_d chosen on the host so that
// a . delta < 0 for every true dependency of every function of the mesh.  All windows of one
// hyperplane t are independent and run in parallel; hyperplanes are separated by a block barrier
// (small fronts: one CTA) or a cooperative grid barrier (large fronts).  Examples: time marching
// (taps on the row above only) a = (1, 0): a whole row per step; Gauss-Seidel / README RK4
// (x[0,-1], x[-1,0]) a = (1, 1): anti-diagonals; a 3x3 window reading x[-1,+1] a = (2, 1).
//
// The row-marching kernel (kx_scan_rowmarch.cu) takes the HBM-bound time-marching case; this one is
// the general path and is bit-identical to the one-thread sweep (scan_sequential_kernel), which is
// kept behind KX_SCAN_SEQUENTIAL=1 as a cross-check.
#include <cooperative_groups.h>

#include "kx_internal.cuh"

namespace cg = cooperative_groups;

namespace kx {

namespace {

constexpr int kWaveThreads = 256;
constexpr int kWaveSingleThreads = 1024;
constexpr int64_t kWaveSingleMax = 8192;  // fronts up to this size run on one CTA (block barrier ~10x cheaper)

struct WaveArgs {
  KxGeom g;
  int64_t n_steps;
  int64_t n_free;                    // threads of work per hyperplane
  int32_t a[KX_MAX_DIMS];            // hyperplane coefficients
  int32_t solved;                    // axis recovered from t (a[solved] != 0) or -1
  int32_t reverse;
};

template <typename TIN>
__device__ __forceinline__ TIN ld_new(const TIN* p) {
  return __ldcg(p);  // results of other threads: L2 is the coherence point
}

// IDX = int32_t when every extent and the window count fit 31 bits (64-bit divisions cost ~100
// instructions each on the GPU); UNIT = all strides are 1 (the writer of a cell needs no division).
template <typename TIN, typename TC, typename IDX, bool UNIT>
__device__ __forceinline__ void wave_window(const WaveArgs& a, const DevProgram<TC>& p, const Strides& ist,
                                            const Strides& ost, const TIN* __restrict__ in, TIN* out,
                                            const TC* __restrict__ coef_dev, const IDX* j) {
  const int nd = a.g.ndim;
  int64_t centre[KX_MAX_DIMS];
  IDX origin[KX_MAX_DIMS];
  int64_t w = 0;
#pragma unroll
  for (int d = 0; d < KX_MAX_DIMS; ++d) {
    if (d < nd) {
      origin[d] = j[d] * (IDX)a.g.stride[d] - (IDX)a.g.lo[d];
      centre[d] = (int64_t)origin[d] + (a.g.ksize[d] - 1) / 2;
      w += (int64_t)j[d] * ost.v[d];
    }
  }
  const int f = select_func(p, centre);
  const int kind = p.func[f].kind;
  const int t0 = p.func[f].tap_begin, t1 = p.func[f].tap_end;
  TC acc = (kind == KX_AFFINE) ? p.func[f].bias : TC(0);
  for (int t = t0; t < t1; ++t) {
    int64_t ioff = 0, woff = 0;
    bool inb = true, writer = true;
    int cmp = 0;
#pragma unroll
    for (int d = 0; d < KX_MAX_DIMS; ++d) {
      if (d < nd) {
        const IDX c = origin[d] + (IDX)p.tap_off[t][d];
        inb = inb && (c >= 0) && (c < (IDX)a.g.in_shape[d]);
        ioff += (int64_t)c * ist.v[d];
        // q = jw * stride when the cell is a window centre
        const IDX q = (IDX)p.tap_off[t][d] - (IDX)((a.g.ksize[d] - 1) / 2) + j[d] * (IDX)a.g.stride[d];
        const IDX jw = UNIT ? q : q / (IDX)a.g.stride[d];
        writer = writer && (q >= 0) && (UNIT || jw * (IDX)a.g.stride[d] == q) && (jw < (IDX)a.g.out_shape[d]);
        woff += (int64_t)jw * ost.v[d];
        if (cmp == 0) cmp = (jw < j[d]) ? -1 : ((jw > j[d]) ? 1 : 0);
      }
    }
    const bool is_new = writer && (a.reverse ? cmp > 0 : cmp < 0);
    TC v;
    if (is_new) v = (TC)ld_new(out + woff);
    else v = inb ? (TC)in[ioff] : TC(0);
    if (kind == KX_AFFINE) {
      const TC c = coef_dev ? coef_dev[t] : p.coef[t];
      acc = mad_wrap(c, v, acc);
    } else if (kind == KX_REDUCE_MAX) {
      acc = (t == t0) ? v : red_max(acc, v);
    } else if (kind == KX_REDUCE_MIN) {
      acc = (t == t0) ? v : red_min(acc, v);
    } else {  // scalar KX_GATHER: x[tap]
      acc = v;
    }
  }
  if (kind == KX_AFFINE && p.func[f].post_div != 0.f) acc = (TC)((float)acc / p.func[f].post_div);
  out[w] = (TIN)acc;  // cast to the carried dtype on write (scan.py:50)
}

template <typename TIN, typename TC, typename IDX, bool UNIT>
__global__ void __launch_bounds__(kWaveSingleThreads)
scan_wavefront_kernel(const __grid_constant__ WaveArgs a, const __grid_constant__ DevProgram<TC> p,
                      const TIN* __restrict__ in, TIN* out, const TC* __restrict__ coef_dev) {
  const int nd = a.g.ndim;
  const Strides ist = suffix_strides(a.g.in_shape, nd);
  const Strides ost = suffix_strides(a.g.out_shape, nd);
  const bool multi = gridDim.x > 1;
  const IDX tid = (IDX)blockIdx.x * (IDX)blockDim.x + (IDX)threadIdx.x;
  const IDX nthreads = (IDX)gridDim.x * (IDX)blockDim.x;
  const IDX n_free = (IDX)a.n_free;
  for (int64_t t = 0; t < a.n_steps; ++t) {
    for (IDX i = tid; i < n_free; i += nthreads) {
      IDX jp[KX_MAX_DIMS];  // sweep-order coordinates (mirrored when reverse)
      IDX r = i;
      int64_t rem = t;
#pragma unroll
      for (int d = KX_MAX_DIMS - 1; d >= 0; --d) {
        if (d < nd && d != a.solved) {
          const IDX n = (IDX)a.g.out_shape[d];
          const IDX qd = r / n;
          jp[d] = r - qd * n;
          r = qd;
          rem -= (int64_t)a.a[d] * jp[d];
        }
      }
      if (a.solved >= 0) {
        const int64_t as = a.a[a.solved];
        if (rem < 0 || rem >= as * a.g.out_shape[a.solved]) continue;
        IDX js = (IDX)rem;
        if (as != 1) {
          js = (IDX)(rem / as);
          if ((int64_t)js * as != rem) continue;
        }
#pragma unroll
        for (int d = 0; d < KX_MAX_DIMS; ++d)
          if (d == a.solved) jp[d] = js;
      }
      IDX j[KX_MAX_DIMS];
#pragma unroll
      for (int d = 0; d < KX_MAX_DIMS; ++d)
        if (d < nd) j[d] = a.reverse ? (IDX)a.g.out_shape[d] - 1 - jp[d] : jp[d];
      wave_window<TIN, TC, IDX, UNIT>(a, p, ist, ost, in, out, coef_dev, j);
    }
    if (t + 1 < a.n_steps) {
      if (multi) {
        cg::this_grid().sync();
      } else {
        __threadfence_block();
        __syncthreads();
      }
    }
  }
}

// smallest non-negative integer hyperplane a with a . delta < 0 for every true dependency
bool plan_wavefront(const KxGeom& g, const KxProgram& prog, int reverse, WaveArgs& a) {
  const int nd = g.ndim;
  int64_t acoef[KX_MAX_DIMS] = {0, 0, 0, 0};
  for (int d = nd - 1; d >= 0; --d) {
    int64_t need = 0;
    for (int f = 0; f < prog.n_funcs; ++f) {
      for (int t = prog.func[f].tap_begin; t < prog.func[f].tap_end; ++t) {
        int64_t delta[KX_MAX_DIMS];
        bool hits = true;
        for (int e = 0; e < nd; ++e) {
          const int off = prog.tap_off[t][e] - (g.ksize[e] - 1) / 2;
          if (off % g.stride[e] != 0) hits = false;
          delta[e] = (off / g.stride[e]) * (reverse ? -1 : 1);
        }
        if (!hits) continue;
        int first = -1;
        for (int e = 0; e < nd; ++e)
          if (delta[e] != 0) {
            first = e;
            break;
          }
        if (first != d || delta[d] >= 0) continue;  // not a true dependency led by axis d
        int64_t rhs = 0;
        for (int e = d + 1; e < nd; ++e) rhs += acoef[e] * delta[e];
        if (rhs >= 0) {
          const int64_t v = rhs / (-delta[d]) + 1;
          if (v > need) need = v;
        }
      }
    }
    acoef[d] = need;
  }
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.reverse = reverse;
  a.n_steps = 1;
  a.solved = -1;
  for (int d = 0; d < nd; ++d) {
    if (acoef[d] > 0x7fffffff) return false;
    a.a[d] = (int32_t)acoef[d];
    a.n_steps += acoef[d] * (g.out_shape[d] - 1);
    // solve for the longest dependent axis: fewest idle threads per hyperplane
    if (acoef[d] != 0 && (a.solved < 0 || g.out_shape[d] >= g.out_shape[a.solved])) a.solved = d;
  }
  a.n_free = 1;
  for (int d = 0; d < nd; ++d)
    if (d != a.solved) a.n_free *= g.out_shape[d];
  return true;
}

template <typename TIN, typename TC, typename IDX, bool UNIT>
int launch_wave_x(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev, void* out,
                  int reverse, cudaStream_t s) {
  static thread_local DevProgram<TC> dp;
  static thread_local WaveArgs a;
  make_dev_program<TC>(prog, g.ndim, dp);
  if (!plan_wavefront(g, prog, reverse, a)) return KX_ERR_UNSUPPORTED;
  if (prod(g.out_shape, g.ndim) <= 0) return KX_OK;
  const TIN* in_p = (const TIN*)in;
  TIN* out_p = (TIN*)out;
  const TC* coef_p = (const TC*)coef_dev;
  auto kern = scan_wavefront_kernel<TIN, TC, IDX, UNIT>;
  if (a.n_free <= kWaveSingleMax || a.n_steps == 1) {
    int threads = kWaveSingleThreads;
    int64_t blocks = 1;
    if (a.n_steps == 1) {  // no dependencies at all: a plain parallel map over the windows
      threads = kWaveThreads;
      blocks = (a.n_free + threads - 1) / threads;
      const int64_t cap = (int64_t)sm_count() * 8;
      if (blocks > cap) blocks = cap;
    } else if (a.n_free < kWaveSingleThreads) {
      threads = (int)((a.n_free + 31) / 32 * 32);
    }
    kern<<<(unsigned)blocks, threads, 0, s>>>(a, dp, in_p, out_p, coef_p);
    return after_launch("scan_wavefront_kernel");
  }
  int per_sm = 0;
  KX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWaveThreads, 0));
  if (per_sm < 1) return fail(KX_ERR_CUDA, "scan_wavefront_kernel does not fit an SM");
  int64_t blocks = (a.n_free + kWaveThreads - 1) / kWaveThreads;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  void* params[] = {(void*)&a, (void*)&dp, (void*)&in_p, (void*)&out_p, (void*)&coef_p};
  KX_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)blocks), dim3(kWaveThreads), params, 0, s));
  return after_launch("scan_wavefront_kernel");
}


// ---------------------------------------------------------------------------------------------
// Warp-skewed wavefront for 2-D (and 1-D) scans whose functions read cells of the row being
// written (Gauss-Seidel sweeps, recurrences along a row, README RK4): 'same'-style borders (every
// cell is a window centre), stride 1.  The hyperplane kernel above pays a block / grid barrier plus
// an L2 round trip per anti-diagonal (~3.7 us per step); here
//   * a WARP owns a band of 32 consecutive rows, lane r = row i0 + r, and lane r is `a0` columns
//     behind lane r-1 (the hyperplane skew), so at step s it computes column s - a0*r: the warp IS the
//     wavefront and the only synchronisation per step is __syncwarp;
//   * new values of the band travel through a per-warp shared-memory history (the last 32 columns of
//     each row), old values come from the input, so a step costs one window evaluation, not a
//     round trip to L2;
//   * bands are pipelined over warps / CTAs: band b starts as soon as the last row of band b-1 is
//     ahead of it (lane 31 publishes its progress in global memory every 16 columns, lane 0 of the next
//     band polls it; rows above the band are read from `out` with ld.global.cg).  Bands are handed
//     out by an atomic ticket, so a resident warp only ever waits for bands that already started.
// Total time ~ (nx + a0*ny) window evaluations of one warp instead of that many grid barriers.
constexpr int kHistW = 32;

struct WaveRowsArgs {
  int32_t ni, nj, k0, k1, c0, c1;      // shape, window, centre offsets
  int32_t a0;                           // column skew per row (hyperplane coefficient of axis 0)
  int32_t r_above;                      // largest column offset read on earlier rows
  int32_t reverse, nbands, two_d;
  int32_t pub_mask;                     // progress is published every pub_mask + 1 columns
  int32_t* ticket;                      // [1] next band, then [nbands] progress of each band's last row
};

template <typename TIN, typename TC>
__global__ void __launch_bounds__(128)
scan_wavefront_rows_kernel(const __grid_constant__ WaveRowsArgs a, const __grid_constant__ DevProgram<TC> p,
                           const TIN* __restrict__ in, TIN* out, const TC* __restrict__ coef_dev) {
  __shared__ TIN hist[4][32][kHistW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  volatile int32_t* progress = a.ticket + 1;
  const int sgn = a.reverse ? -1 : 1;
  for (;;) {
    int band = 0;
    if (lane == 0) band = atomicAdd(a.ticket, 1);
    band = __shfl_sync(0xffffffffu, band, 0);
    if (band >= a.nbands) break;
    const int il = band * 32 + lane;                   // logical (sweep-order) row of this lane
    const bool row_ok = il < a.ni;
    const int ia = a.reverse ? a.ni - 1 - il : il;     // actual row
    int avail = band == 0 ? a.nj : 0;                  // completed columns of the previous band's last row
    const int nsteps = a.nj + 31 * a.a0;
    for (int s = 0; s < nsteps; ++s) {
      if (band > 0) {
        const int need = min(s + a.r_above + 1, a.nj);
        if (need > avail) {                            // warp-uniform
          if (lane == 0) {
            int v = progress[band - 1];
            while (v < need) v = progress[band - 1];
            __threadfence();
            avail = v;
          }
          avail = __shfl_sync(0xffffffffu, avail, 0);
        }
      }
      const int jl = s - a.a0 * lane;                  // logical column
      if (row_ok && jl >= 0 && jl < a.nj) {
        const int ja = a.reverse ? a.nj - 1 - jl : jl;
        int64_t centre[KX_MAX_DIMS] = {0, 0, 0, 0};
        if (a.two_d) {
          centre[0] = ia;
          centre[1] = ja;
        } else {
          centre[0] = ja;
        }
        const int f = select_func(p, centre);
        const int kind = p.func[f].kind;
        const int t0 = p.func[f].tap_begin, t1 = p.func[f].tap_end;
        TC acc = (kind == KX_AFFINE) ? p.func[f].bias : TC(0);
        for (int t = t0; t < t1; ++t) {
          const int di = a.two_d ? (int)p.tap_off[t][0] - a.c0 : 0;
          const int dj = (int)p.tap_off[t][a.two_d ? 1 : 0] - a.c1;
          const int ci = ia + di, cj = ja + dj;        // actual cell
          const int ldi = sgn * di, ldj = sgn * dj;    // offset in sweep order
          const bool inside = ci >= 0 && ci < a.ni && cj >= 0 && cj < a.nj;
          TC v = TC(0);
          if (inside) {
            if (ldi < 0 || (ldi == 0 && ldj < 0)) {    // written earlier in the sweep: the new value
              const int r2 = lane + ldi;
              if (r2 >= 0) v = (TC)hist[warp][r2][(jl + ldj) & (kHistW - 1)];
              else v = (TC)__ldcg(out + (int64_t)ci * a.nj + cj);
            } else {                                   // not written yet (or the centre): the input
              v = (TC)in[(int64_t)ci * a.nj + cj];
            }
          }
          if (kind == KX_AFFINE) {
            const TC c = coef_dev ? coef_dev[t] : p.coef[t];
            acc = mad_wrap(c, v, acc);
          } else if (kind == KX_REDUCE_MAX) {
            acc = (t == t0) ? v : red_max(acc, v);
          } else if (kind == KX_REDUCE_MIN) {
            acc = (t == t0) ? v : red_min(acc, v);
          } else {
            acc = v;
          }
        }
        if (kind == KX_AFFINE && p.func[f].post_div != 0.f) acc = (TC)((float)acc / p.func[f].post_div);
        const TIN stored = (TIN)acc;                   // cast to the carried dtype on write (scan.py:50)
        hist[warp][lane][jl & (kHistW - 1)] = stored;
        out[(int64_t)ia * a.nj + ja] = stored;
      }
      const int j31 = s - 31 * a.a0;                   // column just completed by the band's last row
      const bool pub = j31 >= 0 && (((j31 & a.pub_mask) == a.pub_mask) || j31 == a.nj - 1);
      __syncwarp();
      if (pub && lane == 31) {   // the fence is cumulative over the warp's writes ordered by the barrier above
        __threadfence();
        progress[band] = j31 + 1;
      }
    }
  }
}

// Single-function AFFINE scans with at most kFastTaps taps (Gauss-Seidel sweeps, recurrences): the tap
// table lives in registers (source row in the history, element offset in the arrays, coefficient, new / old
// flag are loop invariants of a lane), the loop over the taps is fully unrolled so their loads overlap, and
// the input rows are prefetched into L1 sixteen columns ahead (a lone warp cannot hide an L2 miss).
constexpr int kFastTaps = 8;

template <typename TIN, typename TC>
__global__ void __launch_bounds__(128)
scan_wavefront_rows_affine_kernel(const __grid_constant__ WaveRowsArgs a, const __grid_constant__ DevProgram<TC> p,
                                  const TIN* __restrict__ in, TIN* out, const TC* __restrict__ coef_dev) {
  __shared__ TIN hist[4][32][kHistW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  volatile int32_t* progress = a.ticket + 1;
  const int sgn = a.reverse ? -1 : 1;
  const int ntaps = p.func[0].tap_end - p.func[0].tap_begin;
  const TC bias = p.func[0].bias;
  const unsigned hist0 = (unsigned)__cvta_generic_to_shared(&hist[warp][0][0]);
  for (;;) {
    int band = 0;
    if (lane == 0) band = atomicAdd(a.ticket, 1);
    band = __shfl_sync(0xffffffffu, band, 0);
    if (band >= a.nbands) break;
    const int il = band * 32 + lane;
    const bool row_ok = il < a.ni;
    const int ia = a.reverse ? a.ni - 1 - il : il;
    // per-lane tap table
    TC coef[kFastTaps];
    int ldj[kFastTaps], goff[kFastTaps], src[kFastTaps];   // src: 0 = zero (row outside), 1 = input, 2 = history, 3 = `out` (band above)
    unsigned hrow[kFastTaps];
#pragma unroll
    for (int t = 0; t < kFastTaps; ++t) {
      coef[t] = TC(0);
      ldj[t] = goff[t] = 0;
      src[t] = 0;
      hrow[t] = hist0;
      if (t < ntaps) {
        const int tt = p.func[0].tap_begin + t;
        const int di = a.two_d ? (int)p.tap_off[tt][0] - a.c0 : 0;
        const int dj = (int)p.tap_off[tt][a.two_d ? 1 : 0] - a.c1;
        const int ldi = sgn * di;
        coef[t] = coef_dev ? coef_dev[tt] : p.coef[tt];
        ldj[t] = sgn * dj;
        goff[t] = di * a.nj + dj;
        const int ci = ia + di;
        if (ci >= 0 && ci < a.ni) {
          if (ldi < 0 || (ldi == 0 && ldj[t] < 0)) {
            const int r2 = lane + ldi;
            src[t] = r2 >= 0 ? 2 : 3;
            hrow[t] = hist0 + (unsigned)((r2 >= 0 ? r2 : 0) * kHistW * 4);
          } else {
            src[t] = 1;
          }
        }
      }
    }
    int avail = band == 0 ? a.nj : 0;
    const int nsteps = a.nj + 31 * a.a0;
    const int64_t row_base = (int64_t)ia * a.nj;
    const unsigned hown = hist0 + (unsigned)(lane * kHistW * 4);
    for (int s = 0; s < nsteps; ++s) {
      if (band > 0) {
        const int need = min(s + a.r_above + 1, a.nj);
        if (need > avail) {
          if (lane == 0) {
            int v = progress[band - 1];
            while (v < need) v = progress[band - 1];
            __threadfence();
            avail = v;
          }
          avail = __shfl_sync(0xffffffffu, avail, 0);
        }
      }
      const int jl = s - a.a0 * lane;
      if (row_ok && jl >= 0 && jl < a.nj) {
        const int ja = a.reverse ? a.nj - 1 - jl : jl;
        const int64_t cell = row_base + ja;
        TC v[kFastTaps];
#pragma unroll
        for (int t = 0; t < kFastTaps; ++t) {           // all loads first: they are independent
          v[t] = TC(0);
          if (t < ntaps) {
            const int cl = jl + ldj[t];                 // logical column of the tap
            const bool inside = cl >= 0 && cl < a.nj;
            if (inside && src[t] == 2) {
              int raw;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(hrow[t] + (unsigned)((cl & (kHistW - 1)) * 4)));
              v[t] = (TC) reinterpret_cast<const TIN&>(raw);
            } else if (inside && src[t] == 3) {
              v[t] = (TC)__ldcg(out + cell + goff[t]);
            } else if (inside && src[t] == 1) {
              v[t] = (TC)in[cell + goff[t]];
              // touch the line 16 columns further along the sweep
              const int cp = cl + 16;
              if (cp < a.nj) asm volatile("prefetch.global.L1 [%0];" ::"l"(in + cell + goff[t] + sgn * 16));
            }
          }
        }
        TC acc = bias;
#pragma unroll
        for (int t = 0; t < kFastTaps; ++t)
          if (t < ntaps) acc = mad_wrap(coef[t], v[t], acc);
        if (p.func[0].post_div != 0.f) acc = (TC)((float)acc / p.func[0].post_div);
        const TIN stored = (TIN)acc;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(hown + (unsigned)((jl & (kHistW - 1)) * 4)),
                     "r"(reinterpret_cast<const int&>(stored))
                     : "memory");
        out[cell] = stored;
      }
      const int j31 = s - 31 * a.a0;
      const bool pub = j31 >= 0 && (((j31 & a.pub_mask) == a.pub_mask) || j31 == a.nj - 1);
      __syncwarp();
      if (pub && lane == 31) {
        __threadfence();
        progress[band] = j31 + 1;
      }
    }
  }
}

template <typename TIN, typename TC>
int try_wave_rows(const KxGeom& g, const KxProgram& prog, const WaveArgs& plan, const void* in, const void* coef_dev,
                  void* state, void* out, int reverse, cudaStream_t s) {
  if (getenv("KX_SCAN_NO_ROWS")) return KX_ERR_UNSUPPORTED;   // A/B switch: hyperplane kernel only
  if (g.ndim < 1 || g.ndim > 2 || !state) return KX_ERR_UNSUPPORTED;
  const int two_d = g.ndim == 2;
  for (int d = 0; d < g.ndim; ++d)
    if (g.stride[d] != 1 || g.lo[d] != (g.ksize[d] - 1) / 2 || g.out_shape[d] != g.in_shape[d]) return KX_ERR_UNSUPPORTED;
  const int dj_axis = g.ndim - 1;
  if (plan.a[dj_axis] != 1) return KX_ERR_UNSUPPORTED;        // no in-row dependency: rows are parallel
  const int64_t ni = two_d ? g.in_shape[0] : 1, nj = g.in_shape[dj_axis];
  if (ni * nj >= 0x40000000 || nj < 1) return KX_ERR_UNSUPPORTED;
  static thread_local WaveRowsArgs a;
  memset(&a, 0, sizeof(a));
  a.ni = (int)ni;
  a.nj = (int)nj;
  a.two_d = two_d;
  a.k0 = two_d ? g.ksize[0] : 1;
  a.k1 = g.ksize[dj_axis];
  a.c0 = (a.k0 - 1) / 2;
  a.c1 = (a.k1 - 1) / 2;
  a.a0 = two_d ? plan.a[0] : 0;
  a.reverse = reverse;
  a.nbands = (int)((ni + 31) / 32);
  int max_up = 0, max_back = 0, r_above = 0;
  const int sgn = reverse ? -1 : 1;
  for (int f = 0; f < prog.n_funcs; ++f)
    for (int t = prog.func[f].tap_begin; t < prog.func[f].tap_end; ++t) {
      const int ldi = sgn * (two_d ? prog.tap_off[t][0] - a.c0 : 0), ldj = sgn * (prog.tap_off[t][dj_axis] - a.c1);
      if (ldi < 0) {
        if (-ldi > max_up) max_up = -ldi;
        if (ldj > r_above) r_above = ldj;
      }
      if ((ldi < 0 || (ldi == 0 && ldj < 0)) && -ldj > max_back) max_back = -ldj;
    }
  if (a.a0 * max_up + max_back + 1 >= kHistW || max_up > 31) return KX_ERR_UNSUPPORTED;
  a.r_above = r_above;
  a.pub_mask = 15;
  if (const char* e = getenv("KX_WAVE_PUB")) a.pub_mask = atoi(e) > 0 ? atoi(e) - 1 : 15;   // tuning knob (power of two)
  a.ticket = (int32_t*)state;
  if ((int64_t)(a.nbands + 1) * 4 > scan_state_elems(g) * 4) return KX_ERR_UNSUPPORTED;
  KX_CUDA_CHECK(cudaMemsetAsync(state, 0, (size_t)(a.nbands + 1) * 4, s));
  static thread_local DevProgram<TC> dp;
  make_dev_program<TC>(prog, g.ndim, dp);
  int64_t blocks = (a.nbands + 3) / 4;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;                              // persistent: warps keep taking tickets
  const bool fast = prog.n_funcs == 1 && prog.func[0].kind == KX_AFFINE &&
                    prog.func[0].tap_end - prog.func[0].tap_begin <= kFastTaps && !getenv("KX_WAVE_GENERIC");
  if (fast)
    scan_wavefront_rows_affine_kernel<TIN, TC><<<(unsigned)blocks, 128, 0, s>>>(a, dp, (const TIN*)in, (TIN*)out,
                                                                               (const TC*)coef_dev);
  else
    scan_wavefront_rows_kernel<TIN, TC><<<(unsigned)blocks, 128, 0, s>>>(a, dp, (const TIN*)in, (TIN*)out,
                                                                        (const TC*)coef_dev);
  return after_launch("scan_wavefront_rows_kernel");
}

template <typename TIN, typename TC>
int launch_wave_t(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev, void* state,
                  void* out, int reverse, cudaStream_t s) {
  if (prod(g.out_shape, g.ndim) > 0) {
    static thread_local WaveArgs plan;
    if (plan_wavefront(g, prog, reverse, plan)) {
      const int rc = try_wave_rows<TIN, TC>(g, prog, plan, in, coef_dev, state, out, reverse, s);
      if (rc != KX_ERR_UNSUPPORTED) return rc;
    }
  }
  bool small = prod(g.out_shape, g.ndim) < 0x40000000 && prod(g.in_shape, g.ndim) < 0x40000000, unit = true;
  for (int d = 0; d < g.ndim; ++d) {
    if (g.stride[d] != 1) unit = false;
    if (g.in_shape[d] + g.lo[d] + g.hi[d] + g.ksize[d] >= 0x40000000) small = false;
  }
  if (small && unit) return launch_wave_x<TIN, TC, int32_t, true>(g, prog, in, coef_dev, out, reverse, s);
  if (small) return launch_wave_x<TIN, TC, int32_t, false>(g, prog, in, coef_dev, out, reverse, s);
  return launch_wave_x<TIN, TC, int64_t, false>(g, prog, in, coef_dev, out, reverse, s);
}

}  // namespace

int launch_scan_wavefront(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev, void* state,
                          void* out, int reverse, cudaStream_t s) {
  bool fl = false;  // coef_dev carries the compute type
  for (int f = 0; f < prog.n_funcs; ++f)
    if (prog.func[f].kind == KX_AFFINE && !prog.func[f].coef_is_int) fl = true;
  if (g.in_dtype == KX_F32) return launch_wave_t<float, float>(g, prog, in, coef_dev, state, out, reverse, s);
  if (g.in_dtype == KX_I32 && !fl) return launch_wave_t<int32_t, int32_t>(g, prog, in, coef_dev, state, out, reverse, s);
  if (g.in_dtype == KX_I32) return launch_wave_t<int32_t, float>(g, prog, in, coef_dev, state, out, reverse, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
