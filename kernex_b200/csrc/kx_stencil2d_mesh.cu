// Fast path: MULTI-FUNCTION kmap meshes (`F[slice] = fn` region assignment, kernel_interface.py:85-90)
// whose functions are all AFFINE, over the last two axes, stride 1: interior stencil + boundary
// conditions + source blocks in ONE pass.  The reference evaluates every function for every window
// under vmap and selects with lax.switch (kernex/_src/map.py:115-118); the generic kernel here
// re-derives the region of every window from the key table (0.04 of the HBM roofline).
//
// Same row-marching skeleton as kx_stencil2d.cu (a thread owns 4 output columns, keeps the last KY
// input rows in a register queue, 128-bit loads).  Region selection follows select_func
// (kernex/_src/utils.py:338-382: keys scanned from the LAST inserted, first match wins, no match =
// function 0); a key matches a window iff its row part (plus the leading batch axes) and its column part match, so
// both parts are 32-bit masks over the keys and per output m = row_mask & col_mask,
// function = m ? func_of_key[31 - clz(m)] : 0.
// Tiles (a band's rows x a warp's 128 columns) that ONE function covers are classified up front from any / all masks
// and run the plain stencil loop; only tiles cut by a region boundary select a function per output (see the kernel).
// Accumulation order: bias, then the window in row-major order (as kx_stencil2d.cu).
#include <algorithm>
#include <climits>
#include <type_traits>

#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kMT = 32;         // threads per block: ONE warp, so a slow (cut) tile holds no other warp's slot --
                                // with 128-thread blocks the 3 % boundary-column tiles cost 7 % (3 idle warps each)
constexpr int kMCols = 4;
constexpr int kMaxNF = KX_MAX_FUNCS;   // functions per mesh on this path
constexpr int kMaxTH = 128;     // rows per band
constexpr int kMaxPrio = 16;

// A region key in the form the kernel tests: per axis lo <= c < hi and c % step == 0 in 32-bit arithmetic
// (int -> [a, a+1), slice -> [start, end) with its step, wildcard / axis not in the key -> everything).
struct MeshKey {
  int lo[KX_MAX_DIMS], hi[KX_MAX_DIMS], step[KX_MAX_DIMS];
};

template <typename T>
struct MeshArgs {
  const T* in;
  T* out;
  int64_t in_batch, out_batch;
  int hin, win, hout, wout;
  int ly, lx, th;
  int n_funcs, n_keys;
  int split;                             // 1: warps whose tile ONE function covers take the plain stencil loop
  int ay, ax, ndim;                      // axes of the 2-D window inside the key items
  int ntx, nbands;                       // tile columns (one warp = one block each) and row bands
  int n_prio;                            // tile columns a vertical region edge runs through: dispatched first
  int prio[kMaxPrio];                    // (ascending)
  int n_prio_y;                          // row bands a horizontal region edge runs through: first among the bands
  int prio_y[kMaxPrio];                  // (ascending)
  uint32_t div_m, div_s;                 // magic division by (ntx - n_prio): q = (n * div_m) >> div_s for n < 2^31
  int64_t lead_shape[KX_MAX_DIMS];       // leading (batch) axes, for keys that constrain them
  uint32_t umask;                        // tap mask of the function with the most taps (tested as a uniform predicate)
  uint32_t mask[kMaxNF];                 // tap mask per function
  T bias[kMaxNF];
  T coef[kMaxNF][9];
  unsigned char key_func[KX_MAX_KEYS];
  MeshKey key[KX_MAX_KEYS];
};

// the coordinates fit 31 bits (checked by the launcher); unit steps skip the division
__device__ __forceinline__ bool item_match(const MeshKey& k, int d, int c) {
  const int st = k.step[d];
  return (c >= k.lo[d]) && (c < k.hi[d]) && (st == 1 || c % st == 0);   // C remainder: exact for negative c too
}

template <typename T>
__device__ __forceinline__ void ld4(const T* p, T (&d)[4]) {
  const int4 raw = __ldg(reinterpret_cast<const int4*>(p));
  d[0] = reinterpret_cast<const T&>(raw.x);
  d[1] = reinterpret_cast<const T&>(raw.y);
  d[2] = reinterpret_cast<const T&>(raw.z);
  d[3] = reinterpret_cast<const T&>(raw.w);
}

// Classification of a WARP's tile (the band's rows x the warp's 128 columns) by interval arithmetic, lane k <-> key k:
// returns the function that covers ALL windows of the tile, or -1 when a region boundary cuts it (or might: strided
// items are treated conservatively -- "matches somewhere" if the interval overlaps, "matches everywhere" never).
// A key matches a window iff every axis item matches, so with per-key any / all bits over the tile the last-inserted
// key that matches anywhere decides (select_func scans the keys from the last to the first): it covers the whole tile
// -> its function; only part of it -> mixed.  ~30 instructions per lane, no shared memory, no barrier.
template <typename T>
__device__ __forceinline__ int warp_tile_function(const MeshArgs<T>& a, int lane, int rc0, int rc1, int cc0, int cc1,
                                                  const int (&lead)[KX_MAX_DIMS]) {
  bool any = false, all = false;
  if (lane < a.n_keys) {
    const MeshKey& k = a.key[lane];
    any = all = true;
#pragma unroll 1
    for (int d = 0; d < a.ndim; ++d) {
      int c0, c1;                                   // inclusive coordinate range of the tile's window centres on axis d
      if (d == a.ax) c0 = cc0, c1 = cc1;
      else if (d == a.ay) c0 = rc0, c1 = rc1;
      else c0 = c1 = lead[d];
      const int lo = k.lo[d], hi = k.hi[d], st = k.step[d];
      if (c0 == c1) {                               // a single coordinate: exact
        const bool m = (c0 >= lo) && (c0 < hi) && (st == 1 || c0 % st == 0);
        any = any && m;
        all = all && m;
      } else {
        any = any && (c1 >= lo) && (c0 < hi);
        all = all && (c0 >= lo) && (c1 < hi) && (st == 1);
      }
    }
  }
  const uint32_t any_m = __ballot_sync(0xffffffffu, any), all_m = __ballot_sync(0xffffffffu, all);
  if (any_m == 0u) return 0;                        // no key matches anywhere: function 0 (utils.py:379)
  const int k = 31 - __clz(any_m);
  return (all_m >> k) & 1u ? (int)a.key_func[k] : -1;
}

// ONE launch, one WARP per block and per tile (no block barrier, no shared memory); one row-marching loop (`march`):
//   * a tile ONE function covers -- the interior of a PDE mesh, nearly all tiles -- runs the plain stencil loop of
//     kx_stencil2d.cu (U rows in flight); for functions 0..3 the function index is a template constant, so bias,
//     coefficients and tap mask are constant-bank operands of the FMAs;
//   * a tile a region edge runs through keeps the column part of every key per owned column (32-bit masks) and takes
//     the row part of all keys per row with one ballot (lane k <-> key k).  While that ballot does not change -- every
//     row except the one or two a horizontal edge runs through -- the tile's MAIN function (coefficients in registers)
//     and the per-column functions stay as they are: a row costs the plain loop's FMAs plus, on the few lanes whose
//     column belongs to another function, a second evaluation with coefficients from the parameter bank;
//   * tiles wholly inside the array (INNER) use one row pointer advanced by the pitch and unconditional 128-bit loads /
//     stores; tiles on the rim predicate every vector.
// Both accumulate bias first, then the window in row-major order.
// How it got here, 3-function 8192^2 mesh (profiles/r2_mesh_*, r2e_mesh_probe.txt; the single-function kernel on that
// grid: 87 us = 0.94 of the roofline): round-1 kernel (block-wide mask prologue, 78 registers) 124 us = 0.66;
// per-warp classification + a separate one-row-in-flight loop for cut tiles 106-112 us = 0.73-0.77 (two launches,
// helper streams, programmatic dependent launch: all slower).  What the remaining gap was, measured piece by piece:
//   - per row load the compiler re-derived column index, batch offset and bounds predicates (48 instructions per row
//     load against 20 FMAs, 1.05 instructions per output vs 0.91 in stencil2d_kernel) -> INNER path; with it the kernel
//     wants 94 registers: capped at 64 (32 warps / SM; the few spills land in the cut-tile paths) 97 us, at 72 99 us,
//     uncapped 112 us; two rows in flight 105 us;
//   - 128-thread blocks: a slow cut warp keeps three finished warps' slots -> one-warp blocks: a mesh WITHOUT cut
//     tiles 92.7 -> 88.9 us;
//   - the right boundary tile of the last band is the LAST block of an x-fastest grid and a cut tile takes 2-3x as
//     long as a covered one: boundary columns cost +7.5 us of pure tail (boundary rows: +2) -> linear grid, the tile
//     columns / bands holding a region edge first: 97 -> 93 us = 0.88 (columns only: 91 us = 0.90, no cut tile: 0.92).
// VW = 4: rows 16-byte aligned (SHIFT = (-lx) mod 4); VW = 1: scalar loads (SHIFT = 0); U = rows fetched back to back
template <typename T, int KY, int KX, int SHIFT, int VW, int U, int MB>
__global__ void __launch_bounds__(kMT, MB * (128 / kMT))
stencil2d_mesh_kernel(const __grid_constant__ MeshArgs<T> a) {
  constexpr int NV = (SHIFT + KX + kMCols - 1 + VW - 1) / VW;
  constexpr int NE = NV * VW;
  const int tid = threadIdx.x, lane = tid & 31;
  // Block order.  A tile a VERTICAL region edge runs through (the boundary columns of a PDE mesh, the rim of a source
  // block) evaluates a second function on some lanes and takes 2-3x as long as a covered tile; in the natural x-fastest
  // order the right boundary tile of the last band is the LAST block of the grid and the kernel ends with its latency
  // (measured, 8192^2: boundary columns +7.5 us on an 89 us kernel, boundary rows +0).  So the grid is linear: first the
  // tiles of those columns for every band, then everything else band by band.
  int bx, by;
  {
    const uint32_t l = blockIdx.x, first = (uint32_t)a.nbands * (uint32_t)a.n_prio;
    if (l < first) {
      by = (int)(l / (uint32_t)a.n_prio);
      bx = a.prio[l - (uint32_t)by * a.n_prio];
    } else {
      const uint32_t l2 = l - first, w = (uint32_t)(a.ntx - a.n_prio);
      by = (int)(((uint64_t)l2 * a.div_m) >> a.div_s);
      bx = (int)(l2 - (uint32_t)by * w);
#pragma unroll 1
      for (int k = 0; k < a.n_prio; ++k)
        if (a.prio[k] <= bx) ++bx;
    }
    // bands a horizontal region edge runs through (every tile selects per row there; +2 us as the last band): first
    if (by < a.n_prio_y) {
      by = a.prio_y[by];
    } else if (a.n_prio_y > 0) {
      by -= a.n_prio_y;
#pragma unroll 1
      for (int k = 0; k < a.n_prio_y; ++k)
        if (a.prio_y[k] <= by) ++by;
    }
  }
  const int j0 = (bx * kMT + tid) * kMCols;
  const int jw = (bx * kMT + (tid & ~31)) * kMCols;                   // the warp's first output column
  if (jw >= a.wout) return;                                           // whole warp outside
  const int r0 = by * a.th, r1 = min(r0 + a.th, a.hout);
  const T* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int cy = (KY - 1) / 2, cx = (KX - 1) / 2;

  int lead[KX_MAX_DIMS] = {0, 0, 0, 0};
  {
    int64_t b = blockIdx.z;
    for (int d = a.ndim - 3; d >= 0; --d) {
      lead[d] = (int)(b % a.lead_shape[d]);
      b /= a.lead_shape[d];
    }
  }
  int fn = -1;
  if (a.split)
    fn = warp_tile_function<T>(a, lane, r0 - a.ly + cy, r1 - 1 - a.ly + cy, jw - a.lx + cx,
                               min(jw + 32 * kMCols, a.wout) - 1 - a.lx + cx, lead);
  const int xa = j0 - a.lx - SHIFT;
  const int n_valid = min(kMCols, a.wout - j0);                       // <= 0 for lanes past the last column

  auto load_row = [&](int iy, T(&dst)[NE]) {
    const bool row_ok = (iy >= 0) && (iy < a.hin) && n_valid > 0;
    const T* rp = in + (int64_t)iy * a.win;
    if (VW == 4) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = xa + 4 * v;
        T t4[4] = {T(0), T(0), T(0), T(0)};
        if (row_ok && c >= 0 && c + 3 < a.win) ld4(rp + c, t4);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[4 * v + e] = t4[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int c = xa + e;
        dst[e] = (row_ok && c >= 0 && c < a.win) ? __ldg(rp + c) : T(0);
      }
    }
  };
  auto store_row = [&](int r, const T(&acc)[kMCols]) {
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
      for (int i = 0; i < kMCols; ++i)
        if (i < n_valid) o[i] = acc[i];
    }
  };

  // ---- one row-marching loop for both kinds of tile ----
  // `cut` tiles carry the column part of every key per owned column (col_mask) and lane k's copy of key k's row part;
  // the row part of all keys is ONE ballot per row (rm).  While rm does not change -- every row of a tile except the
  // one or two a horizontal region edge runs through -- the function of each owned column (fi_pack) and the tile's
  // MAIN function (the one lane 0's second column uses; its coefficients sit in registers like the covering function
  // of a uniform tile) stay as they are, so a row costs the uniform loop's FMAs plus, on the few lanes whose column
  // belongs to another function (`slow` bits: a boundary column, the rim of a source block), a second evaluation
  // with that function's coefficients from the parameter bank.
  const bool cut = fn < 0;
  uint32_t col_mask[kMCols] = {0u, 0u, 0u, 0u};
  bool lead_ok = false;
  int ylo = 0, yhi = 0, yst = 1;
  if (cut) {
#pragma unroll
    for (int i = 0; i < kMCols; ++i) {
      const int ccol = (j0 + i) - a.lx + cx;
      uint32_t m = 0;
#pragma unroll 1
      for (int k = 0; k < a.n_keys; ++k)
        if (item_match(a.key[k], a.ax, ccol)) m |= 1u << k;
      col_mask[i] = m;
    }
    lead_ok = lane < a.n_keys;
    if (lead_ok) {
      const MeshKey& k = a.key[lane];
#pragma unroll 1
      for (int d = 0; d < a.ndim; ++d)
        if (d != a.ax && d != a.ay) lead_ok = lead_ok && item_match(k, d, lead[d]);
      if (a.ay >= 0) ylo = k.lo[a.ay], yhi = k.hi[a.ay], yst = k.step[a.ay];
    }
  }
  int fmain = cut ? -1 : fn;
  uint32_t um = 0u, slow = 0u, fi_pack = 0u, rm_prev = 0u;
  bool have_rm = false;
  T cb = T(0);
  T cc[KY * KX];
#pragma unroll
  for (int t = 0; t < KY * KX; ++t) cc[t] = T(0);
  auto set_main = [&](int f) {                       // warp-uniform
    um = a.mask[f];
    cb = a.bias[f];
#pragma unroll
    for (int t = 0; t < KY * KX; ++t) cc[t] = a.coef[f][t];
  };

  // INNER: every input row and every aligned vector the warp's tile reads exists and the output rows are 16-byte
  // aligned (all but the tiles on the rim of the array): one row pointer advanced by the row pitch, unconditional
  // 128-bit loads and stores.  (With the bounds tests per vector the compiler re-derived column index, batch offset
  // and predicates for EVERY row load under the 64-register cap: 48 instructions per row load against 20 FMAs.)
  // FSEL >= 0: the covering function's index is a compile-time constant, so its coefficients, bias and tap mask are
  // constant-bank operands of the FMAs (no registers); FSEL < 0: they sit in cc / cb / um (set_main)
  auto march = [&](auto inner_c, auto cut_c, auto fsel_c) {
    constexpr bool INNER = decltype(inner_c)::value;
    constexpr bool CUT = decltype(cut_c)::value;
    constexpr int FSEL = decltype(fsel_c)::value;
    constexpr int UU = CUT ? (U < 2 ? U : 2) : U;     // rows in flight: the cut tiles carry the region state instead
    const T* lp = in + ((int64_t)(r0 - a.ly) * a.win + xa);          // next input row, this thread's first vector
    T* op = out + ((int64_t)r0 * a.wout + j0);                        // next output row
    auto next_row = [&](int iy, T(&dst)[NE]) {
      if constexpr (INNER) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          T t4[4];
          ld4(lp + 4 * v, t4);
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[4 * v + e] = t4[e];
        }
      } else {
        load_row(iy, dst);
      }
      lp += a.win;
    };
    T q[KY - 1 + UU][NE];
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky) next_row(r0 - a.ly + ky, q[ky]);
    for (int r = r0; r < r1; r += UU) {
#pragma unroll
      for (int u = 0; u < UU; ++u)
        if (r + u < r1) next_row(r + u - a.ly + KY - 1, q[KY - 1 + u]);
#pragma unroll
      for (int u = 0; u < UU; ++u) {
        if (r + u < r1) {
          if constexpr (CUT) {
            const int rc = r + u - a.ly + cy;
            const bool rmatch = lead_ok && (a.ay < 0 || ((rc >= ylo) && (rc < yhi) && (yst == 1 || rc % yst == 0)));
            const uint32_t rm = __ballot_sync(0xffffffffu, rmatch);
            if (!have_rm || rm != rm_prev) {           // warp-uniform: a horizontal region edge (or the first row)
              have_rm = true;
              rm_prev = rm;
              int f_[kMCols];
#pragma unroll
              for (int i = 0; i < kMCols; ++i) {
                const uint32_t m = rm & col_mask[i];
                f_[i] = m ? (int)a.key_func[31 - __clz(m)] : 0;
              }
              const int fm = __shfl_sync(0xffffffffu, f_[1], 0);
              if (fm != fmain) {
                fmain = fm;
                set_main(fm);
              }
              slow = 0u, fi_pack = 0u;
#pragma unroll
              for (int i = 0; i < kMCols; ++i) {
                if (f_[i] != fmain) slow |= 1u << i;
                fi_pack |= (uint32_t)f_[i] << (8 * i);
              }
            }
          }
          T acc[kMCols];
          const uint32_t umx = FSEL >= 0 ? a.mask[FSEL >= 0 ? FSEL : 0] : um;
#pragma unroll
          for (int i = 0; i < kMCols; ++i) acc[i] = FSEL >= 0 ? a.bias[FSEL >= 0 ? FSEL : 0] : cb;
#pragma unroll
          for (int ky = 0; ky < KY; ++ky)
#pragma unroll
            for (int kx = 0; kx < KX; ++kx)
              if (umx & (1u << (ky * KX + kx))) {
                const T c = FSEL >= 0 ? a.coef[FSEL >= 0 ? FSEL : 0][ky * KX + kx] : cc[ky * KX + kx];
#pragma unroll
                for (int i = 0; i < kMCols; ++i) acc[i] = mad_wrap(c, q[u + ky][SHIFT + kx + i], acc[i]);
              }
          if (CUT && slow) {                           // only the lanes a vertical region edge runs through
#pragma unroll
            for (int i = 0; i < kMCols; ++i)
              if (slow & (1u << i)) {
                const int fi = (int)((fi_pack >> (8 * i)) & 255u);
                const uint32_t tm = a.mask[fi];
                T v = a.bias[fi];
#pragma unroll
                for (int ky = 0; ky < KY; ++ky)
#pragma unroll
                  for (int kx = 0; kx < KX; ++kx)
                    if (tm & (1u << (ky * KX + kx))) v = mad_wrap(a.coef[fi][ky * KX + kx], q[u + ky][SHIFT + kx + i], v);
                acc[i] = v;
              }
          }
          if constexpr (INNER) {
            int4 raw;
            raw.x = reinterpret_cast<const int&>(acc[0]);
            raw.y = reinterpret_cast<const int&>(acc[1]);
            raw.z = reinterpret_cast<const int&>(acc[2]);
            raw.w = reinterpret_cast<const int&>(acc[3]);
            __stcs(reinterpret_cast<int4*>(op), raw);
            op += a.wout;
          } else {
            store_row(r + u, acc);
          }
        }
      }
#pragma unroll
      for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
        for (int e = 0; e < NE; ++e) q[ky][e] = q[UU + ky][e];
      }
    }
  };
  // warp-uniform
  const int xw = jw - a.lx - SHIFT;                                   // the warp's first loaded column
  const bool inner = VW == 4 && (r0 - a.ly >= 0) && (r1 - 1 - a.ly + KY - 1 < a.hin) && xw >= 0 &&
                     (xw + 31 * kMCols + NE <= a.win) && (jw + 32 * kMCols <= a.wout) && (a.wout % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  using No = std::false_type;
  using Yes = std::true_type;
  using Dyn = std::integral_constant<int, -1>;
  if (!cut) {
    if (!inner) {
      set_main(fn);
      march(No{}, No{}, Dyn{});
    } else if (fn == 0) march(Yes{}, No{}, std::integral_constant<int, 0>{});
    else if (fn == 1) march(Yes{}, No{}, std::integral_constant<int, 1>{});
    else if (fn == 2 && kMaxNF > 2) march(Yes{}, No{}, std::integral_constant<int, (kMaxNF > 2 ? 2 : 0)>{});
    else if (fn == 3 && kMaxNF > 3) march(Yes{}, No{}, std::integral_constant<int, (kMaxNF > 3 ? 3 : 0)>{});
    else {
      set_main(fn);
      march(Yes{}, No{}, Dyn{});
    }
  } else {
    if (inner) march(Yes{}, Yes{}, Dyn{});
    else march(No{}, Yes{}, Dyn{});
  }
}

template <typename T, int KY, int KX, int U, int MB>
int launch_mesh_v(const MeshArgs<T>& a, dim3 grid, bool vec, cudaStream_t s) {
  const int shift = vec ? (((-a.lx) % 4) + 4) % 4 : 0;
  if (!vec) stencil2d_mesh_kernel<T, KY, KX, 0, 1, U, MB><<<grid, kMT, 0, s>>>(a);
  else if (shift == 0) stencil2d_mesh_kernel<T, KY, KX, 0, 4, U, MB><<<grid, kMT, 0, s>>>(a);
  else if (shift == 1) stencil2d_mesh_kernel<T, KY, KX, 1, 4, U, MB><<<grid, kMT, 0, s>>>(a);
  else if (shift == 2) stencil2d_mesh_kernel<T, KY, KX, 2, 4, U, MB><<<grid, kMT, 0, s>>>(a);
  else stencil2d_mesh_kernel<T, KY, KX, 3, 4, U, MB><<<grid, kMT, 0, s>>>(a);
  return after_launch("stencil2d_mesh_kernel");
}

// Measured on the 3-function 8192^2 mesh (profiles/r2e_mesh_probe.txt): 4 rows in flight at 64 registers (32 one-warp
// blocks per SM; a handful of spills in the cut-tile paths only) 97 us, at 72 registers 99, uncapped (94 registers) 112;
// 2 rows in flight 105-126.  KX_MESH_U=2 keeps the two-row variant for A/B.
template <typename T, int KY, int KX>
int launch_mesh(const MeshArgs<T>& a, dim3 grid, bool vec, cudaStream_t s) {
  const char* e = getenv("KX_MESH_U");
  if (e && atoi(e) == 2) return launch_mesh_v<T, KY, KX, 2, 8>(a, grid, vec, s);
  return launch_mesh_v<T, KY, KX, 4, 8>(a, grid, vec, s);
}

template <typename T>
int run_mesh(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim, ax = nd - 1, ay = nd - 2;
  const int KY = ay >= 0 ? g.ksize[ay] : 1, KX = g.ksize[ax];
  static thread_local MeshArgs<T> a;
  memset(&a, 0, sizeof(a));
  int64_t lead = 1;
  for (int d = 0; d < nd - 2; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    a.lead_shape[d] = g.in_shape[d];
    lead *= g.in_shape[d];
  }
  const int64_t batch = m.batch * lead;
  if (batch > 65535 || m.batch != 1) return KX_ERR_UNSUPPORTED;
  for (int f = 0; f < p.n_funcs; ++f) {
    const KxFunc& fn = p.func[f];
    a.bias[f] = (T)fn.bias;
    for (int t = fn.tap_begin; t < fn.tap_end; ++t) {
      for (int d = 0; d < nd - 2; ++d)
        if (p.tap_off[t][d] != 0) return KX_ERR_UNSUPPORTED;
      const int slot = (ay >= 0 ? p.tap_off[t][ay] : 0) * KX + p.tap_off[t][ax];
      if (a.mask[f] & (1u << slot)) return KX_ERR_UNSUPPORTED;   // duplicate tap: generic kernel
      a.mask[f] |= 1u << slot;
      a.coef[f][slot] = (T)p.coef[t];
    }
  }
  a.n_funcs = p.n_funcs;
  a.n_keys = p.n_keys;
  // the uniform-predicate path serves the function with the most taps (ties: the first); its mask is `umask`
  a.umask = a.mask[0];
  for (int f = 1; f < p.n_funcs; ++f)
    if (__builtin_popcount(a.mask[f]) > __builtin_popcount(a.umask)) a.umask = a.mask[f];
  for (int k = 0; k < p.n_keys; ++k) {
    a.key_func[k] = (unsigned char)p.key[k].func;
    for (int d = 0; d < KX_MAX_DIMS; ++d) {
      int64_t lo = INT32_MIN, hi = INT32_MAX, st = 1;
      if (d < p.key[k].n_items) {
        const KxKeyItem& it = p.key[k].item[d];
        if (it.kind == KX_KEY_EQ) {
          lo = it.a;
          hi = it.a + 1;
        } else if (it.kind == KX_KEY_RANGE || it.kind == KX_KEY_RANGE_STEP) {
          lo = it.a;
          hi = it.b;
          if (it.kind == KX_KEY_RANGE_STEP) st = it.step < 0 ? -(int64_t)it.step : it.step;
          if (st == 0) return KX_ERR_UNSUPPORTED;   // left to the generic kernel
        }
      }
      lo = std::max<int64_t>(lo, INT32_MIN);
      hi = std::min<int64_t>(hi, INT32_MAX);
      if (lo >= hi) lo = 1, hi = 0;                 // empty range
      a.key[k].lo[d] = (int)lo;
      a.key[k].hi[d] = (int)hi;
      a.key[k].step[d] = (int)st;
    }
  }
  a.ay = ay;
  a.ax = ax;
  a.ndim = nd;
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.hin = ay >= 0 ? (int)g.in_shape[ay] : 1;
  a.win = (int)g.in_shape[ax];
  a.hout = ay >= 0 ? (int)g.out_shape[ay] : 1;
  a.wout = (int)g.out_shape[ax];
  a.ly = ay >= 0 ? g.lo[ay] : 0;
  a.lx = g.lo[ax];
  a.in_batch = (int64_t)a.hin * a.win;
  a.out_batch = (int64_t)a.hout * a.wout;
  // band height, measured on the 3-function 8192^2 mesh with the per-warp classification (U = 4): 6 rows 0.63 of the
  // roofline, 8 0.69, 10 0.70, 12 0.72, 14 0.72, 16 0.71, 24 0.67, 32 0.63, 48 0.58 (round 1, block-wide mask
  // prologue: 32 rows were best)
  a.th = a.hout >= 64 ? 12 : 8;
  if (const char* e = getenv("KX_MESH_TH")) a.th = atoi(e) > 0 ? atoi(e) : a.th;
  if (a.th > kMaxTH) return KX_ERR_UNSUPPORTED;
  const bool vec = (a.win % 4 == 0) && ((reinterpret_cast<uintptr_t>(m.in) & 15) == 0) && (a.in_batch % 4 == 0);
  dim3 grid;
  {
    // tile columns holding a column edge of some key (window-centre coordinate c = j - lx + cx; edges at lo and hi)
    const int tw = kMT * kMCols, cxx = (KX - 1) / 2;
    a.ntx = (a.wout + tw - 1) / tw;
    a.nbands = (a.hout + a.th - 1) / a.th;
    if ((int64_t)a.ntx * a.nbands > 0x7fffffffll) return KX_ERR_UNSUPPORTED;
    int cand[2 * KX_MAX_KEYS], nc = 0;
    for (int k = 0; k < a.n_keys; ++k)
      for (int e = 0; e < 2; ++e) {
        const int64_t c = e == 0 ? a.key[k].lo[ax] : a.key[k].hi[ax];   // first column inside / first column outside
        const int64_t j = c + a.lx - cxx;                                 // output column of that centre
        if (j <= 0 || j >= a.wout) continue;                              // the edge is outside the output
        if (j % tw == 0) continue;                                        // the edge coincides with a tile edge
        cand[nc++] = (int)(j / tw);
      }
    std::sort(cand, cand + nc);
    nc = (int)(std::unique(cand, cand + nc) - cand);
    a.n_prio = 0;
    if (nc <= kMaxPrio && nc < a.ntx && !getenv("KX_MESH_NO_PRIO"))
      for (int i = 0; i < nc; ++i) a.prio[a.n_prio++] = cand[i];
    nc = 0;
    const int cyy = (KY - 1) / 2;
    for (int k = 0; k < a.n_keys && ay >= 0; ++k)
      for (int e = 0; e < 2; ++e) {
        const int64_t c = e == 0 ? a.key[k].lo[ay] : a.key[k].hi[ay];   // first row inside / first row outside
        const int64_t r = c + a.ly - cyy;                                 // output row of that centre
        if (r <= 0 || r >= a.hout || r % a.th == 0) continue;             // outside the output / on a band edge
        cand[nc++] = (int)(r / a.th);
      }
    std::sort(cand, cand + nc);
    nc = (int)(std::unique(cand, cand + nc) - cand);
    a.n_prio_y = 0;
    if (nc <= kMaxPrio && nc < a.nbands && !getenv("KX_MESH_NO_PRIO"))
      for (int i = 0; i < nc; ++i) a.prio_y[a.n_prio_y++] = cand[i];
    const uint32_t d = (uint32_t)(a.ntx - a.n_prio);                      // >= 1
    uint32_t sh = 0;
    while ((1ull << sh) < d) ++sh;
    a.div_m = (uint32_t)(((1ull << (31 + sh)) / d) + 1);
    a.div_s = 31 + sh;
    grid = dim3((unsigned)(a.ntx * a.nbands), 1, (unsigned)batch);
  }
  a.split = getenv("KX_MESH_NO_UNIFORM") ? 0 : 1;   // 0: every tile on the general loop (cross-check)
  if (KY == 3 && KX == 3) return launch_mesh<T, 3, 3>(a, grid, vec, s);
  if (KY == 1 && KX == 3) return launch_mesh<T, 1, 3>(a, grid, vec, s);
  if (KY == 1 && KX == 1) return launch_mesh<T, 1, 1>(a, grid, vec, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace

int try_stencil2d_mesh(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (p.n_funcs < 2 || p.n_funcs > kMaxNF || m.fn_elems != 1 || m.coef_dev || m.use_box) return KX_ERR_UNSUPPORTED;
  if (getenv("KX_NO_MESH")) return KX_ERR_UNSUPPORTED;   // cross-check switch: generic kernel
  if (g.in_dtype != g.out_dtype) return KX_ERR_UNSUPPORTED;
  bool all_int = true;
  for (int f = 0; f < p.n_funcs; ++f) {
    if (p.func[f].kind != KX_AFFINE || p.func[f].post_div != 0.0) return KX_ERR_UNSUPPORTED;
    if (!p.func[f].coef_is_int) all_int = false;
  }
  const int nd = g.ndim;
  if (g.stride[nd - 1] != 1 || (nd >= 2 && g.stride[nd - 2] != 1)) return KX_ERR_UNSUPPORTED;
  for (int d = 0; d < nd; ++d)
    if (g.in_shape[d] > 0x3fffffff || g.out_shape[d] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype == KX_F32) return run_mesh<float>(m, p, s);
  if (g.in_dtype == KX_I32 && all_int) return run_mesh<int32_t>(m, p, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
