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

#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kMT = 128;        // threads per block
constexpr int kMCols = 4;
constexpr int kMaxNF = KX_MAX_FUNCS;   // functions per mesh on this path
constexpr int kMaxTH = 128;     // rows per band

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

// ONE launch, every warp on its own (no block barrier, no shared memory), one tile per warp:
//   * a warp whose tile ONE function covers -- the interior of a PDE mesh, nearly all tiles -- runs the plain stencil
//     loop of kx_stencil2d.cu (U rows in flight) with that function's coefficients in registers and a warp-uniform tap
//     mask;
//   * a warp whose tile is cut by a region boundary (with KX_MESH_NO_UNIFORM=1: every warp) selects the function per
//     output: row part of every key per row by one ballot (lane k <-> key k), column part per owned column in a 32-bit
//     mask, function = key_func[31 - clz(row & col)] or 0, coefficients from the parameter bank.  This loop keeps ONE
//     row in flight and a KY-row queue, so it fits in the registers of the fast loop; it is latency-bound, but its few
//     warps (~4 % of an 8192^2 mesh: the boundary columns and rows, the rim of a source block) run under the others.
// Both accumulate bias first, then the window in row-major order.
// How it got here, 3-function 8192^2 mesh (profiles/r2_mesh_*; the single-function kernel on that grid: 87 us = 0.94 of
// the roofline): round-1 kernel (block-wide mask prologue worth ~18 % of its instructions, per-thread coefficient
// caching, 78 registers) 124 us = 0.66; uniform / cut tiles as two launches back to back 112-130 us (the uniform pass
// alone reaches the single-function kernel's speed, the cut tiles need 28-45 us of pure latency whatever the grid);
// overlapped through a helper stream 112 us (~10 us per event hop); as programmatic dependent launches 128 us (the
// dependent cannot start before EVERY block of the first launch has begun); this kernel 106 us = 0.77 (dispatching the
// column blocks that hold a boundary warp first: 110 us).
// VW = 4: rows 16-byte aligned (SHIFT = (-lx) mod 4); VW = 1: scalar loads (SHIFT = 0); U = rows fetched back to back
template <typename T, int KY, int KX, int SHIFT, int VW, int U>
__global__ void __launch_bounds__(kMT)
stencil2d_mesh_kernel(const __grid_constant__ MeshArgs<T> a) {
  constexpr int NV = (SHIFT + KX + kMCols - 1 + VW - 1) / VW;
  constexpr int NE = NV * VW;
  const int tid = threadIdx.x, lane = tid & 31;
  const int bx = blockIdx.x, by = blockIdx.y;
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

  if (fn >= 0) {
    // ---- the tile belongs to ONE function ----
    T q[KY - 1 + U][NE];
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky) load_row(r0 - a.ly + ky, q[ky]);
    const uint32_t um = a.mask[fn];                  // warp-uniform
    const T cb = a.bias[fn];
    T cc[KY * KX];
#pragma unroll
    for (int t = 0; t < KY * KX; ++t) cc[t] = a.coef[fn][t];
    for (int r = r0; r < r1; r += U) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r + u < r1) load_row(r + u - a.ly + KY - 1, q[KY - 1 + u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + u < r1) {
          T acc[kMCols];
#pragma unroll
          for (int i = 0; i < kMCols; ++i) acc[i] = cb;
#pragma unroll
          for (int ky = 0; ky < KY; ++ky)
#pragma unroll
            for (int kx = 0; kx < KX; ++kx)
              if (um & (1u << (ky * KX + kx))) {
#pragma unroll
                for (int i = 0; i < kMCols; ++i) acc[i] = mad_wrap(cc[ky * KX + kx], q[u + ky][SHIFT + kx + i], acc[i]);
              }
          store_row(r + u, acc);
        }
      }
#pragma unroll
      for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
        for (int e = 0; e < NE; ++e) q[ky][e] = q[U + ky][e];
      }
    }
    return;
  }

  // ---- a region boundary cuts the tile ----
  // column part of every key, per owned column; lane k's copy of key k's other axes for the per-row ballot
  uint32_t col_mask[kMCols];
#pragma unroll
  for (int i = 0; i < kMCols; ++i) {
    const int ccol = (j0 + i) - a.lx + cx;
    uint32_t m = 0;
#pragma unroll 1
    for (int k = 0; k < a.n_keys; ++k)
      if (item_match(a.key[k], a.ax, ccol)) m |= 1u << k;
    col_mask[i] = m;
  }
  bool lead_ok = lane < a.n_keys;
  int ylo = 0, yhi = 0, yst = 1;
  if (lead_ok) {
    const MeshKey& k = a.key[lane];
#pragma unroll 1
    for (int d = 0; d < a.ndim; ++d)
      if (d != a.ax && d != a.ay) lead_ok = lead_ok && item_match(k, d, lead[d]);
    if (a.ay >= 0) ylo = k.lo[a.ay], yhi = k.hi[a.ay], yst = k.step[a.ay];
  }
  T w[KY][NE];                                       // the window rows of the current output row
#pragma unroll
  for (int ky = 0; ky < KY - 1; ++ky) load_row(r0 - a.ly + ky, w[ky]);
#pragma unroll 1
  for (int r = r0; r < r1; ++r) {
    load_row(r - a.ly + KY - 1, w[KY - 1]);
    const int rc = r - a.ly + cy;
    const bool rmatch = lead_ok && (a.ay < 0 || ((rc >= ylo) && (rc < yhi) && (yst == 1 || rc % yst == 0)));
    const uint32_t rm = __ballot_sync(0xffffffffu, rmatch);
    T acc[kMCols];
#pragma unroll
    for (int i = 0; i < kMCols; ++i) {
      const uint32_t m = rm & col_mask[i];
      const int fi = m ? (int)a.key_func[31 - __clz(m)] : 0;
      const uint32_t tm = a.mask[fi];
      T v = a.bias[fi];
#pragma unroll
      for (int ky = 0; ky < KY; ++ky)
#pragma unroll
        for (int kx = 0; kx < KX; ++kx)
          if (tm & (1u << (ky * KX + kx))) v = mad_wrap(a.coef[fi][ky * KX + kx], w[ky][SHIFT + kx + i], v);
      acc[i] = v;
    }
    store_row(r, acc);
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
      for (int e = 0; e < NE; ++e) w[ky][e] = w[ky + 1][e];
    }
  }
}

template <typename T, int KY, int KX, int U>
int launch_mesh_v(const MeshArgs<T>& a, dim3 grid, bool vec, cudaStream_t s) {
  const int shift = vec ? (((-a.lx) % 4) + 4) % 4 : 0;
  if (!vec) stencil2d_mesh_kernel<T, KY, KX, 0, 1, U><<<grid, kMT, 0, s>>>(a);
  else if (shift == 0) stencil2d_mesh_kernel<T, KY, KX, 0, 4, U><<<grid, kMT, 0, s>>>(a);
  else if (shift == 1) stencil2d_mesh_kernel<T, KY, KX, 1, 4, U><<<grid, kMT, 0, s>>>(a);
  else if (shift == 2) stencil2d_mesh_kernel<T, KY, KX, 2, 4, U><<<grid, kMT, 0, s>>>(a);
  else stencil2d_mesh_kernel<T, KY, KX, 3, 4, U><<<grid, kMT, 0, s>>>(a);
  return after_launch("stencil2d_mesh_kernel");
}

// four rows in flight per thread: 566 GLUP/s on the 3-function 8192^2 mesh, two rows 549 (KX_MESH_U=2 for A/B)
template <typename T, int KY, int KX>
int launch_mesh(const MeshArgs<T>& a, dim3 grid, bool vec, cudaStream_t s) {
  const char* e = getenv("KX_MESH_U");
  if (e && atoi(e) == 2) return launch_mesh_v<T, KY, KX, 2>(a, grid, vec, s);
  return launch_mesh_v<T, KY, KX, 4>(a, grid, vec, s);
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
  while ((a.hout + a.th - 1) / a.th > 65535) a.th *= 2;
  if (a.th > kMaxTH) return KX_ERR_UNSUPPORTED;
  const bool vec = (a.win % 4 == 0) && ((reinterpret_cast<uintptr_t>(m.in) & 15) == 0) && (a.in_batch % 4 == 0);
  dim3 grid((a.wout + kMT * kMCols - 1) / (kMT * kMCols), (a.hout + a.th - 1) / a.th, (unsigned)batch);
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
