// Fast path: MULTI-FUNCTION kmap meshes (`F[slice] = fn` region assignment, kernel_interface.py:85-90)
// whose functions are all AFFINE, over the last two axes, stride 1: interior stencil + boundary
// conditions + source blocks in ONE pass.  The reference evaluates every function for every window
// under vmap and selects with lax.switch (kernex/_src/map.py:115-118); the generic kernel here
// re-derives the region of every window from the key table (0.04 of the HBM roofline).
//
// Same row-marching skeleton as kx_stencil2d.cu (a thread owns 4 output columns, keeps the last KY
// input rows in a register queue, 128-bit loads).  Region selection follows select_func
// (kernex/_src/utils.py:338-382: keys scanned from the LAST inserted, first match wins, no match =
// function 0) but is factored so that the row loop only does bit operations:
//   * column part of every key, per owned column: a 32-bit mask, computed once per thread;
//   * row part of every key (plus the leading batch axes), per row of the band: a 32-bit mask in
//     shared memory, computed once per block;
//   * per output: m = row_mask & col_mask, function = m ? func_of_key[31 - clz(m)] : 0.
// A thread keeps the coefficients (bias, KY*KX taps, tap mask) of ONE function in registers: the one its 4
// columns agree on, reloaded from the parameter bank only when the selection changes (at region boundaries).
// Inside a region the row loop is therefore the plain stencil loop with register coefficients, whatever the
// number of functions; a thread whose columns disagree (vertical region boundary) evaluates each column with
// that column's own coefficients.  Accumulation order: bias, then the window in row-major order (as
// kx_stencil2d.cu).
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
  int ay, ax, ndim;                      // axes of the 2-D window inside the key items
  int64_t lead_shape[KX_MAX_DIMS];       // leading (batch) axes, for keys that constrain them
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

// VW = 4: rows 16-byte aligned (SHIFT = (-lx) mod 4); VW = 1: scalar loads (SHIFT = 0)
// U = rows fetched back to back
template <typename T, int KY, int KX, int SHIFT, int VW, int U>
__global__ void __launch_bounds__(kMT)
stencil2d_mesh_kernel(const __grid_constant__ MeshArgs<T> a) {
  constexpr int NV = (SHIFT + KX + kMCols - 1 + VW - 1) / VW;
  constexpr int NE = NV * VW;
  __shared__ uint32_t row_mask[kMaxTH];
  const int tid = threadIdx.x;
  const int j0 = (blockIdx.x * kMT + tid) * kMCols;
  const int r0 = blockIdx.y * a.th, r1 = min(r0 + a.th, a.hout);
  const T* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int cy = (KY - 1) / 2, cx = (KX - 1) / 2;

  // ---- row part of the keys (and the leading axes), once per block ----
  if (tid < a.th) {
    const int rc = (r0 + tid) - a.ly + cy;                  // window-centre row in array coordinates
    int lead[KX_MAX_DIMS] = {0, 0, 0, 0};
    int64_t b = blockIdx.z;
    for (int d = a.ndim - 3; d >= 0; --d) {
      lead[d] = (int)(b % a.lead_shape[d]);
      b /= a.lead_shape[d];
    }
    uint32_t m = 0;
#pragma unroll 1
    for (int k = 0; k < a.n_keys; ++k) {
      bool ok = true;
#pragma unroll 1
      for (int d = 0; d < a.ndim; ++d) {
        if (d == a.ax) continue;
        ok = ok && item_match(a.key[k], d, d == a.ay ? rc : lead[d]);
      }
      if (ok) m |= 1u << k;
    }
    row_mask[tid] = m;
  }
  // ---- column part of the keys, once per thread ----
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
  __syncthreads();
  if (j0 >= a.wout) return;
  const int xa = j0 - a.lx - SHIFT;
  const int n_valid = min(kMCols, a.wout - j0);

  auto load_row = [&](int iy, T(&dst)[NE]) {
    const bool row_ok = (iy >= 0) && (iy < a.hin);
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

  T q[KY - 1 + U][NE];
#pragma unroll
  for (int ky = 0; ky < KY - 1; ++ky) load_row(r0 - a.ly + ky, q[ky]);
  int f[kMCols] = {0, 0, 0, 0};       // function of each owned column in the current row
  int cur_f = -1;                     // function whose coefficients sit in cb / cc / cm
  T cb = T(0), cc[KY * KX];
  uint32_t cm = 0;
#pragma unroll
  for (int t = 0; t < KY * KX; ++t) cc[t] = T(0);
  uint32_t rm_prev = 0;               // row mask the selection was computed for
  bool have_sel = false;

  for (int r = r0; r < r1; r += U) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (r + u < r1) load_row(r + u - a.ly + KY - 1, q[KY - 1 + u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u < r1) {
        const uint32_t rm = row_mask[r + u - r0];
        if (!have_sel || rm != rm_prev) {   // block-uniform: the selection only changes at a row boundary of some region
          rm_prev = rm;
          have_sel = true;
#pragma unroll
          for (int i = 0; i < kMCols; ++i) {
            const uint32_t m = rm & col_mask[i];
            f[i] = m ? (int)a.key_func[31 - __clz(m)] : 0;
          }
        }
        T acc[kMCols];
        if (f[0] == f[1] && f[1] == f[2] && f[2] == f[3]) {
          if (f[0] != cur_f) {   // entering another region: this thread's coefficient registers follow
            cur_f = f[0];
            cb = a.bias[cur_f];
            cm = a.mask[cur_f];
#pragma unroll
            for (int t = 0; t < KY * KX; ++t) cc[t] = a.coef[cur_f][t];
          }
#pragma unroll
          for (int i = 0; i < kMCols; ++i) acc[i] = cb;
#pragma unroll
          for (int ky = 0; ky < KY; ++ky)
#pragma unroll
            for (int kx = 0; kx < KX; ++kx)
              if (cm & (1u << (ky * KX + kx))) {
#pragma unroll
                for (int i = 0; i < kMCols; ++i) acc[i] = mad_wrap(cc[ky * KX + kx], q[u + ky][SHIFT + kx + i], acc[i]);
              }
        } else {
          // vertical region boundary inside the thread's 4 columns: every column with its own function
#pragma unroll
          for (int i = 0; i < kMCols; ++i) {
            const int fi = f[i];
            const uint32_t m = a.mask[fi];
            T v = a.bias[fi];
#pragma unroll
            for (int ky = 0; ky < KY; ++ky)
#pragma unroll
              for (int kx = 0; kx < KX; ++kx)
                if (m & (1u << (ky * KX + kx))) v = mad_wrap(a.coef[fi][ky * KX + kx], q[u + ky][SHIFT + kx + i], v);
            acc[i] = v;
          }
        }
        T* o = out + (int64_t)(r + u) * a.wout + j0;
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
      }
    }
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
      for (int e = 0; e < NE; ++e) q[ky][e] = q[U + ky][e];
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
  // band height measured on the 3-function 8192^2 mesh (U = 4): 8 rows 484 GLUP/s, 12 531, 16 550, 24 563, 32 566,
  // 48 529, 64 512, 128 400
  a.th = a.hout >= 512 ? 32 : (a.hout >= 64 ? 16 : 8);
  if (const char* e = getenv("KX_MESH_TH")) a.th = atoi(e) > 0 ? atoi(e) : a.th;
  while ((a.hout + a.th - 1) / a.th > 65535) a.th *= 2;
  if (a.th > kMaxTH) return KX_ERR_UNSUPPORTED;
  const bool vec = (a.win % 4 == 0) && ((reinterpret_cast<uintptr_t>(m.in) & 15) == 0) && (a.in_batch % 4 == 0);
  dim3 grid((a.wout + kMT * kMCols - 1) / (kMT * kMCols), (a.hout + a.th - 1) / a.th, (unsigned)batch);
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
