/*
 * kx_oracle_c.c -- C restatement of the kernex kmap/kscan hot path for the CPU baseline.
 * TEST / BENCH INFRASTRUCTURE ONLY: never linked into or loaded by kernex_b200.
 *
 * Same semantics as oracle/kx_oracle.py (which is pinned by the reference's golden
 * vectors; tests/test_oracle_c.py checks this file against it): window j of axis d covers
 * input coordinates [-lo + j*s, -lo + j*s + k) (kernex/_src/utils.py:205-242), cells
 * outside [0, n) read zero (utils.py:43-53, map.py:85), results are laid out in row-major
 * window order (utils.py:245-280, map.py:98).  Instead of materialising index views and
 * gathering a patch per window (map.py:46) it walks output rows and accumulates one tap
 * at a time over the in-bounds column range, which is what a hand-written CPU stencil does;
 * OpenMP spreads output rows over the host cores.
 *
 *   kxo_affine_f32 / kxo_affine_i32   out = bias + sum_t coef[t] * x[origin + tap[t]]
 *   kxo_patch_f32                     out[j, t] = x[origin(j) + tap[t]]
 *   kxo_rowmarch_f32                  2-D kscan whose functions read only row n-1
 *                                     (kernex/_src/scan.py:84-113 restricted to that class)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KXO_MAX_DIMS 4

typedef struct {
  int32_t ndim;
  int64_t in_shape[KXO_MAX_DIMS], out_shape[KXO_MAX_DIMS];
  int32_t stride[KXO_MAX_DIMS], lo[KXO_MAX_DIMS];
} KxoGeom;

int kxo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define KXO_AFFINE(NAME, T)                                                                         \
  void NAME(const KxoGeom* g, const T* in, T* out, int ntaps, const int32_t* tap /*[ntaps][ndim]*/, \
            const T* coef, T bias, int nthreads) {                                                  \
    const int nd = g->ndim;                                                                         \
    int64_t istr[KXO_MAX_DIMS], rows = 1;                                                           \
    istr[nd - 1] = 1;                                                                               \
    for (int d = nd - 2; d >= 0; --d) istr[d] = istr[d + 1] * g->in_shape[d + 1];                   \
    for (int d = 0; d < nd - 1; ++d) rows *= g->out_shape[d];                                       \
    const int64_t wout = g->out_shape[nd - 1], win = g->in_shape[nd - 1];                           \
    const int sx = g->stride[nd - 1], lx = g->lo[nd - 1];                                           \
    if (nthreads < 1) nthreads = 1;                                                                 \
    _Pragma("omp parallel for schedule(static) num_threads(nthreads)")                              \
    for (int64_t r = 0; r < rows; ++r) {                                                            \
      int64_t j[KXO_MAX_DIMS], rem = r;                                                             \
      for (int d = nd - 2; d >= 0; --d) { j[d] = rem % g->out_shape[d]; rem /= g->out_shape[d]; }   \
      T* orow = out + r * wout;                                                                     \
      for (int64_t x = 0; x < wout; ++x) orow[x] = bias;                                            \
      for (int t = 0; t < ntaps; ++t) {                                                             \
        int ok = 1;                                                                                 \
        int64_t base = 0;                                                                           \
        for (int d = 0; d < nd - 1; ++d) {                                                          \
          const int64_t c = j[d] * g->stride[d] - g->lo[d] + tap[t * nd + d];                       \
          if (c < 0 || c >= g->in_shape[d]) ok = 0;                                                 \
          base += c * istr[d];                                                                      \
        }                                                                                           \
        if (!ok) continue; /* whole row of this tap lies in the zero pad */                         \
        const int64_t off = tap[t * nd + nd - 1] - lx; /* input col = x*sx + off */                 \
        int64_t x0 = 0, x1 = wout;                                                                  \
        if (off < 0) x0 = (-off + sx - 1) / sx;                                                     \
        if ((x1 - 1) * sx + off >= win) x1 = (win - 1 - off) / sx + 1;                              \
        const T c = coef[t];                                                                        \
        const T* irow = in + base + off;                                                            \
        if (sx == 1) {                                                                              \
          for (int64_t x = x0; x < x1; ++x) orow[x] += c * irow[x];                                 \
        } else {                                                                                    \
          for (int64_t x = x0; x < x1; ++x) orow[x] += c * irow[x * sx];                            \
        }                                                                                           \
      }                                                                                             \
    }                                                                                               \
  }

KXO_AFFINE(kxo_affine_f32, float)
KXO_AFFINE(kxo_affine_i32, int32_t)

void kxo_patch_f32(const KxoGeom* g, const float* in, float* out, int ntaps, const int32_t* tap, int nthreads) {
  const int nd = g->ndim;
  int64_t istr[KXO_MAX_DIMS], nwin = 1;
  istr[nd - 1] = 1;
  for (int d = nd - 2; d >= 0; --d) istr[d] = istr[d + 1] * g->in_shape[d + 1];
  for (int d = 0; d < nd; ++d) nwin *= g->out_shape[d];
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t w = 0; w < nwin; ++w) {
    int64_t o[KXO_MAX_DIMS], rem = w;
    for (int d = nd - 1; d >= 0; --d) {
      o[d] = (rem % g->out_shape[d]) * g->stride[d] - g->lo[d];
      rem /= g->out_shape[d];
    }
    for (int t = 0; t < ntaps; ++t) {
      int ok = 1;
      int64_t off = 0;
      for (int d = 0; d < nd; ++d) {
        const int64_t c = o[d] + tap[t * nd + d];
        if (c < 0 || c >= g->in_shape[d]) ok = 0;
        off += c * istr[d];
      }
      out[w * ntaps + t] = ok ? in[off] : 0.0f;
    }
  }
}

/* 2-D kscan, kernel (3,3), border ((1,1),(1,1)), stride 1, functions of the form
 *   region r (cols [c0,c1), rows [r0,r1), later regions win):  out = bias + a0*u[n-1][i-1] + a1*u[n-1][i] + a2*u[n-1][i+1]
 * where u[-1][*] and the pad columns are zero.  Rows are sequential, columns parallel. */
typedef struct {
  int64_t r0, r1, c0, c1;
  float a[3], bias;
} KxoRegion;

void kxo_rowmarch_f32(int64_t nt, int64_t nx, int nreg, const KxoRegion* reg, float* out, int nthreads) {
  float* prev = (float*)calloc((size_t)nx + 2, sizeof(float));
  if (nthreads < 1) nthreads = 1;
  for (int64_t n = 0; n < nt; ++n) {
    float* row = out + n * nx;
    for (int q = 0; q < nreg; ++q) {
      const KxoRegion R = reg[q];
      if (n < R.r0 || n >= R.r1) continue;
#pragma omp parallel for schedule(static) num_threads(nthreads)
      for (int64_t i = R.c0; i < R.c1; ++i)
        row[i] = R.bias + R.a[0] * prev[i] + R.a[1] * prev[i + 1] + R.a[2] * prev[i + 2];
    }
    memcpy(prev + 1, row, (size_t)nx * sizeof(float));
  }
  free(prev);
}

/* ------------------------------------------------------------------------------------------
 * kxo_map_views_f32 -- the reference's kmap ALGORITHM, step by step, instead of the streaming
 * stencil above (which is what a hand-written CPU code does, not what kernex does):
 *   1. jnp.pad with ALL padding appended at the end of each axis (kernex/_src/utils.py:43-53,
 *      map.py:85); the low-side pad cells are reached through negative-index wraparound;
 *   2. _generate_views (utils.py:74-90): per axis an int32 table [out_d][k_d] of window indices
 *      -lo + j*s + i (general_arange, utils.py:205-242), then the Cartesian product over the
 *      axes in row-major window order (general_product, utils.py:259-280): N_win x k_d int32 per
 *      axis, MATERIALISED, rebuilt on every call (kernel_interface.py:130-157);
 *   3. per window (vmap): gather the patch array[ix_(*view)] (map.py:46, utils.py:184-192), roll it
 *      by -((k-1)//2) per axis when `relative` (utils.py:147-181), then evaluate the window
 *      function on the patch: sum_t coef[t] * patch[tap[t]] with taps in PATCH coordinates (after
 *      the roll) -- the caller lists every term the user wrote, zero coefficients included.
 * Used by bench.py's reference arm / cpu_baseline as the closest CPU stand-in for kernex on
 * JAX's CPU backend (JAX itself is not installable here).  2-D only (config 2 of BASELINE.json).
 * `views` is caller-provided scratch of n_win * (k0 + k1) int32; `padded` of (n0+p0)*(n1+p1) floats.
 */
void kxo_map_views_f32(const KxoGeom* g, const int32_t* ksize, const int32_t* hi, const float* in, float* out,
                       int ntaps, const int32_t* tap /*[ntaps][2], patch coords*/, const float* coef, float bias,
                       int relative, float* padded, int32_t* views, int nthreads) {
  if (g->ndim != 2) return;
  if (nthreads < 1) nthreads = 1;
  const int64_t n0 = g->in_shape[0], n1 = g->in_shape[1];
  const int k0 = ksize[0], k1 = ksize[1];
  const int64_t p0 = (g->lo[0] > 0 ? g->lo[0] : 0) + (hi[0] > 0 ? hi[0] : 0);
  const int64_t p1 = (g->lo[1] > 0 ? g->lo[1] : 0) + (hi[1] > 0 ? hi[1] : 0);
  const int64_t m0 = n0 + p0, m1 = n1 + p1;             /* padded shape */
  const int64_t o0 = g->out_shape[0], o1 = g->out_shape[1], nwin = o0 * o1;
  /* 1. pad at the end */
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t r = 0; r < m0; ++r) {
    float* dst = padded + r * m1;
    if (r < n0) {
      memcpy(dst, in + r * n1, (size_t)n1 * sizeof(float));
      for (int64_t c = n1; c < m1; ++c) dst[c] = 0.0f;
    } else {
      for (int64_t c = 0; c < m1; ++c) dst[c] = 0.0f;
    }
  }
  /* 2. views: product of the per-axis index tables, materialised per window */
  int32_t* v0 = views;                  /* [nwin][k0] */
  int32_t* v1 = views + nwin * k0;      /* [nwin][k1] */
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t w = 0; w < nwin; ++w) {
    const int64_t j0 = w / o1, j1 = w % o1;
    for (int i = 0; i < k0; ++i) v0[w * k0 + i] = (int32_t)(-g->lo[0] + j0 * g->stride[0] + i);
    for (int i = 0; i < k1; ++i) v1[w * k1 + i] = (int32_t)(-g->lo[1] + j1 * g->stride[1] + i);
  }
  /* 3. gather -> roll -> function, one window at a time */
  const int r0 = relative ? (k0 - 1) / 2 : 0, r1 = relative ? (k1 - 1) / 2 : 0;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t w = 0; w < nwin; ++w) {
    float patch[25], rolled[25];
    for (int a = 0; a < k0; ++a) {
      int64_t ia = v0[w * k0 + a];
      if (ia < 0) ia += m0;             /* negative index = from the end (the appended zeros) */
      for (int b = 0; b < k1; ++b) {
        int64_t ib = v1[w * k1 + b];
        if (ib < 0) ib += m1;
        patch[a * k1 + b] = padded[ia * m1 + ib];
      }
    }
    for (int a = 0; a < k0; ++a)       /* roll by -r: rolled[a] = patch[(a + r) mod k] */
      for (int b = 0; b < k1; ++b) rolled[a * k1 + b] = patch[((a + r0) % k0) * k1 + (b + r1) % k1];
    float acc = bias;
    for (int t = 0; t < ntaps; ++t) acc += coef[t] * rolled[tap[2 * t] * k1 + tap[2 * t + 1]];
    out[w] = acc;
  }
}
