// extern "C" entry points of libkx_b200 (declared in include/kx_b200.h): validation and
// dispatch to the kernel templates.  Every entry point only enqueues on `stream`.
#include <algorithm>
#include <vector>
#include "kx_internal.cuh"

using namespace kx;

namespace {

int validate(const KxGeom* g, const KxProgram* p) {
  if (!g || !p) return fail(KX_ERR_INVALID, "null geometry / program");
  if (g->ndim < 1 || g->ndim > KX_MAX_DIMS) return fail(KX_ERR_INVALID, "ndim=%d outside [1,%d]", g->ndim, KX_MAX_DIMS);
  if ((g->in_dtype != KX_F32 && g->in_dtype != KX_I32) || (g->out_dtype != KX_F32 && g->out_dtype != KX_I32))
    return fail(KX_ERR_INVALID, "unknown dtype in=%d out=%d", g->in_dtype, g->out_dtype);
  for (int d = 0; d < g->ndim; ++d) {
    if (g->in_shape[d] < 0 || g->ksize[d] < 1 || g->stride[d] < 1)
      return fail(KX_ERR_INVALID, "axis %d: shape=%lld ksize=%d stride=%d", d, (long long)g->in_shape[d], g->ksize[d],
                  g->stride[d]);
    int64_t span = g->in_shape[d] + g->lo[d] + g->hi[d] - g->ksize[d];
    int64_t expect = span >= 0 ? span / g->stride[d] + 1 : 0;
    if (expect < 0) expect = 0;
    if (g->out_shape[d] != expect)
      return fail(KX_ERR_INVALID, "axis %d: out_shape=%lld but (n+lo+hi-k)/s+1=%lld", d, (long long)g->out_shape[d],
                  (long long)expect);
  }
  if (p->n_funcs < 1 || p->n_funcs > KX_MAX_FUNCS) return fail(KX_ERR_INVALID, "n_funcs=%d", p->n_funcs);
  if (p->n_taps < 0 || p->n_taps > KX_MAX_TAPS) return fail(KX_ERR_INVALID, "n_taps=%d", p->n_taps);
  if (p->n_keys < 0 || p->n_keys > KX_MAX_KEYS) return fail(KX_ERR_INVALID, "n_keys=%d", p->n_keys);
  for (int f = 0; f < p->n_funcs; ++f) {
    const KxFunc& fn = p->func[f];
    if (fn.kind < KX_AFFINE || fn.kind > KX_GATHER) return fail(KX_ERR_INVALID, "function %d: kind=%d", f, fn.kind);
    if (fn.tap_begin < 0 || fn.tap_end < fn.tap_begin || fn.tap_end > p->n_taps)
      return fail(KX_ERR_INVALID, "function %d: taps [%d,%d) of %d", f, fn.tap_begin, fn.tap_end, p->n_taps);
    if (fn.kind != KX_AFFINE && fn.tap_end == fn.tap_begin)
      return fail(KX_ERR_INVALID, "function %d: reduce / gather needs at least one tap", f);
    for (int t = fn.tap_begin; t < fn.tap_end; ++t)
      for (int d = 0; d < g->ndim; ++d)
        if (p->tap_off[t][d] < 0 || p->tap_off[t][d] >= g->ksize[d])
          return fail(KX_ERR_INVALID, "tap %d axis %d: offset %d outside the window [0,%d)", t, d, p->tap_off[t][d],
                      g->ksize[d]);
  }
  for (int k = 0; k < p->n_keys; ++k) {
    if (p->key[k].func < 0 || p->key[k].func >= p->n_funcs || p->key[k].n_items < 0 || p->key[k].n_items > g->ndim)
      return fail(KX_ERR_INVALID, "region key %d is malformed", k);
    if (k > 0 && p->key[k].func < p->key[k - 1].func)
      return fail(KX_ERR_INVALID, "region keys must be grouped by function in insertion order");
  }
  if (p->n_funcs > 1)
    for (int f = 0; f < p->n_funcs; ++f)
      if (p->func[f].kind == KX_GATHER && p->func[f].tap_end - p->func[f].tap_begin != 1)
        return fail(KX_ERR_INVALID, "vector-valued functions cannot be part of a multi-function program");
  return KX_OK;
}

void fill_map_args(MapArgs& a, const KxGeom& g, const KxProgram& p, const void* in, const void* coef_dev,
                   int64_t coef_batch_stride, void* out, int64_t batch) {
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.in = in;
  a.coef_dev = coef_dev;
  a.out = out;
  a.n_win = prod(g.out_shape, g.ndim);
  a.in_elems = prod(g.in_shape, g.ndim);
  a.fn_elems = (p.n_funcs == 1 && p.func[0].kind == KX_GATHER) ? p.func[0].tap_end - p.func[0].tap_begin : 1;
  a.out_batch_stride = a.n_win * a.fn_elems;
  a.coef_batch_stride = coef_batch_stride;
  a.batch = batch;
  int64_t pitch = 1;
  for (int d = g.ndim - 1; d >= 0; --d) {
    a.out_pitch[d] = pitch;
    pitch *= g.out_shape[d];
  }
  a.out_base = 0;
}

int dispatch_map(const MapArgs& a, const KxProgram& prog, cudaStream_t s);

// Multi-function meshes that none of the fast templates takes (3-D windows -- a heat equation with Dirichlet faces --,
// window reductions, wide windows, more than 8 functions): `F[slice] = fn` regions are boxes, so the function that
// covers the bulk of the domain runs as a SINGLE-function map over every window (on whatever fast template takes it:
// 0.8-1.0 of the HBM roofline instead of the generic kernel's 0.05), and the boxes of the OTHER functions' keys are
// recomputed afterwards by the generic kernel with the full region selection of kernex/_src/utils.py:338-382 (so key
// priority and overlaps need no special care: a recomputed cell always gets the function the reference selects).
// Declines (-> one generic launch over everything) for strided keys, vector-valued functions, device coefficients,
// batches, when a window could match no key while the bulk function is not function 0, or when the boxes are more
// than 30 % of the domain.  KX_NO_DECOMPOSE=1 switches it off (cross-check).
int try_mesh_decompose(const MapArgs& a, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = a.g;
  const int nd = g.ndim;
  if (p.n_funcs < 2 || a.batch != 1 || a.fn_elems != 1 || a.use_box || a.coef_dev || a.in_bcast) return KX_ERR_UNSUPPORTED;
  if (getenv("KX_NO_DECOMPOSE")) return KX_ERR_UNSUPPORTED;
  for (int f = 0; f < p.n_funcs; ++f)
    if (p.func[f].kind == KX_GATHER) return KX_ERR_UNSUPPORTED;
  struct Box {
    int64_t j0[KX_MAX_DIMS], j1[KX_MAX_DIMS];
    int64_t vol;
    bool whole;
  };
  auto ceil_div = [](int64_t n, int64_t d) { return n >= 0 ? (n + d - 1) / d : -((-n) / d); };
  static thread_local Box box[KX_MAX_KEYS];
  const int64_t kFar = (int64_t)1 << 40;
  for (int k = 0; k < p.n_keys; ++k) {
    Box& b = box[k];
    b.vol = 1;
    b.whole = true;
    for (int d = 0; d < nd; ++d) {
      int64_t lo_j = 0, hi_j = g.out_shape[d];
      if (d < p.key[k].n_items && p.key[k].item[d].kind != KX_KEY_ANY) {
        const KxKeyItem& it = p.key[k].item[d];
        if (it.kind == KX_KEY_RANGE_STEP && it.step != 1 && it.step != -1) return KX_ERR_UNSUPPORTED;   // c % 1 == 0 always
        int64_t ca = it.a, cb = it.kind == KX_KEY_EQ ? it.a + 1 : it.b;     // window centres c with ca <= c < cb
        ca = std::max(ca, -kFar);
        cb = std::min(cb, kFar);
        // c(j) = j * stride - lo + (k - 1) / 2
        const int64_t sh = (int64_t)g.lo[d] - (g.ksize[d] - 1) / 2, st = g.stride[d];
        lo_j = std::max(lo_j, ceil_div(ca + sh, st));
        hi_j = std::min(hi_j, ceil_div(cb + sh, st));
      }
      if (hi_j < lo_j) hi_j = lo_j;
      b.j0[d] = lo_j;
      b.j1[d] = hi_j;
      b.vol *= hi_j - lo_j;
      b.whole = b.whole && lo_j == 0 && hi_j == g.out_shape[d];
    }
  }
  // the bulk function: the one the reference selects for the window in the middle of the domain
  int dom = 0;
  {
    int64_t centre[KX_MAX_DIMS];
    for (int d = 0; d < nd; ++d) centre[d] = (g.out_shape[d] / 2) * g.stride[d] - g.lo[d] + (g.ksize[d] - 1) / 2;
    for (int k = p.n_keys - 1; k >= 0; --k) {
      bool ok = true;
      for (int d = 0; d < p.key[k].n_items && d < nd; ++d) {
        const KxKeyItem& it = p.key[k].item[d];
        if (it.kind == KX_KEY_EQ) ok = ok && centre[d] == it.a;
        else if (it.kind == KX_KEY_RANGE || it.kind == KX_KEY_RANGE_STEP) ok = ok && it.a <= centre[d] && centre[d] < it.b;
      }
      if (ok) {
        dom = p.key[k].func;
        break;
      }
    }
  }
  bool covered = dom == 0;           // a window no key matches gets function 0 (utils.py:379)
  for (int k = 0; k < p.n_keys; ++k) covered = covered || box[k].whole;
  if (!covered) return KX_ERR_UNSUPPORTED;
  // cells to recompute: the box of every key of another function, minus the boxes of the LATER-inserted keys of the
  // bulk function (they win there: `F[:, :, :] = bc` followed by `F[1:-1, 1:-1, 1:-1] = step` leaves the six faces)
  struct Cut {
    int64_t j0[KX_MAX_DIMS], j1[KX_MAX_DIMS];
  };
  std::vector<Cut> todo, cur, nxt;
  for (int k = 0; k < p.n_keys; ++k) {
    if (p.key[k].func == dom || box[k].vol == 0) continue;
    cur.clear();
    Cut c0;
    for (int d = 0; d < KX_MAX_DIMS; ++d) c0.j0[d] = d < nd ? box[k].j0[d] : 0, c0.j1[d] = d < nd ? box[k].j1[d] : 1;
    cur.push_back(c0);
    for (int q = k + 1; q < p.n_keys && !cur.empty(); ++q) {
      if (p.key[q].func != dom || box[q].vol == 0) continue;
      nxt.clear();
      for (const Cut& b : cur) {
        bool overlap = true;
        for (int d = 0; d < nd; ++d) overlap = overlap && box[q].j0[d] < b.j1[d] && box[q].j1[d] > b.j0[d];
        if (!overlap) {
          nxt.push_back(b);
          continue;
        }
        Cut r = b;                                   // peel off the slabs of b outside box q, axis by axis
        for (int d = 0; d < nd; ++d) {
          if (r.j0[d] < box[q].j0[d]) {
            Cut lo = r;
            lo.j1[d] = box[q].j0[d];
            nxt.push_back(lo);
            r.j0[d] = box[q].j0[d];
          }
          if (r.j1[d] > box[q].j1[d]) {
            Cut hi = r;
            hi.j0[d] = box[q].j1[d];
            nxt.push_back(hi);
            r.j1[d] = box[q].j1[d];
          }
        }
      }
      cur.swap(nxt);
      if (cur.size() > 256) return KX_ERR_UNSUPPORTED;
    }
    todo.insert(todo.end(), cur.begin(), cur.end());
    if (todo.size() > 256) return KX_ERR_UNSUPPORTED;
  }
  int64_t patch = 0;
  for (const Cut& b : todo) {
    int64_t v = 1;
    for (int d = 0; d < nd; ++d) v *= b.j1[d] - b.j0[d];
    patch += v;
  }
  if (patch * 10 > a.n_win * 3) return KX_ERR_UNSUPPORTED;

  static thread_local KxProgram p1;
  memset(&p1, 0, sizeof(p1));
  p1.n_funcs = 1;
  p1.func[0] = p.func[dom];
  p1.n_taps = p.func[dom].tap_end - p.func[dom].tap_begin;
  p1.func[0].tap_begin = 0;
  p1.func[0].tap_end = p1.n_taps;
  for (int t = 0; t < p1.n_taps; ++t) {
    for (int d = 0; d < KX_MAX_DIMS; ++d) p1.tap_off[t][d] = p.tap_off[p.func[dom].tap_begin + t][d];
    p1.coef[t] = p.coef[p.func[dom].tap_begin + t];
  }
  int rc = dispatch_map(a, p1, s);
  if (rc) return rc;
  for (const Cut& b : todo) {
    MapArgs ab = a;
    ab.n_win = 1;
    for (int d = 0; d < nd; ++d) {
      ab.g.out_shape[d] = b.j1[d] - b.j0[d];
      ab.n_win *= ab.g.out_shape[d];
      ab.g.lo[d] = (int32_t)(g.lo[d] - b.j0[d] * g.stride[d]);          // window j of the box = window j0 + j of the map
      ab.out_base += b.j0[d] * a.out_pitch[d];
    }
    if (ab.n_win == 0) continue;
    rc = launch_generic_map(ab, p, s);
    if (rc) return rc;
  }
  return KX_OK;
}

// fast templates first, the generic kernel as the universal backstop
int dispatch_map(const MapArgs& a, const KxProgram& prog, cudaStream_t s) {
  int rc = try_stencil2d(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_stencil2d_mesh(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_conv2d(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_convc(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_stencil3d(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_patch2d(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_patch3d(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_patch(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_pool2d(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_pool3d(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_mesh_decompose(a, prog, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  return launch_generic_map(a, prog, s);
}

// The input-VJP of a stride-1 AFFINE stencil is itself an AFFINE stencil of the cotangent:
//   dx[i] = sum_t c_t * g[i + lo - tap_t]  =  stencil(g) with taps (k-1-tap_t), borders (k-1-lo, k-1-hi)
// so it runs on the same roofline-tuned kernels as the forward pass.
bool transpose_program(const KxGeom& g, const KxProgram& p, KxGeom& gt, KxProgram& pt) {
  for (int d = 0; d < g.ndim; ++d)
    if (g.stride[d] != 1) return false;
  memset(&gt, 0, sizeof(gt));
  gt.ndim = g.ndim;
  gt.in_dtype = gt.out_dtype = KX_F32;
  for (int d = 0; d < g.ndim; ++d) {
    gt.in_shape[d] = g.out_shape[d];
    gt.out_shape[d] = g.in_shape[d];
    gt.ksize[d] = g.ksize[d];
    gt.stride[d] = 1;
    gt.lo[d] = g.ksize[d] - 1 - g.lo[d];
    gt.hi[d] = g.ksize[d] - 1 - g.hi[d];
  }
  memset(&pt, 0, sizeof(pt));
  pt.n_funcs = 1;
  pt.n_taps = p.func[0].tap_end - p.func[0].tap_begin;
  pt.func[0] = p.func[0];
  pt.func[0].tap_begin = 0;
  pt.func[0].tap_end = pt.n_taps;
  pt.func[0].bias = 0.0;
  for (int t = 0; t < pt.n_taps; ++t) {
    const int src = p.func[0].tap_begin + t;
    for (int d = 0; d < g.ndim; ++d) pt.tap_off[t][d] = (int16_t)(g.ksize[d] - 1 - p.tap_off[src][d]);
    pt.coef[t] = p.coef[src];
  }
  return true;
}

// conv-as-kmap over the WHOLE leading axis: window (C, KY, KX) with ksize[0] == in_shape[0] (output depth 1), stride 1,
// every tap present in row-major (c, ky, kx) order -- tests_and_benchmarks/test_benchmark_conv2d.py, C up to 64
bool full_depth_conv(const KxGeom& g, const KxProgram& p, int& C, int& NT) {
  if (g.ndim != 3 || g.in_shape[0] < 2 || g.ksize[0] != g.in_shape[0] || g.lo[0] != 0 || g.hi[0] != 0) return false;
  for (int d = 0; d < 3; ++d)
    if (g.stride[d] != 1) return false;
  C = g.ksize[0];
  NT = g.ksize[1] * g.ksize[2];
  const KxFunc& f = p.func[0];
  if (f.tap_end - f.tap_begin != C * NT || f.post_div != 0.0) return false;
  for (int t = 0; t < C * NT; ++t) {
    const int16_t* o = p.tap_off[f.tap_begin + t];
    if ((o[0] * g.ksize[1] + o[1]) * g.ksize[2] + o[2] != t) return false;
  }
  return true;
}

// dx of a full-depth conv: channel c of dx is the 2-D flipped-tap stencil of the (single-plane) cotangent with the
// weights of channel c.  Device weights: ONE batched stencil2d launch (batch = C, every item reads the same g, item c
// takes weights [c*NT, (c+1)*NT) in reversed order).  Host weights: one 2-D launch per channel.
int vjp_input_full_depth(const KxGeom& g, const KxProgram& p, int C, int NT, const void* gout, const void* coef_dev,
                         void* dx, cudaStream_t s) {
  static thread_local KxGeom gt;
  static thread_local KxProgram pt;
  memset(&gt, 0, sizeof(gt));
  gt.ndim = 2;
  gt.in_dtype = gt.out_dtype = KX_F32;
  for (int d = 0; d < 2; ++d) {
    gt.in_shape[d] = g.out_shape[d + 1];
    gt.out_shape[d] = g.in_shape[d + 1];
    gt.ksize[d] = g.ksize[d + 1];
    gt.stride[d] = 1;
    gt.lo[d] = g.ksize[d + 1] - 1 - g.lo[d + 1];
    gt.hi[d] = g.ksize[d + 1] - 1 - g.hi[d + 1];
  }
  memset(&pt, 0, sizeof(pt));
  pt.n_funcs = 1;
  pt.n_taps = NT;
  pt.func[0].kind = KX_AFFINE;
  pt.func[0].tap_begin = 0;
  pt.func[0].tap_end = NT;
  pt.func[0].coef_is_int = p.func[0].coef_is_int;
  for (int t = 0; t < NT; ++t) {   // tap t = (ky, kx) row-major reads the flipped offset: table order is REVERSED slot order
    pt.tap_off[t][0] = (int16_t)(g.ksize[1] - 1 - t / g.ksize[2]);
    pt.tap_off[t][1] = (int16_t)(g.ksize[2] - 1 - t % g.ksize[2]);
  }
  const int first = p.func[0].tap_begin;
  const int64_t plane_out = gt.out_shape[0] * gt.out_shape[1];
  if (plane_out <= 0 || gt.in_shape[0] * gt.in_shape[1] <= 0) return KX_ERR_UNSUPPORTED;
  MapArgs a;
  if (coef_dev) {
    fill_map_args(a, gt, pt, gout, (const float*)coef_dev + first, NT, dx, C);
    a.in_bcast = 1;
    return try_stencil2d(a, pt, s);   // KX_ERR_UNSUPPORTED: the caller falls back to the generic kernel
  }
  for (int c = 0; c < C; ++c) {
    for (int t = 0; t < NT; ++t) pt.coef[t] = p.coef[first + c * NT + t];
    fill_map_args(a, gt, pt, gout, nullptr, 0, (float*)dx + (int64_t)c * plane_out, 1);
    const int rc = dispatch_map(a, pt, s);
    if (rc) return rc;
  }
  return KX_OK;
}

}  // namespace

// rows x width 32-bit words, pitches in words (cudaMemcpy2D / 3D device-to-device copies of unaligned pitched rows ran at
// ~0.3 TB/s here)
__global__ void __launch_bounds__(256) copy2d_kernel(uint32_t* __restrict__ dst, int64_t dpitch, const uint32_t* __restrict__ src,
                                                     int64_t spitch, int width, int64_t rows) {
  const int64_t total = rows * width;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    dst[r * dpitch + c] = __ldg(src + r * spitch + c);
  }
}

int copy2d(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width, int64_t rows, cudaStream_t s) {
  if (width <= 0 || rows <= 0) return KX_OK;
  if (dpitch == width && spitch == width) {
    KX_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)(rows * width) * 4, cudaMemcpyDeviceToDevice, s));
    return KX_OK;
  }
  int64_t blocks = (rows * width + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  copy2d_kernel<<<(unsigned)blocks, 256, 0, s>>>((uint32_t*)dst, dpitch, (const uint32_t*)src, spitch, (int)width, rows);
  return after_launch("copy2d_kernel");
}

// dst <- src on the box [lo, hi) of a C-contiguous array of rank <= 4 (missing leading axes: extent 1)
struct BoxCopy {
  int64_t lo[4], ext[4], pitch[4];
};
__global__ void __launch_bounds__(256) box_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                                                       const __grid_constant__ BoxCopy b) {
  const int64_t total = b.ext[0] * b.ext[1] * b.ext[2] * b.ext[3];
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t r = i, off = 0;
#pragma unroll
    for (int d = 3; d >= 0; --d) {
      const int64_t j = r % b.ext[d];
      r /= b.ext[d];
      off += (b.lo[d] + j) * b.pitch[d];
    }
    dst[off] = __ldg(src + off);
  }
}

// every cell OUTSIDE the box [lo, hi) of an nd-dimensional array: dst <- src, as 2 nd slabs (axis d: below / above the
// box, restricted to the box on the axes before d so that no cell is copied twice)
int copy_outside_box(void* dst, const void* src, const int64_t* shape, int nd, const int64_t* lo, const int64_t* hi, cudaStream_t s) {
  for (int d = 0; d < nd; ++d) {
    for (int side = 0; side < 2; ++side) {
      BoxCopy b;
      int64_t pitch = 1, total = 1;
      for (int q = 3; q >= 0; --q) {
        const int a = q - (4 - nd);                 // array axis of slot q (negative: padding slot)
        if (a < 0) {
          b.lo[q] = 0, b.ext[q] = 1, b.pitch[q] = 0;
          continue;
        }
        b.pitch[q] = pitch;
        pitch *= shape[a];
        if (a < d) b.lo[q] = lo[a], b.ext[q] = hi[a] - lo[a];
        else if (a == d) b.lo[q] = side == 0 ? 0 : hi[a], b.ext[q] = side == 0 ? lo[a] : shape[a] - hi[a];
        else b.lo[q] = 0, b.ext[q] = shape[a];
        total *= b.ext[q];
      }
      if (total <= 0) continue;
      int64_t blocks = (total + 255) / 256;
      const int64_t cap = (int64_t)sm_count() * 16;
      if (blocks > cap) blocks = cap;
      box_copy_kernel<<<(unsigned)blocks, 256, 0, s>>>((uint32_t*)dst, (const uint32_t*)src, b);
      const int rc = after_launch("box_copy_kernel");
      if (rc) return rc;
    }
  }
  return KX_OK;
}

// the cells of an H x W plane OUTSIDE the box [y0, y1) x [x0, x1): dst <- src (the ring an offset scan leaves untouched)
__global__ void __launch_bounds__(256) ring_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int H, int W,
                                                        int y0, int y1, int x0, int x1) {
  const int ring_w = x0 + (W - x1);                      // ring cells of a row that crosses the box
  const int64_t full_rows = (int64_t)y0 + (H - y1);
  const int64_t total = full_rows * W + (int64_t)(y1 - y0) * ring_w;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t r, c;
    if (i < full_rows * W) {
      r = i / W;
      c = i - r * W;
      if (r >= y0) r += y1 - y0;
    } else {
      const int64_t k = i - full_rows * W;
      r = y0 + k / ring_w;
      c = k - (k / ring_w) * ring_w;
      if (c >= x0) c += x1 - x0;
    }
    dst[r * W + c] = __ldg(src + r * W + c);
  }
}

// kscan / sscan over a 3-D array whose function reads only the PLANE ABOVE the cell it writes (axis-0 offset -1 from the
// window centre): explicit time marching with two space axes -- the reference's own diffusion test
// (tests/test_interface.py:28-97, an sscan with kernel (3, 3, 3)), the 3-D analogue of BASELINE config 4.  Window
// plane n reads state plane n + c0 - 1 and writes the window centres of state plane n + c0, so output plane n is a 2-D
// kmap of a plane that is complete before the step starts: ONE launch of the 2-D templates per plane (stencil2d_kernel
// at 0.9-1.0 of the HBM roofline) instead of the wavefront kernel's plane-per-grid-barrier sweep (26 GLUP/s on 2048^2
// planes).  When the window centres cover the whole plane ('same'-style borders) the source of step n is simply the
// zero-padded output plane n - 1; otherwise (offset scans: the border cells keep their input values) the padded state
// array of the generic scan is kept up to date with one strided copy of the output plane per step.
int try_scan_planemarch(const KxGeom& g, const KxProgram& p, const void* in, const void* coef_dev, void* state, void* out,
                        int reverse, cudaStream_t s) {
  if (getenv("KX_NO_PLANEMARCH") || reverse || coef_dev || g.ndim != 3 || p.n_funcs < 1) return KX_ERR_UNSUPPORTED;
  for (int d = 0; d < 3; ++d)
    if (g.stride[d] != 1 || g.lo[d] < 0 || g.hi[d] < 0) return KX_ERR_UNSUPPORTED;
  const int c0 = (g.ksize[0] - 1) / 2, c1 = (g.ksize[1] - 1) / 2, c2 = (g.ksize[2] - 1) / 2;
  // every function: AFFINE, taps (if any: constants are boundary / initial conditions) in the plane above the centre
  bool any_tap = false;
  for (int fi = 0; fi < p.n_funcs; ++fi) {
    const KxFunc& f = p.func[fi];
    if (f.kind != KX_AFFINE) return KX_ERR_UNSUPPORTED;
    if (g.in_dtype == KX_I32 && !f.coef_is_int) return KX_ERR_UNSUPPORTED;
    for (int t = f.tap_begin; t < f.tap_end; ++t) {
      if (p.tap_off[t][0] != c0 - 1) return KX_ERR_UNSUPPORTED;
      any_tap = true;
    }
  }
  if (!any_tap) return KX_ERR_UNSUPPORTED;
  const int64_t D = g.in_shape[0], H = g.in_shape[1], W = g.in_shape[2];
  const int64_t P0 = D + g.lo[0] + g.hi[0], P1 = H + g.lo[1] + g.hi[1], P2 = W + g.lo[2] + g.hi[2];
  const int64_t o0 = g.out_shape[0], o1 = g.out_shape[1], o2 = g.out_shape[2];
  if (o0 < 1 || o1 < 1 || o2 < 1) return KX_ERR_UNSUPPORTED;
  if (o1 * o2 < 128 * 128) return KX_ERR_UNSUPPORTED;          // tiny planes: two launches per plane cost more than the sweep
  const bool full_cover = o1 == H && o2 == W && g.lo[1] == c1 && g.lo[2] == c2;
  // region keys are window-centre coordinates of the UNPADDED array: the 2-D maps below keep them only when their
  // geometry is the unpadded plane's ('same'-style borders), so meshes take the full-cover form only
  if (p.n_funcs > 1 && !full_cover) return KX_ERR_UNSUPPORTED;
  (void)P0;
  // the 2-D program: every function with its taps projected onto the plane; the keys are filled in per plane
  static thread_local KxProgram p2;
  memset(&p2, 0, sizeof(p2));
  p2.n_funcs = p.n_funcs;
  p2.n_taps = p.n_taps;
  for (int fi = 0; fi < p.n_funcs; ++fi) p2.func[fi] = p.func[fi];
  for (int t = 0; t < p.n_taps; ++t) {
    p2.tap_off[t][0] = p.tap_off[t][1];
    p2.tap_off[t][1] = p.tap_off[t][2];
    p2.coef[t] = p.coef[t];
  }
  // keys of plane n: those whose axis-0 item matches the plane's window-centre coordinate, with that item dropped
  // (utils.py:302-335: a key matches iff every item present matches; zip truncation: absent items match)
  uint32_t cur_mask = 0xffffffffu;
  bool have_mask = false;
  auto set_plane_keys = [&](int64_t n) {
    if (p.n_funcs == 1) return;
    const int64_t cen = n - g.lo[0] + c0;
    uint32_t mask = 0;
    for (int k = 0; k < p.n_keys; ++k) {
      bool m = true;
      if (p.key[k].n_items >= 1) {
        const KxKeyItem& it = p.key[k].item[0];
        if (it.kind == KX_KEY_EQ) m = cen == it.a;
        else if (it.kind == KX_KEY_RANGE) m = it.a <= cen && cen < it.b;
        else if (it.kind == KX_KEY_RANGE_STEP) m = it.a <= cen && cen < it.b && (it.step == 0 ? false : cen % it.step == 0);
      }
      if (m) mask |= 1u << k;
    }
    if (have_mask && mask == cur_mask) return;
    have_mask = true;
    cur_mask = mask;
    p2.n_keys = 0;
    for (int k = 0; k < p.n_keys; ++k) {
      if (!(mask & (1u << k))) continue;
      KxKey& q = p2.key[p2.n_keys++];
      memset(&q, 0, sizeof(q));
      q.func = p.key[k].func;
      q.n_items = p.key[k].n_items > 0 ? p.key[k].n_items - 1 : 0;
      for (int d = 0; d < q.n_items; ++d) q.item[d] = p.key[k].item[d + 1];
    }
  };
  KxGeom g2;
  memset(&g2, 0, sizeof(g2));
  g2.ndim = 2;
  g2.in_dtype = g2.out_dtype = g.in_dtype;
  g2.ksize[0] = g.ksize[1], g2.ksize[1] = g.ksize[2];
  g2.stride[0] = g2.stride[1] = 1;
  g2.out_shape[0] = o1, g2.out_shape[1] = o2;
  const size_t esz = 4;
  char* outp = (char*)out;
  if (full_cover) {
    // source of step n: the zero-padded output plane n - 1 (step 0: input plane c0 - 1 - lo0, or the zero pad)
    g2.in_shape[0] = H, g2.in_shape[1] = W;
    g2.lo[0] = g.lo[1], g2.hi[0] = g.hi[1], g2.lo[1] = g.lo[2], g2.hi[1] = g.hi[2];
    for (int64_t n = 0; n < o0; ++n) {
      const void* src;
      if (n > 0) {
        src = outp + (size_t)(n - 1) * o1 * o2 * esz;
      } else {
        const int64_t pl = c0 - 1 - g.lo[0];
        if (pl >= 0 && pl < D) {
          src = (const char*)in + (size_t)pl * H * W * esz;
        } else {
          KX_CUDA_CHECK(cudaMemsetAsync(state, 0, (size_t)H * W * esz, s));
          src = state;
        }
      }
      set_plane_keys(n);
      MapArgs a;
      fill_map_args(a, g2, p2, src, nullptr, 0, outp + (size_t)n * o1 * o2 * esz, 1);
      const int rc = dispatch_map(a, p2, s);
      if (rc) return rc;
    }
    return KX_OK;
  }
  // general borders: keep the padded state array (pad cells and uncovered cells stay as they are)
  char* st = (char*)state;
  if (P1 != H || P2 != W || P0 != D) KX_CUDA_CHECK(cudaMemsetAsync(state, 0, (size_t)P0 * P1 * P2 * esz, s));
  for (int64_t z = 0; z < D; ++z) {                   // in -> the interior of the padded state (one strided copy per plane
    const int rc = copy2d(st + ((size_t)((z + g.lo[0]) * P1 + g.lo[1]) * P2 + g.lo[2]) * esz, P2,   // unless nothing is padded)
                          (const char*)in + (size_t)z * H * W * esz, W, W, H, s);
    if (rc) return rc;
    if (P1 == H && P2 == W) {                         // contiguous: the rest in one copy
      KX_CUDA_CHECK(cudaMemcpyAsync(st + (size_t)(g.lo[0] + 1) * P1 * P2 * esz, (const char*)in + (size_t)H * W * esz,
                                    (size_t)(D - 1) * H * W * esz, cudaMemcpyDeviceToDevice, s));
      break;
    }
  }
  g2.in_shape[0] = P1, g2.in_shape[1] = P2;          // 'valid' windows over the padded plane
  for (int64_t n = 0; n < o0; ++n) {
    const void* src = st + (size_t)(n + c0 - 1) * P1 * P2 * esz;
    char* dst = outp + (size_t)n * o1 * o2 * esz;
    MapArgs a;
    fill_map_args(a, g2, p2, src, nullptr, 0, dst, 1);
    const int rc = dispatch_map(a, p2, s);
    if (rc) return rc;
    if (n + 1 < o0) {    // the window centres of state plane n + c0: read by the next step
      const int rc2 = copy2d(st + ((size_t)((n + c0) * P1 + c1) * P2 + c2) * esz, P2, dst, o2, o2, o1, s);
      if (rc2) return rc2;
    }
  }
  return KX_OK;
}

// sscan (offset scan: the result is a copy of the input whose offset interior is overwritten, kernex/_src/scan.py:118-147)
// of the same time-marching shape: result plane u = the 2-D offset map of result plane u - 1, fused into ONE pass per
// plane -- stencil2d_kernel in its smap form ('same' geometry, cells outside the offset interior copy the window
// centre) writes straight into the result array, and the ring outside the interior, which must keep the INPUT values
// of plane u rather than those of plane u - 1, is patched by ring_copy_kernel.  8 B per cell, no state array, no
// write-back.  Returns KX_ERR_UNSUPPORTED for anything else (the caller runs kx_scan + the strided write-back).
int scan_offset_planemarch(const KxGeom& g, const KxProgram& p, const void* in, const void* coef_dev, const int32_t* off_lo,
                           void* out_full, cudaStream_t s) {
  if (getenv("KX_NO_PLANEMARCH") || coef_dev || g.ndim != 3 || p.n_funcs != 1) return KX_ERR_UNSUPPORTED;
  const KxFunc& f = p.func[0];
  if (f.kind != KX_AFFINE || f.tap_end <= f.tap_begin || f.post_div != 0.0) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype || (g.in_dtype == KX_I32 && !f.coef_is_int)) return KX_ERR_UNSUPPORTED;
  int c[3], off_hi[3];
  for (int d = 0; d < 3; ++d) {
    c[d] = (g.ksize[d] - 1) / 2;
    if (g.stride[d] != 1 || g.lo[d] != c[d] - off_lo[d] || g.lo[d] < 0 || g.hi[d] < 0) return KX_ERR_UNSUPPORTED;
    off_hi[d] = g.ksize[d] - 1 - c[d] - g.hi[d];                     // windows end off_hi cells before the array does
    if (off_lo[d] < 0 || off_hi[d] < 0) return KX_ERR_UNSUPPORTED;
  }
  if (off_lo[0] < 1) return KX_ERR_UNSUPPORTED;                      // the first written plane reads a real plane
  for (int t = f.tap_begin; t < f.tap_end; ++t)
    if (p.tap_off[t][0] != c[0] - 1) return KX_ERR_UNSUPPORTED;
  const int64_t D = g.in_shape[0], H = g.in_shape[1], W = g.in_shape[2];
  if (H * W < 128 * 128 || H > 0x3fffffff || W > 0x3fffffff) return KX_ERR_UNSUPPORTED;
  const int64_t u0 = off_lo[0], u1 = D - off_hi[0];                  // written planes [u0, u1)
  const int y0 = off_lo[1], y1 = (int)H - off_hi[1], x0 = off_lo[2], x1 = (int)W - off_hi[2];
  if (u1 <= u0 || y1 <= y0 || x1 <= x0) return KX_ERR_UNSUPPORTED;
  static thread_local KxProgram p2;
  memset(&p2, 0, sizeof(p2));
  p2.n_funcs = 1;
  p2.func[0] = f;
  p2.n_taps = f.tap_end - f.tap_begin;
  p2.func[0].tap_begin = 0;
  p2.func[0].tap_end = p2.n_taps;
  for (int t = 0; t < p2.n_taps; ++t) {
    p2.tap_off[t][0] = p.tap_off[f.tap_begin + t][1];
    p2.tap_off[t][1] = p.tap_off[f.tap_begin + t][2];
    p2.coef[t] = p.coef[f.tap_begin + t];
  }
  KxGeom g2;
  memset(&g2, 0, sizeof(g2));
  g2.ndim = 2;
  g2.in_dtype = g2.out_dtype = g.in_dtype;
  for (int d = 0; d < 2; ++d) {
    g2.in_shape[d] = g2.out_shape[d] = g.in_shape[d + 1];
    g2.ksize[d] = g.ksize[d + 1];
    g2.stride[d] = 1;
    g2.lo[d] = c[d + 1];
    g2.hi[d] = g.ksize[d + 1] - 1 - c[d + 1];
  }
  const size_t plane = (size_t)H * W * 4;
  char* res = (char*)out_full;
  const char* src_in = (const char*)in;
  const bool ring = y0 > 0 || x0 > 0 || y1 < H || x1 < W;
  int64_t ring_cells = (int64_t)(y0 + (H - y1)) * W + (int64_t)(y1 - y0) * (x0 + (W - x1));
  for (int64_t u = u0; u < u1; ++u) {
    MapArgs a;
    fill_map_args(a, g2, p2, u == u0 ? (const void*)(src_in + (size_t)(u - 1) * plane) : (const void*)(res + (size_t)(u - 1) * plane),
                  nullptr, 0, res + (size_t)u * plane, 1);
    a.use_box = 1;
    a.box_lo[0] = y0, a.box_hi[0] = y1, a.box_lo[1] = x0, a.box_hi[1] = x1;
    const int rc = try_stencil2d(a, p2, s);
    if (rc) return rc;                               // KX_ERR_UNSUPPORTED at the first plane: nothing was launched yet
    if (ring) {
      int64_t blocks = (ring_cells + 255) / 256;
      if (blocks > 2048) blocks = 2048;
      ring_copy_kernel<<<(unsigned)blocks, 256, 0, s>>>((uint32_t*)(res + (size_t)u * plane), (const uint32_t*)(src_in + (size_t)u * plane),
                                                         (int)H, (int)W, y0, y1, x0, x1);
      const int rc2 = after_launch("ring_copy_kernel");
      if (rc2) return rc2;
    }
  }
  // the planes an offset scan never writes keep the input
  if (u0 > 0) KX_CUDA_CHECK(cudaMemcpyAsync(res, src_in, (size_t)u0 * plane, cudaMemcpyDeviceToDevice, s));
  if (u1 < D) KX_CUDA_CHECK(cudaMemcpyAsync(res + (size_t)u1 * plane, src_in + (size_t)u1 * plane, (size_t)(D - u1) * plane, cudaMemcpyDeviceToDevice, s));
  return KX_OK;
}

__global__ void __launch_bounds__(256) cast_i32_f32_kernel(const int32_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t nv = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nv; i += (int64_t)gridDim.x * 256) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(in) + i);
    __stcs(reinterpret_cast<float4*>(out) + i, make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w));
  }
  for (int64_t i = (nv << 2) + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) out[i] = (float)in[i];
}

extern "C" {

// int32 -> float32, round to nearest even like XLA's convert: the first thing JAX does to an int32 window when the
// function's result is float (jnp.mean of integers, a Python-float coefficient): lets those calls take the float32 fast
// templates instead of the generic kernel (engine.py _run_map).  in / out: device, 16-byte aligned, n elements.
int kx_cast_i32_f32(const void* in, void* out, int64_t n, void* stream) {
  if (n <= 0) return KX_OK;
  if (!in || !out) return fail(KX_ERR_INVALID, "null device pointer");
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(KX_ERR_INVALID, "unaligned buffer");
  DeviceGuard guard(out);
  int64_t blocks = ((n >> 2) + 255) / 256 + 1;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cast_i32_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const int32_t*)in, (float*)out, n);
  return after_launch("cast_i32_f32_kernel");
}

int kx_abi_version(void) { return KX_ABI_VERSION; }
const char* kx_last_error(void) { return err_buf(); }
uint64_t kx_launch_count(void) { return launch_counter().load(); }
const char* kx_last_kernel(void) { return last_kernel_name().load(std::memory_order_relaxed); }

int kx_map(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev, int64_t coef_batch_stride,
           void* out, int64_t batch, void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (batch < 1) return fail(KX_ERR_INVALID, "batch=%lld", (long long)batch);
  if (!in || !out) return fail(KX_ERR_INVALID, "null device pointer");
  DeviceGuard guard(out);
  MapArgs a;
  fill_map_args(a, *geom, *prog, in, coef_dev, coef_batch_stride, out, batch);
  if (a.n_win == 0) return KX_OK;
  return dispatch_map(a, *prog, (cudaStream_t)stream);
}

int kx_map_offset(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev,
                  const int32_t* off0, void* out, void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (!in || !out || !off0) return fail(KX_ERR_INVALID, "null pointer");
  if (geom->in_dtype != geom->out_dtype)
    return fail(KX_ERR_INVALID, "offset maps write into a copy of the input: dtypes must agree");
  for (int f = 0; f < prog->n_funcs; ++f)
    if (prog->func[f].kind == KX_GATHER && prog->func[f].tap_end - prog->func[f].tap_begin != 1)
      return fail(KX_ERR_INVALID, "offset map output must be scalar");
  DeviceGuard guard(out);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t in_elems = prod(geom->in_shape, geom->ndim);
  if (in_elems <= 0) return KX_OK;
  // Fused pass: with stride 1 the offset map is a 'same'-padded map over EVERY cell in which the
  // cells outside [off0, off0 + out_shape) keep their input value (window centre), so the copy of
  // the input and the strided write-back disappear (8 B per cell instead of 16).
  bool geo_ok = true;                     // stride 1 and borders that are exactly what the offsets imply
  for (int d = 0; d < geom->ndim; ++d)
    if (geom->stride[d] != 1 || geom->lo[d] != (geom->ksize[d] - 1) / 2 - off0[d]) geo_ok = false;
  const bool unit = geo_ok && prog->n_funcs == 1 && prog->func[0].kind == KX_AFFINE;   // the one-pass smap form of stencil2d
  if (geo_ok) {
    KxGeom gs = *geom;
    MapArgs a;
    for (int d = 0; d < geom->ndim; ++d) {
      gs.lo[d] = (geom->ksize[d] - 1) / 2;
      gs.hi[d] = geom->ksize[d] / 2;
      gs.out_shape[d] = geom->in_shape[d];
    }
    fill_map_args(a, gs, *prog, in, coef_dev, 0, out, 1);
    a.use_box = 1;
    for (int d = 0; d < geom->ndim; ++d) {
      a.box_lo[d] = off0[d];
      a.box_hi[d] = off0[d] + geom->out_shape[d];
    }
    rc = unit ? try_stencil2d(a, *prog, s) : KX_ERR_UNSUPPORTED;
    if (rc != KX_ERR_UNSUPPORTED) return rc;
    // Other windows and programs (3-D stencils: a Jacobi step written as smap; window reductions; `F[slice] = fn` meshes): the SAME map over every cell on whatever fast
    // template takes it -- the cells outside the offset interior come out as zero-padded window values -- and the
    // ring outside the interior is then restored from the input, slab by slab.  8 B per cell + the ring instead of
    // copy + generic kernel + strided write-back (0.05 of the HBM roofline).
    if (!getenv("KX_NO_FUSED_OFFSET") && in_elems >= (1 << 16)) {
      MapArgs a2;
      fill_map_args(a2, gs, *prog, in, coef_dev, 0, out, 1);
      rc = dispatch_map(a2, *prog, s);
      if (rc) return rc;
      int64_t blo[KX_MAX_DIMS], bhi[KX_MAX_DIMS];
      for (int d = 0; d < geom->ndim; ++d) blo[d] = off0[d], bhi[d] = off0[d] + geom->out_shape[d];
      return copy_outside_box(out, in, geom->in_shape, geom->ndim, blo, bhi, s);
    }
  }
  KX_CUDA_CHECK(cudaMemcpyAsync(out, in, (size_t)in_elems * 4, cudaMemcpyDeviceToDevice, s));
  MapArgs a;
  fill_map_args(a, *geom, *prog, in, coef_dev, 0, out, 1);
  if (a.n_win == 0) return KX_OK;
  // window j of axis d lands at input coordinate off0[d] + j*stride[d] of the copy
  int64_t pitch = 1;
  a.out_base = 0;
  for (int d = geom->ndim - 1; d >= 0; --d) {
    a.out_pitch[d] = pitch * geom->stride[d];
    a.out_base += pitch * off0[d];
    if (off0[d] < 0 || off0[d] + (geom->out_shape[d] - 1) * (int64_t)geom->stride[d] >= geom->in_shape[d])
      return fail(KX_ERR_INVALID, "axis %d: offset write-back leaves the array", d);
    pitch *= geom->in_shape[d];
  }
  a.fn_elems = 1;
  return launch_generic_map(a, *prog, s);
}

int kx_map_vjp_input(const KxGeom* geom, const KxProgram* prog, const void* g, const void* coef_dev, void* dx,
                     void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (prog->n_funcs != 1 || prog->func[0].kind != KX_AFFINE || geom->in_dtype != KX_F32 || geom->out_dtype != KX_F32)
    return fail(KX_ERR_UNSUPPORTED, "VJP is implemented for single-function float32 AFFINE programs");
  if (!g || !dx) return fail(KX_ERR_INVALID, "null device pointer");
  DeviceGuard guard(dx);
  {
    int C = 0, NT = 0;
    if (full_depth_conv(*geom, *prog, C, NT) && (C > 4 || NT != 9)) {   // C <= 4 with 3x3 windows: the conv row ring takes the transposed program
      rc = vjp_input_full_depth(*geom, *prog, C, NT, g, coef_dev, dx, (cudaStream_t)stream);
      if (rc != KX_ERR_UNSUPPORTED) return rc;
    }
  }
  static thread_local KxGeom gt;
  static thread_local KxProgram pt;
  const int first = prog->func[0].tap_begin;
  if (transpose_program(*geom, *prog, gt, pt) && prod(gt.out_shape, gt.ndim) > 0 && prod(gt.in_shape, gt.ndim) > 0) {
    MapArgs a;
    const void* cd = coef_dev ? (const void*)((const float*)coef_dev + first) : nullptr;
    fill_map_args(a, gt, pt, g, cd, 0, dx, 1);
    return dispatch_map(a, pt, (cudaStream_t)stream);
  }
  return launch_generic_vjp_input(*geom, *prog, g, coef_dev, dx, (cudaStream_t)stream);
}

int64_t kx_vjp_coef_scratch_bytes(const KxGeom* geom, const KxProgram* prog) {
  if (!geom || !prog) return 0;
  int64_t a = generic_vjp_coef_scratch_bytes(*geom, *prog), b = conv_wgrad_scratch_bytes(*geom, *prog);
  {
    const int64_t c = wgrad_rows_scratch_bytes(*geom, *prog);
    a = c > a ? c : a;
  }
  int C = 0, NT = 0;
  if (prog->n_funcs == 1 && full_depth_conv(*geom, *prog, C, NT) && C > 4 && NT == 9) {   // groups of 4 channels, one at a time
    KxGeom gs = *geom;
    gs.in_shape[0] = gs.ksize[0] = 4;
    static thread_local KxProgram ps;
    memset(&ps, 0, sizeof(ps));
    ps.n_funcs = 1;
    ps.n_taps = 36;
    ps.func[0] = prog->func[0];
    ps.func[0].tap_begin = 0;
    ps.func[0].tap_end = 36;
    const int64_t c = conv_wgrad_scratch_bytes(gs, ps);
    b = c > b ? c : b;
  }
  return a > b ? a : b;
}

int kx_map_vjp_coef(const KxGeom* geom, const KxProgram* prog, const void* x, const void* g, void* scratch,
                    void* dcoef, void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (prog->n_funcs != 1 || prog->func[0].kind != KX_AFFINE || geom->in_dtype != KX_F32 || geom->out_dtype != KX_F32)
    return fail(KX_ERR_UNSUPPORTED, "VJP is implemented for single-function float32 AFFINE programs");
  if (!x || !g || !scratch || !dcoef) return fail(KX_ERR_INVALID, "null device pointer");
  DeviceGuard guard(dcoef);
  rc = try_conv_wgrad(*geom, *prog, x, g, scratch, dcoef, (cudaStream_t)stream);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  {
    // full-depth conv with more than 4 channels: the fused window-correlation kernel per group of <= 4 channels
    // (the cotangent plane is re-read once per group: (4 C + C) / (4 C + 4) of the ideal traffic)
    int C = 0, NT = 0;
    if (full_depth_conv(*geom, *prog, C, NT) && C > 4 && NT == 9) {
      static thread_local KxGeom gs;
      static thread_local KxProgram ps;
      const int first = prog->func[0].tap_begin;
      const int64_t plane = geom->in_shape[1] * geom->in_shape[2];
      (void)first;
      for (int c0 = 0; c0 < C; c0 += 4) {
        const int cn = C - c0 < 4 ? C - c0 : 4;
        gs = *geom;
        gs.in_shape[0] = gs.ksize[0] = cn;
        memset(&ps, 0, sizeof(ps));
        ps.n_funcs = 1;
        ps.n_taps = cn * 9;
        ps.func[0] = prog->func[0];
        ps.func[0].tap_begin = 0;
        ps.func[0].tap_end = cn * 9;
        for (int t = 0; t < cn * 9; ++t) {
          ps.tap_off[t][0] = (int16_t)(t / 9);
          ps.tap_off[t][1] = prog->tap_off[first + c0 * 9 + t][1];
          ps.tap_off[t][2] = prog->tap_off[first + c0 * 9 + t][2];
        }
        rc = try_conv_wgrad(gs, ps, (const float*)x + (int64_t)c0 * plane, g, scratch, (float*)dcoef + c0 * 9,
                            (cudaStream_t)stream);
        if (rc) break;
      }
      if (rc != KX_ERR_UNSUPPORTED) return rc;
    }
  }
  rc = try_wgrad_rows(*geom, *prog, x, g, scratch, dcoef, (cudaStream_t)stream);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  return launch_generic_vjp_coef(*geom, *prog, x, g, scratch, dcoef, (cudaStream_t)stream);
}

int64_t kx_scan_state_bytes(const KxGeom* geom) {
  if (!geom || geom->ndim < 1 || geom->ndim > KX_MAX_DIMS) return 0;
  return scan_state_elems(*geom) * 4;
}

int kx_scan(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev, void* state, void* out,
            int32_t reverse, void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (geom->in_dtype != geom->out_dtype) return fail(KX_ERR_INVALID, "kscan returns the carried dtype");
  for (int f = 0; f < prog->n_funcs; ++f)
    if (prog->func[f].kind == KX_GATHER && prog->func[f].tap_end - prog->func[f].tap_begin != 1)
      return fail(KX_ERR_INVALID, "scan operation output must be scalar");
  if (!in || !out || !state) return fail(KX_ERR_INVALID, "null device pointer");
  DeviceGuard guard(out);
  cudaStream_t s = (cudaStream_t)stream;
  rc = try_scan_rowmarch(*geom, *prog, in, coef_dev, out, reverse, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  rc = try_scan_planemarch(*geom, *prog, in, coef_dev, state, out, reverse, s);
  if (rc != KX_ERR_UNSUPPORTED) return rc;
  if (!getenv("KX_SCAN_SEQUENTIAL")) {  // cross-check switch: the one-thread sweep of kx_generic.cu
    rc = launch_scan_wavefront(*geom, *prog, in, coef_dev, state, out, reverse, s);
    if (rc != KX_ERR_UNSUPPORTED) return rc;
  }
  return launch_generic_scan(*geom, *prog, in, coef_dev, state, out, reverse, s);
}

int kx_scan_offset(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev, const int32_t* off0,
                   void* out, void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (!in || !out || !off0) return fail(KX_ERR_INVALID, "null pointer");
  DeviceGuard guard(out);
  rc = scan_offset_planemarch(*geom, *prog, in, coef_dev, off0, out, (cudaStream_t)stream);
  if (rc == KX_ERR_UNSUPPORTED) return fail(KX_ERR_UNSUPPORTED, "no fused offset-scan template: run kx_scan and write the values back");
  return rc;
}

int kx_scan_rowmarch_shard(const KxGeom* geom, const KxProgram* prog, int64_t col_base, int64_t global_nx, void* out,
                           void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (!out) return fail(KX_ERR_INVALID, "null device pointer");
  if (geom->ndim != 2 || col_base + geom->in_shape[1] > global_nx + (int64_t)geom->in_shape[0] * 4 || global_nx < 1)
    return fail(KX_ERR_INVALID, "shard [%lld, %lld) does not belong to a scan of width %lld", (long long)col_base,
                (long long)(col_base + (geom->ndim == 2 ? geom->in_shape[1] : 0)), (long long)global_nx);
  DeviceGuard guard(out);
  rc = scan_rowmarch_shard(*geom, *prog, nullptr, out, 0, col_base, global_nx, (cudaStream_t)stream, 0, -1);
  if (rc == KX_ERR_UNSUPPORTED)
    return fail(KX_ERR_UNSUPPORTED,
                "row-march sharding needs a 2-D float32 kscan with 'same'-style borders whose functions read only "
                "the row above");
  return rc;
}

int kx_scan_rowmarch_rows_per_launch(void) { return scan_rowmarch_rows_per_launch(); }

int kx_scan_rowmarch_shard_rows(const KxGeom* geom, const KxProgram* prog, int64_t col_base, int64_t global_nx,
                                int32_t row_begin, int32_t row_end, void* out, void* stream) {
  int rc = validate(geom, prog);
  if (rc) return rc;
  if (!out) return fail(KX_ERR_INVALID, "null device pointer");
  if (geom->ndim != 2 || global_nx < 1 || row_begin < 0 || row_end > geom->in_shape[0] || row_begin > row_end)
    return fail(KX_ERR_INVALID, "rows [%d, %d) of a %lld-row shard", row_begin, row_end,
                (long long)(geom->ndim == 2 ? geom->in_shape[0] : 0));
  DeviceGuard guard(out);
  rc = scan_rowmarch_shard(*geom, *prog, nullptr, out, 0, col_base, global_nx, (cudaStream_t)stream, row_begin, row_end);
  if (rc == KX_ERR_UNSUPPORTED)
    return fail(KX_ERR_UNSUPPORTED,
                "row-march sharding needs a 2-D float32 kscan with 'same'-style borders whose functions read only "
                "the row above");
  return rc;
}

}  // extern "C"
