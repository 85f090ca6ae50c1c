// Host-side planning of the stencil2d fast path (kernel: kx_stencil2d_impl.cuh, instantiations: kx_stencil2d_k*.cu).
#include <cmath>

#include "kx_stencil2d_impl.cuh"

namespace kx {

namespace {

using namespace s2;

template <typename T>
int launch_t(const S2Args<T>& a, int ky, int kx, int op, int batch, int vw, cudaStream_t s) {
  switch (ky) {
    case 1: return s2_launch_k1(a, kx, op, batch, vw, s);
    case 2: return s2_launch_k2(a, kx, op, batch, vw, s);
    case 3: return s2_launch_k3(a, kx, op, batch, vw, s);
    case 4: return s2_launch_k4(a, kx, op, batch, vw, s);
    case 5: return s2_launch_k5(a, kx, op, batch, vw, s);
    case 7: return s2_launch_k7(a, kx, op, batch, vw, s);
    case 9: return s2_launch_k9(a, kx, op, batch, vw, s);
    case 11: return s2_launch_k11(a, kx, op, batch, vw, s);
    case 13: return s2_launch_k13(a, kx, op, batch, vw, s);
    case 15: return s2_launch_k15(a, kx, op, batch, vw, s);
  }
  return KX_ERR_UNSUPPORTED;
}

template <typename T>
int run(const MapArgs& m, const KxProgram& p, int ay, int ax, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int KY = ay >= 0 ? g.ksize[ay] : 1, KX = g.ksize[ax];
  S2Args<T> a;
  memset(&a, 0, sizeof(a));
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.hin = ay >= 0 ? (int)g.in_shape[ay] : 1;
  a.win = (int)g.in_shape[ax];
  a.hout = ay >= 0 ? (int)g.out_shape[ay] : 1;
  a.wout = (int)g.out_shape[ax];
  a.ly = ay >= 0 ? g.lo[ay] : 0;
  a.lx = g.lo[ax];
  a.in_batch = m.in_bcast ? 0 : (int64_t)a.hin * a.win;   // in_bcast: every batch item reads the SAME input
  a.out_batch = (int64_t)a.hout * a.wout;
  a.bias = (T)f.bias;
  if (m.use_box) {
    a.use_box = 1;
    a.by0 = ay >= 0 ? (int)m.box_lo[ay] : 0;
    a.by1 = ay >= 0 ? (int)m.box_hi[ay] : 1;
    a.bx0 = (int)m.box_lo[ax];
    a.bx1 = (int)m.box_hi[ax];
  }
  // band height, measured on C2 (16384^2, 3x3): 4 rows 762 GLUP/s, 8 822, 12 824 (VJP 834), 16 802, 32 776, 48 761:
  // short bands keep the concurrently active blocks on few DRAM pages and their halo re-reads in L2
  a.th = a.hout >= 512 ? 12 : (a.hout >= 64 ? 16 : 8);
  if (const char* e = getenv("KX_S2_TH")) a.th = atoi(e) > 0 ? atoi(e) : a.th;   // tuning knob
  while ((a.hout + a.th - 1) / a.th > 65535) a.th *= 2;
  const bool dev = m.coef_dev != nullptr;
  T dense[kMaxK * kMaxK];
  for (int i = 0; i < kMaxK * kMaxK; ++i) dense[i] = T(0);
  // device tables are indexed by tap slot; the kernel wants a dense KY*KX table: only
  // usable directly when the program lists every window element in row-major order
  bool dense_order = (f.tap_end - f.tap_begin) == KY * KX, rev_order = dense_order;
  bool seen[kMaxK * kMaxK] = {false};
  int n_seen = 0;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int oy = ay >= 0 ? p.tap_off[t][ay] : 0, ox = p.tap_off[t][ax];
    const int slot = oy * KX + ox;
    if (seen[slot]) return KX_ERR_UNSUPPORTED;  // duplicate tap: leave it to the generic kernel
    seen[slot] = true;
    ++n_seen;
    if (slot < 32) a.mask |= 1u << slot;
    else if (slot < 64) a.mask_hi |= 1u << (slot - 32);
    dense[slot] = (T)p.coef[t];
    if (slot != t - f.tap_begin) dense_order = false;
    if (slot != KY * KX - 1 - (t - f.tap_begin)) rev_order = false;   // a transposed program: flipped taps, same table
  }
  const bool full = n_seen == KY * KX;
  if (dev) {
    if (!dense_order && !rev_order) return KX_ERR_UNSUPPORTED;
    a.coef_rev = dense_order ? 0 : 1;
    a.coef_dev = (const T*)m.coef_dev;
    a.coef_batch = m.coef_batch_stride;
  }
  for (int i = 0; i < KY * KX; ++i) a.coef[i] = dense[i];   // (kSep overwrites its two vectors below)
  const uintptr_t base = reinterpret_cast<uintptr_t>(m.in);
  int vw = 1;
  if (a.win % 4 == 0 && (base & 15) == 0 && a.in_batch % 4 == 0) vw = 4;
  else if (a.win % 2 == 0 && (base & 7) == 0 && a.in_batch % 2 == 0) vw = 2;
  int op = kAffine;
  if (f.kind != KX_AFFINE) {
    // whole-window reductions only (np.max(x) / np.min(x)); subsets stay on the generic kernel
    if (dev || !full || m.use_box) return KX_ERR_UNSUPPORTED;
    op = f.kind == KX_REDUCE_MAX ? kMax : kMin;
  } else if (!dev && !m.use_box && full && !getenv("KX_S2_NO_SEP")) {
    bool same = true;   // np.sum / np.mean / box blurs: one common coefficient -> separable row / column sums
    for (int i = 1; i < KY * KX; ++i) same = same && (dense[i] == dense[0]);
    if (same && s2_has_box(KY, KX)) op = kBox;
    else if (std::is_same<T, float>::value && s2_has_sep(KY, KX)) {
      // outer-product tables (Gaussian blur = outer(w, w) / sum, README.md:281-284): w[ky][kx] = cy[ky] * cx[kx] with
      // cy = column / pivot and cx = row through the largest entry, accepted when every entry is reproduced to
      // 4e-7 relative (fp32 rounding of the user's own products); the result then differs from the dense form
      // by <= ~5e-7 * sum|c x|, well inside the stated 1e-5 tolerance
      int pi = 0;
      for (int i = 1; i < KY * KX; ++i)
        if (fabs((double)dense[i]) > fabs((double)dense[pi])) pi = i;
      const int py = pi / KX, px = pi % KX;
      const double piv = (double)dense[pi];
      bool sep = piv != 0.0;
      for (int ky = 0; ky < KY && sep; ++ky)
        for (int kx = 0; kx < KX && sep; ++kx) {
          const double w = (double)dense[ky * KX + kx];
          const double r = (double)dense[ky * KX + px] / piv * (double)dense[py * KX + kx];
          sep = fabs(w - r) <= 4e-7 * fabs(w) + 1e-37;
        }
      if (sep) {
        op = kSep;
        for (int ky = 0; ky < KY; ++ky) a.coef[ky] = (T)((double)dense[ky * KX + px] / piv);
        for (int kx = 0; kx < KX; ++kx) a.coef[kMaxK + kx] = dense[py * KX + kx];
      }
    }
  }
  if (op == kAffine && !s2_has_affine(KY, KX)) return KX_ERR_UNSUPPORTED;   // dense form: windows up to 7x7, 1 x 15
  a.post_mul = f.post_div != 0.0 ? (float)(1.0 / f.post_div) : 0.f;
  if constexpr (std::is_same<T, float>::value) {
    if (op == kSep && vw == 4) {   // square odd windows >= 7x7: the TMA row-ring kernel (kx_sep2d.cu)
      Sep2dReq r;
      memset(&r, 0, sizeof(r));
      r.in = a.in;
      r.out = a.out;
      r.in_batch = a.in_batch;
      r.out_batch = a.out_batch;
      r.batch = batch;
      r.hin = a.hin, r.win = a.win, r.hout = a.hout, r.wout = a.wout, r.ly = a.ly, r.lx = a.lx;
      r.KY = KY, r.KX = KX;
      r.bias = a.bias;
      r.post_mul = a.post_mul;
      for (int ky = 0; ky < KY && ky < 15; ++ky) r.cy[ky] = a.coef[ky];
      for (int kx = 0; kx < KX && kx < 15; ++kx) r.cx[kx] = a.coef[kMaxK + kx];
      const int rc = try_sep2d(r, s);
      if (rc != KX_ERR_UNSUPPORTED) return rc;
    }
  }
  return launch_t<T>(a, KY, KX, op, (int)batch, vw, s);
}

}  // namespace

int try_stencil2d(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (p.n_funcs != 1 || p.func[0].kind == KX_GATHER) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype) return KX_ERR_UNSUPPORTED;
  if (p.func[0].post_div != 0.0 && (p.func[0].kind != KX_AFFINE || g.in_dtype != KX_F32)) return KX_ERR_UNSUPPORTED;
  if (m.fn_elems != 1) return KX_ERR_UNSUPPORTED;
  if (getenv("KX_S2_LEGACY_SHAPES")) {   // A/B switch: only what the first version covered (the rest: pool2d / generic)
    const int ky = g.ndim >= 2 ? g.ksize[g.ndim - 2] : 1, kx = g.ksize[g.ndim - 1];
    const bool old = (ky == 1 && (kx <= 3 || kx == 5)) || (ky == 2 && kx <= 3) || (ky == 3 && kx <= 3) ||
                     (ky == 5 && (kx == 1 || kx == 5));
    if (!old || p.func[0].kind != KX_AFFINE || p.func[0].post_div != 0.0) return KX_ERR_UNSUPPORTED;
  }
  const int nd = g.ndim;
  const int ax = nd - 1, ay = nd - 2;  // ay = -1 for 1-D arrays
  // leading axes must be pure batch axes: window 1, stride 1, no border
  int64_t lead = 1;
  for (int d = 0; d < nd - 2; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    if (m.use_box && (m.box_lo[d] != 0 || m.box_hi[d] != g.in_shape[d])) return KX_ERR_UNSUPPORTED;
    lead *= g.in_shape[d];
  }
  if (g.stride[ax] != 1 || (ay >= 0 && g.stride[ay] != 1)) return KX_ERR_UNSUPPORTED;
  if (g.ksize[ax] > kMaxK || (ay >= 0 && g.ksize[ay] > kMaxK)) return KX_ERR_UNSUPPORTED;
  for (int d = 0; d < nd; ++d)
    if (g.in_shape[d] > 0x3fffffff || g.out_shape[d] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
  if (m.coef_dev && lead > 1 && m.coef_batch_stride != 0) return KX_ERR_UNSUPPORTED;
  const int64_t batch = m.batch * lead;
  if (batch > 65535) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype == KX_F32) return run<float>(m, p, ay, ax, batch, s);
  if (g.in_dtype == KX_I32) {
    if (p.func[0].kind == KX_AFFINE && !p.func[0].coef_is_int) return KX_ERR_UNSUPPORTED;
    return run<int32_t>(m, p, ay, ax, batch, s);
  }
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
