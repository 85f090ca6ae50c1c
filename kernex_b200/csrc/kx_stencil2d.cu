// Fast path: single-function AFFINE stencils over the last two axes, stride 1, f32 / i32.
// (2-D 5-point Laplacian, 3x3 blur / conv, 1-D windows as H = 1, leading axes with a
// 1-wide window folded into the batch.)
//
// HBM-bound design (8 B per lattice update = read once + write once):
//   * each thread owns 4 consecutive output columns and MARCHES down the rows of its
//     block's band, keeping the last KY input rows in a register queue, so every input
//     row is fetched once per thread column; the x-halo and the row halo between bands
//     are re-read through L1 / L2, never from DRAM;
//   * input rows are fetched as 128-bit aligned vectors (ld.global.nc.v4); the window
//     misalignment ((-lo_x) mod 4) is a template parameter so the realignment is pure
//     register renaming; U rows are fetched back to back before any arithmetic so each
//     warp keeps U*NV 512-byte requests in flight;
//   * zero padding (kernex/_src/utils.py:43-53) is a predicate on whole vectors, never a
//     materialised pad; negative borders just move the window origin;
//   * coefficients sit in kernel-parameter constant memory (FFMA constant operands) or,
//     for device weights, in registers; a tap mask skips taps the user never wrote so
//     non-finite inputs cannot leak through a 0 * x product.
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kTX = 128;      // threads per block, each owning 4 output columns
constexpr int kCols = 4;
constexpr int kMaxK = 5;

template <typename T>
struct S2Args {
  const T* in;
  T* out;
  const T* coef_dev;           // optional [KY*KX] device table (dense, row-major)
  int64_t in_batch, out_batch, coef_batch;
  int hin, win, hout, wout;
  int ly, lx;                  // left borders
  int th;                      // output rows per block
  uint32_t mask;               // bit ky*KX+kx set = tap present
  int use_box, by0, by1, bx0, bx1;  // offset maps: outputs outside the box copy the window centre
  T bias;
  T coef[kMaxK * kMaxK];
};

template <typename T>
struct Vec4 {
  T v[4];
};

template <typename T>
__device__ __forceinline__ Vec4<T> ld_vec4(const T* p) {
  Vec4<T> r;
  const int4 raw = __ldg(reinterpret_cast<const int4*>(p));
  r.v[0] = reinterpret_cast<const T&>(raw.x);
  r.v[1] = reinterpret_cast<const T&>(raw.y);
  r.v[2] = reinterpret_cast<const T&>(raw.z);
  r.v[3] = reinterpret_cast<const T&>(raw.w);
  return r;
}

template <typename T>
__device__ __forceinline__ void st_row4(T* p, const T (&a)[4], int n_valid) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
  if (n_valid == 4 && (addr & 15) == 0) {
    int4 raw;
    raw.x = reinterpret_cast<const int&>(a[0]);
    raw.y = reinterpret_cast<const int&>(a[1]);
    raw.z = reinterpret_cast<const int&>(a[2]);
    raw.w = reinterpret_cast<const int&>(a[3]);
    __stcs(reinterpret_cast<int4*>(p), raw);
  } else if (n_valid == 4 && (addr & 7) == 0) {
    int2 lo, hi;
    lo.x = reinterpret_cast<const int&>(a[0]);
    lo.y = reinterpret_cast<const int&>(a[1]);
    hi.x = reinterpret_cast<const int&>(a[2]);
    hi.y = reinterpret_cast<const int&>(a[3]);
    __stcs(reinterpret_cast<int2*>(p), lo);
    __stcs(reinterpret_cast<int2*>(p) + 1, hi);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n_valid) p[i] = a[i];
  }
}

// VW = elements per load: 4 when the input rows are 16-byte aligned, 2 when they are 8-byte aligned
// (even widths such as the 16382-wide cotangent of the C2 backward pass), else 1.
// SHIFT = (-lx) mod VW: misalignment of the window inside the thread's aligned row buffer.
template <typename T, int KY, int KX, int SHIFT, int VW, int U>
__global__ void __launch_bounds__(kTX)
stencil2d_kernel(const __grid_constant__ S2Args<T> a) {
  constexpr int NV = (SHIFT + KX + kCols - 1 + VW - 1) / VW;  // aligned vectors per row per thread
  constexpr int NE = NV * VW;
  const int j0 = (blockIdx.x * kTX + threadIdx.x) * kCols;  // first output column
  if (j0 >= a.wout) return;
  const int r0 = blockIdx.y * a.th;
  const int r1 = min(r0 + a.th, a.hout);
  const T* __restrict__ in = a.in + (int64_t)blockIdx.z * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)blockIdx.z * a.out_batch;
  const int xa = j0 - a.lx - SHIFT;  // column of element 0 of the thread's row buffer
  const int n_valid = min(kCols, a.wout - j0);

  T coef[KY * KX];
  if (a.coef_dev) {
    const T* cd = a.coef_dev + (int64_t)blockIdx.z * a.coef_batch;
#pragma unroll
    for (int i = 0; i < KY * KX; ++i) coef[i] = cd[i];
  } else {
#pragma unroll
    for (int i = 0; i < KY * KX; ++i) coef[i] = a.coef[i];
  }

  auto load_row = [&](int iy, T(&dst)[NE]) {
    const bool row_ok = (iy >= 0) && (iy < a.hin);
    const T* rp = in + (int64_t)iy * a.win;
    if (VW == 4) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = xa + 4 * v;
        if (row_ok && c >= 0 && c + 3 < a.win) {  // win % 4 == 0: a vector is all in or all out
          const Vec4<T> t = ld_vec4(rp + c);
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[4 * v + e] = t.v[e];
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[4 * v + e] = T(0);
        }
      }
    } else if (VW == 2) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = xa + 2 * v;
        if (row_ok && c >= 0 && c + 1 < a.win) {  // win % 2 == 0: a pair is all in or all out
          const int2 raw = __ldg(reinterpret_cast<const int2*>(rp + c));
          dst[2 * v] = reinterpret_cast<const T&>(raw.x);
          dst[2 * v + 1] = reinterpret_cast<const T&>(raw.y);
        } else {
          dst[2 * v] = T(0);
          dst[2 * v + 1] = T(0);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < KX + kCols - 1; ++e) {
        const int c = xa + e;
        dst[e] = (row_ok && c >= 0 && c < a.win) ? __ldg(rp + c) : T(0);
      }
    }
  };

  T q[KY - 1 + U][NE];
#pragma unroll
  for (int ky = 0; ky < KY - 1; ++ky) load_row(r0 - a.ly + ky, q[ky]);

  for (int r = r0; r < r1; r += U) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (r + u < r1) load_row(r + u - a.ly + KY - 1, q[KY - 1 + u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u < r1) {
        T acc[kCols];
#pragma unroll
        for (int i = 0; i < kCols; ++i) acc[i] = a.bias;
#pragma unroll
        for (int ky = 0; ky < KY; ++ky) {
#pragma unroll
          for (int kx = 0; kx < KX; ++kx) {
            if (a.mask & (1u << (ky * KX + kx))) {
#pragma unroll
              for (int i = 0; i < kCols; ++i)
                acc[i] = mad_wrap(coef[ky * KX + kx], q[u + ky][SHIFT + kx + i], acc[i]);
            }
          }
        }
        if (a.use_box) {  // block-uniform: smap / offset_kernel_map fused into one pass
          const bool rin = (r + u >= a.by0) && (r + u < a.by1);
#pragma unroll
          for (int i = 0; i < kCols; ++i)
            if (!(rin && j0 + i >= a.bx0 && j0 + i < a.bx1)) acc[i] = q[u + (KY - 1) / 2][SHIFT + (KX - 1) / 2 + i];
        }
        st_row4(out + (int64_t)(r + u) * a.wout + j0, acc, n_valid);
      }
    }
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
      for (int e = 0; e < NE; ++e) q[ky][e] = q[U + ky][e];
    }
  }
}

template <typename T, int KY, int KX>
int launch_kk(const S2Args<T>& a, int batch, int vw, cudaStream_t s) {
  constexpr int U = (KY * KX >= 15) ? 2 : 4;
  dim3 grid((a.wout + kTX * kCols - 1) / (kTX * kCols), (a.hout + a.th - 1) / a.th, batch);
  const int shift = vw > 1 ? (((-a.lx) % vw) + vw) % vw : 0;
  if (vw == 1) stencil2d_kernel<T, KY, KX, 0, 1, U><<<grid, kTX, 0, s>>>(a);
  else if (vw == 2 && shift == 0) stencil2d_kernel<T, KY, KX, 0, 2, U><<<grid, kTX, 0, s>>>(a);
  else if (vw == 2) stencil2d_kernel<T, KY, KX, 1, 2, U><<<grid, kTX, 0, s>>>(a);
  else if (shift == 0) stencil2d_kernel<T, KY, KX, 0, 4, U><<<grid, kTX, 0, s>>>(a);
  else if (shift == 1) stencil2d_kernel<T, KY, KX, 1, 4, U><<<grid, kTX, 0, s>>>(a);
  else if (shift == 2) stencil2d_kernel<T, KY, KX, 2, 4, U><<<grid, kTX, 0, s>>>(a);
  else stencil2d_kernel<T, KY, KX, 3, 4, U><<<grid, kTX, 0, s>>>(a);
  return after_launch("stencil2d_kernel");
}

template <typename T>
int launch_t(const S2Args<T>& a, int ky, int kx, int batch, int vw, cudaStream_t s) {
#define KX_CASE(Y, X) \
  if (ky == Y && kx == X) return launch_kk<T, Y, X>(a, batch, vw, s);
  KX_CASE(1, 1) KX_CASE(1, 2) KX_CASE(1, 3) KX_CASE(1, 5)
  KX_CASE(2, 1) KX_CASE(2, 2) KX_CASE(2, 3)
  KX_CASE(3, 1) KX_CASE(3, 2) KX_CASE(3, 3)
  KX_CASE(5, 1) KX_CASE(5, 5)
#undef KX_CASE
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
  a.in_batch = (int64_t)a.hin * a.win;
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
  bool dense_order = (f.tap_end - f.tap_begin) == KY * KX;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int oy = ay >= 0 ? p.tap_off[t][ay] : 0, ox = p.tap_off[t][ax];
    const int slot = oy * KX + ox;
    if (a.mask & (1u << slot)) return KX_ERR_UNSUPPORTED;  // duplicate tap: leave it to the generic kernel
    a.mask |= 1u << slot;
    dense[slot] = (T)p.coef[t];
    if (slot != t - f.tap_begin) dense_order = false;
  }
  if (dev) {
    if (!dense_order) return KX_ERR_UNSUPPORTED;
    a.coef_dev = (const T*)m.coef_dev;
    a.coef_batch = m.coef_batch_stride;
  }
  for (int i = 0; i < KY * KX; ++i) a.coef[i] = dense[i];
  const uintptr_t base = reinterpret_cast<uintptr_t>(m.in);
  int vw = 1;
  if (a.win % 4 == 0 && (base & 15) == 0 && a.in_batch % 4 == 0) vw = 4;
  else if (a.win % 2 == 0 && (base & 7) == 0 && a.in_batch % 2 == 0) vw = 2;
  return launch_t<T>(a, KY, KX, (int)batch, vw, s);
}

}  // namespace

int try_stencil2d(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (p.n_funcs != 1 || p.func[0].kind != KX_AFFINE || p.func[0].post_div != 0.0) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype) return KX_ERR_UNSUPPORTED;
  if (m.fn_elems != 1) return KX_ERR_UNSUPPORTED;
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
    if (!p.func[0].coef_is_int) return KX_ERR_UNSUPPORTED;
    return run<int32_t>(m, p, ay, ax, batch, s);
  }
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
