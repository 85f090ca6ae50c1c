// Fast path: single-function AFFINE stencils over the last three axes, stride 1, f32 / i32
// (3-D 7-point Laplacian, conv-as-kmap with a (C,3,3) window over a channel-first image).
//
// 2.5-D blocking, HBM-bound (8 B per lattice update):
//   * a block owns a (kTYB*RY rows) x (128 columns) tile of the y/x plane and MARCHES
//     along axis 0; every thread owns 4 consecutive columns x RY rows and keeps the last
//     KZ planes of its footprint in a register queue, so each input plane is fetched from
//     L2/HBM once per block (plus the y halo rows, (KY-1)/(kTYB*RY) extra) and re-used KZ
//     times from registers;
//   * within a plane the rows a thread shares with its y neighbours and the x halo come
//     from L1 (the block touches them in the same z step);
//   * loads are 128-bit aligned vectors with the window misalignment as a template
//     parameter (see kx_stencil2d.cu); zero padding is a predicate;
//   * STAR (all taps on the axes through the window centre, e.g. the 7-point Laplacian)
//     is a compile-time flag: off-axis rows / planes are then dead code, which shrinks the
//     register queue to the centre plane plus one 4-vector per neighbour plane and row.
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kTXB = 32;   // threads along x (each 4 columns)
constexpr int kTYB = 8;    // threads along y
constexpr int kCols = 4;
constexpr int kMaxTaps3 = 125;

template <typename T>
struct S3Args {
  const T* in;
  T* out;
  const T* coef_dev;          // optional dense [KZ*KY*KX] device table
  int64_t in_batch, out_batch;
  int din, hin, win, dout, hout, wout;
  int lz, ly, lx;
  int zc;                     // output planes per block
  int nzc;                    // z chunks per batch item
  T bias;
  uint32_t mask[4];           // bit (kz*KY+ky)*KX+kx
  T coef[kMaxTaps3];
};

template <typename T>
__device__ __forceinline__ void ld4(const T* p, T (&dst)[4]) {
  const int4 raw = __ldg(reinterpret_cast<const int4*>(p));
  dst[0] = reinterpret_cast<const T&>(raw.x);
  dst[1] = reinterpret_cast<const T&>(raw.y);
  dst[2] = reinterpret_cast<const T&>(raw.z);
  dst[3] = reinterpret_cast<const T&>(raw.w);
}

template <typename T>
__device__ __forceinline__ void st4(T* p, const T (&a)[4], int n_valid) {
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

template <typename T, int KZ, int KY, int KX, int SHIFT, bool VEC, int RY, bool STAR>
__global__ void __launch_bounds__(kTXB * kTYB)
stencil3d_kernel(const __grid_constant__ S3Args<T> a) {
  constexpr int NV = (SHIFT + KX + kCols - 1 + 3) / 4;
  constexpr int NE = NV * 4;
  constexpr int NR = RY + KY - 1;       // input rows of one plane a thread touches
  constexpr int CZ = (KZ - 1) / 2, CY = (KY - 1) / 2, CX = (KX - 1) / 2;

  const int j0 = (blockIdx.x * kTXB + threadIdx.x) * kCols;
  const int i0 = (blockIdx.y * kTYB + threadIdx.y) * RY;
  if (j0 >= a.wout || i0 >= a.hout) return;
  const int b = blockIdx.z / a.nzc;
  const int z0 = (blockIdx.z - b * a.nzc) * a.zc;
  const int z1 = min(z0 + a.zc, a.dout);
  const T* __restrict__ in = a.in + (int64_t)b * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)b * a.out_batch;
  const int xa = j0 - a.lx - SHIFT;
  const int n_valid = min(kCols, a.wout - j0);

  T coef[KZ * KY * KX];
  if (a.coef_dev) {
#pragma unroll
    for (int i = 0; i < KZ * KY * KX; ++i) coef[i] = a.coef_dev[i];
  } else {
#pragma unroll
    for (int i = 0; i < KZ * KY * KX; ++i) coef[i] = a.coef[i];
  }

  auto load_plane = [&](int iz, T(&dst)[NR][NE]) {
    const bool z_ok = (iz >= 0) && (iz < a.din);
    const T* pp = in + (int64_t)iz * a.hin * a.win;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int iy = i0 - a.ly + r;
      const bool row_ok = z_ok && (iy >= 0) && (iy < a.hin);
      const T* rp = pp + (int64_t)iy * a.win;
      if (VEC) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = xa + 4 * v;
          T t4[4] = {T(0), T(0), T(0), T(0)};
          if (row_ok && c >= 0 && c + 3 < a.win) ld4(rp + c, t4);
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[r][4 * v + e] = t4[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < KX + kCols - 1; ++e) {
          const int c = xa + e;
          dst[r][e] = (row_ok && c >= 0 && c < a.win) ? __ldg(rp + c) : T(0);
        }
      }
    }
  };

  T q[KZ][NR][NE];
#pragma unroll
  for (int kz = 0; kz < KZ - 1; ++kz) load_plane(z0 - a.lz + kz, q[kz]);

  for (int z = z0; z < z1; ++z) {
    // (a one-plane register prefetch was tried: 124 registers, 2 blocks/SM, 8 % slower)
    load_plane(z - a.lz + KZ - 1, q[KZ - 1]);
#pragma unroll
    for (int ry = 0; ry < RY; ++ry) {
      if (i0 + ry < a.hout) {
        T acc[kCols];
#pragma unroll
        for (int i = 0; i < kCols; ++i) acc[i] = a.bias;
#pragma unroll
        for (int kz = 0; kz < KZ; ++kz) {
#pragma unroll
          for (int ky = 0; ky < KY; ++ky) {
#pragma unroll
            for (int kx = 0; kx < KX; ++kx) {
              constexpr bool dummy = true;
              (void)dummy;
              const bool on_axis = ((kz == CZ) + (ky == CY) + (kx == CX)) >= 2;
              if (STAR && !on_axis) continue;
              const int slot = (kz * KY + ky) * KX + kx;
              if (a.mask[slot >> 5] & (1u << (slot & 31))) {
#pragma unroll
                for (int i = 0; i < kCols; ++i)
                  acc[i] = mad_wrap(coef[slot], q[kz][ry + ky][SHIFT + kx + i], acc[i]);
              }
            }
          }
        }
        st4(out + ((int64_t)z * a.hout + (i0 + ry)) * a.wout + j0, acc, n_valid);
      }
    }
#pragma unroll
    for (int kz = 0; kz < KZ - 1; ++kz) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int e = 0; e < NE; ++e) q[kz][r][e] = q[kz + 1][r][e];
      }
    }
  }
}

template <typename T, int KZ, int KY, int KX, int RY, bool STAR>
int launch_kkk(S3Args<T>& a, int batch, bool vec, cudaStream_t s) {
  const int tiles_x = (a.wout + kTXB * kCols - 1) / (kTXB * kCols);
  const int tiles_y = (a.hout + kTYB * RY - 1) / (kTYB * RY);
  // enough blocks for a few waves (2 resident blocks / SM), but long z marches for reuse
  const int64_t cols = (int64_t)tiles_x * tiles_y * batch;
  // >= ~16 waves of 2 resident blocks per SM keeps the last partial wave under a few percent
  const int64_t want = (int64_t)sm_count() * 2 * 16;
  int nzc = (int)((want + cols - 1) / cols);
  if (nzc < 1) nzc = 1;
  int zc = (a.dout + nzc - 1) / nzc;
  // short z chunks keep concurrently resident blocks on a compact slab so tile halos hit L2
  // (benchmarks/exp3d.cu: 32-64 planes beat full-depth marches by 40 %)
  static const int zc_cap = getenv("KX_S3_ZC") ? atoi(getenv("KX_S3_ZC")) : 64;
  if (zc > zc_cap && zc_cap > 0) zc = zc_cap;
  if (zc < 16) zc = a.dout < 16 ? a.dout : 16;
  nzc = (a.dout + zc - 1) / zc;
  a.zc = zc;
  a.nzc = nzc;
  if ((int64_t)nzc * batch > 65535 || tiles_y > 65535) return KX_ERR_UNSUPPORTED;
  dim3 grid(tiles_x, tiles_y, nzc * batch), block(kTXB, kTYB);
  const int shift = vec ? (((-a.lx) % 4) + 4) % 4 : 0;
  if (!vec) stencil3d_kernel<T, KZ, KY, KX, 0, false, RY, STAR><<<grid, block, 0, s>>>(a);
  else if (shift == 0) stencil3d_kernel<T, KZ, KY, KX, 0, true, RY, STAR><<<grid, block, 0, s>>>(a);
  else if (shift == 1) stencil3d_kernel<T, KZ, KY, KX, 1, true, RY, STAR><<<grid, block, 0, s>>>(a);
  else if (shift == 2) stencil3d_kernel<T, KZ, KY, KX, 2, true, RY, STAR><<<grid, block, 0, s>>>(a);
  else stencil3d_kernel<T, KZ, KY, KX, 3, true, RY, STAR><<<grid, block, 0, s>>>(a);
  return after_launch("stencil3d_kernel");
}

template <typename T>
int run3(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int az = g.ndim - 3, ay = g.ndim - 2, ax = g.ndim - 1;
  const int KZ = g.ksize[az], KY = g.ksize[ay], KX = g.ksize[ax];
  if (KZ * KY * KX > kMaxTaps3) return KX_ERR_UNSUPPORTED;
  S3Args<T> a;
  memset(&a, 0, sizeof(a));
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.din = (int)g.in_shape[az];
  a.hin = (int)g.in_shape[ay];
  a.win = (int)g.in_shape[ax];
  a.dout = (int)g.out_shape[az];
  a.hout = (int)g.out_shape[ay];
  a.wout = (int)g.out_shape[ax];
  a.lz = g.lo[az];
  a.ly = g.lo[ay];
  a.lx = g.lo[ax];
  a.in_batch = (int64_t)a.din * a.hin * a.win;
  a.out_batch = (int64_t)a.dout * a.hout * a.wout;
  a.bias = (T)f.bias;
  const int cz = (KZ - 1) / 2, cy = (KY - 1) / 2, cx = (KX - 1) / 2;
  bool star = true, dense_order = (f.tap_end - f.tap_begin) == KZ * KY * KX;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int oz = p.tap_off[t][az], oy = p.tap_off[t][ay], ox = p.tap_off[t][ax];
    const int slot = (oz * KY + oy) * KX + ox;
    if (a.mask[slot >> 5] & (1u << (slot & 31))) return KX_ERR_UNSUPPORTED;  // duplicate tap
    a.mask[slot >> 5] |= 1u << (slot & 31);
    a.coef[slot] = (T)p.coef[t];
    if (((oz == cz) + (oy == cy) + (ox == cx)) < 2) star = false;
    if (slot != t - f.tap_begin) dense_order = false;
  }
  if (m.coef_dev) {
    if (!dense_order || m.coef_batch_stride != 0) return KX_ERR_UNSUPPORTED;
    a.coef_dev = (const T*)m.coef_dev;
    star = false;
  }
  const bool vec = (a.win % 4 == 0) && ((reinterpret_cast<uintptr_t>(m.in) & 15) == 0);
#define KX_CASE3(Z, Y, X, RYS, RYD)                                              \
  if (KZ == Z && KY == Y && KX == X) {                                           \
    if (star) return launch_kkk<T, Z, Y, X, RYS, true>(a, (int)batch, vec, s);   \
    return launch_kkk<T, Z, Y, X, RYD, false>(a, (int)batch, vec, s);            \
  }
  KX_CASE3(3, 3, 3, 2, 1)
  KX_CASE3(2, 2, 2, 2, 2)
#undef KX_CASE3
  if (KZ == 5 && KY == 5 && KX == 5 && star) return launch_kkk<T, 5, 5, 5, 1, true>(a, (int)batch, vec, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace

int try_stencil3d(const MapArgs& m, const KxProgram& p0, cudaStream_t s) {
  const KxGeom& g = m.g;
  if (g.ndim < 3) return KX_ERR_UNSUPPORTED;
  if (p0.n_funcs != 1 || p0.func[0].kind != KX_AFFINE) return KX_ERR_UNSUPPORTED;
  // A final true division ((bias + sum) / d: np.mean of a 3-D window, `(...) / 6.0` in a Jacobi step) is folded into
  // the host coefficients: c / d and bias / d rounded once from double (<= 0.5 ulp each, i.e. ~6e-8 * sum|c x| away from
  // the divided form -- the 2-D kernels multiply by the rounded reciprocal instead; both far inside the 1e-5 bound).
  // Before, these programs were declined here and ran on the generic kernel (0.04 of the HBM roofline).
  static thread_local KxProgram scaled;
  const KxProgram* pp = &p0;
  if (p0.func[0].post_div != 0.0) {
    if (g.in_dtype != KX_F32 || m.coef_dev) return KX_ERR_UNSUPPORTED;
    scaled = p0;
    const double d = p0.func[0].post_div;
    scaled.func[0].post_div = 0.0;
    scaled.func[0].bias = (double)(float)(p0.func[0].bias / d);
    for (int t = p0.func[0].tap_begin; t < p0.func[0].tap_end; ++t) scaled.coef[t] = (double)(float)(p0.coef[t] / d);
    pp = &scaled;
  }
  const KxProgram& p = *pp;
  if (g.in_dtype != g.out_dtype || m.fn_elems != 1) return KX_ERR_UNSUPPORTED;
  int64_t lead = 1;
  for (int d = 0; d < g.ndim - 3; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    lead *= g.in_shape[d];
  }
  for (int d = g.ndim - 3; d < g.ndim; ++d) {
    if (g.stride[d] != 1) return KX_ERR_UNSUPPORTED;
    if (g.in_shape[d] > 0x3fffffff || g.out_shape[d] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
  }
  if (g.ksize[g.ndim - 3] == 1) return KX_ERR_UNSUPPORTED;  // a 2-D problem: stencil2d folds it
  const int64_t batch = m.batch * lead;
  if (m.coef_dev && batch > 1 && m.coef_batch_stride != 0) return KX_ERR_UNSUPPORTED;
  if (batch <= 65535) {
    // 7-point-style stars take the cp.async shared-memory ring (kx_stencil3d_star.cu)
    const int rc = try_stencil3d_star(m, p, batch, s);
    if (rc != KX_ERR_UNSUPPORTED) return rc;
  }
  if (batch <= 65535) {
    // dense 3x3x3 windows with host weights: plane-stationary kernel (kx_stencil3d_dense.cu)
    const int rc = try_stencil3d_dense(m, p, batch, s);
    if (rc != KX_ERR_UNSUPPORTED) return rc;
  }
  {
    // dense 3x3x3 windows: the conv row-ring kernel with the three z planes as channels (kx_conv2d.cu)
    const int rc = try_conv_dense3d(m, p, batch, s);
    if (rc != KX_ERR_UNSUPPORTED) return rc;
  }
  if (g.in_dtype == KX_F32) return run3<float>(m, p, batch, s);
  if (g.in_dtype == KX_I32 && p.func[0].coef_is_int) return run3<int32_t>(m, p, batch, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
