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
//
// This header holds the kernel template; kx_stencil2d.cu the host-side planning; kx_stencil2d_k<KY>.cu the
// instantiations of one window height each.
#pragma once

#include <type_traits>

#include "kx_internal.cuh"

namespace kx {
namespace s2 {

constexpr int kTX = 128;      // threads per block, each owning 4 output columns
constexpr int kCols = 4;
constexpr int kMaxK = 15;

template <typename T>
struct S2Args {
  const T* in;
  T* out;
  const T* coef_dev;           // optional [KY*KX] device table (dense, row-major; coef_rev: in REVERSED order)
  int coef_rev;
  int64_t in_batch, out_batch, coef_batch;
  int hin, win, hout, wout;
  int ly, lx;                  // left borders
  int th;                      // output rows per block
  uint32_t mask, mask_hi;      // bit ky*KX+kx set = tap present (kAffine: at most 49 taps; two words keep the tests 32-bit)
  int use_box, by0, by1, bx0, bx1;  // offset maps: outputs outside the box copy the window centre
  float post_mul;              // != 0: result = (bias + sum) * post_mul, post_mul = 1 / n (window mean), f32 only
  T bias;
  T coef[kMaxK * kMaxK];       // kAffine: dense KY*KX table; kBox: coef[0]; kSep: coef[ky] and coef[kMaxK + kx]
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
// OP = kAffine: bias + sum of coef * tap over the masked window; kMax / kMin: reduction over the WHOLE window;
// kBox: bias + c * (sum of the WHOLE window), for programs whose taps all carry the same coefficient c (np.sum,
// np.mean, box blurs) -- summed rows first, then columns (fp32: another order than the row-major FMA chain, inside
// the stated 1e-5 * sum|c x| tolerance; int32: exact);
// kSep: bias + sum_ky cy[ky] * (sum_kx cx[kx] * tap), for fp32 coefficient tables that are an outer product
// cy (x) cx to within fp32 rounding (Gaussian blurs, README.md:277-294; Sobel-like filters): KY + KX FMAs per output
// instead of KY * KX, which keeps windows up to 15x15 on the HBM roofline instead of the FMA pipe.
enum S2Op { kAffine = 0, kMax = 1, kMin = 2, kBox = 3, kSep = 4 };

// max / min propagate NaN like np.max / jnp.max (max.NaN.f32); kBox adds
template <typename T, int OP>
__device__ __forceinline__ T s2_fold(T acc, T v) {
  if constexpr (OP == kBox) {
    if constexpr (std::is_same<T, float>::value) return acc + v;
    else return (T)((uint32_t)acc + (uint32_t)v);
  } else {
    return OP == kMax ? red_max(acc, v) : red_min(acc, v);
  }
}

#include "kx_stencil2d_minb.inc"
// Which instantiations carry the INNER path (and the register cap that goes with it).  Measured on B200, 16384^2 fp32,
// fraction of the HBM roofline without -> with (profiles/r2e_s2_inner.txt): 1x7 moving average 0.70 -> 0.92, 5x5 box blur
// 0.90 -> 0.92, 5x5 max filter 0.92 -> 0.94, dense 7x7 0.37 -> 0.44; but the small windows LOSE: 3x3 5-point (C2)
// 1.02 -> 0.99 and its 8-byte aligned backward pass 1.01 -> 0.91 (the cap alone: spills where ptxas used to
// rematerialise), 3x3 max filter 1.02 -> 0.89, separable 4x4 0.92 -> 0.83.  So: row windows of 5 taps and more and
// windows of 25 taps and more; every other instantiation compiles exactly as before (minBlocks 0 = not specified;
// minBlocks 1 is NOT the same: ptxas then stops rematerialising and the 3x3 kernels take 85-92 registers instead of 54).
__host__ __device__ constexpr bool s2_inner_path(int KY, int KX, int VW, int OP) {
  return VW == 4 && ((KY == 1 && KX >= 5) || KY * KX >= 25);
}

constexpr int s2_minb(bool is_float, int KY, int KX, int SHIFT, int VW, int U, int OP) {
  const unsigned key = (((((((is_float ? 1u : 0u) * 16 + KY) * 16 + KX) * 4 + SHIFT) * 8 + VW) * 8 + U) * 8 + OP);
  for (unsigned i = 0; i < sizeof(kS2MinB) / sizeof(kS2MinB[0]); ++i)
    if (kS2MinB[i][0] == key) return (int)kS2MinB[i][1];
  return 1;   // an instantiation the table does not know: no cap
}

template <typename T, int KY, int KX, int SHIFT, int VW, int U, int OP>
__global__ void __launch_bounds__(kTX, s2_inner_path(KY, KX, VW, OP) ? s2_minb(std::is_same<T, float>::value, KY, KX, SHIFT, VW, U, OP) : 0)
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

  T coef[OP == kAffine ? KY * KX : 1];
  if (OP != kAffine) {
    coef[0] = T(0);   // unused
  } else if (a.coef_dev) {
    const T* cd = a.coef_dev + (int64_t)blockIdx.z * a.coef_batch;
#pragma unroll
    for (int i = 0; i < KY * KX; ++i) coef[i] = cd[a.coef_rev ? KY * KX - 1 - i : i];   // reversed: a transposed program
  } else {
#pragma unroll
    for (int i = 0; i < KY * KX; ++i) coef[i] = a.coef[i];
  }

  // INNER (warp-uniform, see below): every row and every aligned vector the warp reads exists -> one 64-bit multiply-add
  // and NV unconditional 128-bit loads per row.  With the bounds tests the compiler re-derives column index, batch
  // offset and predicates per vector (about 12 instructions per loaded vector, measured on the mesh kernel: 48 per row
  // load against 20 FMAs), which is what made smap, the 1-D FIR filters and the small grids issue-bound (75-80 % issue
  // active at 0.71-0.94 of the roofline).
  const T* __restrict__ colp = in + xa;            // element 0 of the thread's row buffer in row 0 (dereferenced only where valid)
  // Separable operators (whole-window max / min, and sums with one common coefficient) queue the HORIZONTAL fold
  // of every input row (4 values per thread) instead of the row itself: KX-1 + KY-1 operations per output
  // instead of KY*KX-1, and a queue of 4-wide rows.
  constexpr bool SEP = OP != kAffine;
  constexpr int QW = SEP ? kCols : NE;
  auto hfold = [&](const T(&row)[NE], T(&dst)[QW]) {   // SEP only: horizontal fold of one input row
#pragma unroll
    for (int i = 0; i < kCols; ++i) {
      T v;
      if constexpr (OP == kSep) {
        if constexpr (std::is_same<T, float>::value) {
          if (i & 1) continue;      // outputs (i, i + 1) are folded together below
          float v0 = a.coef[kMaxK] * row[SHIFT + i], v1 = a.coef[kMaxK] * row[SHIFT + i + 1];
#pragma unroll
          for (int kx = 1; kx < KX; ++kx) {
            const float wx = a.coef[kMaxK + kx];
            ffma2(v0, v1, wx, wx, row[SHIFT + kx + i], row[SHIFT + kx + i + 1]);
          }
          dst[i] = v0;
          dst[i + 1] = v1;
          continue;
        }
        v = a.coef[kMaxK] * row[SHIFT + i];
#pragma unroll
        for (int kx = 1; kx < KX; ++kx) v = mad_wrap(a.coef[kMaxK + kx], row[SHIFT + kx + i], v);
      } else {
        v = row[SHIFT + i];
#pragma unroll
        for (int kx = 1; kx < KX; ++kx) v = s2_fold<T, OP>(v, row[SHIFT + kx + i]);
      }
      dst[i < QW ? i : 0] = v;
    }
  };

  T* __restrict__ outp = out + j0;
  const bool st_al = (a.wout % 4 == 0) && (a.out_batch % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);   // every output row 16-byte aligned
  auto march = [&](auto inner_c) {
    constexpr bool INNER = decltype(inner_c)::value;
    auto load_row = [&](int iy, T(&dst)[NE]) {
      const bool row_ok = (iy >= 0) && (iy < a.hin);
      const T* rp = in + (int64_t)iy * a.win;
      if constexpr (INNER && VW == 4) {
        const T* ip = colp + (int64_t)iy * a.win;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const Vec4<T> t = ld_vec4(ip + 4 * v);
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[(4 * v + e) % NE] = t.v[e];
        }
      } else if (VW == 4) {
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
  
    T q[KY - 1 + U][QW];
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky) {
      if constexpr (!SEP) {
        load_row(r0 - a.ly + ky, q[ky]);
      } else {
        T row[NE];
        load_row(r0 - a.ly + ky, row);
        hfold(row, q[ky]);
      }
    }
  
    for (int r = r0; r < r1; r += U) {
      if constexpr (!SEP) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (r + u < r1) load_row(r + u - a.ly + KY - 1, q[KY - 1 + u]);
      } else {
        T raw[U][NE];   // all U rows are requested before the first fold waits for one
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (r + u < r1) load_row(r + u - a.ly + KY - 1, raw[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (r + u < r1) hfold(raw[u], q[KY - 1 + u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + u < r1) {
          T acc[kCols];
          if constexpr (!SEP) {
#pragma unroll
            for (int i = 0; i < kCols; ++i) acc[i] = a.bias;
#pragma unroll
            for (int ky = 0; ky < KY; ++ky) {
#pragma unroll
              for (int kx = 0; kx < KX; ++kx) {
                if ((ky * KX + kx < 32) ? (a.mask & (1u << ((ky * KX + kx) & 31))) : (a.mask_hi & (1u << ((ky * KX + kx) & 31)))) {
#pragma unroll
                  for (int i = 0; i < kCols; ++i)
                    acc[i] = mad_wrap(coef[ky * KX + kx], q[u + ky][SHIFT + kx + i], acc[i]);
                }
              }
            }
          } else {
            if constexpr (OP == kSep && std::is_same<T, float>::value) {
              // vertical fold as packed FFMA2: two output columns per instruction, the weight is a broadcast scalar
              // (same rounding as the scalar chain; the wide windows are issue-bound: 74 % issue-active at 15x15)
#pragma unroll
              for (int i = 0; i < kCols; ++i) acc[i] = a.bias;
#pragma unroll
              for (int ky = 0; ky < KY; ++ky) {
                const float wy = a.coef[ky];
                ffma2(acc[0], acc[1], wy, wy, q[u + ky][0], q[u + ky][1]);
                ffma2(acc[2], acc[3], wy, wy, q[u + ky][2], q[u + ky][3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < kCols; ++i) {
                if constexpr (OP == kSep) {
                  T v = a.bias;
#pragma unroll
                  for (int ky = 0; ky < KY; ++ky) v = mad_wrap(a.coef[ky], q[u + ky][i], v);
                  acc[i] = v;
                } else {
                  T v = q[u][i];
#pragma unroll
                  for (int ky = 1; ky < KY; ++ky) v = s2_fold<T, OP>(v, q[u + ky][i]);
                  acc[i] = (OP == kBox) ? mad_wrap(a.coef[0], v, a.bias) : v;
                }
              }
            }
          }
          if (std::is_same<T, float>::value && (OP == kAffine || OP == kBox || OP == kSep) && a.post_mul != 0.f) {
            // block-uniform: window mean.  One multiply by the host-rounded reciprocal (<= 1.5 ulp from the true
            // quotient, exact for power-of-two windows) instead of ~10 instructions of IEEE division per output:
            // the 5x5 box blur went from 0.74 to 0.9x of peak
#pragma unroll
            for (int i = 0; i < kCols; ++i) acc[i] = (T)((float)acc[i] * a.post_mul);
          }
          if constexpr (!SEP) {
            if (a.use_box) {  // block-uniform: smap / offset_kernel_map fused into one pass
              const bool rin = (r + u >= a.by0) && (r + u < a.by1);
#pragma unroll
              for (int i = 0; i < kCols; ++i)
                if (!(rin && j0 + i >= a.bx0 && j0 + i < a.bx1)) acc[i] = q[u + (KY - 1) / 2][SHIFT + (KX - 1) / 2 + i];
            }
          }
          if constexpr (INNER) {
            T* o = outp + (int64_t)(r + u) * a.wout;
            if (st_al) {
              int4 raw4;
              raw4.x = reinterpret_cast<const int&>(acc[0]);
              raw4.y = reinterpret_cast<const int&>(acc[1]);
              raw4.z = reinterpret_cast<const int&>(acc[2]);
              raw4.w = reinterpret_cast<const int&>(acc[3]);
              __stcs(reinterpret_cast<int4*>(o), raw4);
            } else {
              st_row4(o, acc, 4);
            }
          } else {
            st_row4(out + (int64_t)(r + u) * a.wout + j0, acc, n_valid);
          }
        }
      }
#pragma unroll
      for (int ky = 0; ky < KY - 1; ++ky) {
#pragma unroll
        for (int e = 0; e < QW; ++e) q[ky][e] = q[U + ky][e];
      }
    }
  };
  // warp-uniform: the rows r0 - ly .. r1 - 1 - ly + KY - 1 exist, so do the aligned vectors of lanes 0 and 31, and every
  // lane owns 4 output columns
  const int jw = (blockIdx.x * kTX + (threadIdx.x & ~31)) * kCols;
  const int xw = jw - a.lx - SHIFT;
  const bool inner = VW == 4 && (r0 - a.ly >= 0) && (r1 - 1 - a.ly + KY - 1 < a.hin) && (xw >= 0) &&
                     (xw + 31 * kCols + NE <= a.win) && (jw + 32 * kCols <= a.wout);
  if constexpr (s2_inner_path(KY, KX, VW, OP)) {
    if (inner) march(std::true_type{});
    else march(std::false_type{});
  } else {
    march(std::false_type{});
  }
}

template <typename T, int KY, int KX, int OP>
int launch_kk(const S2Args<T>& a, int batch, int vw, cudaStream_t s) {
  constexpr int U = ((OP == kAffine && (KY * KX >= 15 || KX >= 9)) || (OP != kAffine && KX >= 7)) ? 2 : 4;
  dim3 grid((a.wout + kTX * kCols - 1) / (kTX * kCols), (a.hout + a.th - 1) / a.th, batch);
  const int shift = vw > 1 ? (((-a.lx) % vw) + vw) % vw : 0;
  if (vw == 1) stencil2d_kernel<T, KY, KX, 0, 1, U, OP><<<grid, kTX, 0, s>>>(a);
  else if (vw == 2) {
    if constexpr (OP == kAffine) {
      if (shift == 0) stencil2d_kernel<T, KY, KX, 0, 2, U, OP><<<grid, kTX, 0, s>>>(a);
      else stencil2d_kernel<T, KY, KX, 1, 2, U, OP><<<grid, kTX, 0, s>>>(a);
    } else {
      return KX_ERR_UNSUPPORTED;
    }
  } else if (shift == 0) stencil2d_kernel<T, KY, KX, 0, 4, U, OP><<<grid, kTX, 0, s>>>(a);
  else if (shift == 1) stencil2d_kernel<T, KY, KX, 1, 4, U, OP><<<grid, kTX, 0, s>>>(a);
  else if (shift == 2) stencil2d_kernel<T, KY, KX, 2, 4, U, OP><<<grid, kTX, 0, s>>>(a);
  else stencil2d_kernel<T, KY, KX, 3, 4, U, OP><<<grid, kTX, 0, s>>>(a);
  return after_launch("stencil2d_kernel");
}

// The separable operators come in the 16-byte aligned and the scalar-load variants only.  Box sums and outer-product
// tables are instantiated where they save work; windows larger than 7x7 exist in separable form only.
template <int KY, int KX>
struct S2Has {
  static constexpr bool affine = (KY <= 7 && KX <= 7) || (KY == 1 && KX <= kMaxK);   // 1-row FIR filters up to 15 taps
  static constexpr bool box = KY >= 2 && KY * KX >= 7;
  static constexpr bool sep = KY >= 2 && KX >= 2 && KY * KX >= 15;
};
inline bool s2_has_box(int ky, int kx) { return ky >= 2 && ky * kx >= 7; }   // 1-row windows gain nothing
inline bool s2_has_affine(int ky, int kx) { return (ky <= 7 && kx <= 7) || (ky == 1 && kx <= kMaxK); }
inline bool s2_has_sep(int ky, int kx) { return ky >= 2 && kx >= 2 && ky * kx >= 15; }

template <typename T, int KY, int KX>
int launch_op(const S2Args<T>& a, int op, int batch, int vw, cudaStream_t s) {
  if constexpr (S2Has<KY, KX>::affine) {
    if (op == kAffine) return launch_kk<T, KY, KX, kAffine>(a, batch, vw, s);
  }
  if (vw == 2) vw = 1;
  if (op == kMax) return launch_kk<T, KY, KX, kMax>(a, batch, vw, s);
  if (op == kMin) return launch_kk<T, KY, KX, kMin>(a, batch, vw, s);
  if constexpr (S2Has<KY, KX>::box) {
    if (op == kBox) return launch_kk<T, KY, KX, kBox>(a, batch, vw, s);
  }
  if constexpr (S2Has<KY, KX>::sep && std::is_same<T, float>::value) {
    if (op == kSep) return launch_kk<T, KY, KX, kSep>(a, batch, vw, s);
  }
  return KX_ERR_UNSUPPORTED;
}

// One translation unit per window height (kx_stencil2d_k<KY>.cu) so the instantiations compile in parallel.
#define KX_S2_DECLARE(KY)                                                                                     \
  int s2_launch_k##KY(const S2Args<float>& a, int kx, int op, int batch, int vw, cudaStream_t s);            \
  int s2_launch_k##KY(const S2Args<int32_t>& a, int kx, int op, int batch, int vw, cudaStream_t s);
KX_S2_DECLARE(1) KX_S2_DECLARE(2) KX_S2_DECLARE(3) KX_S2_DECLARE(4) KX_S2_DECLARE(5) KX_S2_DECLARE(7)
KX_S2_DECLARE(9) KX_S2_DECLARE(11) KX_S2_DECLARE(13) KX_S2_DECLARE(15)
#undef KX_S2_DECLARE


}  // namespace s2
}  // namespace kx
