// Generic kernels: any rank <= 4, any stride / border (incl. negative), multi-function
// region tables, f32 / i32.  One thread per output element; window taps are gathered
// straight from global memory (L1/L2 serve the overlap).  These are the universal
// correct path; the roofline-tuned templates live in kx_stencil2d.cu / kx_stencil3d.cu /
// kx_patch.cu / kx_scan_rowmarch.cu and take over whenever they apply.
//
// Semantics follow kernex/_src/map.py:84-122 (kmap), kernex/_src/scan.py:84-113 (kscan)
// and kernex/_src/utils.py:302-382 (region -> function selection).
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kThreads = 256;

// value of one window function at window `origin`, reading `src` with zero fill
// outside [0, shape) (src_off shifts origin into src coordinates for the scan state)
template <typename TIN, typename TC, bool kBounds>
__device__ __forceinline__ TC eval_func(const DevProgram<TC>& p, int f, const TIN* __restrict__ src,
                                        const int64_t* shape, const Strides& st, const int64_t* origin,
                                        const TC* __restrict__ coef_dev) {
  const int kind = p.func[f].kind;
  const int t0 = p.func[f].tap_begin, t1 = p.func[f].tap_end;
  TC acc = (kind == KX_AFFINE) ? p.func[f].bias : TC(0);
  for (int t = t0; t < t1; ++t) {
    int64_t off = 0;
    bool inb = true;
#pragma unroll
    for (int d = 0; d < KX_MAX_DIMS; ++d) {
      if (d < p.ndim) {
        const int64_t c = origin[d] + p.tap_off[t][d];
        if (kBounds) inb = inb && (c >= 0) && (c < shape[d]);
        off += c * st.v[d];
      }
    }
    const TC v = inb ? (TC)src[off] : TC(0);
    if (kind == KX_AFFINE) {
      const TC c = coef_dev ? coef_dev[t] : p.coef[t];
      acc = mad_wrap(c, v, acc);
    } else if (kind == KX_REDUCE_MAX) {
      acc = (t == t0) ? v : red_max(acc, v);
    } else {  // KX_REDUCE_MIN (KX_GATHER is handled by the caller)
      acc = (t == t0) ? v : red_min(acc, v);
    }
  }
  if (kind == KX_AFFINE && p.func[f].post_div != 0.f) acc = (TC)((float)acc / p.func[f].post_div);
  return acc;
}

// ------------------------------------------------------------------ kmap ----
template <typename TIN, typename TC>
__global__ void __launch_bounds__(kThreads)
generic_map_kernel(const __grid_constant__ MapArgs a, const __grid_constant__ DevProgram<TC> p) {
  const int nd = a.g.ndim;
  const Strides ist = suffix_strides(a.g.in_shape, nd);
  const int64_t per_batch = a.n_win * a.fn_elems;
  const int64_t total = per_batch * a.batch;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int64_t b = i / per_batch;
    int64_t r = i - b * per_batch;
    const int e = (int)(r % a.fn_elems);
    int64_t w = r / a.fn_elems;
    int64_t origin[KX_MAX_DIMS], centre[KX_MAX_DIMS];
    int64_t oidx = a.out_base;
#pragma unroll
    for (int d = KX_MAX_DIMS - 1; d >= 0; --d) {
      if (d < nd) {
        const int64_t j = w % a.g.out_shape[d];
        w /= a.g.out_shape[d];
        origin[d] = (int64_t)j * a.g.stride[d] - a.g.lo[d];
        centre[d] = origin[d] + (a.g.ksize[d] - 1) / 2;
        oidx += j * a.out_pitch[d];
      }
    }
    const TIN* src = (const TIN*)a.in + b * a.in_elems;
    const TC* cdev = a.coef_dev ? (const TC*)a.coef_dev + b * a.coef_batch_stride : nullptr;
    TC* dst = (TC*)a.out + b * a.out_batch_stride;
    const int f = select_func(p, centre);
    if (p.func[f].kind == KX_GATHER) {
      const int t = p.func[f].tap_begin + e;
      int64_t off = 0;
      bool inb = true;
#pragma unroll
      for (int d = 0; d < KX_MAX_DIMS; ++d) {
        if (d < nd) {
          const int64_t c = origin[d] + p.tap_off[t][d];
          inb = inb && (c >= 0) && (c < a.g.in_shape[d]);
          off += c * ist.v[d];
        }
      }
      dst[oidx * a.fn_elems + e] = inb ? (TC)src[off] : TC(0);
    } else {
      dst[oidx] = eval_func<TIN, TC, true>(p, f, src, a.g.in_shape, ist, origin, cdev);
    }
  }
}

template <typename TIN, typename TC>
int launch_map_t(const MapArgs& a, const KxProgram& prog, cudaStream_t s) {
  static thread_local DevProgram<TC> dp;  // 10 KB: keep it off the stack
  make_dev_program<TC>(prog, a.g.ndim, dp);
  const int64_t total = a.n_win * a.fn_elems * a.batch;
  if (total <= 0) return KX_OK;
  int64_t blocks = (total + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  generic_map_kernel<TIN, TC><<<(unsigned)blocks, kThreads, 0, s>>>(a, dp);
  return after_launch("generic_map_kernel");
}

// ------------------------------------------------------------- VJP (input) ----
// dx[i] = sum_t coef[t] * g[j(i,t)],  j_d = (i_d - tap_d + lo_d) / s_d when divisible & in range
template <typename TC>
__global__ void __launch_bounds__(kThreads)
generic_vjp_input_kernel(const __grid_constant__ KxGeom g, const __grid_constant__ DevProgram<TC> p,
                         const TC* __restrict__ gout, const TC* __restrict__ coef_dev, TC* __restrict__ dx,
                         int64_t n_in) {
  const int nd = g.ndim;
  const Strides ost = suffix_strides(g.out_shape, nd);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_in; i += (int64_t)gridDim.x * kThreads) {
    int64_t coord[KX_MAX_DIMS];
    int64_t r = i;
#pragma unroll
    for (int d = KX_MAX_DIMS - 1; d >= 0; --d) {
      if (d < nd) {
        coord[d] = r % g.in_shape[d];
        r /= g.in_shape[d];
      }
    }
    TC acc = 0;
    for (int t = p.func[0].tap_begin; t < p.func[0].tap_end; ++t) {
      int64_t off = 0;
      bool ok = true;
#pragma unroll
      for (int d = 0; d < KX_MAX_DIMS; ++d) {
        if (d < nd) {
          const int64_t num = coord[d] - p.tap_off[t][d] + g.lo[d];
          const int64_t j = num / g.stride[d];
          ok = ok && (num >= 0) && (j * g.stride[d] == num) && (j < g.out_shape[d]);
          off += j * ost.v[d];
        }
      }
      if (ok) acc = fmaf(coef_dev ? coef_dev[t] : p.coef[t], gout[off], acc);
    }
    // forward = (bias + sum) / post_div  =>  the cotangent carries the same final division
    if (p.func[0].post_div != 0.f) acc = (TC)((float)acc / p.func[0].post_div);
    dx[i] = acc;
  }
}

// -------------------------------------------------------------- VJP (coef) ----
// stage 1: partial[t][b] = sum over the block's windows of g[j] * x[origin(j)+tap[t]]
__global__ void __launch_bounds__(kThreads)
generic_vjp_coef_stage1(const __grid_constant__ KxGeom g, const __grid_constant__ DevProgram<float> p,
                        const float* __restrict__ x, const float* __restrict__ gout, float* __restrict__ partial,
                        int64_t n_win) {
  const int nd = g.ndim;
  const int t = p.func[0].tap_begin + blockIdx.y;
  const Strides ist = suffix_strides(g.in_shape, nd);
  float acc = 0.f;
  for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < n_win; w += (int64_t)gridDim.x * kThreads) {
    int64_t r = w, off = 0;
    bool inb = true;
#pragma unroll
    for (int d = KX_MAX_DIMS - 1; d >= 0; --d) {
      if (d < nd) {
        const int64_t j = r % g.out_shape[d];
        r /= g.out_shape[d];
        const int64_t c = j * g.stride[d] - g.lo[d] + p.tap_off[t][d];
        inb = inb && (c >= 0) && (c < g.in_shape[d]);
        off += c * ist.v[d];
      }
    }
    if (inb) acc = fmaf(gout[w], x[off], acc);
  }
  __shared__ float red[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    partial[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

// stage 2: fixed-order sum of the per-block partials (deterministic)
__global__ void generic_vjp_coef_stage2(const float* __restrict__ partial, float* __restrict__ dcoef, int nblocks,
                                        float post_div) {
  const int t = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 32) acc += partial[(int64_t)t * nblocks + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) dcoef[t] = post_div != 0.f ? acc / post_div : acc;
}

int vjp_coef_blocks(int64_t n_win) {
  int64_t b = (n_win + kThreads * 8 - 1) / (kThreads * 8);
  if (b < 1) b = 1;
  if (b > 1024) b = 1024;
  return (int)b;
}

// ------------------------------------------------------------------ kscan ----
struct ScanArgs {
  KxGeom g;
  int64_t pshape[KX_MAX_DIMS];   // padded state shape
  int32_t padlo[KX_MAX_DIMS];    // max(lo, 0): offset of the array inside the state
  int64_t n_win;
};

template <typename TIN>
__global__ void __launch_bounds__(kThreads)
scan_init_state_kernel(const __grid_constant__ ScanArgs a, const TIN* __restrict__ in, TIN* __restrict__ state,
                       int64_t n_state) {
  const int nd = a.g.ndim;
  const Strides ist = suffix_strides(a.g.in_shape, nd);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_state; i += (int64_t)gridDim.x * kThreads) {
    int64_t r = i, off = 0;
    bool inb = true;
#pragma unroll
    for (int d = KX_MAX_DIMS - 1; d >= 0; --d) {
      if (d < nd) {
        const int64_t c = r % a.pshape[d] - a.padlo[d];
        r /= a.pshape[d];
        inb = inb && (c >= 0) && (c < a.g.in_shape[d]);
        off += c * ist.v[d];
      }
    }
    state[i] = inb ? in[off] : TIN(0);
  }
}

// one scan step for window `w` (row-major linear index): select, read state, write centre
template <typename TIN, typename TC>
__device__ __forceinline__ void scan_step(const ScanArgs& a, const DevProgram<TC>& p, const Strides& sst,
                                          TIN* state, TIN* out, const TC* coef_dev, int64_t w) {
  const int nd = a.g.ndim;
  int64_t origin[KX_MAX_DIMS], centre[KX_MAX_DIMS];
  int64_t r = w, coff = 0;
#pragma unroll
  for (int d = KX_MAX_DIMS - 1; d >= 0; --d) {
    if (d < nd) {
      const int64_t j = r % a.g.out_shape[d];
      r /= a.g.out_shape[d];
      const int64_t o = j * a.g.stride[d] - a.g.lo[d];      // array coordinates
      centre[d] = o + (a.g.ksize[d] - 1) / 2;
      origin[d] = o + a.padlo[d];                           // state coordinates (always >= 0)
      coff += (centre[d] + a.padlo[d]) * sst.v[d];
    }
  }
  const int f = select_func(p, centre);
  TC v;
  if (p.func[f].kind == KX_GATHER) {  // scalar identity-like function: x[tap]
    int64_t off = 0;
    const int t = p.func[f].tap_begin;
#pragma unroll
    for (int d = 0; d < KX_MAX_DIMS; ++d)
      if (d < nd) off += (origin[d] + p.tap_off[t][d]) * sst.v[d];
    v = (TC)state[off];
  } else {
    v = eval_func<TIN, TC, false>(p, f, state, a.pshape, sst, origin, coef_dev);
  }
  const TIN stored = (TIN)v;  // cast to the carried dtype on write (scan.py:50)
  state[coff] = stored;
  out[w] = stored;
}

// strictly sequential sweep: one thread, exact reference order
template <typename TIN, typename TC>
__global__ void scan_sequential_kernel(const __grid_constant__ ScanArgs a, const __grid_constant__ DevProgram<TC> p,
                                       TIN* state, TIN* out, const TC* coef_dev, int reverse) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Strides sst = suffix_strides(a.pshape, a.g.ndim);
  if (!reverse) {
    for (int64_t w = 0; w < a.n_win; ++w) scan_step<TIN, TC>(a, p, sst, state, out, coef_dev, w);
  } else {
    for (int64_t w = a.n_win - 1; w >= 0; --w) scan_step<TIN, TC>(a, p, sst, state, out, coef_dev, w);
  }
}

// all windows of one axis-0 hyperplane in parallel (legal when no function reads another
// cell of the hyperplane being written -- see plane_parallel_ok)
template <typename TIN, typename TC>
__global__ void __launch_bounds__(kThreads)
scan_plane_kernel(const __grid_constant__ ScanArgs a, const __grid_constant__ DevProgram<TC> p, TIN* state,
                  TIN* out, const TC* coef_dev, int64_t j0, int64_t plane) {
  const Strides sst = suffix_strides(a.pshape, a.g.ndim);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < plane; i += (int64_t)gridDim.x * kThreads)
    scan_step<TIN, TC>(a, p, sst, state, out, coef_dev, j0 * plane + i);
}

// A tap is hazardous when it reads, inside the hyperplane currently being written, a
// cell other than the window's own centre: that cell may or may not have been
// overwritten depending on sweep order, so only the sequential kernel is exact.
bool plane_parallel_ok(const KxGeom& g, const KxProgram& prog) {
  if (g.ndim < 2) return false;
  for (int f = 0; f < prog.n_funcs; ++f) {
    for (int t = prog.func[f].tap_begin; t < prog.func[f].tap_end; ++t) {
      const int d0 = prog.tap_off[t][0] - (g.ksize[0] - 1) / 2;
      if (d0 != 0) continue;
      for (int d = 1; d < g.ndim; ++d)
        if (prog.tap_off[t][d] != (g.ksize[d] - 1) / 2) return false;
    }
  }
  return true;
}

template <typename TIN, typename TC>
int launch_scan_t(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev, void* state,
                  void* out, int reverse, cudaStream_t s) {
  static thread_local DevProgram<TC> dp;
  make_dev_program<TC>(prog, g.ndim, dp);
  ScanArgs a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  int64_t n_state = 1;
  for (int d = 0; d < g.ndim; ++d) {
    a.padlo[d] = g.lo[d] > 0 ? g.lo[d] : 0;
    a.pshape[d] = g.in_shape[d] + a.padlo[d] + (g.hi[d] > 0 ? g.hi[d] : 0);
    n_state *= a.pshape[d];
  }
  a.n_win = prod(g.out_shape, g.ndim);
  if (a.n_win <= 0) return KX_OK;
  {
    int64_t blocks = (n_state + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    scan_init_state_kernel<TIN><<<(unsigned)blocks, kThreads, 0, s>>>(a, (const TIN*)in, (TIN*)state, n_state);
    int rc = after_launch("scan_init_state_kernel");
    if (rc) return rc;
  }
  if (plane_parallel_ok(g, prog)) {
    const int64_t plane = a.n_win / g.out_shape[0];
    int64_t blocks = (plane + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    for (int64_t step = 0; step < g.out_shape[0]; ++step) {
      const int64_t j0 = reverse ? g.out_shape[0] - 1 - step : step;
      scan_plane_kernel<TIN, TC><<<(unsigned)blocks, kThreads, 0, s>>>(a, dp, (TIN*)state, (TIN*)out,
                                                                       (const TC*)coef_dev, j0, plane);
      int rc = after_launch("scan_plane_kernel");
      if (rc) return rc;
    }
    return KX_OK;
  }
  scan_sequential_kernel<TIN, TC><<<1, 32, 0, s>>>(a, dp, (TIN*)state, (TIN*)out, (const TC*)coef_dev, reverse);
  return after_launch("scan_sequential_kernel");
}

bool program_has_float_coefs(const KxProgram& prog) {
  for (int f = 0; f < prog.n_funcs; ++f)
    if (prog.func[f].kind == KX_AFFINE && !prog.func[f].coef_is_int) return true;
  return false;
}

}  // namespace

// --------------------------------------------------------------- launchers ----

int launch_generic_map(const MapArgs& a, const KxProgram& prog, cudaStream_t s) {
  const int in = a.g.in_dtype, out = a.g.out_dtype;
  if (in == KX_F32 && out == KX_F32) return launch_map_t<float, float>(a, prog, s);
  if (in == KX_I32 && out == KX_I32) return launch_map_t<int32_t, int32_t>(a, prog, s);
  if (in == KX_I32 && out == KX_F32) return launch_map_t<int32_t, float>(a, prog, s);
  if (in == KX_F32 && out == KX_I32) return launch_map_t<float, int32_t>(a, prog, s);
  return fail(KX_ERR_UNSUPPORTED, "dtype combination in=%d out=%d is not supported", in, out);
}

// ---------------------------------------------------- VJP (input), strided windows ----
// Cotangent of the input of a STRIDED dense window (average / sum pooling, strided weighted windows): the gradient of
// the README's avg-pool example.  generic_vjp_input_kernel tests every tap of every input element with 64-bit divisions
// (0.013-0.018 of the HBM roofline on 3x3 stride-2 pooling of 8192^2); here a thread owns 4 consecutive input elements
// of a row and walks ONLY the windows that contain them: input coordinate c of axis d lies in window j at tap t iff
// j * S + t = c + lo, so t runs over (c + lo) mod S, +S, ... < K while j runs down from (c + lo) / S -- at most
// ceil(K / S) windows per axis, 32-bit arithmetic throughout.  Same accumulation as the generic kernel (taps in
// row-major order, one FMA each, the forward's final division last), so the two agree bit for bit.
constexpr int kPVMaxTaps = 128;
struct PVArgs {
  const float* g;
  float* dx;
  int in_shape[3], out_shape[3], K[3], S[3], lo[3];   // trailing three axes (missing leading ones: size 1, K = S = 1)
  int64_t in_elems, out_elems;                        // per batch item
  float post_div;
  float coef[kPVMaxTaps];                             // dense [K0][K1][K2]
};

// SS > 0: the stride of the last two axes is the compile-time constant SS (division = shifts / a multiply); 0: run time
// KK > 0 (needs SS > 0): the window of the last two axes is KK x KK: the tap loops unroll into ceil(KK / SS) predicated steps
template <int VEC, int SS, int KK>
__global__ void __launch_bounds__(kThreads)
strided_vjp_input_kernel(const __grid_constant__ PVArgs a) {
  const int S1 = SS > 0 ? SS : a.S[1], S2 = SS > 0 ? SS : a.S[2];
  const int K1 = KK > 0 ? KK : a.K[1], K2 = KK > 0 ? KK : a.K[2];
  constexpr int NI = (KK > 0 && SS > 0) ? (KK + SS - 1) / SS : 1;
  const int W = a.in_shape[2], H = a.in_shape[1];
  const unsigned nvec = (unsigned)(a.in_elems / VEC);
  const float* __restrict__ g = a.g + (int64_t)blockIdx.y * a.out_elems;
  float* __restrict__ dx = a.dx + (int64_t)blockIdx.y * a.in_elems;
  for (unsigned v = blockIdx.x * kThreads + threadIdx.x; v < nvec; v += gridDim.x * kThreads) {
    const unsigned r = v * VEC;
    const int x = (int)(r % (unsigned)W);
    const unsigned yz = r / (unsigned)W;
    const int y = (int)(yz % (unsigned)H), z = (int)(yz / (unsigned)H);
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    const int qz = z + a.lo[0], qy = y + a.lo[1];
    if (qz >= 0 && qy >= 0) {
      const int tz0 = a.S[0] == 1 ? 0 : qz % a.S[0], jz0 = a.S[0] == 1 ? qz : qz / a.S[0];
      for (int tz = tz0, jz = jz0; tz < a.K[0] && jz >= 0; tz += a.S[0], --jz) {
        if (jz >= a.out_shape[0]) continue;
        if constexpr (KK > 0 && SS > 0) {
          const int ty0 = qy % SS, jy0 = qy / SS;
#pragma unroll
          for (int iy = 0; iy < NI; ++iy) {
            const int ty = ty0 + iy * SS, jy = jy0 - iy;
            if (ty >= KK || jy < 0 || jy >= a.out_shape[1]) continue;
            const float* grow = g + ((int64_t)jz * a.out_shape[1] + jy) * a.out_shape[2];
            const float* crow = a.coef + (tz * KK + ty) * KK;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const int qx = x + e + a.lo[2];
              if (qx < 0) continue;
              const int tx0 = qx % SS, jx0 = qx / SS;
#pragma unroll
              for (int ix = 0; ix < NI; ++ix) {
                const int tx = tx0 + ix * SS, jx = jx0 - ix;
                if (tx < KK && jx >= 0 && jx < a.out_shape[2]) acc[e] = fmaf(crow[tx], __ldg(grow + jx), acc[e]);
              }
            }
          }
        } else {
          for (int ty = qy % S1, jy = qy / S1; ty < K1 && jy >= 0; ty += S1, --jy) {
            if (jy >= a.out_shape[1]) continue;
            const float* grow = g + ((int64_t)jz * a.out_shape[1] + jy) * a.out_shape[2];
            const float* crow = a.coef + (tz * K1 + ty) * K2;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const int qx = x + e + a.lo[2];
              if (qx < 0) continue;
              for (int tx = qx % S2, jx = qx / S2; tx < K2 && jx >= 0; tx += S2, --jx)
                if (jx < a.out_shape[2]) acc[e] = fmaf(crow[tx], __ldg(grow + jx), acc[e]);
            }
          }
        }
      }
    }
    if (a.post_div != 0.f) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] = acc[e] / a.post_div;
    }
    if (VEC == 4) {
      __stcs(reinterpret_cast<float4*>(dx + r), make_float4(acc[0], acc[1], acc[2], acc[3]));
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) dx[r + e] = acc[e];
    }
  }
}

// dense host-coefficient windows over the trailing <= 3 axes, pure batch axes in front; anything else -> generic kernel
int try_strided_vjp_input(const KxGeom& g, const KxProgram& p, const void* gout, const void* coef_dev, void* dx,
                          cudaStream_t s) {
  if (coef_dev || p.n_funcs != 1 || p.func[0].kind != KX_AFFINE || getenv("KX_NO_STRIDED_VJP")) return KX_ERR_UNSUPPORTED;
  const int nd = g.ndim, nw = nd < 3 ? nd : 3, lead = nd - nw;
  int64_t batch = 1;
  for (int d = 0; d < lead; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    batch *= g.in_shape[d];
  }
  static thread_local PVArgs a;
  memset(&a, 0, sizeof(a));
  int64_t ntap = 1;
  a.in_elems = a.out_elems = 1;
  for (int i = 0; i < 3; ++i) {
    const int d = lead + i - (3 - nw);
    if (i < 3 - nw) {
      a.in_shape[i] = a.out_shape[i] = a.K[i] = a.S[i] = 1;
      a.lo[i] = 0;
    } else {
      if (g.in_shape[d] > 0x3fffffff || g.out_shape[d] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
      a.in_shape[i] = (int)g.in_shape[d];
      a.out_shape[i] = (int)g.out_shape[d];
      a.K[i] = g.ksize[d];
      a.S[i] = g.stride[d];
      a.lo[i] = g.lo[d];
    }
    ntap *= a.K[i];
    a.in_elems *= a.in_shape[i];
    a.out_elems *= a.out_shape[i];
  }
  const KxFunc& f = p.func[0];
  if (ntap > kPVMaxTaps || f.tap_end - f.tap_begin != ntap) return KX_ERR_UNSUPPORTED;      // dense windows only
  if (a.in_elems <= 0 || a.in_elems > 0xfffffff0ll || batch > 65535 || batch < 1) return KX_ERR_UNSUPPORTED;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    int idx = 0;
    for (int i = 3 - nw; i < 3; ++i) idx = idx * a.K[i] + p.tap_off[t][lead + i - (3 - nw)];
    if (t - f.tap_begin != idx) return KX_ERR_UNSUPPORTED;   // row-major tap order = the generic kernel's accumulation order
    a.coef[idx] = (float)p.coef[t];
  }
  a.g = (const float*)gout;
  a.dx = (float*)dx;
  a.post_div = (float)f.post_div;
  if (a.out_elems == 0) {
    KX_CUDA_CHECK(cudaMemsetAsync(dx, 0, (size_t)(a.in_elems * batch) * 4, s));
    return KX_OK;
  }
  const bool vec = a.in_shape[2] % 4 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0;
  const int64_t nvec = a.in_elems / (vec ? 4 : 1);
  int64_t blocks = (nvec + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)batch);
  const int ss = (a.S[1] == a.S[2] || a.in_shape[1] == 1) ? a.S[2] : 0;   // 1-D windows: axis 1 has one row, any stride does
  if (a.in_shape[1] == 1) a.S[1] = a.S[2];
  const int kk = (a.K[1] == a.K[2] && a.in_shape[1] > 1) ? a.K[2] : 0;
  if (vec && ss == 2 && kk == 2) strided_vjp_input_kernel<4, 2, 2><<<grid, kThreads, 0, s>>>(a);
  else if (vec && ss == 2 && kk == 3) strided_vjp_input_kernel<4, 2, 3><<<grid, kThreads, 0, s>>>(a);
  else if (vec && ss == 2 && kk == 4) strided_vjp_input_kernel<4, 2, 4><<<grid, kThreads, 0, s>>>(a);
  else if (vec && ss == 3 && kk == 3) strided_vjp_input_kernel<4, 3, 3><<<grid, kThreads, 0, s>>>(a);
  else if (vec && ss == 4 && kk == 4) strided_vjp_input_kernel<4, 4, 4><<<grid, kThreads, 0, s>>>(a);
  else if (vec && ss == 2) strided_vjp_input_kernel<4, 2, 0><<<grid, kThreads, 0, s>>>(a);
  else if (vec && ss == 3) strided_vjp_input_kernel<4, 3, 0><<<grid, kThreads, 0, s>>>(a);
  else if (vec && ss == 4) strided_vjp_input_kernel<4, 4, 0><<<grid, kThreads, 0, s>>>(a);
  else if (vec) strided_vjp_input_kernel<4, 0, 0><<<grid, kThreads, 0, s>>>(a);
  else if (ss == 2) strided_vjp_input_kernel<1, 2, 0><<<grid, kThreads, 0, s>>>(a);
  else strided_vjp_input_kernel<1, 0, 0><<<grid, kThreads, 0, s>>>(a);
  return after_launch("strided_vjp_input_kernel");
}

int launch_generic_vjp_input(const KxGeom& g, const KxProgram& prog, const void* gout, const void* coef_dev,
                             void* dx, cudaStream_t s) {
  {
    const int rc = try_strided_vjp_input(g, prog, gout, coef_dev, dx, s);
    if (rc != KX_ERR_UNSUPPORTED) return rc;
  }
  static thread_local DevProgram<float> dp;
  make_dev_program<float>(prog, g.ndim, dp);
  const int64_t n_in = prod(g.in_shape, g.ndim);
  if (n_in <= 0) return KX_OK;
  int64_t blocks = (n_in + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  generic_vjp_input_kernel<float><<<(unsigned)blocks, kThreads, 0, s>>>(g, dp, (const float*)gout,
                                                                       (const float*)coef_dev, (float*)dx, n_in);
  return after_launch("generic_vjp_input_kernel");
}

int64_t generic_vjp_coef_scratch_bytes(const KxGeom& g, const KxProgram& prog) {
  const int64_t n_win = prod(g.out_shape, g.ndim);
  return (int64_t)vjp_coef_blocks(n_win) * prog.n_taps * (int64_t)sizeof(float);
}

int launch_generic_vjp_coef(const KxGeom& g, const KxProgram& prog, const void* x, const void* gout,
                            void* scratch, void* dcoef, cudaStream_t s) {
  static thread_local DevProgram<float> dp;
  make_dev_program<float>(prog, g.ndim, dp);
  const int64_t n_win = prod(g.out_shape, g.ndim);
  const int ntaps = prog.func[0].tap_end - prog.func[0].tap_begin;
  if (ntaps <= 0) return KX_OK;
  const int nb = vjp_coef_blocks(n_win);
  dim3 grid(nb, ntaps);
  generic_vjp_coef_stage1<<<grid, kThreads, 0, s>>>(g, dp, (const float*)x, (const float*)gout, (float*)scratch,
                                                    n_win);
  int rc = after_launch("generic_vjp_coef_stage1");
  if (rc) return rc;
  generic_vjp_coef_stage2<<<ntaps, 32, 0, s>>>((const float*)scratch, (float*)dcoef, nb, (float)prog.func[0].post_div);
  return after_launch("generic_vjp_coef_stage2");
}

int64_t scan_state_elems(const KxGeom& g) {
  int64_t n = 1;
  for (int d = 0; d < g.ndim; ++d)
    n *= g.in_shape[d] + (g.lo[d] > 0 ? g.lo[d] : 0) + (g.hi[d] > 0 ? g.hi[d] : 0);
  return n;
}

int launch_generic_scan(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev,
                        void* state, void* out, int reverse, cudaStream_t s) {
  const bool fl = program_has_float_coefs(prog);  // coef_dev carries the compute type
  if (g.in_dtype == KX_F32) return launch_scan_t<float, float>(g, prog, in, coef_dev, state, out, reverse, s);
  if (g.in_dtype == KX_I32 && !fl)
    return launch_scan_t<int32_t, int32_t>(g, prog, in, coef_dev, state, out, reverse, s);
  if (g.in_dtype == KX_I32) return launch_scan_t<int32_t, float>(g, prog, in, coef_dev, state, out, reverse, s);
  return fail(KX_ERR_UNSUPPORTED, "scan dtype %d is not supported", g.in_dtype);
}

}  // namespace kx
