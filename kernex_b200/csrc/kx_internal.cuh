// Internal launcher interfaces shared by the translation units of libkx_b200.
#pragma once

#include "kx_common.cuh"

namespace kx {

// Device-side form of KxProgram with coefficients already in the compute type.
template <typename TC>
struct DevProgram {
  int32_t n_funcs, n_keys, n_taps, ndim;
  struct Func {
    int32_t kind, tap_begin, tap_end;
    TC bias;
    float post_div;
  } func[KX_MAX_FUNCS];
  KxKey key[KX_MAX_KEYS];
  int16_t tap_off[KX_MAX_TAPS][KX_MAX_DIMS];
  TC coef[KX_MAX_TAPS];
};

template <typename TC>
inline void make_dev_program(const KxProgram& p, int ndim, DevProgram<TC>& d) {
  d.n_funcs = p.n_funcs;
  d.n_keys = p.n_keys;
  d.n_taps = p.n_taps;
  d.ndim = ndim;
  for (int f = 0; f < p.n_funcs; ++f) {
    d.func[f].kind = p.func[f].kind;
    d.func[f].tap_begin = p.func[f].tap_begin;
    d.func[f].tap_end = p.func[f].tap_end;
    d.func[f].bias = (TC)p.func[f].bias;
    d.func[f].post_div = (float)p.func[f].post_div;
  }
  for (int k = 0; k < p.n_keys; ++k) d.key[k] = p.key[k];
  for (int t = 0; t < p.n_taps; ++t) {
    for (int a = 0; a < KX_MAX_DIMS; ++a) d.tap_off[t][a] = p.tap_off[t][a];
    d.coef[t] = (TC)p.coef[t];
  }
}

// ---- region -> function selection (utils.py:338-382 + lax.switch clamp) ----
template <typename TC>
__device__ __forceinline__ int select_func(const DevProgram<TC>& p, const int64_t* centre) {
  if (p.n_funcs <= 1) return 0;
  // keys are stored grouped by function in insertion order: scanning them backwards
  // finds the last-inserted function with any matching key
  for (int k = p.n_keys - 1; k >= 0; --k) {
    const KxKey& key = p.key[k];
    bool ok = true;
    for (int d = 0; d < key.n_items; ++d) {
      const KxKeyItem& it = key.item[d];
      const int64_t c = centre[d];
      bool m = true;
      if (it.kind == KX_KEY_EQ) m = (c == it.a);
      else if (it.kind == KX_KEY_RANGE) m = (it.a <= c) && (c < it.b);
      else if (it.kind == KX_KEY_RANGE_STEP) m = (it.a <= c) && (c < it.b) && (c % it.step == 0);
      ok = ok && m;
    }
    if (ok) return key.func;
  }
  return 0;  // nothing matched: first-inserted function
}

struct Strides {
  int64_t v[KX_MAX_DIMS];
};

__device__ __forceinline__ Strides suffix_strides(const int64_t* shape, int nd) {
  Strides s;
  int64_t acc = 1;
#pragma unroll
  for (int d = KX_MAX_DIMS - 1; d >= 0; --d) {
    if (d < nd) {
      s.v[d] = acc;
      acc *= shape[d];
    } else {
      s.v[d] = 0;
    }
  }
  return s;
}

struct MapArgs {
  KxGeom g;
  const void* in;
  const void* coef_dev;
  void* out;
  int64_t n_win;               // windows per batch item
  int64_t in_elems;            // input elements per batch item
  int64_t out_batch_stride;    // output elements per batch item
  int64_t coef_batch_stride;
  int64_t batch;
  int64_t out_pitch[KX_MAX_DIMS];  // element pitch of the window index along each axis
  int64_t out_base;
  int32_t fn_elems;            // 1, or n_taps for KX_GATHER
  int32_t in_bcast;            // 1: all batch items read the same input array (stencil2d only; others must decline)
  // offset maps (smap) fused into one pass: windows whose centre lies outside [box_lo, box_hi) copy the
  // centre input element instead of evaluating the function ('same'-style geometry only)
  int32_t use_box;
  int64_t box_lo[KX_MAX_DIMS], box_hi[KX_MAX_DIMS];
};

// generic (any rank <= 4, stride, border, multi-function) kernels -- kx_generic.cu
int launch_generic_map(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int launch_generic_vjp_input(const KxGeom& g, const KxProgram& prog, const void* gout, const void* coef_dev,
                             void* dx, cudaStream_t s);
int64_t generic_vjp_coef_scratch_bytes(const KxGeom& g, const KxProgram& prog);
int launch_generic_vjp_coef(const KxGeom& g, const KxProgram& prog, const void* x, const void* gout,
                            void* scratch, void* dcoef, cudaStream_t s);
int64_t scan_state_elems(const KxGeom& g);
int launch_generic_scan(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev,
                        void* state, void* out, int reverse, cudaStream_t s);

// fast paths; each returns KX_ERR_UNSUPPORTED (without touching the error string)
// when its template does not cover the request so the caller can fall through
int try_stencil2d(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int try_stencil2d_mesh(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int try_stencil3d(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int try_stencil3d_star(const MapArgs& a, const KxProgram& prog, int64_t batch, cudaStream_t s);
int try_patch(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int try_patch2d(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int try_patch3d(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
// general stride-1 weight gradient (kx_wgrad.cu)
int64_t wgrad_rows_scratch_bytes(const KxGeom& g, const KxProgram& prog);
int try_wgrad_rows(const KxGeom& g, const KxProgram& prog, const void* x, const void* gout, void* scratch, void* dcoef,
                   cudaStream_t s);
int try_conv2d(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int try_convc(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int try_conv_dense3d(const MapArgs& a, const KxProgram& prog, int64_t batch, cudaStream_t s);
int try_stencil3d_dense(const MapArgs& a, const KxProgram& prog, int64_t batch, cudaStream_t s);
int try_pool2d(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
// wide separable windows (kx_sep2d.cu), requested by kx_stencil2d.cu once it has factored the coefficient table
struct Sep2dReq {
  const float* in;
  float* out;
  int64_t in_batch, out_batch, batch;
  int hin, win, hout, wout, ly, lx;
  int KY, KX, th;
  float bias, post_mul;
  float cy[15], cx[15];
};
int try_sep2d(const Sep2dReq& req, cudaStream_t s);
int try_pool3d(const MapArgs& a, const KxProgram& prog, cudaStream_t s);
int64_t conv_wgrad_scratch_bytes(const KxGeom& g, const KxProgram& prog);
int try_conv_wgrad(const KxGeom& g, const KxProgram& prog, const void* x, const void* gout, void* scratch,
                   void* dcoef, cudaStream_t s);
int launch_scan_wavefront(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev, void* state,
                          void* out, int reverse, cudaStream_t s);
int try_scan_rowmarch(const KxGeom& g, const KxProgram& prog, const void* in, const void* coef_dev,
                      void* out, int reverse, cudaStream_t s);
int scan_rowmarch_shard(const KxGeom& g, const KxProgram& prog, const void* coef_dev, void* out, int reverse,
                        int64_t col_base, int64_t gnx, cudaStream_t s, int row_begin, int row_end);
int scan_rowmarch_rows_per_launch();

}  // namespace kx
