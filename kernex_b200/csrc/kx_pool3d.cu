// Dispatch of the 3-D pooling fast path (kernel template in kx_pool3d_impl.cuh, instantiations in kx_pool3d_k*s*.cu).
#include "kx_pool3d_impl.cuh"

namespace kx {

int try_pool3d(const MapArgs& m, const KxProgram& p, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  if (nd < 3 || p.n_funcs != 1 || m.fn_elems != 1 || p.func[0].kind == KX_GATHER) return KX_ERR_UNSUPPORTED;
  if (g.in_dtype != g.out_dtype || m.coef_dev) return KX_ERR_UNSUPPORTED;
  const KxFunc& f = p.func[0];
  if (f.kind == KX_AFFINE && f.post_div != 0.0 && g.in_dtype != KX_F32) return KX_ERR_UNSUPPORTED;
  int64_t lead = 1;
  for (int d = 0; d < nd - 3; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return KX_ERR_UNSUPPORTED;
    lead *= g.in_shape[d];
  }
  const int K = g.ksize[nd - 1], S = g.stride[nd - 1];
  for (int d = nd - 3; d < nd; ++d) {
    if (g.ksize[d] != K || g.stride[d] != S) return KX_ERR_UNSUPPORTED;          // cubic windows and strides only
    if (g.in_shape[d] > 0x3fffffff || g.out_shape[d] > 0x3fffffff) return KX_ERR_UNSUPPORTED;
  }
  if (f.tap_end - f.tap_begin != K * K * K) return KX_ERR_UNSUPPORTED;            // dense windows only
  // TMA bulk copies need 16-byte aligned rows
  if (g.in_shape[nd - 1] % 4 != 0 || (reinterpret_cast<uintptr_t>(m.in) & 15) != 0) return KX_ERR_UNSUPPORTED;
  const int64_t batch = m.batch * lead;
  if (K == 2 && S == 2) return p3::run_pool3_ks<2, 2>(m, p, batch, s);
  if (K == 3 && S == 2) return p3::run_pool3_ks<3, 2>(m, p, batch, s);
  if (K == 3 && S == 3) return p3::run_pool3_ks<3, 3>(m, p, batch, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
