// stencil2d_kernel instantiations for 9x9 windows: separable forms only (see kx_stencil2d_impl.cuh).
#include "kx_stencil2d_impl.cuh"

namespace kx {
namespace s2 {

int s2_launch_k9(const S2Args<float>& a, int kx, int op, int batch, int vw, cudaStream_t s) {
  return kx == 9 ? launch_op<float, 9, 9>(a, op, batch, vw, s) : KX_ERR_UNSUPPORTED;
}
int s2_launch_k9(const S2Args<int32_t>& a, int kx, int op, int batch, int vw, cudaStream_t s) {
  return kx == 9 ? launch_op<int32_t, 9, 9>(a, op, batch, vw, s) : KX_ERR_UNSUPPORTED;
}

}  // namespace s2
}  // namespace kx
