// stencil2d_kernel instantiations for windows 1 row high (see kx_stencil2d_impl.cuh).
#include "kx_stencil2d_impl.cuh"

namespace kx {
namespace s2 {

namespace {
template <typename T>
int launch_row(const S2Args<T>& a, int kx, int op, int batch, int vw, cudaStream_t s) {
  switch (kx) {
    case 1: return launch_op<T, 1, 1>(a, op, batch, vw, s);
    case 2: return launch_op<T, 1, 2>(a, op, batch, vw, s);
    case 3: return launch_op<T, 1, 3>(a, op, batch, vw, s);
    case 4: return launch_op<T, 1, 4>(a, op, batch, vw, s);
    case 5: return launch_op<T, 1, 5>(a, op, batch, vw, s);
    case 6: return launch_op<T, 1, 6>(a, op, batch, vw, s);
    case 7: return launch_op<T, 1, 7>(a, op, batch, vw, s);
    case 8: return launch_op<T, 1, 8>(a, op, batch, vw, s);
    case 9: return launch_op<T, 1, 9>(a, op, batch, vw, s);
    case 11: return launch_op<T, 1, 11>(a, op, batch, vw, s);
    case 13: return launch_op<T, 1, 13>(a, op, batch, vw, s);
    case 15: return launch_op<T, 1, 15>(a, op, batch, vw, s);
  }
  return KX_ERR_UNSUPPORTED;
}
}  // namespace

int s2_launch_k1(const S2Args<float>& a, int kx, int op, int batch, int vw, cudaStream_t s) {
  return launch_row<float>(a, kx, op, batch, vw, s);
}
int s2_launch_k1(const S2Args<int32_t>& a, int kx, int op, int batch, int vw, cudaStream_t s) {
  return launch_row<int32_t>(a, kx, op, batch, vw, s);
}

}  // namespace s2
}  // namespace kx
