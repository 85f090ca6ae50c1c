// pool3d_kernel instantiations for 3x3x3 windows with stride 2 (see kx_pool3d_impl.cuh).
#include "kx_pool3d_impl.cuh"

namespace kx {
namespace p3 {
KX_POOL3_INSTANTIATE(3, 2)
}  // namespace p3
}  // namespace kx
