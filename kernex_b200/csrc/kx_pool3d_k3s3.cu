// pool3d_kernel instantiations for 3x3x3 windows with stride 3 (see kx_pool3d_impl.cuh).
#include "kx_pool3d_impl.cuh"

namespace kx {
namespace p3 {
KX_POOL3_INSTANTIATE(3, 3)
}  // namespace p3
}  // namespace kx
