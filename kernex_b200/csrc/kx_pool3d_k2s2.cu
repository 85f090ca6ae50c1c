// pool3d_kernel instantiations for 2x2x2 windows with stride 2 (see kx_pool3d_impl.cuh).
#include "kx_pool3d_impl.cuh"

namespace kx {
namespace p3 {
KX_POOL3_INSTANTIATE(2, 2)
}  // namespace p3
}  // namespace kx
