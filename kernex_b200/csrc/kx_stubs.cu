// Placeholders for fast paths that are not written yet: they decline every request so the
// generic kernels run.  Each is replaced by a real translation unit as it lands.
#include "kx_internal.cuh"
namespace kx {
#ifndef KX_HAVE_ROWMARCH
int try_scan_rowmarch(const KxGeom&, const KxProgram&, const void*, const void*, void*, int, cudaStream_t) {
  return KX_ERR_UNSUPPORTED;
}
#endif
}  // namespace kx
