// Intentionally empty: every fast path declared in kx_internal.cuh now has a real
// translation unit (kx_stencil2d.cu, kx_stencil3d.cu, kx_patch.cu, kx_scan_rowmarch.cu).
#include "kx_internal.cuh"
