// Dense 3x3x3 windows (27-point stencils, conv3d-as-kmap with HOST weights), stride 1, f32 / i32, rows 16-byte
// aligned.  Measured on B200 at 1024^3: 626 GLUP/s = 0.77 of the HBM roofline, against 489 (0.60) for the conv row
// ring with the three z planes as channels (kx_conv2d.cu run_dense3d, where every input plane travels L2 -> SM three
// times); that path still serves device-resident weights and is selected for host weights by KX_NO_DENSE3D_PLANES=1
// (cross-check).  This kernel keeps the
// 2.5-D blocking of kx_stencil3d_star.cu -- a block owns a 16 x 128 tile and marches 32 output planes, every input
// plane tile is fetched ONCE with 16-byte cp.async into a 3-stage shared-memory ring two planes ahead of the math --
// and is PLANE-STATIONARY: when input plane p lands, a thread reads its 4 rows x 6 columns once and feeds three
// accumulator sets,
//     out[p]   = bias + W[kz=0] * P      (started)
//     out[p-1] +=       W[kz=1] * P
//     out[p-2] +=       W[kz=2] * P      (complete: stored)
// so there is no queue of planes in registers (24 plane values + 24 accumulators per thread) and the 27 FMAs per
// output run as packed FFMA2 for fp32.  The plane loop is unrolled by three so the accumulator sets rotate by
// renaming.  Accumulation order per output: bias, then the window in row-major (kz, ky, kx) order = the order of
// stencil3d_kernel and of the oracle's tap list.
#include <type_traits>

#include "kx_internal.cuh"

namespace kx {

namespace {

// tile 8 rows x 256 columns (two warps side by side, 1 KB contiguous per row segment), 256 threads, 64-plane z chunks,
// wave-sized tile groups whose z chunk runs fastest: the geometry measured best on kx_stencil3d_star.cu (see there)
constexpr int kDWX = 2, kDYB = 8, kDRY = 2;
constexpr int kDX = 128 * kDWX, kDY = kDYB / kDWX * kDRY;
constexpr int kDThreads = 32 * kDYB;
constexpr int kDStages = 3;
constexpr int kDZC = 64;

template <typename T>
struct DenseArgs {
  const T* in;
  T* out;
  int64_t in_batch, out_batch;
  int din, hin, win, dout, hout, wout;
  int lz, ly, lx;
  int nzc, zc;
  int tiles_x, ntiles, group;   // block -> (z chunk, tile) mapping, see block_coords in kx_stencil3d_star.cu
  T bias;
  T c[27];              // row-major (kz, ky, kx)
};

__device__ __forceinline__ void d_cp_async16(unsigned smem_addr, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void d_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void d_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
__device__ __forceinline__ void d_lds4(const T* p, T (&d)[4]) {
  const int4 raw = *reinterpret_cast<const int4*>(p);
  d[0] = reinterpret_cast<const T&>(raw.x);
  d[1] = reinterpret_cast<const T&>(raw.y);
  d[2] = reinterpret_cast<const T&>(raw.z);
  d[3] = reinterpret_cast<const T&>(raw.w);
}

template <typename T>
__device__ __forceinline__ void d_store4(T* o, const T (&acc)[4], int n_valid) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
  if (n_valid == 4 && (addr & 15) == 0) {
    int4 raw;
    raw.x = reinterpret_cast<const int&>(acc[0]);
    raw.y = reinterpret_cast<const int&>(acc[1]);
    raw.z = reinterpret_cast<const int&>(acc[2]);
    raw.w = reinterpret_cast<const int&>(acc[3]);
    __stcs(reinterpret_cast<int4*>(o), raw);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n_valid) o[i] = acc[i];
  }
}

template <typename T, int SHIFT>
__global__ void __launch_bounds__(kDThreads)
stencil3d_dense_kernel(const __grid_constant__ DenseArgs<T> a) {
  constexpr int NV = (SHIFT + 6 + 3) / 4;          // aligned 4-vectors covering the 6 needed columns
  constexpr int SW = kDX + 4 * NV;                  // tile width (columns), multiple of 4
  constexpr int SR = kDY + 2;
  constexpr int CH = SW / 4;
  constexpr int NCH = (SR * CH + kDThreads - 1) / kDThreads;   // 16-byte chunks per thread per plane
  __shared__ __align__(16) T smem[kDStages * SR * SW];

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  // linear block id -> (batch, tile group, z chunk, tile in group): the z chunk of a wave-sized group of tiles runs
  // faster than the group index, so the planes two consecutive chunks share are L2 hits
  int b, zchunk, tile_x, tile_y;
  {
    const unsigned per_batch = (unsigned)a.ntiles * (unsigned)a.nzc;
    unsigned l = blockIdx.x;
    b = (int)(l / per_batch);
    l -= (unsigned)b * per_batch;
    const unsigned gsz = (unsigned)a.group * (unsigned)a.nzc;
    const unsigned ngroups = ((unsigned)a.ntiles + a.group - 1) / a.group;
    unsigned g = l / gsz;
    if (g > ngroups - 1) g = ngroups - 1;
    const unsigned r = l - g * gsz;
    const unsigned in_group = min((unsigned)a.group, (unsigned)a.ntiles - g * a.group);
    zchunk = (int)(r / in_group);
    const unsigned tile = g * a.group + (r - (unsigned)zchunk * in_group);
    tile_y = (int)(tile / (unsigned)a.tiles_x);
    tile_x = (int)(tile - (unsigned)tile_y * a.tiles_x);
  }
  const int z0 = zchunk * a.zc;
  const int z1 = min(z0 + a.zc, a.dout);
  const int x0 = tile_x * kDX, y0 = tile_y * kDY;
  const T* __restrict__ in = a.in + (int64_t)b * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)b * a.out_batch;
  const int gz0 = z0 - a.lz;                        // global plane of ring plane 0
  const int nplanes = (z1 - z0) + 2;
  const int64_t plane_elems = (int64_t)a.hin * a.win;

  int src_off[NCH];        // element offset inside a plane, or -1 = outside the array (zero fill)
  unsigned dst_off[NCH];   // byte offset inside a stage
  {
    const int gx0 = x0 - a.lx - SHIFT;             // global column of tile column 0 (multiple of 4)
    const int gy0 = y0 - a.ly;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int i = tid + k * kDThreads;
      const int r = i / CH, c = (i - r * CH) * 4;
      const int gy = gy0 + r, gx = gx0 + c;
      const bool ok = (i < SR * CH) && (gy >= 0) && (gy < a.hin) && (gx >= 0) && (gx + 3 < a.win);
      src_off[k] = ok ? gy * a.win + gx : -1;
      dst_off[k] = (i < SR * CH) ? (unsigned)((r * SW + c) * sizeof(T)) : 0xffffffffu;
    }
  }
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);

  auto issue = [&](int pl) {
    const int gz = gz0 + pl;
    const bool z_ok = (gz >= 0) && (gz < a.din);
    const T* src = in + (int64_t)gz * plane_elems;
    const unsigned dst = smem_base + (unsigned)((pl % kDStages) * SR * SW * sizeof(T));
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      if (dst_off[k] != 0xffffffffu) {
        const bool ok = z_ok && (src_off[k] >= 0);
        d_cp_async16(dst + dst_off[k], ok ? (const void*)(src + src_off[k]) : (const void*)in, ok ? 16 : 0);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < kDStages - 1; ++s) {
    if (s < nplanes) issue(s);
    d_cp_commit();
  }

  const int jx = 4 * tx + 128 * (ty % kDWX);   // thread's first output column inside the tile
  const int iy = (ty / kDWX) * kDRY;            // thread's first output row inside the tile
  const int ox = x0 + jx;
  const int n_valid = min(4, a.wout - ox);
  const bool col_ok = ox < a.wout;
  T* orow = out + ((int64_t)z0 * a.hout + (y0 + iy)) * a.wout + ox;   // advanced by one plane per stored plane
  const int64_t out_plane = (int64_t)a.hout * a.wout;

  // rows iy .. iy+RY+1 of a plane tile: the 6 window columns SHIFT+jx .. SHIFT+jx+5 of each
  auto read_plane = [&](int pl, T (&dst)[kDRY + 2][6]) {
    const T* t = smem + (pl % kDStages) * SR * SW + iy * SW + jx;
#pragma unroll
    for (int r = 0; r < kDRY + 2; ++r) {
      T w[4 * NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        T d[4];
        d_lds4(t + r * SW + 4 * q, d);
#pragma unroll
        for (int e = 0; e < 4; ++e) w[4 * q + e] = d[e];
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) dst[r][e] = w[SHIFT + e];
    }
  };

  // acc[s][r][i]: three output planes in flight
  T acc[3][kDRY][4];

  // one kz slice of the window applied to plane P, accumulated into `dst` (INIT: dst starts from the bias)
  auto apply = [&](const T (&P)[kDRY + 2][6], int kz, T (&dst)[kDRY][4], bool init) {
#pragma unroll
    for (int r = 0; r < kDRY; ++r) {
      if (init) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[r][i] = a.bias;
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const T wk = a.c[(kz * 3 + ky) * 3 + kx];
          if constexpr (std::is_same<T, float>::value) {   // two packed FMAs instead of four scalar ones
            ffma2(dst[r][0], dst[r][1], wk, wk, P[r + ky][kx], P[r + ky][kx + 1]);
            ffma2(dst[r][2], dst[r][3], wk, wk, P[r + ky][kx + 2], P[r + ky][kx + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[r][i] = mad_wrap(wk, P[r + ky][kx + i], dst[r][i]);
          }
        }
      }
    }
  };

  // ring step for plane pl with phase PH = pl % 3 (compile time): set (PH) starts output pl, set (PH+2)%3 is output
  // pl-1, set (PH+1)%3 is output pl-2 and completes here
  auto step = [&](auto phase, int pl) {
    constexpr int PH = decltype(phase)::value;
    d_cp_wait<kDStages - 2>();
    __syncthreads();            // plane pl landed for all threads; the stage read last step is free
    if (pl + kDStages - 1 < nplanes) issue(pl + kDStages - 1);
    d_cp_commit();
    T P[kDRY + 2][6];
    read_plane(pl, P);
    apply(P, 2, acc[(PH + 1) % 3], false);
    if (pl >= 2) {
#pragma unroll
      for (int r = 0; r < kDRY; ++r)
        if (col_ok && (y0 + iy + r < a.hout)) d_store4(orow + (int64_t)r * a.wout, acc[(PH + 1) % 3][r], n_valid);
      orow += out_plane;
    }
    apply(P, 1, acc[(PH + 2) % 3], false);
    apply(P, 0, acc[PH], true);
  };

#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int r = 0; r < kDRY; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[s][r][i] = T(0);   // sets touched before their first `init` are never stored

  int pl = 0;
  for (; pl + 2 < nplanes; pl += 3) {   // unrolled by three: the accumulator sets swap roles by name
    step(std::integral_constant<int, 0>{}, pl);
    step(std::integral_constant<int, 1>{}, pl + 1);
    step(std::integral_constant<int, 2>{}, pl + 2);
  }
  if (pl < nplanes) step(std::integral_constant<int, 0>{}, pl);
  if (pl + 1 < nplanes) step(std::integral_constant<int, 1>{}, pl + 1);
}

template <typename T>
int run_dense(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int az = g.ndim - 3, ay = g.ndim - 2, ax = g.ndim - 1;
  if (f.tap_end - f.tap_begin != 27) return KX_ERR_UNSUPPORTED;
  static thread_local DenseArgs<T> a;
  memset(&a, 0, sizeof(a));
  uint32_t seen = 0;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int slot = (p.tap_off[t][az] * 3 + p.tap_off[t][ay]) * 3 + p.tap_off[t][ax];
    if (slot < 0 || slot >= 27 || (seen & (1u << slot))) return KX_ERR_UNSUPPORTED;
    seen |= 1u << slot;
    a.c[slot] = (T)p.coef[t];
  }
  a.in = (const T*)m.in;
  a.out = (T*)m.out;
  a.din = (int)g.in_shape[az];
  a.hin = (int)g.in_shape[ay];
  a.win = (int)g.in_shape[ax];
  a.dout = (int)g.out_shape[az];
  a.hout = (int)g.out_shape[ay];
  a.wout = (int)g.out_shape[ax];
  a.lz = g.lo[az];
  a.ly = g.lo[ay];
  a.lx = g.lo[ax];
  if ((int64_t)a.hin * a.win > 0x7ffffff0ll) return KX_ERR_UNSUPPORTED;  // 32-bit in-plane offsets
  a.in_batch = (int64_t)a.din * a.hin * a.win;
  a.out_batch = (int64_t)a.dout * a.hout * a.wout;
  a.bias = (T)f.bias;
  a.zc = kDZC;
  a.nzc = (a.dout + a.zc - 1) / a.zc;
  if (const char* e = getenv("KX_DENSE3D_ZC")) {   // tuning knob
    a.zc = atoi(e) > 0 ? atoi(e) : a.zc;
    a.nzc = (a.dout + a.zc - 1) / a.zc;
  }
  const int tiles_x = (a.wout + kDX - 1) / kDX, tiles_y = (a.hout + kDY - 1) / kDY;
  const int64_t nblocks = (int64_t)tiles_x * tiles_y * a.nzc * batch;
  if (nblocks > 0x7fffffffll || (int64_t)tiles_x * tiles_y > 0x7fffffll) return KX_ERR_UNSUPPORTED;
  a.tiles_x = tiles_x;
  a.ntiles = tiles_x * tiles_y;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stencil3d_dense_kernel<T, 0>, kDThreads, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 2;
  }
  int gy = sm_count() * per_sm / tiles_x;          // one group = whole rows of tiles filling about one wave
  if (gy < 1) gy = 1;
  a.group = getenv("KX_DENSE3D_NO_GROUPS") ? a.ntiles : gy * tiles_x;
  if (a.group > a.ntiles) a.group = a.ntiles;
  dim3 grid((unsigned)nblocks), block(32, kDYB);
  const int shift = (((-a.lx) % 4) + 4) % 4;
  switch (shift) {
    case 0: stencil3d_dense_kernel<T, 0><<<grid, block, 0, s>>>(a); break;
    case 1: stencil3d_dense_kernel<T, 1><<<grid, block, 0, s>>>(a); break;
    case 2: stencil3d_dense_kernel<T, 2><<<grid, block, 0, s>>>(a); break;
    default: stencil3d_dense_kernel<T, 3><<<grid, block, 0, s>>>(a); break;
  }
  return after_launch("stencil3d_dense_kernel");
}

}  // namespace

// Host-coefficient dense 3x3x3 windows with 16-byte aligned rows (KX_NO_DENSE3D_PLANES=1: decline, for cross-checks).
int try_stencil3d_dense(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  if (getenv("KX_NO_DENSE3D_PLANES")) return KX_ERR_UNSUPPORTED;
  if (g.ksize[nd - 3] != 3 || g.ksize[nd - 2] != 3 || g.ksize[nd - 1] != 3) return KX_ERR_UNSUPPORTED;
  if (m.coef_dev) return KX_ERR_UNSUPPORTED;
  if (g.in_shape[nd - 1] % 4 != 0 || (reinterpret_cast<uintptr_t>(m.in) & 15) != 0) return KX_ERR_UNSUPPORTED;
  if (((int64_t)g.in_shape[nd - 3] * g.in_shape[nd - 2] * g.in_shape[nd - 1]) % 4 != 0 && batch > 1)
    return KX_ERR_UNSUPPORTED;
  if (g.in_dtype == KX_F32) return run_dense<float>(m, p, batch, s);
  if (g.in_dtype == KX_I32 && p.func[0].coef_is_int) return run_dense<int32_t>(m, p, batch, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
