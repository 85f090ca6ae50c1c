// Fast path: 3-D 7-point STAR stencils (taps on the three axes through the window centre
// of a 3x3x3 window), stride 1, f32 / i32, rows 16-byte aligned -- BASELINE config 5.
//
// 2.5-D blocking with an asynchronous shared-memory plane ring (first tuned in benchmarks/exp3d.cu):
//   * a block (8 warps) owns an 8-row x 256-column tile and marches kZC = 64 output planes along axis 0;
//   * every input plane tile (10 rows x 264 columns) is fetched ONCE with 16-byte cp.async.cg (LDGSTS: global ->
//     shared, no registers, zero-fill outside the array = the reference's zero padding) into a 3-stage ring, two
//     planes ahead of the math; the per-thread chunk offsets / validity are computed once, outside the plane loop;
//   * each thread owns 4 columns x 2 rows; per plane it reads its rows from shared memory with 128-bit loads and
//     keeps the previous / current plane in registers (z queue); the plane loop is unrolled by two so the queue
//     rotates by renaming instead of moves;
//   * 128-bit streaming stores; the window misalignment (-lo_x) mod 4 and "all seven taps present" are template
//     parameters (the first ncu pass showed the kernel issue-bound at 80 % with per-FMA tap-mask tests: they are
//     compiled out for the full star).
//
// Round-2 measurements on B200 at 2048^3 (one box, back to back; profiles/r2_star_*):
//   16 x 128 tiles, zc 32, blocks in x / y / z-chunk order           720 GLUP/s   DRAM read 36.5 GB (1.063 x algorithmic)
//   + z chunk of a wave-sized tile GROUP runs faster than the group   720          DRAM read 34.6 GB (1.006 x): the
//     (block_coords below)                                                        z-halo planes hit L2 -- but the kernel
//                                                                                 was not limited by the extra bytes
//   + 8 x 256 tiles (1 KB contiguous per row segment, KX_STAR_WX=2)   745
//   + zc 64                                                           765-770     = 0.935-0.94 of the measured copy peak
//   4 x 512 tiles 727-763, a 4-stage ring 666-705, 4 blocks / SM at 64 registers 625, one barrier per two planes 617,
//   cp.async L2::128B / L2::256B prefetch hints: no change.  Without the groups zc 64 gives 749.
//   Per-stage mbarriers instead of wait_group + __syncthreads (cp.async.mbarrier.arrive.noinc for "landed", one
//   arrive per thread for "read"; a warp then waits for data only): 767 vs 770-774 on the same box -- the per-plane
//   block barrier shows up in the stall samples but is not what limits the kernel; removed again.
//   Deeper rings in dynamic shared memory (3 blocks / SM kept): 3 stages 764, 4 stages 695, 5 stages 708, 6 stages 697 --
//   MORE bytes in flight are slower, so the limit is not latency but what DRAM / L2 make of the access stream.
// KX_STAR_WX=1|2|4, KX_STAR_ZC=n, KX_STAR_NO_GROUPS=1 and KX_STAR_GRID3D=1 select the other variants (bit-identical
// results; cross-checks and A/B timing).
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kWarps = 8;                     // 256 threads; a warp covers 128 columns x RY rows
constexpr int kThreads3 = 32 * kWarps;
// tile = (kWarps / WX * RY) rows x (128 * WX) columns: WX = 1, RY = 2 -> 16 x 128; WX = 2, RY = 2 -> 8 x 256
constexpr int kStages = 3;
constexpr int kZC = 64;

template <typename T>
struct StarArgs {
  const T* in;
  T* out;
  int64_t in_batch, out_batch;
  int din, hin, win, dout, hout, wout;
  int lz, ly, lx;
  int nzc, zc;           // z chunks per volume, planes per chunk
  int tiles_x, ntiles;   // xy tiles per plane
  int group;             // tiles per scheduling group (see block_coords)
  uint32_t mask;        // bit t set = tap t present; taps in row-major window order:
  T bias;               // 0:(z-1) 1:(y-1) 2:(x-1) 3:centre 4:(x+1) 5:(y+1) 6:(z+1)
  T c[7];
};

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
__device__ __forceinline__ void lds4(const T* p, T (&d)[4]) {
  const int4 raw = *reinterpret_cast<const int4*>(p);
  d[0] = reinterpret_cast<const T&>(raw.x);
  d[1] = reinterpret_cast<const T&>(raw.y);
  d[2] = reinterpret_cast<const T&>(raw.z);
  d[3] = reinterpret_cast<const T&>(raw.w);
}

template <typename T>
__device__ __forceinline__ void store4(T* o, const T (&acc)[4], int n_valid) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
  if (n_valid == 4 && (addr & 15) == 0) {
    int4 raw;
    raw.x = reinterpret_cast<const int&>(acc[0]);
    raw.y = reinterpret_cast<const int&>(acc[1]);
    raw.z = reinterpret_cast<const int&>(acc[2]);
    raw.w = reinterpret_cast<const int&>(acc[3]);
    __stcs(reinterpret_cast<int4*>(o), raw);
  } else if (n_valid == 4 && (addr & 7) == 0) {
    int2 lo, hi;
    lo.x = reinterpret_cast<const int&>(acc[0]);
    lo.y = reinterpret_cast<const int&>(acc[1]);
    hi.x = reinterpret_cast<const int&>(acc[2]);
    hi.y = reinterpret_cast<const int&>(acc[3]);
    __stcs(reinterpret_cast<int2*>(o), lo);
    __stcs(reinterpret_cast<int2*>(o) + 1, hi);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n_valid) o[i] = acc[i];
  }
}

// (A variant with a 4-stage ring and one block barrier per TWO planes was measured in round 2: 617 vs 720 GLUP/s at
// 2048^3 -- the extra stage costs a resident block per SM -- and removed.)
// Block -> (batch, z chunk, tile) mapping.  Blocks are dispatched in linear order, so the order decides what the
// ~444 resident blocks (one "wave") share in L2.  Tiles are cut into GROUPS of about one wave (whole rows of tiles: a
// compact patch, so the y / x halos of its tiles are read by blocks of the same wave), and inside a group the z chunk
// runs FASTER than the group index: wave w marches chunk k of the group, wave w+1 chunk k+1 of the SAME tiles.  The
// two input planes consecutive chunks share (2 of zc+2: 6.3 % of the reads at zc = 32, ncu r1b: 36.5 GB read vs
// 34.4 GB algorithmic) were the last planes the previous wave touched and are still in L2.
template <typename T>
__device__ __forceinline__ void block_coords(const StarArgs<T>& a, int& b, int& zchunk, int& tx_tile, int& ty_tile) {
  if (a.group == 0) {   // 3-D grid launch: x tiles, y tiles, batch * z chunks (A/B switch KX_STAR_GRID3D=1)
    b = blockIdx.z / a.nzc;
    zchunk = blockIdx.z - b * a.nzc;
    tx_tile = blockIdx.x;
    ty_tile = blockIdx.y;
    return;
  }
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
  ty_tile = (int)(tile / (unsigned)a.tiles_x);
  tx_tile = (int)(tile - (unsigned)ty_tile * a.tiles_x);
}

template <typename T, int SHIFT, bool FULL, int WX, int kRY>
__global__ void __launch_bounds__(kThreads3, kRY == 2 ? 3 : 2)
stencil3d_star_kernel(const __grid_constant__ StarArgs<T> a) {
  constexpr int STG = kStages;
  constexpr int kTX = 128 * WX, kTY = kWarps / WX * kRY;
  constexpr int NV = (SHIFT + 6 + 3) / 4;          // aligned 4-vectors covering the 6 needed columns
  constexpr int SW = kTX + 4 * NV;                  // tile width (columns), multiple of 4
  constexpr int SR = kTY + 2;
  constexpr int CH = SW / 4;
  constexpr int NCH = (SR * CH + kThreads3 - 1) / kThreads3;   // 16-byte chunks per thread per plane
  __shared__ __align__(16) T smem[STG * SR * SW];

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  int b, zchunk, tile_x, tile_y;
  block_coords(a, b, zchunk, tile_x, tile_y);
  const int z0 = zchunk * a.zc;
  const int z1 = min(z0 + a.zc, a.dout);
  const int x0 = tile_x * kTX, y0 = tile_y * kTY;
  const T* __restrict__ in = a.in + (int64_t)b * a.in_batch;
  T* __restrict__ out = a.out + (int64_t)b * a.out_batch;
  const int gz0 = z0 - a.lz;                        // global plane of ring plane 0
  const int nplanes = (z1 - z0) + 2;
  const int64_t plane_elems = (int64_t)a.hin * a.win;

  // this thread's chunks of a plane tile: fixed for the whole march
  int src_off[NCH];        // element offset inside a plane, or -1 = outside the array (zero fill)
  unsigned dst_off[NCH];   // byte offset inside a stage
  {
    const int gx0 = x0 - a.lx - SHIFT;             // global column of tile column 0 (multiple of 4)
    const int gy0 = y0 - a.ly;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int i = tid + k * kThreads3;
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
    const unsigned dst = smem_base + (unsigned)((pl % STG) * SR * SW * sizeof(T));
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      if (dst_off[k] != 0xffffffffu) {
        const bool ok = z_ok && (src_off[k] >= 0);
        cp_async16(dst + dst_off[k], ok ? (const void*)(src + src_off[k]) : (const void*)in, ok ? 16 : 0);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < STG - 1; ++s) {
    if (s < nplanes) issue(s);
    cp_commit();
  }

  const int jx = 4 * tx + 128 * (ty % WX);   // thread's first output column inside the tile
  const int iy = (ty / WX) * kRY;            // thread's first output row inside the tile
  const int ox = x0 + jx;
  const int n_valid = min(4, a.wout - ox);
  const bool col_ok = ox < a.wout;
  T* orow = out + ((int64_t)z0 * a.hout + (y0 + iy)) * a.wout + ox;   // advanced by one plane per step
  const int64_t out_plane = (int64_t)a.hout * a.wout;

  // rows iy .. iy+RY+1 of a plane tile: the 6 window columns SHIFT+jx .. SHIFT+jx+5 of each
  auto read_plane = [&](int pl, T (&dst)[kRY + 2][6]) {
    const T* t = smem + (pl % STG) * SR * SW + iy * SW + jx;
#pragma unroll
    for (int r = 0; r < kRY + 2; ++r) {
      T w[4 * NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        T d[4];
        lds4(t + r * SW + 4 * q, d);
#pragma unroll
        for (int e = 0; e < 4; ++e) w[4 * q + e] = d[e];
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) dst[r][e] = w[SHIFT + e];
    }
  };

  T below[kRY][4];              // plane z-1: centre columns of the thread's own rows
  T bufA[kRY + 2][6], bufB[kRY + 2][6];

  // one ring step: `up` receives plane pl, `mid` holds plane pl-1, `below` plane pl-2
  auto step = [&](int pl, T (&mid)[kRY + 2][6], T (&up)[kRY + 2][6]) {
    cp_wait<STG - 2>();
    __syncthreads();            // plane pl landed for all threads; the stage read last step is free
    if (pl + STG - 1 < nplanes) issue(pl + STG - 1);
    cp_commit();
    read_plane(pl, up);
    if (pl >= 2) {
#pragma unroll
      for (int r = 0; r < kRY; ++r) {
        T acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          T v = a.bias;
          if (FULL || (a.mask & 1u)) v = mad_wrap(a.c[0], below[r][i], v);
          if (FULL || (a.mask & 2u)) v = mad_wrap(a.c[1], mid[r][i + 1], v);
          if (FULL || (a.mask & 4u)) v = mad_wrap(a.c[2], mid[r + 1][i], v);
          if (FULL || (a.mask & 8u)) v = mad_wrap(a.c[3], mid[r + 1][i + 1], v);
          if (FULL || (a.mask & 16u)) v = mad_wrap(a.c[4], mid[r + 1][i + 2], v);
          if (FULL || (a.mask & 32u)) v = mad_wrap(a.c[5], mid[r + 2][i + 1], v);
          if (FULL || (a.mask & 64u)) v = mad_wrap(a.c[6], up[r + 1][i + 1], v);
          acc[i] = v;
        }
        if (col_ok && (y0 + iy + r < a.hout)) store4(orow + (int64_t)r * a.wout, acc, n_valid);
      }
      orow += out_plane;
    }
#pragma unroll
    for (int r = 0; r < kRY; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) below[r][i] = mid[r + 1][i + 1];
    }
  };

  int pl = 0;
  for (; pl + 1 < nplanes; pl += 2) {   // unrolled by two: the plane buffers swap roles by name
    step(pl, bufB, bufA);
    step(pl + 1, bufA, bufB);
  }
  if (pl < nplanes) step(pl, bufB, bufA);
}

template <typename T>
int run_star(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const KxFunc& f = p.func[0];
  const int az = g.ndim - 3, ay = g.ndim - 2, ax = g.ndim - 1;
  StarArgs<T> a;
  memset(&a, 0, sizeof(a));
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    const int oz = p.tap_off[t][az], oy = p.tap_off[t][ay], ox = p.tap_off[t][ax];
    int slot = -1;
    if (oy == 1 && ox == 1) slot = oz == 0 ? 0 : (oz == 1 ? 3 : 6);
    else if (oz == 1 && ox == 1) slot = oy == 0 ? 1 : 5;
    else if (oz == 1 && oy == 1) slot = ox == 0 ? 2 : 4;
    if (slot < 0 || (a.mask & (1u << slot))) return KX_ERR_UNSUPPORTED;  // off-axis or duplicate tap
    a.mask |= 1u << slot;
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
  a.zc = kZC;
  if (const char* e = getenv("KX_STAR_ZC")) a.zc = atoi(e) > 0 ? atoi(e) : a.zc;   // tuning knob
  a.nzc = (a.dout + a.zc - 1) / a.zc;
  int wx = 2, ry = 2;
  if (const char* e = getenv("KX_STAR_WX")) wx = atoi(e) == 2 ? 2 : (atoi(e) == 4 ? 4 : 1);
  const int kTX = 128 * wx, kTY = kWarps / wx * ry;
  const int tiles_x = (a.wout + kTX - 1) / kTX, tiles_y = (a.hout + kTY - 1) / kTY;
  const int64_t nblocks = (int64_t)tiles_x * tiles_y * a.nzc * batch;
  if (nblocks > 0x7fffffffll || (int64_t)tiles_x * tiles_y > 0x7fffffll) return KX_ERR_UNSUPPORTED;
  a.tiles_x = tiles_x;
  a.ntiles = tiles_x * tiles_y;
  // one group = whole rows of tiles filling about one wave of resident blocks (3 per SM)
  const int minb = ry == 2 ? 3 : 2;
  const int wave = sm_count() * minb;
  int gy = wave / tiles_x;
  if (gy < 1) gy = 1;
  a.group = getenv("KX_STAR_NO_GROUPS") ? a.ntiles : gy * tiles_x;              // whole volume = the old x, y, z-chunk order
  if (a.group > a.ntiles) a.group = a.ntiles;
  dim3 grid((unsigned)nblocks), block(32, kWarps);
  if (getenv("KX_STAR_GRID3D") && (int64_t)a.nzc * batch <= 65535 && tiles_y <= 65535) {
    a.group = 0;
    grid = dim3(tiles_x, tiles_y, (unsigned)(a.nzc * batch));
  }
  const int shift = (((-a.lx) % 4) + 4) % 4;
  const bool full = a.mask == 0x7fu;
#define KX_STAR3(S, WXV, RYV)                                                          \
  if (full) stencil3d_star_kernel<T, S, true, WXV, RYV><<<grid, block, 0, s>>>(a);     \
  else stencil3d_star_kernel<T, S, false, WXV, RYV><<<grid, block, 0, s>>>(a);
#define KX_STAR2(S, RYV) \
  if (wx == 2) { KX_STAR3(S, 2, RYV) } else if (wx == 4) { KX_STAR3(S, 4, RYV) } else { KX_STAR3(S, 1, RYV) }
#define KX_STAR(S) KX_STAR2(S, 2)
  switch (shift) {
    case 0: KX_STAR(0) break;
    case 1: KX_STAR(1) break;
    case 2: KX_STAR(2) break;
    default: KX_STAR(3) break;
  }
#undef KX_STAR
#undef KX_STAR2
#undef KX_STAR3
  return after_launch("stencil3d_star_kernel");
}

}  // namespace

// Called first by try_stencil3d; declines (KX_ERR_UNSUPPORTED) anything that is not a
// host-coefficient 3x3x3 star with 16-byte aligned rows.
int try_stencil3d_star(const MapArgs& m, const KxProgram& p, int64_t batch, cudaStream_t s) {
  const KxGeom& g = m.g;
  const int nd = g.ndim;
  if (g.ksize[nd - 3] != 3 || g.ksize[nd - 2] != 3 || g.ksize[nd - 1] != 3) return KX_ERR_UNSUPPORTED;
  if (m.coef_dev) return KX_ERR_UNSUPPORTED;
  if (g.in_shape[nd - 1] % 4 != 0 || (reinterpret_cast<uintptr_t>(m.in) & 15) != 0) return KX_ERR_UNSUPPORTED;
  if (((int64_t)g.in_shape[nd - 3] * g.in_shape[nd - 2] * g.in_shape[nd - 1]) % 4 != 0 && batch > 1)
    return KX_ERR_UNSUPPORTED;
  if (g.in_dtype == KX_F32) return run_star<float>(m, p, batch, s);
  if (g.in_dtype == KX_I32 && p.func[0].coef_is_int) return run_star<int32_t>(m, p, batch, s);
  return KX_ERR_UNSUPPORTED;
}

}  // namespace kx
