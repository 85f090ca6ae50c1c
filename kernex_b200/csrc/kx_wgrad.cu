// Weight gradient of a stride-1 AFFINE window of ANY size over the last two or three axes (conv-as-kmap `sum(x * w)` with
// 5x5 / 7x7 / 3x3x3 ... windows): dW[kz][ky][kx] = sum over windows j of g[j] * x[origin(j) + (kz, ky, kx)] -- the
// correlation of the cotangent with the input (north star: "the weight gradient is a fused window-reduction kernel").
// kx_conv2d.cu has the TMA row-ring form for 3x3 and (C, 3, 3) windows; every other shape ran on generic_vjp_coef
// (one pass over all windows PER TAP with 64-bit index arithmetic: 5x5 on 4096^2 2.2 ms, 3x3x3 on 256^3 3.2 ms).
//
// The transposed row march of kx_stencil2d.cu: a thread owns one column of a band of output rows and keeps the last KY
// input rows of its KX window columns in a register queue; per output row it loads one cotangent value and KX new input
// values (coalesced across the warp, neighbours' loads hit L1) and updates KY * KX accumulators, acc[ky][kx] += g * x.
// (Rotating the queue by renaming -- the row loop unrolled by KY -- was measured: 96 instead of 72 registers at 5x5 and
// 0.51 instead of 0.44 ms on 8192^2; the moves stay.)
// For 3-D windows blockIdx.z enumerates (window plane kz, chunk of output planes): the same 2-D correlation between
// cotangent plane z and input plane z + kz - lz, summed over the chunk.  Reduction: registers -> warp shuffle tree ->
// shared memory -> one partial vector per block; a second kernel adds the partials in a fixed order (deterministic,
// no atomics) and scatters them into the program's tap order (applying the forward's final division).
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kWT = 128;          // threads per block = columns per block
constexpr int kWRows = 64;        // output rows per band
constexpr int kWMaxK = 7;

struct WGArgs {
  const float* x;
  const float* g;
  float* partial;                 // [blocks][KY * KX] (block = linear (bx, by, bz))
  int64_t x_batch, g_batch;       // elements per batch item
  int din, hin, win, dout, hout, wout;
  int lz, ly, lx;
  int KZ, zc, nzc;                // window planes, output planes per chunk, chunks
  int batch;
};

template <int KY, int KX>
__global__ void __launch_bounds__(kWT)
wgrad_rows_kernel(const __grid_constant__ WGArgs a) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x * kWT + tid;                      // output column
  const int r0 = blockIdx.y * kWRows, r1 = min(r0 + kWRows, a.hout);
  int bz = blockIdx.z;
  const int chunk = bz % a.nzc;
  bz /= a.nzc;
  const int kz = bz % a.KZ;
  const int b = bz / a.KZ;
  const float* __restrict__ xb = a.x + (int64_t)b * a.x_batch;
  const float* __restrict__ gb = a.g + (int64_t)b * a.g_batch;
  float acc[KY][KX];
#pragma unroll
  for (int ky = 0; ky < KY; ++ky)
#pragma unroll
    for (int kx = 0; kx < KX; ++kx) acc[ky][kx] = 0.f;
  const bool col_ok = c < a.wout;
  const int cx0 = c - a.lx;                                  // input column of window column 0
  const int z0 = chunk * a.zc, z1 = min(z0 + a.zc, a.dout);
  for (int z = z0; z < z1; ++z) {
    const int iz = z + kz - a.lz;
    if (iz < 0 || iz >= a.din) continue;                     // zero-padded plane: contributes nothing
    const float* __restrict__ xp = xb + (int64_t)iz * a.hin * a.win;
    const float* __restrict__ gp = gb + (int64_t)z * a.hout * a.wout;
    float q[KY][KX];                                         // input rows r - ly .. r - ly + KY - 1 at the window columns
    auto load_row = [&](int iy, float(&dst)[KX]) {
      const bool row_ok = col_ok && iy >= 0 && iy < a.hin;
      const float* rp = xp + (int64_t)iy * a.win;
#pragma unroll
      for (int kx = 0; kx < KX; ++kx) {
        const int cx = cx0 + kx;
        dst[kx] = (row_ok && cx >= 0 && cx < a.win) ? __ldg(rp + cx) : 0.f;
      }
    };
#pragma unroll
    for (int ky = 0; ky < KY - 1; ++ky) load_row(r0 - a.ly + ky, q[ky]);
    for (int r = r0; r < r1; ++r) {
      load_row(r - a.ly + KY - 1, q[KY - 1]);
      const float gv = col_ok ? __ldg(gp + (int64_t)r * a.wout + c) : 0.f;
#pragma unroll
      for (int ky = 0; ky < KY; ++ky)
#pragma unroll
        for (int kx = 0; kx < KX; ++kx) acc[ky][kx] = fmaf(gv, q[ky][kx], acc[ky][kx]);
#pragma unroll
      for (int ky = 0; ky < KY - 1; ++ky)
#pragma unroll
        for (int kx = 0; kx < KX; ++kx) q[ky][kx] = q[ky + 1][kx];
    }
  }
  // block reduction: shuffle tree per accumulator, then the 4 warps through shared memory in warp order
  __shared__ float red[kWT / 32][KY * KX];
#pragma unroll
  for (int ky = 0; ky < KY; ++ky)
#pragma unroll
    for (int kx = 0; kx < KX; ++kx) {
      float v = acc[ky][kx];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][ky * KX + kx] = v;
    }
  __syncthreads();
  if (tid < KY * KX) {
    float v = red[0][tid];
#pragma unroll
    for (int w = 1; w < kWT / 32; ++w) v += red[w][tid];
    const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    a.partial[blk * (KY * KX) + tid] = v;
  }
}

struct TapMap {
  int t[16 * kWMaxK * kWMaxK];      // dense (kz, ky, kx) -> tap number of the program, -1 = absent
};

// dcoef[tap] = sum over the blocks of (batch, kz) of partial[block][ky * KX + kx], blocks in index order
__global__ void __launch_bounds__(256)
wgrad_stage2_kernel(const float* __restrict__ partial, float* __restrict__ dcoef, const __grid_constant__ TapMap tap_of, int n_dense,
                    int kykx, int KZ, int nzc, int batch, int64_t blocks_xy, float post_div) {
  // one warp per dense coefficient (kz, ky, kx)
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (d >= n_dense) return;
  const int kz = d / kykx, t2 = d - kz * kykx;
  float v = 0.f;
  for (int b = 0; b < batch; ++b)
    for (int ch = 0; ch < nzc; ++ch) {
      const int64_t bz = ((int64_t)b * KZ + kz) * nzc + ch;
      const float* p = partial + bz * blocks_xy * kykx + t2;
      float s = 0.f;
      for (int64_t i = lane; i < blocks_xy; i += 32) s += p[i * kykx];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      v += s;
    }
  if (lane == 0) {
    const int t = tap_of.t[d];
    if (t >= 0) dcoef[t] = post_div != 0.f ? v / post_div : v;
  }
}

struct WGPlan {
  bool ok;
  int nd3;                 // 2 or 3 window axes
  int KZ, KY, KX;
  int zc, nzc;
  int64_t batch;
  dim3 grid;
  int64_t blocks_xy;
};

WGPlan plan_wgrad(const KxGeom& g, const KxProgram& p) {
  WGPlan w{};                  // ok = false
  const int nd = g.ndim;
  if (nd < 2 || p.n_funcs != 1 || p.func[0].kind != KX_AFFINE || g.in_dtype != KX_F32 || g.out_dtype != KX_F32) return w;
  if (getenv("KX_NO_WGRAD_ROWS")) return w;
  // window axes: the last two, plus the third-last when its window is wider than 1
  const int nw = (nd >= 3 && g.ksize[nd - 3] > 1) ? 3 : 2;
  int64_t batch = 1;
  for (int d = 0; d < nd - nw; ++d) {
    if (g.ksize[d] != 1 || g.stride[d] != 1 || g.lo[d] != 0 || g.hi[d] != 0) return w;
    batch *= g.in_shape[d];
  }
  for (int d = nd - nw; d < nd; ++d)
    if (g.stride[d] != 1 || g.in_shape[d] > 0x3fffffff || g.out_shape[d] > 0x3fffffff || g.out_shape[d] < 1) return w;
  w.KZ = nw == 3 ? g.ksize[nd - 3] : 1;
  w.KY = g.ksize[nd - 2];
  w.KX = g.ksize[nd - 1];
  if (w.KY > kWMaxK || w.KX > kWMaxK || w.KZ > 16 || w.KY != w.KX || (w.KY != 3 && w.KY != 5 && w.KY != 7 && w.KY != 2 && w.KY != 4))
    return w;
  const int64_t dout = nw == 3 ? g.out_shape[nd - 3] : 1;
  const int64_t gx = (g.out_shape[nd - 1] + kWT - 1) / kWT, gy = (g.out_shape[nd - 2] + kWRows - 1) / kWRows;
  // output planes per chunk: enough blocks for a few waves, few enough partials for the second stage
  int64_t nzc = 1;
  const int64_t want = (int64_t)sm_count() * 8;
  while (nzc < dout && gx * gy * batch * w.KZ * nzc < want) nzc *= 2;
  if (nzc > dout) nzc = dout;
  w.zc = (int)((dout + nzc - 1) / nzc);
  w.nzc = (int)((dout + w.zc - 1) / w.zc);
  const int64_t gz = batch * w.KZ * w.nzc;
  if (gz > 65535 || gy > 65535 || batch > 4096) return w;
  w.batch = batch;
  w.nd3 = nw;
  w.grid = dim3((unsigned)gx, (unsigned)gy, (unsigned)gz);
  w.blocks_xy = gx * gy;
  w.ok = true;
  return w;
}

}  // namespace

int64_t wgrad_rows_scratch_bytes(const KxGeom& g, const KxProgram& p) {
  const WGPlan w = plan_wgrad(g, p);
  if (!w.ok) return 0;
  const int64_t n_dense = (int64_t)w.KZ * w.KY * w.KX;
  (void)n_dense;
  return (int64_t)w.grid.x * w.grid.y * w.grid.z * w.KY * w.KX * 4 + 256;
}

int try_wgrad_rows(const KxGeom& g, const KxProgram& p, const void* x, const void* gout, void* scratch, void* dcoef,
                   cudaStream_t s) {
  const WGPlan w = plan_wgrad(g, p);
  if (!w.ok) return KX_ERR_UNSUPPORTED;
  const int nd = g.ndim, nw = w.nd3;
  const KxFunc& f = p.func[0];
  const int n_dense = w.KZ * w.KY * w.KX;
  // tap table: dense (kz, ky, kx) -> tap number of the program (absent taps: -1); duplicate taps -> generic kernel
  static thread_local TapMap tap_of;
  for (int i = 0; i < 16 * kWMaxK * kWMaxK; ++i) tap_of.t[i] = -1;
  for (int t = f.tap_begin; t < f.tap_end; ++t) {
    for (int d = 0; d < nd - nw; ++d)
      if (p.tap_off[t][d] != 0) return KX_ERR_UNSUPPORTED;
    const int kz = nw == 3 ? p.tap_off[t][nd - 3] : 0;
    const int idx = (kz * w.KY + p.tap_off[t][nd - 2]) * w.KX + p.tap_off[t][nd - 1];
    if (tap_of.t[idx] >= 0) return KX_ERR_UNSUPPORTED;
    tap_of.t[idx] = t - f.tap_begin;
  }
  WGArgs a;
  memset(&a, 0, sizeof(a));
  a.x = (const float*)x;
  a.g = (const float*)gout;
  a.partial = (float*)scratch;
  a.din = nw == 3 ? (int)g.in_shape[nd - 3] : 1;
  a.hin = (int)g.in_shape[nd - 2];
  a.win = (int)g.in_shape[nd - 1];
  a.dout = nw == 3 ? (int)g.out_shape[nd - 3] : 1;
  a.hout = (int)g.out_shape[nd - 2];
  a.wout = (int)g.out_shape[nd - 1];
  a.lz = nw == 3 ? g.lo[nd - 3] : 0;
  a.ly = g.lo[nd - 2];
  a.lx = g.lo[nd - 1];
  a.x_batch = (int64_t)a.din * a.hin * a.win;
  a.g_batch = (int64_t)a.dout * a.hout * a.wout;
  a.KZ = w.KZ;
  a.zc = w.zc;
  a.nzc = w.nzc;
  a.batch = (int)w.batch;
#define KX_WG(K) wgrad_rows_kernel<K, K><<<w.grid, kWT, 0, s>>>(a)
  switch (w.KY) {
    case 2: KX_WG(2); break;
    case 3: KX_WG(3); break;
    case 4: KX_WG(4); break;
    case 5: KX_WG(5); break;
    default: KX_WG(7); break;
  }
#undef KX_WG
  int rc = after_launch("wgrad_rows_kernel");
  if (rc) return rc;
  wgrad_stage2_kernel<<<(n_dense + 7) / 8, 256, 0, s>>>((const float*)scratch, (float*)dcoef, tap_of, n_dense, w.KY * w.KX, w.KZ,
                                                        w.nzc, (int)w.batch, w.blocks_xy, (float)f.post_div);
  return after_launch("wgrad_stage2_kernel");
}

}  // namespace kx
