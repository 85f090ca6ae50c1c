// Fast path for kscan: 2-D time-marching stencils (README.md:231-266 linear convection,
// BASELINE config 4).
//
// The reference runs kscan as ONE lax.scan over all nt*nx windows carrying the padded
// array (kernex/_src/scan.py:97-113).  When every window function reads only the row
// ABOVE the cell it writes (centre-relative row offset -1) the sweep order inside a row
// is irrelevant: row n is a pure function of row n-1, which was fully produced earlier
// in the sweep, and -- with 'same'-style borders on both axes -- the input array is never
// read at all (row -1 and the side columns are the zero pad, every other cell is
// overwritten before it is read).  The kernel is then WRITE-only: 4 B per lattice update.
//
// Temporal blocking: one launch advances kT rows.  A block owns a contiguous span of
// columns plus a redundant halo of kT*R columns per side that shrinks by R per row (the
// dependency cone), so blocks never talk to each other inside a launch; the next launch
// re-reads only the last row.  Inside a block each thread keeps its 4*G columns of the
// current row in registers, rows are exchanged through a double-buffered shared-memory
// line (one barrier per row), and each row is written with 128-bit streaming stores.
//
// F[slice] = fn regions (kernex/_src/utils.py:338-382 selection rule) are resolved per
// thread into per-column coefficient registers (bias, a[-2..2]) whenever the marching row
// crosses a row breakpoint of the region keys, so the inner loop is branch-free FMAs and
// constants (boundary / initial conditions) are just "all a = 0".
#include "kx_internal.cuh"

namespace kx {

namespace {

constexpr int kRMThreads = 256;
constexpr int kT = 64;          // rows per launch
constexpr int kMaxR = 2;        // |column offset| supported
constexpr int kMaxRowIv = 16;

struct RMFunc {
  float bias;
  float a[2 * kMaxR + 1];       // coefficient of prev[i + d], d = -kMaxR..kMaxR
};

struct RMArgs {
  float* out;
  int64_t nx;                   // local width (columns held by this call)
  int64_t col_base, gnx;        // global column of local column 0, global width (axis-1 sharding)
  int nt;
  int n0, n1;                   // rows advanced by this launch
  int own;                      // owned columns per block (multiple of 4)
  int ntiles;                   // fat kernel: column tiles (a block loops over tile, tile+grid, ...)
  int halo_l, halo_r;           // redundant columns per side (multiples of 4)
  int n_funcs, n_keys;
  int n_breaks;
  int row_break[kMaxRowIv + 1]; // sorted row breakpoints, row_break[0] = 0
  RMFunc func[KX_MAX_FUNCS];
  KxKey key[KX_MAX_KEYS];
};

__device__ __forceinline__ int rm_select(const RMArgs& a, int64_t row, int64_t col) {
  if (a.n_funcs <= 1) return 0;
  for (int k = a.n_keys - 1; k >= 0; --k) {
    const KxKey& key = a.key[k];
    bool ok = true;
    for (int d = 0; d < key.n_items; ++d) {
      const KxKeyItem& it = key.item[d];
      const int64_t c = d == 0 ? row : col;
      bool m = true;
      if (it.kind == KX_KEY_EQ) m = (c == it.a);
      else if (it.kind == KX_KEY_RANGE) m = (it.a <= c) && (c < it.b);
      else if (it.kind == KX_KEY_RANGE_STEP) m = (it.a <= c) && (c < it.b) && (c % it.step == 0);
      ok = ok && m;
    }
    if (ok) return key.func;
  }
  return 0;
}

template <int G, int RL, int RR>
__global__ void __launch_bounds__(kRMThreads)
scan_rowmarch_kernel(const __grid_constant__ RMArgs a) {
  constexpr int kEdge = 2 * (G * (kRMThreads / 32) + 2);   // 2 floats per warp slot (+ a zero slot per side)
  __shared__ float edge[2][2 * kEdge];                     // [parity][right edges | left edges]
  const int tid = threadIdx.x;
  const int64_t span0 = (int64_t)blockIdx.x * a.own - a.halo_l;  // first processed column
  const int64_t own0 = (int64_t)blockIdx.x * a.own;
  const int64_t own1 = min(own0 + a.own, a.nx);

  float cur[G][4];
  float cb[G][4], cl2[G][4], cl1[G][4], cc[G][4], cr1[G][4], cr2[G][4];
  (void)cl2; (void)cr2; (void)cl1; (void)cr1;

  // previous row: zero pad for n0 == 0, else the last row of the previous launch
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int64_t c0 = span0 + (int64_t)g * kRMThreads * 4 + tid * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) cur[g][e] = 0.f;
    if (a.n0 > 0) {
      const float* prow = a.out + (int64_t)(a.n0 - 1) * a.nx;
      if (c0 >= 0 && c0 + 3 < a.nx && ((a.nx & 3) == 0)) {
        const float4 v = *reinterpret_cast<const float4*>(prow + c0);
        cur[g][0] = v.x; cur[g][1] = v.y; cur[g][2] = v.z; cur[g][3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c0 + e >= 0 && c0 + e < a.nx) cur[g][e] = prow[c0 + e];
      }
    }
  }
  for (int i = tid; i < 4 * kEdge; i += kRMThreads) (&edge[0][0])[i] = 0.f;   // incl. the zero slots
  __syncthreads();

  int iv = 0;
  while (iv + 1 < a.n_breaks && a.row_break[iv + 1] <= a.n0) ++iv;
  int next_break = a.n0;   // force a coefficient refresh at the first row

  // loop invariants the compiler otherwise re-derives every row (ncu/SASS of the first version: three
  // S2R, ~10 instructions of 64-bit address arithmetic and 5 local-memory loads of spilled
  // coefficients per row in an issue-bound loop)
  int lane = tid & 31;
  const int wrp = tid >> 5;
  asm volatile("" : "+r"(lane));   // pin it: the compiler otherwise re-reads SR_TID.X every row
  int store_mode[G];        // 2 = aligned 128-bit store, 1 = element-wise (ragged right edge), 0 = halo column: no store
  float* optr[G];           // &out[n][c0] of the row being written
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int64_t c0 = span0 + (int64_t)g * kRMThreads * 4 + tid * 4;
    store_mode[g] = (c0 >= own0 && c0 < own1) ? ((c0 + 3 < own1 && ((a.nx & 3) == 0)) ? 2 : 1) : 0;
    optr[g] = a.out + (int64_t)a.n0 * a.nx + c0;
  }
  unsigned e_wr[G], e_rl[G], e_wl[G], e_rr[G];   // shared-memory byte addresses of this lane's edge slots (parity 0)
  {
    const unsigned e0 = (unsigned)__cvta_generic_to_shared(&edge[0][0]);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int slot = 1 + g * (kRMThreads / 32) + wrp;      // slot 0 and the last slot stay zero
      e_wr[g] = e0 + (unsigned)(2 * slot) * 4;               // my right edge -> read by the warp on my right
      e_rl[g] = e0 + (unsigned)(2 * (slot - 1)) * 4;         // right edge of the warp on my left
      e_wl[g] = e0 + (unsigned)(kEdge + 2 * slot) * 4;
      e_rr[g] = e0 + (unsigned)(kEdge + 2 * (slot + 1)) * 4;
      asm volatile("" : "+r"(e_wr[g]), "+r"(e_rl[g]), "+r"(e_wl[g]), "+r"(e_rr[g]));   // keep them in registers
    }
  }
  constexpr unsigned kParityBytes = (unsigned)(2 * kEdge * 4);
  unsigned par = (a.n0 & 1) ? kParityBytes : 0u;

  int n = a.n0;
  while (n < a.n1) {
    {   // row n starts a new interval of the region keys: refresh the per-column coefficients
      while (iv + 1 < a.n_breaks && a.row_break[iv + 1] <= n) ++iv;
      next_break = (iv + 1 < a.n_breaks) ? a.row_break[iv + 1] : 0x7fffffff;
#pragma unroll
      for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t c = span0 + (int64_t)g * kRMThreads * 4 + tid * 4 + e;
          const int64_t gc = a.col_base + c;
          const bool held = !(c < 0 || c >= a.nx || gc < 0 || gc >= a.gnx);   // else: pad / not-held column, reads as zero
          const RMFunc& f = a.func[held ? rm_select(a, n, gc) : 0];           // read straight from the parameter bank
          cb[g][e] = held ? f.bias : 0.f;
          cl2[g][e] = held ? f.a[0] : 0.f;
          cl1[g][e] = held ? f.a[1] : 0.f;
          cc[g][e] = held ? f.a[2] : 0.f;
          cr1[g][e] = held ? f.a[3] : 0.f;
          cr2[g][e] = held ? f.a[4] : 0.f;
        }
      }
    }
    const int stop = min(next_break, a.n1);
#pragma unroll 2
    for (; n < stop; ++n) {   // branch-free steady state between two breakpoints
    // neighbours across thread boundaries: warp shuffles inside a warp; only the two edge
    // lanes of every warp go through shared memory (one barrier per row, double buffered)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const bool lane_first = lane == 0, lane_last = lane == 31;
      if (RL >= 1 && lane_last) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(e_wr[g] + par), "f"(cur[g][3]) : "memory");
        if (RL >= 2) asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(e_wr[g] + par), "f"(cur[g][2]) : "memory");
      }
      if (RR >= 1 && lane_first) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(e_wl[g] + par), "f"(cur[g][0]) : "memory");
        if (RR >= 2) asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(e_wl[g] + par), "f"(cur[g][1]) : "memory");
      }
    }
    __syncthreads();
    float nxt[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const bool lane_first = lane == 0, lane_last = lane == 31;
      float w[4 + 2 * kMaxR];  // prev[base-2 .. base+5]
      w[2] = cur[g][0]; w[3] = cur[g][1]; w[4] = cur[g][2]; w[5] = cur[g][3];
      w[0] = w[1] = w[6] = w[7] = 0.f;
      if (RL >= 1) {
        w[1] = __shfl_up_sync(0xffffffffu, cur[g][3], 1);
        if (lane_first) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[1]) : "r"(e_rl[g] + par) : "memory");
      }
      if (RL >= 2) {
        w[0] = __shfl_up_sync(0xffffffffu, cur[g][2], 1);
        if (lane_first) asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(w[0]) : "r"(e_rl[g] + par) : "memory");
      }
      if (RR >= 1) {
        w[6] = __shfl_down_sync(0xffffffffu, cur[g][0], 1);
        if (lane_last) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[6]) : "r"(e_rr[g] + par) : "memory");
      }
      if (RR >= 2) {
        w[7] = __shfl_down_sync(0xffffffffu, cur[g][1], 1);
        if (lane_last) asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(w[7]) : "r"(e_rr[g] + par) : "memory");
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = cb[g][e];
        if (RL >= 2) v = fmaf(cl2[g][e], w[e + 0], v);
        if (RL >= 1) v = fmaf(cl1[g][e], w[e + 1], v);
        v = fmaf(cc[g][e], w[e + 2], v);
        if (RR >= 1) v = fmaf(cr1[g][e], w[e + 3], v);
        if (RR >= 2) v = fmaf(cr2[g][e], w[e + 4], v);
        nxt[g][e] = v;
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int e = 0; e < 4; ++e) cur[g][e] = nxt[g][e];
      if (store_mode[g] == 2) {
        __stcs(reinterpret_cast<float4*>(optr[g]), make_float4(nxt[g][0], nxt[g][1], nxt[g][2], nxt[g][3]));
      } else if (store_mode[g] == 1) {
        const int64_t c0 = span0 + (int64_t)g * kRMThreads * 4 + tid * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c0 + e < own1) optr[g][e] = nxt[g][e];
      }
      optr[g] += a.nx;
    }
    par ^= kParityBytes;
    }
  }
}

template <int G>
int launch_g(const RMArgs& a, int rl, int rr, unsigned blocks, cudaStream_t s) {
  if (rl <= 1 && rr == 0) scan_rowmarch_kernel<G, 1, 0><<<blocks, kRMThreads, 0, s>>>(a);
  else if (rl <= 1 && rr <= 1) scan_rowmarch_kernel<G, 1, 1><<<blocks, kRMThreads, 0, s>>>(a);
  else scan_rowmarch_kernel<G, 2, 2><<<blocks, kRMThreads, 0, s>>>(a);
  return after_launch("scan_rowmarch_kernel");
}

// ---- "fat" variant: one 512-thread block per SM owning ~16K columns -----------------------
// benchmarks/expwrite.cu: the same 2-D write pattern runs at 6.0 TB/s with 1024 blocks of 2048
// columns but at 7.1 TB/s with <= 128 blocks of 16384 columns (one long write stream per SM, 16
// blocks per GPC), so wide grids use this kernel.  32 columns per thread would need 3 coefficient
// registers per column; instead a group of 4 columns shares ONE coefficient set (regions are
// rectangles, so almost every group is uniform) and the few groups a region boundary cuts
// through fetch per-column sets from a small shared-memory exception table.
constexpr int kFatThreads = 512, kFatG = 8, kFatSpan = kFatThreads * 4 * kFatG;   // 16384 columns
constexpr int kMaxExc = 96;

template <int RL, int RR>
__global__ void __launch_bounds__(kFatThreads, 1)
scan_rowmarch_fat_kernel(const __grid_constant__ RMArgs a) {
  constexpr int G = kFatG, NW = kFatThreads / 32;
  constexpr int kEdge = 2 * (G * NW + 2);
  __shared__ float edge[2][2 * kEdge];
  __shared__ float exc[kMaxExc][4][6];      // [slot][column][bias, a(-2..2)]
  __shared__ int exc_count;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
  const int64_t span0 = (int64_t)tile * a.own - a.halo_l;
  const int64_t own0 = (int64_t)tile * a.own;
  const int64_t own1 = min(own0 + a.own, a.nx);

  float cur[G][4];
  float ub[G], ul2[G], ul1[G], uc[G], ur1[G], ur2[G];
  int xs[G];
  (void)ul2; (void)ur2; (void)ul1; (void)ur1;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int64_t c0 = span0 + (int64_t)g * kFatThreads * 4 + tid * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) cur[g][e] = 0.f;
    xs[g] = -1;
    if (a.n0 > 0) {
      const float* prow = a.out + (int64_t)(a.n0 - 1) * a.nx;
      if (c0 >= 0 && c0 + 3 < a.nx && ((a.nx & 3) == 0)) {
        const float4 v = *reinterpret_cast<const float4*>(prow + c0);
        cur[g][0] = v.x; cur[g][1] = v.y; cur[g][2] = v.z; cur[g][3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c0 + e >= 0 && c0 + e < a.nx) cur[g][e] = prow[c0 + e];
      }
    }
  }
  for (int i = tid; i < 4 * kEdge; i += kFatThreads) (&edge[0][0])[i] = 0.f;
  __syncthreads();

  int iv = 0;
  while (iv + 1 < a.n_breaks && a.row_break[iv + 1] <= a.n0) ++iv;
  int next_break = a.n0;

  for (int n = a.n0; n < a.n1; ++n) {
    if (n == next_break) {   // block-uniform: resolve the region functions of this row interval
      while (iv + 1 < a.n_breaks && a.row_break[iv + 1] <= n) ++iv;
      next_break = (iv + 1 < a.n_breaks) ? a.row_break[iv + 1] : 0x7fffffff;
      if (tid == 0) exc_count = 0;
      __syncthreads();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        int fi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t c = span0 + (int64_t)g * kFatThreads * 4 + tid * 4 + e;
          const int64_t gc = a.col_base + c;
          fi[e] = (c < 0 || c >= a.nx || gc < 0 || gc >= a.gnx) ? -1 : rm_select(a, n, gc);
        }
        const bool uniform = fi[0] == fi[1] && fi[1] == fi[2] && fi[2] == fi[3];
        RMFunc f;
        f.bias = 0.f;
#pragma unroll
        for (int d = 0; d < 2 * kMaxR + 1; ++d) f.a[d] = 0.f;
        if (uniform) {
          if (fi[0] >= 0) f = a.func[fi[0]];
          xs[g] = -1;
        } else {
          const int slot = atomicAdd(&exc_count, 1);
          xs[g] = slot < kMaxExc ? slot : kMaxExc - 1;   // the host guarantees it fits
          if (slot < kMaxExc) {
            for (int e = 0; e < 4; ++e) {
              RMFunc fe;
              fe.bias = 0.f;
              for (int d = 0; d < 2 * kMaxR + 1; ++d) fe.a[d] = 0.f;
              if (fi[e] >= 0) fe = a.func[fi[e]];
              exc[slot][e][0] = fe.bias;
              for (int d = 0; d < 2 * kMaxR + 1; ++d) exc[slot][e][1 + d] = fe.a[d];
            }
          }
        }
        ub[g] = f.bias;
        ul2[g] = f.a[0]; ul1[g] = f.a[1]; uc[g] = f.a[2]; ur1[g] = f.a[3]; ur2[g] = f.a[4];
      }
      __syncthreads();
    }
    float* eb = edge[n & 1];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int slot = 1 + g * NW + wrp;
      if (RL >= 1 && lane == 31) {
        eb[2 * slot] = cur[g][3];
        if (RL >= 2) eb[2 * slot + 1] = cur[g][2];
      }
      if (RR >= 1 && lane == 0) {
        eb[kEdge + 2 * slot] = cur[g][0];
        if (RR >= 2) eb[kEdge + 2 * slot + 1] = cur[g][1];
      }
    }
    __syncthreads();
    float* orow = a.out + (int64_t)n * a.nx;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int slot = 1 + g * NW + wrp;
      float w[4 + 2 * kMaxR];
      w[2] = cur[g][0]; w[3] = cur[g][1]; w[4] = cur[g][2]; w[5] = cur[g][3];
      w[0] = w[1] = w[6] = w[7] = 0.f;
      if (RL >= 1) {
        w[1] = __shfl_up_sync(0xffffffffu, cur[g][3], 1);
        if (lane == 0) w[1] = eb[2 * (slot - 1)];
      }
      if (RL >= 2) {
        w[0] = __shfl_up_sync(0xffffffffu, cur[g][2], 1);
        if (lane == 0) w[0] = eb[2 * (slot - 1) + 1];
      }
      if (RR >= 1) {
        w[6] = __shfl_down_sync(0xffffffffu, cur[g][0], 1);
        if (lane == 31) w[6] = eb[kEdge + 2 * (slot + 1)];
      }
      if (RR >= 2) {
        w[7] = __shfl_down_sync(0xffffffffu, cur[g][1], 1);
        if (lane == 31) w[7] = eb[kEdge + 2 * (slot + 1) + 1];
      }
      float v[4];
      if (xs[g] < 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float t = ub[g];
          if (RL >= 2) t = fmaf(ul2[g], w[e + 0], t);
          if (RL >= 1) t = fmaf(ul1[g], w[e + 1], t);
          t = fmaf(uc[g], w[e + 2], t);
          if (RR >= 1) t = fmaf(ur1[g], w[e + 3], t);
          if (RR >= 2) t = fmaf(ur2[g], w[e + 4], t);
          v[e] = t;
        }
      } else {
        const float(*x)[6] = exc[xs[g]];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float t = x[e][0];
          if (RL >= 2) t = fmaf(x[e][1], w[e + 0], t);
          if (RL >= 1) t = fmaf(x[e][2], w[e + 1], t);
          t = fmaf(x[e][3], w[e + 2], t);
          if (RR >= 1) t = fmaf(x[e][4], w[e + 3], t);
          if (RR >= 2) t = fmaf(x[e][5], w[e + 4], t);
          v[e] = t;
        }
      }
      const int64_t c0 = span0 + (int64_t)g * kFatThreads * 4 + tid * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) cur[g][e] = v[e];
      if (c0 >= own0 && c0 < own1) {
        if (c0 + 3 < own1 && ((a.nx & 3) == 0)) {
          __stcs(reinterpret_cast<float4*>(orow + c0), make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c0 + e < own1) orow[c0 + e] = v[e];
        }
      }
    }
  }
  __syncthreads();   // the shared tables are reused by the next tile
  }
}

int launch_fat(const RMArgs& a, int rl, int rr, unsigned blocks, cudaStream_t s) {
  if (rl <= 1 && rr == 0) scan_rowmarch_fat_kernel<1, 0><<<blocks, kFatThreads, 0, s>>>(a);
  else if (rl <= 1 && rr <= 1) scan_rowmarch_fat_kernel<1, 1><<<blocks, kFatThreads, 0, s>>>(a);
  else scan_rowmarch_fat_kernel<2, 2><<<blocks, kFatThreads, 0, s>>>(a);
  return after_launch("scan_rowmarch_kernel");
}

bool add_break(int* br, int& n, int64_t v, int nt) {
  if (v <= 0 || v >= nt) return true;
  for (int i = 0; i < n; ++i)
    if (br[i] == (int)v) return true;
  if (n >= kMaxRowIv) return false;
  br[n++] = (int)v;
  return true;
}

}  // namespace

// `col_base` / `gnx`: the call produces columns [col_base, col_base + g.in_shape[1]) of a scan whose
// global width is gnx (region keys and the zero pad are in GLOBAL coordinates).  Columns the call
// does not hold read as zero, so the first nt*RL (last nt*RR) local columns of a shard that does not
// start (end) at the global edge are only valid if the caller overlaps shards by that much.
int scan_rowmarch_shard(const KxGeom& g, const KxProgram& p, const void* coef_dev, void* out, int reverse,
                        int64_t col_base, int64_t gnx, cudaStream_t s, int row_begin, int row_end) {
  if (g.ndim != 2 || g.in_dtype != KX_F32 || reverse || coef_dev) return KX_ERR_UNSUPPORTED;
  const int c0 = (g.ksize[0] - 1) / 2, c1 = (g.ksize[1] - 1) / 2;
  // every in-bounds cell is a window centre exactly once ('same'-style borders), stride 1
  for (int d = 0; d < 2; ++d) {
    const int c = (g.ksize[d] - 1) / 2;
    if (g.stride[d] != 1 || g.lo[d] != c || g.out_shape[d] != g.in_shape[d]) return KX_ERR_UNSUPPORTED;
  }
  if (g.in_shape[0] > 0x7ffffff0 || g.in_shape[0] < 1 || g.in_shape[1] < 1) return KX_ERR_UNSUPPORTED;
  static thread_local RMArgs a;
  memset(&a, 0, sizeof(a));
  int rl = 0, rr = 0;
  for (int f = 0; f < p.n_funcs; ++f) {
    const KxFunc& fn = p.func[f];
    if (fn.kind != KX_AFFINE || fn.post_div != 0.0) return KX_ERR_UNSUPPORTED;
    a.func[f].bias = (float)fn.bias;
    for (int t = fn.tap_begin; t < fn.tap_end; ++t) {
      const int d0 = p.tap_off[t][0] - c0, d1 = p.tap_off[t][1] - c1;
      if (d0 != -1 || d1 < -kMaxR || d1 > kMaxR) return KX_ERR_UNSUPPORTED;  // must read the row above only
      a.func[f].a[d1 + kMaxR] += (float)p.coef[t];
      if (-d1 > rl) rl = -d1;
      if (d1 > rr) rr = d1;
    }
  }
  const int nt = (int)g.in_shape[0];
  if (row_end < 0 || row_end > nt) row_end = nt;
  if (row_begin < 0 || row_begin > row_end) return KX_ERR_UNSUPPORTED;
  a.n_breaks = 1;
  a.row_break[0] = 0;
  for (int k = 0; k < p.n_keys; ++k) {
    a.key[k] = p.key[k];
    if (p.key[k].n_items < 1) continue;
    const KxKeyItem& it = p.key[k].item[0];
    bool ok = true;
    if (it.kind == KX_KEY_EQ) ok = add_break(a.row_break, a.n_breaks, it.a, nt) && add_break(a.row_break, a.n_breaks, it.a + 1, nt);
    else if (it.kind == KX_KEY_RANGE || (it.kind == KX_KEY_RANGE_STEP && (it.step == 1 || it.step == -1)))
      ok = add_break(a.row_break, a.n_breaks, it.a, nt) && add_break(a.row_break, a.n_breaks, it.b, nt);
    else if (it.kind == KX_KEY_RANGE_STEP) return KX_ERR_UNSUPPORTED;  // strided row keys: generic path
    if (!ok) return KX_ERR_UNSUPPORTED;
  }
  for (int i = 1; i < a.n_breaks; ++i)  // insertion sort
    for (int j = i; j > 0 && a.row_break[j] < a.row_break[j - 1]; --j) {
      const int t = a.row_break[j];
      a.row_break[j] = a.row_break[j - 1];
      a.row_break[j - 1] = t;
    }
  a.n_funcs = p.n_funcs;
  a.n_keys = p.n_keys;
  a.out = (float*)out;
  a.nx = g.in_shape[1];
  a.col_base = col_base;
  a.gnx = gnx;
  a.nt = nt;
  a.halo_l = ((kT * rl + 3) / 4) * 4;
  a.halo_r = ((kT * rr + 3) / 4) * 4;
  const int64_t nx = a.nx;
  // wide grids: one fat block per SM (see scan_rowmarch_fat_kernel); needs rectangular column keys
  {
    bool col_steps = false;
    for (int k = 0; k < p.n_keys; ++k)
      if (p.key[k].n_items >= 2 && p.key[k].item[1].kind == KX_KEY_RANGE_STEP && p.key[k].item[1].step != 1 &&
          p.key[k].item[1].step != -1)
        col_steps = true;
    const int fat_own = kFatSpan - a.halo_l - a.halo_r;
    const int64_t ntiles = (nx + fat_own - 1) / fat_own;
    // measured on B200 (C4, nx = 2^21): thin 6.65 ms, fat 7.96 ms -- with one 16-warp block per SM the
    // fat kernel is issue/latency-bound, so it stays opt-in (KX_RM_FAT=1) until it is tuned further
    const bool fat_on = getenv("KX_RM_FAT") != nullptr;
    if (!col_steps && fat_on && ntiles >= 96 && ntiles < 0x7fffffff && 4 * p.n_keys + 8 <= kMaxExc) {
      a.own = fat_own;
      a.ntiles = (int)ntiles;
      const unsigned grid = ntiles <= 148 ? (unsigned)ntiles : 128u;
      for (int n0 = row_begin; n0 < row_end; n0 += kT) {
        a.n0 = n0;
        a.n1 = n0 + kT < row_end ? n0 + kT : row_end;
        int rc = launch_fat(a, rl, rr, grid, s);
        if (rc) return rc;
      }
      return KX_OK;
    }
  }
  // otherwise: pick G (columns per thread = 4G) for the thin kernel
  // 1024-column blocks (5 resident per SM, 48 registers) measured 2.7 % faster than 2048-column
  // ones at C4; wider blocks only pay when the redundant halo would eat > 1/8 of a narrow one
  int G = (a.halo_l + a.halo_r <= 128) ? 1 : 2;
  if (nx <= 2048) G = 1;
  if (getenv("KX_RM_G")) G = atoi(getenv("KX_RM_G")) == 1 ? 1 : 2;
  const int span = kRMThreads * 4 * G;
  int own = span - a.halo_l - a.halo_r;
  if (own < 256) return KX_ERR_UNSUPPORTED;
  a.own = own;
  const int64_t blocks = (nx + own - 1) / own;
  if (blocks > 0x7fffffff) return KX_ERR_UNSUPPORTED;
  for (int n0 = row_begin; n0 < row_end; n0 += kT) {
    a.n0 = n0;
    a.n1 = n0 + kT < row_end ? n0 + kT : row_end;
    int rc = (G == 1) ? launch_g<1>(a, rl, rr, (unsigned)blocks, s) : launch_g<2>(a, rl, rr, (unsigned)blocks, s);
    if (rc) return rc;
  }
  return KX_OK;
}

int try_scan_rowmarch(const KxGeom& g, const KxProgram& p, const void* in, const void* coef_dev, void* out,
                      int reverse, cudaStream_t s) {
  (void)in;
  return scan_rowmarch_shard(g, p, coef_dev, out, reverse, 0, g.ndim == 2 ? g.in_shape[1] : 0, s, 0, -1);
}

int scan_rowmarch_rows_per_launch() { return kT; }

}  // namespace kx
