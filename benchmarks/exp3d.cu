// Standalone tuning harness for the 3-D 7-point stencil (BASELINE config 5).  Not part of the
// library: variants that win here are folded into kernex_b200/csrc/kx_stencil3d.cu.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo benchmarks/exp3d.cu -o build/exp3d
//   build/exp3d [N] [reps]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct P {
  const float* in; float* out;
  int n;        // input is n^3, output (n-2)^3 ('valid')
  int zc;       // output planes per block
  float c0, c1; // centre / neighbour coefficient
};

// ---------------------------------------------------------------- reference (naive)
__global__ void ref_kernel(P p) {
  const int no = p.n - 2;
  const int64_t total = (int64_t)no * no * no;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = i % no, y = (i / no) % no, z = i / ((int64_t)no * no);
    const int64_t n = p.n, c = ((int64_t)(z + 1) * n + (y + 1)) * n + (x + 1);
    float a = 0.f;
    a = fmaf(p.c1, p.in[c - n * n], a);
    a = fmaf(p.c1, p.in[c - n], a);
    a = fmaf(p.c1, p.in[c - 1], a);
    a = fmaf(p.c0, p.in[c], a);
    a = fmaf(p.c1, p.in[c + 1], a);
    a = fmaf(p.c1, p.in[c + n], a);
    a = fmaf(p.c1, p.in[c + n * n], a);
    p.out[i] = a;
  }
}

// ---------------------------------------------------------------- V1: cp.async smem ring
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// block: 32 x TYB threads, thread = 4 cols x RY rows; tile TX=128 cols, TY=TYB*RY rows.
// smem tile per stage: (TY+2) rows x (TX+8) floats; tile col 0 <-> input col x0 (x0 % 4 == 0).
// output col j in [x0, x0+128) needs input cols j..j+2 (centre j+1).
template <int TYB, int RY, int STAGES>
__global__ void __launch_bounds__(32 * TYB)
v1_kernel(P p) {
  constexpr int TX = 128, TY = TYB * RY, SW = TX + 8, SR = TY + 2;
  extern __shared__ __align__(16) float smem[];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int n = p.n, no = n - 2;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int z0 = blockIdx.z * p.zc, z1 = min(z0 + p.zc, no);
  const int nplanes = (z1 - z0) + 2;  // input planes z0 .. z1+1

  auto issue = [&](int pl) {  // pl = plane index relative to z0
    float* dst = smem + (size_t)(pl % STAGES) * SR * SW;
    const float* src = p.in + (int64_t)(z0 + pl) * n * n;
    constexpr int CH = SW / 4;  // 16-byte chunks per row
    for (int i = tid; i < SR * CH; i += 32 * TYB) {
      const int r = i / CH, c = (i - r * CH) * 4;
      const int gy = y0 + r, gx = x0 + c;
      const bool ok = (gy < n) && (gx + 3 < n);
      cp_async16(dst + r * SW + c, ok ? (const void*)(src + (int64_t)gy * n + gx) : (const void*)p.in, ok ? 16 : 0);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nplanes) issue(s);
    cp_commit();
  }

  // register z-queue: below[ry][4] = centre values of plane z, mid = plane z+1 (full), above from smem
  float below[RY][4], mid_c[RY][6], mid_t[4], mid_b[4];
  const int jx = 4 * tx;           // thread's first tile column
  const int iy = ty * RY;          // thread's first tile row (output row y0+iy <-> input centre row y0+iy+1)

  auto read_plane = [&](const float* t, float (&c)[RY][6], float (&top)[4], float (&bot)[4]) {
    {
      const float4 a = *reinterpret_cast<const float4*>(t + (iy + 0) * SW + jx);
      const float b = t[(iy + 0) * SW + jx + 4];
      top[0] = a.y; top[1] = a.z; top[2] = a.w; top[3] = b;
    }
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(t + (iy + 1 + r) * SW + jx);
      const float2 b = *reinterpret_cast<const float2*>(t + (iy + 1 + r) * SW + jx + 4);
      c[r][0] = a.x; c[r][1] = a.y; c[r][2] = a.z; c[r][3] = a.w; c[r][4] = b.x; c[r][5] = b.y;
    }
    {
      const float4 a = *reinterpret_cast<const float4*>(t + (iy + 1 + RY) * SW + jx);
      const float b = t[(iy + 1 + RY) * SW + jx + 4];
      bot[0] = a.y; bot[1] = a.z; bot[2] = a.w; bot[3] = b;
    }
  };

  for (int pl = 0; pl < nplanes; ++pl) {
    cp_wait<STAGES - 2>();
    __syncthreads();  // plane pl has landed for everyone; stage (pl-1)%STAGES is free for reuse
    if (pl + STAGES - 1 < nplanes) issue(pl + STAGES - 1);
    cp_commit();
    const float* t = smem + (size_t)(pl % STAGES) * SR * SW;
    if (pl == 0) {
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(t + (iy + 1 + r) * SW + jx);
        const float b = t[(iy + 1 + r) * SW + jx + 4];
        below[r][0] = a.y; below[r][1] = a.z; below[r][2] = a.w; below[r][3] = b;
      }
    } else if (pl == 1) {
      read_plane(t, mid_c, mid_t, mid_b);
    } else {
      // plane pl is "above"; output plane z = z0 + pl - 2
      float up_c[RY][6], up_t[4], up_b[4];
      read_plane(t, up_c, up_t, up_b);
      const int z = z0 + pl - 2;
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        const int oy = y0 + iy + r, ox = x0 + jx;
        float acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = 0.f;
          a = fmaf(p.c1, below[r][i], a);
          a = fmaf(p.c1, (r == 0) ? mid_t[i] : mid_c[r - 1][i + 1], a);
          a = fmaf(p.c1, mid_c[r][i], a);
          a = fmaf(p.c0, mid_c[r][i + 1], a);
          a = fmaf(p.c1, mid_c[r][i + 2], a);
          a = fmaf(p.c1, (r == RY - 1) ? mid_b[i] : mid_c[r + 1][i + 1], a);
          a = fmaf(p.c1, up_c[r][i + 1], a);
          acc[i] = a;
        }
        if (oy < no && ox < no) {
          float* o = p.out + ((int64_t)z * no + oy) * no + ox;
          if (ox + 3 < no && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            __stcs(reinterpret_cast<float4*>(o), make_float4(acc[0], acc[1], acc[2], acc[3]));
          } else if (ox + 3 < no && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
            __stcs(reinterpret_cast<float2*>(o), make_float2(acc[0], acc[1]));
            __stcs(reinterpret_cast<float2*>(o) + 1, make_float2(acc[2], acc[3]));
          } else {
            for (int i = 0; i < 4; ++i) if (ox + i < no) o[i] = acc[i];
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RY; ++r) {
#pragma unroll
        for (int i = 0; i < 4; ++i) below[r][i] = mid_c[r][i + 1];
#pragma unroll
        for (int i = 0; i < 6; ++i) mid_c[r][i] = up_c[r][i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { mid_t[i] = up_t[i]; mid_b[i] = up_b[i]; }
    }
  }
}

template <int TYB, int RY, int STAGES>
float run_v1(P p, int reps, const char* name, const float* ref, int64_t nout) {
  constexpr int TX = 128, TY = TYB * RY;
  const size_t smem = (size_t)STAGES * (TY + 2) * (TX + 8) * sizeof(float);
  CK(cudaFuncSetAttribute(v1_kernel<TYB, RY, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int no = p.n - 2;
  int nzc = (no + p.zc - 1) / p.zc;
  dim3 grid((no + TX - 1) / TX, (no + TY - 1) / TY, nzc), block(32, TYB);
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v1_kernel<TYB, RY, STAGES>, 32 * TYB, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemset(p.out, 0, nout * sizeof(float)));
  v1_kernel<TYB, RY, STAGES><<<grid, block, smem>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  // verify on a sample
  int bad = 0;
  if (ref) {
    std::vector<float> a(1 << 20), b(1 << 20);
    for (int64_t off : {(int64_t)0, nout / 2 - (1 << 19), nout - (1 << 20)}) {
      CK(cudaMemcpy(a.data(), p.out + off, sizeof(float) << 20, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(b.data(), ref + off, sizeof(float) << 20, cudaMemcpyDeviceToHost));
      for (int i = 0; i < (1 << 20); ++i) if (a[i] != b[i]) ++bad;
    }
  }
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    v1_kernel<TYB, RY, STAGES><<<grid, block, smem>>>(p);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double lups = (double)no * no * no;
  printf("%-28s zc=%4d grid=(%d,%d,%d) smem=%zuKB occ=%d  %8.3f ms  %7.1f GLUP/s  %6.0f GB/s  mismatches=%d\n", name, p.zc,
         grid.x, grid.y, grid.z, smem >> 10, occ, best, lups / best / 1e6, lups * 8 / best / 1e6, bad);
  return best;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1024;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  const int64_t nin = (int64_t)n * n * n, no = n - 2, nout = no * no * no;
  float *in, *out, *ref;
  CK(cudaMalloc(&in, nin * sizeof(float)));
  CK(cudaMalloc(&out, nout * sizeof(float)));
  CK(cudaMalloc(&ref, nout * sizeof(float)));
  {
    std::vector<float> h(1 << 22);
    uint32_t s = 12345;
    for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (float)((s >> 8) & 0xffff) / 65536.f - 0.5f; }
    for (int64_t off = 0; off < nin; off += (1 << 22)) {
      const int64_t cnt = (nin - off) < (1 << 22) ? (nin - off) : (1 << 22);
      CK(cudaMemcpy(in + off, h.data(), cnt * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  P p{in, ref, n, 64, -6.f, 1.f};
  ref_kernel<<<148 * 16, 256>>>(p);
  CK(cudaDeviceSynchronize());
  p.out = out;
  for (int zc : {16, 32, 48, 64, 96, 128}) {
    p.zc = zc;
    run_v1<8, 2, 2>(p, reps, "v1 TY=16 RY=2 S=2", ref, nout);
    run_v1<8, 2, 3>(p, reps, "v1 TY=16 RY=2 S=3", ref, nout);
    run_v1<16, 2, 3>(p, reps, "v1 TY=32 RY=2 TYB=16 S=3", ref, nout);
    run_v1<16, 2, 4>(p, reps, "v1 TY=32 RY=2 TYB=16 S=4", ref, nout);
    run_v1<16, 1, 3>(p, reps, "v1 TY=16 RY=1 TYB=16 S=3", ref, nout);
    run_v1<8, 4, 4>(p, reps, "v1 TY=32 RY=4 S=4", ref, nout);
  }
  return 0;
}
