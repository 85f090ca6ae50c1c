// Write-pattern probe for the row-marching kscan: how fast can 2-D sub-array writes go?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a benchmarks/expwrite.cu -o build/expwrite
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// each block owns `span` columns and writes `rows` rows (pitch nx) -- the kscan pattern
template <bool SYNC, bool CS>
__global__ void __launch_bounds__(256) rows_kernel(float* out, long nx, int span, int row0, int rows) {
  const long c0 = (long)blockIdx.x * span;
  for (int r = 0; r < rows; ++r) {
    float* o = out + (long)(row0 + r) * nx + c0;
    for (int c = threadIdx.x * 4; c < span; c += 1024) {
      if (c0 + c + 3 < nx) {
        float4 v = make_float4(r, c, 1.f, 2.f);
        if (CS) __stcs(reinterpret_cast<float4*>(o + c), v); else *reinterpret_cast<float4*>(o + c) = v;
      }
    }
    if (SYNC) __syncthreads();
  }
}

template <bool SYNC, bool CS>
void run(float* out, long nx, int nt, int span, int T, const char* name) {
  const int blocks = (int)((nx + span - 1) / span);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    for (int n0 = 0; n0 < nt; n0 += T) rows_kernel<SYNC, CS><<<blocks, 256>>>(out, nx, span, n0, T);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("%-26s span=%6d blocks=%6d T=%4d  %8.3f ms  %7.0f GB/s\n", name, span, blocks, T, best, (double)nx * nt * 4 / best / 1e6);
}

int main() {
  const long nx = 1L << 21; const int nt = 4096;
  float* out; CK(cudaMalloc(&out, nx * nt * sizeof(float)));
  CK(cudaMemset(out, 0, nx * nt * sizeof(float)));
  for (int T : {64}) for (int span : {2048, 8192, 12288, 16320, 16384, 16448, 20480, 32768, 65536}) {
    run<true, true>(out, nx, nt, span, T, "sync   .cs");
  }
  return 0;
}
