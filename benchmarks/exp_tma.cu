// Microbenchmark: per-warp TMA bulk-copy (cp.async.bulk) streaming throughput on B200 as a function
// of copy size, ring depth and resident warps.  Each warp streams `rows` rows of `bytes` bytes
// (row pitch `pitch` bytes) through a private ring and touches one word per lane of every row.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o exp_tma exp_tma.cu && ./exp_tma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../kernex_b200/csrc/kx_tma.cuh"
using namespace kx;

template <int D>
__global__ void __launch_bounds__(128) stream_kernel(const float* in, float* sink, int bytes, int64_t pitch_f, int rows,
                                                     int copies, int64_t plane_f, int split) {
  extern __shared__ __align__(128) unsigned char dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stage_bytes = bytes * copies;
  unsigned ring = (unsigned)__cvta_generic_to_shared(dyn) + (unsigned)(warp * D * stage_bytes);
  __shared__ __align__(8) uint64_t bars[4 * D];
  unsigned mbar = (unsigned)__cvta_generic_to_shared(bars) + warp * D * 8;
  if (lane < D) mbar_init(mbar + lane * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncwarp();
  // image of pitch_f floats per row; blockIdx.x / warp pick the column strip, blockIdx.y the band of `rows` rows;
  // `copies` > 1 = that many planes (conv channels) 2^28 bytes apart
  const int64_t gw = (int64_t)blockIdx.x * 4 + warp;
  const float* src = in + gw * (bytes / 4) + (int64_t)blockIdx.y * rows * pitch_f;
  float acc = 0.f;
  auto issue = [&](int r) {
    if (lane == 0 && r < rows) {
      const unsigned bar = mbar + (r % D) * 8;
      mbar_expect_tx(bar, stage_bytes);
      const int sub = bytes / split;
      for (int c = 0; c < copies; ++c)
        for (int q = 0; q < split; ++q)
          bulk_g2s(ring + (r % D) * stage_bytes + c * bytes + q * sub, src + (int64_t)r * pitch_f + c * plane_f + q * (sub / 4),
                   sub, bar);
    }
  };
  for (int r = 0; r < D - 1; ++r) issue(r);
  for (int r = 0; r < rows; ++r) {
    __syncwarp();
    issue(r + D - 1);
    mbar_wait(mbar + (r % D) * 8, (r / D) & 1);
    acc += lds_1<float>(ring + (r % D) * stage_bytes + lane * 4);
  }
  if (acc == 123.456f) sink[0] = acc;
}

template <int D>
float run(const float* in, float* sink, dim3 blocks, int bytes, int64_t pitch_f, int rows, int copies, int64_t plane_f, int split = 1) {
  size_t smem = (size_t)4 * D * bytes * copies;
  cudaFuncSetAttribute(stream_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  stream_kernel<D><<<blocks, 128, smem>>>(in, sink, bytes, pitch_f, rows, copies, plane_f, split);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) stream_kernel<D><<<blocks, 128, smem>>>(in, sink, bytes, pitch_f, rows, copies, plane_f, split);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return ms / 5;
}

int main() {
  const int64_t total = (int64_t)2 << 30;  // 2 GiB streamed per launch
  float* in;
  float* sink;
  cudaMalloc(&in, total + (1 << 20));
  cudaMalloc(&sink, 4);
  cudaMemset(in, 0, total + (1 << 20));
  const int64_t W = 8192;                      // floats per image row (32 KB pitch, as in the 8192^2 configs)
  for (int split : {1, 2, 4, 8}) {
    const int copies = 1, rows = 32;
    const int64_t H = total / 4 / W;
    for (int bytes : {1024, 2048}) {
      for (int D : {3, 4, 6}) {
        dim3 blocks((unsigned)(W * 4 / (4 * bytes)), (unsigned)(H / rows));
        float ms = 0;
        if (D == 3) ms = run<3>(in, sink, blocks, bytes, W, rows, copies, H * W, split);
        if (D == 4) ms = run<4>(in, sink, blocks, bytes, W, rows, copies, H * W, split);
        if (D == 6) ms = run<6>(in, sink, blocks, bytes, W, rows, copies, H * W, split);
        printf("split=%d bytes=%4d depth=%d smem/block=%6d  %.3f ms  %.0f GB/s\n", split, bytes, D, 4 * D * bytes, ms,
               (double)total / (ms * 1e-3) / 1e9);
      }
    }
  }
  return 0;
}
