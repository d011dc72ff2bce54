// How fast can ONE SM stream a private chunk out of L2, and how much re-read data does the L2 keep between launches?
// Each of G CTAs sums its own `chunk` bytes (float4 loads, 8 in flight per thread, 512 threads); the launch is repeated over `nbuf`
// different buffers cyclically, so the re-read distance is nbuf * G * chunk bytes.  Prints per-SM and aggregate GB/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_stream_bench tools/l2_stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) stream_kernel(const float4* __restrict__ src, size_t chunk4, float* out) {
  const float4* p = src + (size_t)blockIdx.x * chunk4;
  float acc = 0.f;
  for (size_t i = threadIdx.x; i < chunk4; i += 512 * 8) {
    float4 x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (i + j * 512 < chunk4) ? __ldg(p + i + j * 512) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += x[j].x + x[j].y + x[j].z + x[j].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}
int main(int argc, char** argv) {
  float* out;
  cudaMalloc(&out, 4);
  const int reps = 400;
  for (int G : {16, 64, 148}) {
    for (size_t chunk_kb : {64, 270, 1024}) {
      for (int nbuf : {1, 2, 3, 4, 6, 8}) {
        const size_t chunk = chunk_kb * 1024, total = (size_t)G * chunk;
        if (total * nbuf > (size_t)400 << 20) continue;
        float4* buf;
        cudaMalloc(&buf, total * nbuf);
        cudaMemset(buf, 0, total * nbuf);
        for (int i = 0; i < 2 * nbuf; ++i) stream_kernel<<<G, 512>>>(buf + (size_t)(i % nbuf) * total / 16, chunk / 16, out);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) stream_kernel<<<G, 512>>>(buf + (size_t)(i % nbuf) * total / 16, chunk / 16, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1000.0 / reps;
        printf("G=%3d chunk=%4zu KB re-read distance=%6.1f MB: %7.2f us/launch  %6.1f GB/s per SM  %5.2f TB/s total\n", G, chunk_kb,
               total * nbuf / 1048576.0, us, chunk / us / 1e3, total / us / 1e6);
        cudaFree(buf);
      }
    }
  }
  return 0;
}
