// Why do the decoder's tcgen05.mma chains run at ~60 cycles per instruction when the isolated chain of tools/microbench.cu
// reaches ~33 (A operand in tensor memory)?  Same chain, with the conditions of the kernel switched on one at a time:
//   fill      operand data: zeros (as in microbench.cu) or bf16 1.0
//   atoms     4 shared-memory atoms revisited (descriptors repeat) or 8 distinct ones (as the decoder's 8 K atoms)
//   spinners  extra warps of the CTA spinning in mbarrier.try_wait on the barrier the chain commits to (the decoder's 8 epilogue
//             warps do exactly that while the chain runs)
//   nop       1: the B operand in the no-swizzle INTERLEAVE layout with N = 16 (the attention CTA's context reduction)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I las_pytorch_b200/csrc -I include tools/microbench_mma_insitu.cu -o tools/microbench_mma_insitu.bin
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "umma.cuh"

using namespace las;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int N, int ATMEM, int INTERLEAVE>
__global__ void __launch_bounds__(512, 1) chain_kernel(long long* out, int chain, int reps, uint32_t fill, int natoms, int spinners) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sa = base;                       // natoms x 16 KB (128 rows x 64 bf16, SW128)
  uint8_t* sb = base + natoms * 16384;      // natoms x N*128
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int total16 = (natoms * 16384 + natoms * N * 128) / 16;
  for (int i = threadIdx.x; i < total16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(fill, fill, fill, fill);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) ptx::tmem_alloc(&slot, 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp < 4) {  // A operand region of tensor memory: columns 256..511
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fill;
    for (int c = 256; c < 512; c += 8) ptx::tmem_st_32x32b_x8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) {
    const UmmaLayout la{1, 0, 1024, 16384};
    const UmmaLayout lb = INTERLEAVE ? UmmaLayout{0, 256, 128, 0} : UmmaLayout{1, 0, 1024, (uint32_t)N * 128u};
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a0 = ptx::smem_u32(sa), b0 = ptx::smem_u32(sb);
    for (int r = 0; r < reps; ++r) {
      __syncwarp();
      const long long t0 = clock64();
      if (ptx::elect_one()) {
        for (int at = 0; at < chain / 4; ++at) {
          const int ai = at % natoms;
          const uint32_t a_addr = a0 + ai * 16384, b_addr = b0 + ai * N * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = INTERLEAVE ? umma_smem_desc(lb, b0, ((at * 4 + k) % 16) * 16) : umma_smem_desc(lb, b_addr, k * 16);  // 16 distinct K steps (8 KB)
            if (ATMEM) ptx::umma_bf16_ts(tmem, tmem + 256 + (uint32_t)((at * 4 + k) % 32) * 8, bd, idesc, (at | k) != 0);
            else ptx::umma_bf16(tmem, umma_smem_desc(la, a_addr, k * 16), bd, idesc, (at | k) != 0);
          }
        }
      }
      __syncwarp();
      const long long t1 = clock64();
      if (ptx::elect_one()) ptx::umma_commit(&bar);
      __syncwarp();
      ptx::mbar_wait(&bar, (uint32_t)(r & 1));
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (warp >= 4 && warp < 4 + spinners) {
    for (int r = 0; r < reps; ++r) ptx::mbar_wait(&bar, (uint32_t)(r & 1));  // what the decoder's epilogue warps do during the chain
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  CK(cudaMalloc(&out, 64));
  long long h[2];
  auto run = [&](auto kern, const char* name, int N, int chain) {
    for (int natoms : {4, 8})
      for (uint32_t fill : {0u, 0x3F803F80u})
        for (int spinners : {0, 8}) {
          const size_t smem = (size_t)natoms * 16384 + (size_t)natoms * N * 128 + 2048;
          CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          kern<<<1, 512, smem>>>(out, chain, 5, fill, natoms, spinners);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
          printf("%-22s N=%3d chain=%2d atoms=%d data=%s spinning warps=%d : issue %5lld  complete %5lld cycles  (%.1f / instruction)\n", name, N, chain,
                 natoms, fill ? "ones " : "zeros", spinners, h[0], h[1], (double)h[1] / chain);
        }
  };
  run(chain_kernel<64, 0, 0>, "SS  (LSTM CTA form)", 64, 32);
  run(chain_kernel<64, 1, 0>, "TS  (A in TMEM)", 64, 32);
  run(chain_kernel<16, 1, 1>, "TS  N=16 interleave", 16, 52);
  return 0;
}
