// Stand-alone sm_100a micro-benchmarks behind the decoder / recurrence design choices (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I las_pytorch_b200/csrc -I include tools/microbench.cu -o gpurun_out/microbench -lcuda
// 1. mma_chain: cycles for a chain of tcgen05.mma (M=128, N, K=16) issued by one thread, as a function of chain length
//    and of the number of independent accumulators the chain is spread over (issue time and completion time).
// 2. hop: latency of one cross-SM hand-off "32 producer CTAs write a [64 x 512] bf16 matrix, release a counter;
//    32 consumer CTAs acquire it and bring the whole matrix into shared memory" for TMA vs plain loads and for three
//    ways of releasing.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "umma.cuh"

using namespace las;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------------ 1. MMA chain
template <int N, int NACC, int ATMEM, int M = 128>
__global__ void __launch_bounds__(128, 1) mma_chain_kernel(long long* out, int chain, int reps) {
  constexpr int nacc = NACC, a_tmem = ATMEM;
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;                  // 4 atoms x 16 KB
  uint8_t* sb = base + 4 * 16384;      // 4 atoms x N*128
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    // whole warp walks the loop with warp-uniform values; only the tcgen05 instructions are predicated on one elected lane
    const UmmaLayout la{1, 0, 1024, 16384}, lb{1, 0, 1024, (uint32_t)N * 128u};
    const uint32_t idesc = umma_idesc_bf16(M, N);
    const uint32_t a0 = ptx::smem_u32(sa), b0 = ptx::smem_u32(sb);
    // accumulators at columns [0, nacc*N); A-in-TMEM operand at columns 384.. (64 columns = K 128)
    for (int r = 0; r < reps; ++r) {
      __syncwarp();
      const long long t0 = clock64();
      for (int at = 0; at < chain / 4; ++at) {
        if (ptx::elect_one()) {
          const uint32_t a_addr = a0 + (at & 3) * 16384, b_addr = b0 + (at & 3) * N * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = at * 4 + k;
            const uint32_t d = tmem + (uint32_t)(((NACC <= 4 ? k % NACC : (at & 1) * 4 + k)) * N);
            if (a_tmem) ptx::umma_bf16_ts(d, tmem + 384 + ((at & 1) * 64 + k * 16) / 2, umma_smem_desc(lb, b_addr, k * 16), idesc, i >= nacc);
            else ptx::umma_bf16(d, umma_smem_desc(la, a_addr, k * 16), umma_smem_desc(lb, b_addr, k * 16), idesc, i >= nacc);
          }
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (ptx::elect_one()) ptx::umma_commit(&bar);
      __syncwarp();
      ptx::mbar_wait(&bar, (uint32_t)(r & 1));
      const long long t2 = clock64();
      if (threadIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}


// ------------------------------------------------------------------------------------------------ 1b. two issuing warps
// Does the ~57-cycle SS issue interval belong to the issuing thread or to the tensor pipe?  Two warps issue half of the chain
// each (own accumulator, own commit barrier); time until both halves have completed.
template <int N>
__global__ void __launch_bounds__(128, 1) mma_two_issuers_kernel(long long* out, int chain, int reps, int issuers) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sa = base;
  uint8_t* sb = base + 4 * 16384;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  __shared__ long long t_done[2];
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int w = threadIdx.x >> 5;
  const UmmaLayout la{1, 0, 1024, 16384}, lb{1, 0, 1024, (uint32_t)N * 128u};
  const uint32_t idesc = umma_idesc_bf16(128, N);
  const uint32_t a0 = ptx::smem_u32(sa), b0 = ptx::smem_u32(sb);
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    if (w < issuers) {
      const int per = chain / issuers / 4;  // atoms per issuer
      for (int at = 0; at < per; ++at) {
        if (ptx::elect_one()) {
          const uint32_t a_addr = a0 + ((at + w) & 3) * 16384, b_addr = b0 + ((at + w) & 3) * N * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem + w * N, umma_smem_desc(la, a_addr, k * 16), umma_smem_desc(lb, b_addr, k * 16), idesc, (at | k) != 0);
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::umma_commit(&bar[w]);
      __syncwarp();
      ptx::mbar_wait(&bar[w], (uint32_t)(r & 1));
      if ((threadIdx.x & 31) == 0) t_done[w] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = issuers == 2 ? (t_done[0] > t_done[1] ? t_done[0] : t_done[1]) : t_done[0];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ 2. hop
constexpr int HOP_ROWS = 64, HOP_K = 512, HOP_NCTA = 32, HOP_THREADS = 512;
struct HopParams {
  CUtensorMap tm[2][2];        // [group][parity]
  __nv_bfloat16* buf[2][2];    // [group][parity] [64, 512]
  uint32_t* ctr;               // [2] counters, 32 words apart
  long long* out;
  int rounds, load_mode, sig_mode;
};
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(uint32_t* p, uint32_t v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_relaxed(uint32_t* p, uint32_t v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(HOP_THREADS, 1) hop_kernel(const __grid_constant__ HopParams p) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full;
  const int g = blockIdx.x / HOP_NCTA, nb = blockIdx.x % HOP_NCTA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { ptx::mbar_init(&full, 1); ptx::fence_mbar_init(); }
  __syncthreads();
  uint32_t* my_ctr = p.ctr + g * 32;
  const uint32_t* in_ctr = p.ctr + (1 - g) * 32;
  const uint32_t per_round = (p.sig_mode == 1) ? HOP_NCTA * 8u : HOP_NCTA;
  long long t_start = 0;
  float sink = 0.f;
  for (int r = 0; r < p.rounds; ++r) {
    const int par = r & 1;
    // group 0 consumes what group 1 produced in round r-1; group 1 consumes what group 0 produced in round r
    const int need = (g == 0) ? r : r + 1;
    if (blockIdx.x == 0 && tid == 0 && r == 8) t_start = gtimer();
    if (need > 0) {
      const int src_par = (g == 0) ? ((r - 1) & 1) : par;
      if (p.load_mode == 2) {
        if (tid == 0) {
          while (ld_relaxed(in_ctr) < (uint32_t)need * per_round) {}
          (void)ld_acquire(in_ctr);
        }
        __syncthreads();
      } else if (p.load_mode == 0 || p.load_mode == 3 || p.load_mode == 4) {
        if (warp == 0) {
          if (lane == 0) {
            while (ld_relaxed(in_ctr) < (uint32_t)need * per_round) {}
            (void)ld_acquire(in_ctr);
          }
          __syncwarp();
          if (p.load_mode != 3) asm volatile("fence.proxy.async.global;" ::: "memory");
          const int nbox = p.load_mode == 4 ? 1 : HOP_K / 64;
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&full, HOP_ROWS * 128 * nbox);
            for (int i = 0; i < nbox; ++i) ptx::tma_load_2d(sm + i * 8192, &p.tm[1 - g][src_par], &full, i * 64, 0);
          }
          __syncwarp();
        }
        ptx::mbar_wait(&full, (uint32_t)((need - 1) & 1));
      } else {
        if (tid == 0) {
          while (ld_relaxed(in_ctr) < (uint32_t)need * per_round) {}
          (void)ld_acquire(in_ctr);
        }
        __syncthreads();
        const uint4* src = reinterpret_cast<const uint4*>(p.buf[1 - g][src_par]);
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + tid + j * HOP_THREADS);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = tid + j * HOP_THREADS;       // 16-byte chunk index in the [64, 512] matrix (64 chunks per row)
          const int row = idx >> 6, ch = idx & 63;
          const int atom = ch >> 3, c8 = ch & 7;
          *reinterpret_cast<uint4*>(sm + atom * 8192 + (row >> 3) * 1024 + (row & 7) * 128 + ((c8 ^ (row & 7)) << 4)) = v[j];
        }
        ptx::fence_proxy_async_smem();
        __syncthreads();
      }
      sink += reinterpret_cast<const float*>(sm)[tid];
    }
    // produce: this CTA's 16 columns of all 64 rows (what an LSTM CTA's epilogue writes), 256 threads x 8 bytes
    if (tid < 256) {
      const int row = tid >> 2, q = tid & 3;
      uint2 val = make_uint2((uint32_t)r, (uint32_t)tid);
      if (p.sig_mode != 3) *reinterpret_cast<uint2*>(p.buf[g][par] + (size_t)row * HOP_K + nb * 16 + q * 4) = val;
      if (p.sig_mode == 0 || p.sig_mode == 3) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) red_release(my_ctr, 1u);
      } else if (p.sig_mode == 1) {
        __syncwarp();
        if (lane == 0) red_release(my_ctr, 1u);
      } else {
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) red_relaxed(my_ctr, 1u);
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    p.out[0] = gtimer() - t_start;
    p.out[1] = (long long)sink;
  }
}


// ------------------------------------------------------------------------------------------------ 3. signalling only
struct SigParams { uint32_t* ctr; long long* out; int rounds, poll_mode, rel_mode, ncta, store; float* scratch; };
__global__ void __launch_bounds__(HOP_THREADS, 1) sig_kernel(const SigParams p) {
  const int g = blockIdx.x / p.ncta;
  const int tid = threadIdx.x;
  uint32_t* my_ctr = p.ctr + g * 32;
  const uint32_t* in_ctr = p.ctr + (1 - g) * 32;
  long long t_start = 0;
  for (int r = 0; r < p.rounds; ++r) {
    const int need = (g == 0) ? r : r + 1;
    if (blockIdx.x == 0 && tid == 0 && r == 8) t_start = gtimer();
    if (need > 0) {
      if (tid == 0) {
        const uint32_t target = (uint32_t)need * p.ncta;
        if (p.poll_mode == 0) { while (ld_relaxed(in_ctr) < target) {} (void)ld_acquire(in_ctr); }
        else if (p.poll_mode == 1) { while (ld_acquire(in_ctr) < target) {} }
        else if (p.poll_mode == 2) { while (ld_relaxed(in_ctr) < target) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
        else { while (ld_relaxed(in_ctr) < target) {} }
      }
      __syncthreads();
    }
    if (p.store && tid < 256) p.scratch[(size_t)blockIdx.x * 256 + tid] = (float)r;
    if (tid < 256) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) {
        if (p.rel_mode == 0) red_release(my_ctr, 1u);
        else if (p.rel_mode == 1) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); red_relaxed(my_ctr, 1u); }
        else red_relaxed(my_ctr, 1u);
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0) p.out[0] = gtimer() - t_start;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  long long* out;
  CK(cudaMalloc(&out, 64));
  long long h[2];
  printf("== mma_chain: cycles (issue / complete) for `chain` tcgen05.mma M=128 K=16 over `nacc` accumulators\n");
  auto run = [&](auto kern, int N, int nacc, int a_tmem) {
    for (int chain : {4, 16, 36, 64}) {
      const size_t smem = 4 * 16384 + 4 * N * 128 + 2048;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<1, 128, smem>>>(out, chain, 5);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
      printf("a_tmem=%d N=%3d chain=%2d nacc=%d : issue %5lld  complete %5lld cycles\n", a_tmem, N, chain, nacc, h[0], h[1]);
    }
  };
  run(mma_chain_kernel<16, 1, 0>, 16, 1, 0);
  run(mma_chain_kernel<16, 4, 0>, 16, 4, 0);
  run(mma_chain_kernel<16, 8, 0>, 16, 8, 0);
  run(mma_chain_kernel<64, 1, 0>, 64, 1, 0);
  run(mma_chain_kernel<64, 2, 0>, 64, 2, 0);
  run(mma_chain_kernel<64, 4, 0>, 64, 4, 0);
  run(mma_chain_kernel<128, 1, 0>, 128, 1, 0);
  run(mma_chain_kernel<256, 1, 0>, 256, 1, 0);
  printf("-- M=64 SS:\n");
  run(mma_chain_kernel<64, 1, 0, 64>, 64, 1, 0);
  run(mma_chain_kernel<32, 1, 0, 64>, 32, 1, 0);
  run(mma_chain_kernel<16, 1, 0, 64>, 16, 1, 0);
  run(mma_chain_kernel<128, 1, 0, 64>, 128, 1, 0);
  printf("-- M=128 TS:\n");
  run(mma_chain_kernel<16, 1, 1>, 16, 1, 1);
  run(mma_chain_kernel<16, 4, 1>, 16, 4, 1);
  run(mma_chain_kernel<64, 1, 1>, 64, 1, 1);
  run(mma_chain_kernel<64, 4, 1>, 64, 4, 1);

  printf("== two issuers: cycles until a chain of 32 SS tcgen05.mma (M=128, N=64, K=16) has completed\n");
  for (int issuers = 1; issuers <= 2; ++issuers) {
    const size_t smem2 = 4 * 16384 + 4 * 64 * 128 + 2048;
    CK(cudaFuncSetAttribute(mma_two_issuers_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    mma_two_issuers_kernel<64><<<1, 128, smem2>>>(out, 32, 5, issuers);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("issuing warps=%d : %lld cycles\n", issuers, h[0]);
  }
  printf("== hop: ns per hand-off (32 producer CTAs -> 32 consumer CTAs, [64 x 512] bf16)\n");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  HopParams p;
  memset(&p, 0, sizeof(p));
  for (int g = 0; g < 2; ++g)
    for (int k = 0; k < 2; ++k) {
      CK(cudaMalloc(&p.buf[g][k], HOP_ROWS * HOP_K * 2));
      CK(cudaMemset(p.buf[g][k], 0, HOP_ROWS * HOP_K * 2));
      cuuint64_t gdim[2] = {HOP_K, HOP_ROWS};
      cuuint64_t gstr[1] = {HOP_K * 2};
      cuuint32_t box[2] = {64, HOP_ROWS};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&p.tm[g][k], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.buf[g][k], gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); return 1; }
    }
  CK(cudaMalloc(&p.ctr, 64 * 4));
  p.out = out;
  p.rounds = 208;
  const size_t smem = 64 * 1024 + 2048;
  CK(cudaFuncSetAttribute(hop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* ln[5] = {"TMA 8 boxes", "ld.cg + st.shared", "none (signal only)", "TMA 8, no proxy fence", "TMA 1 box"};
  const char* sn[4] = {"bar + 1 red.release", "per-warp red.release", "threadfence + bar + red.relaxed", "no data stores, red.release"};
  for (int lm = 0; lm < 5; ++lm)
    for (int sg = 0; sg < 4; ++sg) {
      if (sg == 1 || sg == 2) continue;
      p.load_mode = lm;
      p.sig_mode = sg;
      CK(cudaMemset(p.ctr, 0, 64 * 4));
      hop_kernel<<<2 * HOP_NCTA, HOP_THREADS, smem>>>(p);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
      printf("load=%-18s signal=%-32s : %7.1f ns per hop\n", ln[lm], sn[sg], (double)h[0] / (2.0 * (p.rounds - 8)));
    }
  printf("== sig: ns per hop, signalling only (ncta producers -> ncta consumers)\n");
  {
    SigParams q;
    q.ctr = p.ctr; q.out = out; q.rounds = 208;
    CK(cudaMalloc(&q.scratch, 128 * 256 * 4));
    const char* pn[4] = {"relaxed spin + ld.acquire", "ld.acquire spin", "relaxed spin + fence.acq_rel", "relaxed spin only (unordered)"};
    const char* rn[3] = {"red.release", "fence.acq_rel + red.relaxed", "red.relaxed only (unordered)"};
    for (int ncta : {1, 16, 32, 64})
      for (int store = 0; store < 2; ++store)
        for (int pm = 0; pm < 4; ++pm)
          for (int rm = 0; rm < 3; ++rm) {
            if (ncta != 32 && !(pm == 0 && rm == 0) && !(pm == 1 && rm == 0) && !(pm == 3 && rm == 2)) continue;
            q.ncta = ncta; q.poll_mode = pm; q.rel_mode = rm; q.store = store;
            CK(cudaMemset(p.ctr, 0, 64 * 4));
            sig_kernel<<<2 * ncta, HOP_THREADS>>>(q);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
            printf("ncta=%2d stores=%d poll=%-30s release=%-30s : %7.1f ns per hop\n", ncta, store, pn[pm], rn[rm], (double)h[0] / (2.0 * (q.rounds - 8)));
          }
  }
  return 0;
}
