// sm_100a primitives used by the LAS_MODE_BF16 kernels: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05
// (TMEM alloc, UMMA, commit, TMEM load), cluster / DSMEM helpers, and the shared-memory operand layouts the
// UMMA descriptors describe.  Inline PTX only.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace las {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Watchdog for the unbounded CROSS-CTA waits of the persistent kernels (release/acquire counters, flag-in-data slots, GEMM
// tile flags): a wait that has not been satisfied after LAS_SPIN_LIMIT_NS traps, so a protocol bug or a lost peer surfaces
// as a CUDA launch failure (-> RuntimeError in the host mirror) instead of hanging the device.  The trap kills the whole
// grid, so the CTA-local mbarrier waits (which only ever wait for this CTA's own TMA / MMA completions or for a peer that
// is itself behind a guarded wait) need no guard of their own -- measured: guarding them as well costs 4 % per decoder step.
// The clock is only read once every 2^14 failed polls.
#ifndef LAS_SPIN_LIMIT_NS
#define LAS_SPIN_LIMIT_NS 20000000000ll
#endif
#ifndef LAS_GUARD_MBAR
#define LAS_GUARD_MBAR 0
#endif
#ifndef LAS_GUARD_LL
#define LAS_GUARD_LL 1
#endif
struct SpinGuard {
  uint32_t n = 0;
  long long t0 = 0;
  __device__ __forceinline__ void tick() {
    if ((++n & 0x3FFFu) == 0) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > LAS_SPIN_LIMIT_NS) __trap();
    }
  }
};

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if LAS_GUARD_MBAR
  SpinGuard g;
  while (!mbar_try_wait(bar, parity)) g.tick();
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
// cluster-scope acquire: pairs with remote arrives / complete_tx from other CTAs of the cluster
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
#if LAS_GUARD_MBAR
  SpinGuard g;
  while (!mbar_try_wait_cluster(bar, parity)) g.tick();
#else
  while (!mbar_try_wait_cluster(bar, parity)) {
  }
#endif
}
// arrive on the barrier at shared::cluster address `remote_bar` (from mapa), release at cluster scope
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t remote_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(remote_bar), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of the (converged) warp returns true; lets the compiler keep single-thread tcgen05 / TMA issue on the
// uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------- cluster
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync() {
  cluster_arrive();
  cluster_wait();
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// bulk copy own smem -> smem of another CTA in the cluster; completes `bytes` on that CTA's mbarrier
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                     uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(remote_bar)
               : "memory");
}

// ---------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier of this CTA.  c0 = innermost coordinate.
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 2-D tiled store shared -> global (bulk async group)
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] . B[smem desc]; bf16 x bf16 -> fp32; one thread issues for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, A operand read from tensor memory (lane = row, two bf16 per 32-bit column along K)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM, shape 32x32b: thread t of the warp writes lane (lane_base + t), 8 consecutive columns
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive (count 1) on `bar` when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM -> registers, shape 32x32b: thread t of the warp reads lane (lane_base + t), `N` consecutive columns.
__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

}  // namespace ptx

// =========================================================================================================
// UMMA operand layouts in shared memory (K-major operands, 16-bit elements) and their descriptors.
//
// INTERLEAVE (no swizzle): 8x8-element "core matrices" of 128 contiguous bytes (8 rows x 16 B).  Core matrices
//   adjacent along K are `lbo` bytes apart, 8-row groups adjacent along M/N are `sbo` bytes apart.
// SW128: rows of 64 elements (128 B); 8-row groups of 1024 B are `sbo` bytes apart (normally 1024); inside a
//   1024-byte group the 16-byte chunk index is XORed with (row % 8).  K extents beyond 64 elements use further
//   "atoms" (separate tiles).  This is what TMA's SWIZZLE_128B writes for a {64, rows} box.
// =========================================================================================================
struct UmmaLayout {
  uint32_t sw128;  // 0 = INTERLEAVE, 1 = SW128
  uint32_t lbo;    // bytes (INTERLEAVE only)
  uint32_t sbo;    // bytes
  uint32_t atom_bytes;  // SW128 only: bytes between successive 64-element K atoms
};

// byte offset of element (row, k) of a K-major operand
__host__ __device__ __forceinline__ uint32_t umma_offset(const UmmaLayout& L, uint32_t row, uint32_t k) {
  if (L.sw128) {
    const uint32_t atom = k >> 6, kk = k & 63;
    return atom * L.atom_bytes + (row >> 3) * L.sbo + (row & 7) * 128 + ((((kk >> 3) ^ (row & 7)) & 7) << 4) + (kk & 7) * 2;
  }
  return (k >> 3) * L.lbo + (row >> 3) * L.sbo + (row & 7) * 16 + (k & 7) * 2;
}

// 64-bit shared-memory matrix descriptor for the 16-element K slice starting at element column k0
__device__ __forceinline__ uint64_t umma_smem_desc(const UmmaLayout& L, uint32_t base_smem_addr, uint32_t k0) {
  uint32_t addr;
  uint64_t lbo, sbo, type;
  if (L.sw128) {
    addr = base_smem_addr + (k0 >> 6) * L.atom_bytes + (k0 & 63) * 2;
    lbo = 1;  // ignored for swizzled K-major layouts
    sbo = L.sbo >> 4;
    type = 2;  // SWIZZLE_128B
  } else {
    addr = base_smem_addr + (k0 >> 3) * L.lbo;
    lbo = L.lbo >> 4;
    sbo = L.sbo >> 4;
    type = 0;  // SWIZZLE_NONE (interleave)
  }
  return (uint64_t)((addr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46) /* descriptor version 1 (sm_100) */ |
         (type << 61);
}

// 32-bit instruction descriptor: bf16 (or, f16 != 0, IEEE fp16) A/B, both K-major, fp32 accumulate, shape M x N
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, int f16 = 0) {
  const uint32_t fmt = f16 ? 0u : 1u;  // kind::f16 operand format field: 0 = F16, 1 = BF16
  return (1u << 4) /* D = f32 */ | (fmt << 7) /* A */ | (fmt << 10) /* B */ | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace las
