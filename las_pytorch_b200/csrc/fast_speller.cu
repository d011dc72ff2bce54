// LAS_MODE_BF16 speller: the whole attention-decoder step loop (model/las_model.py:209-236) as ONE persistent
// cooperative kernel.  No launch, host sync or PCIe traffic between decode steps.
//
// CTA roles (one CTA per SM, all co-resident; cudaLaunchAttributeCooperative guarantees it):
//   * LSTM CTAs: layer l, block nb owns 16 hidden units (64 gate columns, unit-major / gate-minor).  Its slice of
//     [W_hh | W_ih] (bf16, 128-byte-swizzled K-major atoms) stays in shared memory for all S steps as the B operand;
//     the A operand is the activation matrix [batch (M = 128 TMEM lanes), K] streamed per step by TMA through a
//     4-stage ring: first the layer's own h_{s-1} (available early), then the critical input ([word | context] for
//     layer 0, the lower layer's fresh h otherwise).  tcgen05.mma accumulates both parts in TMEM; the epilogue thread
//     of batch row b holds that row's cell state c in registers (fp32) and writes h (bf16 operand copy + fp32 copy).
//   * attention CTAs: one per utterance.  W_phi, W_cd (bf16) and psi(enc)[b] (fp32) are resident in shared memory;
//     per step: q = relu(W_phi h + b), energies, length-masked softmax, context (streams enc[b] as bf16, 16-byte
//     loads), character distribution, log-softmax, argmax / teacher forcing, and the next LSTM input row.
// Hand-off between roles is through global-memory buffers (double-buffered by step parity) and monotonically
// increasing release/acquire counters; TMA reads of freshly written activations are ordered by fence.proxy.async.
#include <cuda.h>
#include <string.h>

#include "las_fast.cuh"
#include "las_kernels.cuh"
#include "umma.cuh"

namespace las {

int make_tmap_bf16_box(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);  // fast_gemm.cu

namespace {

constexpr int DEC_NW = 64;        // gate columns per LSTM CTA (16 hidden units x 4 gates)
constexpr int DEC_UNITS = DEC_NW / 4;
constexpr int DEC_MAX_STAGES = 16;
constexpr int DEC_THREADS = 512;
constexpr int DEC_VP = 64;        // one-hot / word columns padded to one 64-wide K atom
constexpr int WATOM_BYTES = DEC_NW * 128;  // B atom: 64 gate columns x 64 bf16
constexpr int MAX_SL = 4;

struct DecParams {
  CUtensorMap tm_h[MAX_SL][2];  // hbuf[l][parity]  bf16 [B, Hs]
  CUtensorMap tm_x[2];          // xbuf[parity]     bf16 [B, VP + E]
  const uint8_t* w_img[MAX_SL]; // [ncl][atoms_l] swizzled 64x64 bf16 atoms: h part first, then input part
  const float* bias[MAX_SL];    // [ncl*64] b_ih + b_hh in CTA column order
  __nv_bfloat16* hbuf[MAX_SL][2];
  float* hf32[2];               // top layer h, fp32 [B, Hs]
  __nv_bfloat16* xbuf[2];
  const float* c_init;          // nullable [sl, B, Hs]
  float* h_out;                 // nullable [sl, B, Hs]
  float* c_out;                 // nullable [sl, B, Hs]
  const __nv_bfloat16* enc;     // [B, U, E] bf16
  const float* psi;             // [B, U, D] fp32
  const __nv_bfloat16* w_phi;   // [D, Hs] bf16
  const float* b_phi;
  const __nv_bfloat16* w_cd;    // [V, Hs + E] bf16
  const float* b_cd;
  const float* gt_dense;        // nullable [B, gt_steps, V]
  const int32_t* gt_index;      // nullable [B, gt_steps]
  const int32_t* enc_lengths;   // nullable [B]
  float* logp;                  // [S, B, V]
  float* attn;                  // nullable [S, B, U]
  int32_t* tokens;              // nullable [S, B]
  float* word_out;              // nullable [B, V]
  float* ctx_out;               // nullable [B, E]
  uint32_t* sync;               // counters, 32 uint32 apart: [l] = h_ready[l], [MAX_SL] = ctx_ready
  int B;      // utterances handled by this launch (one attention CTA each)
  int Bfull;  // batch pitch of the caller's tensors; this launch covers utterances [b0, b0 + B)
  int b0;
  int U, E, Hs, sl, V, D, steps, decode_mode, relu, gt_steps, ncl, k_in_smem;
  int ctx_tmem;  // 1: enc[b]^T is resident in the attention CTA's tensor memory and the context is a UMMA
  int nstages, stage_bytes;  // activation ring: stage = [box_rows (64 or 128) batch rows x 64 bf16], 128-byte swizzled
  long long* trace;  // nullable test hook: [3 roles][32 steps][8] globaltimer stamps (layer-0 CTA 0, top-layer CTA 0, attention CTA 0)
};

__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DEC_TRACE(role, slot) do { if (p.trace && s < 32) p.trace[((role) * 32 + s) * 8 + (slot)] = gtimer(); } while (0)

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin with relaxed loads (an acquire load also invalidates L1 on every poll), then one acquire load to synchronise
__device__ __forceinline__ void wait_counter(const uint32_t* ctr, uint32_t target) {
  while (ld_relaxed(ctr) < target) {
  }
  (void)ld_acquire(ctr);
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint32_t* counter(const DecParams& p, int idx) { return p.sync + idx * 32; }

__device__ __forceinline__ float block_max_512(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < DEC_THREADS / 32; ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum_512(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < DEC_THREADS / 32; ++i) r += red[i];
  return r;
}

// ------------------------------------------------------------------------------------------------------------
// LSTM role
// ------------------------------------------------------------------------------------------------------------
__device__ void lstm_role(const DecParams& p, uint8_t* smem, int l, int nb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = (p.Hs + 63) / 64;
  const int nx = (l == 0) ? (DEC_VP + p.E + 63) / 64 : (p.Hs + 63) / 64;
  const int natoms = nh + nx;
  const int NBUF = p.nstages, STAGE_BYTES = p.stage_bytes;  // NBUF >= max(nh, nx): one slot per atom of a part
  // Activation buffer first, weights right behind it: with 64-row slots the UMMA (M = 128) also reads the 8 KB that
  // follow a slot (the next slot or the first weight atom); those rows only feed accumulator lanes >= 64, never read.
  uint8_t* abuf = smem;                                  // NBUF x STAGE_BYTES
  uint8_t* wsm = abuf + (size_t)NBUF * STAGE_BYTES;      // natoms x 8 KB
  float* bias_s = reinterpret_cast<float*>(wsm + (size_t)natoms * WATOM_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + DEC_NW);
  uint64_t* full = bars;                         // [NBUF] one per slot: TMA bytes landed
  uint64_t* part_empty = bars + DEC_MAX_STAGES;  // all MMAs of the previous part have read the buffer
  uint64_t* tmem_full = part_empty + 1;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  constexpr int EPI_WARPS = 8, EPI_THREADS = EPI_WARPS * 32;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) ptx::mbar_init(&full[i], 1);
    ptx::mbar_init(part_empty, 1);
    ptx::mbar_init(tmem_full, 1);
    ptx::mbar_init(tmem_empty, EPI_THREADS);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 128);
  {  // resident weight slice + bias
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img[l] + (size_t)nb * natoms * WATOM_BYTES);
    uint4* dst = reinterpret_cast<uint4*>(wsm);
    for (int i = threadIdx.x; i < natoms * WATOM_BYTES / 16; i += DEC_THREADS) dst[i] = src[i];
    if (threadIdx.x < DEC_NW) bias_s[threadIdx.x] = p.bias[l][nb * DEC_NW + threadIdx.x];
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int S = p.steps;
  uint32_t* my_ready = counter(p, l);
  const int trole = (nb == 0 && l == 0) ? 0 : ((nb == 0 && l == p.sl - 1) ? 1 : -1);

  if (warp == 0) {
    // ============================ TMA producer ============================
    // Each step has two parts: (0) the layer's own h_{s-1}, complete once every CTA of this layer finished step s-1
    // (available long before it is needed), and (1) the critical input: [word | context] of step s-1 for layer 0, the
    // lower layer's h of THIS step otherwise.  All atoms of a part are issued back to back into their own slots.
    const uint32_t* own_ctr = counter(p, l);
    const uint32_t* in_ctr = (l == 0) ? counter(p, MAX_SL) : counter(p, l - 1);
    int n = 0;  // running part index
    for (int s = 0; s < S; ++s) {
      const int par = s & 1;
      for (int part = 0; part < 2; ++part, ++n) {
        if (n > 0) ptx::mbar_wait(part_empty, (uint32_t)((n - 1) & 1));
        if (lane == 0) {
          if (part == 0) wait_counter(own_ctr, (uint32_t)s * p.ncl);
          else wait_counter(in_ctr, (l == 0) ? (uint32_t)s * p.B : (uint32_t)(s + 1) * p.ncl);
          if (part == 1 && trole >= 0) DEC_TRACE(trole, 0);
        }
        __syncwarp();
        fence_proxy_async_global();  // other SMs' generic-proxy stores (acquired above) -> this warp's TMA reads
        const CUtensorMap* tm = (part == 0) ? &p.tm_h[l][par] : ((l == 0) ? &p.tm_x[par] : &p.tm_h[l - 1][par ^ 1]);
        const int na = part == 0 ? nh : nx;
        if (ptx::elect_one()) {
          for (int i = 0; i < na; ++i) {
            ptx::mbar_arrive_expect_tx(&full[i], STAGE_BYTES);
            ptx::tma_load_2d(abuf + (size_t)i * STAGE_BYTES, tm, &full[i], i * 64, 0);
          }
        }
        __syncwarp();
        if (part == 1 && lane == 0 && trole >= 0) DEC_TRACE(trole, 1);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    const UmmaLayout la{1, 0, 1024, (uint32_t)STAGE_BYTES}, lb{1, 0, 1024, WATOM_BYTES};
    const uint32_t idesc = umma_idesc_bf16(128, DEC_NW);
    const uint32_t a0 = ptx::smem_u32(abuf), w_addr = ptx::smem_u32(wsm);
    uint32_t phase_bits = 0;  // per-slot phase parity
    for (int s = 0; s < S; ++s) {
      ptx::mbar_wait(tmem_empty, (uint32_t)((s & 1) ^ 1));
      ptx::tc_fence_after();
      for (int part = 0; part < 2; ++part) {
        const int na = part == 0 ? nh : nx;
        const uint32_t d = tmem + (part == 0 ? 0u : (uint32_t)DEC_NW);  // separate accumulators for the two parts
        for (int i = 0; i < na; ++i) {
          ptx::mbar_wait(&full[i], (phase_bits >> i) & 1u);
          phase_bits ^= 1u << i;
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t a_addr = a0 + i * STAGE_BYTES, b_addr = w_addr + (part == 0 ? i : nh + i) * WATOM_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(d, umma_smem_desc(la, a_addr, k * 16), umma_smem_desc(lb, b_addr, k * 16), idesc, !(i == 0 && k == 0));
            if (i == na - 1) {
              ptx::umma_commit(part_empty);
              if (part == 1) ptx::umma_commit(tmem_full);
            }
          }
          __syncwarp();
        }
        if (part == 0 && lane == 0 && trole >= 0) DEC_TRACE(trole, 6);
      }
      if (lane == 0 && trole >= 0) DEC_TRACE(trole, 2);
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ============================ epilogue: gates, cell state, h ============================
    // 8 warps: TMEM lane quadrant = warp & 3 (batch rows), column half = (warp - 2) / 4 (8 of the CTA's 16 units each)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int b = q * 32 + lane;                       // batch row = TMEM lane
    const int u0 = nb * DEC_UNITS + half * 8;          // first hidden unit of this thread
    const int c0 = half * 32;                          // first accumulator column
    const bool live = b < p.B;
    const bool warp_live = q * 32 < p.B;
    float c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = (live && p.c_init) ? p.c_init[((size_t)l * p.Bfull + p.b0 + b) * p.Hs + u0 + i] : 0.f;
    const bool top = (l == p.sl - 1);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_wait(tmem_full, (uint32_t)(s & 1));
      if (warp == 2 && lane == 0 && trole >= 0) DEC_TRACE(trole, 3);
      ptx::tc_fence_after();
      float h[8];
      if (warp_live) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t a0[16], a1[16];
          ptx::tmem_ld_32x32b_x16(tmem + ((uint32_t)(q * 32) << 16) + c0 + ch * 16, a0);
          ptx::tmem_ld_32x32b_x16(tmem + ((uint32_t)(q * 32) << 16) + DEC_NW + c0 + ch * 16, a1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = c0 + ch * 16 + u * 4;
            const float pi = __uint_as_float(a0[u * 4 + 0]) + __uint_as_float(a1[u * 4 + 0]) + bias_s[col + 0];
            const float pf = __uint_as_float(a0[u * 4 + 1]) + __uint_as_float(a1[u * 4 + 1]) + bias_s[col + 1];
            const float pg = __uint_as_float(a0[u * 4 + 2]) + __uint_as_float(a1[u * 4 + 2]) + bias_s[col + 2];
            const float po = __uint_as_float(a0[u * 4 + 3]) + __uint_as_float(a1[u * 4 + 3]) + bias_s[col + 3];
            const int ui = ch * 4 + u;
            const float cn = sigmoid_fast(pf) * c[ui] + sigmoid_fast(pi) * tanh_fast(pg);
            c[ui] = cn;
            h[ui] = sigmoid_fast(po) * tanh_fast(cn);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(tmem_empty);
      if (live) {
        const int np = (s + 1) & 1;
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 t = __floats2bfloat162_rn(h[2 * i], h[2 * i + 1]);
          pk[i] = *reinterpret_cast<const uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(p.hbuf[l][np] + (size_t)b * p.Hs + u0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (top) {
          float4* df = reinterpret_cast<float4*>(p.hf32[np] + (size_t)b * p.Hs + u0);
          df[0] = make_float4(h[0], h[1], h[2], h[3]);
          df[1] = make_float4(h[4], h[5], h[6], h[7]);
        }
        if (s == S - 1) {
          if (p.h_out) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p.h_out[((size_t)l * p.Bfull + p.b0 + b) * p.Hs + u0 + i] = h[i];
          }
          if (p.c_out) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p.c_out[((size_t)l * p.Bfull + p.b0 + b) * p.Hs + u0 + i] = c[i];
          }
        }
      }
      if (warp == 2 && lane == 0 && trole >= 0) DEC_TRACE(trole, 4);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // all rows stored; the release below is cumulative over the barrier
      if (warp == 2 && lane == 0) {
        red_release_add(my_ready, 1u);
        if (trole >= 0) DEC_TRACE(trole, 5);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------------------
// attention role (one utterance)
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int att_kstride(int D) { return ((D + 3) & ~3) + 4; }  // padded psi row (floats), 16-byte multiple

__device__ void attention_role(const DecParams& p, uint8_t* smem, int b) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = DEC_THREADS / 32;
  const int Hs = p.Hs, E = p.E, U = p.U, D = p.D, V = p.V, KC = p.Hs + p.E;
  const int KS = att_kstride(D);
  const int gb = p.b0 + b;  // utterance index in the caller's tensors
  // ---- shared-memory carve-up
  const int WPS = Hs + 16, WCS = KC + 32;  // padded row strides (bf16): shift consecutive rows by 32 / 64 bytes across the banks
  __nv_bfloat16* s_wphi = reinterpret_cast<__nv_bfloat16*>(smem);             // [D][WPS]
  __nv_bfloat16* s_wcd = s_wphi + (size_t)D * WPS;                             // [V][WCS]
  float* s_f = reinterpret_cast<float*>(s_wcd + (((size_t)V * WCS + 7) & ~(size_t)7));
  float* s_h = s_f;                 // [Hs]
  float* s_ctx = s_h + Hs;          // [E]   (contiguous after s_h: [h | ctx] is the character-distribution input)
  float* s_q = s_ctx + E;           // [KS] (zero padded)
  float* s_score = s_q + KS;        // [U]
  float* s_logit = s_score + U;     // [V]
  float* s_bphi = s_logit + V;      // [D]
  float* s_bcd = s_bphi + D;        // [V]
  float* s_red = s_bcd + V;         // [32]
  float* s_part = s_f + (((size_t)Hs + E + KS + D + U + 2 * V + 32 + 3) & ~(size_t)3);  // [nrg][E] <= 4096 floats, 16-byte aligned
  float* s_k = s_part + 4096;       // [U][KS] when k_in_smem
  const int ncg = E / 8;            // 8-column groups of enc
  const int nrg = DEC_THREADS / ncg;  // row groups working in parallel
  // tensor-memory context path: s_part's space holds the mbarrier, the TMEM slot and the score operand instead
  uint64_t* ctx_bar = reinterpret_cast<uint64_t*>(s_part);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_part + 2);
  uint8_t* s_bop = reinterpret_cast<uint8_t*>(s_part + 4);   // [2*nks core-K][2][8 rows][16 B]: row 0 = scores (bf16)
  float* s_lh = s_logit;                                      // h-part of the logits is accumulated in place
  const int nks = (U + 15) / 16;                              // UMMA K steps over the encoder axis
  const int CU = nks * 8;                                     // TMEM columns of one 128-feature tile of enc^T
  const int NT = E / 128;                                     // feature tiles

  for (int i = tid; i < D * Hs; i += DEC_THREADS) s_wphi[(size_t)(i / Hs) * WPS + (i % Hs)] = p.w_phi[i];
  for (int i = tid; i < V * KC; i += DEC_THREADS) s_wcd[(size_t)(i / KC) * WCS + (i % KC)] = p.w_cd[i];
  for (int i = tid; i < KS; i += DEC_THREADS) s_q[i] = 0.f;
  for (int i = tid; i < D; i += DEC_THREADS) s_bphi[i] = p.b_phi[i];
  for (int i = tid; i < V; i += DEC_THREADS) s_bcd[i] = p.b_cd[i];
  const float* psib = p.psi + (size_t)gb * U * D;
  if (p.k_in_smem) {
    for (int i = tid; i < U * KS; i += DEC_THREADS) {
      const int u = i / KS, d = i % KS;
      s_k[i] = d < D ? psib[(size_t)u * D + d] : 0.f;
    }
  }
  const int ulen = p.enc_lengths ? min(max(p.enc_lengths[gb], 1), U) : U;
  const __nv_bfloat16* encb = p.enc + (size_t)b * U * E;
  const uint32_t* h_ctr = counter(p, p.sl - 1);
  uint32_t* ctx_ctr = counter(p, MAX_SL);
  const int dchunk = ((D + 3) / 4 + 3) & ~3;  // psi columns per lane of a 4-lane row team, multiple of 4
  uint32_t tmem = 0;
  if (p.ctx_tmem) {
    // enc[b]^T -> tensor memory, once: tile t holds features [128t, 128t+128) as TMEM lanes, encoder steps along the
    // columns (two bf16 per 32-bit column) = the A operand of  ctx^T[E,1] = enc^T[E,U] . score[U,1]
    if (tid == 0) {
      ptx::mbar_init(ctx_bar, 1);
      ptx::fence_mbar_init();
    }
    if (warp == 0) ptx::tmem_alloc(tmem_slot, 512);
    for (int i = tid; i < nks * 512 / 16; i += DEC_THREADS) reinterpret_cast<uint4*>(s_bop)[i] = make_uint4(0, 0, 0, 0);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    tmem = *tmem_slot;
    const int qd = warp & 3;
    for (int t = warp >> 2; t < NT; t += NWARP / 4) {
      const __nv_bfloat16* col = encb + t * 128 + qd * 32 + lane;
      for (int ks = 0; ks < nks; ++ks) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u0 = ks * 16 + 2 * i;
          const unsigned short lo = u0 < U ? *reinterpret_cast<const unsigned short*>(col + (size_t)u0 * E) : 0;
          const unsigned short hi = u0 + 1 < U ? *reinterpret_cast<const unsigned short*>(col + (size_t)(u0 + 1) * E) : 0;
          v[i] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        ptx::tmem_st_32x32b_x8(tmem + ((uint32_t)(qd * 32) << 16) + t * CU + ks * 8, v);
      }
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();

  for (int s = 0; s < p.steps; ++s) {
    const int np = (s + 1) & 1;
    if (tid == 0) {
      wait_counter(h_ctr, (uint32_t)(s + 1) * p.ncl);  // acquire; the CTA barrier below extends it to the other threads
      if (b == 0) DEC_TRACE(2, 0);
    }
    __syncthreads();
    for (int k = tid; k < Hs; k += DEC_THREADS) s_h[k] = __ldcg(p.hf32[np] + (size_t)b * Hs + k);
    __syncthreads();

    // q = act(W_phi . h + b_phi)   (model/las_model.py:278): 8 lanes per output, interleaved 4-byte columns
    {
      const int part = tid & 7;
      const float2* hv = reinterpret_cast<const float2*>(s_h);
      for (int d = tid >> 3; d < ((D + 3) & ~3); d += DEC_THREADS / 8) {
        float a0 = 0.f, a1 = 0.f;
        if (d < D) {
          const __nv_bfloat162* wr = reinterpret_cast<const __nv_bfloat162*>(s_wphi + (size_t)d * WPS);
          int kp = part;
          for (; kp + 8 < Hs / 2; kp += 16) {
            const float2 w0 = __bfloat1622float2(wr[kp]), w1 = __bfloat1622float2(wr[kp + 8]);
            const float2 x0 = hv[kp], x1 = hv[kp + 8];
            a0 = fmaf(w0.x, x0.x, a0); a0 = fmaf(w0.y, x0.y, a0);
            a1 = fmaf(w1.x, x1.x, a1); a1 = fmaf(w1.y, x1.y, a1);
          }
          for (; kp < Hs / 2; kp += 8) {
            const float2 w0 = __bfloat1622float2(wr[kp]);
            const float2 x0 = hv[kp];
            a0 = fmaf(w0.x, x0.x, a0); a0 = fmaf(w0.y, x0.y, a0);
          }
        }
        float acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (part == 0 && d < D) {
          acc += s_bphi[d];
          s_q[d] = p.relu ? fmaxf(acc, 0.f) : acc;
        }
      }
    }
    __syncthreads();
    if (tid == 0 && b == 0) DEC_TRACE(2, 1);

    // energy[u] = <q, psi[b,u,:]>  (:289-291).  A team of 4 lanes (lane, lane^8, lane^16, lane^24) shares one encoder
    // step, each lane taking a contiguous quarter of the D columns with 16-byte loads; a warp covers 8 steps per pass
    // and the 8 lanes of a quarter-warp hit 8 different rows -> conflict-free with the padded row stride.
    {
      const int urow = lane & 7, part = lane >> 3;
      const int d0 = part * dchunk;
      for (int ub = warp * 8; ub < U; ub += NWARP * 8) {
        const int u = ub + urow;
        float acc = 0.f;
        if (u < U) {
          if (p.k_in_smem) {
            const float* kr = s_k + (size_t)u * KS;
            float acc2 = 0.f;
            for (int d = d0; d < d0 + dchunk && d < KS - 4; d += 4) {  // q and psi rows are zero padded to KS
              const float4 kv = *reinterpret_cast<const float4*>(kr + d);
              const float4 qv = *reinterpret_cast<const float4*>(s_q + d);
              acc = fmaf(qv.x, kv.x, acc); acc2 = fmaf(qv.y, kv.y, acc2); acc = fmaf(qv.z, kv.z, acc); acc2 = fmaf(qv.w, kv.w, acc2);
            }
            acc += acc2;
          } else {
            const float* kr = psib + (size_t)u * D;
            for (int d = d0; d < d0 + dchunk && d < D; ++d) acc = fmaf(s_q[d], kr[d], acc);
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (part == 0 && u < U) s_score[u] = (u < ulen) ? acc : -INFINITY;
      }
    }
    __syncthreads();
    if (tid == 0 && b == 0) DEC_TRACE(2, 2);

    // softmax over encoder steps (:292): every warp reduces max / sum redundantly with shuffles (no CTA barriers)
    {
      float m = -INFINITY;
      for (int u = lane; u < U; u += 32) m = fmaxf(m, s_score[u]);
      m = warp_max(m);
      float ssum = 0.f;
      for (int u = lane; u < U; u += 32) ssum += __expf(s_score[u] - m);
      ssum = warp_sum(ssum);
      const float inv = 1.0f / ssum;
      __syncthreads();  // everyone has read the raw energies
      for (int u = tid; u < U; u += DEC_THREADS) {
        const float a = __expf(s_score[u] - m) * inv;
        s_score[u] = a;
        if (p.attn) p.attn[((size_t)s * p.Bfull + gb) * U + u] = a;
      }
    }
    __syncthreads();
    if (tid == 0 && b == 0) DEC_TRACE(2, 3);

    if (p.ctx_tmem) {
      // context via the tensor core: scores (bf16) are row 0 of a 16-row K-major operand in shared memory,
      // D_t[128 features, 16] = enc^T tile (TMEM) . scores^T ; column 0 of each accumulator is the context
      for (int u = tid; u < U; u += DEC_THREADS) {
        const __nv_bfloat16 a = __float2bfloat16_rn(s_score[u]);
        *reinterpret_cast<__nv_bfloat16*>(s_bop + (u >> 3) * 256 + (u & 7) * 2) = a;
      }
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (warp == 0) {
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const UmmaLayout lb{0, 256, 128, 0};
          const uint32_t idesc = umma_idesc_bf16(128, 16);
          const uint32_t bop = ptx::smem_u32(s_bop);
          for (int t = 0; t < NT; ++t)
            for (int ks = 0; ks < nks; ++ks)
              ptx::umma_bf16_ts(tmem + NT * CU + t * 16, tmem + t * CU + ks * 8, umma_smem_desc(lb, bop, ks * 16), idesc, ks != 0);
          ptx::umma_commit(ctx_bar);
        }
        __syncwarp();
      }
      // meanwhile: the h half of the character distribution, W_cd[:, :Hs] . h  (16 lanes per output)
      {
        const int part = tid & 15;
        const float2* xv = reinterpret_cast<const float2*>(s_h);
        for (int v = tid >> 4; v < ((V + 1) & ~1); v += DEC_THREADS / 16) {
          float a0 = 0.f, a1 = 0.f;
          if (v < V) {
            const __nv_bfloat162* wr = reinterpret_cast<const __nv_bfloat162*>(s_wcd + (size_t)v * WCS);
            int kp = part;
            for (; kp + 16 < Hs / 2; kp += 32) {
              const float2 w0 = __bfloat1622float2(wr[kp]), w1 = __bfloat1622float2(wr[kp + 16]);
              const float2 x0 = xv[kp], x1 = xv[kp + 16];
              a0 = fmaf(w0.x, x0.x, a0); a0 = fmaf(w0.y, x0.y, a0);
              a1 = fmaf(w1.x, x1.x, a1); a1 = fmaf(w1.y, x1.y, a1);
            }
            for (; kp < Hs / 2; kp += 16) {
              const float2 w0 = __bfloat1622float2(wr[kp]);
              const float2 x0 = xv[kp];
              a0 = fmaf(w0.x, x0.x, a0); a0 = fmaf(w0.y, x0.y, a0);
            }
          }
          float acc = a0 + a1;
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          acc += __shfl_xor_sync(0xffffffffu, acc, 2);
          acc += __shfl_xor_sync(0xffffffffu, acc, 4);
          acc += __shfl_xor_sync(0xffffffffu, acc, 8);
          if (part == 0 && v < V) s_lh[v] = acc + s_bcd[v];
        }
      }
      ptx::mbar_wait(ctx_bar, (uint32_t)(s & 1));
      ptx::tc_fence_after();
      {
        const int qd = warp & 3;
        for (int t = warp >> 2; t < NT; t += NWARP / 4) {
          const uint32_t r = ptx::tmem_ld_32x32b_x1(tmem + ((uint32_t)(qd * 32) << 16) + NT * CU + t * 16);
          ptx::tmem_ld_wait();
          s_ctx[t * 128 + qd * 32 + lane] = __uint_as_float(r);
        }
      }
      ptx::tc_fence_before();
      __syncthreads();
    } else {
      // context[e] = sum_u score[u] * enc[b,u,e]  (:293-297): 16-byte bf16 loads, nrg row groups in parallel, 8 loads in flight
      {
        const int rg = tid / ncg, cg = tid % ncg;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        if (rg < nrg) {
          const uint4* col = reinterpret_cast<const uint4*>(encb + cg * 8);
          const int rstride = E / 8;  // uint4 per row
          for (int u = rg; u < ulen; u += 8 * nrg) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int uu = u + j * nrg;
              v[j] = (uu < ulen) ? __ldg(col + (size_t)uu * rstride) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int uu = u + j * nrg;
              const float a = (uu < ulen) ? s_score[uu] : 0.f;
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v[j]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = __bfloat1622float2(h2[i]);
                acc[2 * i] = fmaf(a, f.x, acc[2 * i]);
                acc[2 * i + 1] = fmaf(a, f.y, acc[2 * i + 1]);
              }
            }
          }
          float4* dst = reinterpret_cast<float4*>(s_part + (size_t)rg * E + cg * 8);
          dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
      __syncthreads();
      for (int e = tid; e < E; e += DEC_THREADS) {
        float acc = 0.f;
        for (int rg = 0; rg < nrg; ++rg) acc += s_part[(size_t)rg * E + e];
        s_ctx[e] = acc;
      }
      __syncthreads();
    }
    if (tid == 0 && b == 0) DEC_TRACE(2, 4);

    // logits = W_cd . [h || context] + b_cd ; log_softmax  (:181-182)
    {  // 16 lanes per output, interleaved 4-byte columns.  Tensor-memory path: only the context half is left to add.
      const int part = tid & 15;
      const int k_lo = p.ctx_tmem ? Hs / 2 : 0;
      const float2* xv = reinterpret_cast<const float2*>(s_h);  // s_h and s_ctx are contiguous: [h || context]
      for (int v = tid >> 4; v < ((V + 1) & ~1); v += DEC_THREADS / 16) {
        float a0 = 0.f, a1 = 0.f;
        if (v < V) {
          const __nv_bfloat162* wr = reinterpret_cast<const __nv_bfloat162*>(s_wcd + (size_t)v * WCS);
          int kp = k_lo + part;
          for (; kp + 16 < KC / 2; kp += 32) {
            const float2 w0 = __bfloat1622float2(wr[kp]), w1 = __bfloat1622float2(wr[kp + 16]);
            const float2 x0 = xv[kp], x1 = xv[kp + 16];
            a0 = fmaf(w0.x, x0.x, a0); a0 = fmaf(w0.y, x0.y, a0);
            a1 = fmaf(w1.x, x1.x, a1); a1 = fmaf(w1.y, x1.y, a1);
          }
          for (; kp < KC / 2; kp += 16) {
            const float2 w0 = __bfloat1622float2(wr[kp]);
            const float2 x0 = xv[kp];
            a0 = fmaf(w0.x, x0.x, a0); a0 = fmaf(w0.y, x0.y, a0);
          }
        }
        float acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        if (part == 0 && v < V) s_logit[v] = p.ctx_tmem ? s_lh[v] + acc : acc + s_bcd[v];
      }
    }
    __syncthreads();
    // every warp computes the log-sum-exp and the argmax redundantly (shuffles only); warp 0 writes the outputs
    int best;
    {
      float lm = -INFINITY;
      for (int v = lane; v < V; v += 32) lm = fmaxf(lm, s_logit[v]);
      lm = warp_max(lm);
      float ls = 0.f;
      for (int v = lane; v < V; v += 32) ls += __expf(s_logit[v] - lm);
      ls = warp_sum(ls);
      const float lse = lm + __logf(ls);
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int v = lane; v < V; v += 32) {
        const float lp = s_logit[v] - lse;
        if (warp == 0) p.logp[((size_t)s * p.Bfull + gb) * V + v] = lp;
        if (lp > bv) { bv = lp; bi = v; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      best = bi;
      if (tid == 0 && p.tokens) p.tokens[(size_t)s * p.Bfull + gb] = bi;
      if (tid == 0 && b == 0) DEC_TRACE(2, 5);

      // next LSTM input row: [word (padded to 64) | context], bf16  (:216-227, :236)
      __nv_bfloat16* xr = p.xbuf[np] + (size_t)b * (DEC_VP + E);
      const bool last = (s == p.steps - 1);
      for (int i = tid; i < DEC_VP + E; i += DEC_THREADS) {
        float val;
        if (i < DEC_VP) {
          if (i >= V) val = 0.f;
          else if (p.gt_dense) val = p.gt_dense[((size_t)gb * p.gt_steps + s) * V + i];
          else if (p.gt_index) val = (p.gt_index[(size_t)gb * p.gt_steps + s] == i) ? 1.f : 0.f;
          else if (p.decode_mode == LAS_DECODE_RAW) val = s_logit[i] - lse;
          else val = (i == best) ? 1.f : 0.f;
          if (last && p.word_out && i < V) p.word_out[(size_t)gb * V + i] = val;
        } else {
          val = s_ctx[i - DEC_VP];
          if (last && p.ctx_out) p.ctx_out[(size_t)gb * E + (i - DEC_VP)] = val;
        }
        xr[i] = __float2bfloat16_rn(val);
      }
    }
    __syncthreads();  // all rows written (and s_logit / s_ctx no longer needed); the release below is cumulative over it
    if (tid == 0) {
      red_release_add(ctx_ctr, 1u);
      if (b == 0) DEC_TRACE(2, 6);
    }
  }
  if (p.ctx_tmem) {
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 512);
  }
}

__global__ void __launch_bounds__(DEC_THREADS, 1) speller_decode_persistent_kernel(const __grid_constant__ DecParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_lstm = p.sl * p.ncl;
  if ((int)blockIdx.x < n_lstm) lstm_role(p, smem, blockIdx.x / p.ncl, blockIdx.x % p.ncl);
  else attention_role(p, smem, blockIdx.x - n_lstm);
}

// ---- pack kernels ------------------------------------------------------------------------------------------
// LSTM layer weights -> per-CTA swizzled atoms.  K order: [h part (Hs) | input part], each padded to 64-wide atoms.
__global__ void pack_dec_w_kernel(const float* w_ih, const float* w_hh, uint8_t* img, int l, int Hs, int E, int V, int ncl) {
  const int nh = (Hs + 63) / 64;
  const int nx = (l == 0) ? (DEC_VP + E + 63) / 64 : (Hs + 63) / 64;
  const int natoms = nh + nx;
  const int Kx = (l == 0) ? V + E : Hs;  // row length of w_ih
  const size_t n = (size_t)ncl * natoms * DEC_NW * 64;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % 64);
    const int rr = (int)((i / 64) % DEC_NW);
    const int at = (int)((i / (64 * DEC_NW)) % natoms);
    const int nb = (int)(i / ((size_t)64 * DEC_NW * natoms));
    const int unit = nb * DEC_UNITS + (rr >> 2), gate = rr & 3;
    const size_t srow = (size_t)gate * Hs + unit;
    float val = 0.f;
    if (at < nh) {
      const int k = at * 64 + kk;
      if (k < Hs) val = w_hh[srow * Hs + k];
    } else {
      const int kx = (at - nh) * 64 + kk;
      if (l == 0) {
        if (kx < DEC_VP) { if (kx < V) val = w_ih[srow * Kx + kx]; }
        else if (kx - DEC_VP < E) val = w_ih[srow * Kx + V + (kx - DEC_VP)];
      } else if (kx < Hs) {
        val = w_ih[srow * Kx + kx];
      }
    }
    const size_t off = ((size_t)nb * natoms + at) * WATOM_BYTES + (size_t)(rr >> 3) * 1024 + (rr & 7) * 128 + ((((kk >> 3) ^ (rr & 7)) & 7) << 4) +
                       (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(val);
  }
}
__global__ void pack_dec_bias_kernel(const float* b_ih, const float* b_hh, float* dst, int Hs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * Hs) return;
  const int nb = i / DEC_NW, rr = i % DEC_NW;
  const int unit = nb * DEC_UNITS + (rr >> 2), gate = rr & 3;
  dst[i] = b_ih[gate * Hs + unit] + b_hh[gate * Hs + unit];
}

// initial decoder input / state in the kernel's operand formats
__global__ void dec_init_kernel(DecParams p, const float* enc_f32, const float* word_in, const float* ctx_in, const float* h_in) {
  const int b = blockIdx.x, gb = p.b0 + blockIdx.x;
  __nv_bfloat16* xr = p.xbuf[0] + (size_t)b * (DEC_VP + p.E);
  for (int i = threadIdx.x; i < DEC_VP + p.E; i += blockDim.x) {
    float v;
    if (i < DEC_VP) v = (i < p.V) ? (word_in ? word_in[(size_t)gb * p.V + i] : (i == 0 ? 1.f : 0.f)) : 0.f;  // <sos> = index 0
    else v = ctx_in ? ctx_in[(size_t)gb * p.E + (i - DEC_VP)] : enc_f32[(size_t)gb * p.U * p.E + (i - DEC_VP)];  // enc[:,0,:]
    xr[i] = __float2bfloat16_rn(v);
  }
  for (int l = 0; l < p.sl; ++l)
    for (int i = threadIdx.x; i < p.Hs; i += blockDim.x)
      p.hbuf[l][0][(size_t)b * p.Hs + i] = __float2bfloat16_rn(h_in ? h_in[((size_t)l * p.Bfull + gb) * p.Hs + i] : 0.f);
}

int g_dec_ctx_tmem = 1;  // las_debug_set_option(2, v)

struct Shape {
  int ncl, natoms[MAX_SL];
  size_t w_bytes[MAX_SL];
};
Shape shape_of(const las_speller_dims* d) {
  Shape s;
  s.ncl = d->Hs / DEC_UNITS;
  for (int l = 0; l < d->sl && l < MAX_SL; ++l) {
    const int nh = (d->Hs + 63) / 64;
    const int nx = (l == 0) ? (DEC_VP + d->E + 63) / 64 : (d->Hs + 63) / 64;
    s.natoms[l] = nh + nx;
    s.w_bytes[l] = (size_t)s.ncl * s.natoms[l] * WATOM_BYTES;
  }
  return s;
}
struct RingCfg {
  int box_rows, stage_bytes, nstages;
  size_t smem;
};
// ring geometry for a launch covering `rows` utterances: as many stages as shared memory allows (up to 12), so that a
// whole step's critical input can be in flight at once
RingCfg ring_cfg(const las_speller_dims* d, int rows) {
  const Shape s = shape_of(d);
  int mx = 0;
  for (int l = 0; l < d->sl; ++l) mx = s.natoms[l] > mx ? s.natoms[l] : mx;
  RingCfg r;
  (void)rows;
  r.box_rows = 64;
  r.stage_bytes = r.box_rows * 128;
  const size_t fixed = (size_t)mx * WATOM_BYTES + DEC_NW * 4 + (DEC_MAX_STAGES + 4) * 8 + 64;
  // one slot per atom of the larger part (own-h part: ceil(Hs/64); input part: (64 + E)/64 for layer 0)
  int need = (d->Hs + 63) / 64;
  const int nx0 = (DEC_VP + d->E + 63) / 64;
  if (nx0 > need) need = nx0;
  r.nstages = need;
  r.smem = fixed + (size_t)r.nstages * r.stage_bytes + (r.box_rows == 64 ? 0 : 0);
  return r;
}
size_t att_smem(const las_speller_dims* d, bool k_in) {
  const size_t KC = (size_t)d->Hs + d->E;
  size_t b = (size_t)d->D * (d->Hs + 16) * 2 + ((size_t)d->V * (KC + 32) + 8) * 2;
  b += 4 * ((size_t)d->Hs + d->E + att_kstride(d->D) + d->D + d->U + 2 * d->V + 32 + 4 + 4096);
  if (k_in) b += 4 * (size_t)d->U * att_kstride(d->D);
  return b + 64;
}
int supported(const las_speller_dims* d) {
  LAS_REQUIRE(d->sl <= MAX_SL, "LAS_MODE_BF16 speller supports at most %d layers (sl=%d)", MAX_SL, d->sl);
  LAS_REQUIRE(d->Hs % 16 == 0 && d->Hs <= 512, "LAS_MODE_BF16 speller needs hidden_size %% 16 == 0 and <= 512 (Hs=%d); use LAS_MODE_FP32", d->Hs);
  LAS_REQUIRE(d->V <= DEC_VP, "LAS_MODE_BF16 speller supports vocabularies up to %d (V=%d)", DEC_VP, d->V);
  LAS_REQUIRE(d->E % 8 == 0 && d->E / 8 <= DEC_THREADS, "LAS_MODE_BF16 speller needs E %% 8 == 0 and E <= 4096 (E=%d)", d->E);
  LAS_REQUIRE(att_smem(d, false) <= 220 * 1024 && ring_cfg(d, 64).smem <= 224 * 1024 && ring_cfg(d, 64).nstages <= DEC_MAX_STAGES,
              "LAS_MODE_BF16 speller: model does not fit shared memory (U=%d, Hs=%d)", d->U, d->Hs);
  return LAS_OK;
}

struct SpellerPackFast {
  uint8_t* w_img[MAX_SL];
  float* bias[MAX_SL];
  __nv_bfloat16* w_phi;
  __nv_bfloat16* w_cd;
  size_t bytes;
};
SpellerPackFast pack_layout(const las_speller_dims* d, void* base) {
  SpellerPackFast p;
  const Shape s = shape_of(d);
  Carver cv(base);
  for (int l = 0; l < d->sl && l < MAX_SL; ++l) {
    p.w_img[l] = cv.take<uint8_t>(s.w_bytes[l]);
    p.bias[l] = cv.take<float>(4 * (size_t)d->Hs);
  }
  p.w_phi = cv.take<__nv_bfloat16>((size_t)d->D * d->Hs);
  p.w_cd = cv.take<__nv_bfloat16>((size_t)d->V * (d->Hs + d->E) + 8);
  p.bytes = cv.total();
  return p;
}
struct SpellerWsFast {
  __nv_bfloat16* enc_bf16;
  __nv_bfloat16* hbuf[MAX_SL][2];
  float* hf32[2];
  __nv_bfloat16* xbuf[2];
  uint32_t* sync;
  size_t bytes;
};
SpellerWsFast ws_layout(const las_speller_dims* d, void* base) {
  SpellerWsFast w;
  Carver cv(base);
  w.enc_bf16 = cv.take<__nv_bfloat16>((size_t)d->B * d->U * d->E);
  for (int l = 0; l < MAX_SL; ++l)
    for (int k = 0; k < 2; ++k) w.hbuf[l][k] = cv.take<__nv_bfloat16>((size_t)d->B * d->Hs);
  for (int k = 0; k < 2; ++k) w.hf32[k] = cv.take<float>((size_t)d->B * d->Hs);
  for (int k = 0; k < 2; ++k) w.xbuf[k] = cv.take<__nv_bfloat16>((size_t)d->B * (DEC_VP + d->E));
  w.sync = cv.take<uint32_t>(32 * (MAX_SL + 1));
  w.bytes = cv.total();
  return w;
}

}  // namespace

bool fast_available() { return true; }
void fast_set_option_speller(int key, int value) {
  if (key == 2) g_dec_ctx_tmem = value;
}

size_t fast_speller_packed_bytes(const las_speller_dims* d) { return pack_layout(d, nullptr).bytes; }
size_t fast_speller_workspace_bytes(const las_speller_dims* d, int) { return ws_layout(d, nullptr).bytes; }

int fast_speller_pack(const las_speller_weights* w, const las_speller_dims* d, void* packed_fast, cudaStream_t st) {
  LAS_TRY(supported(d));
  const SpellerPackFast pk = pack_layout(d, packed_fast);
  const Shape s = shape_of(d);
  for (int l = 0; l < d->sl; ++l) {
    const las_lstm_weights& lw = w->rnn_host[l];
    pack_dec_w_kernel<<<592, 256, 0, st>>>(lw.w_ih, lw.w_hh, pk.w_img[l], l, d->Hs, d->E, d->V, s.ncl);
    LAS_LAUNCH_OK("pack_dec_w_kernel");
    pack_dec_bias_kernel<<<(4 * d->Hs + 255) / 256, 256, 0, st>>>(lw.b_ih, lw.b_hh, pk.bias[l], d->Hs);
    LAS_LAUNCH_OK("pack_dec_bias_kernel");
  }
  LAS_TRY(launch_f32_to_bf16(w->w_phi, pk.w_phi, (size_t)d->D * d->Hs, st));
  LAS_TRY(launch_f32_to_bf16(w->w_cd, pk.w_cd, (size_t)d->V * (d->Hs + d->E), st));
  return LAS_OK;
}

int fast_speller_decode(const las_decode_io* io, const void* packed_f32, const void* packed_fast, const las_speller_dims* d, int steps,
                        int decode_mode, int relu, void* ws_f32, void* ws_fast, cudaStream_t st) {
  LAS_TRY(supported(d));
  const Shape s = shape_of(d);
  const int n_lstm = d->sl * s.ncl;
  const int nsm = sm_count();
  LAS_REQUIRE(n_lstm + 1 <= nsm, "LAS_MODE_BF16 speller: %d LSTM CTAs do not fit %d SMs", n_lstm, nsm);
  // utterances per persistent launch: one attention CTA each, and at most 64 (activation slots hold 64 batch rows)
  const int max_b = (nsm - n_lstm) < 64 ? (nsm - n_lstm) : 64;
  const SpellerPackFast pk = pack_layout(d, const_cast<void*>(packed_fast));
  // fp32 block of the pack (las_api.cu layout): psi / phi / cd weights and biases in the reference's own shapes
  struct F32View { const float *w_psi, *b_psi, *b_phi, *b_cd; } fv;
  {
    Carver cv(const_cast<void*>(packed_f32));
    const size_t G = 4 * (size_t)d->Hs;
    for (int l = 0; l < d->sl; ++l) {
      const size_t Kx = (l == 0) ? (size_t)d->V + d->E : (size_t)d->Hs;
      cv.take<float>(G * Kx); cv.take<float>(G * d->Hs); cv.take<float>(G); cv.take<float>(G);
    }
    cv.take<float>((size_t)d->D * d->Hs);
    fv.b_phi = cv.take<float>(d->D);
    fv.w_psi = cv.take<float>((size_t)d->D * d->E);
    fv.b_psi = cv.take<float>(d->D);
    cv.take<float>((size_t)d->V * (d->Hs + d->E));
    fv.b_cd = cv.take<float>(d->V);
  }
  float* psi_ws = static_cast<float*>(ws_f32);  // first buffer of the fp32 workspace layout: psi [B,U,D]
  const float* psi = io->psi;
  if (!psi) {
    ProfScope ps("speller.psi", st);
    LAS_TRY(launch_sgemm_nt_bias(io->enc, d->E, fv.w_psi, d->E, fv.b_psi, psi_ws, d->D, d->B * d->U, d->D, d->E, relu != 0, st));
    psi = psi_ws;
  }

  for (int b0 = 0; b0 < d->B; b0 += max_b) {
    const int Bc = (d->B - b0) < max_b ? (d->B - b0) : max_b;
    las_speller_dims dc = *d;
    dc.B = Bc;
    const SpellerWsFast w = ws_layout(&dc, ws_fast);
    DecParams p;
    memset(&p, 0, sizeof(p));
    p.B = Bc; p.U = d->U; p.E = d->E; p.Hs = d->Hs; p.sl = d->sl; p.V = d->V; p.D = d->D;
    p.steps = steps; p.decode_mode = decode_mode; p.relu = relu; p.gt_steps = io->gt_steps; p.ncl = s.ncl;
    p.k_in_smem = att_smem(d, true) <= 220 * 1024;
    {
      const int nks = (d->U + 15) / 16, NT = d->E / 128;
      p.ctx_tmem = (d->E % 128 == 0) && (NT * (nks * 8 + 16) <= 512) && (nks * 512 + 16 <= 4096 * 4) && g_dec_ctx_tmem;
    }
    const RingCfg rc = ring_cfg(d, Bc);
    p.nstages = rc.nstages;
    p.stage_bytes = rc.stage_bytes;
    for (int l = 0; l < d->sl; ++l) {
      p.w_img[l] = pk.w_img[l];
      p.bias[l] = pk.bias[l];
      for (int k = 0; k < 2; ++k) {
        p.hbuf[l][k] = w.hbuf[l][k];
        LAS_TRY(make_tmap_bf16_box(&p.tm_h[l][k], w.hbuf[l][k], Bc, d->Hs, d->Hs, rc.box_rows));
      }
    }
    for (int k = 0; k < 2; ++k) {
      p.hf32[k] = w.hf32[k];
      p.xbuf[k] = w.xbuf[k];
      LAS_TRY(make_tmap_bf16_box(&p.tm_x[k], w.xbuf[k], Bc, DEC_VP + d->E, DEC_VP + d->E, rc.box_rows));
    }
    const size_t so = (size_t)b0;  // batch offset into caller tensors
    p.Bfull = d->B;
    p.b0 = b0;
    p.c_init = io->c_state;
    p.h_out = io->h_state;
    p.c_out = io->c_state;
    p.enc = w.enc_bf16;
    p.psi = psi;
    p.w_phi = pk.w_phi; p.b_phi = fv.b_phi; p.w_cd = pk.w_cd; p.b_cd = fv.b_cd;
    p.gt_dense = io->gt_dense;
    p.gt_index = io->gt_index;
    p.enc_lengths = io->enc_lengths;
    p.logp = io->logp; p.attn = io->attn; p.tokens = io->tokens;
    p.word_out = io->word; p.ctx_out = io->context;
    p.sync = w.sync;
    p.trace = fast_get_trace() ? fast_get_trace() + 512 : nullptr;

    {
      ProfScope ps("speller.prepare", st);
      LAS_TRY(launch_f32_to_bf16(io->enc + so * d->U * d->E, w.enc_bf16, (size_t)Bc * d->U * d->E, st));
      LAS_CUDA_OK(cudaMemsetAsync(w.sync, 0, sizeof(uint32_t) * 32 * (MAX_SL + 1), st));
      dec_init_kernel<<<Bc, 256, 0, st>>>(p, io->enc, io->word, io->context, io->h_state);
      LAS_LAUNCH_OK("dec_init_kernel");
    }
    ProfScope ps("speller.steps", st);
    const size_t smem_l = rc.smem, smem_a = att_smem(d, p.k_in_smem != 0);
    const size_t smem = (smem_l > smem_a ? smem_l : smem_a) + 1024;
    LAS_CUDA_OK(cudaFuncSetAttribute(speller_decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_lstm + Bc);
    cfg.blockDim = dim3(DEC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    LAS_CUDA_OK(cudaLaunchKernelEx(&cfg, speller_decode_persistent_kernel, p));
    count_launch();
  }
  return LAS_OK;
}

}  // namespace las
