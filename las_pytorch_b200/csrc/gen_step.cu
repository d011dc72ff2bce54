// Generic tensor-core decoder step, fused (LAS_MODE_BF16 / LAS_MODE_F16 beyond what the persistent decoder keeps on chip: the
// reference's shipped config/librispeech-config.yaml:13-34 with its 1024-wide speller cells, GRU / RNN cells, use_mlp_in_attention
// False).  One decoder step (model/las_model.py:178-238) is `layers` launches of gen_cell_step_kernel + one of attend_cluster_kernel:
//
//   gen_cell_step_kernel   one stacked-cell layer: pre[r, b] = W[r, :] . [x_b | h_prev_b]  (tcgen05, weights streamed from L2 through a
//                          TMA ring, fp32 accumulation in tensor memory) + bias -> cell update in fp32 -> h as fp32 (attention, state
//                          output) AND as 16-bit operand rows of the GEMMs that consume it next (this layer's next step, the next
//                          layer's current step): no pre-activation round trip, no operand-building kernel.
//   attend_cluster_kernel  phi, energies, masked softmax, context, character distribution, feedback for one utterance by a cluster of
//                          up to 8 CTAs: the encoder steps are cut into 8 fixed slices (flash-decoding style partial softmax / partial
//                          context per slice, combined in slice order through distributed shared memory), phi's and the character
//                          MLP's rows are dealt to the CTAs.  The slices, and every summation order, are the same for every cluster
//                          size, so an utterance's result does not depend on the batch it is decoded in.
//
// Launches are chained with programmatic dependent launch: a kernel's prologue, and for the cell kernel the K blocks that multiply the
// layer's OWN previous state (written a whole step earlier), run under the tail of the kernel in front of it.
#include <cuda.h>
#include <string.h>

#include "attend_tail.cuh"
#include "las_fast.cuh"
#include "umma.cuh"

namespace las {

namespace {

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// =========================================================================================================
// gen_cell_step_kernel.  CTA c owns 64 / G consecutive cells and all G gate rows of each (LSTM: i,f,g,o; GRU: r,z,n_x,n_h; RNN: 1),
// i.e. 64 rows of the packed [R, Kp] weights (kernels_f32.cu gen_pack_w_kernel), fetched as G boxes of 64/G rows per 64-column K
// block.  The weights are the A operand (M = 128 in the instruction; rows 64-127 of the tile are whatever follows in shared
// memory and land in tensor-memory lanes nobody reads -- M = 128 costs the same and keeps the lane = row layout), the activations
// [B, Kp] the B operand (N = B rounded up to 16, rows beyond B are TMA zero fill).
//   warp 0 (one thread)  TMA producer, ring of `stages` x (8 KB weights + N x 128 B activations)
//   warp 1 (one thread)  tcgen05.mma issuer; tensor-memory allocation
//   all warps            epilogue: lanes 0-63 -> shared memory (warps 4, 5), then one (utterance, cell) per thread
// K blocks [kb_first, k_blocks) multiply the layer's own previous state and are issued first, before griddepcontrol.wait.
// =========================================================================================================
constexpr int GS_ROWS = 64, GS_THREADS = 256, GS_WBYTES = GS_ROWS * 128;

struct GenStepArgs {
  const float* bias;     // [R]
  float* c;              // [B, H] (LSTM)
  const float* h_prev;   // [B, H] row stride h_ld (GRU)
  long long h_ld;
  float* h_out;          // [B, H] row stride hout_ld
  long long hout_ld;
  __nv_bfloat16* o1;     // nullable: operand rows that receive h (this layer's next step), already offset to the h columns
  long long o1_ld;
  __nv_bfloat16* o2;     // nullable: the next layer's x columns
  long long o2_ld;
  int B, H, G, cell, Kp, NB, stages, f16, kb_first;
};

__global__ void __launch_bounds__(GS_THREADS, 1)
gen_cell_step_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_a, const GenStepArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const ring = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = GS_WBYTES + p.NB * 128;
  uint64_t* const full = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * stage_bytes + GS_WBYTES);  // + slack for the M=128 view
  uint64_t* const empty = full + p.stages;
  uint64_t* const done = empty + p.stages;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cells = GS_ROWS / p.G, cell0 = blockIdx.x * cells;
  const int k_blocks = (p.Kp + 63) / 64;
  uint32_t ncols = 32;
  while ((int)ncols < p.NB) ncols <<= 1;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tm_w);
    ptx::prefetch_tensormap(&tm_a);
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, ncols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int n_indep = k_blocks - p.kb_first;
      for (int i = 0; i < k_blocks; ++i) {
        if (i == n_indep) {  // everything from here on reads what the previous launch wrote
          griddep_wait();
          griddep_launch();
        }
        const int kb = i < n_indep ? p.kb_first + i : i - n_indep;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* const dst = ring + (size_t)stage * stage_bytes;
        ptx::mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
        for (int g = 0; g < p.G; ++g) ptx::tma_load_2d(dst + g * cells * 128, &tm_w, &full[stage], kb * 64, g * p.H + cell0);
        ptx::tma_load_2d(dst + GS_WBYTES, &tm_a, &full[stage], kb * 64, 0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (n_indep >= k_blocks) {
        griddep_wait();
        griddep_launch();
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const UmmaLayout lw{1, 0, 1024, 0};
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.NB, p.f16);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < k_blocks; ++i) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t w_addr = ptx::smem_u32(ring + (size_t)stage * stage_bytes), a_addr = w_addr + GS_WBYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16(tmem, umma_smem_desc(lw, w_addr, k * 16), umma_smem_desc(lw, a_addr, k * 16), idesc, (i | k) != 0);
        ptx::umma_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(done);
    }
    __syncwarp();
  }
  ptx::mbar_wait(done, 0);
  ptx::tc_fence_after();
  griddep_wait();  // (already satisfied: the dependent K blocks were loaded after the producer's wait)

  // accumulator rows 0-63 -> shared memory (the ring is dead: every copy has landed and every MMA has read its operands)
  float* const pre = reinterpret_cast<float*>(ring);
  const int ldp = p.NB + 1;
  if (warp == 4 || warp == 5) {
    const int row = (warp - 4) * 32 + lane;
    for (int c0 = 0; c0 < p.NB; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld_32x32b_x16(tmem + ((uint32_t)((warp - 4) * 32) << 16) + (uint32_t)c0, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) pre[row * ldp + c0 + j] = __uint_as_float(v[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();

  const int H = p.H;
  for (int i = tid; i < p.B * cells; i += GS_THREADS) {
    const int b = i / cells, j = i % cells, cell = cell0 + j;
    if (cell >= H) continue;
    auto P = [&](int g) { return pre[(g * cells + j) * ldp + b] + p.bias[g * H + cell]; };
    float h;
    if (p.cell == LAS_CELL_LSTM) {
      const float ig = sigmoid_precise(P(0)), fg = sigmoid_precise(P(1)), gg = tanhf(P(2)), og = sigmoid_precise(P(3));
      const float cn = fg * p.c[(size_t)b * H + cell] + ig * gg;
      p.c[(size_t)b * H + cell] = cn;
      h = og * tanhf(cn);
    } else if (p.cell == LAS_CELL_GRU) {
      const float hp = p.h_prev ? p.h_prev[(long long)b * p.h_ld + cell] : 0.f;
      const float rg = sigmoid_precise(P(0)), zg = sigmoid_precise(P(1));
      const float ng = tanhf(P(2) + rg * P(3));
      h = (1.0f - zg) * ng + zg * hp;
    } else {
      h = tanhf(P(0));
    }
    p.h_out[(long long)b * p.hout_ld + cell] = h;
    const __nv_bfloat16 ho = op_from_f32(h, p.f16);
    if (p.o1) p.o1[(long long)b * p.o1_ld + cell] = ho;
    if (p.o2) p.o2[(long long)b * p.o2_ld + cell] = ho;
  }
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, ncols);
}

// =========================================================================================================
// attend_cluster_kernel (heads == 1).  Cluster of C in {1,2,4,8} CTAs per utterance, rank r:
//   q       rows [r*D/C ..) of phi, broadcast to the peers                                                  -- cluster barrier 1
//   slices  encoder steps in AC_SLICES fixed slices; rank r owns slices [r*8/C, (r+1)*8/C): energies, slice maximum m_i,
//           p = exp(e - m_i), slice sum s_i (both broadcast), partial context part_i[E] = sum_u p[u] enc[b,u,:]   -- cluster barrier 2
//   combine M = max m_i, w_i = exp(m_i - M), S = sum s_i w_i (slice order), context = (sum_i w_i part_i) / S: every CTA reads
//           all eight partial contexts (its own and its peers', through distributed shared memory); scores of the own slices
//   logits  rows v = r, r+C, .. of the character MLP as four K segments each, summed in segment order, sent to rank 0  -- barrier 3
//   rank 0  log-softmax, NLL term, token, fed-back word (attend_tail)
// =========================================================================================================
constexpr int AC_SLICES = 8, AC_THREADS = 512, AC_SEGS = 4;

__global__ void __launch_bounds__(AC_THREADS, 1) attend_cluster_kernel(const AttendArgs a, const int C) {
  extern __shared__ float sm[];
  const int nown = AC_SLICES / C;
  float* s_state = sm;                       // Hs
  float* s_q = s_state + a.Hs;               // D
  float* s_p = s_q + a.D;                    // U (own slices only)
  float* s_ms = s_p + a.U;                   // 2 * AC_SLICES: {m_i, s_i}, a full copy in every CTA
  float* s_w = s_ms + 2 * AC_SLICES;         // AC_SLICES + 1: w_i, 1/S
  float* s_ctx = s_w + AC_SLICES + 4;        // E
  float* s_logit = s_ctx + a.E;              // V (rank 0 collects)
  float* s_lpart = s_logit + a.V;            // rows per CTA * AC_SEGS
  float* s_red = s_lpart + ((a.V + C - 1) / C) * AC_SEGS;  // 32
  float* s_part = s_red + 32;                // nown * E
  const int r = (C > 1) ? (int)ptx::cluster_ctarank() : 0;
  const int b = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = AC_THREADS >> 5;

  griddep_wait();
  griddep_launch();

  for (int k = tid; k < a.Hs; k += AC_THREADS) s_state[k] = a.state[(size_t)b * a.state_ld + k];
  __syncthreads();

  // ---- q = act(W_phi . state + b_phi)   (model/las_model.py:278; :283-285 without the MLP)
  if (!a.w_phi) {
    for (int d = tid; d < a.D; d += AC_THREADS) s_q[d] = s_state[d];
  } else {
    const int dq = (a.D + C - 1) / C, d_end = min(a.D, (r + 1) * dq);
    for (int d = r * dq + wid; d < d_end; d += nw) {
      const float* wr = a.w_phi + (size_t)d * a.Hs;
      float acc = 0.f;
      for (int k = lane; k < a.Hs; k += 32) acc = fmaf(wr[k], s_state[k], acc);
      acc = warp_sum(acc) + a.b_phi[d];
      if (a.relu) acc = fmaxf(acc, 0.f);
      if (C == 1) {
        if (lane == 0) s_q[d] = acc;
      } else if (lane < C) {
        st_cluster_f32(ptx::mapa(ptx::smem_u32(&s_q[d]), (uint32_t)lane), acc);
      }
    }
  }
  if (C > 1) ptx::cluster_sync(); else __syncthreads();

  // ---- energies of the own slices   (:289-291)
  const int ulen = a.enc_lengths ? min(max(a.enc_lengths[b], 1), a.U) : a.U;
  const int us = (a.U + AC_SLICES - 1) / AC_SLICES;
  const int u_begin = min(a.U, r * nown * us), u_end = min(a.U, (r + 1) * nown * us);
  const float* psib = a.psi + (size_t)b * a.U * a.D;
  const float* encb = a.enc + (size_t)b * a.U * a.E;
  for (int u = u_begin + wid; u < u_end; u += nw) {
    const float* pr = psib + (size_t)u * a.D;
    float acc = 0.f;
    for (int d = lane; d < a.D; d += 32) acc = fmaf(s_q[d], pr[d], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_p[u] = (u < ulen) ? acc : -INFINITY;
  }
  __syncthreads();
  // slice maximum, exponentials, slice sum: one warp per own slice
  for (int li = wid; li < nown; li += nw) {
    const int i = r * nown + li;
    const int s0 = min(a.U, i * us), s1 = min(a.U, (i + 1) * us);
    float m = -INFINITY;
    for (int u = s0 + lane; u < s1; u += 32) m = fmaxf(m, s_p[u]);
    m = warp_max(m);
    float sum = 0.f;
    for (int u = s0 + lane; u < s1; u += 32) {
      const float e = (m == -INFINITY) ? 0.f : expf(s_p[u] - m);
      s_p[u] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (C == 1) {
      if (lane == 0) { s_ms[2 * i] = m; s_ms[2 * i + 1] = sum; }
    } else if (lane < C) {
      st_cluster_f32(ptx::mapa(ptx::smem_u32(&s_ms[2 * i]), (uint32_t)lane), m);
      st_cluster_f32(ptx::mapa(ptx::smem_u32(&s_ms[2 * i + 1]), (uint32_t)lane), sum);
    }
  }
  __syncthreads();
  // ---- partial contexts of the own slices: two features per thread, encoder steps ascending   (:293-297)
  for (int li = 0; li < nown; ++li) {
    const int i = r * nown + li;
    const int s0 = min(a.U, i * us), s1 = min(a.U, (i + 1) * us);
    for (int e = 2 * tid; e < a.E; e += 2 * AC_THREADS) {
      float ax = 0.f, ay = 0.f;
      int u = s0;
      if (e + 1 < a.E && (a.E & 1) == 0) {
        for (; u + 4 <= s1; u += 4) {
          float2 x[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float2*>(encb + (size_t)(u + j) * a.E + e);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ax = fmaf(s_p[u + j], x[j].x, ax);
            ay = fmaf(s_p[u + j], x[j].y, ay);
          }
        }
        for (; u < s1; ++u) {
          const float2 x = *reinterpret_cast<const float2*>(encb + (size_t)u * a.E + e);
          ax = fmaf(s_p[u], x.x, ax);
          ay = fmaf(s_p[u], x.y, ay);
        }
        s_part[li * a.E + e] = ax;
        s_part[li * a.E + e + 1] = ay;
      } else {
        for (; u < s1; ++u) {
          ax = fmaf(s_p[u], encb[(size_t)u * a.E + e], ax);
          if (e + 1 < a.E) ay = fmaf(s_p[u], encb[(size_t)u * a.E + e + 1], ay);
        }
        s_part[li * a.E + e] = ax;
        if (e + 1 < a.E) s_part[li * a.E + e + 1] = ay;
      }
    }
  }
  if (C > 1) ptx::cluster_sync(); else __syncthreads();

  // ---- combine   (softmax :292 over all slices)
  if (tid == 0) {
    float M = -INFINITY;
    for (int i = 0; i < AC_SLICES; ++i) M = fmaxf(M, s_ms[2 * i]);
    float S = 0.f;
    for (int i = 0; i < AC_SLICES; ++i) {
      const float w = (s_ms[2 * i] == -INFINITY) ? 0.f : expf(s_ms[2 * i] - M);
      s_w[i] = w;
      S = fmaf(s_ms[2 * i + 1], w, S);
    }
    s_w[AC_SLICES] = 1.0f / S;
  }
  __syncthreads();
  const float inv = s_w[AC_SLICES];
  for (int e = tid; e < a.E; e += AC_THREADS) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < AC_SLICES; ++i) {
      const uint32_t local = ptx::smem_u32(&s_part[(i % nown) * a.E + e]);
      const float x = (C == 1) ? s_part[(i % nown) * a.E + e] : ld_cluster_f32(ptx::mapa(local, (uint32_t)(i / nown)));
      acc = fmaf(s_w[i], x, acc);
    }
    acc *= inv;
    s_ctx[e] = acc;
    if (r == 0) {
      a.ctx_out[(size_t)b * a.ctx_ld + e] = acc;
      if (a.op_out) a.op_out[(size_t)b * a.op_ld + a.V + e] = op_from_f32(acc, a.op_f16);
    }
  }
  if (a.score_out) {
    for (int u = u_begin + tid; u < u_end; u += AC_THREADS) a.score_out[(size_t)b * a.U + u] = s_p[u] * s_w[u / us] * inv;
  }
  __syncthreads();
  if (!a.w_cd) {
    if (C > 1) ptx::cluster_sync();  // nobody leaves while a peer may still read its partial contexts
    return;
  }

  // ---- logits = W_cd . [state || context] + b_cd   (:181): rows r, r+C, ..; AC_SEGS K segments per row
  const int KC = a.Hs + a.E;
  const int rows_own = (a.V > r) ? (a.V - r + C - 1) / C : 0;
  const int seg_len = (KC + AC_SEGS - 1) / AC_SEGS;
  for (int item = wid; item < rows_own * AC_SEGS; item += nw) {
    const int lr = item / AC_SEGS, sg = item % AC_SEGS, v = r + lr * C;
    const float* wr = a.w_cd + (size_t)v * KC;
    const int k0 = sg * seg_len, k1 = min(KC, k0 + seg_len);
    float acc = 0.f;
    for (int k = k0 + lane; k < k1; k += 32) acc = fmaf(wr[k], k < a.Hs ? s_state[k] : s_ctx[k - a.Hs], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_lpart[item] = acc;
  }
  __syncthreads();
  for (int lr = tid; lr < rows_own; lr += AC_THREADS) {
    const int v = r + lr * C;
    float acc = a.b_cd[v];
    for (int sg = 0; sg < AC_SEGS; ++sg) acc += s_lpart[lr * AC_SEGS + sg];
    if (C == 1) s_logit[v] = acc;
    else st_cluster_f32(ptx::mapa(ptx::smem_u32(&s_logit[v]), 0u), acc);
  }
  if (C > 1) ptx::cluster_sync(); else __syncthreads();
  if (r != 0) return;
  attend_tail(a, b, s_logit, s_red);
}

int g_gen_fused = 1;    // las_debug_set_option(12, 0): the unfused generic step (GEMM + cell + operand kernels, one-CTA attention)
int g_gen_cluster = 0;  // las_debug_set_option(13, C): force the attention cluster size (0 = by batch)

}  // namespace

void fast_set_option_gen(int key, int value) {
  if (key == 12) g_gen_fused = value;
  if (key == 13) g_gen_cluster = value;
}
bool gen_step_fused(int B) { return g_gen_fused != 0 && B <= 256; }

int gen_step_make_maps(GenStepMaps* m, const __nv_bfloat16* w, const __nv_bfloat16* act0, const __nv_bfloat16* act1, int R, int G, int B, int Kp) {
  int nb = (B + 15) & ~15;
  LAS_REQUIRE(nb <= 256, "fused generic decoder step covers at most 256 utterances per launch (B=%d)", B);
  LAS_TRY(make_tmap_bf16_box(&m->w, w, R, Kp, Kp, GS_ROWS / G));
  LAS_TRY(make_tmap_bf16_box(&m->a[0], act0, B, Kp, Kp, nb));
  LAS_TRY(make_tmap_bf16_box(&m->a[1], act1, B, Kp, Kp, nb));
  return LAS_OK;
}

int launch_gen_cell_step(const GenStepMaps& m, int parity, const float* bias, float* c, const float* h_prev, long long h_ld, float* h_out,
                         long long hout_ld, __nv_bfloat16* o1, long long o1_ld, __nv_bfloat16* o2, long long o2_ld, int B, int H, int cell,
                         int Kxp, int Kp, bool pdl, cudaStream_t st) {
  GenStepArgs p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.c = c; p.h_prev = h_prev; p.h_ld = h_ld; p.h_out = h_out; p.hout_ld = hout_ld;
  p.o1 = o1; p.o1_ld = o1_ld; p.o2 = o2; p.o2_ld = o2_ld;
  p.B = B; p.H = H; p.G = (cell == LAS_CELL_RNN) ? 1 : 4; p.cell = cell; p.Kp = Kp;
  p.NB = (B + 15) & ~15;
  p.stages = p.NB <= 64 ? 8 : (p.NB <= 128 ? 6 : 4);
  p.f16 = op_f16();
  const int k_blocks = (Kp + 63) / 64;
  p.kb_first = (Kxp + 63) / 64 < k_blocks ? (Kxp + 63) / 64 : k_blocks;
  const int cells = GS_ROWS / p.G;
  const size_t smem = 1024 + (size_t)p.stages * (GS_WBYTES + p.NB * 128) + GS_WBYTES + 8 * (2 * p.stages + 1) + 16;
  LAS_CUDA_OK(cudaFuncSetAttribute(gen_cell_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((H + cells - 1) / cells));
  cfg.blockDim = dim3(GS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  LAS_CUDA_OK(cudaLaunchKernelEx(&cfg, gen_cell_step_kernel, m.w, m.a[parity & 1], p));
  LAS_LAUNCH_OK("gen_cell_step_kernel");
  return LAS_OK;
}

int launch_attend_cluster(const AttendArgs& a, bool pdl, cudaStream_t st) {
  LAS_REQUIRE(a.heads == 1, "attend_cluster_kernel is the single-head form (heads=%d)", a.heads);
  int C = g_gen_cluster;
  if (C != 1 && C != 2 && C != 4 && C != 8) {
    const int sms = sm_count();
    C = a.B * 8 <= sms ? 8 : (a.B * 4 <= sms ? 4 : (a.B * 2 <= sms ? 2 : 1));
  }
  const int nown = AC_SLICES / C;
  size_t head = (size_t)a.Hs + a.D + a.U + 2 * AC_SLICES + AC_SLICES + 4 + a.E + a.V + (size_t)((a.V + C - 1) / C) * AC_SEGS + 32;
  const size_t smem = sizeof(float) * (head + (size_t)nown * a.E);
  LAS_REQUIRE(smem <= 200 * 1024, "attention step needs %zu bytes of shared memory (U=%d E=%d)", smem, a.U, a.E);
  LAS_CUDA_OK(cudaFuncSetAttribute(attend_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(a.B * C));
  cfg.blockDim = dim3(AC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (C > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)C;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  LAS_CUDA_OK(cudaLaunchKernelEx(&cfg, attend_cluster_kernel, a, C));
  LAS_LAUNCH_OK("attend_cluster_kernel");
  return LAS_OK;
}

}  // namespace las
