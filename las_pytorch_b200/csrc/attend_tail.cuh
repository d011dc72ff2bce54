// Shared by the two attention-step kernels (kernels_f32.cu attend_f32_kernel, gen_step.cu attend_cluster_kernel): block reductions and
// the tail of a decoder step for one utterance -- log-softmax, NLL term, argmax / draw, the fed-back word
// (model/las_model.py:181-182, :216-236).
#pragma once
#include "las_kernels.cuh"

namespace las {

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += red[i];  // fixed order -> deterministic
  return r;
}

// s_logit[V] holds the raw logits of utterance b (already visible to the whole CTA); s_red is a 32-float scratch.
__device__ __forceinline__ void attend_tail(const AttendArgs& a, int b, float* s_logit, float* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (a.V <= 128) {  // character-level vocabularies: one warp, no block barriers
    if (wid == 0) {
      float lm = -INFINITY;
      for (int v = lane; v < a.V; v += 32) lm = fmaxf(lm, s_logit[v]);
      lm = warp_max(lm);
      float ls = 0.f;
      for (int v = lane; v < a.V; v += 32) ls += expf(s_logit[v] - lm);
      ls = warp_sum(ls);
      const float lse = lm + logf(ls);
      for (int v = lane; v < a.V; v += 32) {
        const float lp = s_logit[v] - lse;
        s_logit[v] = lp;
        a.logp_out[(size_t)b * a.V + v] = lp;
      }
    }
  } else {
    float lm = -INFINITY;
    for (int v = tid; v < a.V; v += blockDim.x) lm = fmaxf(lm, s_logit[v]);
    lm = block_reduce_max(lm, s_red);
    float ls = 0.f;
    for (int v = tid; v < a.V; v += blockDim.x) ls += expf(s_logit[v] - lm);
    ls = block_reduce_sum(ls, s_red);
    const float lse = lm + logf(ls);
    for (int v = tid; v < a.V; v += blockDim.x) {
      const float lp = s_logit[v] - lse;
      s_logit[v] = lp;
      a.logp_out[(size_t)b * a.V + v] = lp;
    }
  }
  __syncthreads();
  if (a.nll_term_out && tid == 0) {  // NLLLoss(ignore_index=0) term of this (step, utterance)
    const int lab = a.nll_label_step ? a.nll_label_step[(size_t)b * a.nll_label_ld] : 0;
    a.nll_term_out[b] = (lab > 0 && lab < a.V) ? -s_logit[lab] : 0.f;
  }

  // argmax (lowest index wins ties, as torch.topk / argmax do on a row)   (:225)
  int best = 0;
  if (wid == 0) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < a.V; v += 32) {
      const float x = s_logit[v];
      if (x > bv) { bv = x; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    best = bi;
    if (lane == 0) {
      // decode_mode 2 (:229-234): the word fed back (and reported) is a draw from Categorical(probs = log-probs)
      if (a.decode_mode == LAS_DECODE_SAMPLE && !a.gt_dense_step && !a.gt_index_step)
        best = las_sample_logp_as_probs(s_logit, a.V, las_uniform(a.sample_seed, (uint32_t)a.step, (uint32_t)b));
      s_red[0] = __int_as_float(best);
      if (a.token_out) a.token_out[b] = best;
    }
  }
  __syncthreads();
  best = __float_as_int(s_red[0]);

  // next input word   (:216-227); op_out (nullable) receives the same row in the 16-bit GEMM operand format
  if (a.word_out) {
    float* wo = a.word_out + (size_t)b * a.word_ld;
    __nv_bfloat16* oo = a.op_out ? a.op_out + (size_t)b * a.op_ld : nullptr;
    const float* g = a.gt_dense_step ? a.gt_dense_step + (long long)b * a.gt_ld : nullptr;
    const bool by_index = !g && a.gt_index_step;
    const int gi = by_index ? a.gt_index_step[(size_t)b * a.gt_index_ld] : -1;
    for (int v = tid; v < a.V; v += blockDim.x) {
      float w;
      if (g) w = g[v];
      else if (by_index) w = (v == gi) ? 1.f : 0.f;
      else if (a.decode_mode == LAS_DECODE_RAW) w = s_logit[v];
      else w = (v == best) ? 1.f : 0.f;
      wo[v] = w;
      if (oo) oo[v] = op_from_f32(w, a.op_f16);
    }
  }
}

}  // namespace las
