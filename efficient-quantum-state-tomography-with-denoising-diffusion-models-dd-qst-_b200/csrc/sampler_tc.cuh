// Persistent tcgen05 reverse sampler (DDQST_PRECISION_BF16): declarations shared with api.cu.
#pragma once
#include "common.cuh"

namespace ddqst {

int64_t sampler_tc_workspace_bytes(const ddqst_dims* d, int64_t batch);
int sampler_tc_supported(const ddqst_dims* d);   // DDQST_OK or DDQST_EUNSUPPORTED (+ message)

int sampler_tc_sample(const ddqst_dims* d, const char* pack, const PackLayout& pl, const float* sched, int mode,
                      const int32_t* basis_ids, int32_t n_bases, int64_t spb, int64_t shot_offset, uint64_t seed,
                      void* out_packed, uint32_t* out_hist, void* workspace, int64_t ws_bytes, cudaStream_t s);

int sampler_tc_step(const ddqst_dims* d, const char* pack, const PackLayout& pl, const float* sched, int mode,
                    int32_t basis_id, int32_t t, int64_t shots, int64_t shot_offset, uint64_t seed,
                    const uint16_t* x_t, uint16_t* x_prev, float* logits_out, void* workspace, int64_t ws_bytes,
                    cudaStream_t s);

int sampler_tc_forward(const ddqst_dims* d, const char* pack, const PackLayout& pl, const uint16_t* x,
                       const int32_t* t, const int32_t* basis, int64_t batch, float* logits, void* workspace,
                       int64_t ws_bytes, cudaStream_t s);

// training step (train.cu: fp32 CUDA cores; train_tc.cu: bf16 tcgen05)
int64_t train_workspace_bytes(const ddqst_dims* d, int64_t batch);
int64_t train_tc_workspace_bytes(const ddqst_dims* d, int64_t batch);
int train_tc_abort_fetch();
int recon_tc_abort_fetch();   // recon.cu (ring Jacobi inbox barriers)

// shared by the fp32 reverse-step kernel and the tcgen05 epilogue so both draw identically
// logit(q, c) supplies logits[q][c]; returns the packed x_{t-1}
template <typename LogitFn>
__device__ __forceinline__ uint32_t reverse_step_bits(int N, int T, const float* __restrict__ sched, int mode, int t,
                                                      uint64_t seed, uint32_t basis, uint64_t shot, uint32_t xt,
                                                      LogitFn logit) {
  const float* betas = sched;
  const float* Qp = sched + (T + 1) + (int64_t)(t - 1) * 4;
  const float beta = betas[t], one_m = 1.0f - beta;
  uint32_t out = 0;
  Philox4 p{}, p2{};
  for (int q = 0; q < N; ++q) {
    float l0 = logit(q, 0), l1 = logit(q, 1);
    float m = fmaxf(l0, l1), e0 = expf(l0 - m), e1 = expf(l1 - m), s = e0 + e1;
    float p0 = __fdiv_rn(e0, s), p1 = __fdiv_rn(e1, s);
    if (mode == DDQST_MODE_POSTERIOR) {
      if ((q & 3) == 0) p = stream_block(seed, basis, t, DDQST_SITE_POSTERIOR, shot, q >> 2);
      uint32_t bit = (xt >> q) & 1u;
      float tr0 = bit ? beta : one_m, tr1 = bit ? one_m : beta;
      // torch.matmul([p0,p1], Q_bar[t-1]) on CPU rounds as fma(p1, Q[1][k], p0*Q[0][k]) (measured)
      float pr0 = __fmaf_rn(p1, Qp[2], __fmul_rn(p0, Qp[0]));
      float pr1 = __fmaf_rn(p1, Qp[3], __fmul_rn(p0, Qp[1]));
      float w0 = __fmul_rn(tr0, pr0), w1 = __fmul_rn(tr1, pr1);
      float den = __fadd_rn(__fadd_rn(w0, w1), 1e-8f);
      out |= draw_bit(word_to_uniform(lane_of(p, q)), __fdiv_rn(w0, den), __fdiv_rn(w1, den)) << q;
    } else {
      if ((q & 3) == 0) {
        p = stream_block(seed, basis, t, DDQST_SITE_X0HAT, shot, q >> 2);
        if (t > 1) p2 = stream_block(seed, basis, t, DDQST_SITE_RENOISE, shot, q >> 2);
      }
      uint32_t x0 = draw_bit(word_to_uniform(lane_of(p, q)), p0, p1);
      if (t > 1) x0 = draw_bit(word_to_uniform(lane_of(p2, q)), Qp[0 * 2 + x0], Qp[1 * 2 + x0]);   // Q[t-1][to, from]
      out |= x0 << q;
    }
  }
  return out;
}

}  // namespace ddqst
