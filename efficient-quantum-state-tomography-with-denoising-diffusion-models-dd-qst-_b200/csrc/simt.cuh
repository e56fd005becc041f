// CUDA-core fp32 building blocks: strided SGEMM with fused epilogues, FiLM / input-collapse kernels,
// reverse-step draw kernels.  They implement the "exact" (DDQST_PRECISION_FP32) path and the training
// step, and they prepare the packed tables the tcgen05 sampler consumes.
#pragma once
#include "common.cuh"

namespace ddqst {

enum Epilogue {
  EPI_NONE = 0,       // C = acc
  EPI_BIAS = 1,       // C = acc + bias[j]
  EPI_BIAS_SILU = 2,  // C = silu(acc + bias[j])
  EPI_RES_SILU = 3,   // C = silu(R[i,j] + acc + bias[j])          (ResBlock tail, RQC/model.py:24)
  EPI_BIAS_RELU = 4,  // C = relu(acc + bias[j])                   (NB c6:78-84)
  EPI_ACCUM = 5,      // C += acc
  EPI_BIAS_PRE = 6    // C = acc + bias[j], pre[i,j] = same (keeps the pre-activation for backward)
};

struct GemmArgs {
  const float* A; int64_t a_rs, a_cs;   // A(i,k) = A[i*a_rs + k*a_cs]
  const float* B; int64_t b_rs, b_cs;   // B(k,j) = B[k*b_rs + j*b_cs]
  float* C; int64_t ldc;                // C(i,j) = C[i*ldc + j]
  const float* bias;                    // [N] or null
  const float* R; int64_t ldr;          // residual (EPI_RES_SILU)
  float* aux; int64_t ldaux;            // optional second output: pre-activation (acc + bias [+R])
  int M, N, K;
  int epi;
  float alpha;                          // acc scaled by alpha before the epilogue
};

int launch_sgemm(const GemmArgs& g, cudaStream_t s);

// who supplies (t, basis, shot) for row i of a chunk
struct RowCtx {
  const int32_t* t_arr;       // per-row timesteps or null
  int32_t t_uniform;
  const int32_t* basis_arr;   // per-row basis ids or null
  const int32_t* basis_ids;   // per-slot basis ids (sample()): slot = (row0+i)/spb
  int64_t spb;                // shots per basis
  int64_t row0;               // global row of chunk row 0
  int64_t shot_offset;
};

__device__ __forceinline__ int32_t row_t(const RowCtx& c, int64_t i) { return c.t_arr ? c.t_arr[i] : c.t_uniform; }
__device__ __forceinline__ int32_t row_basis(const RowCtx& c, int64_t i) {
  if (c.basis_arr) return c.basis_arr[i];
  return c.basis_ids[(c.row0 + i) / c.spb];
}
__device__ __forceinline__ uint64_t row_shot(const RowCtx& c, int64_t i) {
  return (uint64_t)(c.shot_offset + (c.spb > 0 ? (c.row0 + i) % c.spb : (c.row0 + i)));
}

// h = c0 + sum_q bit_q D[q,:]  and  a = h*(1+gamma)+beta for block 0
int launch_input_film(const ddqst_dims* d, const char* pack, const PackLayout& pl, const uint16_t* x, RowCtx ctx,
                      int64_t rows, float* h, float* a, cudaStream_t s);
// a = h*(1+gamma)+beta for block `blk`
int launch_film(const ddqst_dims* d, const char* pack, const PackLayout& pl, int blk, RowCtx ctx, int64_t rows,
                const float* h, float* a, cudaStream_t s);

// reverse step on logits[rows, N, 2] (posterior: RQC/diffusion.py:62-79; renoise: SS/diffusion.py:67-80)
int launch_reverse_step(const ddqst_dims* d, const float* sched, int mode, int t, RowCtx ctx, int64_t rows,
                        uint64_t seed, const float* logits, const uint16_t* x_t, uint16_t* x_prev, cudaStream_t s);
int launch_init_bits(int num_qubits, RowCtx ctx, int64_t rows, uint64_t seed, uint16_t* x, cudaStream_t s);

// full denoiser forward on a chunk through the fp32 path; ws holds 3*rows*H floats
int forward_fp32(const ddqst_dims* d, const char* pack, const PackLayout& pl, const uint16_t* x, RowCtx ctx,
                 int64_t rows, float* logits, float* ws, cudaStream_t s);

}  // namespace ddqst
