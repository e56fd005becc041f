// Shared helpers for libddqst: status/error plumbing, Philox4x32-10, packed-state layout.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ddqst.h"

namespace ddqst {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int check_arch();   // DDQST_OK or DDQST_EUNSUPPORTED_ARCH
int num_sms();

#define DDQST_CUDA_OK(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ddqst::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DDQST_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define DDQST_LAUNCH_OK()                                                                     \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ddqst::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DDQST_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define DDQST_REQUIRE(cond, code, ...)                                                        \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ddqst::set_error(__VA_ARGS__);                                                          \
      return (code);                                                                          \
    }                                                                                         \
  } while (0)

#define DDQST_TRY(expr)                                                                       \
  do {                                                                                        \
    int _s = (expr);                                                                          \
    if (_s != DDQST_OK) return _s;                                                            \
  } while (0)

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- Philox4x32-10
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// words of the injected stream for qubits [4*blk, 4*blk+3] of one sample (see include/ddqst.h)
__host__ __device__ __forceinline__ Philox4 stream_block(uint64_t seed, uint32_t stream, uint32_t t, uint32_t site,
                                                          uint64_t sample, uint32_t blk) {
  return philox4x32_10((uint32_t)sample, stream, (t & 0xFFFFu) | (site << 16),
                       blk | ((uint32_t)((sample >> 32) & 0xFFFFFFu) << 8), (uint32_t)seed, (uint32_t)(seed >> 32));
}

__host__ __device__ __forceinline__ uint32_t lane_of(const Philox4& p, int q) {
  int l = q & 3;
  return l == 0 ? p.x : (l == 1 ? p.y : (l == 2 ? p.z : p.w));
}

__host__ __device__ __forceinline__ float word_to_uniform(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }

// draw bit = u*(p0+p1) < p1 with separately rounded fp32 ops (matches oracle.draw_bits)
__device__ __forceinline__ uint32_t draw_bit(float u, float p0, float p1) {
  return __fmul_rn(u, __fadd_rn(p0, p1)) < p1 ? 1u : 0u;
}

// ---------------------------------------------------------------- flat parameter layout
struct ParamLayout {
  int64_t x_emb, in_w, in_b, time_emb, basis_emb;
  int64_t film_w[16], film_b[16], w1[16], b1[16], w2[16], b2[16];
  int64_t head_w, head_b, total;
};
int param_layout(const ddqst_dims* d, ParamLayout* out);   // status

// ---------------------------------------------------------------- packed inference state
// byte offsets inside the pack buffer
struct PackLayout {
  int64_t c0;        // fp32 [H]
  int64_t D;         // fp32 [N,H]
  int64_t Tt;        // fp32 [T+1, L, 2H]
  int64_t Tb;        // fp32 [num_bases, L, 2H]   (film bias folded in)
  int64_t bias1;     // fp32 [L,H]
  int64_t bias2;     // fp32 [L,H]
  int64_t head_b;    // fp32 [16]  (2N padded)
  int64_t w1_f32;    // fp32 [L,H,H]   (copies for the exact path, K-major as in nn.Linear)
  int64_t w2_f32;    // fp32 [L,H,H]
  int64_t head_f32;  // fp32 [NH_PAD,H]
  int64_t w_bf16;    // bf16 [L,2,H,H]
  int64_t head_bf16; // bf16 [NH_PAD,H]
  int64_t dt_bf16;   // bf16 [H,64]: input collapse as an MMA operand, hi/lo split (see sampler_tc.cu)
  int64_t total;
  int head_pad;      // rows of the padded head (multiple of 16, >= 2N)
};
int pack_layout(const ddqst_dims* d, PackLayout* out);     // status

int validate_dims(const ddqst_dims* d);

}  // namespace ddqst
