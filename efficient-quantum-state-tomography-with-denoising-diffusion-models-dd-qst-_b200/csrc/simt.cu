// CUDA-core fp32 kernels (see simt.cuh).
#include "simt.cuh"
#include "sampler_tc.cuh"

namespace ddqst {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

// ------------------------------------------------------------------------------------ SGEMM
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool a_kfast = (g.a_cs == 1), b_nfast = (g.b_cs == 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256, m, k;
      if (a_kfast) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      float v = 0.f;
      if (m0 + m < g.M && k0 + k < g.K) v = g.A[(int64_t)(m0 + m) * g.a_rs + (int64_t)(k0 + k) * g.a_cs];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256, n, k;
      if (b_nfast) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      float v = 0.f;
      if (n0 + n < g.N && k0 + k < g.K) v = g.B[(int64_t)(k0 + k) * g.b_rs + (int64_t)(n0 + n) * g.b_cs];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      if (g.bias) v += g.bias[col];
      if (g.epi == EPI_RES_SILU) v += g.R[(int64_t)row * g.ldr + col];
      if (g.aux) g.aux[(int64_t)row * g.ldaux + col] = v;
      float* c = g.C + (int64_t)row * g.ldc + col;
      switch (g.epi) {
        case EPI_BIAS_SILU:
        case EPI_RES_SILU: *c = silu_f(v); break;
        case EPI_BIAS_RELU: *c = fmaxf(v, 0.f); break;
        case EPI_ACCUM: *c += v; break;
        default: *c = v;
      }
    }
  }
}

int launch_sgemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return DDQST_OK;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  sgemm_kernel<<<grid, 256, 0, s>>>(g);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

// ------------------------------------------------------------------------------------ FiLM / input
__global__ void input_film_kernel(int N, int H, int L, const float* __restrict__ c0, const float* __restrict__ D,
                                  const float* __restrict__ Tt, const float* __restrict__ Tb,
                                  const uint16_t* __restrict__ x, RowCtx ctx, int64_t rows, float* __restrict__ h,
                                  float* __restrict__ a) {
  int64_t i = blockIdx.x;
  if (i >= rows) return;
  uint32_t bits = x[i];
  int t = row_t(ctx, i), b = row_basis(ctx, i);
  const float* tt = Tt + ((int64_t)t * L) * 2 * H;
  const float* tb = Tb + ((int64_t)b * L) * 2 * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = c0[c];
    for (int q = 0; q < N; ++q)
      if ((bits >> q) & 1u) v += D[q * H + c];
    h[i * H + c] = v;
    float gamma = tt[c] + tb[c], beta = tt[H + c] + tb[H + c];
    a[i * H + c] = v * (1.0f + gamma) + beta;
  }
}

__global__ void film_kernel(int H, int L, int blk, const float* __restrict__ Tt, const float* __restrict__ Tb,
                            RowCtx ctx, int64_t rows, const float* __restrict__ h, float* __restrict__ a) {
  int64_t i = blockIdx.x;
  if (i >= rows) return;
  int t = row_t(ctx, i), b = row_basis(ctx, i);
  const float* tt = Tt + ((int64_t)t * L + blk) * 2 * H;
  const float* tb = Tb + ((int64_t)b * L + blk) * 2 * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float gamma = tt[c] + tb[c], beta = tt[H + c] + tb[H + c];
    a[i * H + c] = h[i * H + c] * (1.0f + gamma) + beta;
  }
}

int launch_input_film(const ddqst_dims* d, const char* pack, const PackLayout& pl, const uint16_t* x, RowCtx ctx,
                      int64_t rows, float* h, float* a, cudaStream_t s) {
  if (rows <= 0) return DDQST_OK;
  input_film_kernel<<<(unsigned)rows, 128, 0, s>>>(d->num_qubits, d->hidden_dim, d->num_blocks,
                                                   (const float*)(pack + pl.c0), (const float*)(pack + pl.D),
                                                   (const float*)(pack + pl.Tt), (const float*)(pack + pl.Tb), x, ctx,
                                                   rows, h, a);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int launch_film(const ddqst_dims* d, const char* pack, const PackLayout& pl, int blk, RowCtx ctx, int64_t rows,
                const float* h, float* a, cudaStream_t s) {
  if (rows <= 0) return DDQST_OK;
  film_kernel<<<(unsigned)rows, 128, 0, s>>>(d->hidden_dim, d->num_blocks, blk, (const float*)(pack + pl.Tt),
                                             (const float*)(pack + pl.Tb), ctx, rows, h, a);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int forward_fp32(const ddqst_dims* d, const char* pack, const PackLayout& pl, const uint16_t* x, RowCtx ctx,
                 int64_t rows, float* logits, float* ws, cudaStream_t s) {
  const int H = d->hidden_dim, L = d->num_blocks, N = d->num_qubits;
  float* h = ws;
  float* a = ws + rows * H;
  float* u = ws + 2 * rows * H;
  DDQST_TRY(launch_input_film(d, pack, pl, x, ctx, rows, h, a, s));
  for (int l = 0; l < L; ++l) {
    if (l > 0) DDQST_TRY(launch_film(d, pack, pl, l, ctx, rows, h, a, s));
    GemmArgs g{};
    g.A = a; g.a_rs = H; g.a_cs = 1;
    g.B = (const float*)(pack + pl.w1_f32) + (int64_t)l * H * H; g.b_rs = 1; g.b_cs = H;   // B(k,j) = W[j,k]
    g.C = u; g.ldc = H;
    g.bias = (const float*)(pack + pl.bias1) + l * H;
    g.M = (int)rows; g.N = H; g.K = H; g.epi = EPI_BIAS_SILU; g.alpha = 1.f;
    DDQST_TRY(launch_sgemm(g, s));
    GemmArgs g2{};
    g2.A = u; g2.a_rs = H; g2.a_cs = 1;
    g2.B = (const float*)(pack + pl.w2_f32) + (int64_t)l * H * H; g2.b_rs = 1; g2.b_cs = H;
    g2.C = h; g2.ldc = H; g2.R = h; g2.ldr = H;
    g2.bias = (const float*)(pack + pl.bias2) + l * H;
    g2.M = (int)rows; g2.N = H; g2.K = H; g2.epi = EPI_RES_SILU; g2.alpha = 1.f;
    DDQST_TRY(launch_sgemm(g2, s));
  }
  GemmArgs gh{};
  gh.A = h; gh.a_rs = H; gh.a_cs = 1;
  gh.B = (const float*)(pack + pl.head_f32); gh.b_rs = 1; gh.b_cs = H;
  gh.C = logits; gh.ldc = 2 * N;
  gh.bias = (const float*)(pack + pl.head_b);
  gh.M = (int)rows; gh.N = 2 * N; gh.K = H; gh.epi = EPI_BIAS; gh.alpha = 1.f;
  DDQST_TRY(launch_sgemm(gh, s));
  return DDQST_OK;
}

// ------------------------------------------------------------------------------------ reverse step
__global__ void init_bits_kernel(int N, RowCtx ctx, int64_t rows, uint64_t seed, uint16_t* __restrict__ x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  uint32_t basis = (uint32_t)row_basis(ctx, i);
  uint64_t shot = row_shot(ctx, i);
  uint32_t out = 0;
  Philox4 p{};
  for (int q = 0; q < N; ++q) {
    if ((q & 3) == 0) p = stream_block(seed, basis, 0, DDQST_SITE_INIT, shot, q >> 2);
    out |= (lane_of(p, q) & 1u) << q;
  }
  x[i] = (uint16_t)out;
}

__global__ void reverse_step_kernel(int N, int T, const float* __restrict__ sched, int mode, int t, RowCtx ctx,
                                    int64_t rows, uint64_t seed, const float* __restrict__ logits,
                                    const uint16_t* __restrict__ x_t, uint16_t* __restrict__ x_prev) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* lg = logits + i * 2 * N;
  x_prev[i] = (uint16_t)reverse_step_bits(N, T, sched, mode, t, seed, (uint32_t)row_basis(ctx, i), row_shot(ctx, i),
                                          x_t[i], [&](int q, int c) { return lg[2 * q + c]; });
}

int launch_init_bits(int num_qubits, RowCtx ctx, int64_t rows, uint64_t seed, uint16_t* x, cudaStream_t s) {
  if (rows <= 0) return DDQST_OK;
  init_bits_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(num_qubits, ctx, rows, seed, x);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int launch_reverse_step(const ddqst_dims* d, const float* sched, int mode, int t, RowCtx ctx, int64_t rows,
                        uint64_t seed, const float* logits, const uint16_t* x_t, uint16_t* x_prev, cudaStream_t s) {
  if (rows <= 0) return DDQST_OK;
  reverse_step_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(d->num_qubits, d->num_timesteps, sched, mode, t,
                                                                     ctx, rows, seed, logits, x_t, x_prev);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // namespace ddqst
