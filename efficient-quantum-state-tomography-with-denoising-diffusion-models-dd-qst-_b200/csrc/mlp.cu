// Single-qubit phase (config C1): the notebook's concat-conditioned MLP denoisers and their DDM.
//   SimpleMLP   NB c6:65-102  : cat[x, t_emb(32), b_emb(32)] -> 128 -> 128 -> 2, ReLU
//   UpgradedMLP NB c12:58-94  : cat[x, t_emb(128), b_emb(128)] -> 256 -> 256 -> 256 -> 2, ReLU
//   BitstringDDM NB c6:106-221: p_stay = linspace(1, .5, T+1); forward_diffusion; train_step; sample (x0-hat then re-noise)
// fp32 CUDA-core arithmetic on the shared SGEMM; the reverse step reuses reverse_step_bits with N = 1.
#include "sampler_tc.cuh"
#include "simt.cuh"

namespace ddqst {

struct MlpLayout {
  int64_t time_emb, basis_emb, w[8], b[8], total;
  int in_dim;
};

static int mlp_layout(const ddqst_mlp_dims* d, MlpLayout* out) {
  DDQST_REQUIRE(d != nullptr, DDQST_EINVAL_SHAPE, "dims is NULL");
  DDQST_REQUIRE(d->num_bases >= 1 && d->num_timesteps >= 1 && d->num_timesteps < 65536 && d->embed_dim >= 1 && d->hidden_dim >= 1,
                DDQST_EINVAL_SHAPE, "bad MLP dims");
  DDQST_REQUIRE(d->num_hidden >= 1 && d->num_hidden <= 6, DDQST_EINVAL_SHAPE, "num_hidden=%d outside [1,6]", d->num_hidden);
  const int64_t E = d->embed_dim, H = d->hidden_dim;
  out->in_dim = 1 + 2 * (int)E;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off = align_up(off + n, 4); return o; };
  out->time_emb = take((int64_t)(d->num_timesteps + 1) * E);
  out->basis_emb = take((int64_t)d->num_bases * E);
  for (int l = 0; l < d->num_hidden; ++l) {
    out->w[l] = take(H * (l == 0 ? out->in_dim : H));
    out->b[l] = take(H);
  }
  out->w[d->num_hidden] = take(2 * H);
  out->b[d->num_hidden] = take(2);
  out->total = off;
  return DDQST_OK;
}

// workspace (floats): xin[B,in] | act[l][B,H] (post-ReLU), l < num_hidden | logits[B,2] | dcur[B,H] | dnext[B,max(H,in)] | x bits
static int64_t mlp_ws_floats(const ddqst_mlp_dims* d, int64_t B) {
  const int64_t H = d->hidden_dim, in = 1 + 2 * d->embed_dim;
  return align_up(B * in, 64) + (int64_t)d->num_hidden * align_up(B * H, 64) + align_up(B * 2, 64) + align_up(B * H, 64) +
         align_up(B * (H > in ? H : in), 64) + align_up(B, 64) + 256;
}

__global__ void mlp_gather_kernel(int E, const float* __restrict__ time_emb, const float* __restrict__ basis_emb,
                                  const uint16_t* __restrict__ x, const int32_t* __restrict__ t_arr, int32_t t_uniform,
                                  const int32_t* __restrict__ b_arr, int32_t b_uniform, int64_t B, float* __restrict__ xin) {
  const int64_t i = blockIdx.x;
  if (i >= B) return;
  const int in = 1 + 2 * E;
  const int t = t_arr ? t_arr[i] : t_uniform, b = b_arr ? b_arr[i] : b_uniform;
  if (threadIdx.x == 0) xin[i * in] = (float)(x[i] & 1u);
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    xin[i * in + 1 + e] = time_emb[(int64_t)t * E + e];
    xin[i * in + 1 + E + e] = basis_emb[(int64_t)b * E + e];
  }
}

// dz = dy * (act > 0)
__global__ void relu_backward_kernel(int64_t n, const float* __restrict__ act, float* __restrict__ d) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n && !(act[e] > 0.f)) d[e] = 0.f;
}

__global__ void mlp_emb_scatter_kernel(int E, const float* __restrict__ dxin, const int32_t* __restrict__ t, const int32_t* __restrict__ basis,
                                       float* __restrict__ g_time, float* __restrict__ g_basis) {
  const int64_t i = blockIdx.x;
  const int in = 1 + 2 * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    atomicAdd(g_time + (int64_t)t[i] * E + e, dxin[i * in + 1 + e]);
    atomicAdd(g_basis + (int64_t)basis[i] * E + e, dxin[i * in + 1 + E + e]);
  }
}

__global__ void colsum_rows_kernel(const float* __restrict__ M, int64_t rows, int cols, int64_t ld, float* __restrict__ out) {
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + (threadIdx.x & 31), r0 = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < cols) for (int64_t r = r0; r < rows; r += 8) acc += M[r * ld + c];
  red[r0][threadIdx.x & 31] = acc;
  __syncthreads();
  if (r0 == 0 && c < cols) { float s = 0.f; for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x & 31]; out[c] = s; }
}

struct MlpCtx { const ddqst_mlp_dims* d; MlpLayout L; float* ws; int64_t B; cudaStream_t s; float *xin, *act, *logits, *dcur, *dnext; };

static int mlp_ctx(MlpCtx* c, const ddqst_mlp_dims* d, int64_t B, void* workspace, int64_t ws_bytes, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_TRY(mlp_layout(d, &c->L));
  DDQST_REQUIRE(B >= 1, DDQST_EINVAL_SHAPE, "batch=%lld", (long long)B);
  const int64_t need = mlp_ws_floats(d, B) * 4;
  DDQST_REQUIRE(workspace && ws_bytes >= need, DDQST_EWORKSPACE, "MLP op needs %lld workspace bytes, got %lld", (long long)need, (long long)ws_bytes);
  const int64_t H = d->hidden_dim, in = c->L.in_dim;
  c->d = d; c->B = B; c->s = (cudaStream_t)stream; c->ws = (float*)workspace;
  c->xin = c->ws;
  c->act = c->xin + align_up(B * in, 64);
  c->logits = c->act + (int64_t)d->num_hidden * align_up(B * H, 64);
  c->dcur = c->logits + align_up(B * 2, 64);
  c->dnext = c->dcur + align_up(B * H, 64);
  return DDQST_OK;
}

static int mlp_forward(const MlpCtx& c, const float* params, const uint16_t* x, const int32_t* t_arr, int32_t t_uni,
                       const int32_t* b_arr, int32_t b_uni) {
  const ddqst_mlp_dims* d = c.d;
  const int H = d->hidden_dim, E = d->embed_dim, in = c.L.in_dim;
  const int64_t B = c.B, stride = align_up(B * H, 64);
  mlp_gather_kernel<<<(unsigned)B, 64, 0, c.s>>>(E, params + c.L.time_emb, params + c.L.basis_emb, x, t_arr, t_uni, b_arr, b_uni, B, c.xin);
  DDQST_LAUNCH_OK();
  const float* cur = c.xin;
  int K = in;
  for (int l = 0; l < d->num_hidden; ++l) {
    GemmArgs g{};
    g.A = cur; g.a_rs = K; g.a_cs = 1; g.B = params + c.L.w[l]; g.b_rs = 1; g.b_cs = K;
    g.C = c.act + l * stride; g.ldc = H; g.bias = params + c.L.b[l]; g.M = (int)B; g.N = H; g.K = K; g.epi = EPI_BIAS_RELU; g.alpha = 1.f;
    DDQST_TRY(launch_sgemm(g, c.s));
    cur = c.act + l * stride; K = H;
  }
  GemmArgs g{};
  g.A = cur; g.a_rs = H; g.a_cs = 1; g.B = params + c.L.w[d->num_hidden]; g.b_rs = 1; g.b_cs = H;
  g.C = c.logits; g.ldc = 2; g.bias = params + c.L.b[d->num_hidden]; g.M = (int)B; g.N = 2; g.K = H; g.epi = EPI_BIAS; g.alpha = 1.f;
  return launch_sgemm(g, c.s);
}

// ---- the notebook sampler as ONE launch.  The denoiser sees (x_t, t, basis) with x_t a single bit, so for a fixed basis its output
// takes only 2 T distinct values: logits[t][x_t].  They are computed once, by the same forward kernels as before, on the 2 T rows
// (x, t) = (0,1), (1,1), (0,2), ... -- and the T-step chain of NB c6:189-221 becomes a 2-state Markov chain per sample that never
// leaves the registers: one thread per sample, the table in shared memory, Philox draws exactly as the per-step kernel of round 1 made them.
// (Round 1 launched the gather + (num_hidden + 1) SGEMMs + the draw kernel for every step: 404 SGEMM and 240 elementwise launches
// for T = 100.)  Same logits bit for bit -- every SGEMM output row depends on its own input row only -- hence the same samples.
__global__ void mlp_table_inputs_kernel(int T, uint16_t* __restrict__ x, int32_t* __restrict__ t) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 2 * T) return;
  x[r] = (uint16_t)(r & 1);
  t[r] = 1 + (r >> 1);
}

__global__ void __launch_bounds__(256) mlp_chain_kernel(int T, const float* __restrict__ sched, const float* __restrict__ table, int64_t n,
                                                        uint64_t seed, uint32_t basis, int64_t shot_offset, uint8_t* __restrict__ out,
                                                        uint32_t* __restrict__ hist) {
  extern __shared__ float s_tab[];                    // [2T][2] logits, then the schedule (betas[T+1], Q[T+1][4])
  float* s_sched = s_tab + 4 * T;
  for (int i = threadIdx.x; i < 4 * T; i += blockDim.x) s_tab[i] = table[i];
  for (int i = threadIdx.x; i < 5 * (T + 1); i += blockDim.x) s_sched[i] = sched[i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t ones = 0;
  if (i < n) {
    const uint64_t shot = (uint64_t)(shot_offset + i);
    uint32_t x = stream_block(seed, basis, 0, DDQST_SITE_INIT, shot, 0).x & 1u;
    for (int t = T; t >= 1; --t) {
      const float* lg = s_tab + 4 * (t - 1) + 2 * x;
      x = reverse_step_bits(1, T, s_sched, DDQST_MODE_RENOISE, t, seed, basis, shot, x, [&](int, int cc) { return lg[cc]; });
    }
    if (out) out[i] = (uint8_t)x;
    ones = x;
  }
  if (hist) {
    const uint32_t valid = __ballot_sync(0xFFFFFFFFu, i < n), set = __ballot_sync(0xFFFFFFFFu, ones != 0);
    if ((threadIdx.x & 31) == 0) {
      const uint32_t n1 = __popc(set), n0 = __popc(valid) - n1;
      if (n0) atomicAdd(hist, n0);
      if (n1) atomicAdd(hist + 1, n1);
    }
  }
}

}  // namespace ddqst

using namespace ddqst;

extern "C" {

int64_t ddqst_mlp_param_count(const ddqst_mlp_dims* d, int64_t* offsets_out) {
  MlpLayout L;
  if (mlp_layout(d, &L) != DDQST_OK) return -1;
  if (offsets_out) {
    int k = 0;
    offsets_out[k++] = L.time_emb; offsets_out[k++] = L.basis_emb;
    for (int l = 0; l <= d->num_hidden; ++l) { offsets_out[k++] = L.w[l]; offsets_out[k++] = L.b[l]; }
  }
  return L.total;
}

int64_t ddqst_mlp_workspace_bytes(const ddqst_mlp_dims* d, int64_t batch) {
  MlpLayout L;
  if (mlp_layout(d, &L) != DDQST_OK) return -1;
  return mlp_ws_floats(d, batch < 1 ? 1 : batch) * 4;
}

int ddqst_mlp_forward_saved(const ddqst_mlp_dims* d, const float* params, const uint16_t* x, const int32_t* t, const int32_t* basis,
                            int64_t batch, float* logits_out, void* workspace, int64_t ws_bytes, void* stream) {
  MlpCtx c;
  DDQST_TRY(mlp_ctx(&c, d, batch, workspace, ws_bytes, stream));
  DDQST_REQUIRE(params && x && t && basis, DDQST_EINVAL_SHAPE, "NULL argument");
  DDQST_TRY(mlp_forward(c, params, x, t, 0, basis, 0));
  if (logits_out) DDQST_CUDA_OK(cudaMemcpyAsync(logits_out, c.logits, sizeof(float) * batch * 2, cudaMemcpyDeviceToDevice, c.s));
  return DDQST_OK;
}

int ddqst_mlp_backward_saved(const ddqst_mlp_dims* d, const float* params, const int32_t* t, const int32_t* basis, int64_t batch,
                             const float* dlogits, float* grads, void* workspace, int64_t ws_bytes, void* stream) {
  MlpCtx c;
  DDQST_TRY(mlp_ctx(&c, d, batch, workspace, ws_bytes, stream));
  DDQST_REQUIRE(params && t && basis && dlogits && grads, DDQST_EINVAL_SHAPE, "NULL argument");
  const int H = d->hidden_dim, E = d->embed_dim, in = c.L.in_dim, nh = d->num_hidden;
  const int64_t B = batch, stride = align_up(B * H, 64);
  cudaStream_t s = c.s;
  DDQST_CUDA_OK(cudaMemsetAsync(grads, 0, sizeof(float) * c.L.total, s));
  auto wgrad = [&](const float* dY, int Nout, const float* X, int K, float* dW) {
    GemmArgs g{};
    g.A = dY; g.a_rs = 1; g.a_cs = Nout; g.B = X; g.b_rs = K; g.b_cs = 1; g.C = dW; g.ldc = K;
    g.M = Nout; g.N = K; g.K = (int)B; g.epi = EPI_NONE; g.alpha = 1.f;
    return launch_sgemm(g, s);
  };
  auto dgrad = [&](const float* dY, int Nout, const float* W, int K, float* dX) {
    GemmArgs g{};
    g.A = dY; g.a_rs = Nout; g.a_cs = 1; g.B = W; g.b_rs = K; g.b_cs = 1; g.C = dX; g.ldc = K;
    g.M = (int)B; g.N = K; g.K = Nout; g.epi = EPI_NONE; g.alpha = 1.f;
    return launch_sgemm(g, s);
  };
  // output layer
  const float* h_last = c.act + (nh - 1) * stride;
  DDQST_TRY(wgrad(dlogits, 2, h_last, H, grads + c.L.w[nh]));
  colsum_rows_kernel<<<1, 256, 0, s>>>(dlogits, B, 2, 2, grads + c.L.b[nh]);
  DDQST_LAUNCH_OK();
  DDQST_TRY(dgrad(dlogits, 2, params + c.L.w[nh], H, c.dcur));
  float* dcur = c.dcur;
  float* dnext = c.dnext;
  for (int l = nh - 1; l >= 0; --l) {
    relu_backward_kernel<<<(unsigned)((B * H + 255) / 256), 256, 0, s>>>(B * H, c.act + l * stride, dcur);
    DDQST_LAUNCH_OK();
    const float* X = l == 0 ? c.xin : c.act + (l - 1) * stride;
    const int K = l == 0 ? in : H;
    DDQST_TRY(wgrad(dcur, H, X, K, grads + c.L.w[l]));
    colsum_rows_kernel<<<(H + 31) / 32, 256, 0, s>>>(dcur, B, H, H, grads + c.L.b[l]);
    DDQST_LAUNCH_OK();
    DDQST_TRY(dgrad(dcur, H, params + c.L.w[l], K, dnext));
    float* tmp = dcur; dcur = dnext; dnext = tmp;
  }
  mlp_emb_scatter_kernel<<<(unsigned)B, 64, 0, s>>>(E, dcur, t, basis, grads + c.L.time_emb, grads + c.L.basis_emb);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_mlp_sample(const ddqst_mlp_dims* d, const float* params, const float* sched, int32_t basis_id, int64_t n,
                     int64_t shot_offset, uint64_t seed, uint8_t* out_bits, uint32_t* out_hist, void* workspace, int64_t ws_bytes,
                     void* stream) {
  DDQST_TRY(check_arch());
  if (n == 0) return DDQST_OK;
  DDQST_REQUIRE(n > 0 && params && sched, DDQST_EINVAL_SHAPE, "bad argument");
  DDQST_REQUIRE(basis_id >= 0 && basis_id < d->num_bases, DDQST_EINVAL_SHAPE, "basis_id=%d", basis_id);
  const int T = d->num_timesteps;
  const int64_t rows = 2 * (int64_t)T;
  const int64_t ctx_bytes = ddqst_mlp_workspace_bytes(d, rows);
  if (ctx_bytes < 0) return DDQST_EINVAL_SHAPE;
  const int64_t need = ctx_bytes + align_up(rows * 2, 256) + align_up(rows * 4, 256);
  DDQST_REQUIRE(workspace && ws_bytes >= need, DDQST_EWORKSPACE, "mlp_sample needs %lld workspace bytes, got %lld", (long long)need, (long long)ws_bytes);
  MlpCtx c;
  DDQST_TRY(mlp_ctx(&c, d, rows, workspace, ctx_bytes, stream));
  uint16_t* x_rows = (uint16_t*)((char*)workspace + ctx_bytes);
  int32_t* t_rows = (int32_t*)((char*)workspace + ctx_bytes + align_up(rows * 2, 256));
  mlp_table_inputs_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, c.s>>>(T, x_rows, t_rows);
  DDQST_LAUNCH_OK();
  DDQST_TRY(mlp_forward(c, params, x_rows, t_rows, 0, nullptr, basis_id));            // logits[2T][2]: row 2 (t-1) + x
  const size_t smem = sizeof(float) * (4 * (size_t)T + 5 * ((size_t)T + 1));
  DDQST_REQUIRE(smem <= 48 * 1024, DDQST_EUNSUPPORTED, "num_timesteps=%d: the logit table does not fit shared memory", T);
  mlp_chain_kernel<<<(unsigned)((n + 255) / 256), 256, smem, c.s>>>(T, sched, c.logits, n, seed, (uint32_t)basis_id, shot_offset, out_bits, out_hist);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // extern "C"
