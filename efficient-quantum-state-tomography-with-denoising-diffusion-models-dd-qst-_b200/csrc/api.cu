// C-ABI entry points: layouts, weight packing, forward, sampling dispatch, noising, bit packing.
#include <stdarg.h>
#include <string.h>

#include "sampler_tc.cuh"
#include "simt.cuh"

namespace ddqst {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_arch() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    set_error("no CUDA device available; libddqst has no CPU fallback");
    return DDQST_EUNSUPPORTED_ARCH;
  }
  if (major != 10) {
    set_error("device compute capability %d.x is not supported: libddqst is built for sm_100a only", major);
    return DDQST_EUNSUPPORTED_ARCH;
  }
  return DDQST_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

int validate_dims(const ddqst_dims* d) {
  DDQST_REQUIRE(d != nullptr, DDQST_EINVAL_SHAPE, "dims is NULL");
  DDQST_REQUIRE(d->num_qubits >= 1 && d->num_qubits <= 16, DDQST_EINVAL_SHAPE, "num_qubits=%d outside [1,16]", d->num_qubits);
  DDQST_REQUIRE(d->num_bases >= 1, DDQST_EINVAL_SHAPE, "num_bases=%d", d->num_bases);
  DDQST_REQUIRE(d->num_timesteps >= 1 && d->num_timesteps < 65536, DDQST_EINVAL_SHAPE, "num_timesteps=%d", d->num_timesteps);
  DDQST_REQUIRE(d->embed_dim >= 1 && d->hidden_dim >= 1, DDQST_EINVAL_SHAPE, "embed_dim=%d hidden_dim=%d", d->embed_dim, d->hidden_dim);
  DDQST_REQUIRE(d->num_blocks >= 1 && d->num_blocks <= 16, DDQST_EINVAL_SHAPE, "num_blocks=%d outside [1,16]", d->num_blocks);
  DDQST_REQUIRE(d->variant == DDQST_VARIANT_A || d->variant == DDQST_VARIANT_B, DDQST_EINVAL_SHAPE, "variant=%d", d->variant);
  return DDQST_OK;
}

int param_layout(const ddqst_dims* d, ParamLayout* out) {
  DDQST_TRY(validate_dims(d));
  const int64_t N = d->num_qubits, E = d->embed_dim, H = d->hidden_dim, L = d->num_blocks;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off = align_up(off + n, 4); return o; };
  if (d->variant == DDQST_VARIANT_B) { out->x_emb = take(2 * E); out->in_w = take(H * N * E); }
  else { out->x_emb = -1; out->in_w = take(H * N); }
  out->in_b = take(H);
  out->time_emb = take((int64_t)(d->num_timesteps + 1) * E);
  out->basis_emb = take((int64_t)d->num_bases * E);
  for (int l = 0; l < L; ++l) {
    out->film_w[l] = take(2 * H * 2 * E); out->film_b[l] = take(2 * H);
    out->w1[l] = take(H * H); out->b1[l] = take(H);
    out->w2[l] = take(H * H); out->b2[l] = take(H);
  }
  out->head_w = take(2 * N * H);
  out->head_b = take(2 * N);
  out->total = off;
  return DDQST_OK;
}

int pack_layout(const ddqst_dims* d, PackLayout* out) {
  DDQST_TRY(validate_dims(d));
  const int64_t N = d->num_qubits, H = d->hidden_dim, L = d->num_blocks, T = d->num_timesteps;
  out->head_pad = (int)align_up(2 * N, 16);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 1024); return o; };
  out->c0 = take(4 * H);
  out->D = take(4 * N * H);
  out->Tt = take(4 * (T + 1) * L * 2 * H);
  out->Tb = take(4 * (int64_t)d->num_bases * L * 2 * H);
  out->bias1 = take(4 * L * H);
  out->bias2 = take(4 * L * H);
  out->head_b = take(4 * 32);
  out->w1_f32 = take(4 * L * H * H);
  out->w2_f32 = take(4 * L * H * H);
  out->head_f32 = take(4 * (int64_t)out->head_pad * H);
  out->w_bf16 = take(2 * L * 2 * H * H);
  out->head_bf16 = take(2 * (int64_t)out->head_pad * H);
  out->dt_bf16 = take(2 * H * 64);
  out->total = off;
  return DDQST_OK;
}

// ------------------------------------------------------------------------------------ pack kernels
__global__ void collapse_input_kernel(int variant, int N, int E, int H, const float* __restrict__ x_emb,
                                      const float* __restrict__ in_w, const float* __restrict__ in_b,
                                      float* __restrict__ c0, float* __restrict__ D) {
  // one thread per hidden unit: c0[h] = b[h] + sum_q W[h,qE:(q+1)E].emb0 ; D[q,h] = W[h,q-slice].(emb1-emb0)
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float c = in_b[h];
  for (int q = 0; q < N; ++q) {
    if (variant == DDQST_VARIANT_B) {
      float d0 = 0.f, d1 = 0.f;
      for (int e = 0; e < E; ++e) {
        float w = in_w[(int64_t)h * N * E + q * E + e];
        d0 = fmaf(w, x_emb[e], d0);
        d1 = fmaf(w, x_emb[E + e], d1);
      }
      c += d0;
      D[q * H + h] = d1 - d0;
    } else {
      D[q * H + h] = in_w[(int64_t)h * N + q];
    }
  }
  c0[h] = c;
}

__global__ void copy_convert_kernel(const float* __restrict__ src, float* __restrict__ dst_f32,
                                    __nv_bfloat16* __restrict__ dst_bf16, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = src[i];
  if (dst_f32) dst_f32[i] = v;
  if (dst_bf16) dst_bf16[i] = __float2bfloat16_rn(v);
}

// Dt[h][k]: k<N hi(D[k][h]); k==N hi(c0[h]); 16<=k<16+N lo(D[k-16][h]); k==16+N lo(c0[h]); else 0
__global__ void build_dt_kernel(int N, int H, const float* __restrict__ c0, const float* __restrict__ D,
                                __nv_bfloat16* __restrict__ dt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H * 64) return;
  int h = e / 64, k = e % 64;
  int kk = k >= 16 ? k - 16 : k;
  float out = 0.f;
  if (k < 32 && kk <= N && N <= 15) {
    float v = kk < N ? D[kk * H + h] : c0[h];
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out = k < 16 ? __bfloat162float(hi) : v - __bfloat162float(hi);
  }
  dt[e] = __float2bfloat16_rn(out);
}

static int copy_convert(const float* src, float* f32, __nv_bfloat16* bf16, int64_t n, cudaStream_t s) {
  copy_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, f32, bf16, n);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // namespace ddqst

using namespace ddqst;

extern "C" {

const char* ddqst_last_error(void) { return g_err; }
int ddqst_version(void) { return 200; }

int64_t ddqst_param_count(const ddqst_dims* d, int64_t* offsets_out) {
  ParamLayout p;
  if (param_layout(d, &p) != DDQST_OK) return -1;
  if (offsets_out) {
    int k = 0;
    offsets_out[k++] = p.x_emb; offsets_out[k++] = p.in_w; offsets_out[k++] = p.in_b;
    offsets_out[k++] = p.time_emb; offsets_out[k++] = p.basis_emb;
    for (int l = 0; l < d->num_blocks; ++l) {
      offsets_out[k++] = p.film_w[l]; offsets_out[k++] = p.film_b[l];
      offsets_out[k++] = p.w1[l]; offsets_out[k++] = p.b1[l];
      offsets_out[k++] = p.w2[l]; offsets_out[k++] = p.b2[l];
    }
    offsets_out[k++] = p.head_w; offsets_out[k++] = p.head_b;
  }
  return p.total;
}

int64_t ddqst_pack_bytes(const ddqst_dims* d) {
  PackLayout p;
  if (pack_layout(d, &p) != DDQST_OK) return -1;
  return p.total;
}

int ddqst_pack_weights(const ddqst_dims* d, const float* params, void* pack_v, void* stream) {
  DDQST_TRY(check_arch());
  ParamLayout pr;
  PackLayout pl;
  DDQST_TRY(param_layout(d, &pr));
  DDQST_TRY(pack_layout(d, &pl));
  DDQST_REQUIRE(params && pack_v, DDQST_EINVAL_SHAPE, "params/pack is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  char* pack = (char*)pack_v;
  const int N = d->num_qubits, E = d->embed_dim, H = d->hidden_dim, L = d->num_blocks, T = d->num_timesteps;
  collapse_input_kernel<<<(H + 127) / 128, 128, 0, s>>>(d->variant, N, E, H, pr.x_emb >= 0 ? params + pr.x_emb : nullptr,
                                                        params + pr.in_w, params + pr.in_b, (float*)(pack + pl.c0),
                                                        (float*)(pack + pl.D));
  DDQST_LAUNCH_OK();
  build_dt_kernel<<<(H * 64 + 255) / 256, 256, 0, s>>>(N, H, (const float*)(pack + pl.c0), (const float*)(pack + pl.D),
                                                        (__nv_bfloat16*)(pack + pl.dt_bf16));
  DDQST_LAUNCH_OK();
  for (int l = 0; l < L; ++l) {
    // Tt[:, l, :] = time_emb . Wf[:, :E]^T        (cond = [t_emb || b_emb], RQC/model.py:62)
    GemmArgs g{};
    g.A = params + pr.time_emb; g.a_rs = E; g.a_cs = 1;
    g.B = params + pr.film_w[l]; g.b_rs = 1; g.b_cs = 2 * E;
    g.C = (float*)(pack + pl.Tt) + (int64_t)l * 2 * H; g.ldc = (int64_t)L * 2 * H;
    g.M = T + 1; g.N = 2 * H; g.K = E; g.epi = EPI_NONE; g.alpha = 1.f;
    DDQST_TRY(launch_sgemm(g, s));
    // Tb[:, l, :] = basis_emb . Wf[:, E:]^T + bf
    GemmArgs g2{};
    g2.A = params + pr.basis_emb; g2.a_rs = E; g2.a_cs = 1;
    g2.B = params + pr.film_w[l] + E; g2.b_rs = 1; g2.b_cs = 2 * E;
    g2.C = (float*)(pack + pl.Tb) + (int64_t)l * 2 * H; g2.ldc = (int64_t)L * 2 * H;
    g2.bias = params + pr.film_b[l];
    g2.M = d->num_bases; g2.N = 2 * H; g2.K = E; g2.epi = EPI_BIAS; g2.alpha = 1.f;
    DDQST_TRY(launch_sgemm(g2, s));
    __nv_bfloat16* wb = (__nv_bfloat16*)(pack + pl.w_bf16) + (int64_t)l * 2 * H * H;
    DDQST_TRY(copy_convert(params + pr.w1[l], (float*)(pack + pl.w1_f32) + (int64_t)l * H * H, wb, (int64_t)H * H, s));
    DDQST_TRY(copy_convert(params + pr.w2[l], (float*)(pack + pl.w2_f32) + (int64_t)l * H * H, wb + (int64_t)H * H, (int64_t)H * H, s));
    DDQST_TRY(copy_convert(params + pr.b1[l], (float*)(pack + pl.bias1) + l * H, nullptr, H, s));
    DDQST_TRY(copy_convert(params + pr.b2[l], (float*)(pack + pl.bias2) + l * H, nullptr, H, s));
  }
  DDQST_CUDA_OK(cudaMemsetAsync(pack + pl.head_f32, 0, 4 * (int64_t)pl.head_pad * H, s));
  DDQST_CUDA_OK(cudaMemsetAsync(pack + pl.head_bf16, 0, 2 * (int64_t)pl.head_pad * H, s));
  DDQST_CUDA_OK(cudaMemsetAsync(pack + pl.head_b, 0, 4 * 32, s));
  DDQST_TRY(copy_convert(params + pr.head_w, (float*)(pack + pl.head_f32), (__nv_bfloat16*)(pack + pl.head_bf16), (int64_t)2 * N * H, s));
  DDQST_TRY(copy_convert(params + pr.head_b, (float*)(pack + pl.head_b), nullptr, 2 * N, s));
  return DDQST_OK;
}

// ------------------------------------------------------------------------------------ workspace
static int64_t fp32_rows_bytes(const ddqst_dims* d) {
  // h, a, u [H] fp32 + logits [2N] fp32 + two x buffers
  return (int64_t)d->hidden_dim * 12 + (int64_t)d->num_qubits * 8 + 8;
}
static const int64_t kFp32ChunkRows = 32768;

int64_t ddqst_workspace_bytes(int op, const ddqst_dims* d, int64_t batch, int precision) {
  switch (op) {
    case DDQST_OP_FORWARD:
    case DDQST_OP_SAMPLE: {
      if (validate_dims(d) != DDQST_OK) return -1;
      if (precision == DDQST_PRECISION_BF16) return sampler_tc_workspace_bytes(d, batch);
      int64_t rows = batch < kFp32ChunkRows ? batch : kFp32ChunkRows;
      if (rows < 1) rows = 1;
      return align_up(rows * fp32_rows_bytes(d), 256) + 4096;
    }
    case DDQST_OP_LINEAR_INVERSION: {
      // batch = n_slots ; W int32 [n_slots, 2^N]
      { int64_t a = batch * ((int64_t)4 << d->num_qubits), b = (int64_t)8 << (2 * d->num_qubits); return align_up(a > b ? a : b, 256) + 256; }
    }
    case DDQST_OP_PSD:
    case DDQST_OP_METRICS: {
      int64_t dim = (int64_t)1 << d->num_qubits;
      return 2 * 16 * dim * dim + 64 * dim + 4096;
    }
    case DDQST_OP_FIDELITY_MIXED: {
      int64_t dim = (int64_t)1 << d->num_qubits;
      return 6 * 16 * dim * dim + 128 * dim + 8192;
    }
    case DDQST_OP_TRAIN:
      return precision == DDQST_PRECISION_BF16 ? train_tc_workspace_bytes(d, batch) : train_workspace_bytes(d, batch);
    default:
      set_error("unknown op %d", op);
      return -1;
  }
}

// ------------------------------------------------------------------------------------ forward
int ddqst_denoiser_forward(const ddqst_dims* d, const void* pack, int precision, const uint16_t* x_packed,
                           const int32_t* t, const int32_t* basis, int64_t batch, float* logits, void* workspace,
                           int64_t ws_bytes, void* stream) {
  DDQST_TRY(check_arch());
  PackLayout pl;
  DDQST_TRY(pack_layout(d, &pl));
  DDQST_REQUIRE(batch >= 0, DDQST_EINVAL_SHAPE, "batch=%lld", (long long)batch);
  if (batch == 0) return DDQST_OK;
  DDQST_REQUIRE(pack && x_packed && t && basis && logits, DDQST_EINVAL_SHAPE, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == DDQST_PRECISION_BF16)
    return sampler_tc_forward(d, (const char*)pack, pl, x_packed, t, basis, batch, logits, workspace, ws_bytes, s);
  DDQST_REQUIRE(precision == DDQST_PRECISION_FP32, DDQST_EINVAL_SHAPE, "precision=%d", precision);
  const int64_t per_row = fp32_rows_bytes(d);
  int64_t chunk = (ws_bytes - 4096) / per_row;
  DDQST_REQUIRE(workspace && chunk >= 1, DDQST_EWORKSPACE, "workspace of %lld bytes is too small", (long long)ws_bytes);
  if (chunk > batch) chunk = batch;
  for (int64_t r0 = 0; r0 < batch; r0 += chunk) {
    int64_t rows = batch - r0 < chunk ? batch - r0 : chunk;
    RowCtx ctx{t + r0, 0, basis + r0, nullptr, 0, r0, 0};
    DDQST_TRY(forward_fp32(d, (const char*)pack, pl, x_packed + r0, ctx, rows, logits + r0 * 2 * d->num_qubits,
                           (float*)workspace, s));
  }
  return DDQST_OK;
}

// ------------------------------------------------------------------------------------ sampling (fp32 path)
namespace ddqst {

__global__ void emit_kernel(const uint16_t* __restrict__ x, int64_t rows, int N, int64_t row0, int64_t spb,
                            void* __restrict__ out_packed, uint32_t* __restrict__ out_hist) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  uint32_t v = x[i];
  if (out_packed) {
    if (N <= 8) ((uint8_t*)out_packed)[row0 + i] = (uint8_t)v;
    else ((uint16_t*)out_packed)[row0 + i] = (uint16_t)v;
  }
  if (out_hist) atomicAdd(out_hist + (((row0 + i) / spb) << N) + v, 1u);
}

static int sample_fp32(const ddqst_dims* d, const char* pack, const PackLayout& pl, const float* sched, int mode,
                       const int32_t* basis_ids, int32_t n_bases, int64_t spb, int64_t shot_offset, uint64_t seed,
                       void* out_packed, uint32_t* out_hist, void* workspace, int64_t ws_bytes, cudaStream_t s) {
  const int N = d->num_qubits, H = d->hidden_dim, T = d->num_timesteps;
  const int64_t total = (int64_t)n_bases * spb;
  const int64_t per_row = fp32_rows_bytes(d);
  int64_t chunk = (ws_bytes - 4096) / per_row;
  DDQST_REQUIRE(workspace && chunk >= 1, DDQST_EWORKSPACE, "workspace of %lld bytes is too small", (long long)ws_bytes);
  if (chunk > total) chunk = total;
  float* act = (float*)workspace;
  float* logits = act + 3 * chunk * H;
  uint16_t* xa = (uint16_t*)(logits + chunk * 2 * N);
  uint16_t* xb = xa + align_up(chunk, 2);
  for (int64_t r0 = 0; r0 < total; r0 += chunk) {
    int64_t rows = total - r0 < chunk ? total - r0 : chunk;
    RowCtx ctx{nullptr, 0, nullptr, basis_ids, spb, r0, shot_offset};
    DDQST_TRY(launch_init_bits(N, ctx, rows, seed, xa, s));
    uint16_t *cur = xa, *nxt = xb;
    for (int t = T; t >= 1; --t) {
      ctx.t_uniform = t;
      DDQST_TRY(forward_fp32(d, pack, pl, cur, ctx, rows, logits, act, s));
      DDQST_TRY(launch_reverse_step(d, sched, mode, t, ctx, rows, seed, logits, cur, nxt, s));
      uint16_t* tmp = cur; cur = nxt; nxt = tmp;
    }
    emit_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(cur, rows, N, r0, spb, out_packed, out_hist);
    DDQST_LAUNCH_OK();
  }
  return DDQST_OK;
}

}  // namespace ddqst

int ddqst_sample(const ddqst_dims* d, const void* pack, const float* sched, int mode, int precision,
                 const int32_t* basis_ids, int32_t n_bases, int64_t shots_per_basis, int64_t shot_offset,
                 uint64_t seed, void* out_packed, uint32_t* out_hist, void* workspace, int64_t ws_bytes,
                 void* stream) {
  DDQST_TRY(check_arch());
  PackLayout pl;
  DDQST_TRY(pack_layout(d, &pl));
  DDQST_REQUIRE(n_bases >= 0 && shots_per_basis >= 0 && shot_offset >= 0, DDQST_EINVAL_SHAPE, "negative count");
  DDQST_REQUIRE(mode == DDQST_MODE_POSTERIOR || mode == DDQST_MODE_RENOISE, DDQST_EINVAL_SHAPE, "mode=%d", mode);
  if (n_bases == 0 || shots_per_basis == 0) return DDQST_OK;
  DDQST_REQUIRE(pack && sched && basis_ids, DDQST_EINVAL_SHAPE, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == DDQST_PRECISION_BF16)
    return sampler_tc_sample(d, (const char*)pack, pl, sched, mode, basis_ids, n_bases, shots_per_basis, shot_offset,
                             seed, out_packed, out_hist, workspace, ws_bytes, s);
  DDQST_REQUIRE(precision == DDQST_PRECISION_FP32, DDQST_EINVAL_SHAPE, "precision=%d", precision);
  return sample_fp32(d, (const char*)pack, pl, sched, mode, basis_ids, n_bases, shots_per_basis, shot_offset, seed,
                     out_packed, out_hist, workspace, ws_bytes, s);
}

int ddqst_sample_step(const ddqst_dims* d, const void* pack, const float* sched, int mode, int precision,
                      int32_t basis_id, int32_t t, int64_t shots, int64_t shot_offset, uint64_t seed,
                      const uint16_t* x_t, uint16_t* x_prev, float* logits_out, void* workspace, int64_t ws_bytes,
                      void* stream) {
  DDQST_TRY(check_arch());
  PackLayout pl;
  DDQST_TRY(pack_layout(d, &pl));
  DDQST_REQUIRE(t >= 1 && t <= d->num_timesteps, DDQST_EINVAL_SHAPE, "t=%d outside [1,%d]", t, d->num_timesteps);
  DDQST_REQUIRE(basis_id >= 0 && basis_id < d->num_bases, DDQST_EINVAL_SHAPE, "basis_id=%d", basis_id);
  if (shots == 0) return DDQST_OK;
  DDQST_REQUIRE(pack && sched && x_t && x_prev && workspace, DDQST_EINVAL_SHAPE, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == DDQST_PRECISION_BF16)
    return sampler_tc_step(d, (const char*)pack, pl, sched, mode, basis_id, t, shots, shot_offset, seed, x_t, x_prev,
                           logits_out, workspace, ws_bytes, s);
  const int N = d->num_qubits, H = d->hidden_dim;
  const int64_t need = shots * fp32_rows_bytes(d) + 4096;
  DDQST_REQUIRE(ws_bytes >= need, DDQST_EWORKSPACE, "sample_step needs %lld workspace bytes, got %lld", (long long)need, (long long)ws_bytes);
  float* act = (float*)workspace;
  float* logits = act + 3 * shots * H;
  int32_t* bid = (int32_t*)(logits + shots * 2 * N);
  DDQST_CUDA_OK(cudaMemcpyAsync(bid, &basis_id, sizeof(int32_t), cudaMemcpyHostToDevice, s));
  RowCtx ctx{nullptr, t, nullptr, bid, shots, 0, shot_offset};
  DDQST_TRY(forward_fp32(d, (const char*)pack, pl, x_t, ctx, shots, logits, act, s));
  DDQST_TRY(launch_reverse_step(d, sched, mode, t, ctx, shots, seed, logits, x_t, x_prev, s));
  if (logits_out) DDQST_CUDA_OK(cudaMemcpyAsync(logits_out, logits, sizeof(float) * shots * 2 * N, cudaMemcpyDeviceToDevice, s));
  return DDQST_OK;
}

// ------------------------------------------------------------------------------------ q_sample
namespace ddqst {
extern "C++" {
// NBLK = ceil(N / 4) Philox blocks of qubit draws per sample, a compile-time count: the qubit loop unrolls, the word of a qubit is a
// fixed lane of its block, and the two transition rows of Q[t] are read once as a float4
template <int NBLK>
__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ Q, int T, int N, int cumulative,
                                const uint16_t* __restrict__ x0, const int32_t* __restrict__ t_in, int64_t batch,
                                int64_t row_offset, uint64_t seed, uint32_t stream_id,
                                const int64_t* __restrict__ stream_dev, uint16_t* __restrict__ xt,
                                int32_t* __restrict__ t_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  if (stream_dev) stream_id = (uint32_t)stream_dev[0];     // step counter kept on the device (CUDA-graph replay)
  uint64_t row = (uint64_t)(row_offset + i);
  int t;
  if (t_in) t = t_in[i];
  else {
    Philox4 pt = stream_block(seed, stream_id, 0, DDQST_SITE_TSTEP, row, 0);
    t = 1 + (int)(((uint64_t)(pt.x >> 8) * (uint64_t)T) >> 24);
  }
  if (t_out) t_out[i] = t;
  const float4 qt = __ldg(reinterpret_cast<const float4*>(Q) + t);       // [a][b] row-major: x = [0][0], y = [0][1], z = [1][0], w = [1][1]
  // from-bit 0 / 1: (p0, p1) = cumulative ? Q_bar[t][from][to 0 / 1] : Q[t][to 0 / 1][from]
  const float p0_f0 = qt.x, p1_f0 = cumulative ? qt.y : qt.z;
  const float p0_f1 = cumulative ? qt.z : qt.y, p1_f1 = qt.w;
  const float s_f0 = __fadd_rn(p0_f0, p1_f0), s_f1 = __fadd_rn(p0_f1, p1_f1);
  const uint32_t bits = x0[i];
  uint32_t out = 0;
#pragma unroll
  for (int blk = 0; blk < NBLK; ++blk) {
    if (blk * 4 < N) {
      const Philox4 p = stream_block(seed, stream_id, 0, DDQST_SITE_QSAMPLE, row, blk);
      const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = blk * 4 + j;
        if (q < N) {
          const bool from1 = (bits >> q) & 1u;
          const float u = word_to_uniform(w[j]);
          out |= (__fmul_rn(u, from1 ? s_f1 : s_f0) < (from1 ? p1_f1 : p1_f0) ? 1u : 0u) << q;     // draw_bit(u, p0, p1)
        }
      }
    }
  }
  xt[i] = (uint16_t)out;
}

static void launch_q_sample(const float* Q, int T, int N, int cumulative, const uint16_t* x0, const int32_t* t_in, int64_t batch,
                            int64_t row_offset, uint64_t seed, uint32_t stream_id, const int64_t* stream_dev, uint16_t* xt, int32_t* t_out,
                            cudaStream_t s) {
  const unsigned grid = (unsigned)((batch + 255) / 256);
  switch ((N + 3) / 4) {
    case 1: q_sample_kernel<1><<<grid, 256, 0, s>>>(Q, T, N, cumulative, x0, t_in, batch, row_offset, seed, stream_id, stream_dev, xt, t_out); break;
    case 2: q_sample_kernel<2><<<grid, 256, 0, s>>>(Q, T, N, cumulative, x0, t_in, batch, row_offset, seed, stream_id, stream_dev, xt, t_out); break;
    case 3: q_sample_kernel<3><<<grid, 256, 0, s>>>(Q, T, N, cumulative, x0, t_in, batch, row_offset, seed, stream_id, stream_dev, xt, t_out); break;
    default: q_sample_kernel<4><<<grid, 256, 0, s>>>(Q, T, N, cumulative, x0, t_in, batch, row_offset, seed, stream_id, stream_dev, xt, t_out); break;
  }
}
}  // extern "C++"

__global__ void pack_bits_kernel(const int64_t* __restrict__ bits, int64_t batch, int N, uint16_t* __restrict__ packed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  uint32_t v = 0;
  for (int q = 0; q < N; ++q) v |= (uint32_t)(bits[i * N + q] & 1) << q;
  packed[i] = (uint16_t)v;
}

__global__ void unpack_bits_kernel(const void* __restrict__ packed, int elem_bytes, int64_t total, int N,
                                   int64_t* __restrict__ bits) {
  // one thread per PAIR of output elements: coalesced 16-byte stores (the int64[B,N] output is 64x the input bytes, so
  // the store width is what matters); an odd total leaves one element for the last thread
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e = 2 * p;
  if (e >= total) return;
  auto bit_of = [&](int64_t el) -> long long {
    const int64_t i = el / N;
    const int q = (int)(el - i * N);
    const uint32_t v = elem_bytes == 1 ? ((const uint8_t*)packed)[i] : ((const uint16_t*)packed)[i];
    return (long long)((v >> q) & 1u);
  };
  if (e + 1 < total) *reinterpret_cast<longlong2*>(bits + e) = make_longlong2(bit_of(e), bit_of(e + 1));
  else bits[e] = bit_of(e);
}
}  // namespace ddqst

int ddqst_q_sample(const float* Q, int32_t num_timesteps, int32_t num_qubits, int cumulative,
                   const uint16_t* x0_packed, const int32_t* t, int64_t batch, int64_t row_offset, uint64_t seed,
                   uint32_t stream_id, uint16_t* xt_packed, int32_t* t_out, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(num_qubits >= 1 && num_qubits <= 16 && num_timesteps >= 1 && batch >= 0, DDQST_EINVAL_SHAPE, "bad shape");
  if (batch == 0) return DDQST_OK;
  DDQST_REQUIRE(Q && x0_packed && xt_packed, DDQST_EINVAL_SHAPE, "NULL argument");
  DDQST_REQUIRE(((uintptr_t)Q & 15) == 0, DDQST_EINVAL_SHAPE, "the transition table must be 16-byte aligned");
  ddqst::launch_q_sample(Q, num_timesteps, num_qubits, cumulative, x0_packed, t, batch, row_offset, seed, stream_id, nullptr, xt_packed, t_out,
                         (cudaStream_t)stream);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_q_sample_dev(const float* Q, int32_t num_timesteps, int32_t num_qubits, int cumulative,
                       const uint16_t* x0_packed, const int32_t* t, int64_t batch, int64_t row_offset, uint64_t seed,
                       const int64_t* stream_id_dev, uint16_t* xt_packed, int32_t* t_out, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(num_qubits >= 1 && num_qubits <= 16 && num_timesteps >= 1 && batch >= 0, DDQST_EINVAL_SHAPE, "bad shape");
  if (batch == 0) return DDQST_OK;
  DDQST_REQUIRE(Q && x0_packed && xt_packed && stream_id_dev, DDQST_EINVAL_SHAPE, "NULL argument");
  DDQST_REQUIRE(((uintptr_t)Q & 15) == 0, DDQST_EINVAL_SHAPE, "the transition table must be 16-byte aligned");
  ddqst::launch_q_sample(Q, num_timesteps, num_qubits, cumulative, x0_packed, t, batch, row_offset, seed, 0u, stream_id_dev, xt_packed, t_out,
                         (cudaStream_t)stream);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_pack_bits(const int64_t* bits, int64_t batch, int32_t num_qubits, uint16_t* packed, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(num_qubits >= 1 && num_qubits <= 16 && batch >= 0, DDQST_EINVAL_SHAPE, "bad shape");
  if (batch == 0) return DDQST_OK;
  pack_bits_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, batch, num_qubits, packed);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

namespace ddqst {
// N = 8, uint8 shots (the C4 shape): four threads per shot, each writing qubits 2j, 2j+1 of it as one 16-byte store -- a warp stores
// 512 contiguous bytes per instruction, reads 8 bytes, and does no integer division
__global__ void __launch_bounds__(256) unpack_bits8_kernel(const uint8_t* __restrict__ packed, int64_t batch, int64_t* __restrict__ bits) {
  const int64_t th = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t shot = th >> 2;
  if (shot >= batch) return;
  const uint32_t v = packed[shot] >> (2 * (int)(th & 3));
  *reinterpret_cast<longlong2*>(bits + 2 * th) = make_longlong2((long long)(v & 1u), (long long)((v >> 1) & 1u));
}
}  // namespace ddqst

int ddqst_unpack_bits(const void* packed, int elem_bytes, int64_t batch, int32_t num_qubits, int64_t* bits, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(num_qubits >= 1 && num_qubits <= 16 && batch >= 0 && (elem_bytes == 1 || elem_bytes == 2), DDQST_EINVAL_SHAPE, "bad shape");
  if (batch == 0) return DDQST_OK;
  int64_t total = batch * num_qubits;
  DDQST_REQUIRE(((uintptr_t)bits & 15) == 0, DDQST_EINVAL_SHAPE, "bits must be 16-byte aligned");
  static int fast8 = -1;
  if (fast8 < 0) { const char* e = getenv("DDQST_UNPACK_FAST8"); fast8 = (e && e[0] == '0') ? 0 : 1; }
  if (fast8 && num_qubits == 8 && elem_bytes == 1)
    ddqst::unpack_bits8_kernel<<<(unsigned)((4 * batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)packed, batch, bits);
  else
    unpack_bits_kernel<<<(unsigned)(((total + 1) / 2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(packed, elem_bytes, total, num_qubits, bits);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_sample_host(const ddqst_dims* d, const void* pack, const float* sched, int mode, int precision,
                      const int32_t* basis_ids_host, int32_t n_bases, int64_t shots_per_basis, int64_t shot_offset,
                      uint64_t seed, void* out_packed_host, uint32_t* out_hist_host, void* dev_scratch,
                      int64_t dev_scratch_bytes, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_TRY(validate_dims(d));
  cudaStream_t s = (cudaStream_t)stream;
  const int N = d->num_qubits;
  const int64_t total = (int64_t)n_bases * shots_per_basis;
  const int eb = N <= 8 ? 1 : 2;
  // scratch: [basis ids][hist][packed][workspace]
  int64_t off_ids = 0, off_hist = align_up(4 * (int64_t)n_bases, 256);
  int64_t off_packed = off_hist + align_up(out_hist_host ? ((int64_t)n_bases << N) * 4 : 0, 256);
  int64_t off_ws = off_packed + align_up(out_packed_host ? total * eb : 0, 256);
  DDQST_REQUIRE(dev_scratch && dev_scratch_bytes > off_ws, DDQST_EWORKSPACE, "device scratch too small");
  char* base = (char*)dev_scratch;
  DDQST_CUDA_OK(cudaMemcpyAsync(base + off_ids, basis_ids_host, 4 * (int64_t)n_bases, cudaMemcpyHostToDevice, s));
  if (out_hist_host) DDQST_CUDA_OK(cudaMemsetAsync(base + off_hist, 0, ((int64_t)n_bases << N) * 4, s));
  DDQST_TRY(ddqst_sample(d, pack, sched, mode, precision, (const int32_t*)(base + off_ids), n_bases, shots_per_basis,
                         shot_offset, seed, out_packed_host ? base + off_packed : nullptr,
                         out_hist_host ? (uint32_t*)(base + off_hist) : nullptr, base + off_ws,
                         dev_scratch_bytes - off_ws, stream));
  if (out_packed_host) DDQST_CUDA_OK(cudaMemcpyAsync(out_packed_host, base + off_packed, total * eb, cudaMemcpyDeviceToHost, s));
  if (out_hist_host) DDQST_CUDA_OK(cudaMemcpyAsync(out_hist_host, base + off_hist, ((int64_t)n_bases << N) * 4, cudaMemcpyDeviceToHost, s));
  DDQST_CUDA_OK(cudaStreamSynchronize(s));
  return DDQST_OK;
}

// ------------------------------------------------------------------------------------ self tests
namespace ddqst {
__global__ void philox_selftest_kernel(const uint32_t* __restrict__ ck, int64_t n, uint32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox4 p = philox4x32_10(ck[i * 6], ck[i * 6 + 1], ck[i * 6 + 2], ck[i * 6 + 3], ck[i * 6 + 4], ck[i * 6 + 5]);
  out[i * 4] = p.x; out[i * 4 + 1] = p.y; out[i * 4 + 2] = p.z; out[i * 4 + 3] = p.w;
}
}  // namespace ddqst

int ddqst_selftest_philox(const uint32_t* ctr_key, int64_t n, uint32_t* out, void* stream) {
  DDQST_TRY(check_arch());
  if (n <= 0) return DDQST_OK;
  philox_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctr_key, n, out);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // extern "C"
