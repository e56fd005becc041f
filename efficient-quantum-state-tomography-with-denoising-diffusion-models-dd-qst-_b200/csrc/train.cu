// Training step (T1): forward with saved activations, mean cross-entropy, backward into the flat
// gradient buffer, and the fused Adam / AdamW update.  fp32 CUDA-core arithmetic (the reference trains
// in fp32, RQC/main.py:105-115).
#include "sampler_tc.cuh"
#include "simt.cuh"

namespace ddqst {

// ---- workspace layout (floats), B = batch, per block l:
//   cond[B,2E] | xemb[B,N*E] (variant B) or xf[B,N] (variant A) | h_in[L+1][B,H] | gb[L][B,2H] | a[L][B,H]
//   | z1[L][B,H] | u[L][B,H] | z2[L][B,H] | logits[B,2N] | dlogits[B,2N] | dh[B,H] | dtmp[B,H] | du[B,H]
//   | dgb[B,2H] | dcond[B,2E] | dxemb[B,N*E]
struct TrainWs {
  int64_t cond, xin, h, gb, a, z1, u, z2, logits, dlogits, dh, dtmp, du, dgb, dcond, dxin, loss_part, total;
};

static void train_ws_layout(const ddqst_dims* d, int64_t B, TrainWs* w) {
  const int64_t N = d->num_qubits, E = d->embed_dim, H = d->hidden_dim, L = d->num_blocks;
  const int64_t xin = d->variant == DDQST_VARIANT_B ? N * E : N;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off = align_up(off + n, 64); return o; };
  w->cond = take(B * 2 * E); w->xin = take(B * xin); w->h = take((L + 1) * B * H); w->gb = take(L * B * 2 * H);
  w->a = take(L * B * H); w->z1 = take(L * B * H); w->u = take(L * B * H); w->z2 = take(L * B * H);
  w->logits = take(B * 2 * N); w->dlogits = take(B * 2 * N); w->dh = take(B * H); w->dtmp = take(B * H);
  w->du = take(B * H); w->dgb = take(B * 2 * H); w->dcond = take(B * 2 * E); w->dxin = take(B * xin);
  w->loss_part = take(1024);
  w->total = off;
}

int64_t train_workspace_bytes(const ddqst_dims* d, int64_t batch) {
  if (validate_dims(d) != DDQST_OK) return -1;
  TrainWs w;
  train_ws_layout(d, batch < 1 ? 1 : batch, &w);
  return w.total * 4 + 256;
}

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float dsilu_f(float v) { float s = sigmoid_f(v); return s * (1.0f + v * (1.0f - s)); }

// cond = [time_emb[t] || basis_emb[basis]]; x input: token embedding (B) or float bits (A)
__global__ void gather_inputs_kernel(int variant, int N, int E, const float* __restrict__ x_emb,
                                     const float* __restrict__ time_emb, const float* __restrict__ basis_emb,
                                     const uint16_t* __restrict__ xt, const int32_t* __restrict__ t,
                                     const int32_t* __restrict__ basis, float* __restrict__ cond, float* __restrict__ xin) {
  const int64_t i = blockIdx.x;
  const uint32_t bits = xt[i];
  const float* te = time_emb + (int64_t)t[i] * E;
  const float* be = basis_emb + (int64_t)basis[i] * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) { cond[i * 2 * E + e] = te[e]; cond[i * 2 * E + E + e] = be[e]; }
  if (variant == DDQST_VARIANT_B) {
    for (int j = threadIdx.x; j < N * E; j += blockDim.x) { int q = j / E, e = j - q * E; xin[i * N * E + j] = x_emb[((bits >> q) & 1u) * E + e]; }
  } else {
    for (int q = threadIdx.x; q < N; q += blockDim.x) xin[i * N + q] = (float)((bits >> q) & 1u);
  }
}

__global__ void film_apply_kernel(int64_t n, int H, const float* __restrict__ h, const float* __restrict__ gb, float* __restrict__ a) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int64_t i = e / H; int c = (int)(e - i * H);
  a[e] = h[e] * (1.0f + gb[i * 2 * H + c]) + gb[i * 2 * H + H + c];
}

// mean CE over B*N (RQC/main.py:110) and dlogits = (softmax - onehot) * scale / (B*N)
__global__ void ce_kernel(int64_t B, int N, const float* __restrict__ logits, const uint16_t* __restrict__ x0,
                          float scale, float* __restrict__ dlogits, float* __restrict__ loss_part) {
  __shared__ float red[256];
  float acc = 0.f;
  const int64_t total = B * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = e / N; int q = (int)(e - i * N);
    float l0 = logits[e * 2], l1 = logits[e * 2 + 1];
    float m = fmaxf(l0, l1), e0 = expf(l0 - m), e1 = expf(l1 - m), s = e0 + e1;
    uint32_t y = (x0[i] >> q) & 1u;
    acc += (m + logf(s)) - (y ? l1 : l0);
    float p0 = e0 / s, p1 = e1 / s, k = scale / (float)total;
    dlogits[e * 2] = (p0 - (y ? 0.f : 1.f)) * k;
    dlogits[e * 2 + 1] = (p1 - (y ? 1.f : 0.f)) * k;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) loss_part[blockIdx.x] = red[0];
}

__global__ void loss_finish_kernel(const float* __restrict__ part, int n, float inv_total, float* __restrict__ out) {
  // single thread, fixed order: deterministic
  if (threadIdx.x == 0 && blockIdx.x == 0) { double s = 0.0; for (int i = 0; i < n; ++i) s += part[i]; out[0] = (float)(s * inv_total); }
}

// dz = dy * silu'(z)   (in place on dy allowed)
__global__ void dsilu_kernel(int64_t n, const float* __restrict__ dy, const float* __restrict__ z, float* __restrict__ dz) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) dz[e] = dy[e] * dsilu_f(z[e]);
}

// FiLM backward: a = h*(1+g)+b  ->  dg = da*h, db = da, dh_acc += da*(1+g)
__global__ void film_backward_kernel(int64_t n, int H, const float* __restrict__ da, const float* __restrict__ h,
                                     const float* __restrict__ gb, float* __restrict__ dgb, float* __restrict__ dh_acc) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int64_t i = e / H; int c = (int)(e - i * H);
  float d = da[e];
  dgb[i * 2 * H + c] = d * h[e];
  dgb[i * 2 * H + H + c] = d;
  dh_acc[e] += d * (1.0f + gb[i * 2 * H + c]);
}

__global__ void colsum_kernel(const float* __restrict__ M, int64_t rows, int cols, int64_t ld, float* __restrict__ out, int accumulate) {
  // one block per column group of 32; fixed-order tree inside the block -> deterministic
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + (threadIdx.x & 31), r0 = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < cols) for (int64_t r = r0; r < rows; r += 8) acc += M[r * ld + c];
  red[r0][threadIdx.x & 31] = acc;
  __syncthreads();
  if (r0 == 0 && c < cols) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x & 31];
    out[c] = accumulate ? out[c] + s : s;
  }
}

// embedding gradients: rows of time_emb / basis_emb / x_emb receive sums over the samples that used them
__global__ void emb_scatter_kernel(int variant, int N, int E, const float* __restrict__ dcond, const float* __restrict__ dxin,
                                   const uint16_t* __restrict__ xt, const int32_t* __restrict__ t, const int32_t* __restrict__ basis,
                                   float* __restrict__ g_time, float* __restrict__ g_basis, float* __restrict__ g_xemb) {
  const int64_t i = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    atomicAdd(g_time + (int64_t)t[i] * E + e, dcond[i * 2 * E + e]);
    atomicAdd(g_basis + (int64_t)basis[i] * E + e, dcond[i * 2 * E + E + e]);
  }
  if (variant == DDQST_VARIANT_B) {
    const uint32_t bits = xt[i];
    for (int j = threadIdx.x; j < N * E; j += blockDim.x) { int q = j / E, e = j - q * E; atomicAdd(g_xemb + ((bits >> q) & 1u) * E + e, dxin[i * N * E + j]); }
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr, float b1, float b2, float eps, float wd, int decoupled, float gscale,
                            float bc1, float bc2_sqrt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float grad = g[i] * gscale, param = p[i];
  if (decoupled) param *= (1.0f - lr * wd);            // torch AdamW: param.mul_(1 - lr*wd)
  else if (wd != 0.f) grad += wd * param;              // torch Adam: grad.add(param, alpha=wd)
  float mi = m[i] + (grad - m[i]) * (1.0f - b1);       // exp_avg.lerp_(grad, 1-beta1)
  float vi = v[i] * b2 + (1.0f - b2) * grad * grad;    // exp_avg_sq.mul_(b2).addcmul_(grad, grad, 1-b2)
  m[i] = mi; v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = param - (lr / bc1) * (mi / denom);
}

static int ew_grid(int64_t n) { return (int)((n + 255) / 256); }

}  // namespace ddqst

using namespace ddqst;

namespace ddqst {

struct TrainCtx {
  const ddqst_dims* d; ParamLayout pr; TrainWs w; float* ws; int64_t B; cudaStream_t s;
};

static int train_ctx(TrainCtx* c, const ddqst_dims* d, int64_t batch, void* workspace, int64_t ws_bytes, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_TRY(param_layout(d, &c->pr));
  DDQST_REQUIRE(batch >= 1, DDQST_EINVAL_SHAPE, "batch=%lld", (long long)batch);
  train_ws_layout(d, batch, &c->w);
  DDQST_REQUIRE(workspace && ws_bytes >= c->w.total * 4, DDQST_EWORKSPACE, "train step needs %lld workspace bytes, got %lld",
                (long long)(c->w.total * 4), (long long)ws_bytes);
  c->d = d; c->ws = (float*)workspace; c->B = batch; c->s = (cudaStream_t)stream;
  return DDQST_OK;
}

// forward with saved activations (RQC/model.py:51-70); logits land in ws[w.logits]
static int forward_saved(const TrainCtx& c, const float* params, const uint16_t* xt_packed, const int32_t* t, const int32_t* basis) {
  const ddqst_dims* d = c.d; const ParamLayout& pr = c.pr; const TrainWs& w = c.w; float* ws = c.ws; cudaStream_t s = c.s;
  const int N = d->num_qubits, E = d->embed_dim, H = d->hidden_dim, L = d->num_blocks;
  const int64_t B = c.B;
  const int xin_dim = d->variant == DDQST_VARIANT_B ? N * E : N;
  float *cond = ws + w.cond, *xin = ws + w.xin, *hbuf = ws + w.h, *gb = ws + w.gb, *abuf = ws + w.a, *z1 = ws + w.z1,
        *ubuf = ws + w.u, *z2 = ws + w.z2, *logits = ws + w.logits;
  auto lin = [&](const float* A, int64_t lda, const float* W, int K, int Nout, const float* bias, float* C, int64_t ldc,
                 int epi, const float* R, float* aux) {
    GemmArgs g{};                                   // C = epi(A[B,K] . W[Nout,K]^T + bias)
    g.A = A; g.a_rs = lda; g.a_cs = 1; g.B = W; g.b_rs = 1; g.b_cs = K; g.C = C; g.ldc = ldc; g.bias = bias;
    g.R = R; g.ldr = H; g.aux = aux; g.ldaux = ldc; g.M = (int)B; g.N = Nout; g.K = K; g.epi = epi; g.alpha = 1.f;
    return launch_sgemm(g, s);
  };
  gather_inputs_kernel<<<(unsigned)B, 128, 0, s>>>(d->variant, N, E, pr.x_emb >= 0 ? params + pr.x_emb : nullptr,
                                                   params + pr.time_emb, params + pr.basis_emb, xt_packed, t, basis, cond, xin);
  DDQST_LAUNCH_OK();
  DDQST_TRY(lin(xin, xin_dim, params + pr.in_w, xin_dim, H, params + pr.in_b, hbuf, H, EPI_BIAS, nullptr, nullptr));
  for (int l = 0; l < L; ++l) {
    float* h_in = hbuf + (int64_t)l * B * H;
    float* h_out = hbuf + (int64_t)(l + 1) * B * H;
    float* gbl = gb + (int64_t)l * B * 2 * H;
    DDQST_TRY(lin(cond, 2 * E, params + pr.film_w[l], 2 * E, 2 * H, params + pr.film_b[l], gbl, 2 * H, EPI_BIAS, nullptr, nullptr));
    film_apply_kernel<<<ew_grid(B * H), 256, 0, s>>>(B * H, H, h_in, gbl, abuf + (int64_t)l * B * H);
    DDQST_LAUNCH_OK();
    DDQST_TRY(lin(abuf + (int64_t)l * B * H, H, params + pr.w1[l], H, H, params + pr.b1[l], ubuf + (int64_t)l * B * H, H,
                  EPI_BIAS_SILU, nullptr, z1 + (int64_t)l * B * H));
    DDQST_TRY(lin(ubuf + (int64_t)l * B * H, H, params + pr.w2[l], H, H, params + pr.b2[l], h_out, H, EPI_RES_SILU, h_in,
                  z2 + (int64_t)l * B * H));
  }
  DDQST_TRY(lin(hbuf + (int64_t)L * B * H, H, params + pr.head_w, H, 2 * N, params + pr.head_b, logits, 2 * N, EPI_BIAS, nullptr, nullptr));
  return DDQST_OK;
}

// backward from dlogits (ws[w.dlogits]) through the saved activations into grads (OVERWRITTEN)
static int backward_saved(const TrainCtx& c, const float* params, const uint16_t* xt_packed, const int32_t* t,
                          const int32_t* basis, float* grads) {
  const ddqst_dims* d = c.d; const ParamLayout& pr = c.pr; const TrainWs& w = c.w; float* ws = c.ws; cudaStream_t s = c.s;
  const int N = d->num_qubits, E = d->embed_dim, H = d->hidden_dim, L = d->num_blocks;
  const int64_t B = c.B;
  const int xin_dim = d->variant == DDQST_VARIANT_B ? N * E : N;
  float *cond = ws + w.cond, *xin = ws + w.xin, *hbuf = ws + w.h, *gb = ws + w.gb, *abuf = ws + w.a, *z1 = ws + w.z1,
        *ubuf = ws + w.u, *z2 = ws + w.z2, *dlogits = ws + w.dlogits, *dh = ws + w.dh,
        *dtmp = ws + w.dtmp, *du = ws + w.du, *dgb = ws + w.dgb, *dcond = ws + w.dcond, *dxin = ws + w.dxin;
  DDQST_CUDA_OK(cudaMemsetAsync(grads, 0, sizeof(float) * pr.total, s));
  auto wgrad = [&](const float* dY, int64_t ldy, int Nout, const float* X, int64_t ldx, int K, float* dW) {
    GemmArgs g{};                                   // dW[Nout,K] = dY[B,Nout]^T . X[B,K]
    g.A = dY; g.a_rs = 1; g.a_cs = ldy; g.B = X; g.b_rs = ldx; g.b_cs = 1; g.C = dW; g.ldc = K;
    g.M = Nout; g.N = K; g.K = (int)B; g.epi = EPI_NONE; g.alpha = 1.f;
    return launch_sgemm(g, s);
  };
  auto dgrad = [&](const float* dY, int64_t ldy, int Nout, const float* W, int K, float* dX, int64_t ldx, int epi) {
    GemmArgs g{};                                   // dX[B,K] (=|+=) dY[B,Nout] . W[Nout,K]
    g.A = dY; g.a_rs = ldy; g.a_cs = 1; g.B = W; g.b_rs = K; g.b_cs = 1; g.C = dX; g.ldc = ldx;
    g.M = (int)B; g.N = K; g.K = Nout; g.epi = epi; g.alpha = 1.f;
    return launch_sgemm(g, s);
  };
  auto bgrad = [&](const float* dY, int64_t ldy, int Nout, float* db) {
    colsum_kernel<<<(Nout + 31) / 32, 256, 0, s>>>(dY, B, Nout, ldy, db, 0);
    return cudaGetLastError() == cudaSuccess ? DDQST_OK : DDQST_ECUDA;
  };
  DDQST_TRY(wgrad(dlogits, 2 * N, 2 * N, hbuf + (int64_t)L * B * H, H, H, grads + pr.head_w));
  DDQST_TRY(bgrad(dlogits, 2 * N, 2 * N, grads + pr.head_b));
  DDQST_TRY(dgrad(dlogits, 2 * N, 2 * N, params + pr.head_w, H, dh, H, EPI_NONE));
  DDQST_CUDA_OK(cudaMemsetAsync(dcond, 0, sizeof(float) * B * 2 * E, s));
  for (int l = L - 1; l >= 0; --l) {
    const float* h_in = hbuf + (int64_t)l * B * H;
    const float* gbl = gb + (int64_t)l * B * 2 * H;
    // h_out = silu(z2), z2 = h_in + W2 u + b2
    dsilu_kernel<<<ew_grid(B * H), 256, 0, s>>>(B * H, dh, z2 + (int64_t)l * B * H, dh);      // dh := dz2 (also the residual grad)
    DDQST_LAUNCH_OK();
    DDQST_TRY(wgrad(dh, H, H, ubuf + (int64_t)l * B * H, H, H, grads + pr.w2[l]));
    DDQST_TRY(bgrad(dh, H, H, grads + pr.b2[l]));
    DDQST_TRY(dgrad(dh, H, H, params + pr.w2[l], H, du, H, EPI_NONE));
    dsilu_kernel<<<ew_grid(B * H), 256, 0, s>>>(B * H, du, z1 + (int64_t)l * B * H, du);      // du := dz1
    DDQST_LAUNCH_OK();
    DDQST_TRY(wgrad(du, H, H, abuf + (int64_t)l * B * H, H, H, grads + pr.w1[l]));
    DDQST_TRY(bgrad(du, H, H, grads + pr.b1[l]));
    DDQST_TRY(dgrad(du, H, H, params + pr.w1[l], H, dtmp, H, EPI_NONE));                      // dtmp := da
    film_backward_kernel<<<ew_grid(B * H), 256, 0, s>>>(B * H, H, dtmp, h_in, gbl, dgb, dh); // dh += da*(1+g)
    DDQST_LAUNCH_OK();
    DDQST_TRY(wgrad(dgb, 2 * H, 2 * H, cond, 2 * E, 2 * E, grads + pr.film_w[l]));
    DDQST_TRY(bgrad(dgb, 2 * H, 2 * H, grads + pr.film_b[l]));
    DDQST_TRY(dgrad(dgb, 2 * H, 2 * H, params + pr.film_w[l], 2 * E, dcond, 2 * E, EPI_ACCUM));
  }
  DDQST_TRY(wgrad(dh, H, H, xin, xin_dim, xin_dim, grads + pr.in_w));
  DDQST_TRY(bgrad(dh, H, H, grads + pr.in_b));
  if (d->variant == DDQST_VARIANT_B) DDQST_TRY(dgrad(dh, H, H, params + pr.in_w, xin_dim, dxin, xin_dim, EPI_NONE));
  emb_scatter_kernel<<<(unsigned)B, 128, 0, s>>>(d->variant, N, E, dcond, dxin, xt_packed, t, basis, grads + pr.time_emb,
                                                 grads + pr.basis_emb, pr.x_emb >= 0 ? grads + pr.x_emb : nullptr);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // namespace ddqst

using namespace ddqst;

extern "C" {

int ddqst_forward_saved(const ddqst_dims* d, const float* params, const uint16_t* xt_packed, const int32_t* t,
                        const int32_t* basis, int64_t batch, float* logits_out, void* workspace, int64_t ws_bytes, void* stream) {
  TrainCtx c;
  DDQST_TRY(train_ctx(&c, d, batch, workspace, ws_bytes, stream));
  DDQST_REQUIRE(params && xt_packed && t && basis, DDQST_EINVAL_SHAPE, "NULL argument");
  DDQST_TRY(forward_saved(c, params, xt_packed, t, basis));
  if (logits_out) DDQST_CUDA_OK(cudaMemcpyAsync(logits_out, c.ws + c.w.logits, sizeof(float) * batch * 2 * d->num_qubits, cudaMemcpyDeviceToDevice, c.s));
  return DDQST_OK;
}

int ddqst_backward_saved(const ddqst_dims* d, const float* params, const uint16_t* xt_packed, const int32_t* t,
                         const int32_t* basis, int64_t batch, const float* dlogits, float* grads, void* workspace,
                         int64_t ws_bytes, void* stream) {
  TrainCtx c;
  DDQST_TRY(train_ctx(&c, d, batch, workspace, ws_bytes, stream));
  DDQST_REQUIRE(params && xt_packed && t && basis && dlogits && grads, DDQST_EINVAL_SHAPE, "NULL argument");
  DDQST_CUDA_OK(cudaMemcpyAsync(c.ws + c.w.dlogits, dlogits, sizeof(float) * batch * 2 * d->num_qubits, cudaMemcpyDeviceToDevice, c.s));
  return backward_saved(c, params, xt_packed, t, basis, grads);
}

int ddqst_train_forward_backward(const ddqst_dims* d, const float* params, const uint16_t* xt_packed,
                                 const uint16_t* x0_packed, const int32_t* t, const int32_t* basis, int64_t batch,
                                 float loss_scale, float* grads, float* loss_out, void* workspace, int64_t ws_bytes,
                                 void* stream) {
  TrainCtx c;
  DDQST_TRY(train_ctx(&c, d, batch, workspace, ws_bytes, stream));
  DDQST_REQUIRE(params && xt_packed && x0_packed && t && basis && grads && loss_out, DDQST_EINVAL_SHAPE, "NULL argument");
  DDQST_TRY(forward_saved(c, params, xt_packed, t, basis));
  const int N = d->num_qubits;
  const int64_t B = batch;
  const int ce_blocks = (int)((B * N + 255) / 256 > 1024 ? 1024 : (B * N + 255) / 256);
  ce_kernel<<<ce_blocks, 256, 0, c.s>>>(B, N, c.ws + c.w.logits, x0_packed, loss_scale, c.ws + c.w.dlogits, c.ws + c.w.loss_part);
  DDQST_LAUNCH_OK();
  loss_finish_kernel<<<1, 32, 0, c.s>>>(c.ws + c.w.loss_part, ce_blocks, 1.0f / (float)(B * N), loss_out);
  DDQST_LAUNCH_OK();
  return backward_saved(c, params, xt_packed, t, basis, grads);
}

int ddqst_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, float grad_scale,
                    void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(n >= 0 && step >= 1, DDQST_EINVAL_SHAPE, "n=%lld step=%lld", (long long)n, (long long)step);
  if (n == 0) return DDQST_OK;
  DDQST_REQUIRE(params && grads && exp_avg && exp_avg_sq, DDQST_EINVAL_SHAPE, "NULL argument");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                            weight_decay, decoupled, grad_scale, (float)bc1, (float)sqrt(bc2));
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // extern "C"
