// Mixed-precision front end of the Hermitian eigensolver (included by recon.cu inside namespace ddqst, after the fp64 odd-even
// kernel whose machinery -- register-resident column pairs, st.async inboxes, mbarrier hand-over -- it reuses).
//
// Cyclic Jacobi on a linear-inversion rho spends sweeps 2..7 in a slow, roughly halving phase while the ~2^N clustered noise
// eigenvalues separate (benchmarks/jacobi_sweeps.py); only the last two sweeps need fp64.  So:
//   1. the sweeps run in FP32 first (half the exchanged bytes, twice the FMA rate) until a sweep starts below ~3e-5 relative
//      off-diagonal (fp32 cannot resolve much further),
//   2. the fp32 eigenvector estimate V is orthonormalised in fp64 by one Newton-Schulz step  V <- (3 I - V V^H) V / 2
//      (orthogonality error 1e-6 -> 1e-12),
//   3. G = A' V is formed in fp64 (columns nearly orthogonal already) and the fp64 kernel finishes: 2-3 sweeps instead of 8-9.
// The result is a plain fp64 one-sided Jacobi solution -- G = A' V W with V W unitary to 1e-12 -- so accuracy is unchanged.
#pragma once

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ void jr_send_f(uint32_t raddr, float2 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "f"(v.x), "f"(v.y), "r"(rbar) : "memory");
}

__device__ __forceinline__ void jr_send_f4(uint32_t raddr, float2 u, float2 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(raddr), "f"(u.x), "f"(u.y), "f"(v.x), "f"(v.y), "r"(rbar) : "memory");
}
// element e of a lane's column slice sits at row (e >> 1) * 64 + 2 * lane + (e & 1): rows come in adjacent PAIRS, so a lane moves
// 16 bytes per st.async / LDS / LDG (half the instructions of the 8-byte form; the st.async issue rate bounds the exchange)
__device__ __forceinline__ int jr_row_f(int lane, int e) { return (e >> 1) * 64 + 2 * lane + (e & 1); }

template <int EPL>
__global__ void __launch_bounds__(kJcThreads) jacobi_oddeven_f32_kernel(float2* __restrict__ GT, int n, int max_sweeps, float tol,
                                                                        JacobiCtl* ctl) {
  extern __shared__ __align__(16) uint8_t jr_smem[];
  constexpr int COLB = EPL * 32 * 8;
  constexpr int WPC = kJcThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int h = n / 2, k = (int)crank * WPC + warp;
  const bool active = k < h;
  const uint32_t smem0 = smem_u32(jr_smem);
  const uint32_t bars0 = smem0 + WPC * 3 * COLB;
  auto inbox = [&](int w, int par) { return smem0 + (uint32_t)((w * 3 + par) * COLB); };
  auto bar = [&](int w, int par) { return bars0 + (uint32_t)((w * 2 + par) * 8); };
  uint8_t* park = jr_smem + (warp * 3 + 2) * COLB;
  const uint32_t colbytes = (uint32_t)n * 8u;
  const bool recv_right = active && k <= h - 2;
  const bool recv_left = active && k >= 1;
  if (lane == 0) {
    mbar_init(bar(warp, 0), 1); mbar_init(bar(warp, 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (recv_left) mbar_expect_tx(bar(warp, 0), colbytes);
    if (recv_right) mbar_expect_tx(bar(warp, 1), colbytes);
  }
  float2 a[EPL], b[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int i = jr_row_f(lane, e);
    const bool ok = active && i < n;
    a[e] = ok ? GT[(int64_t)(2 * k) * n + i] : make_float2(0.f, 0.f);
    b[e] = ok ? GT[(int64_t)(2 * k + 1) * n + i] : make_float2(0.f, 0.f);
  }
  uint32_t left_addr = 0, left_bar = 0, right_addr = 0, right_bar = 0;
  if (recv_left) {
    left_addr = jr_mapa(inbox((k - 1) % WPC, 1), (uint32_t)((k - 1) / WPC));
    left_bar = jr_mapa(bar((k - 1) % WPC, 1), (uint32_t)((k - 1) / WPC));
  }
  if (recv_right) {
    right_addr = jr_mapa(inbox((k + 1) % WPC, 0), (uint32_t)((k + 1) / WPC));
    right_bar = jr_mapa(bar((k + 1) % WPC, 0), (uint32_t)((k + 1) / WPC));
  }
  __syncwarp();
  jc_cluster_barrier();
  int sweep = 0;
  uint32_t g = 0;
  for (; sweep < max_sweeps; ++sweep) {
    int rot = 0;
    float worst = 0.f;
    if (active) {
      for (int step = 0; step < n; ++step, ++g) {
        const bool odd = (g & 1u) != 0u;
        if (!odd || k <= h - 2) {
          float sa = 0.f, sb = 0.f, gr = 0.f, gi = 0.f;
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            sa += a[e].x * a[e].x + a[e].y * a[e].y;
            sb += b[e].x * b[e].x + b[e].y * b[e].y;
            gr += a[e].x * b[e].x + a[e].y * b[e].y;
            gi += a[e].x * b[e].y - a[e].y * b[e].x;
          }
          sa = warp_sum_f(sa); sb = warp_sum_f(sb); gr = warp_sum_f(gr); gi = warp_sum_f(gi);
          const float g2 = gr * gr + gi * gi;
          worst = fmaxf(worst, g2 / (sa * sb));
          float c = 1.f, s = 0.f, pr = 1.f, pi = 0.f;
          const bool on = g2 > tol * tol * sa * sb && g2 > 1e-30f;
          if (on) {
            const float inv_g = rsqrtf(g2);
            const float zeta = 0.5f * (sb - sa) * inv_g, az = fabsf(zeta);
            const float at = az < 1e4f ? __frcp_rn(az + sqrtf(fmaf(az, az, 1.f))) : 0.5f * __frcp_rn(az);
            const float t = zeta >= 0.f ? at : -at;
            c = rsqrtf(fmaf(t, t, 1.f)); s = c * t;
            pr = gr * inv_g; pi = -gi * inv_g;
            ++rot;
          }
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            const float2 a0 = a[e], b0 = b[e];
            if (on) {
              const float2 br = make_float2(b0.x * pr - b0.y * pi, b0.x * pi + b0.y * pr);
              a[e] = make_float2(s * a0.x + c * br.x, s * a0.y + c * br.y);
              b[e] = make_float2(c * a0.x - s * br.x, c * a0.y - s * br.y);
            } else {
              a[e] = b0; b[e] = a0;
            }
          }
        }
        if (!odd) {
          if (k >= 1) {
#pragma unroll
            for (int e = 0; e < EPL; e += 2) { const int i = jr_row_f(lane, e); if (i < n) jr_send_f4(left_addr + (uint32_t)i * 8u, a[e], a[e + 1], left_bar); }
          } else {
#pragma unroll
            for (int e = 0; e < EPL; e += 2) { const int i = jr_row_f(lane, e); if (i < n) *reinterpret_cast<float4*>(park + i * 8) = make_float4(a[e].x, a[e].y, a[e + 1].x, a[e + 1].y); }
          }
#pragma unroll
          for (int e = 0; e < EPL; ++e) a[e] = b[e];
          if (recv_right) {
            mbar_wait(bar(warp, 1), (g >> 1) & 1u, 64);
            const uint8_t* in = jr_smem + (warp * 3 + 1) * COLB;
#pragma unroll
            for (int e = 0; e < EPL; e += 2) { const int i = jr_row_f(lane, e); if (i < n) { const float4 v = *reinterpret_cast<const float4*>(in + i * 8); b[e] = make_float2(v.x, v.y); b[e + 1] = make_float2(v.z, v.w); } }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(bar(warp, 1), colbytes);
          }
        } else {
          if (k <= h - 2) {
#pragma unroll
            for (int e = 0; e < EPL; e += 2) { const int i = jr_row_f(lane, e); if (i < n) jr_send_f4(right_addr + (uint32_t)i * 8u, b[e], b[e + 1], right_bar); }
          }
#pragma unroll
          for (int e = 0; e < EPL; ++e) b[e] = a[e];
          if (recv_left) {
            mbar_wait(bar(warp, 0), (g >> 1) & 1u, 65);
            const uint8_t* in = jr_smem + (warp * 3 + 0) * COLB;
#pragma unroll
            for (int e = 0; e < EPL; e += 2) { const int i = jr_row_f(lane, e); if (i < n) { const float4 v = *reinterpret_cast<const float4*>(in + i * 8); a[e] = make_float2(v.x, v.y); a[e + 1] = make_float2(v.z, v.w); } }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(bar(warp, 0), colbytes);
          } else {
            __syncwarp();
#pragma unroll
            for (int e = 0; e < EPL; e += 2) { const int i = jr_row_f(lane, e); if (i < n) { const float4 v = *reinterpret_cast<const float4*>(park + i * 8); a[e] = make_float2(v.x, v.y); a[e + 1] = make_float2(v.z, v.w); } }
          }
        }
      }
      if (lane == 0 && rot > 0) atomicAdd(&ctl->rotations[sweep], rot);
      if (lane == 0 && sweep < 48) atomicMax(&ctl->max_ratio2[sweep], __float_as_uint(worst));
    }
    jc_cluster_barrier();
    const int total = __ldcg(&ctl->rotations[sweep]);
    if (total == 0) { ++sweep; break; }
    if (sweep < 48 && __uint_as_float(__ldcg(&ctl->max_ratio2[sweep])) < __ldcg(&ctl->stop_ratio2)) { ++sweep; break; }
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int i = jr_row_f(lane, e);
      if (i < n) { GT[(int64_t)(2 * k) * n + i] = a[e]; GT[(int64_t)(2 * k + 1) * n + i] = b[e]; }
    }
  }
  if (k == 0 && lane == 0) ctl->sweeps_done = sweep;
  jc_cluster_barrier();
}

template <int EPL>
static int launch_jacobi_oddeven_f32(float2* GT, int n, int max_sweeps, float tol, JacobiCtl* ctl, cudaStream_t s, bool* launched) {
  constexpr int smem = (kJcThreads / 32) * 3 * EPL * 32 * 8 + 256;
  *launched = false;
  static bool attr_set = false;
  if (!attr_set) {
    DDQST_CUDA_OK(cudaFuncSetAttribute(jacobi_oddeven_f32_kernel<EPL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    DDQST_CUDA_OK(cudaFuncSetAttribute(jacobi_oddeven_f32_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  int csize = (n / 2 + kJcThreads / 32 - 1) / (kJcThreads / 32);
  if (csize < 1) csize = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)csize);
  cfg.blockDim = dim3(kJcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int fits = 0;
  if (cudaOccupancyMaxActiveClusters(&fits, jacobi_oddeven_f32_kernel<EPL>, &cfg) != cudaSuccess || fits < 1) {
    (void)cudaGetLastError();
    return DDQST_OK;
  }
  DDQST_CUDA_OK(cudaLaunchKernelEx(&cfg, jacobi_oddeven_f32_kernel<EPL>, GT, n, max_sweeps, tol, ctl));
  *launched = true;
  return DDQST_OK;
}

__global__ void eig_set_stop_kernel(JacobiCtl* ctl, float stop_ratio2) { if (threadIdx.x == 0) ctl->stop_ratio2 = stop_ratio2; }

// ---- glue kernels (all n x n, row-major; R holds the eigenvector estimates as ROWS, like VT / GT)
__global__ void eig_to_f32_kernel(const double2* __restrict__ src, float2* __restrict__ dst, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < total) dst[e] = make_float2((float)src[e].x, (float)src[e].y);
}
// R[j][:] = GT32[j][:] / ||GT32[j][:]||  (fp64 out); one block per row; also re-zeroes the control block for the fp64 phase
__global__ void eig_normalise_rows_kernel(const float2* __restrict__ G32, int n, double2* __restrict__ R, JacobiCtl* ctl, float stop_ratio2) {
  __shared__ double scratch[96];
  const int j = blockIdx.x;
  double a = 0.0, z1 = 0.0, z2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float2 v = G32[(int64_t)j * n + i]; a += (double)v.x * v.x + (double)v.y * v.y; }
  block_sum3(a, z1, z2, scratch);
  const double inv = a > 0.0 ? 1.0 / sqrt(a) : 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float2 v = G32[(int64_t)j * n + i]; R[(int64_t)j * n + i] = make_double2(v.x * inv, v.y * inv); }
  if (j == 0 && threadIdx.x == 0) {
    ctl->f32_sweeps = ctl->sweeps_done;
    ctl->f32_last_ratio2 = ctl->sweeps_done >= 1 && ctl->sweeps_done <= 48 ? __uint_as_float(ctl->max_ratio2[ctl->sweeps_done - 1]) : 0.f;
    ctl->sweeps_done = 0;
    ctl->stop_ratio2 = stop_ratio2;
    for (int i = 0; i < 64; ++i) ctl->rotations[i] = 0;
    for (int i = 0; i < 48; ++i) ctl->max_ratio2[i] = 0u;
  }
}
// OP 0: C = A . B^H            (M = R R^H)
// OP 1: C = 1.5 A2 - 0.5 A . B (R' = 1.5 R - 0.5 M R ; A = M, B = R, A2 = R)
// OP 2: C[j][i] = sum_k A[j][k] conj(H[k][i]) + sigma A[j][i]   (GT = R conj(A') with A' = H + sigma I; H Hermitian, symmetrised as in init)
// OP 4: C = A . B^T            (no conjugate; the change of basis of the mixed-state fidelity)
template <int OP>
__global__ void eig_zgemm_kernel(const double2* __restrict__ A, const double2* __restrict__ B, const double2* __restrict__ A2, int n,
                                 const JacobiCtl* ctl, double2* __restrict__ C) {
  __shared__ double2 ta[16][17], tb[16][17];
  const int r = blockIdx.y * 16 + threadIdx.y, c = blockIdx.x * 16 + threadIdx.x;
  double ax = 0.0, ay = 0.0;
  for (int k0 = 0; k0 < n; k0 += 16) {
    const int ka = k0 + threadIdx.x, kb = k0 + threadIdx.y;
    ta[threadIdx.y][threadIdx.x] = (r < n && ka < n) ? A[(int64_t)r * n + ka] : make_double2(0, 0);
    double2 bv = make_double2(0, 0);
    if (OP == 0 || OP == 4) {                         // B^H[k][c] = conj(B[c][k]): read B[c0+ty... coalesced along k
      const int cc = blockIdx.x * 16 + threadIdx.y, kk = k0 + threadIdx.x;
      if (cc < n && kk < n) { const double2 t = B[(int64_t)cc * n + kk]; bv = make_double2(t.x, OP == 0 ? -t.y : t.y); }
      tb[threadIdx.x][threadIdx.y] = bv;              // tb[k][c]
    } else if (OP == 1) {
      if (kb < n && c < n) bv = B[(int64_t)kb * n + c];
      tb[threadIdx.y][threadIdx.x] = bv;
    } else {
      if (kb < n && c < n) {                          // conj of the symmetrised Hermitian input: 0.5 (conj(H[k][c]) + H[c][k])
        const double2 u = B[(int64_t)kb * n + c], w = B[(int64_t)c * n + kb];
        bv = make_double2(0.5 * (u.x + w.x), 0.5 * (-u.y + w.y));
        if (kb == c) bv.y = 0.0;
      }
      tb[threadIdx.y][threadIdx.x] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const double2 a = ta[threadIdx.y][k], b = tb[k][threadIdx.x];
      ax += a.x * b.x - a.y * b.y;
      ay += a.x * b.y + a.y * b.x;
    }
    __syncthreads();
  }
  if (r < n && c < n) {
    const int64_t e = (int64_t)r * n + c;
    if (OP == 0 || OP == 4) C[e] = make_double2(ax, ay);
    else if (OP == 1) { const double2 v = A2[e]; C[e] = make_double2(1.5 * v.x - 0.5 * ax, 1.5 * v.y - 0.5 * ay); }
    else { const double2 v = A[e]; const double sg = ctl->sigma; C[e] = make_double2(ax + sg * v.x, ay + sg * v.y); }
  }
}

// M[j][k] *= sqrt(max(l_j, 0) max(l_k, 0)): rows / columns of clipped eigenvalues become EXACT zeros (see fidelity_in_eigenbasis)
__global__ void eig_scale_sqrt_kernel(double2* __restrict__ M, const double* __restrict__ ev, int n) {
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e / n), k = (int)(e % n);
    const double lj = ev[j], lk = ev[k];
    const double f = (lj > 0.0 && lk > 0.0) ? sqrt(lj * lk) : 0.0;
    const double2 v = M[e];
    M[e] = f > 0.0 ? make_double2(f * v.x, f * v.y) : make_double2(0.0, 0.0);
  }
}
