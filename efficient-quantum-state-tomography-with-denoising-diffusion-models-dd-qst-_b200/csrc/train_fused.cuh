// Fused forward + data-gradient pass of the training step (T1, RQC/main.py:105-113) -- included by train_tc.cu inside
// namespace ddqst (uses its make_map3 and the PTX wrappers of tc_ptx.cuh).
//
// ONE persistent launch replaces the 19 dependent GEMM launches of the per-layer path (input projection, 2L forward GEMMs,
// head + cross-entropy, head data gradient, 2L backward GEMMs).  It has the structure of the reverse sampler
// (sampler_pair.cuh): a cluster of two CTAs owns 2 x 128 batch rows and walks the whole network forwards and then backwards
// without leaving the SM pair --
//
//   MMA        tcgen05.mma.cta_group::2.kind::f16, M = 256 (128 rows per CTA), N = 128, bf16 x bf16 -> fp32 in TMEM (512 columns =
//              the 128 x H accumulator).  Forward GEMMs read W[out,in] K-major; the data-gradient GEMMs read the SAME bf16
//              shadow weights as MN-major operands ([k = out rows][n = in columns], one 128-byte-swizzled 64-wide group per CTA),
//              so no transposed copy exists.  Weight tiles stream from L2 by TMA through a 4 x 16 KB ring per CTA.
//   A operand  the activation / gradient tile, written by the 16 epilogue warps straight into shared memory in the UMMA
//              K-major 128B-swizzled layout; the next GEMM starts on chunk n as soon as chunk n is published (chunk
//              pipelining, see sampler_pair.cuh), so the tensor pipe works under the epilogue.
//   epilogues  forward: FiLM GEMM + bias -> gamma|beta, bias, FiLM h(1+gamma)+beta, SiLU, residual, head: softmax
//              cross-entropy + dlogits;  backward: silu'(z1), silu'(s), FiLM backward (dgamma = da*h, dbeta = da,
//              dh += da(1+gamma)); the residual stream / residual gradient are re-read from tile-private saves, not kept in registers.
//
// What leaves the SM pair, once, as bf16 (the operands of the weight-gradient GEMMs, which reduce over the batch and
// therefore stay a separate grouped launch): a_l, u_l, h_L, dz1_l, dz2_l, dgamma|dbeta_l, dh_0, dlogits; plus what the
// backward half re-reads (silu'(z1_l), h_{l+1}, silu'(s_l), h_0, gamma|beta -- in a tile-private coalesced layout).  No fp32 activation
// round trips.
//
// GEMM sequence per tile (6L+3 with the FiLM phase in-kernel, else 4L+3): [Wfilm_l gamma half, beta half] l=0..L-1 (K = 2E, A = the gathered
// cond tile), input (K=32 hi/lo table), [W1_l, W2_l] l=0..L-1, head, head^T, [W2_l^T, W1_l^T] l=L-1..0.
// Debug aids: DDQST_FT_DEBUG bit 0 = clock stamps (ddqst_debug_ft_stamps), bits 1-4 switch off the bulk stores / tile-private stores /
// per-row dgamma|dbeta stores / tile-private loads (timing ablations only: results are wrong with them set).
#pragma once

// kFtEpiWarps / kFtEpiThreads / kFtColsPerWarp / kFtBatches / kFtThreads: see train_tc.cu (the FiLM GEMM epilogue writes the
// tile-private layout too and needs them)
constexpr int kFtRing = 4;
constexpr int kFtStage = 16384;

// clock64 stamps of CTA 0, first tile (DDQST_FT_DEBUG=1): [2g], [2g+1] = epilogue sweep g got its first accumulator chunk / finished;
// [100+2g], [101+2g] = MMA warp issued the first / committed the last MMA of GEMM g.  Read with ddqst_debug_ft_stamps().
static __device__ long long g_ft_dbg[256];

struct FusedParams {
  int N, L, head_pad, iters, dbg;
  int64_t B, n_tiles;
  const uint16_t* xt;            // noised bits x_t [B]
  const uint16_t* x0;            // clean bits [B] (cross-entropy targets)
  __nv_bfloat16* gb;             // [L][2][tile-private] gamma_l, beta_l (bias folded): written by the FiLM GEMM launch, or -- film_kc > 0 --
                                 // by this kernel itself (FiLM phase at the start of every tile)
  int film_kc, E;                // FiLM in-kernel: number of 128-column K chunks of cond = [t_emb | b_emb] (2E / 128), embed dim
  const int32_t* t; const int32_t* basis;                   // [B] timestep / basis index of every row
  const __nv_bfloat16 *temb, *bemb;                         // bf16 shadow of time_emb [T+1, E], basis_emb [num_bases, E]
  const float* film_b; int64_t film_b_stride;               // fp32 film.net bias of block l: film_b + l * film_b_stride, [2H]
  const float* b1; const float* b2; int64_t bias_stride;   // fp32 parameters: b1 + l * bias_stride
  const float* head_b;
  float scale;                   // loss_scale / (B * N)
  __nv_bfloat16 *dgb, *dlog;     // row-major [L][B][2H] (dgamma | dbeta) and [B][32]: stored by the threads
  // row-major act / hL / dz / dh0 (TcWs arrays, inputs of the weight-gradient launch) are written by TMA through the maps
  __nv_bfloat16 *d1s, *hs, *dsv, *h0s, *dsp;   // tile-private saves: silu'(z1_l) [L], h_{l+1} [L], silu'(s_l) [L], h_0, residual gradient
  int64_t priv_elems;            // elements of one tile-private array = padded tiles * 128 * H
  float* loss_part;              // [n_tiles * 4] per-warp partial sums of the per-row cross-entropy
};

template <int H>
__host__ __device__ constexpr int ft_smem_bytes(int L) {
  return 1024 + (H / 64) * 16384 + kFtRing * kFtStage + 2 * L * H * 4 + 256;
}

__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// f(n, k, wait_chunk or -1, commit_chunk or -1) in the issue order of the chunk-pipelined GEMM (see sampler_pair.cuh)
template <typename F>
__device__ __forceinline__ void ft_for_each_item(int NCH, F f) {
  for (int c = 0; c < NCH - 1; ++c) {
    bool first = true;
    for (int n = 0; n < c; ++n) { f(n, c, first ? c : -1, -1); first = false; }
    for (int k = 0; k <= c; ++k) { f(c, k, first ? c : -1, -1); first = false; }
  }
  const int c = NCH - 1;
  bool first = true;
  for (int k = 0; k < c; ++k) { f(c, k, first ? c : -1, -1); first = false; }
  for (int n = 0; n < c; ++n) { f(n, c, first ? c : -1, n); first = false; }
  f(c, c, first ? c : -1, c);
}

__device__ __forceinline__ void ft_ldg16(const __nv_bfloat16* p, bool ok, uint32_t (&v)[8]) {
  if (ok) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0u;
  }
}
__device__ __forceinline__ void ft_stg16(__nv_bfloat16* p, bool ok, const uint32_t (&o)[8]) {
  if (ok) {
    reinterpret_cast<uint4*>(p)[0] = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}
// one 32-byte store per thread (256-bit st.global, sm_100): half the L1TEX tag cycles of two 16-byte stores for the per-row arrays
__device__ __forceinline__ void ft_stg32(__nv_bfloat16* p, bool ok, const uint32_t (&o)[8]) {
  if (ok)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
}
// sigmoid(z) from one tanh.approx: 0.5 + 0.5 tanh(z/2)
__device__ __forceinline__ float ft_sigmoid(float z) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
  return fmaf(0.5f, th, 0.5f);
}
__device__ __forceinline__ float ft_dsilu(float z) { const float s = ft_sigmoid(z); return s * fmaf(z, 1.0f - s, 1.0f); }

template <int H>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFtThreads, 1)
train_fused_kernel(const __grid_constant__ CUtensorMap map_dt, const __grid_constant__ CUtensorMap map_w1k,
                   const __grid_constant__ CUtensorMap map_w2k, const __grid_constant__ CUtensorMap map_w1m,
                   const __grid_constant__ CUtensorMap map_w2m, const __grid_constant__ CUtensorMap map_headk,
                   const __grid_constant__ CUtensorMap map_headm, const __grid_constant__ CUtensorMap map_wf,
                   const __grid_constant__ CUtensorMap map_act,
                   const __grid_constant__ CUtensorMap map_hL, const __grid_constant__ CUtensorMap map_dz,
                   const __grid_constant__ CUtensorMap map_dh0, const FusedParams P) {
  constexpr int NCH = H / 128;
  constexpr int A_BYTES = (H / 64) * 16384;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) atomicCAS(&g_tc_abort, 0, 98);
  uint8_t* sA = smem;
  uint8_t* sRing = sA + A_BYTES;
  float* sB1 = (float*)(sRing + kFtRing * kFtStage);
  float* sB2 = sB1 + P.L * H;
  uint64_t* bars = (uint64_t*)(sB2 + P.L * H);
  // bars: [0..3] full (leader's used), [4..7] empty, [8..11] acc_done, [12..15] ready (leader's used)
  uint32_t* tmem_slot = (uint32_t*)(bars + 16);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 4), bar_acc = smem_u32(bars + 8),
                 bar_ready = smem_u32(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = P.L, N = P.N;
  const uint32_t crank = cluster_ctarank();
  const int hp2 = P.head_pad / 2;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_full + 8 * i, 2);
      mbar_init(bar_empty + 8 * i, 1);
      mbar_init(bar_acc + 8 * i, 1);
      mbar_init(bar_ready + 8 * i, 2 * kFtEpiWarps);    // epilogue warps x 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kFtEpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid < kFtEpiThreads) {
    for (int i = tid; i < L * H; i += kFtEpiThreads) {
      const int l = i / H, j = i - l * H;
      sB1[i] = P.b1[(int64_t)l * P.bias_stride + j];
      sB2[i] = P.b2[(int64_t)l * P.bias_stride + j];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *((volatile uint32_t*)tmem_slot);

  if (warp >= kFtEpiWarps) {
    if (warp == kFtEpiWarps) {
      // =============================== TMA producer (both CTAs: own half of every B tile) ===============================
      const uint32_t elected = elect_one();
      if (elected) {
        tma_prefetch_desc(&map_dt); tma_prefetch_desc(&map_w1k); tma_prefetch_desc(&map_w2k); tma_prefetch_desc(&map_w1m);
        tma_prefetch_desc(&map_w2m); tma_prefetch_desc(&map_headk); tma_prefetch_desc(&map_headm);
      }
      uint32_t cnt = 0;
      auto acquire = [&](uint32_t bytes_both) -> uint32_t {
        uint32_t s = cnt % kFtRing, ph = (cnt / kFtRing) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u, 51);
        if (elected) {
          if (crank == 0) mbar_expect_tx(bar_full + 8 * s, bytes_both);
          else mbar_arrive_leader(bar_full + 8 * s);
        }
        ++cnt;
        return s;
      };
      for (int it = 0; it < P.iters; ++it) {
        if (P.film_kc > 0) {                                              // FiLM phase: gamma_l (rows 0..H of Wfilm_l), beta_l (rows H..2H)
          for (int g = 0; g < 2 * L; ++g)
            for (int n = 0; n < NCH; ++n)
              for (int kc = 0; kc < P.film_kc; ++kc) {
                uint32_t s = acquire(2 * 16384);
                uint32_t dst = smem_u32(sRing + s * kFtStage);
                const int row = (g & 1) * H + n * 128 + (int)crank * 64;
                if (elected) {
                  tma_load_3d_2sm(dst, &map_wf, bar_full + 8 * s, (2 * kc) * 64, row, g >> 1);
                  tma_load_3d_2sm(dst + 8192, &map_wf, bar_full + 8 * s, (2 * kc + 1) * 64, row, g >> 1);
                }
              }
        }
        for (int n = 0; n < NCH; ++n) {                                   // input table, K block 0 only
          uint32_t s = acquire(2 * 8192);
          if (elected) tma_load_3d_2sm(smem_u32(sRing + s * kFtStage), &map_dt, bar_full + 8 * s, 0, n * 128 + (int)crank * 64, 0);
        }
        for (int g = 0; g < 2 * L; ++g) {                                 // forward: K-major tiles of W1_l / W2_l
          const CUtensorMap* mp = (g & 1) ? &map_w2k : &map_w1k;
          const int l = g >> 1;
          ft_for_each_item(NCH, [&](int n, int k, int, int) {
            uint32_t s = acquire(2 * 16384);
            uint32_t dst = smem_u32(sRing + s * kFtStage);
            const int row = n * 128 + (int)crank * 64;
            if (elected) {
              tma_load_3d_2sm(dst, mp, bar_full + 8 * s, (2 * k) * 64, row, l);
              tma_load_3d_2sm(dst + 8192, mp, bar_full + 8 * s, (2 * k + 1) * 64, row, l);
            }
          });
        }
        for (int k = 0; k < NCH; ++k) {                                   // head, K chunk k
          uint32_t s = acquire((uint32_t)(2 * 2 * hp2 * 128));
          uint32_t dst = smem_u32(sRing + s * kFtStage);
          if (elected) {
            tma_load_3d_2sm(dst, &map_headk, bar_full + 8 * s, (2 * k) * 64, (int)crank * hp2, 0);
            tma_load_3d_2sm(dst + hp2 * 128, &map_headk, bar_full + 8 * s, (2 * k + 1) * 64, (int)crank * hp2, 0);
          }
        }
        for (int n = 0; n < NCH; ++n) {                                   // head^T: [k = class rows][n = 64 hidden columns]
          uint32_t s = acquire((uint32_t)(2 * P.head_pad * 128));
          if (elected) tma_load_3d_2sm(smem_u32(sRing + s * kFtStage), &map_headm, bar_full + 8 * s, n * 128 + (int)crank * 64, 0, 0);
        }
        for (int g = 2 * L - 1; g >= 0; --g) {                            // backward: MN-major tiles of W2_l, W1_l
          const CUtensorMap* mp = (g & 1) ? &map_w2m : &map_w1m;
          const int l = g >> 1;
          ft_for_each_item(NCH, [&](int n, int k, int, int) {
            uint32_t s = acquire(2 * 16384);
            if (elected) tma_load_3d_2sm(smem_u32(sRing + s * kFtStage), mp, bar_full + 8 * s, n * 128 + (int)crank * 64, k * 128, l);
          });
        }
      }
    } else if (warp == kFtEpiWarps + 1 && crank == 0) {
      // =============================== MMA issuer (leader CTA only) ===============================
      const uint32_t elected = elect_one();
      uint32_t cnt = 0, slot = 0;
      const bool mstamp = (P.dbg & 1) && blockIdx.x == 0 && lane == 0;
      const uint32_t idesc = umma_idesc_bf16_m(256, 128), idesc_head = umma_idesc_bf16_m(256, P.head_pad);
      const uint32_t idesc_mn = idesc | (1u << 16);                       // B operand MN-major
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sRing));
      const uint64_t b_desc0_mn = umma_desc_mn_sw128(smem_u32(sRing), 16384);
      auto stage_wait = [&]() -> uint32_t {
        uint32_t s = cnt % kFtRing, ph = (cnt / kFtRing) & 1u;
        mbar_wait_cluster(bar_full + 8 * s, ph, 53);
        tc_fence_after();
        return s;
      };
      auto stage_release = [&](uint32_t s) {
        if (elected) umma2_commit_mc(bar_empty + 8 * s, 3);
        ++cnt;
      };
      auto commit_all = [&]() {
        if (elected)
          for (int n = 0; n < NCH; ++n) umma2_commit_mc(bar_acc + 8 * n, 3);
        __syncwarp();
        if (mstamp) g_ft_dbg[101 + 2 * (slot % 50)] = clock64();
        ++slot;
      };
      for (int it = 0; it < P.iters; ++it) {
        // ---- FiLM phase: 2L GEMMs [256 x H] = cond[256 x 2E] . Wfilm_l[half]^T; chunk n may be overwritten once ready[n] says the
        //      previous stage's epilogue has read it
        if (P.film_kc > 0) {
          for (int g = 0; g < 2 * L; ++g) {
            for (int n = 0; n < NCH; ++n) {
              mbar_wait_cluster(bar_ready + 8 * n, slot & 1u, 59);
              tc_fence_after();
              for (int kc = 0; kc < P.film_kc; ++kc) {
                uint32_t s = stage_wait();
                if (elected) {
                  const uint64_t ad = desc_adv(a_desc0, (uint32_t)kc * 32768u);
                  const uint64_t bd = desc_adv(b_desc0, s * kFtStage);
#pragma unroll
                  for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      umma2_bf16(tmem_base + n * 128, desc_adv(ad, h * 16384 + j * 32), desc_adv(bd, h * 8192 + j * 32), idesc,
                                 (uint32_t)((kc | h | j) != 0));
                }
                stage_release(s);
              }
              if (elected) umma2_commit_mc(bar_acc + 8 * n, 3);
              __syncwarp();
            }
            ++slot;
          }
        }
        // ---- input GEMM (K = 32)
        mbar_wait_cluster(bar_ready, slot & 1u, 52);
        tc_fence_after();
        for (int n = 0; n < NCH; ++n) {
          uint32_t s = stage_wait();
          if (elected) {
            const uint64_t bd = desc_adv(b_desc0, s * kFtStage);
#pragma unroll
            for (int j = 0; j < 2; ++j)
              umma2_bf16(tmem_base + n * 128, desc_adv(a_desc0, j * 32), desc_adv(bd, j * 32), idesc, j > 0);
          }
          stage_release(s);
        }
        commit_all();
        // ---- forward hidden GEMMs, chunk pipelined
        for (int g = 0; g < 2 * L; ++g) {
          ft_for_each_item(NCH, [&](int n, int k, int wait_c, int commit_n) {
            if (wait_c >= 0) { mbar_wait_cluster(bar_ready + 8 * wait_c, slot & 1u, 54); tc_fence_after(); if (wait_c == 0 && mstamp) g_ft_dbg[100 + 2 * (slot % 50)] = clock64(); }
            uint32_t s = stage_wait();
            if (elected) {
              const uint64_t ad = desc_adv(a_desc0, (uint32_t)k * 32768u);
              const uint64_t bd = desc_adv(b_desc0, s * kFtStage);
              const uint32_t d_tmem = tmem_base + n * 128;
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  umma2_bf16(d_tmem, desc_adv(ad, h * 16384 + j * 32), desc_adv(bd, h * 8192 + j * 32), idesc,
                             (uint32_t)((k | h | j) != 0));
            }
            stage_release(s);
            if (commit_n >= 0 && elected) umma2_commit_mc(bar_acc + 8 * commit_n, 3);
            __syncwarp();
          });
          if (mstamp) g_ft_dbg[101 + 2 * (slot % 50)] = clock64();
          ++slot;
        }
        // ---- head GEMM: K chunk k needs ready[k]
        for (int k = 0; k < NCH; ++k) {
          mbar_wait_cluster(bar_ready + 8 * k, slot & 1u, 56);
          tc_fence_after();
          uint32_t s = stage_wait();
          if (elected) {
            const uint64_t ad = desc_adv(a_desc0, (uint32_t)k * 32768u);
            const uint64_t bd = desc_adv(b_desc0, s * kFtStage);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma2_bf16(tmem_base, desc_adv(ad, h * 16384 + j * 32), desc_adv(bd, h * hp2 * 128 + j * 32), idesc_head,
                           (uint32_t)((k | h | j) != 0));
          }
          stage_release(s);
        }
        commit_all();
        // ---- head^T: dh_L = dlogits . W_head  (K = head_pad, A = dlogits in K block 0)
        mbar_wait_cluster(bar_ready, slot & 1u, 57);
        tc_fence_after();
        for (int n = 0; n < NCH; ++n) {
          uint32_t s = stage_wait();
          if (elected) {
            const uint64_t bd = desc_adv(b_desc0_mn, s * kFtStage);
            for (int j = 0; j < P.head_pad / 16; ++j)
              umma2_bf16(tmem_base + n * 128, desc_adv(a_desc0, j * 32), desc_adv(bd, j * 2048), idesc_mn, j > 0);
          }
          stage_release(s);
        }
        commit_all();
        // ---- backward hidden GEMMs (W2_l^T, W1_l^T), chunk pipelined, B tiles MN-major [128 k rows][64 n columns] per CTA
        for (int g = 0; g < 2 * L; ++g) {
          ft_for_each_item(NCH, [&](int n, int k, int wait_c, int commit_n) {
            if (wait_c >= 0) { mbar_wait_cluster(bar_ready + 8 * wait_c, slot & 1u, 58); tc_fence_after(); if (wait_c == 0 && mstamp) g_ft_dbg[100 + 2 * (slot % 50)] = clock64(); }
            uint32_t s = stage_wait();
            if (elected) {
              const uint64_t ad = desc_adv(a_desc0, (uint32_t)k * 32768u);
              const uint64_t bd = desc_adv(b_desc0_mn, s * kFtStage);
              const uint32_t d_tmem = tmem_base + n * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                umma2_bf16(d_tmem, desc_adv(ad, (j >> 2) * 16384 + (j & 3) * 32), desc_adv(bd, j * 2048), idesc_mn,
                           (uint32_t)((k | j) != 0));
            }
            stage_release(s);
            if (commit_n >= 0 && elected) umma2_commit_mc(bar_acc + 8 * commit_n, 3);
            __syncwarp();
          });
          if (mstamp) g_ft_dbg[101 + 2 * (slot % 50)] = clock64();
          ++slot;
        }
      }
    }
  } else {
    // =============================== epilogue / compute warps ===============================
    // Global-memory traffic of the epilogues.  TMEM pins "thread = batch row", and a warp-wide 16-byte access to 32 different
    // rows of a row-major [B, H] array costs 32 tag cycles in L1TEX (one per 128-byte line): with ~10 such arrays per layer
    // that alone was 25 us per layer (measured, first version of this kernel).  So:
    //  * arrays only this kernel reads back (z1, s, h0, the residual gradient, gamma|beta from the FiLM GEMM) use a
    //    TILE-PRIVATE layout: per (tile, chunk, column group, 16-column batch, lane quarter) one 1 KB block
    //    [half][lane][8 bf16], so a warp's 16-byte accesses are 512 contiguous bytes (4 lines instead of 32);
    //  * arrays the weight-gradient GEMMs consume row-major (a_l, u_l, h_L, dz1_l, dz2_l, dh_0) are not stored by the
    //    threads at all: they ARE the next A operand, which sits in shared memory in exactly the layout a SWIZZLE_128B TMA
    //    box has, so each warp's lane 0 issues a bulk tensor store of the [32 rows x 64 columns] the warp just finished;
    //  * only dgamma|dbeta (row-major for the FiLM weight gradient) go out as per-row stores.
    const int lq = warp & 3, cs = warp >> 2;            // TMEM lane quarter, column group
    const int m = lq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    const bool worker = (cs == 0);
    uint32_t slot = 0;
    const bool stamp = (P.dbg & 1) && blockIdx.x == 0 && tid == 0;
    constexpr int kGroups = kFtEpiWarps / 4;

    auto signal_ready = [&](int c) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_ready + 8 * c);
    };
    auto wait_acc = [&](int n, int code) {
      mbar_wait(bar_acc + 8 * n, slot & 1u, code);
      tc_fence_after();
    };
    auto store_sA = [&](int c0, const uint32_t (&o)[8]) {
      const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
    };
    auto epi_barrier = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kFtEpiThreads) : "memory"); };
    // every warp's lane 0 issues (and waits for) the bulk stores of the warp's own rows; a CTA-wide drain is needed only where
    // OTHER warps overwrite the region (K block 0: the next tile's input rows, the dlogits)
    auto stores_drained = [&]() {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      epi_barrier();
    };

    for (int it = 0; it < P.iters; ++it) {
      const int64_t pair = (int64_t)(blockIdx.x >> 1) + (int64_t)it * (gridDim.x >> 1);
      const int64_t tile = pair * 2 + crank;
      const int64_t grow = tile * 128 + m;
      const bool tile_ok = tile < P.n_tiles;
      const bool valid = tile_ok && grow < P.B;
      const int64_t rowG = grow * 2 * H;                    // element offset of this row in [B, 2H] arrays
      const int row0 = (int)(tile * 128);                   // first batch row of the tile (TMA coordinate)
      // tile-private arrays: element offset of (chunk n, batch b) for this thread; the second 8 columns are 256 elements on
      auto poff = [&](int n, int b) -> int64_t {
        return ((((tile * NCH + n) * kGroups + cs) * kFtBatches + b) * 4 + lq) * 512 + lane * 8;
      };
      auto ldp = [&](const __nv_bfloat16* base, int n, int b, uint32_t (&v)[8]) {      // plain (coherent) loads: same-kernel data
        const uint4* p = reinterpret_cast<const uint4*>(base + poff(n, b));
        uint4 x = make_uint4(0, 0, 0, 0), y = x;
        if (tile_ok && !(P.dbg & 16)) { x = p[0]; y = p[32]; }
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
      };
      auto stp = [&](__nv_bfloat16* base, int n, int b, const uint32_t (&o)[8]) {
        uint4* p = reinterpret_cast<uint4*>(base + poff(n, b));
        if (tile_ok && !(P.dbg & 4)) { p[0] = make_uint4(o[0], o[1], o[2], o[3]); p[32] = make_uint4(o[4], o[5], o[6], o[7]); }
      };
      // One epilogue sweep: kFtBatches * NCH batches of 16 columns per thread.  `pre(n, b, la, lb, lc)` issues the global loads
      // of a batch (they do not depend on the accumulator: batch i+1's fly under batch i's TMEM read and math);
      // `body(n, b, c0, r, la, lb, lc)` consumes them and writes the next A operand into shared memory.  When a chunk is
      // complete, thread 0 stores it to `omap` (coordinate z) with TMA.
      // The batch loop is deliberately NOT unrolled: unrolled, the 4L+3 sweeps are ~400 KB of straight-line SASS that every warp
      // executes exactly once per tile -- pure instruction streaming from L2 (first-occurrence sweeps measured 3x slower than
      // repeats).  Rolled, one sweep is a few hundred instructions reused 4 NCH times.
      auto sweep = [&](int code, bool sig, const CUtensorMap* omap, int oz, auto pre, auto body) {
        // operand prefetch: one batch ahead.  With 8 epilogue warps (168 registers) the next batch's loads are issued into a
        // second register set before this batch's TMEM read; with 16 warps (96 registers, 4 warps per scheduler to hide the
        // latency instead) they are issued into the SAME registers right after this batch's math has consumed them.
        constexpr bool kDouble = kFtEpiWarps <= 8;
        uint32_t la[8], lb[8], lc[8], ld[8], na[8], nb[8], nc[8], nd[8];
        pre(0, 0, la, lb, lc, ld);
#pragma unroll 1
        for (int idx = 0; idx < kFtBatches * NCH; ++idx) {
          const int n = idx / kFtBatches, b = idx % kFtBatches;
          const int c0 = n * 128 + cs * kFtColsPerWarp + b * 16;
          if (kDouble && idx + 1 < kFtBatches * NCH) pre((idx + 1) / kFtBatches, (idx + 1) % kFtBatches, na, nb, nc, nd);
          if (b == 0) {
            wait_acc(n, code);
            if (n == 0 && stamp) g_ft_dbg[2 * (slot % 50)] = clock64();
          }
          uint32_t r[16];
          tmem_ld16(t_lane + c0, r);
          tmem_wait_ld16(r);
          body(n, b, c0, r, la, lb, lc, ld);
          if (kDouble) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { la[i] = na[i]; lb[i] = nb[i]; lc[i] = nc[i]; ld[i] = nd[i]; }
          } else if (idx + 1 < kFtBatches * NCH) {
            pre((idx + 1) / kFtBatches, (idx + 1) % kFtBatches, la, lb, lc, ld);
          }
          if (b == kFtBatches - 1) {
            if (sig) signal_ready(n);
            else { fence_async_smem(); __syncwarp(); }
            if (omap && !(P.dbg & 2)) {
              // rows 32 lq .. of K block kb form a [32 x 64] box of the swizzled A tile (32-row slabs keep the 128-byte swizzle
              // phase).  It was written by this warp alone (8 epilogue warps) or by this warp and its neighbour column group (16
              // warps: a 64-thread named barrier joins the two); the owner's lane 0 stores it with TMA -- no CTA-wide barrier --
              // and owns the warp's bulk groups: before the region is written again (next sweep, same chunk) that store must have
              // finished READING shared memory, so at most NCH - 1 younger groups may stay pending.
              constexpr int kWarpsPerBox = 64 / kFtColsPerWarp;
              const int kb = 2 * n + cs / kWarpsPerBox;
              if (kWarpsPerBox == 2) asm volatile("bar.sync %0, 64;" ::"r"(2 + lq * 2 + (cs >> 1)) : "memory");
              if (lane == 0 && tile_ok && (cs % kWarpsPerBox) == 0) {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                             ::"l"(omap), "r"(smem_u32(sA + kb * 16384 + lq * 4096)), "r"(kb * 64), "r"(row0 + lq * 32), "r"(oz)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NCH - 1) : "memory");
              }
              if (kWarpsPerBox == 2) asm volatile("bar.sync %0, 64;" ::"r"(2 + lq * 2 + (cs >> 1)) : "memory");
              else __syncwarp();
            }
          }
        }
        if (stamp) g_ft_dbg[2 * (slot % 50) + 1] = clock64();
        ++slot;
      };
      auto no_pre = [&](int, int, uint32_t (&)[8], uint32_t (&)[8], uint32_t (&)[8], uint32_t (&)[8]) {};

      stores_drained();                                     // the previous tile's last bulk stores have read shared memory
      if (P.film_kc > 0) {
        // ---- FiLM phase (RQC/model.py:9-10): cond = [time_emb[t] | basis_emb[basis]] is gathered into the A-operand buffer (K blocks
        //      0 .. 2E/64), then 2L GEMMs against Wfilm_l produce gamma_l / beta_l, which the sweeps below write -- bias added -- into
        //      the tile-private arrays the later epilogues read.  Replaces a separate GEMM launch of B x 2E x 2HL (14 waves of
        //      4-K-block CTAs at batch 8192, 118 us).
        {
          const int tt = valid ? P.t[grow] : 0, bb = valid ? P.basis[grow] : 0;
          const int kbt = P.E >> 6;                                      // K blocks of the time half
          for (int kb = cs; kb < 2 * kbt; kb += kGroups) {
            const uint4* src = reinterpret_cast<const uint4*>(kb < kbt ? P.temb + (int64_t)tt * P.E + kb * 64
                                                                       : P.bemb + (int64_t)bb * P.E + (kb - kbt) * 64);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = valid ? __ldg(src + ch) : make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) signal_ready(c);
        for (int g = 0; g < 2 * L; ++g) {
          const float* fb = P.film_b + (int64_t)(g >> 1) * P.film_b_stride + (g & 1) * H;
          __nv_bfloat16* dst = P.gb + (int64_t)g * P.priv_elems;
          sweep(67, g + 1 < 2 * L, nullptr, 0, no_pre,
                [&](int n, int b, int c0, const uint32_t (&r)[16], const uint32_t (&)[8], const uint32_t (&)[8], const uint32_t (&)[8], const uint32_t (&)[8]) {
                  uint32_t o[8];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(fb + c0) + i);
                    o[2 * i] = pack_bf16(__uint_as_float(r[4 * i]) + bv.x, __uint_as_float(r[4 * i + 1]) + bv.y);
                    o[2 * i + 1] = pack_bf16(__uint_as_float(r[4 * i + 2]) + bv.z, __uint_as_float(r[4 * i + 3]) + bv.w);
                  }
                  stp(dst, n, b, o);
                });
        }
        // every warp has seen the last FiLM accumulator chunk complete, i.e. all reads of cond by the tensor core are done:
        // K block 0 may now take the input rows
      }
      // ---- input rows: [bits, 1, 0.. | bits, 1, 0..] (hi / lo halves of the collapsed input table), K columns 0..31
      if (worker) {
        const uint32_t xbits = valid ? P.xt[grow] : 0u;
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int k0 = 2 * i, k1 = 2 * i + 1;
          uint32_t lo = k0 < N ? ((xbits >> k0) & 1u) : (k0 == N ? 1u : 0u);
          uint32_t hi = k1 < N ? ((xbits >> k1) & 1u) : (k1 == N ? 1u : 0u);
          w[i] = (lo ? 0x3F80u : 0u) | ((hi ? 0x3F80u : 0u) << 16);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          *reinterpret_cast<uint4*>(sA + a_chunk_off(0, m, 2 * half)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(sA + a_chunk_off(0, m, 2 * half + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) signal_ready(c);

      // ================================================= forward =================================================
      // ---- E_in: h0 (saved); a_0 = h0 (1 + gamma_0) + beta_0
      {
        const __nv_bfloat16* gam = P.gb;                    // [L][2][tile-private]
        const __nv_bfloat16* bet = P.gb + P.priv_elems;
        sweep(60, true, &map_act, 0,
              [&](int n, int b, uint32_t (&la)[8], uint32_t (&lb)[8], uint32_t (&)[8], uint32_t (&)[8]) { ldp(gam, n, b, la); ldp(bet, n, b, lb); },
              [&](int n, int b, int c0, const uint32_t (&r)[16], const uint32_t (&la)[8], const uint32_t (&lb)[8], const uint32_t (&)[8], const uint32_t (&)[8]) {
                uint32_t o[8], hs[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
                  hs[i] = pack_bf16(v0, v1);
                  o[i] = pack_bf16(fmaf(v0, 1.0f + bf16_lo(la[i]), bf16_lo(lb[i])), fmaf(v1, 1.0f + bf16_hi(la[i]), bf16_hi(lb[i])));
                }
                store_sA(c0, o);
                stp(P.h0s, n, b, hs);
              });
      }
      for (int l = 0; l < L; ++l) {
        // ---- E1: z1 = acc + b1; u = silu(z1); silu'(z1) saved for the backward half (the same tanh gives both)
        {
          const float* hb1 = sB1 + l * H;
          __nv_bfloat16* d1p = P.d1s + (int64_t)l * P.priv_elems;
          sweep(61, true, &map_act, 2 * l + 1, no_pre,
                [&](int n, int b, int c0, const uint32_t (&r)[16], const uint32_t (&)[8], const uint32_t (&)[8], const uint32_t (&)[8], const uint32_t (&)[8]) {
                  uint32_t o[8], dv[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float z0 = __uint_as_float(r[2 * i]) + hb1[c0 + 2 * i], z1 = __uint_as_float(r[2 * i + 1]) + hb1[c0 + 2 * i + 1];
                    const float g0 = ft_sigmoid(z0), g1 = ft_sigmoid(z1);
                    o[i] = pack_bf16(z0 * g0, z1 * g1);
                    dv[i] = pack_bf16(g0 * fmaf(z0, 1.0f - g0, 1.0f), g1 * fmaf(z1, 1.0f - g1, 1.0f));
                  }
                  store_sA(c0, o);
                  stp(d1p, n, b, dv);
                });
        }
        // ---- E2: s = h_l + acc + b2; h_{l+1} = silu(s) and silu'(s) saved; a_{l+1} = FiLM_{l+1}(h) (or h itself before the head).
        //      The residual stream h_l is re-read from what the previous epilogue saved rather than living in 64 registers per
        //      thread across the whole network.
        {
          const float* hb2 = sB2 + l * H;
          const bool last = (l == L - 1);
          const __nv_bfloat16* gam = P.gb + (int64_t)(last ? 0 : l + 1) * 2 * P.priv_elems;
          const __nv_bfloat16* bet = gam + P.priv_elems;
          __nv_bfloat16* hp = P.hs + (int64_t)l * P.priv_elems;
          __nv_bfloat16* dp = P.dsv + (int64_t)l * P.priv_elems;
          const __nv_bfloat16* hsrc = l > 0 ? P.hs + (int64_t)(l - 1) * P.priv_elems : P.h0s;
          sweep(62, true, last ? &map_hL : &map_act, last ? 0 : 2 * (l + 1),
                [&](int n, int b, uint32_t (&la)[8], uint32_t (&lb)[8], uint32_t (&lc)[8], uint32_t (&)[8]) {
                  if (!last) { ldp(gam, n, b, la); ldp(bet, n, b, lb); }
                  ldp(hsrc, n, b, lc);
                },
                [&](int n, int b, int c0, const uint32_t (&r)[16], const uint32_t (&la)[8], const uint32_t (&lb)[8], const uint32_t (&lc)[8], const uint32_t (&)[8]) {
                  uint32_t o[8], hv[8], dv[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float s0 = bf16_lo(lc[i]) + __uint_as_float(r[2 * i]) + hb2[c0 + 2 * i];
                    const float s1 = bf16_hi(lc[i]) + __uint_as_float(r[2 * i + 1]) + hb2[c0 + 2 * i + 1];
                    const float g0 = ft_sigmoid(s0), g1 = ft_sigmoid(s1);
                    const float v0 = s0 * g0, v1 = s1 * g1;
                    hv[i] = pack_bf16(v0, v1);
                    dv[i] = pack_bf16(g0 * fmaf(s0, 1.0f - g0, 1.0f), g1 * fmaf(s1, 1.0f - g1, 1.0f));
                    o[i] = last ? hv[i]
                                : pack_bf16(fmaf(v0, 1.0f + bf16_lo(la[i]), bf16_lo(lb[i])), fmaf(v1, 1.0f + bf16_hi(la[i]), bf16_hi(lb[i])));
                  }
                  store_sA(c0, o);
                  stp(hp, n, b, hv);
                  stp(dp, n, b, dv);
                });
        }
      }
      // ---- head epilogue: logits -> cross-entropy (RQC/main.py:110) -> dlogits (A operand of the head^T GEMM, K block 0)
      wait_acc(0, 63);
      stores_drained();                                     // h_L's bulk store has read K block 0 before dlogits overwrite it
      if (worker) {
        uint32_t r[16], r2[16];
        tmem_ld16(t_lane, r);
        tmem_wait_ld16(r);
        if (P.head_pad > 16) { tmem_ld16(t_lane + 16, r2); tmem_wait_ld16(r2); }
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r2[i] = 0;
        }
        const uint32_t bits = valid ? P.x0[grow] : 0u;
        float lrow = 0.f;
        uint32_t dl[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const bool on = q < N;
          const uint32_t a = q < 8 ? r[2 * q] : r2[2 * (q - 8)], c = q < 8 ? r[2 * q + 1] : r2[2 * (q - 8) + 1];
          const float l0 = __uint_as_float(a) + (on ? P.head_b[2 * q] : 0.f), l1 = __uint_as_float(c) + (on ? P.head_b[2 * q + 1] : 0.f);
          const float mx = fmaxf(l0, l1), e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), sm = e0 + e1;
          const uint32_t y = (bits >> q) & 1u;
          const float inv = __fdividef(1.0f, sm);
          const bool use = on && valid;
          lrow += use ? (mx + __logf(sm)) - (y ? l1 : l0) : 0.f;
          dl[q] = use ? pack_bf16((e0 * inv - (y ? 0.f : 1.f)) * P.scale, (e1 * inv - (y ? 1.f : 0.f)) * P.scale) : 0u;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(sA + a_chunk_off(0, m, ch)) = make_uint4(dl[4 * ch], dl[4 * ch + 1], dl[4 * ch + 2], dl[4 * ch + 3]);
        if (valid) {
          uint4* dp = reinterpret_cast<uint4*>(P.dlog + grow * 32);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) dp[ch] = make_uint4(dl[4 * ch], dl[4 * ch + 1], dl[4 * ch + 2], dl[4 * ch + 3]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lrow += __shfl_xor_sync(0xFFFFFFFFu, lrow, o);
        if (lane == 0 && tile_ok) P.loss_part[tile * 4 + lq] = lrow;
      }
      ++slot;
#pragma unroll
      for (int c = 0; c < NCH; ++c) signal_ready(c);

      // ================================================= backward =================================================
      // ---- B_head: ds_{L-1} = dh_L * silu'(s_{L-1})  (= dz2_{L-1}; also kept tile-private as the residual gradient)
      {
        const __nv_bfloat16* dp = P.dsv + (int64_t)(L - 1) * P.priv_elems;
        sweep(64, true, &map_dz, 2 * (L - 1) + 1,
              [&](int n, int b, uint32_t (&la)[8], uint32_t (&)[8], uint32_t (&)[8], uint32_t (&)[8]) { ldp(dp, n, b, la); },
              [&](int n, int b, int c0, const uint32_t (&r)[16], const uint32_t (&la)[8], const uint32_t (&)[8], const uint32_t (&)[8], const uint32_t (&)[8]) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  o[i] = pack_bf16(__uint_as_float(r[2 * i]) * bf16_lo(la[i]), __uint_as_float(r[2 * i + 1]) * bf16_hi(la[i]));
                store_sA(c0, o);
                stp(P.dsp, n, b, o);
              });
      }
      for (int l = L - 1; l >= 0; --l) {
        // ---- B2: dz1_l = (dz2_l . W2_l) * silu'(z1_l)
        {
          const __nv_bfloat16* d1p = P.d1s + (int64_t)l * P.priv_elems;
          sweep(65, true, &map_dz, 2 * l,
                [&](int n, int b, uint32_t (&la)[8], uint32_t (&)[8], uint32_t (&)[8], uint32_t (&)[8]) { ldp(d1p, n, b, la); },
                [&](int, int, int c0, const uint32_t (&r)[16], const uint32_t (&la)[8], const uint32_t (&)[8], const uint32_t (&)[8], const uint32_t (&)[8]) {
                  uint32_t o[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    o[i] = pack_bf16(__uint_as_float(r[2 * i]) * bf16_lo(la[i]), __uint_as_float(r[2 * i + 1]) * bf16_hi(la[i]));
                  store_sA(c0, o);
                });
        }
        // ---- B1: da_l = dz1_l . W1_l ; dgamma = da * h_l, dbeta = da ; dh_l = ds_l + da (1 + gamma_l) ;
        //          l > 0: ds_{l-1} = dh_l * silu'(s_{l-1}) = dz2_{l-1} ; l == 0: dh_0
        {
          const __nv_bfloat16* gam = P.gb + (int64_t)l * 2 * P.priv_elems;                                   // gamma_l
          const __nv_bfloat16* hsrc = l > 0 ? P.hs + (int64_t)(l - 1) * P.priv_elems : P.h0s;               // h_l
          const __nv_bfloat16* dsrc = P.dsv + (int64_t)(l > 0 ? l - 1 : 0) * P.priv_elems;                   // silu'(s_{l-1})
          __nv_bfloat16* dgp = P.dgb + (int64_t)l * P.B * 2 * H + rowG;
          const bool inner = l > 0;
          sweep(66, inner, inner ? &map_dz : &map_dh0, inner ? 2 * (l - 1) + 1 : 0,
                [&](int n, int b, uint32_t (&la)[8], uint32_t (&lb)[8], uint32_t (&lc)[8], uint32_t (&ld)[8]) {
                  ldp(gam, n, b, la); ldp(hsrc, n, b, lb); ldp(P.dsp, n, b, lc);
                  if (inner) ldp(dsrc, n, b, ld);
                },
                [&](int n, int b, int c0, const uint32_t (&r)[16], const uint32_t (&la)[8], const uint32_t (&lb)[8], const uint32_t (&lc)[8], const uint32_t (&ld)[8]) {
                  uint32_t o[8], dg[8], db[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float da0 = __uint_as_float(r[2 * i]), da1 = __uint_as_float(r[2 * i + 1]);
                    dg[i] = pack_bf16(da0 * bf16_lo(lb[i]), da1 * bf16_hi(lb[i]));
                    db[i] = pack_bf16(da0, da1);
                    float dh0v = fmaf(da0, 1.0f + bf16_lo(la[i]), bf16_lo(lc[i])), dh1v = fmaf(da1, 1.0f + bf16_hi(la[i]), bf16_hi(lc[i]));
                    if (inner) { dh0v *= bf16_lo(ld[i]); dh1v *= bf16_hi(ld[i]); }
                    o[i] = pack_bf16(dh0v, dh1v);
                  }
                  store_sA(c0, o);
                  if (inner) stp(P.dsp, n, b, o);
                  ft_stg32(dgp + c0, valid && !(P.dbg & 8), dg);
                  ft_stg32(dgp + H + c0, valid && !(P.dbg & 8), db);
                });
        }
      }
      // the last sweep (l == 0) does not signal ready[]: the GEMM that follows it is the NEXT tile's input GEMM, whose operand
      // rows are written at the top of the next iteration -- that tile-start signal is this slot's arrival set
      if (it + 1 == P.iters) tc_fence_before();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // bulk stores fully performed before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kFtEpiWarps + 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}
