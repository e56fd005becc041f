// Persistent tcgen05 reverse sampler for sm_100a.
//
// One CTA per SM owns a tile of 128 trajectories of one measurement basis and walks all T reverse steps
// of the D3PM without leaving the SM (RQC/diffusion.py:58-79 / SS/diffusion.py:61-80 + RQC/model.py:51-70):
//
//   x_t bits --(K=32 bf16 hi/lo MMA against the collapsed input table)--> h0            [TMEM]
//   per ResBlock:  a = bf16(h*(1+gamma)+beta)      -> smem (UMMA K-major, 128B swizzle)
//                  acc = a . W1^T   (tcgen05.mma, bf16 x bf16 -> fp32 in TMEM, W tiles streamed by TMA)
//                  u = bf16(silu(acc+b1))          -> smem
//                  acc = u . W2^T
//                  h = silu(h + acc + b2)          (residual stream h kept in REGISTERS as bf16x2)
//   logits = h . Whead^T ; softmax over {0,1}; posterior / re-noise draw with Philox4x32-10 -> x_{t-1}
//
// gamma/beta are table lookups (Tt[t] + Tb[basis], one vector per step because t and the basis are
// uniform across the tile), so the FiLM and input-projection GEMMs of the reference disappear.
//
// Warp roles (576 threads): warps 0-15 epilogue/compute (warp w owns TMEM lanes 32*(w%4).. and hidden
// columns [(w/4)*H/4, ...)), warp 16 TMA producer, warp 17 MMA issuer + TMEM allocator.
#include <cuda.h>
#include <stdlib.h>

#include "sampler_tc.cuh"
#include "simt.cuh"
#include "tc_ptx.cuh"

namespace ddqst {

// ------------------------------------------------------------------------------------ kernel
constexpr int kEpiThreads = 512;
constexpr int kThreads = 640;          // 16 epilogue warps + one producer warpgroup (TMA, MMA, 2 idle)
constexpr int kStages = 4;
constexpr int kStageBytes = 16384;

struct TcParams {
  // packed model
  const float* Tt; const float* Tb; const float* bias1; const float* bias2; const float* head_b;
  const float* sched;
  int N, T, L, mode, head_pad;
  // work
  const int32_t* basis_ids; int32_t n_bases; int64_t spb; int64_t shot_offset; uint64_t seed;
  int t_start, t_end;                 // reverse steps t_start .. t_end (inclusive, descending)
  const uint16_t* x_init;             // if non-null: x_{t_start} per global row instead of the INIT draw
  void* out_packed; int out_elem;     // elem bytes 1/2 (0 = none)
  uint16_t* out_x16;                  // optional uint16 copy (sample_step)
  uint32_t* out_hist;
  float* logits_out;                  // optional [rows, 2N] of the last executed step
  int64_t tiles_per_basis, n_tiles;
  int32_t iters;                      // tile iterations per CTA (uniform so clustered CTAs stay in lock step)
};

template <int H>
struct TcCfg {
  static constexpr int KB = H / 64;                    // K blocks of 64
  static constexpr int NT = H < 128 ? H : 128;         // MMA N per instruction / rows per weight tile
  static constexpr int NC = H / NT;                    // N chunks per GEMM
  static constexpr int CW = H / 4;                     // hidden columns per epilogue thread
  static constexpr int NBATCH = CW / 16;
  static constexpr int A_BYTES = KB * 16384;
  static constexpr int STAGE_TX = NT * 128;            // bytes per weight tile
};

template <int H>
__host__ __device__ constexpr int tc_smem_bytes(int L) {
  return 1024 /*align slack*/ + TcCfg<H>::A_BYTES + kStages * kStageBytes + 2 * L * H * 4 /*g1,beta*/ +
         2 * L * H * 4 /*b1,b2*/ + 256 /*barriers*/;
}

template <int H, int CS>
__global__ void __launch_bounds__(kThreads, 1)
sampler_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_dt,
                  const __grid_constant__ CUtensorMap map_head, const TcParams P) {
  using C = TcCfg<H>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // no static shared memory in this kernel: the dynamic window starts 1024-aligned (checked below)
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) atomicCAS(&g_tc_abort, 0, 99);
  uint8_t* sA = smem;
  uint8_t* sRing = sA + C::A_BYTES;
  float* sG1 = (float*)(sRing + kStages * kStageBytes);   // [L][H] 1+gamma
  float* sBeta = sG1 + P.L * H;                           // [L][H]
  float* sB1 = sBeta + P.L * H;                           // [L][H]
  float* sB2 = sB1 + P.L * H;                             // [L][H]
  uint64_t* bars = (uint64_t*)(sB2 + P.L * H);
  // bars: [0..3] full, [4..7] empty, [8] acc_ready, [9] a_ready ; then tmem base
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_acc = smem_u32(bars + 2 * kStages), bar_a = smem_u32(bars + 2 * kStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = P.L, N = P.N;
  const int n_steps = P.t_start - P.t_end + 1;
  // cluster of CS CTAs: every weight tile is fetched once per cluster (each CTA loads 1/CS of its rows and
  // multicasts it), so a ring slot is free only when ALL CTAs of the cluster have consumed it
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, CS); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_a, kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < kEpiThreads) {
    for (int i = tid; i < L * H; i += kEpiThreads) { sB1[i] = P.bias1[i]; sB2[i] = P.bias2[i]; }
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();        // peers' barriers are initialised before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *((volatile uint32_t*)tmem_slot);

  if (warp >= 16) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 16) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_dt); tma_prefetch_desc(&map_head);
      uint32_t cnt = 0;
      auto acquire = [&](uint32_t bytes) -> uint32_t {
        uint32_t s = cnt % kStages, ph = (cnt / kStages) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u, 1);
        mbar_expect_tx(bar_full + 8 * s, bytes);
        ++cnt;
        return s;
      };
      constexpr int SL = C::NT / CS;                                 // rows of each tile this CTA fetches
      auto load_slice = [&](const CUtensorMap* map, uint32_t s, int col, int row) {
        uint32_t dst = smem_u32(sRing + s * kStageBytes) + crank * SL * 128;
        if (CS > 1) tma_load_2d_mc(dst, map, bar_full + 8 * s, col, row + (int)crank * SL, kMask);
        else tma_load_2d(dst, map, bar_full + 8 * s, col, row);
      };
      for (int it = 0; it < P.iters; ++it) {
        for (int st = 0; st < n_steps; ++st) {
          for (int nc = 0; nc < C::NC; ++nc) {                       // input table
            uint32_t s = acquire(C::STAGE_TX);
            load_slice(&map_dt, s, 0, nc * C::NT);
          }
          for (int g = 0; g < 2 * L; ++g)                            // W1, W2 of every block
            for (int nc = 0; nc < C::NC; ++nc)
              for (int kb = 0; kb < C::KB; ++kb) {
                uint32_t s = acquire(C::STAGE_TX);
                load_slice(&map_w, s, kb * 64, g * H + nc * C::NT);
              }
          {                                                          // head: KB boxes of [head_pad x 64], dealt round-robin
            uint32_t s = acquire((uint32_t)(C::KB * P.head_pad * 128));
            for (int kb = (int)crank; kb < C::KB; kb += CS) {
              uint32_t dst = smem_u32(sRing + s * kStageBytes + kb * P.head_pad * 128);
              if (CS > 1) tma_load_2d_mc(dst, &map_head, bar_full + 8 * s, kb * 64, 0, kMask);
              else tma_load_2d(dst, &map_head, bar_full + 8 * s, kb * 64, 0);
            }
          }
        }
      }
    }
  } else if (warp == 17) {
    // =============================== MMA issuer ===============================
    {
      const uint32_t elected = elect_one();           // warp converged, one lane issues (see elect_one)
      uint32_t cnt = 0, gemm = 0;
      const uint32_t idesc = umma_idesc_bf16(C::NT), idesc_head = umma_idesc_bf16(P.head_pad);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sRing));
      auto release = [&](uint32_t s) {
        if (elected) {
          if (CS > 1) umma_commit_mc(bar_empty + 8 * s, kMask);
          else umma_commit(bar_empty + 8 * s);
        }
        __syncwarp();
      };
      for (int it = 0; it < P.iters; ++it) {
        for (int st = 0; st < n_steps; ++st) {
          // ---- input GEMM: K = 32 (two k-steps of the first K block)
          mbar_wait(bar_a, gemm & 1u, 2);
          tc_fence_after();
          for (int nc = 0; nc < C::NC; ++nc) {
            uint32_t s = cnt % kStages, ph = (cnt / kStages) & 1u;
            mbar_wait(bar_full + 8 * s, ph, 3);
            tc_fence_after();
            if (elected) {
              const uint64_t bd = desc_adv(b_desc0, s * kStageBytes);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                umma_bf16(tmem_base + nc * C::NT, desc_adv(a_desc0, j * 32), desc_adv(bd, j * 32), idesc, j > 0);
            }
            release(s);
            ++cnt;
          }
          if (elected) umma_commit(bar_acc);
          __syncwarp();
          ++gemm;
          // ---- 2L hidden GEMMs
          for (int g = 0; g < 2 * L; ++g) {
            mbar_wait(bar_a, gemm & 1u, 4);
            tc_fence_after();
            for (int nc = 0; nc < C::NC; ++nc)
              for (int kb = 0; kb < C::KB; ++kb) {
                uint32_t s = cnt % kStages, ph = (cnt / kStages) & 1u;
                mbar_wait(bar_full + 8 * s, ph, 5);
                tc_fence_after();
                if (elected) {
                  const uint64_t ad = desc_adv(a_desc0, (uint32_t)kb * 16384u);
                  const uint64_t bd = desc_adv(b_desc0, s * kStageBytes);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    umma_bf16(tmem_base + nc * C::NT, desc_adv(ad, j * 32), desc_adv(bd, j * 32), idesc, (uint32_t)((kb | j) != 0));
                }
                release(s);
                ++cnt;
              }
            if (elected) umma_commit(bar_acc);
          __syncwarp();
            ++gemm;
          }
          // ---- head GEMM: N = head_pad
          mbar_wait(bar_a, gemm & 1u, 6);
          tc_fence_after();
          {
            uint32_t s = cnt % kStages, ph = (cnt / kStages) & 1u;
            mbar_wait(bar_full + 8 * s, ph, 7);
            tc_fence_after();
            if (elected) {
              const uint64_t bd = desc_adv(b_desc0, s * kStageBytes);
              for (int kb = 0; kb < C::KB; ++kb)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  umma_bf16(tmem_base, desc_adv(a_desc0, kb * 16384 + j * 32), desc_adv(bd, kb * P.head_pad * 128 + j * 32),
                            idesc_head, (uint32_t)((kb | j) != 0));
            }
            release(s);
            ++cnt;
          }
          if (elected) umma_commit(bar_acc);
          __syncwarp();
          ++gemm;
        }
      }
    }
  }
  } else {
    // =============================== epilogue / compute warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int lq = warp & 3, cq = warp >> 2;
    const int m = lq * 32 + lane;                       // row of the tile == TMEM lane
    const int cbase = cq * C::CW;
    const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    const bool worker = (cq == 0);                      // one thread per row does the draw
    uint32_t xr[C::CW / 2];                             // residual stream, bf16x2
    uint32_t acc_phase = 0;
    uint32_t xbits = 0;

    for (int it = 0; it < P.iters; ++it) {
      const int64_t tile_raw = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const bool tile_ok = tile_raw < P.n_tiles;                  // padding iterations compute but emit nothing
      const int64_t tile = tile_ok ? tile_raw : 0;
      const int64_t bslot = tile / P.tiles_per_basis;
      const int64_t row_in_basis = (tile % P.tiles_per_basis) * 128 + m;
      const bool valid = tile_ok && row_in_basis < P.spb;
      const uint32_t basis = (uint32_t)P.basis_ids[bslot];
      const uint64_t shot = (uint64_t)(P.shot_offset + row_in_basis);
      const int64_t grow = bslot * P.spb + row_in_basis;         // global output row

      auto write_input_row = [&]() {
        // A row for the input GEMM: [bits, 1, 0.. | bits, 1, 0..] (bf16), K columns 0..31 = chunks 0..3
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
      };

      if (worker) {
        if (P.x_init) xbits = valid ? P.x_init[grow] : 0u;
        else {
          xbits = 0;
          Philox4 p{};
          for (int q = 0; q < N; ++q) {
            if ((q & 3) == 0) p = stream_block(P.seed, basis, 0, DDQST_SITE_INIT, shot, q >> 2);
            xbits |= (lane_of(p, q) & 1u) << q;
          }
        }
        write_input_row();
      }
      fence_async_smem();
      mbar_arrive(bar_a);

      for (int t = P.t_start; t >= P.t_end; --t) {
        // ---- FiLM vectors of this step: (1+gamma, beta) = Tt[t] + Tb[basis]
        {
          const float* tt = P.Tt + (int64_t)t * L * 2 * H;
          const float* tb = P.Tb + (int64_t)basis * L * 2 * H;
          for (int i = tid; i < L * 2 * H; i += kEpiThreads) {
            int l = i / (2 * H), j = i - l * 2 * H;
            float v = __ldg(tt + i) + __ldg(tb + i);
            if (j < H) sG1[l * H + j] = 1.0f + v;
            else sBeta[l * H + j - H] = v;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        }
        // ---- input epilogue: h0 -> residual registers, a = film_0(h0) -> smem
        mbar_wait(bar_acc, acc_phase, 8); acc_phase ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int b = 0; b < C::NBATCH; ++b) {
          uint32_t r[16];
          const int c0 = cbase + b * 16;
          tmem_ld16(t_lane + c0, r);
          tmem_wait_ld16(r);
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
            xr[b * 8 + i] = pack_bf16(v0, v1);
            o[i] = pack_bf16(fmaf(v0, sG1[c0 + 2 * i], sBeta[c0 + 2 * i]), fmaf(v1, sG1[c0 + 2 * i + 1], sBeta[c0 + 2 * i + 1]));
          }
          const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
          *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(bar_a);

        for (int l = 0; l < L; ++l) {
          // ---- E1: u = silu(acc + b1)
          mbar_wait(bar_acc, acc_phase, 9); acc_phase ^= 1u;
          tc_fence_after();
          const float* b1 = sB1 + l * H;
#pragma unroll
          for (int b = 0; b < C::NBATCH; ++b) {
            uint32_t r[16];
            const int c0 = cbase + b * 16;
            tmem_ld16(t_lane + c0, r);
            tmem_wait_ld16(r);
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v0 = silu_fast(__uint_as_float(r[2 * i]) + b1[c0 + 2 * i]);
              float v1 = silu_fast(__uint_as_float(r[2 * i + 1]) + b1[c0 + 2 * i + 1]);
              o[i] = pack_bf16(v0, v1);
            }
            const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
          }
          tc_fence_before();
          fence_async_smem();
          mbar_arrive(bar_a);

          // ---- E2: h = silu(h + acc + b2); a = film_{l+1}(h) (or h itself before the head)
          mbar_wait(bar_acc, acc_phase, 10); acc_phase ^= 1u;
          tc_fence_after();
          const float* b2 = sB2 + l * H;
          const bool last = (l == L - 1);
          const float* g1 = sG1 + (last ? 0 : (l + 1) * H);
          const float* be = sBeta + (last ? 0 : (l + 1) * H);
#pragma unroll
          for (int b = 0; b < C::NBATCH; ++b) {
            uint32_t r[16];
            const int c0 = cbase + b * 16;
            tmem_ld16(t_lane + c0, r);
            tmem_wait_ld16(r);
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t xo = xr[b * 8 + i];
              float v0 = silu_fast(__uint_as_float(r[2 * i]) + b2[c0 + 2 * i] + bf16_lo(xo));
              float v1 = silu_fast(__uint_as_float(r[2 * i + 1]) + b2[c0 + 2 * i + 1] + bf16_hi(xo));
              uint32_t xn = pack_bf16(v0, v1);
              xr[b * 8 + i] = xn;
              o[i] = last ? xn : pack_bf16(fmaf(v0, g1[c0 + 2 * i], be[c0 + 2 * i]), fmaf(v1, g1[c0 + 2 * i + 1], be[c0 + 2 * i + 1]));
            }
            const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
          }
          tc_fence_before();
          fence_async_smem();
          mbar_arrive(bar_a);
        }

        // ---- head epilogue: logits -> softmax -> posterior / re-noise draw -> x_{t-1}
        mbar_wait(bar_acc, acc_phase, 11); acc_phase ^= 1u;
        tc_fence_after();
        if (worker) {
          uint32_t r[16], r2[16];
          tmem_ld16(t_lane, r);
          tmem_wait_ld16(r);
          if (P.head_pad > 16) { tmem_ld16(t_lane + 16, r2); tmem_wait_ld16(r2); }
          else {
#pragma unroll
            for (int i = 0; i < 16; ++i) r2[i] = 0;
          }
          float lg[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { lg[i] = __uint_as_float(r[i]) + P.head_b[i]; lg[16 + i] = __uint_as_float(r2[i]) + P.head_b[16 + i]; }
          if (P.logits_out && valid && t == P.t_end) {
            for (int i = 0; i < 2 * N; ++i) P.logits_out[grow * 2 * N + i] = lg[i];
          }
          xbits = reverse_step_bits(N, P.T, P.sched, P.mode, t, P.seed, basis, shot, xbits,
                                    [&](int q, int c) { return lg[2 * q + c]; });
          if (t > P.t_end) write_input_row();
          else if (valid) {
            if (P.out_elem == 1) ((uint8_t*)P.out_packed)[grow] = (uint8_t)xbits;
            else if (P.out_elem == 2) ((uint16_t*)P.out_packed)[grow] = (uint16_t)xbits;
            if (P.out_x16) P.out_x16[grow] = (uint16_t)xbits;
            if (P.out_hist) atomicAdd(P.out_hist + (bslot << N) + xbits, 1u);
          }
        }
        tc_fence_before();
        if (t > P.t_end) {
          fence_async_smem();
          mbar_arrive(bar_a);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();        // no CTA leaves while a peer may still multicast into it
  if (warp == 17) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
// bf16 matrix [rows, cols] row-major; box [box_rows, 64 cols]; 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn;
  DDQST_TRY(get_encode_fn(&fn));
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DDQST_REQUIRE(r == CUDA_SUCCESS, DDQST_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld box=%d)", (int)r, (long long)rows, (long long)cols, box_rows);
  return DDQST_OK;
}

int sampler_tc_supported(const ddqst_dims* d) {
  const int H = d->hidden_dim;
  DDQST_REQUIRE(H == 64 || H == 128 || H == 256 || H == 512, DDQST_EUNSUPPORTED,
                "tcgen05 sampler needs hidden_dim in {64,128,256,512}, got %d (use DDQST_PRECISION_FP32)", H);
  DDQST_REQUIRE(d->num_qubits <= 15, DDQST_EUNSUPPORTED, "tcgen05 sampler needs num_qubits <= 15");
  DDQST_REQUIRE((int64_t)d->num_blocks * H <= 2048, DDQST_EUNSUPPORTED,
                "tcgen05 sampler keeps FiLM/bias vectors in shared memory: num_blocks*hidden_dim must be <= 2048");
  return DDQST_OK;
}

int64_t sampler_tc_workspace_bytes(const ddqst_dims*, int64_t) { return 4096; }

template <int H, int CS>
static int launch_tc_cs(const ddqst_dims* d, const char* pack, const PackLayout& pl, const TcParams& P0, cudaStream_t s) {
  using C = TcCfg<H>;
  TcParams P = P0;
  CUtensorMap map_w, map_dt, map_head;
  DDQST_TRY(make_map(&map_w, pack + pl.w_bf16, (int64_t)d->num_blocks * 2 * H, H, C::NT / CS));
  DDQST_TRY(make_map(&map_dt, pack + pl.dt_bf16, H, 64, C::NT / CS));
  DDQST_TRY(make_map(&map_head, pack + pl.head_bf16, pl.head_pad, H, pl.head_pad));
  const int smem = tc_smem_bytes<H>(d->num_blocks);
  auto kern = sampler_tc_kernel<H, CS>;
  DDQST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // persistent grid: as many CTAs as can be co-resident (one per SM, whole clusters), never more than the work
  int max_ctas = num_sms() / CS * CS;
  if (CS > 1) {
    static int cached[8] = {0};
    if (!cached[CS]) {
      cfg.gridDim = dim3(max_ctas);
      int nclusters = 0;
      DDQST_CUDA_OK(cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg));
      cached[CS] = nclusters > 0 ? nclusters : 1;
    }
    if (cached[CS] * CS < max_ctas) max_ctas = cached[CS] * CS;
  }
  int64_t want = (P.n_tiles + CS - 1) / CS * CS;
  int grid = (int)(want < max_ctas ? want : max_ctas);
  P.iters = (int32_t)((P.n_tiles + grid - 1) / grid);
  cfg.gridDim = dim3(grid);
  DDQST_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_w, map_dt, map_head, P));
  return DDQST_OK;
}

static int tc_cluster_size() {
  static int cs = 0;
  if (!cs) {
    const char* e = getenv("DDQST_TC_CLUSTER");
    cs = e ? atoi(e) : 2;
    if (cs != 1 && cs != 2 && cs != 4) cs = 2;
  }
  return cs;
}

template <int H>
static int launch_tc(const ddqst_dims* d, const char* pack, const PackLayout& pl, const TcParams& P, cudaStream_t s) {
  switch (tc_cluster_size()) {
    case 1: return launch_tc_cs<H, 1>(d, pack, pl, P, s);
    case 4: return launch_tc_cs<H, 4>(d, pack, pl, P, s);
    default: return launch_tc_cs<H, 2>(d, pack, pl, P, s);
  }
}

#include "sampler_pair.cuh"

static bool use_pair_kernel(int H) {
  const char* e = getenv("DDQST_TC_KERNEL");
  if (e && e[0] == 'v' && e[1] == '1') return false;
  return H % 128 == 0;
}

static int dispatch_tc(const ddqst_dims* d, const char* pack, const PackLayout& pl, TcParams& P, cudaStream_t s) {
  DDQST_TRY(sampler_tc_supported(d));
  P.Tt = (const float*)(pack + pl.Tt); P.Tb = (const float*)(pack + pl.Tb);
  P.bias1 = (const float*)(pack + pl.bias1); P.bias2 = (const float*)(pack + pl.bias2);
  P.head_b = (const float*)(pack + pl.head_b);
  P.N = d->num_qubits; P.T = d->num_timesteps; P.L = d->num_blocks; P.head_pad = pl.head_pad;
  P.tiles_per_basis = (P.spb + 127) / 128;
  P.n_tiles = P.tiles_per_basis * P.n_bases;
  if (P.n_tiles == 0) return DDQST_OK;
  if (use_pair_kernel(d->hidden_dim)) {
    switch (d->hidden_dim) {
      case 128: return launch_pair<128>(d, pack, pl, P, s);
      case 256: return launch_pair<256>(d, pack, pl, P, s);
      default: return launch_pair<512>(d, pack, pl, P, s);
    }
  }
  switch (d->hidden_dim) {
    case 64: return launch_tc<64>(d, pack, pl, P, s);
    case 128: return launch_tc<128>(d, pack, pl, P, s);
    case 256: return launch_tc<256>(d, pack, pl, P, s);
    default: return launch_tc<512>(d, pack, pl, P, s);
  }
}

int sampler_tc_sample(const ddqst_dims* d, const char* pack, const PackLayout& pl, const float* sched, int mode,
                      const int32_t* basis_ids, int32_t n_bases, int64_t spb, int64_t shot_offset, uint64_t seed,
                      void* out_packed, uint32_t* out_hist, void*, int64_t, cudaStream_t s) {
  TcParams P{};
  P.sched = sched; P.mode = mode; P.basis_ids = basis_ids; P.n_bases = n_bases; P.spb = spb;
  P.shot_offset = shot_offset; P.seed = seed; P.t_start = d->num_timesteps; P.t_end = 1;
  P.out_packed = out_packed; P.out_elem = out_packed ? (d->num_qubits <= 8 ? 1 : 2) : 0; P.out_hist = out_hist;
  return dispatch_tc(d, pack, pl, P, s);
}

int sampler_tc_step(const ddqst_dims* d, const char* pack, const PackLayout& pl, const float* sched, int mode,
                    int32_t basis_id, int32_t t, int64_t shots, int64_t shot_offset, uint64_t seed,
                    const uint16_t* x_t, uint16_t* x_prev, float* logits_out, void* workspace, int64_t ws_bytes,
                    cudaStream_t s) {
  DDQST_REQUIRE(workspace && ws_bytes >= 4, DDQST_EWORKSPACE, "sample_step needs a workspace");
  int32_t* bid = (int32_t*)workspace;
  DDQST_CUDA_OK(cudaMemcpyAsync(bid, &basis_id, sizeof(int32_t), cudaMemcpyHostToDevice, s));
  TcParams P{};
  P.sched = sched; P.mode = mode; P.basis_ids = bid; P.n_bases = 1; P.spb = shots; P.shot_offset = shot_offset;
  P.seed = seed; P.t_start = t; P.t_end = t; P.x_init = x_t; P.out_x16 = x_prev; P.logits_out = logits_out;
  return dispatch_tc(d, pack, pl, P, s);
}

int sampler_tc_forward(const ddqst_dims*, const char*, const PackLayout&, const uint16_t*, const int32_t*,
                       const int32_t*, int64_t, float*, void*, int64_t, cudaStream_t) {
  set_error("bf16 forward with per-row (t, basis) is not built: use DDQST_PRECISION_FP32, or ddqst_sample_step for a uniform (t, basis) batch");
  return DDQST_EUNSUPPORTED;
}

// ------------------------------------------------------------------------------------ UMMA self test
// C[128*mt, n] = A[128*mt, k] (fp32 -> bf16 through the epilogue's swizzled store) . W[n, k]^T (bf16 via TMA),
// read back through the epilogue's tcgen05.ld mapping.  One CTA per M tile; k, n multiples of 64, <= 512.
__global__ void __launch_bounds__(kThreads, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap map_w, const float* __restrict__ A, int n, int k,
                     float* __restrict__ Cout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // no static shared memory in this kernel: the dynamic window starts 1024-aligned (checked below)
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) atomicCAS(&g_tc_abort, 0, 99);
  uint8_t* sA = smem;                        // k/64 blocks of 16 KB
  uint8_t* sB = sA + (k / 64) * 16384;       // one 64-row x 64-col tile (8 KB) per (nc, kb), all resident
  uint64_t* bars = (uint64_t*)(sB + (n / 64) * (k / 64) * 8192);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const uint32_t bar_full = smem_u32(bars), bar_acc = smem_u32(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = k / 64, NC = n / 64;
  if (tid == 0) {
    mbar_init(bar_full, 1); mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *((volatile uint32_t*)tmem_slot);
  if (warp == 16 && lane == 0) {
    mbar_expect_tx(bar_full, (uint32_t)(NC * KB * 8192));
    for (int nc = 0; nc < NC; ++nc)
      for (int kb = 0; kb < KB; ++kb)
        tma_load_2d(smem_u32(sB + (nc * KB + kb) * 8192), &map_w, bar_full, kb * 64, nc * 64);
  }
  if (tid < kEpiThreads) {
    // thread (row m, column quarter cq) converts its slice of A to bf16 in the UMMA layout
    const int lq = warp & 3, cq = warp >> 2, m = lq * 32 + lane;
    const float* arow = A + ((int64_t)blockIdx.x * 128 + m) * k;
    for (int c0 = cq * 16; c0 < k; c0 += 64) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pack_bf16(arow[c0 + 2 * i], arow[c0 + 2 * i + 1]);
      const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    fence_async_smem();
  }
  __syncthreads();
  if (warp == 17 && lane == 0) {
    mbar_wait(bar_full, 0, 20);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(64);
    for (int nc = 0; nc < NC; ++nc)
      for (int kb = 0; kb < KB; ++kb)
        for (int j = 0; j < 4; ++j)
          umma_bf16(tmem_base + nc * 64, umma_desc_sw128(smem_u32(sA) + kb * 16384 + j * 32),
                    umma_desc_sw128(smem_u32(sB + (nc * KB + kb) * 8192) + j * 32), idesc, (kb | j) != 0);
    umma_commit(bar_acc);
  }
  if (tid < kEpiThreads) {
    const int lq = warp & 3, cq = warp >> 2, m = lq * 32 + lane;
    mbar_wait(bar_acc, 0, 21);
    tc_fence_after();
    float* crow = Cout + ((int64_t)blockIdx.x * 128 + m) * n;
    for (int c0 = cq * 16; c0 < n; c0 += 64) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + c0, r);
      tmem_wait_ld16(r);
#pragma unroll
      for (int i = 0; i < 16; ++i) crow[c0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}


// 2-CTA variant: C[256*mp, n] with one cta_group::2 MMA (M=256: 128 rows from each CTA of the pair, B split along N:
// each CTA stages n/2 rows of W).  Exercises tcgen05.alloc/mma/commit .cta_group::2, the 2SM TMA form that
// completes on the leader's barrier, and the cross-CTA "operand ready" arrive.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
umma2_selftest_kernel(const __grid_constant__ CUtensorMap map_w, const float* __restrict__ A, int n, int k,
                      float* __restrict__ Cout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // no static shared memory in this kernel: the dynamic window starts 1024-aligned (checked below)
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) atomicCAS(&g_tc_abort, 0, 99);
  uint8_t* sA = smem;
  uint8_t* sB = sA + (k / 64) * 16384;                 // per K block: [n/2 rows x 128 B]
  const int half = n / 2, KB = k / 64;
  uint64_t* bars = (uint64_t*)(sB + KB * half * 128);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const uint32_t bar_full = smem_u32(bars), bar_acc = smem_u32(bars + 1), bar_a = smem_u32(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(bar_full, 2); mbar_init(bar_acc, 1); mbar_init(bar_a, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *((volatile uint32_t*)tmem_slot);
  if (warp == 16 && lane == 0) {
    if (crank == 0) mbar_expect_tx(bar_full, (uint32_t)(2 * KB * half * 128));
    else mbar_arrive_leader(bar_full);
    for (int kb = 0; kb < KB; ++kb)
      tma_load_2d_2sm(smem_u32(sB + kb * half * 128), &map_w, bar_full, kb * 64, (int)crank * half);
  }
  if (tid < kEpiThreads) {
    const int lq = warp & 3, cq = warp >> 2, m = lq * 32 + lane;
    const float* arow = A + ((int64_t)blockIdx.x * 128 + m) * k;
    for (int c0 = cq * 16; c0 < k; c0 += 64) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pack_bf16(arow[c0 + 2 * i], arow[c0 + 2 * i + 1]);
      const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    fence_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    if (tid == 0) mbar_arrive_leader(bar_a);
  }
  if (warp == 17 && lane == 0 && crank == 0) {
    mbar_wait_cluster(bar_a, 0, 30);
    mbar_wait_cluster(bar_full, 0, 31);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16_m(256, n);
    for (int kb = 0; kb < KB; ++kb)
      for (int j = 0; j < 4; ++j)
        umma2_bf16(tmem_base, umma_desc_sw128(smem_u32(sA) + kb * 16384 + j * 32),
                   umma_desc_sw128(smem_u32(sB + kb * half * 128) + j * 32), idesc, (kb | j) != 0);
    umma2_commit_mc(bar_acc, 3);
  }
  if (tid < kEpiThreads) {
    const int lq = warp & 3, cq = warp >> 2, m = lq * 32 + lane;
    mbar_wait(bar_acc, 0, 32);
    tc_fence_after();
    float* crow = Cout + ((int64_t)blockIdx.x * 128 + m) * n;
    for (int c0 = cq * 16; c0 < n; c0 += 64) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + c0, r);
      tmem_wait_ld16(r);
#pragma unroll
      for (int i = 0; i < 16; ++i) crow[c0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 17) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace ddqst

using namespace ddqst;

extern "C" {

int ddqst_selftest_umma2(const float* a, const uint16_t* w_bf16, int32_t m_pairs, int32_t n, int32_t k, float* c, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(m_pairs >= 1 && n >= 16 && n <= 256 && n % 16 == 0 && k >= 64 && k <= 512 && k % 64 == 0, DDQST_EINVAL_SHAPE, "selftest shape");
  const int smem = 1024 + (k / 64) * 16384 + (k / 64) * (n / 2) * 128 + 256;
  DDQST_REQUIRE(smem <= 232448, DDQST_EINVAL_SHAPE, "selftest needs %d bytes of shared memory", smem);
  CUtensorMap map_w;
  DDQST_TRY(make_map(&map_w, w_bf16, n, k, n / 2));
  DDQST_CUDA_OK(cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma2_selftest_kernel<<<2 * m_pairs, kThreads, smem, (cudaStream_t)stream>>>(map_w, a, n, k, c);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_selftest_umma(const float* a, const uint16_t* w_bf16, int32_t m_tiles, int32_t n, int32_t k, float* c, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(m_tiles >= 1 && n >= 64 && n <= 512 && n % 64 == 0 && k >= 64 && k <= 512 && k % 64 == 0, DDQST_EINVAL_SHAPE, "selftest shape");
  const int smem = 1024 + (k / 64) * 16384 + (n / 64) * (k / 64) * 8192 + 256;
  DDQST_REQUIRE(smem <= 232448, DDQST_EINVAL_SHAPE, "selftest needs %d bytes of shared memory", smem);
  CUtensorMap map_w;
  DDQST_TRY(make_map(&map_w, w_bf16, n, k, 64));
  DDQST_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<m_tiles, kThreads, smem, (cudaStream_t)stream>>>(map_w, a, n, k, c);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

// synchronises the device and returns the first pipeline-timeout code recorded by a tcgen05 kernel (0 = none)
int ddqst_debug_tc_status(void) {
  int a = tc_abort_fetch();          // this TU's kernels (samplers, UMMA self tests)
  int b = train_tc_abort_fetch();    // train_tc.cu's GEMM kernels
  int c = recon_tc_abort_fetch();    // recon.cu's ring Jacobi
  return a != 0 ? a : (b != 0 ? b : c);
}

}  // extern "C"
