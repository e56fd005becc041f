// CTA-pair, chunk-pipelined tcgen05 reverse sampler (included by sampler_tc.cu; uses its PTX wrappers).
//
// Same arithmetic as sampler_tc_kernel (see the header of sampler_tc.cu) with two structural changes:
//
//  1. cta_group::2 -- two CTAs (a cluster of 2 SMs) run one M=256 MMA: 128 trajectories from each CTA, the
//     weight tile (B operand) split along N between the two shared memories.  A ring byte now feeds twice the
//     MMA work, which is what the L2->SMEM latency (~1.6k cycles per 16 KB tile, measured) needs to be hidden.
//
//  2. chunk pipelining -- the hidden dimension is cut in NCH = H/128 chunks.  All 16 epilogue warps sweep one
//     128-column accumulator chunk at a time and publish it (ready[c]); the MMA issuer walks the (n,k) tile
//     grid of the NEXT GEMM in an order that only ever needs the chunks already published:
//         group c < NCH-1 : every (n,k) with max(n,k) == c
//         last group      : (c,0..c-1), then (0,c) .. (c-1,c), (c,c), committing acc_done[n] after each (n,c)
//     so the tensor pipe works under the epilogue, and because (c,n) is issued before (n,c) the epilogue may
//     overwrite operand chunk n (a and u share one buffer) as soon as acc_done[n] fires.
#pragma once
// (included inside namespace ddqst)

// DDQST_PAIR_LD32=1: fetch a warp's 32 accumulator columns of a chunk with one tcgen05.ld.x32 instead of two .x16 (one TMEM round trip
// per chunk).  Measured in round 2: 2.75 M bitstrings/s against 2.76-2.83 M -- no gain (the kernel is power-capped: shorter stalls turn
// into lower clocks) and 144 B more spills per thread, so it stays off.
#ifndef DDQST_PAIR_LD32
#define DDQST_PAIR_LD32 0
#endif
constexpr int kRing2 = 4;                 // stages of 16 KB per CTA

template <int H>
struct PairCfg {
  static constexpr int NCH = H / 128;                   // 128-column chunks
  static constexpr int A_BYTES = (H / 64) * 16384;
};

template <int H>
__host__ __device__ constexpr int pair_smem_bytes(int L) {
  return 1024 + PairCfg<H>::A_BYTES + kRing2 * kStageBytes + 4 * L * H * 4 + 256;
}

// f(n, k, wait_chunk or -1, commit_chunk or -1) in issue order
template <typename F>
__device__ __forceinline__ void for_each_item(int NCH, F f) {
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

template <int H>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
sampler_pair_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_dt,
                    const __grid_constant__ CUtensorMap map_head, const TcParams P) {
  using C = PairCfg<H>;
  constexpr int NCH = C::NCH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // no static shared memory in this kernel: the dynamic window starts 1024-aligned (checked below)
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) atomicCAS(&g_tc_abort, 0, 99);
  uint8_t* sA = smem;
  uint8_t* sRing = sA + C::A_BYTES;
  float* sG1 = (float*)(sRing + kRing2 * kStageBytes);
  float* sBeta = sG1 + P.L * H;
  float* sB1 = sBeta + P.L * H;
  float* sB2 = sB1 + P.L * H;
  uint64_t* bars = (uint64_t*)(sB2 + P.L * H);
  // bars: [0..3] full (leader's used), [4..7] empty, [8..11] acc_done, [12..15] ready (leader's used)
  uint32_t* tmem_slot = (uint32_t*)(bars + 16);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 4), bar_acc = smem_u32(bars + 8),
                 bar_ready = smem_u32(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = P.L, N = P.N;
  const int n_steps = P.t_start - P.t_end + 1;
  const uint32_t crank = cluster_ctarank();
  const int hp2 = P.head_pad / 2;                       // head rows staged per CTA

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_full + 8 * i, 2);                   // leader's expect_tx arrive + peer's arrive
      mbar_init(bar_empty + 8 * i, 1);                  // multicast commit from the leader's MMA thread
      mbar_init(bar_acc + 8 * i, 1);
      mbar_init(bar_ready + 8 * i, 32);                 // 16 warps x 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid < kEpiThreads) {
    for (int i = tid; i < L * H; i += kEpiThreads) { sB1[i] = 0.5f * P.bias1[i]; sB2[i] = 0.5f * P.bias2[i]; }   // half biases, see E1
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *((volatile uint32_t*)tmem_slot);

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 16) {
      // =============================== TMA producer (both CTAs: own half of every B tile) ===============================
      {
        const uint32_t elected = elect_one();           // whole warp converged; one lane issues
        if (elected) { tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_dt); tma_prefetch_desc(&map_head); }
        uint32_t cnt = 0;
        auto acquire = [&](uint32_t bytes_both) -> uint32_t {
          uint32_t s = cnt % kRing2, ph = (cnt / kRing2) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u, 1);
          if (elected) {
            if (crank == 0) mbar_expect_tx(bar_full + 8 * s, bytes_both);
            else mbar_arrive_leader(bar_full + 8 * s);
          }
          ++cnt;
          return s;
        };
        for (int it = 0; it < P.iters; ++it) {
          for (int st = 0; st < n_steps; ++st) {
            for (int n = 0; n < NCH; ++n) {                                   // input table, K block 0 only
              uint32_t s = acquire(2 * 8192);
              if (elected) tma_load_2d_2sm(smem_u32(sRing + s * kStageBytes), &map_dt, bar_full + 8 * s, 0, n * 128 + (int)crank * 64);
            }
            for (int g = 0; g < 2 * L; ++g)
              for_each_item(NCH, [&](int n, int k, int, int) {
                uint32_t s = acquire(2 * 16384);
                uint32_t dst = smem_u32(sRing + s * kStageBytes);
                int row = g * H + n * 128 + (int)crank * 64;
                if (elected) {
                  tma_load_2d_2sm(dst, &map_w, bar_full + 8 * s, (2 * k) * 64, row);
                  tma_load_2d_2sm(dst + 8192, &map_w, bar_full + 8 * s, (2 * k + 1) * 64, row);
                }
              });
            for (int k = 0; k < NCH; ++k) {                                   // head, K chunk k
              uint32_t s = acquire((uint32_t)(2 * 2 * hp2 * 128));
              uint32_t dst = smem_u32(sRing + s * kStageBytes);
              if (elected) {
                tma_load_2d_2sm(dst, &map_head, bar_full + 8 * s, (2 * k) * 64, (int)crank * hp2);
                tma_load_2d_2sm(dst + hp2 * 128, &map_head, bar_full + 8 * s, (2 * k + 1) * 64, (int)crank * hp2);
              }
            }
          }
        }
      }
    } else if (warp == 17 && crank == 0) {
      // =============================== MMA issuer (leader CTA only) ===============================
      {
        const uint32_t elected = elect_one();           // whole warp converged; one lane issues
        uint32_t cnt = 0, slot = 0;
        const uint32_t idesc = umma_idesc_bf16_m(256, 128), idesc_head = umma_idesc_bf16_m(256, P.head_pad);
        const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
        const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sRing));
        auto stage_wait = [&]() -> uint32_t {
          uint32_t s = cnt % kRing2, ph = (cnt / kRing2) & 1u;
          mbar_wait_cluster(bar_full + 8 * s, ph, 3);
          tc_fence_after();
          return s;
        };
        auto stage_release = [&](uint32_t s) {
          if (elected) umma2_commit_mc(bar_empty + 8 * s, 3);
          ++cnt;
        };
        for (int it = 0; it < P.iters; ++it) {
          for (int st = 0; st < n_steps; ++st) {
            // ---- input GEMM (K = 32): needs only the input rows (published with ready[0])
            mbar_wait_cluster(bar_ready, slot & 1u, 2);
            tc_fence_after();
            for (int n = 0; n < NCH; ++n) {
              uint32_t s = stage_wait();
              if (elected) {
                const uint64_t bd = desc_adv(b_desc0, s * kStageBytes);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  umma2_bf16(tmem_base + n * 128, desc_adv(a_desc0, j * 32), desc_adv(bd, j * 32), idesc, j > 0);
              }
              stage_release(s);
            }
            if (elected)
              for (int n = 0; n < NCH; ++n) umma2_commit_mc(bar_acc + 8 * n, 3);
            __syncwarp();
            ++slot;
            // ---- hidden GEMMs, chunk pipelined
            for (int g = 0; g < 2 * L; ++g) {
              for_each_item(NCH, [&](int n, int k, int wait_c, int commit_n) {
                if (wait_c >= 0) { mbar_wait_cluster(bar_ready + 8 * wait_c, slot & 1u, 4); tc_fence_after(); }
                uint32_t s = stage_wait();
                if (elected) {
                  const uint64_t ad = desc_adv(a_desc0, (uint32_t)k * 32768u);
                  const uint64_t bd = desc_adv(b_desc0, s * kStageBytes);
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
              ++slot;
            }
            // ---- head GEMM: K chunk k needs ready[k]
            for (int k = 0; k < NCH; ++k) {
              mbar_wait_cluster(bar_ready + 8 * k, slot & 1u, 6);
              tc_fence_after();
              uint32_t s = stage_wait();
              if (elected) {
                const uint64_t ad = desc_adv(a_desc0, (uint32_t)k * 32768u);
                const uint64_t bd = desc_adv(b_desc0, s * kStageBytes);
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    umma2_bf16(tmem_base, desc_adv(ad, h * 16384 + j * 32), desc_adv(bd, h * hp2 * 128 + j * 32), idesc_head,
                               (uint32_t)((k | h | j) != 0));
              }
              stage_release(s);
            }
            if (elected)
              for (int n = 0; n < NCH; ++n) umma2_commit_mc(bar_acc + 8 * n, 3);
            __syncwarp();
            ++slot;
          }
        }
      }
    }
  } else {
    // =============================== epilogue / compute warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int lq = warp & 3, cs = warp >> 2;            // TMEM lane quarter, 32-column sub-chunk
    const int m = lq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    const bool worker = (cs == 0);
    uint32_t xr[NCH][16];                               // residual stream (bf16x2): 32 columns of every chunk
    uint32_t slot = 0;
    uint32_t xbits = 0;

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
    auto store16 = [&](int c0, const uint32_t (&o)[8]) {
      const int kb = c0 >> 6, ch = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(sA + a_chunk_off(kb, m, ch + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
    };

    for (int it = 0; it < P.iters; ++it) {
      const int64_t tile_raw = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const bool tile_ok = tile_raw < P.n_tiles;
      const int64_t tile = tile_ok ? tile_raw : 0;
      const int64_t bslot = tile / P.tiles_per_basis;
      const int64_t row_in_basis = (tile % P.tiles_per_basis) * 128 + m;
      const bool valid = tile_ok && row_in_basis < P.spb;
      const uint32_t basis = (uint32_t)P.basis_ids[bslot];
      const uint64_t shot = (uint64_t)(P.shot_offset + row_in_basis);
      const int64_t grow = bslot * P.spb + row_in_basis;

      auto write_input_row = [&]() {
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
#pragma unroll
      for (int c = 0; c < NCH; ++c) signal_ready(c);

      // FiLM vectors of a step: 1 + gamma and beta for every block, = Tt[t] + Tb[basis].  Step t_start's are loaded here by
      // everyone; the following steps' are loaded by the twelve non-worker warps while the four worker warps run the head
      // epilogue of the step before (every FiLM row is dead once the last block's E2 is done), so the L2 latency of the
      // table rows no longer sits between two steps.
      auto load_film = [&](int t, int first, int stride) {
        const float* tt = P.Tt + (int64_t)t * L * 2 * H;
        const float* tb = P.Tb + (int64_t)basis * L * 2 * H;
        for (int i = first; i < L * 2 * H; i += stride) {
          int l = i / (2 * H), j = i - l * 2 * H;
          float v = __ldg(tt + i) + __ldg(tb + i);
          if (j < H) sG1[l * H + j] = 1.0f + v;
          else sBeta[l * H + j - H] = v;
        }
      };
      load_film(P.t_start, tid, kEpiThreads);

      for (int t = P.t_start; t >= P.t_end; --t) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // this step's FiLM rows are in shared memory
        // ---- input epilogue: h0 -> residual registers, a = film_0(h0)
#pragma unroll
        for (int n = 0; n < NCH; ++n) {
          wait_acc(n, 8);
          const int c0 = n * 128 + cs * 32;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            uint32_t r[16];
            tmem_ld16(t_lane + c0 + b * 16, r);
            tmem_wait_ld16(r);
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = c0 + b * 16 + 2 * i;
              float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
              xr[n][b * 8 + i] = pack_bf16(v0, v1);
              o[i] = pack_bf16(fmaf(v0, sG1[c], sBeta[c]), fmaf(v1, sG1[c + 1], sBeta[c + 1]));
            }
            store16(c0 + b * 16, o);
          }
          signal_ready(n);
        }
        ++slot;

        for (int l = 0; l < L; ++l) {
          // ---- E1: u = silu(acc + b1); sB1/sB2 hold HALF the biases: silu(z) = h + h*tanh(h) with h = z/2 = fma(acc, .5, b/2)
          const float* hb1 = sB1 + l * H;
#pragma unroll
          for (int n = 0; n < NCH; ++n) {
            wait_acc(n, 9);
            const int c0 = n * 128 + cs * 32;
#if DDQST_PAIR_LD32
            uint32_t rr[32];                          // both 16-column batches of the warp's slice in ONE tcgen05.ld: one TMEM round trip per chunk
            tmem_ld32(t_lane + c0, rr);
            tmem_wait_ld32(rr);
#endif
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#if DDQST_PAIR_LD32
              const uint32_t* r = rr + 16 * b;
#else
              uint32_t r[16];
              tmem_ld16(t_lane + c0 + b * 16, r);
              tmem_wait_ld16(r);
#endif
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = c0 + b * 16 + 2 * i;
                o[i] = pack_bf16(silu_half(fmaf(__uint_as_float(r[2 * i]), 0.5f, hb1[c])),
                                 silu_half(fmaf(__uint_as_float(r[2 * i + 1]), 0.5f, hb1[c + 1])));
              }
              store16(c0 + b * 16, o);
            }
            signal_ready(n);
          }
          ++slot;

          // ---- E2: h = silu(h + acc + b2); a = film_{l+1}(h) (or h itself before the head)
          const float* hb2 = sB2 + l * H;
          const bool last = (l == L - 1);
          const float* g1 = sG1 + (last ? 0 : (l + 1) * H);
          const float* be = sBeta + (last ? 0 : (l + 1) * H);
#pragma unroll
          for (int n = 0; n < NCH; ++n) {
            wait_acc(n, 10);
            const int c0 = n * 128 + cs * 32;
#if DDQST_PAIR_LD32
            uint32_t rr[32];
            tmem_ld32(t_lane + c0, rr);
            tmem_wait_ld32(rr);
#endif
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#if DDQST_PAIR_LD32
              const uint32_t* r = rr + 16 * b;
#else
              uint32_t r[16];
              tmem_ld16(t_lane + c0 + b * 16, r);
              tmem_wait_ld16(r);
#endif
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = c0 + b * 16 + 2 * i;
                const uint32_t xo = xr[n][b * 8 + i];
                float v0 = silu_half(fmaf(bf16_lo(xo), 0.5f, fmaf(__uint_as_float(r[2 * i]), 0.5f, hb2[c])));
                float v1 = silu_half(fmaf(bf16_hi(xo), 0.5f, fmaf(__uint_as_float(r[2 * i + 1]), 0.5f, hb2[c + 1])));
                const uint32_t xn = pack_bf16(v0, v1);
                xr[n][b * 8 + i] = xn;
                o[i] = last ? xn : pack_bf16(fmaf(v0, g1[c], be[c]), fmaf(v1, g1[c + 1], be[c + 1]));
              }
              store16(c0 + b * 16, o);
            }
            signal_ready(n);
          }
          ++slot;
        }

        // ---- head epilogue (worker warps); everyone else fetches the next step's FiLM rows meanwhile
        if (!worker && t > P.t_end) load_film(t - 1, tid - 128, kEpiThreads - 128);
        wait_acc(0, 11);
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
        ++slot;
        if (t > P.t_end) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) signal_ready(c);
        } else {
          tc_fence_before();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 17) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int H>
static int launch_pair(const ddqst_dims* d, const char* pack, const PackLayout& pl, const TcParams& P0, cudaStream_t s) {
  TcParams P = P0;
  CUtensorMap map_w, map_dt, map_head;
  DDQST_TRY(make_map(&map_w, pack + pl.w_bf16, (int64_t)d->num_blocks * 2 * H, H, 64));
  DDQST_TRY(make_map(&map_dt, pack + pl.dt_bf16, H, 64, 64));
  DDQST_TRY(make_map(&map_head, pack + pl.head_bf16, pl.head_pad, H, pl.head_pad / 2));
  const int smem = pair_smem_bytes<H>(d->num_blocks);
  auto kern = sampler_pair_kernel<H>;
  DDQST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int max_ctas = num_sms() / 2 * 2;
  int64_t want = (P.n_tiles + 1) / 2 * 2;
  int grid = (int)(want < max_ctas ? want : max_ctas);
  P.iters = (int32_t)((P.n_tiles + grid - 1) / grid);
  kern<<<grid, kThreads, smem, s>>>(map_w, map_dt, map_head, P);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

