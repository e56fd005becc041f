// Tensor-core training step (T1, RQC/main.py:105-115) for sm_100a: every GEMM of the denoiser's forward,
// data-gradient and weight-gradient passes runs on tcgen05 (bf16 x bf16 -> fp32 accumulators in TMEM), fed by
// 3-D TMA tensor maps, with the elementwise work of the reference's autograd graph fused into the epilogues:
//
//   forward   gb_l   = cond . Wfilm_l^T + b            (one launch, z = block)            RQC/model.py:9-10
//             h_0    = xin . Win^T + b ; a_0 = h_0 (1+g_0) + beta_0                        RQC/model.py:53-56,11
//             z1_l   = a_l . W1_l^T + b1 ; u_l = silu(z1_l)                                RQC/model.py:17-19
//             z2_l   = h_l + u_l . W2_l^T + b2 ; h_{l+1} = silu(z2_l) ; a_{l+1} = FiLM     RQC/model.py:24,11
//             logits = h_L . Whead^T + b ; mean CE ; dlogits                              RQC/model.py:70, RQC/main.py:110
//   backward  dgrads with the stored weights read as MN-major operands (no transposed copies),
//             wgrads with BOTH operands MN-major (dY^T . X, reduction over the batch), z-batched over layers,
//             embedding gradients scattered from the epilogue, bias gradients by a deterministic column sum.
//
// Operands are bf16 (a bf16 shadow of the flat fp32 parameter buffer is kept current by the fused Adam kernel);
// accumulation, the saved pre-activations z1/z2, the FiLM vectors and the residual stream stay fp32.
//
// GEMM kernel: one 128 x BN output tile per CTA, BK = 64, 4-stage TMA ring, warp 4 = TMA producer, warp 5 = MMA
// issuer + TMEM allocator, warps 0-3 = epilogue (thread = output row, tcgen05.ld 32x32b).
#include <cuda.h>
#include <stdlib.h>

#include "sampler_tc.cuh"
#include "simt.cuh"
#include "tc_ptx.cuh"

namespace ddqst {

constexpr int kGtThreads = 192;
// thread geometry of the fused forward + data-gradient kernel (train_fused.cuh) and of its tile-private array layout
constexpr int kFtEpiWarps = 16;            // 4 per TMEM lane quarter: warp w owns rows 32*(w%4).., columns (w/4)*32.. of every 128-column chunk
constexpr int kFtEpiThreads = kFtEpiWarps * 32;
constexpr int kFtColsPerWarp = 128 / (kFtEpiWarps / 4);
constexpr int kFtBatches = kFtColsPerWarp / 16;
constexpr int kFtThreads = kFtEpiThreads + 64;   // + warp 8 (TMA) + warp 9 (MMA, TMEM alloc).  10 warps = at most 3 per SM sub-partition,
                                                 // so ptxas may use 168 registers per thread (a 16 + 2-warp layout is capped at 96 and spilled
                                                 // ~500 B per thread into local memory, which misses L1 here: shared memory takes the carve-out)
template <int BN> __host__ __device__ constexpr int gt_stages() { return BN <= 64 ? 8 : (BN <= 128 ? 6 : 4); }
constexpr int kGtMaxZ = 32;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// UMMA shared-memory descriptor for an MN-major operand tile made of 64-wide MN groups, each group = [k rows x 128 B]
// written by TMA SWIZZLE_128B: canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, i.e.
// LBO = byte distance between MN groups, SBO = 1024 (eight 128-byte k rows).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t group_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((group_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}

// how (mn0, k, z) of a tile becomes TMA coordinates {inner, row, batch}
struct TcOperand {
  int mn_major;   // 0: global [mn rows][k cols] (K-major);  1: global [k rows][mn cols] (MN-major)
  int kmod;       // > 0: the reduction index runs over the batch dim too: k -> (k % kmod, batch += k / kmod)
  int zmul;       // batch coordinate = blockIdx.z * zmul (+ k / kmod)
};

enum TcEpi {
  TE_STORE = 0,   // out[z][r, c] = acc (+ bias[z][c]), fp32, masked to M x N
  TE_IN,          // h0 = acc + b -> o0 ; a0 = h0 (1+g) + beta -> b0
  TE_W1,          // z1 = acc + b -> o0 ; u = silu(z1) -> b0
  TE_W2,          // z2 = f0 + acc + b -> o0 ; h = silu(z2) -> o1 ; FiLM(h; f1) or h -> b0
  TE_HEAD,        // logits = acc + b -> o0 ; CE loss part -> o1[cta] ; dlogits -> b0 [B,32]
  TE_BHEAD,       // dz2 = acc * silu'(f0) -> o0 (fp32 residual grad), b0
  TE_BW2,         // dz1 = acc * silu'(f0) -> b0
  TE_BW1,         // da = acc: dgb -> b1 ; dh = f3 + da (1+g) ; (dz2_prev = dh silu'(f2) -> o0, b0) or (dh0 -> b0)
  TE_DCOND        // acc scattered: atomicAdd into time_emb / basis_emb gradient rows
};

struct TcGemm {
  int M, N, K;
  TcOperand a, b;
  // epilogue slots (meaning per TcEpi above)
  const float* bias; int64_t bias_zstride;
  const float *f0, *f1, *f2, *f3;
  float *o0, *o1;
  __nv_bfloat16 *b0, *b1;
  const int32_t *i0, *i1;
  const uint16_t* x0;
  int64_t ld;          // leading dim of [rows, H]-shaped arrays and of TE_STORE's output
  int64_t ldg;         // leading dim of FiLM-shaped arrays (2H)
  int64_t out_zoff[kGtMaxZ];
  int nq, E, flag;     // qubits (TE_HEAD), embed dim (TE_DCOND), variant flag (TE_W2: 1 = FiLM next, TE_BW1: 1 = l > 0)
  float scale;         // TE_HEAD: loss_scale / (B*N)
  long long* dbg;      // optional [8] clock64 stamps of CTA (0,0,0) (self test only)
  long long* trace;    // optional [4] %globaltimer ns of CTA (0,0,0): entry, after griddepcontrol.wait, epilogue done, EPI (benchmarks/train_trace.py)
};

template <int BN>
__host__ __device__ constexpr int gt_stage_bytes() { return 16384 + BN * 128; }
template <int BN>
__host__ __device__ constexpr int gt_smem_bytes() { return 1024 + gt_stages<BN>() * gt_stage_bytes<BN>() + 256; }

__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float dsilu_fast(float v) { float s = sigmoid_fast(v); return s * (1.0f + v * (1.0f - s)); }

__device__ __forceinline__ void ld32(const float* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void st32(float* p, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void st32_bf16(__nv_bfloat16* p, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<uint4*>(p)[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                                pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
}

// one 128 x BN output tile (bx, by) of batch entry bz; the whole CTA calls it
template <int BN, int EPI>
// ks / nks: split-K -- this CTA reduces K blocks [ks * ceil(KB/nks), ...) only and (nks > 1, TE_STORE) ADDS its tile atomically
__device__ __forceinline__ void gemm_tile(const CUtensorMap* mapA_p, const CUtensorMap* mapB_p, const TcGemm& G, int bx, int by, int bz,
                                          int ks = 0, int nks = 1) {
  const CUtensorMap& mapA = *mapA_p;
  const CUtensorMap& mapB = *mapB_p;
  constexpr int STAGE = gt_stage_bytes<BN>();
  constexpr int kGtStages = gt_stages<BN>();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kGtStages * STAGE);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kGtStages + 1);
  float* s_red = (float*)(tmem_slot + 4);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kGtStages), bar_acc = smem_u32(bars + 2 * kGtStages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = bx * 128, n0 = by * BN, z = bz;
  const int KBtot = (G.K + 63) >> 6, kb_per = (KBtot + nks - 1) / nks, kb0 = ks * kb_per;
  const int KB = KBtot - kb0 < kb_per ? KBtot - kb0 : kb_per;      // the host picks nks so that every split is non-empty

  // programmatic dependent launch: let the next kernel of the stream start its prologue now; our own inputs are
  // only touched after griddepcontrol.wait below (= the previous kernel has completed and flushed)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool dbg = G.dbg != nullptr && bx == 0 && by == 0 && bz == 0;
  const bool trace = G.trace != nullptr && bx == 0 && by == 0 && bz == 0 && tid == 0;
  if (dbg && tid == 0) G.dbg[0] = clock64();
  if (trace) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); G.trace[0] = t; }
  if (tid == 0) {
    for (int i = 0; i < kGtStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 4 && elect_one()) { tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *((volatile uint32_t*)tmem_slot);
  if (dbg && tid == 0) G.dbg[1] = clock64();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (dbg && tid == 0) G.dbg[2] = clock64();
  if (trace) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); G.trace[1] = t; }

  if (warp == 4) {
    // =============================== TMA producer ===============================
    const uint32_t elected = elect_one();
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % kGtStages;
      const uint32_t ph = (uint32_t)(kb / kGtStages) & 1u;
      mbar_wait(bar_empty + 8 * s, ph ^ 1u, 40);
      {
        // coordinates are computed by the whole (converged) warp so they stay in uniform registers; only the
        // issue itself is predicated on the elected lane (a divergent block makes ptxas wrap every uniform-datapath
        // instruction in an ELECT / BRA.U.ANY loop)
        const uint32_t full = bar_full + 8 * s;
        const uint32_t dstA = smem_u32(smem + s * STAGE), dstB = dstA + 16384;
        const int kg = (kb0 + kb) * 64;
        int ka = kg, za = z * G.a.zmul, kbb = kg, zb = z * G.b.zmul;
        if (G.a.kmod > 0) { za += kg / G.a.kmod; ka = kg % G.a.kmod; }
        if (G.b.kmod > 0) { zb += kg / G.b.kmod; kbb = kg % G.b.kmod; }
        if (elected) mbar_expect_tx(full, (uint32_t)STAGE);
        if (G.a.mn_major) {
          if (elected) tma_load_3d(dstA, &mapA, full, m0, ka, za);
          if (elected) tma_load_3d(dstA + 8192, &mapA, full, m0 + 64, ka, za);
        } else {
          if (elected) tma_load_3d(dstA, &mapA, full, ka, m0, za);
        }
        if (G.b.mn_major) {
#pragma unroll
          for (int g = 0; g < BN / 64; ++g)
            if (elected) tma_load_3d(dstB + g * 8192, &mapB, full, n0 + g * 64, kbb, zb);
        } else {
          if (elected) tma_load_3d(dstB, &mapB, full, kbb, n0, zb);
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ===============================
    const uint32_t elected = elect_one();
    const uint32_t idesc = umma_idesc_bf16(BN) | ((uint32_t)(G.a.mn_major != 0) << 15) | ((uint32_t)(G.b.mn_major != 0) << 16);
    const uint32_t a_step = G.a.mn_major ? 2048u : 32u, b_step = G.b.mn_major ? 2048u : 32u;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % kGtStages;
      const uint32_t ph = (uint32_t)(kb / kGtStages) & 1u;
      mbar_wait(bar_full + 8 * s, ph, 41);
      tc_fence_after();
      if (dbg && kb == 0 && elected) G.dbg[3] = clock64();
      if (dbg && kb == KB - 1 && elected) G.dbg[4] = clock64();
      {
        const uint32_t aaddr = smem_u32(smem + s * STAGE), baddr = aaddr + 16384;
        const uint64_t ad = G.a.mn_major ? umma_desc_mn_sw128(aaddr, 8192) : umma_desc_sw128(aaddr);
        const uint64_t bd = G.b.mn_major ? umma_desc_mn_sw128(baddr, 8192) : umma_desc_sw128(baddr);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (elected) umma_bf16(tmem_base, desc_adv(ad, j * a_step), desc_adv(bd, j * b_step), idesc, (uint32_t)((kb | j) != 0));
        if (elected) umma_commit(bar_empty + 8 * s);
        if (elected && kb == KB - 1) umma_commit(bar_acc);
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue: thread = output row ===============================
    const int r = m0 + warp * 32 + lane;
    const bool rv = r < G.M;
    // Epilogue operands (residual stream, FiLM vectors, saved pre-activations) do not depend on the accumulator: the
    // loads of a 32-column chunk are all issued together -- for the first chunk before the accumulator wait, so they
    // fly under the main loop -- instead of in dependent load -> use -> store -> load rounds (each round is a full
    // L2 latency with only 128 loading threads per CTA).
    float q0[32], q1[32], q2[32], q3[32];
    auto issue_loads = [&](int c0) {
      if (!rv) return;
      const int c = n0 + c0;
      const int64_t e = (int64_t)r * G.ld + c, ge = (int64_t)r * G.ldg + c;
      if (EPI == TE_IN) { ld32(G.f1 + ge, q0); ld32(G.f1 + ge + G.ld, q1); }
      if (EPI == TE_W2) { ld32(G.f0 + e, q0); if (G.flag) { ld32(G.f1 + ge, q1); ld32(G.f1 + ge + G.ld, q2); } }
      if (EPI == TE_BHEAD || EPI == TE_BW2) ld32(G.f0 + e, q0);
      if (EPI == TE_BW1) { ld32(G.f1 + e, q0); ld32(G.f0 + ge, q1); ld32(G.f3 + e, q2); if (G.flag) ld32(G.f2 + e, q3); }
    };
    issue_loads(0);
    mbar_wait(bar_acc, 0, 42);
    tc_fence_after();
    if (dbg && tid == 0) G.dbg[5] = clock64();
    float loss_acc = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (EPI == TE_HEAD && c0 > 0) break;
      uint32_t raw[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, raw);
      tmem_wait_ld32(raw);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
      const int c = n0 + c0;                      // first global column of this chunk
      if (EPI == TE_STORE) {
        if (rv && c < G.N) {
          const int64_t oe = G.out_zoff[z] + (int64_t)r * G.ld + c;
          const float* bz = G.bias ? G.bias + (int64_t)z * G.bias_zstride + c : nullptr;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            if (c + i < G.N) {      // N is a multiple of 4
              float4 t = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              if (bz && ks == 0) { t.x += bz[i]; t.y += bz[i + 1]; t.z += bz[i + 2]; t.w += bz[i + 3]; }
              if (nks > 1) {                       // split-K partial: the output was zeroed by the caller
                float* o = G.o0 + oe + i;
                atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
              } else if (G.flag == 2) { v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w; }                       // tile-private bf16, below
              else if (G.b0) *reinterpret_cast<uint2*>(G.b0 + oe + i) = make_uint2(pack_bf16(t.x, t.y), pack_bf16(t.z, t.w));   // bf16 output
              else *reinterpret_cast<float4*>(G.o0 + oe + i) = t;
            }
          }
          if (G.flag == 2) {
            // FiLM vectors for the fused kernel (train_fused.cuh): N = 2H columns (gamma | beta), written in the tile-private layout
            // [which][tile][chunk][column group][batch][lane quarter][half][lane][8] -- G.ldg = elements of one such array
            const int Hh = G.N >> 1, nch = Hh >> 7;
            const int which = c / Hh, cc = c - which * Hh, n = cc >> 7, cw = cc & 127, grp = cw / kFtColsPerWarp, b0 = (cw % kFtColsPerWarp) >> 4;
            const int64_t tile = r >> 7;
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
              const int64_t off = G.out_zoff[z] + (int64_t)which * G.ldg +
                                  (((((tile * nch + n) * (kFtEpiWarps / 4) + grp) * kFtBatches + b0 + bb) * 4 + warp) * 512) + lane * 8;
              uint4* p = reinterpret_cast<uint4*>(G.b0 + off);
              const float* w = v + 16 * bb;
              p[0] = make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
              p[32] = make_uint4(pack_bf16(w[8], w[9]), pack_bf16(w[10], w[11]), pack_bf16(w[12], w[13]), pack_bf16(w[14], w[15]));
            }
          }
        }
      } else if (EPI == TE_DCOND) {
        if (rv) {
          const int E = G.E;
          float* gt = G.o0 + (int64_t)G.i0[r] * E;
          float* gbs = G.o1 + (int64_t)G.i1[r] * E;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int cc = c + i;
            if (cc < E) atomicAdd(gt + cc, v[i]);
            else if (cc < 2 * E) atomicAdd(gbs + (cc - E), v[i]);
          }
        }
      } else if (EPI == TE_HEAD) {
        // columns [0, 2*nq) are the logits of row r (class fastest)
        float dl[32];
        float lrow = 0.f;
        const uint32_t bits = rv ? G.x0[r] : 0u;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float l0 = v[2 * q] + (q < G.nq ? G.bias[2 * q] : 0.f), l1 = v[2 * q + 1] + (q < G.nq ? G.bias[2 * q + 1] : 0.f);
          v[2 * q] = l0; v[2 * q + 1] = l1;
          float m = fmaxf(l0, l1), e0 = __expf(l0 - m), e1 = __expf(l1 - m), s = e0 + e1;
          const uint32_t y = (bits >> q) & 1u;
          const float inv = __fdividef(1.0f, s);
          const bool on = q < G.nq;
          lrow += on ? (m + __logf(s)) - (y ? l1 : l0) : 0.f;
          dl[2 * q] = on ? (e0 * inv - (y ? 0.f : 1.f)) * G.scale : 0.f;
          dl[2 * q + 1] = on ? (e1 * inv - (y ? 1.f : 0.f)) * G.scale : 0.f;
        }
        if (rv) {
          loss_acc += lrow;
          float* lo = G.o0 + (int64_t)r * (2 * G.nq);
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < 2 * G.nq) lo[i] = v[i];
          st32_bf16(G.b0 + (int64_t)r * 32, dl);
        }
      } else if (rv) {
        const int64_t e = (int64_t)r * G.ld + c;          // element offset in [rows, H] arrays
        if (EPI == TE_IN) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __ldg(G.bias + c + i);
          st32(G.o0 + e, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], 1.0f + q0[i], q1[i]);
          st32_bf16(G.b0 + e, v);
        } else if (EPI == TE_W1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __ldg(G.bias + c + i);
          st32(G.o0 + e, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= sigmoid_fast(v[i]);
          st32_bf16(G.b0 + e, v);
        } else if (EPI == TE_W2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += q0[i] + __ldg(G.bias + c + i);
          st32(G.o0 + e, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= sigmoid_fast(v[i]);
          st32(G.o1 + e, v);
          if (G.flag) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], 1.0f + q1[i], q2[i]);
          }
          st32_bf16(G.b0 + e, v);
        } else if (EPI == TE_BHEAD) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= dsilu_fast(q0[i]);
          st32(G.o0 + e, v);
          st32_bf16(G.b0 + e, v);
        } else if (EPI == TE_BW2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= dsilu_fast(q0[i]);
          st32_bf16(G.b0 + e, v);
        } else if (EPI == TE_BW1) {
          // v = da.  q0 = h_l, q1 = gamma_l, q2 = residual gradient (dz2_l), q3 = z2_{l-1}
#pragma unroll
          for (int i = 0; i < 32; ++i) q0[i] *= v[i];           // dgamma = da * h_in
          st32_bf16(G.b1 + (int64_t)r * G.ldg + c, q0);
          st32_bf16(G.b1 + (int64_t)r * G.ldg + G.ld + c, v);   // dbeta = da
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], 1.0f + q1[i], q2[i]);
          if (G.flag) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= dsilu_fast(q3[i]);
            st32(G.o0 + e, v);
          }
          st32_bf16(G.b0 + e, v);
        }
      }
      if (c0 + 32 < BN) issue_loads(c0 + 32);        // the q arrays are dead here; these loads fly under the next tcgen05.ld
    }
    if (EPI == TE_HEAD) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xFFFFFFFFu, loss_acc, o);
      if (lane == 0) s_red[warp] = loss_acc;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) G.o1[bx] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    }
  }
  if (dbg && tid == 0) G.dbg[6] = clock64();
  if (trace) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); G.trace[2] = t; G.trace[3] = EPI; }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
  if (dbg && tid == 160) G.dbg[7] = clock64();
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGtThreads)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ TcGemm G) {
  gemm_tile<BN, EPI>(&mapA, &mapB, G, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
}

// Several independent problems (the weight-gradient GEMMs of one step and the FiLM data gradient) in ONE launch: CTA i belongs to the group
// whose [start, start + tiles) range contains i.  Saves the launch + drain of four small kernels per step.
constexpr int kGtMaxGroup = 6;
struct TcGroup {
  CUtensorMap mapA[kGtMaxGroup], mapB[kGtMaxGroup];
  TcGemm g[kGtMaxGroup];
  int start[kGtMaxGroup + 1];
  int tiles_m[kGtMaxGroup], tiles_n[kGtMaxGroup], zcount[kGtMaxGroup], ksplit[kGtMaxGroup];
  int epi[kGtMaxGroup];            // TE_STORE or TE_DCOND
  int n;
};
template <int BN>
__global__ void __launch_bounds__(kGtThreads) gemm_tc_group_kernel(const __grid_constant__ TcGroup P) {
  int i = 0;
  const int lin = (int)blockIdx.x;
#pragma unroll
  for (int k = 1; k < kGtMaxGroup; ++k)
    if (k < P.n && lin >= P.start[k]) i = k;
  const int local = lin - P.start[i];
  const int tm = P.tiles_m[i], tn = P.tiles_n[i];
  const int zc = P.zcount[i], z = (local / (tm * tn)) % zc, ks = local / (tm * tn * zc);
  if (P.epi[i] == TE_DCOND) gemm_tile<BN, TE_DCOND>(&P.mapA[i], &P.mapB[i], P.g[i], local % tm, (local / tm) % tn, z);
  else gemm_tile<BN, TE_STORE>(&P.mapA[i], &P.mapB[i], P.g[i], local % tm, (local / tm) % tn, z, ks, P.ksplit[i]);
}

// ------------------------------------------------------------------------------------ small CUDA-core kernels
// xin = token embeddings (variant B only; x_emb == nullptr for variant A), cond = [time_emb[t] | basis_emb[basis]], ind[b, v*N+q] = (bit_q == v); all bf16
__global__ void gather_tc_kernel(int N, int E, const float* __restrict__ x_emb, const float* __restrict__ time_emb,
                                 const float* __restrict__ basis_emb, const uint16_t* __restrict__ xt,
                                 const int32_t* __restrict__ t, const int32_t* __restrict__ basis,
                                 __nv_bfloat16* __restrict__ xin, __nv_bfloat16* __restrict__ cond, __nv_bfloat16* __restrict__ ind) {
  const int64_t i = blockIdx.x;
  const uint32_t bits = xt[i];
  const float* te = time_emb + (int64_t)t[i] * E;
  const float* be = basis_emb + (int64_t)basis[i] * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    cond[i * 2 * E + e] = __float2bfloat16(te[e]);
    cond[i * 2 * E + E + e] = __float2bfloat16(be[e]);
  }
  if (x_emb) {
    for (int j = threadIdx.x; j < N * E; j += blockDim.x) {
      int q = j / E, e = j - q * E;
      xin[i * N * E + j] = __float2bfloat16(x_emb[((bits >> q) & 1u) * E + e]);
    }
  }
  if (threadIdx.x < 32) {
    int j = threadIdx.x, v = j / N, q = j - v * N;
    float f = (j < 2 * N && ((bits >> q) & 1u) == (uint32_t)v) ? 1.f : 0.f;
    ind[i * 32 + j] = __float2bfloat16(f);
  }
}

// Variant A front end (SS/model.py:56,70: input_proj = Linear(N, H) on x.float()): K = N is far too small for a tensor-core
// tile, so the layer is a CUDA-core elementwise pass with the same fused FiLM epilogue as TE_IN:
//   h0[b,c] = in_b[c] + sum_q bit_q(b) Win[c,q] ;  a0 = h0 (1 + gamma_0) + beta_0
__global__ void in_a_forward_kernel(int64_t B, int N, int H, const float* __restrict__ in_w, const float* __restrict__ in_b,
                                    const uint16_t* __restrict__ xt, const float* __restrict__ gb0, float* __restrict__ h0,
                                    __nv_bfloat16* __restrict__ a0) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * H) return;
  const int64_t b = e / H;
  const int c = (int)(e - b * H);
  const uint32_t bits = xt[b];
  float v = in_b[c];
  for (int q = 0; q < N; ++q) v += ((bits >> q) & 1u) ? in_w[c * N + q] : 0.f;
  h0[e] = v;
  a0[e] = __float2bfloat16(fmaf(v, 1.0f + gb0[b * 2 * H + c], gb0[b * 2 * H + H + c]));
}
// dWin[c,q] = sum_b dh0[b,c] bit_q(b) = S[N + q, c]  (S = ind^T . dh0, rows N.. = the "bit == 1" indicators)
__global__ void in_a_wgrad_kernel(int N, int H, const float* __restrict__ S, float* __restrict__ g_in_w) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H * N) return;
  const int c = e / N, q = e - c * N;
  g_in_w[e] = S[(int64_t)(N + q) * H + c];
}

__global__ void loss_finish_tc_kernel(const float* __restrict__ part, int n, float inv_total, float* __restrict__ out) {
  // one warp; fixed summation tree (lane-strided partial sums in double, then a butterfly) -> deterministic
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += (double)part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(s * inv_total);
}

// Variant B input-projection weight gradient without the B x (N E) x H GEMM: x_in[b, qE+e] = x_emb[bit_q(b)][e], so
//   dWin[h, qE+e] = sum_b dh0[b,h] x_in[b,qE+e] = S[q, h] x_emb[0][e] + S[N+q, h] x_emb[1][e],   S = ind^T . dh0 (rows v N + q)
// (S is computed anyway for the x_emb gradient).  Removes an 8.6 GFLOP GEMM and the 16 MB x_in gather at batch 8192.
__global__ void in_b_wgrad_kernel(int N, int E, int H, const float* __restrict__ S, const float* __restrict__ x_emb, float* __restrict__ g_in_w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)H * N * E) return;
  const int h = (int)(i / (N * E)), j = (int)(i - (int64_t)h * N * E), q = j / E, e = j - q * E;
  g_in_w[i] = S[(int64_t)q * H + h] * x_emb[e] + S[(int64_t)(N + q) * H + h] * x_emb[E + e];
}

// bias gradients: column sums of the bf16 dY arrays.  blockIdx.z splits the rows into chunks whose partial sums are added
// atomically into the (pre-zeroed) gradient -- the reduction over 1024+ rows is latency-bound, so it wants many blocks
struct ColsumTask { const __nv_bfloat16* p; int64_t ld; int cols; float* out; };
struct ColsumTasks { ColsumTask t[56]; };
__global__ void colsum_bf16_kernel(const __grid_constant__ ColsumTasks T, int64_t rows, int64_t rows_per_chunk) {
  const ColsumTask k = T.t[blockIdx.y];
  const int c = blockIdx.x * 64 + (threadIdx.x & 31) * 2;
  if (blockIdx.x * 64 >= k.cols) return;
  __shared__ float red[8][64];
  const int w = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.z * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float a0 = 0.f, a1 = 0.f;
  if (c < k.cols) {
#pragma unroll 4
    for (int64_t r = r0 + w; r < r1; r += 8) {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(k.p + r * k.ld + c);
      a0 += bf16_lo(u); a1 += bf16_hi(u);
    }
  }
  red[w][(threadIdx.x & 31) * 2] = a0; red[w][(threadIdx.x & 31) * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < k.cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(k.out + blockIdx.x * 64 + threadIdx.x, s);
  }
}

// x_emb gradient from S[v*N+q, h] = sum_b [bit_q(b) == v] dh0[b, h]:  g[v, e] = sum_q sum_h S[vN+q, h] Win[h, qE+e]
__global__ void xemb_grad_kernel(int N, int E, int H, const float* __restrict__ S, const float* __restrict__ in_w,
                                 float* __restrict__ g_xemb) {
  const int q = blockIdx.x, h0 = blockIdx.y * 16;
  for (int j = threadIdx.x; j < 2 * E; j += blockDim.x) {
    const int v = j / E, e = j - v * E;
    const float* s = S + (int64_t)(v * N + q) * H + h0;
    const float* w = in_w + (int64_t)h0 * (N * E) + q * E + e;
    float acc = 0.f;
#pragma unroll
    for (int h = 0; h < 16; ++h) acc = fmaf(s[h], w[(int64_t)h * (N * E)], acc);
    atomicAdd(g_xemb + j, acc);
  }
}

// torch Adam / AdamW arithmetic (same as train.cu's adam_kernel) with the step count read from device memory (so
// the whole training step can be replayed from a CUDA graph) and a bf16 shadow of the updated parameters.
// step_dev[0] = completed steps, step_dev[1] = block-arrival counter: the last block to finish bumps the step.
// 4 elements per thread (n4 = n / 4 vectors, the flat buffer is 16-byte aligned); the tail runs scalar.
__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps,
                                          float wd, int decoupled, float gscale, float bc1, float bc2_sqrt) {
  float grad = g * gscale, param = p;
  if (decoupled) param *= (1.0f - lr * wd);            // torch AdamW: param.mul_(1 - lr*wd)
  else if (wd != 0.f) grad += wd * param;              // torch Adam: grad.add(param, alpha=wd)
  float mi = m + (grad - m) * (1.0f - b1);             // exp_avg.lerp_(grad, 1-beta1)
  float vi = v * b2 + (1.0f - b2) * grad * grad;       // exp_avg_sq.mul_(b2).addcmul_(grad, grad, 1-b2)
  m = mi; v = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p = param - (lr / bc1) * (mi / denom);
  return p;
}
__global__ void __launch_bounds__(256)
adam_tc_kernel(float* __restrict__ p, __nv_bfloat16* __restrict__ shadow, const float* __restrict__ g,
               float* __restrict__ m, float* __restrict__ v, int64_t n, int64_t* __restrict__ step_dev,
               float lr, float b1, float b2, float eps, float wd, int decoupled, float gscale) {
  __shared__ float s_bc1, s_bc2s;
  if (threadIdx.x == 0) {
    const double st = (double)(*((volatile int64_t*)step_dev) + 1);
    s_bc1 = (float)(1.0 - pow((double)b1, st));
    s_bc2s = (float)sqrt(1.0 - pow((double)b2, st));
  }
  __syncthreads();
  const float bc1 = s_bc1, bc2_sqrt = s_bc2s;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr, b1, b2, eps, wd, decoupled, gscale, bc1, bc2_sqrt);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr, b1, b2, eps, wd, decoupled, gscale, bc1, bc2_sqrt);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr, b1, b2, eps, wd, decoupled, gscale, bc1, bc2_sqrt);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr, b1, b2, eps, wd, decoupled, gscale, bc1, bc2_sqrt);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16(pp.x, pp.y), pack_bf16(pp.z, pp.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float pv = p[i], mv = m[i], vv = v[i];
    adam_one(pv, g[i], mv, vv, lr, b1, b2, eps, wd, decoupled, gscale, bc1, bc2_sqrt);
    p[i] = pv; m[i] = mv; v[i] = vv;
    if (shadow) shadow[i] = __float2bfloat16(pv);
  }
  __syncthreads();                      // every thread of this block has read the step (through s_bc*) long ago
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long arrived = atomicAdd((unsigned long long*)(step_dev + 1), 1ull);
    if (arrived == (unsigned long long)gridDim.x - 1ull) { step_dev[1] = 0; step_dev[0] += 1; }
  }
}

__global__ void step_inc_kernel(int64_t* step_dev) { if (threadIdx.x == 0 && blockIdx.x == 0) step_dev[0] += 1; }

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}

// ------------------------------------------------------------------------------------ host side
// bf16 tensor {inner, rows, batch}; box {64, box_rows, 1}; 128-byte swizzle; out-of-bounds elements read as zero
static int make_map3(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t batch, int64_t row_stride,
                     int64_t batch_stride, int box_rows) {
  EncodeTiledFn fn;
  DDQST_TRY(get_encode_fn(&fn));
  DDQST_REQUIRE(((uintptr_t)base & 15u) == 0 && (row_stride * 2) % 16 == 0 && (batch_stride * 2) % 16 == 0, DDQST_EUNSUPPORTED,
                "tensor-core training needs 16-byte aligned operands (base %p, row stride %lld, batch stride %lld elements)",
                base, (long long)row_stride, (long long)batch_stride);
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DDQST_REQUIRE(r == CUDA_SUCCESS, DDQST_ECUDA, "cuTensorMapEncodeTiled(3d) failed with CUresult %d (inner=%lld rows=%lld batch=%lld)",
                (int)r, (long long)inner, (long long)rows, (long long)batch);
  return DDQST_OK;
}

// one operand of a GEMM as the host describes it: the matrix [mn, k] per batch entry, stored either way round
struct HostOperand {
  const void* base; int mn_major; int64_t mn, k, batch, ld, batch_stride; int kmod, zmul;
};
static HostOperand op_k(const void* base, int64_t mn, int64_t k, int64_t ld) { return HostOperand{base, 0, mn, k, 1, ld, mn * ld, 0, 0}; }
static HostOperand op_mn(const void* base, int64_t mn, int64_t k, int64_t ld) { return HostOperand{base, 1, mn, k, 1, ld, k * ld, 0, 0}; }

// debugging aid: when a trace buffer is registered, GEMM launch i of a step writes its timestamps to buf[4*i .. 4*i+3]
static long long* g_trace_buf = nullptr;
static int g_trace_idx = 0, g_trace_cap = 0;

// DDQST_TC_PDL=0 launches the GEMMs fully serialised (debugging aid)
static bool tc_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DDQST_TC_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

template <int BN, int EPI>
static int launch_gemm_bn(const HostOperand& A, const HostOperand& B, TcGemm g, int zcount, cudaStream_t s) {
  CUtensorMap ma, mb;
  if (A.mn_major) DDQST_TRY(make_map3(&ma, A.base, A.mn, A.k, A.batch, A.ld, A.batch_stride, 64));
  else DDQST_TRY(make_map3(&ma, A.base, A.k, A.mn, A.batch, A.ld, A.batch_stride, 128));
  if (B.mn_major) DDQST_TRY(make_map3(&mb, B.base, B.mn, B.k, B.batch, B.ld, B.batch_stride, 64));
  else DDQST_TRY(make_map3(&mb, B.base, B.k, B.mn, B.batch, B.ld, B.batch_stride, BN));
  g.a = TcOperand{A.mn_major, A.kmod, A.zmul};
  g.b = TcOperand{B.mn_major, B.kmod, B.zmul};
  if (g_trace_buf && g_trace_idx < g_trace_cap) g.trace = g_trace_buf + 4 * (g_trace_idx++);
  static bool attr_set = false;
  if (!attr_set) {
    DDQST_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes<BN>()));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((g.M + 127) / 128), (unsigned)((g.N + BN - 1) / BN), (unsigned)zcount);
  cfg.blockDim = dim3(kGtThreads);
  cfg.dynamicSmemBytes = gt_smem_bytes<BN>();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  DDQST_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI>, ma, mb, g));
  return DDQST_OK;
}

// ---- grouped TE_STORE launch (BN = 64)
struct GroupBuilder {
  TcGroup grp{};
  int total = 0;
  int bn = 64;                     // N tile of every problem in the group: 128 for large batches (set before the first add())
  bool split_k = false;            // outputs are pre-zeroed gradient buffers: long reductions may be split over CTAs
  int add(const HostOperand& A, const HostOperand& B, TcGemm g, int zcount, int epi = TE_STORE) {
    const int i = grp.n;
    grp.epi[i] = epi;
    if (i >= kGtMaxGroup) { set_error("too many problems in one grouped GEMM launch"); return DDQST_EINVAL_SHAPE; }
    if (A.mn_major) DDQST_TRY(make_map3(&grp.mapA[i], A.base, A.mn, A.k, A.batch, A.ld, A.batch_stride, 64));
    else DDQST_TRY(make_map3(&grp.mapA[i], A.base, A.k, A.mn, A.batch, A.ld, A.batch_stride, 128));
    if (B.mn_major) DDQST_TRY(make_map3(&grp.mapB[i], B.base, B.mn, B.k, B.batch, B.ld, B.batch_stride, 64));
    else DDQST_TRY(make_map3(&grp.mapB[i], B.base, B.k, B.mn, B.batch, B.ld, B.batch_stride, bn));
    g.a = TcOperand{A.mn_major, A.kmod, A.zmul};
    g.b = TcOperand{B.mn_major, B.kmod, B.zmul};
    if (g_trace_buf && g_trace_idx < g_trace_cap) g.trace = g_trace_buf + 4 * (g_trace_idx++);
    grp.g[i] = g;
    grp.tiles_m[i] = (g.M + 127) / 128;
    grp.tiles_n[i] = (g.N + bn - 1) / bn;
    grp.zcount[i] = zcount;
    // split-K for the long reductions (weight gradients reduce over the batch): without it a problem is tiles_m x tiles_n CTAs that
    // each walk the whole batch -- at batch 8192 that is 224 CTAs of 128 K-blocks next to 500 short ones, two badly filled waves.
    // ~32 K-blocks per CTA; the partial tiles are added atomically into the (pre-zeroed) gradient buffer.
    int ks = 1;
    if (epi == TE_STORE && split_k && g.bias == nullptr && g.b0 == nullptr) {
      const int kbt = (g.K + 63) / 64;
      ks = (kbt + 31) / 32;
      while (ks > 1 && ((kbt + ks - 1) / ks) * (ks - 1) >= kbt) --ks;     // every split non-empty
    }
    grp.ksplit[i] = ks;
    grp.start[i] = total;
    total += grp.tiles_m[i] * grp.tiles_n[i] * zcount * ks;
    grp.start[i + 1] = total;
    grp.n = i + 1;
    return DDQST_OK;
  }
  template <int BN>
  int launch_bn(cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
      DDQST_CUDA_OK(cudaFuncSetAttribute(gemm_tc_group_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes<BN>()));
      attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)total);
    cfg.blockDim = dim3(kGtThreads);
    cfg.dynamicSmemBytes = gt_smem_bytes<BN>();
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
    DDQST_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_group_kernel<BN>, grp));
    return DDQST_OK;
  }
  int launch(cudaStream_t s) { return bn == 256 ? launch_bn<256>(s) : (bn == 128 ? launch_bn<128>(s) : launch_bn<64>(s)); }
};

// BN = 128 when the grid still covers the machine (or N is not a multiple of 64-wide tiles anyway), else 64
template <int EPI>
static int launch_gemm(const HostOperand& A, const HostOperand& B, const TcGemm& g, int zcount, cudaStream_t s) {
  const int64_t tiles128 = (int64_t)((g.M + 127) / 128) * ((g.N + 127) / 128) * zcount;
  const bool fused = EPI != TE_STORE && EPI != TE_DCOND && EPI != TE_HEAD;
  const bool can128 = !fused || g.N % 128 == 0;
  static int thr = -1;
  if (thr < 0) { const char* e = getenv("DDQST_TC_BN128_TILES"); thr = e ? atoi(e) : 96; }      // tuning knob: min 128x128 tiles for BN = 128
  if (can128 && EPI != TE_HEAD && tiles128 >= thr) return launch_gemm_bn<128, EPI>(A, B, g, zcount, s);
  return launch_gemm_bn<64, EPI>(A, B, g, zcount, s);
}

// ---- workspace
struct TcWs {
  // bf16 (element offsets into the bf16 region), fp32 (element offsets into the fp32 region)
  int64_t xin, cond, ind, act, hL, dz, dgb, dh0, dlog, d1s, hs, dsv, h0s, dsp, gbh, dt, bf_total;
  int64_t gb, h, z1, z2, logits, dres, S, loss_part, f_total;
};
static void tc_ws_layout(const ddqst_dims* d, int64_t B, TcWs* w) {
  const int64_t N = d->num_qubits, E = d->embed_dim, H = d->hidden_dim, L = d->num_blocks;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off = align_up(off + n, 128); return o; };
  w->xin = take(B * N * E); w->cond = take(B * 2 * E); w->ind = take(B * 32); w->act = take(2 * L * B * H); w->hL = take(B * H);
  w->dz = take(2 * L * B * H); w->dgb = take(L * B * 2 * H); w->dh0 = take(B * H); w->dlog = take(B * 32);
  const int64_t priv = (B + 127) / 128 * 128 * H;     // one tile-private array of the fused pass (rows padded to whole tiles)
  w->d1s = take(L * priv); w->hs = take(L * priv); w->dsv = take(L * priv); w->h0s = take(priv); w->dsp = take(priv); w->gbh = take(L * 2 * priv); w->dt = take(H * 64);
  w->bf_total = off;
  off = 0;
  w->gb = take(L * B * 2 * H); w->h = take((L + 1) * B * H); w->z1 = take(L * B * H); w->z2 = take(L * B * H);
  w->logits = take(B * 2 * N); w->dres = take(B * H); w->S = take(32 * H); w->loss_part = take(4 * ((B + 127) / 128) + 8);
  w->f_total = off;
}

int64_t train_tc_workspace_bytes(const ddqst_dims* d, int64_t batch) {
  if (validate_dims(d) != DDQST_OK) return -1;
  TcWs w;
  tc_ws_layout(d, batch < 1 ? 1 : batch, &w);
  return w.bf_total * 2 + w.f_total * 4 + 1024;
}

int train_tc_abort_fetch() { return tc_abort_fetch(); }

static int train_tc_supported(const ddqst_dims* d, const ParamLayout& pr) {
  DDQST_REQUIRE(d->hidden_dim % 64 == 0 && d->embed_dim % 16 == 0 && d->num_qubits <= 15 && 2 * d->num_blocks <= kGtMaxZ,
                DDQST_EUNSUPPORTED, "tensor-core training needs hidden_dim %% 64 == 0, embed_dim %% 16 == 0, num_qubits <= 15, num_blocks <= 16");
  bool ok = (d->variant == DDQST_VARIANT_A || pr.in_w % 8 == 0) && pr.head_w % 8 == 0;
  for (int l = 0; l < d->num_blocks; ++l) ok = ok && pr.film_w[l] % 8 == 0 && pr.w1[l] % 8 == 0 && pr.w2[l] % 8 == 0;
  DDQST_REQUIRE(ok, DDQST_EUNSUPPORTED, "parameter offsets are not 16-byte aligned in the bf16 shadow");
  return DDQST_OK;
}

#include "train_fused.cuh"

// collapsed input table as an MMA operand (hi/lo split, K = 32), rebuilt from the CURRENT parameters every step:
// h0 = c0 + sum_q x_q D[q] with (variant B) c0 = b + sum_q W_q . emb(0), D[q] = W_q . (emb(1) - emb(0)); (variant A) D[q] = W[:, q].
// Dt[h][k]: k < N hi(D[k][h]); k == N hi(c0[h]); 16 <= k < 16+N lo(D[k-16][h]); k == 16+N lo(c0[h]); else 0.
__global__ void ft_build_dt_kernel(int variant, int N, int E, int H, const float* __restrict__ x_emb, const float* __restrict__ in_w,
                                   const float* __restrict__ in_b, __nv_bfloat16* __restrict__ dt) {
  // one warp per hidden unit h: lanes stride over the embedding index (coalesced reads of W[h, q*E .. q*E+E))
  const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (h >= H) return;
  float c = in_b[h];
  float D[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) D[q] = 0.f;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (q < N) {
      if (variant == DDQST_VARIANT_B) {
        float d0 = 0.f, d1 = 0.f;
        const float* w = in_w + (int64_t)h * N * E + q * E;
        for (int e = lane; e < E; e += 32) { const float wv = w[e]; d0 = fmaf(wv, x_emb[e], d0); d1 = fmaf(wv, x_emb[E + e], d1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xFFFFFFFFu, d0, o); d1 += __shfl_xor_sync(0xFFFFFFFFu, d1, o); }
        c += d0;
        D[q] = d1 - d0;
      } else {
        D[q] = in_w[(int64_t)h * N + q];
      }
    }
  }
  __nv_bfloat16* row = dt + (int64_t)h * 64;
#pragma unroll
  for (int kk2 = 0; kk2 < 2; ++kk2) {
    const int k = lane + 32 * kk2;
    const int kk = k >= 16 ? k - 16 : k;
    float out = 0.f;
    if (k < 32 && kk <= N) {
      float v = c;
#pragma unroll
      for (int q = 0; q < 16; ++q) if (kk == q && q < N) v = D[q];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      out = k < 16 ? __bfloat162float(hi) : v - __bfloat162float(hi);
    }
    row[k] = __float2bfloat16_rn(out);
  }
}

// which forward + data-gradient path a step takes: -1 = by batch size (default), 0 = per-layer GEMM launches, 1 = fused kernel.
// Initialised from DDQST_TRAIN_FUSED; ddqst_debug_train_path() overrides it (tests compare the two paths in one process).
static int g_train_fused_mode = -2;
static int train_fused_mode() {
  if (g_train_fused_mode == -2) { const char* e = getenv("DDQST_TRAIN_FUSED"); g_train_fused_mode = e ? atoi(e) : -1; }
  return g_train_fused_mode;
}
// The fused kernel keeps a 256-row tile on one SM pair, so a batch of B rows occupies B / 128 SMs: below ~3k rows the per-layer
// path, which spreads every GEMM over 64+ SMs, has the shorter critical path (measured at 1024 / 2048 / 4096 / 8192 rows:
// per-layer 0.289 / 0.357 / 0.549 / 1.015 ms, fused 0.343 / 0.384 / 0.474 / 0.629 ms; profiles/r2_train_step.json)
static int train_fused_min_batch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DDQST_TRAIN_FUSED_MIN_BATCH"); v = e ? atoi(e) : 3072; }
  return v;
}
static bool train_fused_enabled(int64_t batch) {
  const int m = train_fused_mode();
  return m == 1 || (m == -1 && batch >= train_fused_min_batch());
}
static bool train_fused_supported(const ddqst_dims* d) {
  const int H = d->hidden_dim;
  if (!(H == 128 || H == 256 || H == 512)) return false;
  if (d->num_qubits > 15) return false;
  const int smem = 1024 + (H / 64) * 16384 + kFtRing * kFtStage + 2 * d->num_blocks * H * 4 + 256;
  return smem <= 227 * 1024;
}

template <int H>
static int launch_train_fused(const ddqst_dims* d, const ParamLayout& pr, const __nv_bfloat16* shadow, const __nv_bfloat16* dt,
                              int64_t blk_stride, __nv_bfloat16* act, __nv_bfloat16* hL, __nv_bfloat16* dz, __nv_bfloat16* dh0,
                              FusedParams P, cudaStream_t s) {
  const int L = d->num_blocks, N = d->num_qubits;
  CUtensorMap m_dt, m_w1k, m_w2k, m_w1m, m_w2m, m_hk, m_hm, m_wf, m_act, m_hL, m_dz, m_dh0;
  const int64_t bs = blk_stride > 0 ? blk_stride : (int64_t)H * H;
  DDQST_TRY(make_map3(&m_dt, dt, 64, H, 1, 64, (int64_t)H * 64, 64));
  DDQST_TRY(make_map3(&m_w1k, shadow + pr.w1[0], H, H, L, H, bs, 64));
  DDQST_TRY(make_map3(&m_w2k, shadow + pr.w2[0], H, H, L, H, bs, 64));
  DDQST_TRY(make_map3(&m_w1m, shadow + pr.w1[0], H, H, L, H, bs, 128));
  DDQST_TRY(make_map3(&m_w2m, shadow + pr.w2[0], H, H, L, H, bs, 128));
  DDQST_TRY(make_map3(&m_hk, shadow + pr.head_w, H, 2 * N, 1, H, (int64_t)2 * N * H, P.head_pad / 2));
  DDQST_TRY(make_map3(&m_hm, shadow + pr.head_w, H, 2 * N, 1, H, (int64_t)2 * N * H, P.head_pad));
  {  // FiLM weights [L][2H rows][2E], K-major boxes of 64 x 64
    const int E2 = 2 * d->embed_dim;
    const int64_t fbs = L > 1 ? pr.film_w[1] - pr.film_w[0] : (int64_t)2 * H * E2;
    DDQST_TRY(make_map3(&m_wf, shadow + pr.film_w[0], E2, 2 * H, L, E2, fbs, 64));
  }
  // outputs stored by TMA straight from the A-operand buffer: row-major [z][B][H] bf16, boxes of 32 rows x 64 columns, one per warp (rows past B are clipped)
  DDQST_TRY(make_map3(&m_act, act, H, P.B, 2 * L, H, P.B * H, 32));
  DDQST_TRY(make_map3(&m_hL, hL, H, P.B, 1, H, P.B * H, 32));
  DDQST_TRY(make_map3(&m_dz, dz, H, P.B, 2 * L, H, P.B * H, 32));
  DDQST_TRY(make_map3(&m_dh0, dh0, H, P.B, 1, H, P.B * H, 32));
  const int smem = ft_smem_bytes<H>(L);
  auto kern = train_fused_kernel<H>;
  static bool attr_set = false;
  if (!attr_set) { DDQST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
  const int64_t pairs = (P.n_tiles + 1) / 2;
  const int max_pairs = num_sms() / 2;
  const int grid_pairs = (int)(pairs < max_pairs ? pairs : max_pairs);
  P.iters = (int)((pairs + grid_pairs - 1) / grid_pairs);
  kern<<<2 * grid_pairs, kFtThreads, smem, s>>>(m_dt, m_w1k, m_w2k, m_w1m, m_w2m, m_hk, m_hm, m_wf, m_act, m_hL, m_dz, m_dh0, P);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

static int train_tc_run(const ddqst_dims* d, const float* params, const __nv_bfloat16* shadow, const uint16_t* xt,
                        const uint16_t* x0, const int32_t* t, const int32_t* basis, int64_t B, float loss_scale,
                        float* grads, float* loss_out, void* workspace, int64_t ws_bytes, cudaStream_t s) {
  ParamLayout pr;
  DDQST_TRY(param_layout(d, &pr));
  DDQST_TRY(train_tc_supported(d, pr));
  TcWs w;
  tc_ws_layout(d, B, &w);
  DDQST_REQUIRE(workspace && ws_bytes >= w.bf_total * 2 + w.f_total * 4 + 1024, DDQST_EWORKSPACE,
                "tensor-core train step needs %lld workspace bytes, got %lld", (long long)(w.bf_total * 2 + w.f_total * 4 + 1024), (long long)ws_bytes);
  const int N = d->num_qubits, E = d->embed_dim, H = d->hidden_dim, L = d->num_blocks, XIN = N * E;
  const bool var_b = d->variant == DDQST_VARIANT_B;
  char* base = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* bf = (__nv_bfloat16*)base;
  float* fp = (float*)(base + align_up(w.bf_total * 2, 1024));
  __nv_bfloat16 *xin = bf + w.xin, *cond = bf + w.cond, *ind = bf + w.ind, *act = bf + w.act, *hL = bf + w.hL, *dz = bf + w.dz,
                *dgb = bf + w.dgb, *dh0 = bf + w.dh0, *dlog = bf + w.dlog;
  float *gb = fp + w.gb, *h = fp + w.h, *z1 = fp + w.z1, *z2 = fp + w.z2, *logits = fp + w.logits, *dres = fp + w.dres,
        *S = fp + w.S, *loss_part = fp + w.loss_part;
  const int64_t BH = B * H, BG = B * 2 * H;
  const int64_t blk_stride = L > 1 ? pr.film_w[1] - pr.film_w[0] : 0;

  auto base_gemm = [&](int M, int Nn, int K) {
    TcGemm g{};
    g.M = M; g.N = Nn; g.K = K; g.ld = H; g.ldg = 2 * H; g.E = E; g.nq = N;
    return g;
  };

  const bool fused = train_fused_enabled(B) && train_fused_supported(d);
  // the fused pass builds h0 from the collapsed input table and gets dWin from S: it never needs the [B, N E] token-embedding matrix
  gather_tc_kernel<<<(unsigned)B, 128, 0, s>>>(N, E, var_b && !fused ? params + pr.x_emb : nullptr, params + pr.time_emb, params + pr.basis_emb,
                                               xt, t, basis, xin, cond, ind);
  DDQST_LAUNCH_OK();

  if (fused) {
    // ---------------------------------------------------------------- fused forward + data-gradient pass (train_fused.cuh)
    __nv_bfloat16 *d1s = bf + w.d1s, *hs = bf + w.hs, *dsv = bf + w.dsv, *h0s = bf + w.h0s, *dsp = bf + w.dsp, *gbh = bf + w.gbh, *dt = bf + w.dt;
    const int64_t priv = (B + 127) / 128 * 128 * H;
    // FiLM GEMMs inside the fused kernel when cond's K = 2E is a whole number of 128-column chunks and the embedding rows are 16-byte
    // aligned in the bf16 shadow; otherwise a separate launch writes the same tile-private arrays
    static int film_env = -1;
    if (film_env < 0) { const char* e = getenv("DDQST_TRAIN_FILM_IN_KERNEL"); film_env = (e && e[0] == '0') ? 0 : 1; }
    const bool film_in_kernel = film_env == 1 && (2 * E) % 128 == 0 && 2 * E <= H && pr.time_emb % 8 == 0 && pr.basis_emb % 8 == 0 && E % 8 == 0;
    if (!film_in_kernel)
    {  // gamma|beta of every block, bf16, tile-private layout: gbh[l] = cond . Wfilm_l^T + bfilm_l
      TcGemm g = base_gemm((int)B, 2 * H, 2 * E);
      g.b0 = gbh; g.ld = 2 * H; g.ldg = priv; g.flag = 2; g.bias = params + pr.film_b[0]; g.bias_zstride = blk_stride;
      for (int l = 0; l < L; ++l) g.out_zoff[l] = (int64_t)l * 2 * priv;
      HostOperand A = op_k(cond, B, 2 * E, 2 * E);
      HostOperand Bo{shadow + pr.film_w[0], 0, 2 * H, 2 * E, L, 2 * E, blk_stride > 0 ? blk_stride : (int64_t)2 * H * 2 * E, 0, 1};
      DDQST_TRY(launch_gemm<TE_STORE>(A, Bo, g, L, s));
    }
    ft_build_dt_kernel<<<(H * 32 + 127) / 128, 128, 0, s>>>(d->variant, N, E, H, var_b ? params + pr.x_emb : nullptr, params + pr.in_w,
                                                       params + pr.in_b, dt);
    DDQST_LAUNCH_OK();
    FusedParams P{};
    P.N = N; P.L = L; P.head_pad = (int)align_up(2 * N, 16); P.B = B; P.n_tiles = (B + 127) / 128;
    P.xt = xt; P.x0 = x0; P.gb = gbh; P.b1 = params + pr.b1[0]; P.b2 = params + pr.b2[0]; P.bias_stride = blk_stride;
    P.head_b = params + pr.head_b; P.scale = loss_scale / (float)(B * N);
    P.film_kc = film_in_kernel ? (2 * E) / 128 : 0; P.E = E; P.t = t; P.basis = basis;
    P.temb = shadow + pr.time_emb; P.bemb = shadow + pr.basis_emb; P.film_b = params + pr.film_b[0]; P.film_b_stride = blk_stride;
    P.dgb = dgb; P.dlog = dlog; P.d1s = d1s; P.hs = hs; P.dsv = dsv; P.h0s = h0s; P.dsp = dsp; P.priv_elems = priv;
    P.loss_part = loss_part;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("DDQST_FT_DEBUG"); dbg = e ? atoi(e) : 0; } P.dbg = dbg; }
    switch (H) {
      case 128: DDQST_TRY(launch_train_fused<128>(d, pr, shadow, dt, blk_stride, act, hL, dz, dh0, P, s)); break;
      case 256: DDQST_TRY(launch_train_fused<256>(d, pr, shadow, dt, blk_stride, act, hL, dz, dh0, P, s)); break;
      default: DDQST_TRY(launch_train_fused<512>(d, pr, shadow, dt, blk_stride, act, hL, dz, dh0, P, s)); break;
    }
    loss_finish_tc_kernel<<<1, 32, 0, s>>>(loss_part, (int)(4 * P.n_tiles), 1.0f / (float)(B * N), loss_out);
    DDQST_LAUNCH_OK();
    DDQST_CUDA_OK(cudaMemsetAsync(grads, 0, sizeof(float) * pr.total, s));
  } else {
  // ------------------------------------------------------------------ forward
  {  // gb[l] = cond . Wfilm_l^T + bfilm_l, all blocks in one launch
    TcGemm g = base_gemm((int)B, 2 * H, 2 * E);
    g.o0 = gb; g.ld = 2 * H; g.bias = params + pr.film_b[0]; g.bias_zstride = blk_stride;
    for (int l = 0; l < L; ++l) g.out_zoff[l] = (int64_t)l * BG;
    HostOperand A = op_k(cond, B, 2 * E, 2 * E);
    HostOperand Bo{shadow + pr.film_w[0], 0, 2 * H, 2 * E, L, 2 * E, blk_stride > 0 ? blk_stride : (int64_t)2 * H * 2 * E, 0, 1};
    DDQST_TRY(launch_gemm<TE_STORE>(A, Bo, g, L, s));
  }
  if (var_b) {  // h0, a0
    TcGemm g = base_gemm((int)B, H, XIN);
    g.bias = params + pr.in_b; g.o0 = h; g.f1 = gb; g.b0 = act;
    DDQST_TRY(launch_gemm<TE_IN>(op_k(xin, B, XIN, XIN), op_k(shadow + pr.in_w, H, XIN, XIN), g, 1, s));
  } else {
    in_a_forward_kernel<<<(unsigned)((B * H + 255) / 256), 256, 0, s>>>(B, N, H, params + pr.in_w, params + pr.in_b, xt, gb, h, act);
    DDQST_LAUNCH_OK();
  }
  for (int l = 0; l < L; ++l) {
    {
      TcGemm g = base_gemm((int)B, H, H);
      g.bias = params + pr.b1[l]; g.o0 = z1 + l * BH; g.b0 = act + (2 * l + 1) * BH;
      DDQST_TRY(launch_gemm<TE_W1>(op_k(act + 2 * l * BH, B, H, H), op_k(shadow + pr.w1[l], H, H, H), g, 1, s));
    }
    {
      TcGemm g = base_gemm((int)B, H, H);
      g.bias = params + pr.b2[l]; g.f0 = h + l * BH; g.o0 = z2 + l * BH; g.o1 = h + (l + 1) * BH;
      g.flag = l + 1 < L; g.f1 = gb + (int64_t)(l + 1) * BG; g.b0 = l + 1 < L ? act + 2 * (l + 1) * BH : hL;
      DDQST_TRY(launch_gemm<TE_W2>(op_k(act + (2 * l + 1) * BH, B, H, H), op_k(shadow + pr.w2[l], H, H, H), g, 1, s));
    }
  }
  const int head_ctas = (int)((B + 127) / 128);
  {  // logits, loss parts, dlogits
    TcGemm g = base_gemm((int)B, 2 * N, H);
    g.bias = params + pr.head_b; g.o0 = logits; g.o1 = loss_part; g.b0 = dlog; g.x0 = x0;
    g.scale = loss_scale / (float)(B * N);
    DDQST_TRY(launch_gemm<TE_HEAD>(op_k(hL, B, H, H), op_k(shadow + pr.head_w, 2 * N, H, H), g, 1, s));
  }
  loss_finish_tc_kernel<<<1, 32, 0, s>>>(loss_part, head_ctas, 1.0f / (float)(B * N), loss_out);
  DDQST_LAUNCH_OK();

  // ------------------------------------------------------------------ backward: data gradients
  // embedding gradients are accumulated with atomics; the alignment padding between tensors must read as zero too
  DDQST_CUDA_OK(cudaMemsetAsync(grads, 0, sizeof(float) * pr.total, s));
  {  // dz2_{L-1} = (dlogits . Whead) * silu'(z2_{L-1})
    TcGemm g = base_gemm((int)B, H, 2 * N);
    g.f0 = z2 + (L - 1) * BH; g.o0 = dres; g.b0 = dz + (2 * (L - 1) + 1) * BH;
    DDQST_TRY(launch_gemm<TE_BHEAD>(op_k(dlog, B, 2 * N, 32), op_mn(shadow + pr.head_w, H, 2 * N, H), g, 1, s));
  }
  for (int l = L - 1; l >= 0; --l) {
    {  // dz1_l = (dz2_l . W2_l) * silu'(z1_l)
      TcGemm g = base_gemm((int)B, H, H);
      g.f0 = z1 + l * BH; g.b0 = dz + 2 * l * BH;
      DDQST_TRY(launch_gemm<TE_BW2>(op_k(dz + (2 * l + 1) * BH, B, H, H), op_mn(shadow + pr.w2[l], H, H, H), g, 1, s));
    }
    {  // da_l = dz1_l . W1_l -> dgb_l, dh_l -> dz2_{l-1} or dh0
      TcGemm g = base_gemm((int)B, H, H);
      g.f0 = gb + (int64_t)l * BG; g.f1 = h + l * BH; g.f3 = dres; g.b1 = dgb + (int64_t)l * BG;
      g.flag = l > 0;
      if (l > 0) { g.f2 = z2 + (l - 1) * BH; g.o0 = dres; g.b0 = dz + (2 * (l - 1) + 1) * BH; }
      else g.b0 = dh0;
      DDQST_TRY(launch_gemm<TE_BW1>(op_k(dz + 2 * l * BH, B, H, H), op_mn(shadow + pr.w1[l], H, H, H), g, 1, s));
    }
  }
  }   // per-layer (non-fused) path
  // ------------------------------------------------------------------ backward: weight gradients (dY^T . X over the batch)
  // six independent problems (five weight gradients + the FiLM data gradient), one grouped launch
  {
    GroupBuilder grp;
    {
      // 128 x 128 tiles once the reduction (= batch) is long enough that operand ingest, not the tile count, bounds the launch:
      // half the A traffic per FLOP (measured: step 0.713 -> 0.629 ms at batch 8192, 0.354 -> 0.343 ms at 1024).  Needs every N of the group to be a multiple of 64.
      static int thr = -1;
      if (thr < 0) { const char* e = getenv("DDQST_TC_GROUP_BN128_BATCH"); thr = e ? atoi(e) : 1024; }
      if (B >= thr && E % 32 == 0) grp.bn = 128;
      // 128 x 256 tiles (DDQST_TC_GROUP_BN256_BATCH=<batch>): built and measured equal to 128 x 128 at batch 8192 (0.569 vs 0.574 ms per step).
      // The grouped launch moves ~1 GB of operands through L2 per step at that batch (every CTA re-reads its A and B panels), i.e. it is
      // L2-bandwidth-bound chip-wide; wider single-CTA tiles trade traffic for fewer CTAs.  Cutting it needs cluster multicast of the
      // shared panels -- left for later; off by default.
      static int thr256 = -1;
      if (thr256 < 0) { const char* e = getenv("DDQST_TC_GROUP_BN256_BATCH"); thr256 = e ? atoi(e) : (1 << 30); }
      if (B >= thr256 && E % 128 == 0 && H % 256 == 0) grp.bn = 256;
      static int sk = -1;
      // measured at batch 8192 (step 0.584 ms without, 0.618 ms with): the atomic epilogues cost more than the better wave
      // balance returns, so split-K stays an opt-in (DDQST_TC_SPLIT_K=1)
      if (sk < 0) { const char* e = getenv("DDQST_TC_SPLIT_K"); sk = (e && e[0] == '1') ? 1 : 0; }
      grp.split_k = sk == 1;
      if (grp.split_k) DDQST_CUDA_OK(cudaMemsetAsync(S, 0, sizeof(float) * 32 * H, s));     // S = ind^T . dh0 is accumulated too
    }
    {  // dcond = sum_l dgb_l . Wfilm_l, scattered into the time / basis embedding gradients; the sum over blocks is a
       // split-K over z (one block per z): the epilogue accumulates with atomics anyway
      TcGemm g = base_gemm((int)B, 2 * E, 2 * H);
      g.o0 = grads + pr.time_emb; g.o1 = grads + pr.basis_emb; g.i0 = t; g.i1 = basis;
      HostOperand A{dgb, 0, B, 2 * H, L, 2 * H, BG, 0, 1};
      HostOperand Bo{shadow + pr.film_w[0], 1, 2 * E, 2 * H, L, 2 * E, blk_stride > 0 ? blk_stride : (int64_t)2 * H * 2 * E, 0, 1};
      DDQST_TRY(grp.add(A, Bo, g, L, TE_DCOND));
    }
    {  // W1_l, W2_l for every block: z = 2l (dz1_l, a_l), 2l+1 (dz2_l, u_l)
      TcGemm g = base_gemm(H, H, (int)B);
      g.o0 = grads;
      for (int l = 0; l < L; ++l) { g.out_zoff[2 * l] = pr.w1[l]; g.out_zoff[2 * l + 1] = pr.w2[l]; }
      HostOperand A{dz, 1, H, B, 2 * L, H, BH, 0, 1};
      HostOperand Bo{act, 1, H, B, 2 * L, H, BH, 0, 1};
      DDQST_TRY(grp.add(A, Bo, g, 2 * L));
    }
    {  // Wfilm_l
      TcGemm g = base_gemm(2 * H, 2 * E, (int)B);
      g.o0 = grads; g.ld = 2 * E;
      for (int l = 0; l < L; ++l) g.out_zoff[l] = pr.film_w[l];
      HostOperand A{dgb, 1, 2 * H, B, L, 2 * H, BG, 0, 1};
      HostOperand Bo{cond, 1, 2 * E, B, 1, 2 * E, B * 2 * E, 0, 0};
      DDQST_TRY(grp.add(A, Bo, g, L));
    }
    if (var_b && !fused) {  // Win (per-layer path; the fused path derives it from S below)
      TcGemm g = base_gemm(H, XIN, (int)B);
      g.o0 = grads + pr.in_w; g.ld = XIN;
      DDQST_TRY(grp.add(op_mn(dh0, H, B, H), op_mn(xin, XIN, B, XIN), g, 1));
    }
    {  // Whead
      TcGemm g = base_gemm(2 * N, H, (int)B);
      g.o0 = grads + pr.head_w; g.ld = H;
      DDQST_TRY(grp.add(op_mn(dlog, 2 * N, B, 32), op_mn(hL, H, B, H), g, 1));
    }
    {  // S = ind^T . dh0  (for the x_emb gradient / the variant-A input weights)
      TcGemm g = base_gemm(2 * N, H, (int)B);
      g.o0 = S; g.ld = H;
      DDQST_TRY(grp.add(op_mn(ind, 2 * N, B, 32), op_mn(dh0, H, B, H), g, 1));
    }
    DDQST_TRY(grp.launch(s));
  }
  // ------------------------------------------------------------------ bias gradients, x_emb gradient
  {
    ColsumTasks T{};
    int n = 0, maxc = 0;
    auto add = [&](const __nv_bfloat16* p, int64_t ld, int cols, float* out) { T.t[n++] = ColsumTask{p, ld, cols, out}; if (cols > maxc) maxc = cols; };
    for (int l = 0; l < L; ++l) {
      add(dz + 2 * l * BH, H, H, grads + pr.b1[l]);
      add(dz + (2 * l + 1) * BH, H, H, grads + pr.b2[l]);
      add(dgb + (int64_t)l * BG, 2 * H, 2 * H, grads + pr.film_b[l]);
    }
    add(dh0, H, H, grads + pr.in_b);
    add(dlog, 32, 2 * N, grads + pr.head_b);
    const int chunks = (int)((B + 127) / 128 < 16 ? (B + 127) / 128 : 16);
    const int64_t per = (B + chunks - 1) / chunks;
    colsum_bf16_kernel<<<dim3((unsigned)((maxc + 63) / 64), (unsigned)n, (unsigned)chunks), 256, 0, s>>>(T, B, per);
    DDQST_LAUNCH_OK();
  }
  if (var_b && fused) {
    in_b_wgrad_kernel<<<(unsigned)(((int64_t)H * N * E + 255) / 256), 256, 0, s>>>(N, E, H, S, params + pr.x_emb, grads + pr.in_w);
    DDQST_LAUNCH_OK();
  }
  if (var_b) xemb_grad_kernel<<<dim3((unsigned)N, (unsigned)(H / 16)), 128, 0, s>>>(N, E, H, S, params + pr.in_w, grads + pr.x_emb);
  else in_a_wgrad_kernel<<<(unsigned)((H * N + 127) / 128), 128, 0, s>>>(N, H, S, grads + pr.in_w);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // namespace ddqst

using namespace ddqst;

extern "C" {

// register (or clear, with NULL) a device buffer of 4*cap int64 for per-GEMM %globaltimer stamps; resets the launch index
int ddqst_debug_tc_trace(long long* buf, int32_t cap) {
  g_trace_buf = buf; g_trace_cap = cap; g_trace_idx = 0;
  return DDQST_OK;
}

// -1: choose by batch size (default), 0: per-layer GEMM launches, 1: fused forward + data-gradient kernel (where supported)
int ddqst_debug_train_path(int mode) {
  DDQST_REQUIRE(mode >= -1 && mode <= 1, DDQST_EINVAL_SHAPE, "mode=%d", mode);
  g_train_fused_mode = mode;
  return DDQST_OK;
}

// debugging aid (DDQST_FT_DEBUG=1): clock64 stamps of the fused training kernel, see train_fused.cuh
int ddqst_debug_ft_stamps(long long* out256_host) {
  if (cudaDeviceSynchronize() != cudaSuccess) return DDQST_ECUDA;
  return cudaMemcpyFromSymbol(out256_host, g_ft_dbg, sizeof(long long) * 256) == cudaSuccess ? DDQST_OK : DDQST_ECUDA;
}

int ddqst_cast_bf16(const float* src, uint16_t* dst, int64_t n, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(n >= 0 && (n == 0 || (src && dst)), DDQST_EINVAL_SHAPE, "bad argument");
  if (n == 0) return DDQST_OK;
  cast_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_train_forward_backward_tc(const ddqst_dims* d, const float* params, const uint16_t* params_bf16,
                                    const uint16_t* xt_packed, const uint16_t* x0_packed, const int32_t* t,
                                    const int32_t* basis, int64_t batch, float loss_scale, float* grads, float* loss_out,
                                    void* workspace, int64_t ws_bytes, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_TRY(validate_dims(d));
  DDQST_REQUIRE(batch >= 1, DDQST_EINVAL_SHAPE, "batch=%lld", (long long)batch);
  DDQST_REQUIRE(params && params_bf16 && xt_packed && x0_packed && t && basis && grads && loss_out, DDQST_EINVAL_SHAPE, "NULL argument");
  return train_tc_run(d, params, (const __nv_bfloat16*)params_bf16, xt_packed, x0_packed, t, basis, batch, loss_scale, grads,
                      loss_out, workspace, ws_bytes, (cudaStream_t)stream);
}

int ddqst_adam_step_dev(float* params, uint16_t* params_bf16, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        int64_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                        float grad_scale, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(n >= 0 && step_dev, DDQST_EINVAL_SHAPE, "n=%lld", (long long)n);
  if (n > 0) {
    DDQST_REQUIRE(params && grads && exp_avg && exp_avg_sq, DDQST_EINVAL_SHAPE, "NULL argument");
    int64_t blocks = ((n >> 2) + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_tc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, (__nv_bfloat16*)params_bf16, grads, exp_avg, exp_avg_sq,
                                                                      n, step_dev, lr, beta1, beta2, eps, weight_decay, decoupled,
                                                                      grad_scale);
    DDQST_LAUNCH_OK();
  } else {
    step_inc_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_dev);
    DDQST_LAUNCH_OK();
  }
  return DDQST_OK;
}

// C[z][M,N] fp32 = A . B^T over bf16 operands given in either storage order (self test of the training GEMM kernel):
// a_mn == 0: A is [M,K] row-major, else [K,M]; b_mn == 0: B is [N,K] row-major, else [K,N].
int ddqst_selftest_gemm_tc_dbg(const uint16_t* a, const uint16_t* b, int a_mn, int b_mn, int32_t m, int32_t n, int32_t k,
                               int32_t batch, float* c, long long* dbg, void* stream);
int ddqst_selftest_gemm_tc(const uint16_t* a, const uint16_t* b, int a_mn, int b_mn, int32_t m, int32_t n, int32_t k,
                           int32_t batch, float* c, void* stream) {
  return ddqst_selftest_gemm_tc_dbg(a, b, a_mn, b_mn, m, n, k, batch, c, nullptr, stream);
}
int ddqst_selftest_gemm_tc_dbg(const uint16_t* a, const uint16_t* b, int a_mn, int b_mn, int32_t m, int32_t n, int32_t k,
                               int32_t batch, float* c, long long* dbg, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(m >= 1 && n >= 4 && n % 4 == 0 && k >= 1 && batch >= 1 && batch <= kGtMaxZ, DDQST_EINVAL_SHAPE, "selftest shape");
  TcGemm g{};
  g.M = m; g.N = n; g.K = k; g.o0 = c; g.ld = n; g.dbg = dbg;
  for (int z = 0; z < batch; ++z) g.out_zoff[z] = (int64_t)z * m * n;
  HostOperand A = a_mn ? HostOperand{a, 1, m, k, batch, m, (int64_t)k * m, 0, 1} : HostOperand{a, 0, m, k, batch, k, (int64_t)m * k, 0, 1};
  HostOperand B = b_mn ? HostOperand{b, 1, n, k, batch, n, (int64_t)k * n, 0, 1} : HostOperand{b, 0, n, k, batch, k, (int64_t)n * k, 0, 1};
  return launch_gemm<TE_STORE>(A, B, g, batch, (cudaStream_t)stream);
}

}  // extern "C"
