// Hermitian eigensolver for 256 < n <= 1024 (N = 9, 10 qubits): the odd-even one-sided Jacobi of recon.cu / eig_mixed.cuh spread over
// n/8 CO-RESIDENT CTAs (included by recon.cu inside namespace ddqst).
//
// A 16-CTA cluster holds at most 256 column pairs in registers, so for n = 1024 the columns live in n/8 = 128 CTAs of one cooperative
// launch: each CTA owns four adjacent pairs of the line, one pair per 64-thread group, both columns of the pair in registers
// (n/64 elements per thread and column).  The step protocol is the one of jacobi_oddeven_kernel -- after an even step the lower column
// goes to the left neighbour, after an odd step the upper column goes to the right neighbour, one column out and one in per step --
// with two kinds of link:
//   * between groups of one CTA: the column is written into the neighbour group's shared-memory inbox and handed over by an mbarrier
//     (64 arrivals);
//   * between CTAs: the column is written into a global-memory mailbox (it stays in L2) as self-validating 32-byte sectors -- seven
//     payload words plus a tag  (step number) xor (the seven words)  -- one 256-bit store per sector.  The receiver spins on the
//     last sector with 256-bit L2 loads, then fetches the others and re-fetches those whose tag does not match: no flag, no fence;
//     a sector torn between two generations fails the check, so nothing is assumed about the atomicity of the 32-byte access.
// A neighbour cannot run ahead by more than one step because it needs this group's output, so inboxes and mailboxes (one per step
// parity) are always drained before they are refilled.  The only grid-wide synchronisation is one counter barrier per SWEEP, where
// every CTA reads the same rotation count / worst ratio and takes the same stop decision.  Every wait is bounded (g_tc_abort).
//
// A step is a latency chain (dots -> reduction -> angle -> rotation -> hand-over), so the chain is kept short:
//   * registers hold P = the column that STAYED after the last step and Q = the one that just ARRIVED (no renaming copies); the
//     rotation first produces the outgoing column, element by element straight into the inbox / mailbox, and only then the staying
//     one -- the hand-over travels while the second half of the rotation runs;
//   * the staying column's norm is accumulated during that second half and carried into the next step's Gram entries;
//   * the four Gram sums are reduced with a 6-shuffle multi-value butterfly, two accumulators per sum;
//   * fp64: the tangent comes from fp32 arithmetic (an inexact angle only leaves a residual of 1e-7 x the off-diagonal it removes),
//     cos and the phase from fp64 Newton refinements that run side by side, so the rotation is unitary to rounding.
#pragma once

constexpr int kJlThreads = 256;          // four pair groups of 64 threads
constexpr int kJlGroups = 4;
constexpr int kJlHeader = 256;           // bytes in front of the mailboxes: the sweep-barrier counter

template <typename R> struct JlVec;
template <> struct JlVec<double> { typedef double2 V; };
template <> struct JlVec<float> { typedef float2 V; };

__device__ __forceinline__ void jl_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ unsigned int jl_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// The 16-byte chunk j of thread t sits at byte (j * 64 + t) * 16 of a column (global or shared): fp64 keeps row j*64 + t per chunk,
// fp32 the row pair 2 (j*64 + t), +1.  In a mailbox the chunk travels in the 32-byte sector (j * 64 + t): payload, tag, 12 bytes unused.
template <typename V> struct JlChunk { static constexpr int PER = 16 / (int)sizeof(V); };
__device__ __forceinline__ uint4 jl_pack(const double2* x) {
  return make_uint4((uint32_t)__double2loint(x[0].x), (uint32_t)__double2hiint(x[0].x), (uint32_t)__double2loint(x[0].y), (uint32_t)__double2hiint(x[0].y));
}
__device__ __forceinline__ uint4 jl_pack(const float2* x) {
  return make_uint4(__float_as_uint(x[0].x), __float_as_uint(x[0].y), __float_as_uint(x[1].x), __float_as_uint(x[1].y));
}
__device__ __forceinline__ void jl_unpack(const uint4& v, double2* x) {
  x[0] = make_double2(__hiloint2double((int)v.y, (int)v.x), __hiloint2double((int)v.w, (int)v.z));
}
__device__ __forceinline__ void jl_unpack(const uint4& v, float2* x) {
  x[0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y)); x[1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
}
template <typename V, int EPL, int TPP = 64> __device__ __forceinline__ void jl_get(const uint8_t* base, int t, V (&x)[EPL]) {
  constexpr int PER = JlChunk<V>::PER;
#pragma unroll
  for (int j = 0; j < EPL / PER; ++j) jl_unpack(*reinterpret_cast<const uint4*>(base + (j * TPP + t) * 16), &x[j * PER]);
}
template <typename V, int EPL, int TPP = 64> __device__ __forceinline__ void jl_put(uint8_t* base, int t, const V (&x)[EPL]) {
  constexpr int PER = JlChunk<V>::PER;
#pragma unroll
  for (int j = 0; j < EPL / PER; ++j) *reinterpret_cast<uint4*>(base + (j * TPP + t) * 16) = jl_pack(&x[j * PER]);
}
// A column slice in a mailbox: the thread's W = EPL * sizeof(V) / 4 words as a stream, seven per 32-byte sector plus the tag
// (step number) xor (the seven words); sector j of thread t at byte (j * 64 + t) * 32.
template <typename V, int EPL, int TPP = 64> struct JlMail {
  static constexpr int WPE = (int)sizeof(V) / 4;        // words per element
  static constexpr int W = EPL * WPE;
  static constexpr int SEC = (W + 6) / 7;
  static constexpr int BYTES = SEC * TPP * 32;
};
__device__ __forceinline__ uint32_t jl_bits(double v, int half) { return half ? (uint32_t)__double2hiint(v) : (uint32_t)__double2loint(v); }
__device__ __forceinline__ uint32_t jl_bits(float v, int) { return __float_as_uint(v); }
__device__ __forceinline__ void jl_st_sector(uint8_t* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.cg.v8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
               ::"r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "l"(p) : "memory");
}
__device__ __forceinline__ void jl_ld_sector(const uint8_t* p, uint32_t (&w)[8]) {
  asm volatile("ld.relaxed.gpu.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ bool jl_sector_ok(const uint32_t (&w)[8], uint32_t gen) {
  return w[7] == (gen ^ w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6]);
}
__device__ __forceinline__ double2 jl_elem(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, const double2*) {
  return make_double2(__hiloint2double((int)w1, (int)w0), __hiloint2double((int)w3, (int)w2));
}
__device__ __forceinline__ float2 jl_elem(uint32_t w0, uint32_t w1, uint32_t, uint32_t, const float2*) {
  return make_float2(__uint_as_float(w0), __uint_as_float(w1));
}
// mailbox -> column slice: fetch every sector, re-fetch the ones whose tag is not this step's.  (Measured: the hand-over is latency,
// ~500 cycles per L2 access at gpu scope, not bandwidth; spinning on one sector first and fetching the rest afterwards costs a third
// access.  DDQST_JL_SPIN_LAST keeps that variant for comparison.)
template <typename V, int EPL, int TPP = 64> __device__ __forceinline__ void jl_collect(const uint8_t* mail, int t, V (&x)[EPL], uint32_t gen, int code) {
  typedef JlMail<V, EPL, TPP> M;
  uint32_t w[M::SEC][8];
  long long start = 0;
  uint32_t spins = 0;
  bool dead = false;
#ifdef DDQST_JL_SPIN_LAST
  jl_ld_sector(mail + ((M::SEC - 1) * TPP + t) * 32, w[M::SEC - 1]);
  while (!jl_sector_ok(w[M::SEC - 1], gen)) {
    if ((++spins & 0xFFu) == 0) {
      if (start == 0) start = clock64();
      if (*((volatile int*)&g_tc_abort) != 0) { dead = true; break; }
      if (clock64() - start > 2000000000LL) { atomicCAS(&g_tc_abort, 0, code); dead = true; break; }
    }
    jl_ld_sector(mail + ((M::SEC - 1) * TPP + t) * 32, w[M::SEC - 1]);
  }
#pragma unroll
  for (int j = 0; j < M::SEC - 1; ++j) jl_ld_sector(mail + (j * TPP + t) * 32, w[j]);
#else
#pragma unroll
  for (int j = 0; j < M::SEC; ++j) jl_ld_sector(mail + (j * TPP + t) * 32, w[j]);
#endif
  while (!dead) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M::SEC; ++j)
      if (!jl_sector_ok(w[j], gen)) { ok = false; jl_ld_sector(mail + (j * TPP + t) * 32, w[j]); }
    if (ok) break;
    if ((++spins & 0xFFu) == 0) {
      if (start == 0) start = clock64();
      if (*((volatile int*)&g_tc_abort) != 0) break;
      if (clock64() - start > 2000000000LL) { atomicCAS(&g_tc_abort, 0, code); break; }
    }
  }
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int i0 = e * M::WPE, i1 = i0 + 1, i2 = i0 + M::WPE - 2, i3 = i0 + M::WPE - 1;
    x[e] = jl_elem(w[i0 / 7][i0 % 7], w[i1 / 7][i1 % 7], w[i2 / 7][i2 % 7], w[i3 / 7][i3 % 7], (const V*)nullptr);
  }
}

// rotation from the pair's Gram entries (sa, sb, gamma = gr + i gi); false = leave the pair alone (c = 1, s = 0, phase 1 is the plain swap)
__device__ __forceinline__ double jl_rsqrt2(double x) {               // fp32 seed + two Newton steps: 22 -> 44 -> 88 bits
  double y = (double)__frsqrt_rn((float)x);
  const double hx = 0.5 * x;
  y = y * fma(-hx, y * y, 1.5);
  y = y * fma(-hx, y * y, 1.5);
  return y;
}
__device__ __forceinline__ bool jl_angle(double sa, double sb, double gr, double gi, double tol, double& c, double& s, double& pr, double& pi) {
  const double g2 = gr * gr + gi * gi;
  if (!(g2 > tol * tol * sa * sb && g2 > 1e-60)) return false;
  if (g2 > 1e-30 && g2 < 1e30) {
    // tangent in fp32 (the difference of the norms is taken in fp64 first); independent of the fp64 refinement of 1/|gamma| below
    const float inv_gf = rsqrtf((float)g2);
    const float zf = (float)(0.5 * (sb - sa)) * inv_gf, az = fabsf(zf);
    const float at = az < 1e4f ? __frcp_rn(az + sqrtf(fmaf(az, az, 1.f))) : 0.5f * __frcp_rn(az);
    const double t = (double)(zf >= 0.f ? at : -at);
    const double inv_g = jl_rsqrt2(g2);
    c = jl_rsqrt2(fma(t, t, 1.0)); s = c * t;
    pr = gr * inv_g; pi = -gi * inv_g;
  } else {
    const double gabs = sqrt(g2);
    const double zeta = (sb - sa) / (2.0 * gabs);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    c = 1.0 / sqrt(1.0 + t * t); s = c * t;
    pr = gr / gabs; pi = -gi / gabs;
  }
  return true;
}
__device__ __forceinline__ bool jl_angle(float sa, float sb, float gr, float gi, float tol, float& c, float& s, float& pr, float& pi) {
  const float g2 = gr * gr + gi * gi;
  if (!(g2 > tol * tol * sa * sb && g2 > 1e-30f)) return false;
  const float inv_g = rsqrtf(g2);
  const float zeta = 0.5f * (sb - sa) * inv_g, az = fabsf(zeta);
  const float at = az < 1e4f ? __frcp_rn(az + sqrtf(fmaf(az, az, 1.f))) : 0.5f * __frcp_rn(az);
  const float t = zeta >= 0.f ? at : -at;
  c = rsqrtf(fmaf(t, t, 1.f)); s = c * t;
  pr = gr * inv_g; pi = -gi * inv_g;
  return true;
}
// four per-lane partial sums -> the four warp totals with 6 shuffles: lanes 0-7 end with v0's total, 8-15 v1's, 16-23 v2's, 24-31 v3's
template <typename R> __device__ __forceinline__ R jl_reduce4(R v0, R v1, R v2, R v3, int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  R k0 = hi16 ? v2 : v0, k1 = hi16 ? v3 : v1;           // kept pair; the other pair goes to the partner half
  const R s0 = hi16 ? v0 : v2, s1 = hi16 ? v1 : v3;
  k0 += __shfl_xor_sync(0xFFFFFFFFu, s0, 16);
  k1 += __shfl_xor_sync(0xFFFFFFFFu, s1, 16);
  R k = hi8 ? k1 : k0;
  const R sx = hi8 ? k0 : k1;
  k += __shfl_xor_sync(0xFFFFFFFFu, sx, 8);
  k += __shfl_xor_sync(0xFFFFFFFFu, k, 4);
  k += __shfl_xor_sync(0xFFFFFFFFu, k, 2);
  k += __shfl_xor_sync(0xFFFFFFFFu, k, 1);
  return k;
}

// Rotation, first half: the outgoing column, straight to its destination (DST 0: shared-memory inbox / park slot, DST 1: mailbox).
//   ODD  step (lower = P, upper = Q):  Q <- Q p;  out = c P - s Q   (rotated lower -> upper position -> right neighbour)
//   EVEN step (lower = Q, upper = P):  P <- P p;  out = s Q + c P   (rotated upper -> lower position -> left neighbour)
template <bool ODD, int DST, bool SWAP, typename V, typename R, int EPL, int TPP = 64>
__device__ __forceinline__ void jl_rotate_out(V (&P)[EPL], V (&Q)[EPL], R c, R s, R pr, R pi, uint8_t* dst, int t, uint32_t gen) {
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    if (ODD) { const V q = Q[e]; Q[e].x = q.x * pr - q.y * pi; Q[e].y = q.x * pi + q.y * pr; }
    else { const V p = P[e]; P[e].x = p.x * pr - p.y * pi; P[e].y = p.x * pi + p.y * pr; }
  }
  auto out = [&](int e, int im) -> R {
    const R p = im ? P[e].y : P[e].x, q = im ? Q[e].y : Q[e].x;
    // with (p, q) the pair after the phase multiply: rotated lower / upper column =  ODD: c p - s q / s p + c q,  EVEN: c q - s p / c p + s q.
    // SWAP (the columns exchange positions after the rotation, the hand-over rule of the line ordering): the rotated lower (ODD) or
    // upper (EVEN) column leaves; without SWAP the rotated version of Q itself moves on.
    return ODD ? (SWAP ? c * p - s * q : s * p + c * q) : (SWAP ? c * p + s * q : c * q - s * p);
  };
  if (DST == 0) {
    constexpr int PER = JlChunk<V>::PER;
#pragma unroll
    for (int j = 0; j < EPL / PER; ++j) {
      V o[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) { o[i].x = out(j * PER + i, 0); o[i].y = out(j * PER + i, 1); }
      *reinterpret_cast<uint4*>(dst + (j * TPP + t) * 16) = jl_pack(o);
    }
  } else {
    typedef JlMail<V, EPL, TPP> M;
#pragma unroll
    for (int j = 0; j < M::SEC; ++j) {
      uint32_t w[8];
      uint32_t tag = gen;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int idx = j * 7 + i;
        if (idx < M::W) {
          const int e = idx / M::WPE, cpt = idx % M::WPE;
          w[i] = jl_bits(out(e, cpt >= M::WPE / 2 ? 1 : 0), cpt & 1);
        } else {
          w[i] = 0u;
        }
        tag ^= w[i];
      }
      w[7] = tag;
      jl_st_sector(dst + (j * TPP + t) * 32, w);
    }
  }
}
// second half: the staying column, in place in P, and its norm for the next step
template <bool ODD, bool SWAP, typename V, typename R, int EPL>
__device__ __forceinline__ R jl_rotate_stay(V (&P)[EPL], const V (&Q)[EPL], R c, R s) {
  R n0 = 0, n1 = 0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    V p;
    // the complement of jl_rotate_out; (P, Q) are the phase-multiplied pair it left behind
    const V a = P[e], b = Q[e];
    if (ODD) {
      if (SWAP) { p.x = s * a.x + c * b.x; p.y = s * a.y + c * b.y; } else { p.x = c * a.x - s * b.x; p.y = c * a.y - s * b.y; }
    } else {
      if (SWAP) { p.x = c * b.x - s * a.x; p.y = c * b.y - s * a.y; } else { p.x = c * a.x + s * b.x; p.y = c * a.y + s * b.y; }
    }
    P[e] = p;
    if (e & 1) n1 += p.x * p.x + p.y * p.y; else n0 += p.x * p.x + p.y * p.y;
  }
  return n0 + n1;
}

#ifdef DDQST_JL_PROFILE
__device__ long long g_jl_prof[32];      // [watch 0: CTA 5 group 1 (interior) | watch 1: CTA 5 group 3 (mailbox receiver)][phase]
#define JL_STAMP(i) do { long long _t = clock64(); _acc[i] += _t - _t0; _t0 = _t; } while (0)
#else
#define JL_STAMP(i) do { } while (0)
#endif

template <typename R, int EPL>
__global__ void __launch_bounds__(kJlThreads, 1) jacobi_line_kernel(typename JlVec<R>::V* __restrict__ GT, int n, int max_sweeps, R tol,
                                                                    JacobiCtl* ctl, uint8_t* __restrict__ comm) {
  typedef typename JlVec<R>::V V;
  extern __shared__ __align__(16) uint8_t jl_smem[];
  constexpr int COLB = EPL * 64 * (int)sizeof(V);
  constexpr int MAILB = JlMail<V, EPL>::BYTES;
  const int tid = threadIdx.x, grp = tid >> 6, t = tid & 63, lane = tid & 31, wig = (tid >> 5) & 1;
  const int c = blockIdx.x, nc = gridDim.x;
  const int h = n / 2, k = c * kJlGroups + grp;        // n = 8 * gridDim.x: every group owns a pair
  // shared memory: inbox[group][parity] (0: from the left, filled after odd steps; 1: from the right, after even steps), pair 0's park
  // slot, reduction scratch [group][step parity][warp][4], mbarriers [group][parity]
  uint8_t* my_in0 = jl_smem + (grp * 2 + 0) * COLB;
  uint8_t* my_in1 = jl_smem + (grp * 2 + 1) * COLB;
  uint8_t* park = jl_smem + 8 * COLB;
  R* red = reinterpret_cast<R*>(jl_smem + 9 * COLB) + grp * 16;
  const uint32_t bars0 = smem_u32(jl_smem + 9 * COLB + kJlGroups * 16 * (int)sizeof(R));
  auto bar = [&](int g_, int par) { return bars0 + (uint32_t)((g_ * 2 + par) * 8); };
  if (t == 0) { mbar_init(bar(grp, 0), 64); mbar_init(bar(grp, 1), 64); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  unsigned int* gbar = reinterpret_cast<unsigned int*>(comm);
  auto mail = [&](int par, int cta) { return comm + kJlHeader + (size_t)(par * nc + cta) * MAILB; };
  // where the outgoing column goes: after even steps to the left (pair 0 parks it), after odd steps to the right
  const bool first = k == 0, last = k == h - 1;
  uint8_t* const out_even = first ? park : grp >= 1 ? jl_smem + ((grp - 1) * 2 + 1) * COLB : mail(1, c - 1);
  uint8_t* const out_odd = grp <= kJlGroups - 2 ? jl_smem + ((grp + 1) * 2 + 0) * COLB : mail(0, last ? c : c + 1);
  const bool even_to_mail = !first && grp == 0, odd_to_mail = grp == kJlGroups - 1;

  V P[EPL], Q[EPL];                                     // P: the column that stayed after the last step, Q: the one that arrived
  jl_get<V, EPL>(reinterpret_cast<const uint8_t*>(GT + (int64_t)(2 * k) * n), t, Q);          // step 0 is even: lower = Q, upper = P
  jl_get<V, EPL>(reinterpret_cast<const uint8_t*>(GT + (int64_t)(2 * k + 1) * n), t, P);
  R carry = 0;                                          // this thread's share of |P|^2
#pragma unroll
  for (int e = 0; e < EPL; ++e) carry += P[e].x * P[e].x + P[e].y * P[e].y;
  int sweep = 0;
  uint32_t g = 0;
#ifdef DDQST_JL_PROFILE
  const int jl_watch = c == 5 ? (grp == 1 ? 0 : grp == 3 ? 1 : -1) : -1;
  long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long _t0 = clock64();
#endif
  for (; sweep < max_sweeps; ++sweep) {
    int rot = 0;
    float worst = 0.f;
    for (int step = 0; step < n; ++step, ++g) {
      const bool odd = (g & 1u) != 0u;
      JL_STAMP(7);
      if (!odd || !last) {                              // in odd steps the last group only holds the idle position n-1
        R q0 = 0, q1 = 0, r0 = 0, r1 = 0, i0 = 0, i1 = 0;                // |Q|^2 and d = conj(P) . Q, two accumulators each
#pragma unroll
        for (int e = 0; e < EPL; e += 2) {
          q0 += Q[e].x * Q[e].x + Q[e].y * Q[e].y;
          r0 += P[e].x * Q[e].x + P[e].y * Q[e].y;
          i0 += P[e].x * Q[e].y - P[e].y * Q[e].x;
          q1 += Q[e + 1].x * Q[e + 1].x + Q[e + 1].y * Q[e + 1].y;
          r1 += P[e + 1].x * Q[e + 1].x + P[e + 1].y * Q[e + 1].y;
          i1 += P[e + 1].x * Q[e + 1].y - P[e + 1].y * Q[e + 1].x;
        }
        JL_STAMP(0);
        const R tot = jl_reduce4<R>(carry, q0 + q1, r0 + r1, i0 + i1, lane);
        R* rp = red + (g & 1u) * 8;
        if ((lane & 7) == 0) rp[wig * 4 + (lane >> 3)] = tot;
        jl_bar_sync(1 + grp, 64);
        const R np = rp[0] + rp[4], nq = rp[1] + rp[5], dr = rp[2] + rp[6], di = rp[3] + rp[7];   // same order in both warps: same angle
        // odd: lower = P, upper = Q, gamma = conj(P).Q = d;  even: lower = Q, upper = P, gamma = conj(d)
        const R sa = odd ? np : nq, sb = odd ? nq : np, gr = dr, gi = odd ? di : -di;
        worst = fmaxf(worst, (float)((gr * gr + gi * gi) / (sa * sb)));
        JL_STAMP(1);
        R cc = 1, ss = 0, pr = 1, pi = 0;
        if (jl_angle(sa, sb, gr, gi, tol, cc, ss, pr, pi)) ++rot;
        JL_STAMP(2);
        if (odd) {
          if (odd_to_mail) {
            jl_rotate_out<true, 1, true, V, R, EPL>(P, Q, cc, ss, pr, pi, out_odd, t, g + 1u);
          } else {
            jl_rotate_out<true, 0, true, V, R, EPL>(P, Q, cc, ss, pr, pi, out_odd, t, 0u);
            mbar_arrive(bar(grp + 1, 0));
          }
          JL_STAMP(3);
          carry = jl_rotate_stay<true, true, V, R, EPL>(P, Q, cc, ss);
        } else {
          if (even_to_mail) {
            jl_rotate_out<false, 1, true, V, R, EPL>(P, Q, cc, ss, pr, pi, out_even, t, g + 1u);
          } else {
            jl_rotate_out<false, 0, true, V, R, EPL>(P, Q, cc, ss, pr, pi, out_even, t, 0u);
            if (!first) mbar_arrive(bar(grp - 1, 1));
          }
          JL_STAMP(3);
          carry = jl_rotate_stay<false, true, V, R, EPL>(P, Q, cc, ss);
        }
        JL_STAMP(4);
      }
      // the arriving column
      if (!odd) {                                       // after an even step: from the right (nothing for the last group)
        if (!last) {
          if (grp <= kJlGroups - 2) {
            mbar_wait(bar(grp, 1), (g >> 1) & 1u, 66);
            jl_get<V, EPL>(my_in1, t, Q);
          } else {
            jl_collect<V, EPL>(mail(1, c), t, Q, g + 1u, 67);
          }
        }
      } else {                                          // after an odd step: from the left (pair 0: the column it parked)
        if (first) {
          jl_get<V, EPL>(park, t, Q);                   // every thread reads back exactly the chunks it wrote
        } else if (grp >= 1) {
          mbar_wait(bar(grp, 0), (g >> 1) & 1u, 68);
          jl_get<V, EPL>(my_in0, t, Q);
        } else {
          jl_collect<V, EPL>(mail(0, c), t, Q, g + 1u, 69);
        }
      }
      JL_STAMP(6);
    }
    if (t == 0) {                                       // the group's threads took identical decisions
      if (rot > 0) atomicAdd(&ctl->rotations[sweep], rot);
      if (sweep < 48) atomicMax(&ctl->max_ratio2[sweep], __float_as_uint(worst));
    }
    // sweep barrier over the grid: monotonic arrival counter
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(gbar, 1u);
      const unsigned int want = (unsigned int)(sweep + 1) * (unsigned int)nc;
      long long start = 0;
      uint32_t spins = 0;
      while (jl_ld_acquire(gbar) < want) {
        if ((++spins & 0xFFu) == 0) {
          if (start == 0) start = clock64();
          if (*((volatile int*)&g_tc_abort) != 0) break;
          if (clock64() - start > 2000000000LL) { atomicCAS(&g_tc_abort, 0, 70); break; }
        }
      }
      __threadfence();
    }
    __syncthreads();
    const int total = __ldcg(&ctl->rotations[sweep]);
    if (total == 0) { ++sweep; break; }
    if (sweep < 48 && __uint_as_float(__ldcg(&ctl->max_ratio2[sweep])) < __ldcg(&ctl->stop_ratio2)) { ++sweep; break; }
    if (*((volatile int*)&g_tc_abort) != 0) { ++sweep; break; }
  }
#ifdef DDQST_JL_PROFILE
  if (jl_watch >= 0 && t == 0)
    for (int i = 0; i < 8; ++i) g_jl_prof[jl_watch * 16 + i] += _acc[i];
#endif
  // after whole sweeps the next step would be an even one: lower position 2k = Q, upper position 2k+1 = P
  jl_put<V, EPL>(reinterpret_cast<uint8_t*>(GT + (int64_t)(2 * k) * n), t, Q);
  jl_put<V, EPL>(reinterpret_cast<uint8_t*>(GT + (int64_t)(2 * k + 1) * n), t, P);
  if (k == 0 && t == 0) ctl->sweeps_done = sweep;
}

template <typename R, int EPL>
static int launch_jacobi_line(typename JlVec<R>::V* GT, int n, int max_sweeps, R tol, JacobiCtl* ctl, uint8_t* comm, int64_t comm_bytes,
                              cudaStream_t s, bool* launched) {
  typedef typename JlVec<R>::V V;
  constexpr int COLB = EPL * 64 * (int)sizeof(V);
  constexpr int smem = 9 * COLB + kJlGroups * 16 * (int)sizeof(R) + kJlGroups * 2 * 8 + 16;
  *launched = false;
  if (n != EPL * 64 || n % 8 != 0) return DDQST_OK;
  const int nc = n / 8;
  const int64_t need = kJlHeader + (int64_t)2 * nc * JlMail<V, EPL>::BYTES;
  if (comm == nullptr || comm_bytes < need) return DDQST_OK;
  static bool attr_set = false;
  if (!attr_set) {
    DDQST_CUDA_OK(cudaFuncSetAttribute(jacobi_line_kernel<R, EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_line_kernel<R, EPL>, kJlThreads, smem) != cudaSuccess ||
      (int64_t)per_sm * num_sms() < nc) {
    (void)cudaGetLastError();
    return DDQST_OK;                                    // not all CTAs can be co-resident: the caller keeps the cooperative kernel
  }
  DDQST_CUDA_OK(cudaMemsetAsync(comm, 0, (size_t)need, s));            // barrier counter 0; all-zero sectors carry tag 0, steps count from 1
  void* args[] = {&GT, &n, &max_sweeps, &tol, &ctl, &comm};
  DDQST_CUDA_OK(cudaLaunchCooperativeKernel((void*)jacobi_line_kernel<R, EPL>, dim3((unsigned)nc), dim3(kJlThreads), args, smem, s));
  *launched = true;
  return DDQST_OK;
}

// ================================================================================================ two-level (block) ordering
// The line kernel pays one L2 hand-over (~1 900 cycles) on the critical path of EVERY step.  Here the line is made of BLOCKS of four
// columns: a CTA owns two adjacent blocks -- group i keeps one column of each, P_i and Q -- and a block step rotates all 16 cross
// pairs in four local steps (the Q columns go round the four groups through shared memory, one __syncthreads per local step);
// only the fourth local step applies the line rule -- swap, lower block to the left neighbour CTA after even block steps, upper block
// to the right after odd ones -- through the mailboxes, so three steps in four never leave the SM.  n/4 block steps bring every pair
// of blocks together once; pairs inside a block are rotated once per sweep, at its start, out of shared memory.
template <typename V, typename R, int EPL, int TPP = 64>
__device__ __forceinline__ bool jl_pair_angle(const V (&P)[EPL], const V (&Q)[EPL], R carry, bool odd, R* rp, int lane, int wig, int grp,
                                              R tol, R& cc, R& ss, R& pr, R& pi, float& worst) {
  R q0 = 0, q1 = 0, r0 = 0, r1 = 0, i0 = 0, i1 = 0;                      // |Q|^2 and d = conj(P) . Q, two accumulators each
#pragma unroll
  for (int e = 0; e < EPL; e += 2) {
    q0 += Q[e].x * Q[e].x + Q[e].y * Q[e].y;
    r0 += P[e].x * Q[e].x + P[e].y * Q[e].y;
    i0 += P[e].x * Q[e].y - P[e].y * Q[e].x;
    q1 += Q[e + 1].x * Q[e + 1].x + Q[e + 1].y * Q[e + 1].y;
    r1 += P[e + 1].x * Q[e + 1].x + P[e + 1].y * Q[e + 1].y;
    i1 += P[e + 1].x * Q[e + 1].y - P[e + 1].y * Q[e + 1].x;
  }
  const R tot = jl_reduce4<R>(carry, q0 + q1, r0 + r1, i0 + i1, lane);
  R np, nq, dr, di;
  if (TPP == 32) {                                      // one warp per pair: the four totals sit in lanes 0 / 8 / 16 / 24
    np = __shfl_sync(0xFFFFFFFFu, tot, 0); nq = __shfl_sync(0xFFFFFFFFu, tot, 8);
    dr = __shfl_sync(0xFFFFFFFFu, tot, 16); di = __shfl_sync(0xFFFFFFFFu, tot, 24);
  } else {
    if ((lane & 7) == 0) rp[wig * 4 + (lane >> 3)] = tot;
    jl_bar_sync(1 + grp, 64);
    np = rp[0] + rp[4]; nq = rp[1] + rp[5]; dr = rp[2] + rp[6]; di = rp[3] + rp[7];
  }
  const R sa = odd ? np : nq, sb = odd ? nq : np, gr = dr, gi = odd ? di : -di;
  worst = fmaxf(worst, (float)((gr * gr + gi * gi) / (sa * sb)));
  cc = 1; ss = 0; pr = 1; pi = 0;
  return jl_angle(sa, sb, gr, gi, tol, cc, ss, pr, pi);
}

template <typename R, int EPL, int G, int TPP>
__global__ void __launch_bounds__(TPP * G, 1) jacobi_block_kernel(typename JlVec<R>::V* __restrict__ GT, int n, int max_sweeps, R tol,
                                                                     JacobiCtl* ctl, uint8_t* __restrict__ comm0, uint8_t* __restrict__ comm1) {
  typedef typename JlVec<R>::V V;
  extern __shared__ __align__(16) uint8_t jl_smem[];
  constexpr int COLB = EPL * TPP * (int)sizeof(V);
  constexpr int MAILB = JlMail<V, EPL, TPP>::BYTES;
  const int tid = threadIdx.x, grp = tid / TPP, t = tid % TPP, lane = tid & 31, wig = t >> 5;
  const int c = blockIdx.x, nc = gridDim.x;
  const bool first = c == 0, last = c == nc - 1;
  // G columns per block (G groups of 64 threads per CTA; G = 4, or 8 for the latency-bound small n where a global hand-over every
  // eighth step pays).  Shared memory: 2G column slots (local inboxes [group][parity]; the within-block phase keeps the CTA's 2G columns
  // there), G park slots (CTA 0 parks its lower block during odd block steps), reduction scratch [group][parity][warp][4]
  auto slot = [&](int i) { return jl_smem + i * COLB; };
  R* red = reinterpret_cast<R*>(jl_smem + 3 * G * COLB) + grp * 16;
  unsigned int* gbar = &ctl->sweep_barrier;
  // mailbox [direction][CTA][group]: direction 0 = filled by the left neighbour (after odd block steps), 1 = by the right one
  auto mail = [&](int dir, int cta) { return (dir ? comm1 : comm0) + (size_t)(cta * G + grp) * MAILB; };

  V P[EPL], Q[EPL];                                     // block step 0 is even: lower block = the Q columns, upper block = the P columns
  jl_get<V, EPL, TPP>(reinterpret_cast<const uint8_t*>(GT + (int64_t)(2 * G * c + grp) * n), t, Q);
  jl_get<V, EPL, TPP>(reinterpret_cast<const uint8_t*>(GT + (int64_t)(2 * G * c + G + grp) * n), t, P);
  R carry = 0;
  uint32_t rc = 0;                                      // reduction-scratch parity
  int sweep = 0;
  uint32_t gb = 0;                                      // block steps so far (mailbox generation)
  const int nblk = n / G;
#ifdef DDQST_JL_PROFILE
  const int jl_watch = (c == 5 && grp == 1) ? 0 : (c == 5 && grp == 3) ? 1 : -1;
  long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long _t0 = clock64();
#endif
  for (; sweep < max_sweeps; ++sweep) {
    int rot = 0;
    float worst = 0.f;
    JL_STAMP(7);
    // ---- pairs inside the two resident blocks: G-1 rounds of a round-robin tournament (circle method) out of shared memory,
    //      block grp / (G/2), pair grp % (G/2)
    jl_put<V, EPL, TPP>(slot(grp), t, Q);
    jl_put<V, EPL, TPP>(slot(G + grp), t, P);
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < G - 1; ++r) {
      const int blk = grp / (G / 2), k = grp % (G / 2);
      int i = k == 0 ? G - 1 : (r + k) % (G - 1), j = k == 0 ? r : (r + (G - 1) - k) % (G - 1);
      if (i > j) { const int tmp = i; i = j; j = tmp; }
      jl_get<V, EPL, TPP>(slot(blk * G + i), t, P);
      jl_get<V, EPL, TPP>(slot(blk * G + j), t, Q);
      R np = 0;
#pragma unroll
      for (int e = 0; e < EPL; ++e) np += P[e].x * P[e].x + P[e].y * P[e].y;
      R cc, ss, pr, pi;
      if (jl_pair_angle<V, R, EPL, TPP>(P, Q, np, true, red + (rc++ & 1u) * 8, lane, wig, grp, tol, cc, ss, pr, pi, worst)) ++rot;
      jl_rotate_out<true, 0, false, V, R, EPL, TPP>(P, Q, cc, ss, pr, pi, slot(blk * G + j), t, 0u);
      (void)jl_rotate_stay<true, false, V, R, EPL>(P, Q, cc, ss);
      jl_put<V, EPL, TPP>(slot(blk * G + i), t, P);
      __syncthreads();
    }
    jl_get<V, EPL, TPP>(slot(grp), t, Q);
    jl_get<V, EPL, TPP>(slot(G + grp), t, P);
    carry = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) carry += P[e].x * P[e].x + P[e].y * P[e].y;
    __syncthreads();                                    // the slots turn back into inboxes
    JL_STAMP(5);
    // ---- block steps
    uint32_t lp = 0;                                    // local inbox parity
#pragma unroll 1
    for (int bs = 0; bs < nblk; ++bs, ++gb) {
      const bool odd = (gb & 1u) != 0u;
      if (!(odd && last)) {                             // in odd block steps the last CTA only holds the idle top block
#pragma unroll 1
        for (int ls = 0; ls < G; ++ls) {
          R cc, ss, pr, pi;
          if (jl_pair_angle<V, R, EPL, TPP>(P, Q, carry, odd, red + (rc++ & 1u) * 8, lane, wig, grp, tol, cc, ss, pr, pi, worst)) ++rot;
          JL_STAMP(0);
          if (ls < G - 1) {
            uint8_t* dst = slot(((grp + G - 1) % G) * 2 + (int)lp);  // the rotated Q moves on to group i-1
            if (odd) { jl_rotate_out<true, 0, false, V, R, EPL, TPP>(P, Q, cc, ss, pr, pi, dst, t, 0u); carry = jl_rotate_stay<true, false, V, R, EPL>(P, Q, cc, ss); }
            else { jl_rotate_out<false, 0, false, V, R, EPL, TPP>(P, Q, cc, ss, pr, pi, dst, t, 0u); carry = jl_rotate_stay<false, false, V, R, EPL>(P, Q, cc, ss); }
            JL_STAMP(1);
            __syncthreads();
            jl_get<V, EPL, TPP>(slot(grp * 2 + (int)lp), t, Q);
            lp ^= 1u;
            JL_STAMP(2);
          } else if (odd) {                             // upper block -> right neighbour
            jl_rotate_out<true, 1, true, V, R, EPL, TPP>(P, Q, cc, ss, pr, pi, mail(0, c + 1), t, gb + 1u);
            carry = jl_rotate_stay<true, true, V, R, EPL>(P, Q, cc, ss);
          } else {                                      // lower block -> left neighbour (CTA 0 parks it)
            if (first) jl_rotate_out<false, 0, true, V, R, EPL, TPP>(P, Q, cc, ss, pr, pi, slot(2 * G + grp), t, 0u);
            else jl_rotate_out<false, 1, true, V, R, EPL, TPP>(P, Q, cc, ss, pr, pi, mail(1, c - 1), t, gb + 1u);
            carry = jl_rotate_stay<false, true, V, R, EPL>(P, Q, cc, ss);
          }
        }
      }
      JL_STAMP(3);
      // the arriving block
      if (!odd) {
        if (!last) jl_collect<V, EPL, TPP>(mail(1, c), t, Q, gb + 1u, 71);
      } else {
        if (first) jl_get<V, EPL, TPP>(slot(2 * G + grp), t, Q);          // every thread reads back exactly the chunks it wrote
        else jl_collect<V, EPL, TPP>(mail(0, c), t, Q, gb + 1u, 72);
      }
      JL_STAMP(4);
    }
    if (t == 0) {
      if (rot > 0) atomicAdd(&ctl->rotations[sweep], rot);
      if (sweep < 48) atomicMax(&ctl->max_ratio2[sweep], __float_as_uint(worst));
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(gbar, 1u);
      const unsigned int want = (unsigned int)(sweep + 1) * (unsigned int)nc;
      long long start = 0;
      uint32_t spins = 0;
      while (jl_ld_acquire(gbar) < want) {
        if ((++spins & 0xFFu) == 0) {
          if (start == 0) start = clock64();
          if (*((volatile int*)&g_tc_abort) != 0) break;
          if (clock64() - start > 2000000000LL) { atomicCAS(&g_tc_abort, 0, 73); break; }
        }
      }
      __threadfence();
    }
    __syncthreads();
    JL_STAMP(6);
    const int total = __ldcg(&ctl->rotations[sweep]);
    if (total == 0) { ++sweep; break; }
    if (sweep < 48 && __uint_as_float(__ldcg(&ctl->max_ratio2[sweep])) < __ldcg(&ctl->stop_ratio2)) { ++sweep; break; }
    if (*((volatile int*)&g_tc_abort) != 0) { ++sweep; break; }
  }
#ifdef DDQST_JL_PROFILE
  if (jl_watch >= 0 && t == 0)
    for (int i = 0; i < 8; ++i) g_jl_prof[jl_watch * 16 + i] += _acc[i];
#endif
  jl_put<V, EPL, TPP>(reinterpret_cast<uint8_t*>(GT + (int64_t)(2 * G * c + grp) * n), t, Q);
  jl_put<V, EPL, TPP>(reinterpret_cast<uint8_t*>(GT + (int64_t)(2 * G * c + G + grp) * n), t, P);
  if (c == 0 && tid == 0) ctl->sweeps_done = sweep;
}

template <typename R, int EPL, int G = 4, int TPP = 64>
static int launch_jacobi_block(typename JlVec<R>::V* GT, int n, int max_sweeps, R tol, JacobiCtl* ctl, uint8_t* comm0, int64_t bytes0,
                               uint8_t* comm1, int64_t bytes1, cudaStream_t s, bool* launched) {
  typedef typename JlVec<R>::V V;
  constexpr int COLB = EPL * TPP * (int)sizeof(V);
  constexpr int smem = 3 * G * COLB + G * 16 * (int)sizeof(R) + 16;
  *launched = false;
  if (n != EPL * TPP || n % (2 * G) != 0 || n / (2 * G) < 2) return DDQST_OK;
  const int nc = n / (2 * G);
  const int64_t need = (int64_t)nc * G * JlMail<V, EPL, TPP>::BYTES;
  if (comm0 == nullptr || comm1 == nullptr || bytes0 < need || bytes1 < need) return DDQST_OK;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(jacobi_block_kernel<R, EPL, G, TPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return DDQST_OK;
    }
    attr_set = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_block_kernel<R, EPL, G, TPP>, TPP * G, smem) != cudaSuccess ||
      (int64_t)per_sm * num_sms() < nc) {
    (void)cudaGetLastError();
    return DDQST_OK;
  }
  DDQST_CUDA_OK(cudaMemsetAsync(comm0, 0, (size_t)need, s));           // all-zero sectors carry tag 0, block steps count from 1
  DDQST_CUDA_OK(cudaMemsetAsync(comm1, 0, (size_t)need, s));
  DDQST_CUDA_OK(cudaMemsetAsync(&ctl->sweep_barrier, 0, sizeof(unsigned int), s));
  void* args[] = {&GT, &n, &max_sweeps, &tol, &ctl, &comm0, &comm1};
  DDQST_CUDA_OK(cudaLaunchCooperativeKernel((void*)jacobi_block_kernel<R, EPL, G, TPP>, dim3((unsigned)nc), dim3(TPP * G), args, smem, s));
  *launched = true;
  return DDQST_OK;
}

// ---- register-tiled fp64 complex GEMM for the glue at n >= 512 (64 x 64 tile per CTA, 4 x 4 outputs per thread); OP as eig_zgemm_kernel
template <int OP>
__global__ void __launch_bounds__(256) eig_zgemm64_kernel(const double2* __restrict__ A, const double2* __restrict__ B, const double2* __restrict__ A2,
                                                          int n, const JacobiCtl* ctl, double2* __restrict__ C) {
  __shared__ double2 As[16][65], Bs[16][65];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  double ax[4][4], ay[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { ax[i][j] = 0.0; ay[i][j] = 0.0; }
  for (int k0 = 0; k0 < n; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // tiles read along k (row-major rows), stored k-major
      const int idx = tid + 256 * i, row = idx >> 4, kk = idx & 15;
      As[kk][row] = A[(int64_t)(r0 + row) * n + k0 + kk];
      if (OP == 0 || OP == 4) { const double2 v = B[(int64_t)(c0 + row) * n + k0 + kk]; Bs[kk][row] = make_double2(v.x, OP == 0 ? -v.y : v.y); }   // B^H / B^T
    }
    if (OP != 0 && OP != 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + 256 * i, kk = idx >> 6, cc = idx & 63;
        const double2 v = B[(int64_t)(k0 + kk) * n + c0 + cc];
        Bs[kk][cc] = OP == 1 ? v : make_double2(0.5 * v.x, -0.5 * v.y);
      }
    }
    if (OP == 2) {                                      // + 0.5 H[c][k]: conj of the symmetrised Hermitian input (jacobi_init_kernel)
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + 256 * i, row = idx >> 4, kk = idx & 15;
        const double2 w = B[(int64_t)(c0 + row) * n + k0 + kk];
        Bs[kk][row].x += 0.5 * w.x; Bs[kk][row].y += 0.5 * w.y;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double2 av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty + 16 * i]; bv[i] = Bs[kk][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ax[i][j] = fma(av[i].x, bv[j].x, fma(-av[i].y, bv[j].y, ax[i][j]));
          ay[i][j] = fma(av[i].x, bv[j].y, fma(av[i].y, bv[j].x, ay[i][j]));
        }
    }
    __syncthreads();
  }
  const double sg = OP == 2 ? ctl->sigma : 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t e = (int64_t)(r0 + ty + 16 * i) * n + c0 + tx + 16 * j;
      if (OP == 0 || OP == 4) C[e] = make_double2(ax[i][j], ay[i][j]);
      else if (OP == 1) { const double2 v = A2[e]; C[e] = make_double2(1.5 * v.x - 0.5 * ax[i][j], 1.5 * v.y - 0.5 * ay[i][j]); }
      else { const double2 v = A[e]; C[e] = make_double2(ax[i][j] + sg * v.x, ay[i][j] + sg * v.y); }
    }
}

template <int OP>
static int launch_eig_zgemm(const double2* A, const double2* B, const double2* A2, int n, const JacobiCtl* ctl, double2* C, cudaStream_t s) {
  if (n >= 512 && n % 64 == 0) {
    eig_zgemm64_kernel<OP><<<dim3(n / 64, n / 64), 256, 0, s>>>(A, B, A2, n, ctl, C);
  } else {
    dim3 grid((n + 15) / 16, (n + 15) / 16), blk(16, 16);
    eig_zgemm_kernel<OP><<<grid, blk, 0, s>>>(A, B, A2, n, ctl, C);
  }
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}
