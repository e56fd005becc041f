// Synthetic measurement data on the device: the stand-in for the reference's Qiskit-Aer data
// generation (SS/data_gen.py:40-63, AS/data_gen.py:59-140, RQC/batch_build_dataset.py:53-144) that bench inputs and
// tests need at N = 8..10, where 3^N bases x 10^6 shots is out of reach for a CPU simulator.
//
//   synth_state_kernel : |0..0> -> special state ('plus', 'ghz') or a brick-wall random circuit (per layer a random
//                        U3 on every qubit, then CZ on alternating neighbour pairs); one CTA, state vector in shared memory.
//   synth_born_kernel  : one CTA per (basis, shot chunk): rotate the state into the measurement basis (H for X, H.Sdg for Y,
//                        SS/data_gen.py:28-33; string position i = qubit i = bit i of the outcome index, Qiskit little
//                        endian), Born probabilities, optional noise (global depolarising mix, independent read-out bit
//                        flips), inclusive CDF, then `shots` inverse-CDF draws from the Philox stream into the histogram.
//
// Everything is fp64; the random stream is Philox4x32-10 with counter (draw pair index, basis, site, 0), key = seed.
#include "common.cuh"

namespace ddqst {

enum { SITE_SYNTH_DRAW = 16, SITE_SYNTH_GATE = 17 };
enum { SYNTH_STATE_ZERO = 0, SYNTH_STATE_PLUS = 1, SYNTH_STATE_GHZ = 2, SYNTH_STATE_RQC = 3 };

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

// amp[s] for all s with bit q: (a0, a1) -> m (a0, a1)^T, m row-major 2x2 complex
__device__ __forceinline__ void apply_1q(double2* amp, int dim, int q, const double2 (&m)[4]) {
  for (int i = threadIdx.x; i < dim / 2; i += blockDim.x) {
    const int lo = ((i >> q) << (q + 1)) | (i & ((1 << q) - 1)), hi = lo | (1 << q);
    const double2 a0 = amp[lo], a1 = amp[hi];
    amp[lo] = make_double2(m[0].x * a0.x - m[0].y * a0.y + m[1].x * a1.x - m[1].y * a1.y,
                           m[0].x * a0.y + m[0].y * a0.x + m[1].x * a1.y + m[1].y * a1.x);
    amp[hi] = make_double2(m[2].x * a0.x - m[2].y * a0.y + m[3].x * a1.x - m[3].y * a1.y,
                           m[2].x * a0.y + m[2].y * a0.x + m[3].x * a1.y + m[3].y * a1.x);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) synth_state_kernel(int N, int kind, int depth, uint64_t seed, double2* __restrict__ psi) {
  extern __shared__ __align__(16) uint8_t synth_smem[];
  double2* amp = reinterpret_cast<double2*>(synth_smem);
  const int dim = 1 << N;
  const double r = 0.70710678118654752440;
  for (int s = threadIdx.x; s < dim; s += blockDim.x) amp[s] = make_double2(s == 0 ? 1.0 : 0.0, 0.0);
  __syncthreads();
  const double2 H[4] = {{r, 0}, {r, 0}, {r, 0}, {-r, 0}};
  if (kind == SYNTH_STATE_PLUS) {
    for (int q = 0; q < N; ++q) apply_1q(amp, dim, q, H);
  } else if (kind == SYNTH_STATE_GHZ) {            // H on qubit 0, CNOT cascade (SS/data_gen.py:22-26): (|0..0> + |1..1>)/sqrt2
    for (int s = threadIdx.x; s < dim; s += blockDim.x) amp[s] = make_double2((s == 0 || s == dim - 1) ? r : 0.0, 0.0);
    __syncthreads();
  } else if (kind == SYNTH_STATE_RQC) {
    for (int layer = 0; layer < depth; ++layer) {
      for (int q = 0; q < N; ++q) {
        // U3(theta, phi, lambda) = [[cos(t/2), -e^{i lambda} sin(t/2)], [e^{i phi} sin(t/2), e^{i(phi+lambda)} cos(t/2)]]
        const Philox4 p = philox4x32_10((uint32_t)layer, (uint32_t)q, (uint32_t)SITE_SYNTH_GATE << 16, 0u, (uint32_t)seed,
                                        (uint32_t)(seed >> 32));
        const double two_pi = 6.283185307179586476925;
        const double theta = acos(1.0 - 2.0 * ((double)p.x * (1.0 / 4294967296.0)));      // Haar-distributed polar angle
        const double phi = two_pi * ((double)p.y * (1.0 / 4294967296.0)), lam = two_pi * ((double)p.z * (1.0 / 4294967296.0));
        const double c = cos(0.5 * theta), s = sin(0.5 * theta);
        const double2 U[4] = {{c, 0.0}, {-cos(lam) * s, -sin(lam) * s}, {cos(phi) * s, sin(phi) * s},
                              {cos(phi + lam) * c, sin(phi + lam) * c}};
        apply_1q(amp, dim, q, U);
      }
      for (int s = threadIdx.x; s < dim; s += blockDim.x) {         // CZ on pairs (q, q+1), q = layer parity, +2, ...
        int sign = 0;
        for (int q = layer & 1; q + 1 < N; q += 2) sign ^= (s >> q) & (s >> (q + 1)) & 1;
        if (sign) amp[s] = make_double2(-amp[s].x, -amp[s].y);
      }
      __syncthreads();
    }
  }
  for (int s = threadIdx.x; s < dim; s += blockDim.x) psi[s] = amp[s];
}

__global__ void __launch_bounds__(256) synth_born_kernel(const double2* __restrict__ psi, int N, const int32_t* __restrict__ basis_ids,
                                                         int64_t shots, int64_t shots_per_chunk, uint64_t seed, double p_depol,
                                                         double p_readout, uint32_t* __restrict__ hist, double* __restrict__ probs_out) {
  extern __shared__ __align__(16) uint8_t synth_smem[];
  const int dim = 1 << N;
  double2* amp = reinterpret_cast<double2*>(synth_smem);
  double* cdf = reinterpret_cast<double*>(synth_smem + (size_t)dim * 16);
  uint32_t* bins = reinterpret_cast<uint32_t*>(synth_smem + (size_t)dim * 24);
  __shared__ double wsum[8];
  __shared__ double carry_s;
  const int slot = blockIdx.x;
  const int basis = basis_ids ? basis_ids[slot] : slot;
  const double r = 0.70710678118654752440;
  for (int s = threadIdx.x; s < dim; s += blockDim.x) { amp[s] = psi[s]; bins[s] = 0; }
  __syncthreads();
  // letter of qubit q = digit q of the basis index in base 3, letter 0 slowest (X=0, Y=1, Z=2)
  int rem = basis;
  for (int q = N - 1; q >= 0; --q) {
    const int letter = rem % 3;
    rem /= 3;
    if (letter == 0) {
      const double2 H[4] = {{r, 0}, {r, 0}, {r, 0}, {-r, 0}};
      apply_1q(amp, dim, q, H);
    } else if (letter == 1) {                       // H . Sdg = [[1, -i], [1, i]] / sqrt2
      const double2 HS[4] = {{r, 0}, {0, -r}, {r, 0}, {0, r}};
      apply_1q(amp, dim, q, HS);
    }
  }
  for (int s = threadIdx.x; s < dim; s += blockDim.x) {
    const double2 a = amp[s];
    cdf[s] = (1.0 - p_depol) * (a.x * a.x + a.y * a.y) + p_depol / (double)dim;
  }
  __syncthreads();
  if (p_readout > 0.0) {                             // independent bit flips on the classical outcome (AS/data_gen.py:44-46)
    for (int q = 0; q < N; ++q) {
      for (int i = threadIdx.x; i < dim / 2; i += blockDim.x) {
        const int lo = ((i >> q) << (q + 1)) | (i & ((1 << q) - 1)), hi = lo | (1 << q);
        const double p0 = cdf[lo], p1 = cdf[hi];
        cdf[lo] = (1.0 - p_readout) * p0 + p_readout * p1;
        cdf[hi] = p_readout * p0 + (1.0 - p_readout) * p1;
      }
      __syncthreads();
    }
  }
  if (probs_out && blockIdx.y == 0)
    for (int s = threadIdx.x; s < dim; s += blockDim.x) probs_out[(int64_t)slot * dim + s] = cdf[s];
  __syncthreads();
  // inclusive scan in place, 256 entries per pass (warp scans + carried offset)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  for (int base = 0; base < dim; base += 256) {
    const int i = base + threadIdx.x;
    double v = i < dim ? cdf[i] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += t; }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    double off = carry_s;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (i < dim) cdf[i] = v + off;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = v + off;
    __syncthreads();
  }
  const double total = cdf[dim - 1];
  // draws [j0, j1) of this basis; pair k = j / 2 shares one Philox block
  const int64_t j0 = (int64_t)blockIdx.y * shots_per_chunk, j1 = min(shots, j0 + shots_per_chunk);
  for (int64_t k = j0 / 2 + threadIdx.x; 2 * k < j1; k += blockDim.x) {
    const Philox4 p = philox4x32_10((uint32_t)k, (uint32_t)basis, ((uint32_t)SITE_SYNTH_DRAW << 16), (uint32_t)(k >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t j = 2 * k + h;
      if (j < j0 || j >= j1) continue;
      const double u = (h == 0 ? u53(p.x, p.y) : u53(p.z, p.w)) * total;
      int a = 0, b = dim - 1;                       // first s with cdf[s] > u
      while (a < b) { const int mid = (a + b) >> 1; if (cdf[mid] > u) b = mid; else a = mid + 1; }
      atomicAdd(bins + a, 1u);
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < dim; s += blockDim.x) {
    const uint32_t c = bins[s];
    if (c) atomicAdd(hist + (int64_t)slot * dim + s, c);
  }
}

}  // namespace ddqst

using namespace ddqst;

extern "C" {

int ddqst_synth_state(int32_t num_qubits, int kind, int32_t depth, uint64_t seed, double* psi, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(num_qubits >= 1 && num_qubits <= 12 && kind >= 0 && kind <= 3 && depth >= 0 && psi, DDQST_EINVAL_SHAPE, "bad argument");
  const size_t smem = (size_t)16 << num_qubits;
  if (smem > 48 * 1024) DDQST_CUDA_OK(cudaFuncSetAttribute(synth_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  synth_state_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(num_qubits, kind, depth, seed, (double2*)psi);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_synth_born_histograms(const double* psi, int32_t num_qubits, const int32_t* basis_ids, int32_t n_bases, int64_t shots,
                                uint64_t seed, double p_depolarizing, double p_readout, uint32_t* hist, double* probs_out,
                                void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(num_qubits >= 1 && num_qubits <= 12 && n_bases >= 0 && shots >= 0 && shots < ((int64_t)1 << 32), DDQST_EINVAL_SHAPE, "bad shape");
  DDQST_REQUIRE(p_depolarizing >= 0.0 && p_depolarizing <= 1.0 && p_readout >= 0.0 && p_readout <= 1.0, DDQST_EINVAL_SHAPE, "noise rate outside [0,1]");
  if (n_bases == 0) return DDQST_OK;
  DDQST_REQUIRE(psi && hist, DDQST_EINVAL_SHAPE, "NULL argument");
  const size_t smem = (size_t)28 << num_qubits;
  if (smem > 48 * 1024) DDQST_CUDA_OK(cudaFuncSetAttribute(synth_born_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // enough CTAs to fill the machine when there are few bases: split the shots of a basis into chunks (even sizes)
  int chunks = 1;
  const int want = num_sms() * 4;
  if (n_bases < want && shots > 4096) {
    chunks = (want + n_bases - 1) / n_bases;
    const int64_t max_chunks = (shots + 4095) / 4096;
    if (chunks > max_chunks) chunks = (int)max_chunks;
    if (chunks > 65535) chunks = 65535;
  }
  int64_t per = (shots + chunks - 1) / chunks;
  per += per & 1;
  synth_born_kernel<<<dim3((unsigned)n_bases, (unsigned)chunks), 256, smem, (cudaStream_t)stream>>>(
      (const double2*)psi, num_qubits, basis_ids, shots, per, seed, p_depolarizing, p_readout, hist, probs_out);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // extern "C"
