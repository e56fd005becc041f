// Dataset unrolling on the device: the reference turns every measurement record's counts dict into
// `count` copies of (bits, basis_idx) in a Python list (RQC/dataset.py:47-65, SS/dataset.py:14-33) and lets a
// DataLoader(shuffle=True) draw batches from it.  Here the counts stay a table hist[n_rows, 2^N] (one row per
// measurement record, outcome index s = sum_q bit_q << q, i.e. the reference's reversed bit order) and a shot is
// addressed by its position p in the canonical unrolled order (row-major, outcomes ascending inside a row):
//
//   counts_scan   : cum[row, s] = inclusive running count inside the row, row_start[row] = shots before the row
//   counts_gather : for output j: p = start + j (optionally sent through a keyed bijection of [0, total) = one epoch's
//                   shuffle), row = last row with row_start <= p, s = first outcome with cum > p - row_start[row];
//                   writes the packed bitstring and the row's basis index.
//
// Nothing of size `total` is ever materialised unless the caller asks for the whole unroll; a training batch costs
// two binary searches per sample (L2-resident tables) and 6 bytes of HBM writes.
#include "common.cuh"

namespace ddqst {

// ---- keyed bijection of [0, total): 4-round balanced Feistel network on 2h bits (4^h >= total) + cycle walking
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
struct FeistelKey { uint32_t k[4]; int h; uint32_t mask; };

__host__ __device__ __forceinline__ FeistelKey feistel_key(uint64_t seed, uint64_t epoch, int64_t total) {
  FeistelKey f;
  int h = 1;
  while (h < 31 && ((int64_t)1 << (2 * h)) < total) ++h;
  f.h = h;
  f.mask = (uint32_t)(((uint64_t)1 << h) - 1u);
  const uint32_t base = (uint32_t)seed ^ fmix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(epoch + 1));
  for (int r = 0; r < 4; ++r) f.k[r] = fmix32(base + (uint32_t)r * 0x85EBCA6Bu);
  return f;
}
__host__ __device__ __forceinline__ int64_t feistel_perm(int64_t i, int64_t total, const FeistelKey& f) {
  uint64_t x = (uint64_t)i;
  do {
    uint32_t L = (uint32_t)(x >> f.h), R = (uint32_t)x & f.mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t nr = L ^ (fmix32(R ^ f.k[r]) & f.mask);
      L = R; R = nr;
    }
    x = ((uint64_t)L << f.h) | R;
  } while ((int64_t)x >= total);
  return (int64_t)x;
}

// one block per row: inclusive scan of the row's 2^N counts, row total out
__global__ void __launch_bounds__(256) counts_scan_kernel(const uint32_t* __restrict__ hist, int dim, uint32_t* __restrict__ cum,
                                                          int64_t* __restrict__ row_total) {
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t carry_s;
  const int64_t row = blockIdx.x;
  const uint32_t* h = hist + row * dim;
  uint32_t* c = cum + row * dim;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < dim; base += 256) {
    const int i = base + threadIdx.x;
    uint32_t v = i < dim ? h[i] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += t; }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    uint32_t off = carry_s;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (i < dim) c[i] = v + off;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = v + off;
    __syncthreads();
  }
  if (threadIdx.x == 0) row_total[row] = (int64_t)carry_s;
}

// single block: row_start[0..n_rows] = exclusive scan of row_total (int64)
__global__ void __launch_bounds__(1024) row_start_kernel(const int64_t* __restrict__ row_total, int64_t n_rows,
                                                         int64_t* __restrict__ row_start) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_rows; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t own = i < n_rows ? row_total[i] : 0;
    int64_t v = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int64_t t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += t; }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    int64_t off = carry_s;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (i < n_rows) row_start[i] = v + off - own;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = v + off;
    __syncthreads();
  }
  if (threadIdx.x == 0) row_start[n_rows] = carry_s;
}

__global__ void __launch_bounds__(256) counts_gather_kernel(const uint32_t* __restrict__ cum, const int64_t* __restrict__ row_start,
                                                            const int32_t* __restrict__ row_basis, int64_t n_rows, int dim,
                                                            int64_t total, int permute, FeistelKey key, int64_t start,
                                                            int64_t count, uint16_t* __restrict__ out_x0,
                                                            int32_t* __restrict__ out_basis, int64_t* __restrict__ out_bits, int N) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  int64_t p = (start + j) % total;
  if (permute) p = feistel_perm(p, total, key);
  // last row with row_start[row] <= p (empty rows share a start with their successor and are skipped by taking the last)
  int64_t lo = 0, hi = n_rows;          // invariant: row_start[lo] <= p < row_start[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(row_start + mid) <= p) lo = mid; else hi = mid;
  }
  const uint32_t r = (uint32_t)(p - __ldg(row_start + lo));
  const uint32_t* c = cum + lo * dim;
  int a = 0, b = dim - 1;               // first s with cum[s] > r
  while (a < b) {
    const int mid = (a + b) >> 1;
    if (__ldg(c + mid) > r) b = mid; else a = mid + 1;
  }
  if (out_x0) out_x0[j] = (uint16_t)a;
  if (out_basis) out_basis[j] = row_basis ? __ldg(row_basis + lo) : (int32_t)lo;
  if (out_bits)
    for (int q = 0; q < N; ++q) out_bits[j * N + q] = (a >> q) & 1;
}

}  // namespace ddqst

using namespace ddqst;

extern "C" {

int ddqst_counts_scan(const uint32_t* hist, int64_t n_rows, int32_t num_qubits, uint32_t* cum, int64_t* row_start,
                      void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(n_rows >= 0 && num_qubits >= 1 && num_qubits <= 16, DDQST_EINVAL_SHAPE, "bad shape");
  DDQST_REQUIRE(row_start && (n_rows == 0 || (hist && cum)), DDQST_EINVAL_SHAPE, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  // row totals are staged in row_start[1..n_rows] and scanned in place (row_start has n_rows + 1 entries)
  if (n_rows > 0) {
    counts_scan_kernel<<<(unsigned)n_rows, 256, 0, s>>>(hist, 1 << num_qubits, cum, row_start + 1);
    DDQST_LAUNCH_OK();
  }
  // the exclusive scan reads total[i] = row_start[i + 1] and writes row_start[i]: a thread only overwrites entries that
  // lower-indexed threads have already consumed within the same 1024-chunk (read into registers before the barrier)
  row_start_kernel<<<1, 1024, 0, s>>>(row_start + 1, n_rows, row_start);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

int ddqst_counts_gather(const uint32_t* cum, const int64_t* row_start, const int32_t* row_basis, int64_t n_rows,
                        int32_t num_qubits, int64_t total, int permute, uint64_t seed, uint64_t epoch, int64_t start,
                        int64_t count, uint16_t* out_x0_packed, int32_t* out_basis, int64_t* out_bits, void* stream) {
  DDQST_TRY(check_arch());
  DDQST_REQUIRE(n_rows >= 1 && num_qubits >= 1 && num_qubits <= 16 && total >= 1 && count >= 0 && start >= 0, DDQST_EINVAL_SHAPE,
                "bad shape (n_rows=%lld total=%lld count=%lld)", (long long)n_rows, (long long)total, (long long)count);
  if (count == 0) return DDQST_OK;
  DDQST_REQUIRE(cum && row_start, DDQST_EINVAL_SHAPE, "NULL argument");
  const FeistelKey key = feistel_key(seed, epoch, total);
  counts_gather_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      cum, row_start, row_basis, n_rows, 1 << num_qubits, total, permute, key, start, count, out_x0_packed, out_basis, out_bits,
      num_qubits);
  DDQST_LAUNCH_OK();
  return DDQST_OK;
}

}  // extern "C"
