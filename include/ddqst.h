/*
 * ddqst.h -- C ABI of libddqst.so: the B200 (sm_100a) implementation of the DD-QST
 * generative-tomography hot path.
 *
 * The reference (anik-m/...DD-QST, pure Python) has no FFI; its boundary is the Python call
 * surface of versions/<phase>/{model,diffusion,reconstruct}.py.  Each entry point below names the
 * reference symbol (file:line, relative to /root/reference/versions/) whose arithmetic it replaces;
 * RQC = RQC_dataset_building_phase, SS = multi_qubit_special_states, NB cK:L = notebook cell K line L.
 * The Python mirror of that surface lives in the package and calls these through ctypes.
 *
 * Conventions
 *   - every pointer is caller-owned DEVICE memory unless the name ends in _host;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronises
 *     the device, allocates nothing, and takes scratch from the caller's workspace
 *     (size from ddqst_workspace_bytes);  *_host entry points are the exception: they copy,
 *     launch and synchronise `stream` themselves (the "host buffers" end-to-end form);
 *   - return 0 on success, a negative ddqst_status otherwise; text from ddqst_last_error()
 *     (thread-local);
 *   - there is NO CPU fallback: a device that is not compute capability 10.x gets
 *     DDQST_EUNSUPPORTED_ARCH.
 *   - bitstrings: outcome index s = sum_q bit_q << q (column q of the reference's [B,N] tensors =
 *     qubit q = bit q).  "packed" = one uint8 per shot when N <= 8, one uint16 when 9 <= N <= 16.
 *   - index contract: timestep arrays (t in [0, num_timesteps]; [1, num_timesteps] for sampling) and basis index
 *     arrays (in [0, num_bases)) that live in DEVICE memory are used to index the embedding / FiLM / schedule
 *     tables unchecked -- validating them would cost a device synchronisation per call.  The caller owns that
 *     contract; the Python mirror enforces it (IndexError, as nn.Embedding / Q_bar[t] raise in the reference,
 *     RQC/model.py:59-61, RQC/diffusion.py:48) on every entry that is not inside a CUDA-graph capture.
 *     Scalar and host-array indices (ddqst_sample_step, ddqst_sample_host, ddqst_mlp_sample) are checked here
 *     and fail with DDQST_EINVAL_SHAPE.
 *   - randomness: Philox4x32-10, counter (shot_lo32, stream, t | site<<16, (q>>2) | shot_hi24<<8),
 *     key = seed; word for qubit q = lane q&3; u = (word>>8) * 2^-24; draw bit = u*(p0+p1) < p1.
 *     `stream` is the basis index when sampling and the step / call counter when noising.
 */
#ifndef DDQST_H_
#define DDQST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  DDQST_OK = 0,
  DDQST_EINVAL_SHAPE = -1,
  DDQST_EUNSUPPORTED_ARCH = -2,
  DDQST_EWORKSPACE = -3,
  DDQST_ECUDA = -4,
  DDQST_EUNSUPPORTED = -5
} ddqst_status;

/* draw sites of the random stream (ctr2 high half) */
enum { DDQST_SITE_INIT = 0, DDQST_SITE_POSTERIOR = 1, DDQST_SITE_X0HAT = 2, DDQST_SITE_RENOISE = 3,
       DDQST_SITE_QSAMPLE = 4, DDQST_SITE_TSTEP = 5 };

/* model variants: ConditionalD3PM front ends */
enum { DDQST_VARIANT_A = 0 /* SS/model.py:56,70  input_proj = Linear(N,H) on x.float() */,
       DDQST_VARIANT_B = 1 /* RQC/model.py:31-35 x_emb(2,E) -> flatten -> Linear(N*E,H)  */ };

/* reverse-sampler modes */
enum { DDQST_MODE_POSTERIOR = 0 /* RQC/diffusion.py:53-80 */, DDQST_MODE_RENOISE = 1 /* SS/diffusion.py:54-82 */ };

/* arithmetic of the denoiser GEMMs */
enum { DDQST_PRECISION_FP32 = 0 /* CUDA-core fp32, the "exact" mode used for end-to-end parity */,
       DDQST_PRECISION_BF16 = 1 /* tcgen05 bf16 x bf16 -> fp32 accumulators in TMEM (production) */ };

/* what `target` points to in ddqst_recon_report */
enum { DDQST_TARGET_NONE = 0, DDQST_TARGET_STATEVECTOR = 1 /* complex128[2^N] */, DDQST_TARGET_MIXED = 2 /* complex128[2^N,2^N] */,
       DDQST_TARGET_RANK_ONE = 3 /* complex128[2^N,2^N] known to be |psi><psi| (RQC/evaluate.py:71 wraps the clean state
                                    vector in a DensityMatrix): F = Tr(sigma rho), no second eigensolve */ };

/* Kronecker convention of get_pauli_matrix */
enum { DDQST_KRON_REVERSED = 0 /* RQC/reconstruct.py:19 label[::-1] */, DDQST_KRON_UNREVERSED = 1 /* SS/reconstruct.py:13-15 */ };

typedef struct {
  int32_t num_qubits, num_bases, num_timesteps, embed_dim, hidden_dim, num_blocks, variant;
} ddqst_dims;   /* ctor arguments of ConditionalD3PM, RQC/model.py:27 */

const char* ddqst_last_error(void);
int ddqst_version(void);

/* ---- parameters: ONE flat fp32 buffer in reference state_dict order (each tensor 16-byte aligned).
 * offsets_out receives element offsets for: x_emb.weight (or -1), input_proj.weight, input_proj.bias,
 * time_emb.weight, basis_emb.weight, then per block {film.net.weight, film.net.bias, net.0.weight,
 * net.0.bias, net.2.weight, net.2.bias}, output_head.weight, output_head.bias;
 * returns the total element count (RQC/model.py:27-49, SS/model.py:47-66). */
int64_t ddqst_param_count(const ddqst_dims* d, int64_t* offsets_out /* [5 + 6*L + 2] or NULL */);

/* ---- packed inference state derived from the parameters (recompute after every optimiser step):
 * input collapse c0[H], D[N,H]; FiLM tables Tt[T+1,L,2H], Tb[num_bases,L,2H] (bias folded into Tb);
 * bf16 copies of net.0 / net.2 / output_head weights; fp32 biases.  RQC/model.py:9-11,53-62. */
int64_t ddqst_pack_bytes(const ddqst_dims* d);
int ddqst_pack_weights(const ddqst_dims* d, const float* params, void* pack, void* stream);

/* ---- M2: ConditionalD3PM.forward (RQC/model.py:51-70, SS/model.py:68-85).
 * x_packed[B] uint16 bit q = qubit q, t[B] int32 in [0,T], basis[B] int32 -> logits[B,N,2] fp32. */
int ddqst_denoiser_forward(const ddqst_dims* d, const void* pack, int precision, const uint16_t* x_packed,
                           const int32_t* t, const int32_t* basis, int64_t batch, float* logits,
                           void* workspace, int64_t ws_bytes, void* stream);

/* ---- D3 / D3': DiscreteDiffusion.p_sample for many bases at once = sample(bases, n_shots).
 * sched: betas[T+1] fp32 followed by Q[T+1,2,2] fp32 (Q_bar for POSTERIOR, marginal Q for RENOISE).
 * For each i < n_bases generates shots_per_basis trajectories for basis_ids[i] with shot indices
 * shot_offset .. shot_offset+shots_per_basis-1.  Outputs (each nullable):
 *   out_packed [n_bases * shots_per_basis]   packed bitstrings (uint8 if N<=8 else uint16)
 *   out_hist   [n_bases, 2^N] uint32         per-basis outcome counts, ACCUMULATED (caller zeroes). */
int ddqst_sample(const ddqst_dims* d, const void* pack, const float* sched, int mode, int precision,
                 const int32_t* basis_ids, int32_t n_bases, int64_t shots_per_basis, int64_t shot_offset,
                 uint64_t seed, void* out_packed, uint32_t* out_hist, void* workspace, int64_t ws_bytes,
                 void* stream);

/* teacher-forced single reverse step (parity tests): x_t -> x_{t-1} for one (basis, t). */
int ddqst_sample_step(const ddqst_dims* d, const void* pack, const float* sched, int mode, int precision,
                      int32_t basis_id, int32_t t, int64_t shots, int64_t shot_offset, uint64_t seed,
                      const uint16_t* x_t, uint16_t* x_prev, float* logits_out /* nullable [shots,N,2] */,
                      void* workspace, int64_t ws_bytes, void* stream);

/* ---- D2 / D2': q_sample forward noising (RQC/diffusion.py:45-51, SS/diffusion.py:27-52).
 * Q[T+1,2,2]; cumulative=1 reads row x0 of Q_bar[t] ([from,to]); 0 reads column x0 of Q[t] ([to,from]).
 * If t == NULL the timesteps are drawn too: t = 1 + floor(u24*T) at site TSTEP (RQC/main.py:107) and
 * written to t_out (nullable).  Q must be 16-byte aligned (each Q[t] is read as one float4). */
int ddqst_q_sample(const float* Q, int32_t num_timesteps, int32_t num_qubits, int cumulative,
                   const uint16_t* x0_packed, const int32_t* t, int64_t batch, int64_t row_offset,
                   uint64_t seed, uint32_t stream_id, uint16_t* xt_packed, int32_t* t_out, void* stream);

/* the same with the stream id read from device memory (stream_id_dev[0]; the device-side step counter that
 * ddqst_adam_step_dev increments), so a whole training step can be captured in a CUDA graph and replayed. */
int ddqst_q_sample_dev(const float* Q, int32_t num_timesteps, int32_t num_qubits, int cumulative,
                       const uint16_t* x0_packed, const int32_t* t, int64_t batch, int64_t row_offset,
                       uint64_t seed, const int64_t* stream_id_dev, uint16_t* xt_packed, int32_t* t_out, void* stream);

/* ---- H0: per-basis histogram of packed bitstrings (elem_bytes 1 or 2), counts ACCUMULATED into
 * hist[2^N] uint32. */
int ddqst_histogram(const void* packed, int elem_bytes, int64_t n, int32_t num_qubits, uint32_t* hist, void* stream);

/* int64[B,N] {0,1} (the reference's sample tensor layout, RQC/dataset.py:64) <-> packed uint16 */
int ddqst_pack_bits(const int64_t* bits, int64_t batch, int32_t num_qubits, uint16_t* packed, void* stream);
int ddqst_unpack_bits(const void* packed, int elem_bytes, int64_t batch, int32_t num_qubits, int64_t* bits, void* stream);

/* ---- R1+R2+R3: linear_inversion before the PSD step (RQC/reconstruct.py:26-46,5-24,56-66).
 * hist[n_slots, 2^N] uint32, shots[n_slots] (row sums; NULL = derived from hist in the kernel; a zero-shot basis
 * gives 0/0 = NaN as np.mean([]) does in the reference), sel[4^N] int32 = histogram
 * slot feeding each Pauli string in product order (-1: none compatible -> 0.0, -2: identity -> 1.0;
 * NULL = complete product-order data: slot = P with I->X).  rho: complex128[2^N,2^N] row-major,
 * OVERWRITTEN.  workspace: n_slots * 2^N int32 (Walsh-Hadamard coefficients). */
int ddqst_linear_inversion(const uint32_t* hist, const int64_t* shots, int32_t n_slots, int32_t num_qubits,
                           const int32_t* sel, int kron, double* rho, void* workspace, int64_t ws_bytes,
                           void* stream);

/* ---- R4: make_positive_semidefinite (RQC/reconstruct.py:48-54): Hermitian eigendecomposition
 * (parallel cyclic Jacobi, fp64), clip, renormalise, rebuild; in place on rho[dim,dim] complex128.
 * evals_out (nullable) [dim] receives the clipped, renormalised spectrum.
 * workspace: 2 * 16 * dim^2 + 8 * dim + 1024 bytes; with 24 * dim^2 bytes more (64 <= dim <= 1024) the Jacobi sweeps start in fp32
 * and only the last 2-3 run in fp64 (same result, ~35 % less time). */
int ddqst_psd_project(double* rho, int32_t dim, double* evals_out, void* workspace, int64_t ws_bytes, void* stream);

/* ---- F1: state_fidelity (qiskit.quantum_info; call sites RQC/evaluate.py:77,87, SS/main.py:127).
 * pure target: <psi|rho|psi>; mixed target: (sum sqrt eig(sqrt(a) b sqrt(a)))^2. out: 1 double. */
int ddqst_fidelity_pure(const double* psi, const double* rho, int32_t dim, double* out, void* stream);
int ddqst_fidelity_mixed(const double* rho_a, const double* rho_b, int32_t dim, double* out,
                         void* workspace, int64_t ws_bytes, void* stream);

/* ---- get_metrics (RQC/reconstruct.py:69-76): out[3] = purity, von Neumann entropy (bits),
 * entanglement entropy of the low num_qubits/2 qubits. */
int ddqst_metrics(const double* rho, int32_t num_qubits, double* out, void* workspace, int64_t ws_bytes, void* stream);

/* ---- R4 + get_metrics + F1 from ONE full eigendecomposition (the evaluation loop RQC/evaluate.py:75-88 calls
 * linear_inversion, state_fidelity and get_metrics back to back on the same rho):
 *   rho[2^N,2^N] complex128: raw Hermitian in, PSD-projected (RQC/reconstruct.py:48-54) out;
 *   target / target_kind: see DDQST_TARGET_*;  evals_out (nullable) [2^N]: clipped, renormalised spectrum;
 *   report[5] = { fidelity (0 when no target), purity Tr rho^2, von Neumann entropy (bits),
 *                 entanglement entropy of the low N/2 qubits (bits), reserved }.
 * Eigensolves: one of size 2^N, one of size 2^(N/2), plus one of size 2^N only for DDQST_TARGET_MIXED.
 * workspace: 3 (5 for DDQST_TARGET_MIXED) * 16 * 4^N + 16 * 2^N + 2048 bytes (+ 24 * 4^N optional, as for ddqst_psd_project). */
int ddqst_recon_report(double* rho, int32_t num_qubits, const double* target, int target_kind, double* evals_out,
                       double* report, void* workspace, int64_t ws_bytes, void* stream);

/* ---- T1: training step (RQC/main.py:105-115): forward with saved activations, mean cross-entropy,
 * backward into the flat gradient buffer (OVERWRITTEN), loss -> loss_out[0].  x_t/t as produced by
 * ddqst_q_sample.  The optimiser step is separate so a gradient all-reduce can sit between. */
int ddqst_train_forward_backward(const ddqst_dims* d, const float* params, const uint16_t* xt_packed,
                                 const uint16_t* x0_packed, const int32_t* t, const int32_t* basis,
                                 int64_t batch, float loss_scale, float* grads, float* loss_out,
                                 void* workspace, int64_t ws_bytes, void* stream);

/* the same step split in two so it can sit behind torch autograd (model(x_t,t,basis) ... loss.backward(),
 * RQC/main.py:109-113): forward keeps its activations in `workspace`; backward consumes them together with
 * dlogits[batch,N,2] and OVERWRITES grads.  The workspace must not be touched in between. */
int ddqst_forward_saved(const ddqst_dims* d, const float* params, const uint16_t* xt_packed, const int32_t* t,
                        const int32_t* basis, int64_t batch, float* logits_out, void* workspace, int64_t ws_bytes,
                        void* stream);
int ddqst_backward_saved(const ddqst_dims* d, const float* params, const uint16_t* xt_packed, const int32_t* t,
                         const int32_t* basis, int64_t batch, const float* dlogits, float* grads, void* workspace,
                         int64_t ws_bytes, void* stream);

/* torch.optim.Adam / AdamW semantics (RQC/main.py:98 Adam lr 1e-3; SS/main.py:77 AdamW lr 1e-4, wd 0.01):
 * step is the 1-based step count; decoupled != 0 selects AdamW. */
int ddqst_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                    int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                    int decoupled, float grad_scale, void* stream);

/* ---- T1 on the tensor cores (DDQST_PRECISION_BF16; RQC model variant): the same step with every GEMM of the
 * forward, data-gradient and weight-gradient passes on tcgen05 (bf16 operands, fp32 accumulation in TMEM, 3-D TMA
 * operands; the stored [out,in] weights serve the backward pass as MN-major operands).  params_bf16 is a bf16 copy of
 * the flat parameter buffer at the same element offsets (ddqst_cast_bf16 once, then kept current by
 * ddqst_adam_step_dev).  workspace: ddqst_workspace_bytes(DDQST_OP_TRAIN, d, batch, DDQST_PRECISION_BF16). */
int ddqst_cast_bf16(const float* src, uint16_t* dst_bf16, int64_t n, void* stream);
int ddqst_train_forward_backward_tc(const ddqst_dims* d, const float* params, const uint16_t* params_bf16,
                                    const uint16_t* xt_packed, const uint16_t* x0_packed, const int32_t* t,
                                    const int32_t* basis, int64_t batch, float loss_scale, float* grads, float* loss_out,
                                    void* workspace, int64_t ws_bytes, void* stream);
/* Adam / AdamW with the count of completed steps in device memory (step_dev: int64[2], [0] = completed steps, incremented
 * by the call, [1] = scratch that must start at 0) and an optional bf16 shadow of the updated parameters (params_bf16 nullable): nothing in the call depends on host state, so it
 * can be replayed from a CUDA graph. */
int ddqst_adam_step_dev(float* params, uint16_t* params_bf16, const float* grads, float* exp_avg, float* exp_avg_sq,
                        int64_t n, int64_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int decoupled, float grad_scale, void* stream);

/* ---- dataset unrolling (RQC/dataset.py:47-65, SS/dataset.py:14-33; call site RQC/main.py:84-92 DataLoader(shuffle=True)).
 * hist[n_rows, 2^N] uint32: one row per measurement record, outcome index s = sum_q bit_q << q (= the reference's
 * reversed counts key).  A shot is addressed by its position in the canonical unrolled order (row-major, outcomes
 * ascending inside a row).  counts_scan: cum[n_rows, 2^N] = inclusive running counts per row, row_start[n_rows + 1] =
 * shots before each row (row_start[n_rows] = total).  counts_gather: for j < count, position p = (start + j) % total,
 * sent through a keyed bijection of [0, total) when permute != 0 (4-round Feistel network + cycle walking, keyed by
 * (seed, epoch): one epoch of DataLoader(shuffle=True)); outputs (each nullable): packed bitstring, basis index of the
 * row (row_basis[row], or the row number when row_basis is NULL), and the reference's int64[count, N] bit layout. */
int ddqst_counts_scan(const uint32_t* hist, int64_t n_rows, int32_t num_qubits, uint32_t* cum, int64_t* row_start,
                      void* stream);
int ddqst_counts_gather(const uint32_t* cum, const int64_t* row_start, const int32_t* row_basis, int64_t n_rows,
                        int32_t num_qubits, int64_t total, int permute, uint64_t seed, uint64_t epoch, int64_t start,
                        int64_t count, uint16_t* out_x0_packed, int32_t* out_basis, int64_t* out_bits, void* stream);

/* ---- synthetic measurement data (stand-in for the Qiskit-Aer generation of SS/data_gen.py:40-63,
 * AS/data_gen.py:59-140, RQC/batch_build_dataset.py:53-144; qiskit is not a dependency of this library).
 * synth_state: psi complex128[2^N] (bit i of the index = qubit i) <- kind 0 |0..0>, 1 |+..+>, 2 GHZ/Bell (SS/data_gen.py:22-26),
 * 3 brick-wall random circuit of `depth` layers (random U3 per qubit, CZ on alternating neighbour pairs; gate angles from
 * the Philox stream keyed by seed).  synth_born_histograms: for each listed basis (product-order index, NULL = 0..n-1)
 * rotate into the basis (H for X, H.Sdg for Y), Born probabilities, noise (p_depolarizing: (1-p) P + p/2^N;
 * p_readout: independent flip of every measured bit), then `shots` inverse-CDF draws -> hist[n_bases, 2^N] uint32
 * ACCUMULATED (caller zeroes); probs_out (nullable) [n_bases, 2^N] float64 receives the sampled distributions. */
int ddqst_synth_state(int32_t num_qubits, int kind, int32_t depth, uint64_t seed, double* psi, void* stream);
int ddqst_synth_born_histograms(const double* psi, int32_t num_qubits, const int32_t* basis_ids, int32_t n_bases,
                                int64_t shots, uint64_t seed, double p_depolarizing, double p_readout, uint32_t* hist,
                                double* probs_out, void* stream);

/* ---- M5 + the notebook DDM (single-qubit phase, config C1): SimpleMLP (NB c6:65-102: embed 32, hidden 128,
 * num_hidden 2) and UpgradedMLP (NB c12:58-94: embed 128, hidden 256, num_hidden 3): cat[x, t_emb, b_emb] -> Linear/ReLU
 * stack -> logits[B,2].  Flat parameters in state_dict order: time_emb.weight, basis_emb.weight, then
 * net.{0,2,..}.{weight,bias}; offsets_out [2 + 2*(num_hidden+1)].  x is one bit per sample (bit 0 of a uint16).
 * forward_saved / backward_saved serve autograd for BitstringDDM.train_step (NB c6:170-187, loss = CE(logits, x_0));
 * mlp_sample is BitstringDDM.sample (NB c6:189-221): sched = [T+1 unused floats][Q[T+1,2,2]] with the notebook's
 * p_stay schedule; output one uint8 per shot and/or counts[2]. */
typedef struct { int32_t num_bases, num_timesteps, embed_dim, hidden_dim, num_hidden; } ddqst_mlp_dims;
int64_t ddqst_mlp_param_count(const ddqst_mlp_dims* d, int64_t* offsets_out);
int64_t ddqst_mlp_workspace_bytes(const ddqst_mlp_dims* d, int64_t batch);
int ddqst_mlp_forward_saved(const ddqst_mlp_dims* d, const float* params, const uint16_t* x, const int32_t* t,
                            const int32_t* basis, int64_t batch, float* logits_out, void* workspace, int64_t ws_bytes,
                            void* stream);
int ddqst_mlp_backward_saved(const ddqst_mlp_dims* d, const float* params, const int32_t* t, const int32_t* basis,
                             int64_t batch, const float* dlogits, float* grads, void* workspace, int64_t ws_bytes,
                             void* stream);
int ddqst_mlp_sample(const ddqst_mlp_dims* d, const float* params, const float* sched, int32_t basis_id, int64_t n,
                     int64_t shot_offset, uint64_t seed, uint8_t* out_bits, uint32_t* out_hist, void* workspace,
                     int64_t ws_bytes, void* stream);

/* ---- workspace sizes */
enum { DDQST_OP_FORWARD = 0, DDQST_OP_SAMPLE = 1, DDQST_OP_LINEAR_INVERSION = 2, DDQST_OP_PSD = 3,
       DDQST_OP_FIDELITY_MIXED = 4, DDQST_OP_TRAIN = 5, DDQST_OP_METRICS = 6 };
int64_t ddqst_workspace_bytes(int op, const ddqst_dims* d, int64_t batch, int precision);

/* ---- host-buffer end-to-end forms (what bench.py's e2e times): everything copied inside the call. */
int ddqst_sample_host(const ddqst_dims* d, const void* pack /* device */, const float* sched /* device */,
                      int mode, int precision, const int32_t* basis_ids_host, int32_t n_bases,
                      int64_t shots_per_basis, int64_t shot_offset, uint64_t seed,
                      void* out_packed_host /* nullable, pinned */, uint32_t* out_hist_host /* nullable */,
                      void* dev_scratch, int64_t dev_scratch_bytes, void* stream);

/* ---- self tests (used by tests/ only) */
int ddqst_selftest_philox(const uint32_t* ctr_key /* [n,6] */, int64_t n, uint32_t* out /* [n,4] */, void* stream);
/* one tcgen05 GEMM C[M=128*mt, N] = A[M,K] bf16 . W[N,K]^T bf16 through the sampler's operand paths */
int ddqst_selftest_umma(const float* a /* [M,K] fp32 */, const uint16_t* w_bf16 /* [N,K] */, int32_t m_tiles,
                        int32_t n, int32_t k, float* c /* [M,N] */, void* stream);
/* the same through one cta_group::2 MMA: M = 256*m_pairs rows, n <= 256, CTA pair shares the B operand */
int ddqst_selftest_umma2(const float* a, const uint16_t* w_bf16, int32_t m_pairs, int32_t n, int32_t k, float* c,
                         void* stream);
/* the training GEMM kernel alone: C[batch][m,n] fp32 = A . B^T over bf16 operands stored either way round
 * (a_mn == 0: A is [m,k] row-major, else [k,m]; b_mn == 0: B is [n,k] row-major, else [k,n]); n % 4 == 0 */
int ddqst_selftest_gemm_tc(const uint16_t* a, const uint16_t* b, int a_mn, int b_mn, int32_t m, int32_t n, int32_t k,
                           int32_t batch, float* c, void* stream);
/* debugging aids of the training GEMM kernel (benchmarks/gemm_tc_stamps.py, benchmarks/train_trace.py):
 * selftest_gemm_tc_dbg additionally writes 8 clock64 stamps of CTA (0,0,0) to dbg (device int64[8]);
 * debug_tc_trace registers (or, with NULL, clears) a device buffer of 4*cap int64: GEMM launch i of the following
 * ddqst_train_forward_backward_tc calls writes its %globaltimer entry / dependency-release / epilogue-done stamps and its
 * epilogue id to buf[4*i .. 4*i+3] (process-wide state, not thread-safe: debugging only). */
int ddqst_selftest_gemm_tc_dbg(const uint16_t* a, const uint16_t* b, int a_mn, int b_mn, int32_t m, int32_t n, int32_t k,
                               int32_t batch, float* c, long long* dbg, void* stream);
int ddqst_debug_tc_trace(long long* buf, int32_t cap);
/* synchronises the device; returns the first pipeline-timeout code a kernel recorded (0 = none).  Every mbarrier / mailbox / grid-barrier
 * wait in the tcgen05 kernels and in the eigensolver is bounded (~1 s); on a timeout the kernel drains with garbage instead of hanging
 * and the first offender's code stays here: 1-49 sampler and training GEMM pipelines, 60-65 one-cluster Jacobi kernels, 66-70 line
 * eigensolver (inbox, mailbox, sweep barrier), 71-73 block eigensolver (mailboxes, sweep barrier). */
int ddqst_debug_tc_status(void);
/* which forward + data-gradient implementation ddqst_train_forward_backward_tc uses: -1 = by batch size (default; the fused
 * persistent kernel of csrc/train_fused.cuh from 3072 rows, per-layer GEMM launches below), 0 = per-layer, 1 = fused. */
int ddqst_debug_train_path(int mode);
/* debugging aid: with DDQST_FT_DEBUG=1 the fused training kernel records clock64 stamps of its first tile (csrc/train_fused.cuh);
 * copies the 256 int64 values to HOST memory (synchronises the device). */
int ddqst_debug_ft_stamps(long long* out256_host);

#ifdef __cplusplus
}
#endif
#endif /* DDQST_H_ */
