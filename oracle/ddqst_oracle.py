"""CPU oracle: a plain numpy / torch-CPU restatement of the DD-QST hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline and
``--impl reference`` legs) may import this module.  The product package never
does; it fails loudly when the CUDA library is missing.

How it is pinned ("parity pinned"):
  * every function below cites the reference file:line it restates
    (paths relative to /root/reference/versions; SS = multi_qubit_special_states,
    RQC = RQC_dataset_building_phase, NB cK:L = notebook cell K line L);
  * ``tests/test_oracle_vs_reference.py`` imports the UNMODIFIED reference
    modules (when /root/reference is mounted, i.e. in the build container),
    injects the same counter-based random stream into ``torch.randint`` /
    ``torch.multinomial`` and requires bit-identical outputs;
  * ``tests/golden/make_golden.py`` ran the reference itself to produce the
    committed fixtures under ``tests/golden/`` which the oracle must reproduce
    on any machine (the GPU box has no /root/reference);
  * the four notebook fidelities (NB c9:38, c10:38, c16:74, notes.pdf p.5)
    are known-answer tests.

The reference never seeds anything, so "identical random
streams" is defined by the injected Philox4x32-10 stream below: both the
oracle and the CUDA kernels derive every random draw from
(seed, stream, t, site, global sample index, qubit) so results do not depend
on batch split, launch geometry or rank count.
"""
from __future__ import annotations

import math
from itertools import product

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# 1. Counter-based random stream (Philox4x32-10, Salmon et al. SC'11)
# ----------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

SITE_INIT = 0       # x_T bits               (RQC/diffusion.py:55, SS/diffusion.py:58)
SITE_POSTERIOR = 1  # posterior draw         (RQC/diffusion.py:79)
SITE_X0HAT = 2      # x0-hat draw            (SS/diffusion.py:71, NB c6:210)
SITE_RENOISE = 3    # re-noise to t-1        (SS/diffusion.py:76 -> :49)
SITE_QSAMPLE = 4    # training noising       (RQC/diffusion.py:50, SS/diffusion.py:49)
SITE_TSTEP = 5      # training timestep draw (RQC/main.py:107)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: uint32[..., 4], key: uint32[..., 2] (broadcastable) -> uint32[..., 4]."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        if r != 9:
            k0 = (k0 + np.uint64(PHILOX_W0)) & mask
            k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def stream_words(seed: int, stream: int, t: int, site: int, sample_idx, n_qubits: int) -> np.ndarray:
    """uint32[B, n_qubits] random words of the injected stream.

    counter = (sample_lo32, stream, t | site<<16, (q>>2) | sample_hi24<<8),
    key = (seed_lo32, seed_hi32); the word for qubit q is output lane q & 3.
    """
    sample_idx = np.asarray(sample_idx, dtype=np.uint64).reshape(-1)
    B = sample_idx.shape[0]
    nblk = (n_qubits + 3) // 4
    ctr = np.zeros((B, nblk, 4), dtype=np.uint32)
    ctr[:, :, 0] = (sample_idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    ctr[:, :, 1] = np.uint32(stream & 0xFFFFFFFF)
    ctr[:, :, 2] = np.uint32((t & 0xFFFF) | ((site & 0xFFFF) << 16))
    hi = ((sample_idx >> np.uint64(32)) & np.uint64(0xFFFFFF)).astype(np.uint32)
    ctr[:, :, 3] = np.arange(nblk, dtype=np.uint32)[None, :] | (hi[:, None] << np.uint32(8))
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    out = philox4x32_10(ctr, key)  # [B, nblk, 4]
    return out.reshape(B, nblk * 4)[:, :n_qubits]


def words_to_uniform(words: np.ndarray) -> np.ndarray:
    """24-bit uniforms in [0,1), exactly representable in fp32."""
    return ((words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def stream_uniforms(seed, stream, t, site, sample_idx, n_qubits) -> np.ndarray:
    return words_to_uniform(stream_words(seed, stream, t, site, sample_idx, n_qubits))


def stream_timesteps(seed, stream, sample_idx, num_timesteps) -> np.ndarray:
    """Injected replacement for ``randint(1, T+1, (B,))`` (RQC/main.py:107): t = 1 + floor(u24 * T / 2^24)."""
    w = stream_words(seed, stream, 0, SITE_TSTEP, sample_idx, 1)[:, 0]
    k = (w >> np.uint32(8)).astype(np.uint64)
    return (1 + ((k * np.uint64(num_timesteps)) >> np.uint64(24))).astype(np.int64)


def draw_bits(probs: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """Injected replacement for ``torch.multinomial(probs[..., 2], 1)``:
    inverse-CDF on the (not necessarily normalised) pair: bit = u*(p0+p1) < p1."""
    p0, p1 = probs[..., 0], probs[..., 1]
    return (u * (p0 + p1) < p1).to(torch.int64)


# ----------------------------------------------------------------------------
# 2. Noise schedules (D1, D1')
# ----------------------------------------------------------------------------
def cosine_schedule(num_timesteps: int):
    """RQC/diffusion.py:15-43: cosine alpha-bar (float64) -> betas fp32 [T+1],
    Q_bar[t] = Q_t @ Q_bar[t-1] accumulated sequentially in fp32."""
    steps = np.arange(num_timesteps + 1, dtype=np.float64) / num_timesteps
    abar = np.cos((steps + 0.008) / 1.008 * np.pi / 2) ** 2
    abar = abar / abar[0]
    vals = [0.0]
    for i in range(1, num_timesteps + 1):
        vals.append(min(1 - abar[i] / abar[i - 1], 0.999))
    betas = torch.tensor(vals, dtype=torch.float32)
    Q_bar = torch.zeros(num_timesteps + 1, 2, 2)
    cur = torch.eye(2)
    Q_bar[0] = cur
    for t in range(1, num_timesteps + 1):
        b = betas[t]
        Qt = torch.tensor([[1 - b, b], [b, 1 - b]])
        cur = Qt @ cur
        Q_bar[t] = cur
    return betas, Q_bar


def linear_schedule(num_timesteps: int):
    """SS/diffusion.py:14-25: beta = linspace(0.001, 0.5, T+1); Q[t] = [[1-b,b],[b,1-b]]
    (used as the MARGINAL x_0 -> x_t channel, no cumulative product)."""
    betas = torch.linspace(0.001, 0.5, num_timesteps + 1)
    Q = torch.zeros(num_timesteps + 1, 2, 2)
    for t in range(num_timesteps + 1):
        b = betas[t]
        Q[t] = torch.tensor([[1 - b, b], [b, 1 - b]])
    return betas, Q


def notebook_schedule(num_timesteps: int):
    """NB c6:117-127: p_stay = linspace(1, 0.5, T+1); Q[0] stays zero."""
    p_stay = torch.linspace(1.0, 0.5, num_timesteps + 1)
    Q = torch.zeros(num_timesteps + 1, 2, 2)
    for t in range(1, num_timesteps + 1):
        p = p_stay[t]
        Q[t] = torch.tensor([[p, 1 - p], [1 - p, p]])
    return 1.0 - p_stay, Q


# ----------------------------------------------------------------------------
# 3. Denoiser forward (M2-M5), functional on a reference state_dict
# ----------------------------------------------------------------------------
def _num_blocks(sd) -> int:
    n = 0
    while f"blocks.{n}.film.net.weight" in sd:
        n += 1
    return n


def denoiser_forward(sd: dict, x: torch.Tensor, t: torch.Tensor, basis: torch.Tensor, num_qubits: int) -> torch.Tensor:
    """logits[B,N,2] fp32.  Variant B (token embedding front end) RQC/model.py:51-70;
    variant A (Linear on x.float()) SS/model.py:68-85; FiLM RQC/model.py:9-11;
    ResBlock RQC/model.py:23-24."""
    if "x_emb.weight" in sd:
        h = F.embedding(x, sd["x_emb.weight"]).reshape(x.shape[0], -1)
    else:
        h = x.float()
    h = F.linear(h, sd["input_proj.weight"], sd["input_proj.bias"])
    cond = torch.cat([F.embedding(t, sd["time_emb.weight"]), F.embedding(basis, sd["basis_emb.weight"])], dim=1)
    for i in range(_num_blocks(sd)):
        p = f"blocks.{i}."
        gamma, beta = F.linear(cond, sd[p + "film.net.weight"], sd[p + "film.net.bias"]).chunk(2, dim=1)
        u = h * (1 + gamma) + beta
        u = F.linear(u, sd[p + "net.0.weight"], sd[p + "net.0.bias"])
        u = F.linear(F.silu(u), sd[p + "net.2.weight"], sd[p + "net.2.bias"])
        h = F.silu(h + u)
    out = F.linear(h, sd["output_head.weight"], sd["output_head.bias"])
    return out.view(-1, num_qubits, 2)


def default_init_state_dict(num_qubits, num_bases, num_timesteps, embed_dim, hidden_dim, num_blocks, variant="B", seed=0):
    """A reference-shaped ``state_dict`` with torch's default initialisation, built from plain ``torch.nn`` modules in
    the reference's construction order (RQC/model.py:27-49; variant A: SS/model.py:47-66) under ``torch.manual_seed(seed)``
    -- so it equals what the reference's own constructor produces under that seed, without importing any product code."""
    import torch.nn as nn
    torch.manual_seed(seed)
    E, H, N = embed_dim, hidden_dim, num_qubits
    mods = {}
    if variant == "B":
        mods["x_emb"] = nn.Embedding(2, E)
        mods["input_proj"] = nn.Linear(N * E, H)
        mods["time_emb"] = nn.Embedding(num_timesteps + 1, E)
        mods["basis_emb"] = nn.Embedding(num_bases, E)
    else:
        mods["time_emb"] = nn.Embedding(num_timesteps + 1, E)
        mods["basis_emb"] = nn.Embedding(num_bases, E)
        mods["input_proj"] = nn.Linear(N, H)
    for i in range(num_blocks):
        mods[f"blocks.{i}.film.net"] = nn.Linear(2 * E, 2 * H)
        mods[f"blocks.{i}.net.0"] = nn.Linear(H, H)
        mods[f"blocks.{i}.net.2"] = nn.Linear(H, H)
    mods["output_head"] = nn.Linear(H, 2 * N)
    sd = {}
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


def notebook_mlp_forward(sd: dict, x: torch.Tensor, t: torch.Tensor, basis: torch.Tensor) -> torch.Tensor:
    """SimpleMLP NB c6:86-102 / UpgradedMLP NB c12:89-94: cat[x.float(), t_emb, b_emb] -> Linear/ReLU stack -> [B,2]."""
    h = torch.cat([x.float().view(-1, 1), F.embedding(t, sd["time_emb.weight"]), F.embedding(basis, sd["basis_emb.weight"])], dim=1)
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith("net.") and k.endswith(".weight")})
    for j, li in enumerate(idx):
        h = F.linear(h, sd[f"net.{li}.weight"], sd[f"net.{li}.bias"])
        if j != len(idx) - 1:
            h = F.relu(h)
    return h


# ----------------------------------------------------------------------------
# 4. Forward noising and reverse sampling (D2, D2', D3, D3')
# ----------------------------------------------------------------------------
def q_sample_cumulative(Q_bar, x_0, t, seed, stream, row_offset=0):
    """RQC/diffusion.py:45-51: probs = Q_bar[t_i][x_0[i]] (rows = from-state), one draw per (sample, qubit)."""
    B, N = x_0.shape
    probs = Q_bar[t][torch.arange(B)[:, None], x_0]            # [B, N, 2]
    u = torch.from_numpy(stream_uniforms(seed, stream, 0, SITE_QSAMPLE, row_offset + np.arange(B), N))
    return draw_bits(probs, u)


def q_sample_marginal(Q, x_0, t, seed, stream, site=SITE_QSAMPLE, t_field=0, row_offset=0):
    """SS/diffusion.py:27-52 (and NB c6:132-168): probs[to] = Q[t_i][to, x_0[i,q]] (column select), per qubit."""
    B, N = x_0.shape
    Qt = Q[t]                                                  # [B, 2(to), 2(from)]
    probs = torch.stack([Qt[torch.arange(B)[:, None], 0, x_0], Qt[torch.arange(B)[:, None], 1, x_0]], dim=-1)
    u = torch.from_numpy(stream_uniforms(seed, stream, t_field, site, row_offset + np.arange(B), N))
    return draw_bits(probs, u)


def init_bits(seed, basis_idx, num_samples, num_qubits, shot_offset=0) -> torch.Tensor:
    """Injected replacement for ``randint(0, 2, (B, N))`` (RQC/diffusion.py:55): low bit of the INIT word."""
    w = stream_words(seed, basis_idx, 0, SITE_INIT, shot_offset + np.arange(num_samples), num_qubits)
    return torch.from_numpy((w & np.uint32(1)).astype(np.int64))


def posterior_step(logits, x_t, beta_t, Q_bar_prev, u):
    """One reverse step of RQC/diffusion.py:62-79 given logits and uniforms; returns (x_{t-1}, norm probs)."""
    p_hat = F.softmax(logits, dim=2)
    one_m = 1 - beta_t
    trans = torch.stack([torch.where(x_t == 0, one_m, beta_t), torch.where(x_t == 0, beta_t, one_m)], dim=-1)
    prior = torch.matmul(p_hat, Q_bar_prev)
    unnorm = trans * prior
    norm = unnorm / (unnorm.sum(dim=-1, keepdim=True) + 1e-8)
    return draw_bits(norm, u), norm


@torch.no_grad()
def p_sample_posterior(sd, betas, Q_bar, num_samples, basis_idx, num_qubits, seed, shot_offset=0,
                       forward=None, trajectory=False):
    """RQC/diffusion.py:53-80 (true D3PM posterior; no special case at t=1)."""
    T = betas.shape[0] - 1
    fwd = forward or (lambda x, t, b: denoiser_forward(sd, x, t, b, num_qubits))
    idx = shot_offset + np.arange(num_samples)
    x_t = init_bits(seed, basis_idx, num_samples, num_qubits, shot_offset)
    b_vec = torch.full((num_samples,), basis_idx, dtype=torch.long)
    traj = [x_t.clone()] if trajectory else None
    for t in reversed(range(1, T + 1)):
        t_vec = torch.full((num_samples,), t, dtype=torch.long)
        logits = fwd(x_t, t_vec, b_vec)
        u = torch.from_numpy(stream_uniforms(seed, basis_idx, t, SITE_POSTERIOR, idx, num_qubits))
        x_t, _ = posterior_step(logits, x_t, betas[t], Q_bar[t - 1], u)
        if trajectory:
            traj.append(x_t.clone())
    return (x_t, traj) if trajectory else x_t


@torch.no_grad()
def p_sample_renoise(sd, Q, num_samples, basis_idx, num_qubits, seed, shot_offset=0, forward=None, trajectory=False):
    """SS/diffusion.py:54-82 ("predict x0 then re-noise to t-1"); NB c6:189-221 is the N=1 case."""
    T = Q.shape[0] - 1
    fwd = forward or (lambda x, t, b: denoiser_forward(sd, x, t, b, num_qubits))
    idx = shot_offset + np.arange(num_samples)
    x_t = init_bits(seed, basis_idx, num_samples, num_qubits, shot_offset)
    b_vec = torch.full((num_samples,), basis_idx, dtype=torch.long)
    traj = [x_t.clone()] if trajectory else None
    for t in reversed(range(1, T + 1)):
        t_vec = torch.full((num_samples,), t, dtype=torch.long)
        probs = F.softmax(fwd(x_t, t_vec, b_vec), dim=2)
        u = torch.from_numpy(stream_uniforms(seed, basis_idx, t, SITE_X0HAT, idx, num_qubits))
        x0_hat = draw_bits(probs, u)
        if t > 1:
            x_t = q_sample_marginal(Q, x0_hat, torch.full_like(t_vec, t - 1), seed, basis_idx,
                                    site=SITE_RENOISE, t_field=t, row_offset=shot_offset)
        else:
            x_t = x0_hat
        if trajectory:
            traj.append(x_t.clone())
    return (x_t, traj) if trajectory else x_t


# ----------------------------------------------------------------------------
# 5. Training step (T1)
# ----------------------------------------------------------------------------
def train_loss(sd, x_t, t, basis, x_0, num_qubits):
    """RQC/main.py:109-110: mean over B*N of -log softmax(logits)[x_0]."""
    logits = denoiser_forward(sd, x_t, t, basis, num_qubits)
    return F.cross_entropy(logits.permute(0, 2, 1), x_0)


def train_step(params: dict, opt: torch.optim.Optimizer, Q_sched, x_0, basis, num_qubits, num_timesteps,
               seed, step, cumulative=True, row_offset=0):
    """RQC/main.py:105-115 with the injected stream: t draw, noising, forward, CE, backward, optimiser step.
    ``params`` maps state_dict names to leaf tensors owned by ``opt``.  Returns (loss, t, x_t)."""
    B = x_0.shape[0]
    t = torch.from_numpy(stream_timesteps(seed, step, row_offset + np.arange(B), num_timesteps))
    if cumulative:
        x_t = q_sample_cumulative(Q_sched, x_0, t, seed, step, row_offset)
    else:
        x_t = q_sample_marginal(Q_sched, x_0, t, seed, step, row_offset=row_offset)
    loss = train_loss(params, x_t, t, basis, x_0, num_qubits)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.detach(), t, x_t


# ----------------------------------------------------------------------------
# 6. Histogram, linear inversion, PSD projection, fidelity, metrics (H0, R1-R4, F1)
# ----------------------------------------------------------------------------
def basis_strings(num_qubits: int):
    """Product order X<Y<Z with qubit 0 as the slowest-varying letter (RQC/dataset.py:43, RQC/evaluate.py:67)."""
    return ["".join(p) for p in product("XYZ", repeat=num_qubits)]


def pack_bits(samples: np.ndarray) -> np.ndarray:
    """[B,N] {0,1} -> outcome index s = sum_i bit_i << i (column i = qubit i = bit i)."""
    samples = np.asarray(samples)
    w = (1 << np.arange(samples.shape[1], dtype=np.int64))
    return (samples.astype(np.int64) * w).sum(axis=1)


def histogram(samples: np.ndarray, num_qubits: int) -> np.ndarray:
    """H0: counts per outcome index, int64[2^N]."""
    return np.bincount(pack_bits(samples), minlength=1 << num_qubits).astype(np.int64)


def pauli_matrix(label: str, reversed_kron: bool = True) -> np.ndarray:
    """RQC/reconstruct.py:5-24 (kron over the REVERSED label: qubit i <-> bit i) or
    SS/reconstruct.py:5-16 (``reversed_kron=False``: label[0] is the most significant factor)."""
    mats = {
        "I": np.array([[1, 0], [0, 1]], dtype=complex),
        "X": np.array([[0, 1], [1, 0]], dtype=complex),
        "Y": np.array([[0, -1j], [1j, 0]], dtype=complex),
        "Z": np.array([[1, 0], [0, -1]], dtype=complex),
    }
    lab = label[::-1] if reversed_kron else label
    m = mats[lab[0]]
    for ch in lab[1:]:
        m = np.kron(m, mats[ch])
    return m


def pauli_coefficient(pauli: str, data: dict) -> float:
    """RQC/reconstruct.py:26-46: 1.0 for the identity; else the FIRST basis (dict order) whose letters match
    on the support; mean over shots of prod(1-2*bit) on the support; 0.0 if none is compatible."""
    if all(c == "I" for c in pauli):
        return 1.0
    for key, samples in data.items():
        if all(p == "I" or p == b for p, b in zip(pauli, key)):
            support = [i for i, c in enumerate(pauli) if c != "I"]
            vals = 1 - 2 * np.asarray(samples)
            return float(np.mean(np.prod(vals[:, support], axis=1)))
    return 0.0


def make_psd(rho: np.ndarray) -> np.ndarray:
    """RQC/reconstruct.py:48-54: eigh, clip negatives, renormalise trace when positive, rebuild."""
    evals, evecs = np.linalg.eigh(rho)
    evals = np.maximum(evals, 0)
    if np.sum(evals) > 0:
        evals = evals / np.sum(evals)
    return (evecs * evals) @ evecs.conj().T


def linear_inversion_literal(data: dict, num_qubits: int, reversed_kron: bool = True, psd: bool = True) -> np.ndarray:
    """RQC/reconstruct.py:56-67: rho = 2^-N sum_P <P> P over all 4^N strings in product order, then PSD."""
    dim = 1 << num_qubits
    rho = np.zeros((dim, dim), dtype=complex)
    for p in product("IXYZ", repeat=num_qubits):
        s = "".join(p)
        rho += pauli_coefficient(s, data) * pauli_matrix(s, reversed_kron)
    rho /= dim
    return make_psd(rho) if psd else rho


def wht_inplace(a: np.ndarray) -> np.ndarray:
    """Unnormalised Walsh-Hadamard transform along the last axis (length 2^N), exact in int64."""
    a = a.copy()
    n = a.shape[-1]
    h = 1
    while h < n:
        a = a.reshape(a.shape[:-1] + (n // (2 * h), 2, h))
        lo, hi = a[..., 0, :].copy(), a[..., 1, :].copy()
        a[..., 0, :], a[..., 1, :] = lo + hi, lo - hi
        a = a.reshape(a.shape[:-3] + (n,))
        h *= 2
    return a


def linear_inversion_hist(hist: np.ndarray, num_qubits: int, reversed_kron: bool = True, psd: bool = True,
                          shots=None) -> np.ndarray:
    """Histogram restatement of R1+R2+R3 (equality with the literal loop is tested):
    W[b,:] = WHT(hist[b,:]); <P> = W[b*(P), mask(P)] / shots_b with b*(P) = P with I->X;
    rho[r, r^xm] += <P> * (-i)^{nY} * (-1)^{popcount(r & zm)} / 2^N.
    ``hist`` is int[3^N, 2^N] in product basis order; bit i of the outcome index = qubit i."""
    N = num_qubits
    dim = 1 << N
    hist = np.asarray(hist, dtype=np.int64)
    W = wht_inplace(hist)
    if shots is None:
        shots = hist.sum(axis=1)
    shots = np.asarray(shots, dtype=np.float64)
    rho = np.zeros((dim, dim), dtype=complex)
    r = np.arange(dim)
    pc = np.array([bin(v).count("1") for v in range(dim)])
    for p in product(range(4), repeat=N):          # 0=I 1=X 2=Y 3=Z, p[i] acts on qubit i
        b = 0
        mask = xm = zm = 0
        nY = 0
        for i, c in enumerate(p):
            b = b * 3 + (0 if c == 0 else c - 1)
            pos = i if reversed_kron else N - 1 - i   # bit position of this factor in the matrix index
            if c != 0:
                mask |= 1 << i                        # data parity always uses column i = qubit i
            if c in (1, 2):
                xm |= 1 << pos
            if c in (2, 3):
                zm |= 1 << pos
            if c == 2:
                nY += 1
        if mask == 0:
            coeff = 1.0
        else:
            coeff = W[b, mask] / shots[b] if shots[b] > 0 else 0.0
        # P[r, r^xm] = prod_i phase; for Y: <r_i|Y|r_i^1> = -i if r_i==0 (row 0) else +i
        # (-i)^{nY} * (-1)^{popcount(r & zm)} with r the ROW index
        phase = ((-1j) ** nY) * (1 - 2 * (pc[r & zm] & 1))
        rho[r, r ^ xm] += coeff * phase
    rho /= dim
    return make_psd(rho) if psd else rho


def _sqrtm_psd_svd(a: np.ndarray) -> np.ndarray:
    u, s, vh = np.linalg.svd(a)
    return (u * np.sqrt(s)) @ vh


def state_fidelity(target, rho: np.ndarray) -> float:
    """qiskit.quantum_info.state_fidelity (third-party, not under /root/reference; qiskit>=1.0.0 per
    RQC/requirements.txt:8).  Published definition (notes.pdf p.9): F = (Tr sqrt(sqrt(r1) r2 sqrt(r1)))^2 =
    ||sqrt(r1) sqrt(r2)||_*^2; for a pure argument it reduces to <psi|rho|psi>.  Call sites:
    RQC/evaluate.py:77,87; SS/main.py:127; NB c9:93."""
    target = np.asarray(target)
    rho = np.asarray(rho)
    if target.ndim == 1 and rho.ndim == 1:
        return float(abs(np.vdot(target, rho)) ** 2)
    if target.ndim == 1:
        return float(np.real(np.vdot(target, rho @ target)))
    if rho.ndim == 1:
        return float(np.real(np.vdot(rho, target @ rho)))
    s1 = _sqrtm_psd_svd(target)
    s2 = _sqrtm_psd_svd(rho)
    return float(np.linalg.norm(s1 @ s2, ord="nuc") ** 2)


def entropy_bits(rho: np.ndarray) -> float:
    """qiskit ``entropy(state, base=2)``: -sum lambda log2 lambda over positive eigenvalues."""
    ev = np.linalg.eigvalsh(rho)
    ev = ev[ev > 0]
    return float(-np.sum(ev * np.log2(ev)))


def partial_trace_high(rho: np.ndarray, num_qubits: int, cut: int) -> np.ndarray:
    """Trace out qubits cut..N-1 (the most significant index bits in qiskit's little-endian convention)."""
    lo, hi = 1 << cut, 1 << (num_qubits - cut)
    return np.einsum("aiaj->ij", rho.reshape(hi, lo, hi, lo))


def get_metrics(rho: np.ndarray, num_qubits: int):
    """RQC/reconstruct.py:69-76: purity Tr rho^2, von Neumann entropy, half-cut entanglement entropy."""
    purity = float(np.real(np.trace(rho @ rho)))
    cut = num_qubits // 2
    return purity, entropy_bits(rho), entropy_bits(partial_trace_high(rho, num_qubits, cut))


def rho_from_single_qubit_counts(cx: dict, cy: dict, cz: dict) -> np.ndarray:
    """NB c9:43-69: rho = (I + <X>X + <Y>Y + <Z>Z)/2 with <P> = p0 - p1."""
    def ev(c):
        tot = sum(c.values())
        return 0.0 if tot == 0 else c.get("0", 0) / tot - c.get("1", 0) / tot
    return 0.5 * (pauli_matrix("I") + ev(cx) * pauli_matrix("X") + ev(cy) * pauli_matrix("Y") + ev(cz) * pauli_matrix("Z"))


# ----------------------------------------------------------------------------
# 7. Synthetic "random-circuit" states and Born histograms (SURVEY 8d inputs)
# ----------------------------------------------------------------------------
def haar_state(num_qubits: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << num_qubits) + 1j * rng.normal(size=1 << num_qubits)
    return v / np.linalg.norm(v)


def born_probabilities_all(psi: np.ndarray, num_qubits: int) -> np.ndarray:
    """Outcome distributions of a pure state in ALL 3^N Pauli bases (product order, letter 0 slowest) -> [3^N, 2^N].
    Same physics as ``born_probabilities`` but expands one qubit at a time (3^N * 2^N * N work instead of 3^N * 4^N)."""
    H = np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
    Sdg = np.array([[1, 0], [0, -1j]], dtype=complex)
    rots = [H, H @ Sdg, np.eye(2, dtype=complex)]
    dim = 1 << num_qubits
    states = np.asarray(psi, dtype=complex).reshape(1, dim)
    for i in range(num_qubits):
        lo = 1 << i
        v = states.reshape(states.shape[0], dim // (2 * lo), 2, lo)
        states = np.stack([np.einsum("ab,nhbl->nhal", r, v) for r in rots], axis=1).reshape(-1, dim)
    p = np.abs(states) ** 2
    return p / p.sum(axis=1, keepdims=True)


def born_probabilities(psi_or_rho: np.ndarray, num_qubits: int, basis: str) -> np.ndarray:
    """Outcome distribution when qubit i is rotated by H (X) or H.Sdg (Y) before a Z measurement
    (RQC/build_dataset.py:94-96); outcome index bit i = qubit i."""
    H = np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
    Sdg = np.array([[1, 0], [0, -1j]], dtype=complex)
    rot = {"X": H, "Y": H @ Sdg, "Z": np.eye(2, dtype=complex)}
    U = np.array([[1.0 + 0j]])
    for i in range(num_qubits):          # qubit i is bit i -> later qubits are more significant kron factors
        U = np.kron(rot[basis[i]], U)
    a = np.asarray(psi_or_rho)
    if a.ndim == 1:
        p = np.abs(U @ a) ** 2
    else:
        p = np.real(np.diag(U @ a @ U.conj().T))
    p = np.maximum(p, 0)
    return p / p.sum()


# ============================================================================================ dataset unrolling
# RQC/dataset.py:47-65 (SS/dataset.py:14-33): every measurement record's counts dict becomes `count` copies of
# (bits[::-1], basis_idx); a DataLoader(shuffle=True) then draws batches (RQC/main.py:84-92).  The restatement keeps
# the counts as a table and addresses shots by their position in the canonical unrolled order.
def counts_rows_from_records(records: list, num_qubits: int):
    """-> (hist int64[n_rows, 2^N], row_basis int64[n_rows], key_order list[list[int]]): one row per measurement
    record in the reference's iteration order (circuit, then measurement); outcome index s = sum_q bit_q << q with
    bits = reversed counts key (RQC/dataset.py:59); unknown bases are skipped (RQC/dataset.py:54);
    key_order[row] = outcome indices in the counts dict's own order (the reference's within-row order)."""
    names = basis_strings(num_qubits)
    b2i = {b: i for i, b in enumerate(names)}
    hist, row_basis, order = [], [], []
    for circ in records:
        for meas in circ.get("measurements", []):
            if meas["basis"] not in b2i:
                continue
            row = np.zeros(1 << num_qubits, dtype=np.int64)
            ko = []
            for key, cnt in meas["counts"].items():
                bits = [int(c) for c in key][::-1]
                s = sum(v << i for i, v in enumerate(bits))
                row[s] += int(cnt)
                ko.append(s)
            hist.append(row)
            row_basis.append(b2i[meas["basis"]])
            order.append(ko)
    return np.array(hist, dtype=np.int64).reshape(-1, 1 << num_qubits), np.array(row_basis, dtype=np.int64), order


def counts_unroll(hist: np.ndarray, row_basis: np.ndarray, num_qubits: int):
    """Canonical unroll (row-major, outcomes ascending inside a row) -> (bits int64[total, N], basis int64[total])."""
    flat = hist.reshape(-1)
    idx = np.repeat(np.arange(flat.size), flat)
    row, s = idx // hist.shape[1], idx % hist.shape[1]
    bits = ((s[:, None] >> np.arange(num_qubits)) & 1).astype(np.int64)
    return bits, np.asarray(row_basis)[row].astype(np.int64)


def _fmix32(x):
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13); x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def feistel_perm(i, total: int, seed: int, epoch: int) -> np.ndarray:
    """Keyed bijection of [0, total): 4-round balanced Feistel network on 2h bits (4^h >= total) with cycle walking;
    the stand-in for one epoch of DataLoader(shuffle=True) (include/ddqst.h, ddqst_counts_gather)."""
    h = 1
    while h < 31 and (1 << (2 * h)) < total:
        h += 1
    mask = np.uint64((1 << h) - 1)
    base = (np.uint64(seed & 0xFFFFFFFF) ^ _fmix32(((seed >> 32) + 0x9E3779B9 * ((epoch + 1) & 0xFFFFFFFF)) & 0xFFFFFFFF))
    keys = [_fmix32((int(base) + r * 0x85EBCA6B) & 0xFFFFFFFF) for r in range(4)]
    x = np.asarray(i, dtype=np.uint64).copy().reshape(-1)
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        L, R = x[todo] >> np.uint64(h), x[todo] & mask
        for r in range(4):
            L, R = R, L ^ (_fmix32(R ^ keys[r]) & mask)
        x[todo] = (L << np.uint64(h)) | R
        todo = x >= np.uint64(total)
    return x.astype(np.int64)


def counts_batch(hist: np.ndarray, row_basis: np.ndarray, num_qubits: int, start: int, count: int, seed: int, epoch: int,
                 permute: bool = True):
    """Positions start..start+count-1 (mod total) of the (shuffled) unrolled dataset -> (packed int64[count], basis)."""
    flat = hist.reshape(-1)
    total = int(flat.sum())
    p = (start + np.arange(count, dtype=np.int64)) % total
    if permute:
        p = feistel_perm(p, total, seed, epoch)
    csum = np.cumsum(flat)
    idx = np.searchsorted(csum, p, side="right")
    return (idx % hist.shape[1]).astype(np.int64), np.asarray(row_basis)[idx // hist.shape[1]].astype(np.int64)


# ============================================================================================ synthetic data (8f-4)
# CPU restatement of csrc/synth.cu: special / random-circuit states and Born-rule sampling in Pauli bases.  The reference
# generates this data with Qiskit Aer (SS/data_gen.py:13-63, AS/data_gen.py:59-140), which is not installable here; the
# physics restated is its circuit: state preparation, then per qubit H (X) or Sdg.H (Y) before a Z measurement.
SITE_SYNTH_DRAW, SITE_SYNTH_GATE = 16, 17


def synth_state(num_qubits: int, kind: str, depth: int = 0, seed: int = 0) -> np.ndarray:
    """'zero' | 'plus' | 'ghz' (= 'bell', SS/data_gen.py:22-26) | 'rqc' (brick-wall: random U3 per qubit, then CZ on
    alternating neighbour pairs, per layer; angles from Philox counter (layer, qubit, SITE_SYNTH_GATE<<16, 0))."""
    dim = 1 << num_qubits
    psi = np.zeros(dim, dtype=complex)
    psi[0] = 1.0
    r = 0.70710678118654752440

    def apply(psi, q, m):
        v = psi.reshape(dim >> (q + 1), 2, 1 << q)
        return np.einsum("ab,hbl->hal", m, v).reshape(dim)

    if kind == "plus":
        for q in range(num_qubits):
            psi = apply(psi, q, np.array([[r, r], [r, -r]], dtype=complex))
    elif kind in ("ghz", "bell"):
        psi[:] = 0
        psi[0] = psi[dim - 1] = r
    elif kind == "rqc":
        s = np.arange(dim)
        for layer in range(depth):
            for q in range(num_qubits):
                w = philox4x32_10(np.array([[layer, q, SITE_SYNTH_GATE << 16, 0]], dtype=np.uint32),
                                  np.array([[seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]], dtype=np.uint32))[0]
                theta = math.acos(1.0 - 2.0 * (float(w[0]) / 4294967296.0))
                phi, lam = 2 * math.pi * (float(w[1]) / 4294967296.0), 2 * math.pi * (float(w[2]) / 4294967296.0)
                c, sn = math.cos(0.5 * theta), math.sin(0.5 * theta)
                U = np.array([[c, -complex(math.cos(lam), math.sin(lam)) * sn],
                              [complex(math.cos(phi), math.sin(phi)) * sn, complex(math.cos(phi + lam), math.sin(phi + lam)) * c]])
                psi = apply(psi, q, U)
            sign = np.zeros(dim, dtype=np.int64)
            for q in range(layer & 1, num_qubits - 1, 2):
                sign ^= (s >> q) & (s >> (q + 1)) & 1
            psi = np.where(sign == 1, -psi, psi)
    elif kind != "zero":
        raise ValueError(kind)
    return psi


def synth_probabilities(psi: np.ndarray, num_qubits: int, basis_ids, p_depol: float = 0.0, p_readout: float = 0.0) -> np.ndarray:
    """Sampled distributions [len(basis_ids), 2^N] (unnormalised exactly as the kernel holds them)."""
    names = basis_strings(num_qubits)
    dim = 1 << num_qubits
    r = 0.70710678118654752440
    rot = {"X": np.array([[r, r], [r, -r]], dtype=complex), "Y": np.array([[r, -1j * r], [r, 1j * r]], dtype=complex)}
    out = np.zeros((len(basis_ids), dim))
    for k, b in enumerate(basis_ids):
        amp = np.asarray(psi, dtype=complex).copy()
        for q in range(num_qubits - 1, -1, -1):
            letter = names[b][q]
            if letter in rot:
                amp = np.einsum("ab,hbl->hal", rot[letter], amp.reshape(dim >> (q + 1), 2, 1 << q)).reshape(dim)
        p = (1.0 - p_depol) * (amp.real ** 2 + amp.imag ** 2) + p_depol / dim
        if p_readout > 0:
            for q in range(num_qubits):
                v = p.reshape(dim >> (q + 1), 2, 1 << q)
                p = np.stack([(1 - p_readout) * v[:, 0] + p_readout * v[:, 1], p_readout * v[:, 0] + (1 - p_readout) * v[:, 1]],
                             axis=1).reshape(dim)
        out[k] = p
    return out


def synth_histograms(probs: np.ndarray, basis_ids, shots: int, seed: int) -> np.ndarray:
    """Inverse-CDF draws from the Philox stream: draw j of a basis uses pair k = j // 2, counter (k_lo, basis,
    SITE_SYNTH_DRAW<<16, k_hi); u = 53 bits of (x,y) for even j, (z,w) for odd j; outcome = first s with cdf[s] > u*cdf[-1]."""
    n_b, dim = probs.shape
    hist = np.zeros((n_b, dim), dtype=np.int64)
    npairs = (shots + 1) // 2
    k = np.arange(npairs, dtype=np.uint64)
    for i, b in enumerate(basis_ids):
        cdf = np.cumsum(probs[i])
        ctr = np.stack([(k & np.uint64(0xFFFFFFFF)).astype(np.uint32), np.full(npairs, b, dtype=np.uint32),
                        np.full(npairs, SITE_SYNTH_DRAW << 16, dtype=np.uint32), (k >> np.uint64(32)).astype(np.uint32)], axis=1)
        key = np.tile(np.array([[seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]], dtype=np.uint32), (npairs, 1))
        w = philox4x32_10(ctr, key).astype(np.uint64)
        u0 = (((w[:, 0] << np.uint64(32)) | w[:, 1]) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        u1 = (((w[:, 2] << np.uint64(32)) | w[:, 3]) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        u = np.stack([u0, u1], axis=1).reshape(-1)[:shots] * cdf[-1]
        idx = np.minimum(np.searchsorted(cdf, u, side="right"), dim - 1)
        hist[i] = np.bincount(idx, minlength=dim)
    return hist
