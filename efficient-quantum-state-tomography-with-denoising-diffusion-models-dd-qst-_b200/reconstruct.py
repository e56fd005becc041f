"""Reconstruction and metrics: the reference's ``reconstruct.py`` surface on top of libddqst.

``linear_inversion(synthetic_data, num_qubits)`` accepts what the reference accepts -- a dict
``basis_str -> int ndarray[shots, N]`` in insertion order (RQC/reconstruct.py:56-67) -- and, natively, a
per-basis histogram tensor ``uint32[3^N, 2^N]`` as produced by ``DiscreteDiffusion.sample``.  The result is a
``DensityMatrix`` holding ``.data`` (complex128 ndarray) like qiskit's, so ``state_fidelity(target, rho)``,
``get_metrics(rho, n)`` read as in RQC/evaluate.py:75-88.  ``convention="unreversed"`` selects the Kronecker
order of SS/reconstruct.py:5-16.
"""
from __future__ import annotations

import ctypes as C
from itertools import product

import numpy as np
import torch

from . import _lib
from .model import pack_bits


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("reconstruction has no CPU path: a B200 (cuda) device is required")
    return torch.device("cuda", torch.cuda.current_device())


class DensityMatrix:
    """Minimal stand-in for qiskit.quantum_info.DensityMatrix: ``.data`` complex128[2^N, 2^N]; the device copy
    is kept so chained native calls do not round-trip through the host."""

    def __init__(self, data):
        self.pure_vector = None                 # set when built from a state vector: lets fidelity skip the mixed-state formula
        if isinstance(data, DensityMatrix):
            self.pure_vector = data.pure_vector
            data = data.tensor if data.tensor is not None else data.data
        if torch.is_tensor(data):
            t = data.to(torch.complex128)
            if t.dim() == 1:
                self.pure_vector = t.contiguous()
                t = torch.outer(t, t.conj())
            self.tensor = t.contiguous()
            self._host = None
        else:
            a = np.asarray(getattr(data, "data", data), dtype=complex)
            if a.ndim == 1:
                self.pure_vector = a.copy()
                a = np.outer(a, a.conj())
            self._host = np.ascontiguousarray(a)
            self.tensor = None

    @property
    def data(self) -> np.ndarray:
        if self._host is None:
            self._host = self.tensor.cpu().numpy()
        return self._host

    def device_tensor(self) -> torch.Tensor:
        if self.tensor is None:
            self.tensor = torch.from_numpy(self.data).to(_dev())
        return self.tensor

    @property
    def dim(self) -> int:
        return (self.tensor if self.tensor is not None else self._host).shape[0]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class Statevector:
    def __init__(self, data):
        self.data = np.asarray(getattr(data, "data", data), dtype=complex).reshape(-1)


def basis_strings(num_qubits: int):
    """Product order X<Y<Z, letter 0 slowest (RQC/dataset.py:43)."""
    return ["".join(p) for p in product("XYZ", repeat=num_qubits)]


def get_pauli_matrix(label: str, convention: str = "reversed") -> np.ndarray:
    """Dense Pauli string (RQC/reconstruct.py:5-24; ``convention='unreversed'``: SS/reconstruct.py:5-16).
    Host-side utility kept for API parity; linear_inversion never materialises these."""
    mats = {"I": np.eye(2, dtype=complex), "X": np.array([[0, 1], [1, 0]], dtype=complex),
            "Y": np.array([[0, -1j], [1j, 0]], dtype=complex), "Z": np.array([[1, 0], [0, -1]], dtype=complex)}
    lab = label[::-1] if convention == "reversed" else label
    out = mats[lab[0]]
    for ch in lab[1:]:
        out = np.kron(out, mats[ch])
    return out


def histogram_samples(samples, num_qubits: int) -> torch.Tensor:
    """int[shots, N] (ndarray or tensor, column q = qubit q) -> uint32[2^N] counts on the device."""
    lib = _lib.load()
    dev = _dev()
    x = samples if torch.is_tensor(samples) else torch.from_numpy(np.ascontiguousarray(samples).astype(np.int64))
    x = x.to(dev)
    hist = torch.zeros(1 << num_qubits, dtype=torch.uint32, device=dev)
    if x.shape[0] == 0:
        return hist
    packed = pack_bits(x, num_qubits)
    _lib.check(lib.ddqst_histogram(_lib.ptr(packed), 2, packed.shape[0], num_qubits, _lib.ptr(hist), _lib.stream_ptr()))
    return hist


def _compatible_slot_table(keys, num_qubits: int) -> np.ndarray:
    """sel[4^N]: for each Pauli string (product order over I<X<Y<Z) the index of the FIRST key (dict order) that
    matches it on its support (RQC/reconstruct.py:32-38); -1 if none, -2 for the identity."""
    N = num_qubits
    code = {"X": 1, "Y": 2, "Z": 3}
    K = np.array([[code[c] for c in k] for k in keys], dtype=np.int8).reshape(len(keys), N)
    n_p = 4 ** N
    sel = np.full(n_p, -1, dtype=np.int32)
    chunk = max(1, (1 << 22) // max(1, len(keys) * N))
    for p0 in range(0, n_p, chunk):
        idx = np.arange(p0, min(n_p, p0 + chunk))
        P = np.stack([(idx // 4 ** (N - 1 - i)) % 4 for i in range(N)], axis=1).astype(np.int8)      # [c, N]
        ok = ((P[:, None, :] == 0) | (P[:, None, :] == K[None, :, :])).all(axis=2) if len(keys) else np.zeros((len(idx), 0), bool)
        first = np.where(ok.any(axis=1), ok.argmax(axis=1), -1) if len(keys) else np.full(len(idx), -1)
        sel[idx] = first
    sel[0] = -2
    return sel


def _prepare(synthetic_data, num_qubits: int):
    """-> (hist uint32[n_slots, 2^N], shots (None: the kernel takes each row's sum, which is the shot count of that
    basis in every input form), sel int32[4^N] or None) on the device."""
    dev = _dev()
    N = num_qubits
    if isinstance(synthetic_data, dict):
        keys = list(synthetic_data.keys())
        hist = torch.zeros(max(len(keys), 1), 1 << N, dtype=torch.uint32, device=dev)
        for i, k in enumerate(keys):
            s = synthetic_data[k]
            if getattr(s, "ndim", 2) == 1:                 # already a counts row uint32/int[2^N] (evaluate.format_raw_counts_for_inversion)
                row = s if torch.is_tensor(s) else torch.from_numpy(np.ascontiguousarray(s))
                row = row.to(dev)
                row = row.view(torch.int32) if row.dtype == torch.uint32 else row.to(torch.int32)
                hist[i] = row.view(torch.uint32)
            else:
                hist[i] = histogram_samples(s, N)
        canonical = keys == basis_strings(N)
        sel = None if canonical else torch.from_numpy(_compatible_slot_table(keys, N)).to(dev)
        return hist[:len(keys)].contiguous(), None, sel
    hist = synthetic_data if torch.is_tensor(synthetic_data) else torch.from_numpy(np.ascontiguousarray(synthetic_data))
    hist = hist.to(dev)
    if hist.dtype != torch.uint32:
        hist = hist.to(torch.int32).view(torch.uint32)
    return hist.contiguous(), None, None


def linear_inversion_raw(synthetic_data, num_qubits: int, convention: str = "reversed") -> torch.Tensor:
    """rho = 2^-N sum_P <P> P before the PSD step (RQC/reconstruct.py:56-66) as complex128[2^N,2^N] on the device."""
    lib = _lib.load()
    hist, shots, sel = _prepare(synthetic_data, num_qubits)
    dev = hist.device
    dim = 1 << num_qubits
    rho = torch.empty(dim, dim, dtype=torch.complex128, device=dev)
    n_slots = hist.shape[0]
    ws = _lib.workspace.get(max(n_slots * dim * 4, 8 * dim * dim) + 256, dev)
    kron = _lib.KRON_REVERSED if convention == "reversed" else _lib.KRON_UNREVERSED
    _lib.check(lib.ddqst_linear_inversion(_lib.ptr(hist) if n_slots else None, _lib.ptr(shots), n_slots,
                                          num_qubits, _lib.ptr(sel), kron, _lib.ptr(rho), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr()))
    return rho


def make_positive_semidefinite(rho) -> DensityMatrix:
    """eigh, clip negatives, renormalise, rebuild (RQC/reconstruct.py:48-54), on the device."""
    lib = _lib.load()
    t = DensityMatrix(rho).device_tensor().clone()
    dim = t.shape[0]
    if dim == 1:
        return DensityMatrix(t)
    nbytes = 2 * 16 * dim * dim + 8 * dim + 1024 + 24 * dim * dim + 4096      # + 24 n^2: lets the eigensolver start its sweeps in fp32
    ws = _lib.workspace.get(nbytes, t.device)
    _lib.check(lib.ddqst_psd_project(_lib.ptr(t), dim, None, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return DensityMatrix(t)


def linear_inversion(synthetic_data, num_qubits: int, convention: str = "reversed") -> DensityMatrix:
    """RQC/reconstruct.py:56-67: linear inversion over all 4^N Pauli strings followed by the PSD projection."""
    return make_positive_semidefinite(linear_inversion_raw(synthetic_data, num_qubits, convention))


def get_coefficient(pauli_str: str, synthetic_data: dict) -> float:
    """<P> from the first compatible basis in dict order (RQC/reconstruct.py:26-46)."""
    if all(c == "I" for c in pauli_str):
        return 1.0
    N = len(pauli_str)
    for key, samples in synthetic_data.items():
        if all(p == "I" or p == b for p, b in zip(pauli_str, key)):
            if samples.shape[0] == 0:
                return float("nan")
            h = histogram_samples(samples, N).view(torch.int32).to(torch.int64).cpu().numpy()
            mask = sum(1 << i for i, c in enumerate(pauli_str) if c != "I")
            par = np.array([bin(s & mask).count("1") & 1 for s in range(1 << N)])
            return float((h * (1 - 2 * par)).sum() / h.sum())
    return 0.0


def state_fidelity(target, rho) -> float:
    """qiskit.quantum_info.state_fidelity as used at RQC/evaluate.py:77,87 / SS/main.py:127: <psi|rho|psi> for a
    state-vector argument, (Tr sqrt(sqrt(a) b sqrt(a)))^2 for two density matrices."""
    lib = _lib.load()
    dev = _dev()

    def as_dev(x):
        if isinstance(x, DensityMatrix):
            return x.device_tensor()
        if torch.is_tensor(x):
            return x.to(dev).to(torch.complex128).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(getattr(x, "data", x), dtype=complex))).to(dev)

    def is_vec(x):
        return isinstance(x, Statevector) or (not isinstance(x, DensityMatrix) and np.ndim(getattr(x, "data", x)) == 1)

    # two density matrices: one that is |psi><psi| (built from a vector, or Tr sigma^2 = 1 to rounding -- RQC/evaluate.py:71 wraps the
    # clean state vector) takes the pure-target formula instead of two eigensolves.  The purity test costs a device sync, so it is
    # only made when neither argument is a vector already.
    if not is_vec(target) and not is_vec(rho):
        for first, second in ((target, rho), (rho, target)):
            t, kind = _target_kind(first, dev)
            if kind == _lib.TARGET_STATEVECTOR:
                return state_fidelity(Statevector(t.cpu().numpy()), second)
            if kind == _lib.TARGET_RANK_ONE:
                return float(torch.sum(t * as_dev(second).transpose(0, 1)).real.item())       # Tr(sigma rho)
    a, b = as_dev(target), as_dev(rho)
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    if a.dim() == 1 and b.dim() == 1:
        return float(abs(torch.vdot(a, b).item()) ** 2)
    if a.dim() == 1 or b.dim() == 1:
        psi, mat = (a, b) if a.dim() == 1 else (b, a)
        _lib.check(lib.ddqst_fidelity_pure(_lib.ptr(psi), _lib.ptr(mat), mat.shape[0], _lib.ptr(out), _lib.stream_ptr()))
        return float(out.item())
    dim = a.shape[0]
    ws = _lib.workspace.get(5 * 16 * dim * dim + 8 * dim + 1024 + 24 * dim * dim + 4096, dev)
    _lib.check(lib.ddqst_fidelity_mixed(_lib.ptr(a), _lib.ptr(b), dim, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return float(out.item())


def get_metrics(rho, num_qubits: int):
    """(purity, von Neumann entropy, half-cut entanglement entropy) as RQC/reconstruct.py:69-76."""
    lib = _lib.load()
    t = DensityMatrix(rho).device_tensor()
    dim = t.shape[0]
    out = torch.zeros(3, dtype=torch.float64, device=t.device)
    ws = _lib.workspace.get(3 * 16 * dim * dim + 8 * dim + 1024 + 24 * dim * dim + 4096, t.device)
    _lib.check(lib.ddqst_metrics(_lib.ptr(t), num_qubits, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    p, s, e = out.cpu().tolist()
    return p, s, e


class ReconReport:
    """What the evaluation loop (RQC/evaluate.py:75-88) needs of one reconstruction: ``rho`` (PSD-projected
    ``DensityMatrix``), ``evals`` (clipped, renormalised spectrum, device tensor), ``fidelity`` (None without a target),
    ``purity``, ``entropy`` (von Neumann, bits), ``entanglement_entropy`` (low N/2 qubits, bits)."""

    def __init__(self, rho, evals, fidelity, purity, entropy, entanglement_entropy):
        self.rho, self.evals, self.fidelity = rho, evals, fidelity
        self.purity, self.entropy, self.entanglement_entropy = purity, entropy, entanglement_entropy

    def metrics(self):
        """The tuple ``get_metrics(rho, n)`` returns (RQC/reconstruct.py:69-76)."""
        return self.purity, self.entropy, self.entanglement_entropy


def _target_kind(target, dev):
    """-> (device tensor or None, DDQST_TARGET_*).  A DensityMatrix that is |psi><psi| to rounding (Tr sigma^2 = 1, as the
    clean-state targets of RQC/evaluate.py:71) is flagged rank one: its Uhlmann fidelity is Tr(sigma rho) exactly."""
    if target is None:
        return None, _lib.TARGET_NONE
    if isinstance(target, Statevector) or (not isinstance(target, DensityMatrix) and np.ndim(getattr(target, "data", target)) == 1):
        v = target if torch.is_tensor(target) else torch.from_numpy(np.ascontiguousarray(np.asarray(getattr(target, "data", target), dtype=complex)))
        return v.to(dev).to(torch.complex128).contiguous(), _lib.TARGET_STATEVECTOR
    dm = target if isinstance(target, DensityMatrix) else DensityMatrix(target)
    if dm.pure_vector is not None:
        v = dm.pure_vector if torch.is_tensor(dm.pure_vector) else torch.from_numpy(np.ascontiguousarray(dm.pure_vector))
        return v.to(dev).to(torch.complex128).contiguous(), _lib.TARGET_STATEVECTOR
    t = dm.device_tensor()
    purity = float(torch.sum(t.real ** 2 + t.imag ** 2).item())          # Tr(sigma^2) of a Hermitian sigma
    trace = float(torch.diagonal(t).real.sum().item())
    return t, (_lib.TARGET_RANK_ONE if abs(purity - 1.0) < 1e-10 and abs(trace - 1.0) < 1e-10 else _lib.TARGET_MIXED)


def recon_report(data, num_qubits: int, target=None, convention: str = "reversed") -> ReconReport:
    """``linear_inversion`` + ``state_fidelity(target, rho)`` + ``get_metrics(rho, n)`` (RQC/evaluate.py:75-78) with ONE full
    eigendecomposition instead of five (ddqst_recon_report).  ``data``: whatever ``linear_inversion`` accepts, or an already
    assembled raw rho (complex tensor / DensityMatrix [2^N, 2^N])."""
    lib = _lib.load()
    dim = 1 << num_qubits
    if isinstance(data, DensityMatrix):
        rho = data.device_tensor().clone()
    elif torch.is_tensor(data) and data.is_complex():
        rho = data.to(_dev()).to(torch.complex128).contiguous().clone()
    else:
        rho = linear_inversion_raw(data, num_qubits, convention)
    dev = rho.device
    tgt, kind = _target_kind(target, dev)
    evals = torch.empty(dim, dtype=torch.float64, device=dev)
    report = torch.empty(5, dtype=torch.float64, device=dev)
    nbytes = (5 if kind == _lib.TARGET_MIXED else 3) * 16 * dim * dim + 16 * dim + 2048 + 24 * dim * dim     # + 24 n^2: fp32 start of the sweeps
    ws = _lib.workspace.get(nbytes, dev)
    _lib.check(lib.ddqst_recon_report(_lib.ptr(rho), num_qubits, _lib.ptr(tgt), kind, _lib.ptr(evals), _lib.ptr(report),
                                      _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    f, p, s, e, _ = report.cpu().tolist()
    return ReconReport(DensityMatrix(rho), evals, None if kind == _lib.TARGET_NONE else f, p, s, e)
