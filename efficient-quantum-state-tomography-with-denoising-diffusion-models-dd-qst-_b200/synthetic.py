"""Synthetic measurement data on the GPU: the stand-in for the reference's Qiskit-Aer data generation
(``generate_synthetic_data`` of SS/data_gen.py:40-63, the noisy variants of AS/data_gen.py:29-57 and the random-circuit
records of RQC/batch_build_dataset.py:53-144) at sizes a CPU simulator cannot reach (3^8 bases x 10^6 shots).

``generate_synthetic_data(num_qubits, state_type, shots)`` keeps the reference's return convention
(list of ``{basis_str, basis_idx, counts}`` + the basis list) when ``as_counts=True``; by default the data stays on the
device as ``uint32[3^N, 2^N]`` -- what ``QuantumStateDataset``/``linear_inversion`` consume without a round trip.
"""
from __future__ import annotations

from itertools import product

import numpy as np
import torch

from . import _lib

_KINDS = {"zero": 0, "plus": 1, "ghz": 2, "bell": 2, "rqc": 3, "random": 3}


def get_basis_combinations(num_qubits: int):
    """All 3^N Pauli basis strings in product order (SS/data_gen.py:9-12)."""
    return ["".join(p) for p in product("XYZ", repeat=num_qubits)]


def synth_state(num_qubits: int, state_type: str = "rqc", depth: int = 8, seed: int = 0, device="cuda") -> torch.Tensor:
    """complex128[2^N] on the device; index bit i = qubit i (Qiskit little endian)."""
    if state_type not in _KINDS:
        raise ValueError(f"state_type must be one of {sorted(_KINDS)}")
    lib = _lib.load()
    psi = torch.empty(1 << num_qubits, dtype=torch.complex128, device=device)
    _lib.check(lib.ddqst_synth_state(num_qubits, _KINDS[state_type], int(depth), int(seed), _lib.ptr(psi), _lib.stream_ptr()))
    return psi


def born_histograms(psi: torch.Tensor, num_qubits: int, shots: int, seed: int = 0, bases=None, noise_type: str = "ideal",
                    error_rate: float = 0.01, return_probs: bool = False):
    """Measure ``psi`` ``shots`` times in each basis (all 3^N when ``bases`` is None) -> uint32[n_bases, 2^N].
    noise_type: 'ideal', 'readout' (every measured bit flips with ``error_rate``, AS/data_gen.py:42-46) or
    'depolarizing' (global mix (1-p) rho + p I/2^N, the C5 mixed-state config)."""
    if noise_type not in ("ideal", "readout", "depolarizing"):
        raise ValueError("noise_type must be 'ideal', 'readout' or 'depolarizing' (gate-level Aer noise models are out of scope)")
    lib = _lib.load()
    dev = psi.device
    n_b = 3 ** num_qubits if bases is None else len(bases)
    ids = None if bases is None else torch.tensor(list(bases), dtype=torch.int32, device=dev)
    hist = torch.zeros(n_b, 1 << num_qubits, dtype=torch.uint32, device=dev)
    probs = torch.empty(n_b, 1 << num_qubits, dtype=torch.float64, device=dev) if return_probs else None
    pd = error_rate if noise_type == "depolarizing" else 0.0
    pr = error_rate if noise_type == "readout" else 0.0
    psi = psi.to(torch.complex128).contiguous()
    _lib.check(lib.ddqst_synth_born_histograms(_lib.ptr(psi), num_qubits, _lib.ptr(ids), n_b, int(shots), int(seed), pd, pr,
                                               _lib.ptr(hist), _lib.ptr(probs), _lib.stream_ptr()))
    return (hist, probs) if return_probs else hist


def counts_records(hist: torch.Tensor, num_qubits: int, bases=None):
    """uint32[n_bases, 2^N] -> the reference's list of {basis_str, basis_idx, counts} (counts keys 'q_{N-1}..q_0')."""
    names = get_basis_combinations(num_qubits)
    ids = list(range(len(names))) if bases is None else list(bases)
    h = hist.view(torch.int32).cpu().numpy()
    out = []
    for row, b in zip(h, ids):
        counts = {format(s, f"0{num_qubits}b"): int(c) for s, c in enumerate(row) if c}
        out.append({"basis_str": names[b], "basis": names[b], "basis_idx": b, "counts": counts})
    return out


def generate_synthetic_data(num_qubits: int, state_type: str, shots_train: int, noise_type: str = "ideal", rqc_depth: int = 8,
                            seed: int = 0, error_rate: float = 0.01, device="cuda", as_counts: bool = False):
    """SS/data_gen.py:40-63 / AS/data_gen.py:190-250 on the device.  -> (data, basis_list, target_state): ``data`` is
    uint32[3^N, 2^N] on the device, or the reference's list of dicts when ``as_counts``; ``target_state`` is complex128[2^N]."""
    psi = synth_state(num_qubits, state_type, rqc_depth, seed, device)
    hist = born_histograms(psi, num_qubits, shots_train, seed, None, noise_type, error_rate)
    basis_list = get_basis_combinations(num_qubits)
    return (counts_records(hist, num_qubits) if as_counts else hist), basis_list, psi
