"""QuantumStateDataset: the reference's dataset surface (RQC/dataset.py:7-78, SS/dataset.py:5-40) with the unrolling
done on the device.

The reference expands every measurement record's counts dict into ``count`` copies of ``(bits[::-1], basis_idx)`` in a
Python list (10 M tuples for the shipped N=3 data) and hands it to ``DataLoader(shuffle=True)``.  Here the counts stay a
table ``hist[n_rows, 2^N]`` on the GPU; ``len()``, ``[idx]`` and ``data_tensor`` / ``basis_tensor`` keep their meaning,
and ``batch(step, batch_size)`` produces training batches directly in the packed layout the train step consumes
(ddqst_counts_scan / ddqst_counts_gather, include/ddqst.h).  Order: the reference's within-record order is the counts
dict's own key order; the canonical order used here is ascending outcome index -- equal as multisets per record.
"""
from __future__ import annotations

import glob
import io
import os
import pickle
import zipfile
from itertools import product

import numpy as np
import torch

from . import _lib


class _QiskitStub:
    """Stand-in for the three qiskit classes the shipped Datapoints/*.pt reference (Statevector, OpShape, Counts)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)


class _Counts(dict):
    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("qiskit"):
            return _Counts if name == "Counts" else type(name, (_QiskitStub,), {})
        return super().find_class(module, name)


def load_circuit_records(path: str) -> list:
    """Decode one ``.pt`` shard of circuit records (torch zip-pickle).  With qiskit installed ``torch.load`` is used, as
    the reference does (RQC/evaluate.py:51); without it the qiskit classes inside are replaced by inert stand-ins."""
    try:
        import qiskit  # noqa: F401
        return torch.load(path, weights_only=False)
    except ImportError:
        pass
    with zipfile.ZipFile(path) as z:
        pkl = [n for n in z.namelist() if n.endswith("data.pkl")][0]
        prefix = pkl[: -len("data.pkl")]

        def persistent_load(pid):
            _, storage_type, key, _, _numel = pid
            dtype = getattr(storage_type, "dtype", None) or torch.uint8
            return torch.frombuffer(bytearray(z.read(f"{prefix}data/{key}")), dtype=dtype).untyped_storage()

        up = _Unpickler(io.BytesIO(z.read(pkl)))
        up.persistent_load = persistent_load
        return up.load()


def state_vector_of(record: dict) -> np.ndarray:
    """``clean_state_vec`` of a record as complex128[2^N] (a qiskit Statevector, its stand-in, or an array)."""
    sv = record["clean_state_vec"]
    return np.asarray(getattr(sv, "_data", getattr(sv, "data", sv)), dtype=np.complex128).reshape(-1)


def counts_table(records: list, num_qubits: int):
    """records -> (hist int64[n_rows, 2^N], row_basis list[int], basis_to_idx): one row per measurement record in the
    reference's iteration order (circuit, then measurement; RQC/dataset.py:50-62)."""
    all_bases = ["".join(p) for p in product("XYZ", repeat=num_qubits)]
    basis_to_idx = {b: i for i, b in enumerate(all_bases)}
    dim = 1 << num_qubits
    rows, row_basis = [], []
    for circ in records:
        measurements = circ.get("measurements", []) if "counts" not in circ else [circ]     # SS phase: flat list of measurements
        for meas in measurements:
            name = meas.get("basis", meas.get("basis_str"))
            if name not in basis_to_idx:
                continue                                                                    # RQC/dataset.py:54
            row = np.zeros(dim, dtype=np.int64)
            for key, cnt in meas["counts"].items():
                # counts keys are 'q_{N-1}..q_0'; bits[::-1] makes column i = qubit i (RQC/dataset.py:57-60), so the
                # outcome index with bit i = qubit i is the key read as a binary number
                row[int(key.replace(" ", ""), 2)] += int(cnt)
            rows.append(row)
            row_basis.append(basis_to_idx[name])
    return np.array(rows, dtype=np.int64).reshape(len(rows), dim), row_basis, basis_to_idx


class QuantumStateDataset:
    def __init__(self, data_input, num_qubits: int, device="cuda", seed: int = 1234):
        """data_input: a list of circuit dicts, a directory of ``.pt`` parts, or one ``.pt`` file (RQC/dataset.py:8-39);
        a list of ``{basis_str, basis_idx, counts}`` measurement dicts (SS/dataset.py:6-12) is accepted too."""
        if isinstance(data_input, list):
            raw = data_input
        elif isinstance(data_input, str):
            if os.path.isdir(data_input):
                raw = []
                for f in sorted(glob.glob(os.path.join(data_input, "*.pt"))):
                    raw.extend(load_circuit_records(f))
            elif os.path.isfile(data_input):
                raw = load_circuit_records(data_input)
            else:
                raise FileNotFoundError(f"Path not found: {data_input}")          # RQC/dataset.py:39
        else:
            raise TypeError("data_input must be a list of records or a path")
        self.num_qubits = int(num_qubits)
        hist, row_basis, self.basis_to_idx = counts_table(raw, self.num_qubits)
        self.device = torch.device(device)
        self.seed = int(seed)
        self.n_rows = hist.shape[0]
        if hist.size and hist.max() >= 2 ** 32:
            raise ValueError("a single outcome count does not fit uint32")
        self._total = int(hist.sum())
        if self._total == 0:
            print("WARNING: Dataset is empty.")                                    # RQC/dataset.py:69
        self.hist = torch.from_numpy(hist.astype(np.uint32).view(np.int32)).view(torch.uint32)
        self.row_basis = torch.tensor(row_basis, dtype=torch.int32)
        self._cum = None
        if self.device.type == "cuda":
            self._upload()

    @classmethod
    def from_counts_table(cls, hist: torch.Tensor, num_qubits: int, row_basis=None, device="cuda", seed: int = 1234):
        """Dataset over an existing counts table ``uint32/int[n_rows, 2^N]`` (e.g. ``generate_synthetic_data``'s device
        output or a sampler histogram) without going through Python dicts; ``row_basis`` defaults to the row index."""
        self = cls([], num_qubits, device="cpu", seed=seed)
        h = hist.detach()
        h = h if h.dtype == torch.uint32 else h.to(torch.int32).view(torch.uint32)
        self.hist = h.clone()
        self.n_rows = int(h.shape[0])
        rb = torch.arange(self.n_rows, dtype=torch.int32) if row_basis is None else torch.as_tensor(row_basis, dtype=torch.int32)
        self.row_basis = rb.clone()
        self._total = int(h.view(torch.int32).to(torch.int64).sum().item())
        return self.to(device)

    # ------------------------------------------------------------------ device tables
    def _upload(self):
        lib = _lib.load()
        self.hist = self.hist.to(self.device).contiguous()
        self.row_basis = self.row_basis.to(self.device).contiguous()
        self._cum = torch.empty_like(self.hist)
        self._row_start = torch.empty(self.n_rows + 1, dtype=torch.int64, device=self.device)
        _lib.check(lib.ddqst_counts_scan(_lib.ptr(self.hist) if self.n_rows else None, self.n_rows, self.num_qubits,
                                         _lib.ptr(self._cum) if self.n_rows else None, _lib.ptr(self._row_start), _lib.stream_ptr()))

    def to(self, device):
        self.device = torch.device(device)
        if self.device.type == "cuda":
            self._upload()
        return self

    def _gather(self, start: int, count: int, permute: bool, epoch: int, want_bits: bool):
        if self._cum is None:
            raise RuntimeError("QuantumStateDataset has no CPU path: construct it with a cuda device")
        if self._total == 0:
            raise IndexError("dataset is empty")
        lib = _lib.load()
        x0 = torch.empty(count, dtype=torch.uint16, device=self.device)
        basis = torch.empty(count, dtype=torch.int32, device=self.device)
        bits = torch.empty(count, self.num_qubits, dtype=torch.int64, device=self.device) if want_bits else None
        _lib.check(lib.ddqst_counts_gather(_lib.ptr(self._cum), _lib.ptr(self._row_start), _lib.ptr(self.row_basis), self.n_rows,
                                           self.num_qubits, self._total, int(permute), self.seed, int(epoch), int(start), int(count),
                                           _lib.ptr(x0), _lib.ptr(basis), _lib.ptr(bits), _lib.stream_ptr()))
        return x0, basis, bits

    # ------------------------------------------------------------------ the reference's surface
    def __len__(self):
        return self._total

    def __getitem__(self, idx):
        """-> (bits[N] int64, basis_idx int64 scalar), canonical unrolled order (RQC/dataset.py:77-78)."""
        idx = int(idx)
        if idx < 0:
            idx += self._total
        if not 0 <= idx < self._total:
            raise IndexError(idx)
        _, basis, bits = self._gather(idx, 1, False, 0, True)
        return bits[0], basis[0].to(torch.int64)

    @property
    def data_tensor(self):
        """int64[len, N] (RQC/dataset.py:66): the whole unroll, materialised on the device on request."""
        return self._gather(0, self._total, False, 0, True)[2]

    @property
    def basis_tensor(self):
        return self._gather(0, self._total, False, 0, False)[1].to(torch.int64)

    # ------------------------------------------------------------------ training batches
    def batches_per_epoch(self, batch_size: int) -> int:
        return (self._total + batch_size - 1) // batch_size

    def batch(self, step: int, batch_size: int, shuffle: bool = True):
        """Batch ``step`` of an endless DataLoader(shuffle=True) stream: positions step*B .. step*B+B-1 of epoch
        ``step*B // len`` (each epoch is its own keyed permutation of the unrolled shots; a batch that straddles an epoch
        boundary wraps inside its starting epoch).  -> (x0_packed uint16[B], basis int32[B]) as train_step consumes them."""
        start = step * batch_size
        epoch, off = divmod(start, self._total)
        x0, basis, _ = self._gather(off, batch_size, shuffle, epoch, False)
        return x0, basis
