"""Harness that runs the UNMODIFIED reference (read-only, /root/reference) on CPU.

TEST INFRASTRUCTURE.  Used only by tests/test_oracle_vs_reference.py and
tests/golden/make_golden.py, and only where /root/reference is mounted (the
build container).  Nothing here travels to the GPU box at run time; the golden
fixtures it produced do.

It (1) stubs ``qiskit.quantum_info`` so ``*/reconstruct.py`` imports unmodified,
(2) loads a reference phase directory as an isolated set of modules,
(3) patches ``torch.randint`` / ``torch.multinomial`` with the injected Philox
stream of ``oracle.ddqst_oracle`` (the reference is unseeded),
(4) decodes ``Datapoints/*.pt`` without qiskit through a stub unpickler.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import json
import os
import pickle
import sys
import types
import zipfile

import numpy as np
import torch

from . import ddqst_oracle as orc

REF_ROOT = "/root/reference"
LOCAL_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # oracle/make_ref.py: verbatim copies of the path's modules
PHASES = {
    "SS": "versions/multi_qubit_special_states",
    "AS": "versions/multi_qubit_any_state",
    "RQC": "versions/RQC_dataset_building_phase",
}
NOTEBOOK = "versions/single_qubit_phase/denoising-with-diffusion-phase-1.ipynb"


def available() -> bool:
    """The whole read-only checkout (notebook, Datapoints, every phase) is mounted -- build container only."""
    return os.path.isdir(os.path.join(REF_ROOT, "versions"))


def path_root() -> str | None:
    """Where the hot-path modules (model / diffusion / reconstruct of the RQC and SS phases) can be imported from:
    the mounted checkout, else the verbatim copy ``oracle/make_ref.py`` placed in ``oracle/_ref`` (which travels to the
    GPU box), else None."""
    if available():
        return REF_ROOT
    if os.path.exists(os.path.join(LOCAL_REF, "MANIFEST.json")):
        return LOCAL_REF
    return None


# ---------------------------------------------------------------- qiskit stub
class _DensityMatrix:
    def __init__(self, data):
        d = getattr(data, "data", data)
        d = np.asarray(d, dtype=complex)
        self.data = np.outer(d, d.conj()) if d.ndim == 1 else d

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class _Statevector:
    def __init__(self, data):
        self.data = np.asarray(getattr(data, "data", data), dtype=complex)


def install_qiskit_stub():
    if "qiskit" in sys.modules and not getattr(sys.modules["qiskit"], "_ddqst_stub", False):
        return
    q = types.ModuleType("qiskit")
    q._ddqst_stub = True
    qi = types.ModuleType("qiskit.quantum_info")
    qi.DensityMatrix = _DensityMatrix
    qi.Statevector = _Statevector
    qi.state_fidelity = lambda a, b: orc.state_fidelity(getattr(a, "data", a), getattr(b, "data", b))
    qi.entropy = lambda r, base=2: orc.entropy_bits(getattr(r, "data", r))

    def _ptrace(rho, qargs):
        d = getattr(rho, "data", rho)
        n = int(np.log2(d.shape[0]))
        keep = [i for i in range(n) if i not in qargs]
        assert keep == list(range(len(keep))), "stub only supports tracing the high qubits"
        return _DensityMatrix(orc.partial_trace_high(d, n, len(keep)))

    qi.partial_trace = _ptrace
    q.quantum_info = qi
    sys.modules["qiskit"] = q
    sys.modules["qiskit.quantum_info"] = qi


# ------------------------------------------------------------- module loading
def load_phase(phase: str, names=("model", "diffusion", "reconstruct")) -> dict:
    """Import reference modules of one phase under private names (phases share file names)."""
    install_qiskit_stub()
    out = {}
    root = path_root()
    if root is None:
        raise FileNotFoundError("neither /root/reference nor oracle/_ref (python oracle/make_ref.py) is present")
    base = os.path.join(root, PHASES[phase])
    for n in names:
        spec = importlib.util.spec_from_file_location(f"_ddqst_ref_{phase}_{n}", os.path.join(base, n + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out[n] = mod
    return out


def load_notebook_classes(cell: int = 6) -> dict:
    """exec the class definitions of a notebook cell (NB c6: SimpleMLP/BitstringDDM; c12: UpgradedMLP)."""
    nb = json.load(open(os.path.join(REF_ROOT, NOTEBOOK)))
    src = "".join(nb["cells"][cell]["source"])
    cut = src.index("# --- Main Execution ---")
    ns = {"__name__": "_ddqst_ref_nb"}
    exec(compile(src[:cut], f"NB c{cell}", "exec"), ns)
    return ns


# ---------------------------------------------------------------- RNG patches
class InjectedStream:
    """Stateful replacement of torch.randint / torch.multinomial for one reference call.

    mode 'posterior'  : RQC p_sample -- 1 randint([B,N]) + T multinomial([B*N,2])
    mode 'renoise'    : SS  p_sample -- 1 randint + per step multinomial([B*N,2]) then N x multinomial([B,2])
    mode 'renoise_nb' : NB  sample   -- same with N=1 and 1-D tensors
    mode 'q_cumulative': RQC q_sample -- B calls of multinomial([N,2])
    mode 'q_marginal' : SS q_sample  -- N calls of multinomial([B,2])
    mode 'train'      : randint(1,T+1,(B,)) then the q_sample pattern (cumulative or marginal)
    """

    def __init__(self, mode, seed, stream, num_qubits, num_timesteps, batch, offset=0, cumulative=True):
        self.mode, self.seed, self.stream = mode, seed, stream
        self.N, self.T, self.B, self.offset = num_qubits, num_timesteps, batch, offset
        self.cumulative = cumulative
        self.idx = offset + np.arange(batch)
        self.t = num_timesteps
        self.sub = 0          # sub-call counter inside a step / q_sample
        self.phase = "x0hat"

    def randint(self, low, high, size, **kw):
        if self.mode in ("posterior", "renoise", "renoise_nb"):
            bits = orc.init_bits(self.seed, self.stream, self.B, self.N, self.offset)
            return bits.reshape(tuple(size))
        if self.mode == "train":
            return torch.from_numpy(orc.stream_timesteps(self.seed, self.stream, self.idx, self.T)).reshape(tuple(size))
        raise AssertionError("unexpected randint")

    def _u(self, t, site):
        if getattr(self, "_u_key", None) != (t, site):
            self._u_key = (t, site)
            self._u_val = torch.from_numpy(orc.stream_uniforms(self.seed, self.stream, t, site, self.idx, self.N))
        return self._u_val

    def multinomial(self, probs, num_samples, *a, **kw):
        assert num_samples == 1
        if self.mode == "posterior":
            u = self._u(self.t, orc.SITE_POSTERIOR).reshape(-1)
            self.t -= 1
            return orc.draw_bits(probs, u).view(-1, 1)
        if self.mode in ("renoise", "renoise_nb"):
            if self.phase == "x0hat":
                u = self._u(self.t, orc.SITE_X0HAT).reshape(-1)
                out = orc.draw_bits(probs, u).view(-1, 1)
                if self.t > 1:
                    self.phase, self.sub = "renoise", 0
                else:
                    self.t -= 1
                return out
            u = self._u(self.t, orc.SITE_RENOISE)[:, self.sub]
            self.sub += 1
            if self.sub == self.N:
                self.phase = "x0hat"
                self.t -= 1
            return orc.draw_bits(probs, u).view(-1, 1)
        if self.mode in ("q_cumulative", "train") and self.cumulative:
            u = self._u(0, orc.SITE_QSAMPLE)[self.sub]
            self.sub += 1
            return orc.draw_bits(probs, u).view(-1, 1)
        if self.mode in ("q_marginal", "train"):
            u = self._u(0, orc.SITE_QSAMPLE)[:, self.sub]
            self.sub += 1
            return orc.draw_bits(probs, u).view(-1, 1)
        raise AssertionError("unexpected multinomial")


@contextlib.contextmanager
def injected(stream: InjectedStream):
    orig_r, orig_m = torch.randint, torch.multinomial
    torch.randint, torch.multinomial = stream.randint, stream.multinomial
    try:
        yield stream
    finally:
        torch.randint, torch.multinomial = orig_r, orig_m


# ------------------------------------------------------- Datapoints unpickler
class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._state = state


class _Counts(dict):
    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("qiskit"):
            if name == "Counts":
                return _Counts
            return type(name, (_Stub,), {})
        return super().find_class(module, name)


def load_datapoints(path: str) -> list:
    """Decode one Datapoints/rqc_N3_data/part_*.pt (torch zip-pickle referencing qiskit classes)."""
    with zipfile.ZipFile(path) as z:
        pkl = [n for n in z.namelist() if n.endswith("data.pkl")][0]
        prefix = pkl[: -len("data.pkl")]

        def persistent_load(pid):
            # ('storage', storage_type, key, location, numel)
            _, storage_type, key, _, numel = pid
            dtype = getattr(storage_type, "dtype", None) or torch.uint8
            raw = z.read(f"{prefix}data/{key}")
            return torch.frombuffer(bytearray(raw), dtype=dtype).untyped_storage()

        up = _RefUnpickler(io.BytesIO(z.read(pkl)))
        up.persistent_load = persistent_load
        return up.load()


def record_to_arrays(rec: dict, num_qubits: int):
    """-> (psi complex128[2^N], hist int64[3^N, 2^N]) with bit i of the outcome index = qubit i
    (the counts keys are 'q_{N-1}...q_0' strings, flipped as in RQC/dataset.py:59, RQC/evaluate.py:27)."""
    sv = rec["clean_state_vec"]
    psi = np.asarray(getattr(sv, "_data", getattr(sv, "data", sv)), dtype=complex).reshape(-1)
    names = orc.basis_strings(num_qubits)
    hist = np.zeros((len(names), 1 << num_qubits), dtype=np.int64)
    for m in rec["measurements"]:
        b = names.index(m["basis"])
        for key, cnt in m["counts"].items():
            bits = [int(c) for c in key][::-1]
            hist[b, sum(v << i for i, v in enumerate(bits))] += int(cnt)
    return psi, hist


def expand_hist_to_samples(hist_row: np.ndarray, num_qubits: int) -> np.ndarray:
    """Histogram row -> int64[shots, N] sample matrix (column i = qubit i), outcome-sorted."""
    idx = np.repeat(np.arange(hist_row.shape[0]), hist_row)
    return ((idx[:, None] >> np.arange(num_qubits)) & 1).astype(np.int64)
