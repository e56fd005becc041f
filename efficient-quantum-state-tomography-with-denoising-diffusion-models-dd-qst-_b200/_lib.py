"""ctypes binding of libddqst.so (include/ddqst.h).  No CPU fallback: if the library is missing and cannot
be built, or a call is made without a CUDA tensor, this raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _build

_HERE = os.path.dirname(os.path.abspath(__file__))

PRECISION_FP32, PRECISION_BF16 = 0, 1
MODE_POSTERIOR, MODE_RENOISE = 0, 1
VARIANT_A, VARIANT_B = 0, 1
KRON_REVERSED, KRON_UNREVERSED = 0, 1
TARGET_NONE, TARGET_STATEVECTOR, TARGET_MIXED, TARGET_RANK_ONE = range(4)
OP_FORWARD, OP_SAMPLE, OP_LINEAR_INVERSION, OP_PSD, OP_FIDELITY_MIXED, OP_TRAIN, OP_METRICS = range(7)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_qubits", "num_bases", "num_timesteps", "embed_dim", "hidden_dim",
                                         "num_blocks", "variant")]


class Mlpdims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_bases", "num_timesteps", "embed_dim", "hidden_dim", "num_hidden")]


_P, _I32, _I64, _U64, _U32, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_float
_DP = C.POINTER(Dims)
_MP = C.POINTER(Mlpdims)

# name -> (restype, argtypes); mirrors include/ddqst.h one to one
SIGNATURES = {
    "ddqst_last_error": (C.c_char_p, []),
    "ddqst_version": (C.c_int, []),
    "ddqst_param_count": (_I64, [_DP, _P]),
    "ddqst_pack_bytes": (_I64, [_DP]),
    "ddqst_pack_weights": (C.c_int, [_DP, _P, _P, _P]),
    "ddqst_denoiser_forward": (C.c_int, [_DP, _P, C.c_int, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "ddqst_sample": (C.c_int, [_DP, _P, _P, C.c_int, C.c_int, _P, _I32, _I64, _I64, _U64, _P, _P, _P, _I64, _P]),
    "ddqst_sample_step": (C.c_int, [_DP, _P, _P, C.c_int, C.c_int, _I32, _I32, _I64, _I64, _U64, _P, _P, _P, _P, _I64, _P]),
    "ddqst_q_sample": (C.c_int, [_P, _I32, _I32, C.c_int, _P, _P, _I64, _I64, _U64, _U32, _P, _P, _P]),
    "ddqst_q_sample_dev": (C.c_int, [_P, _I32, _I32, C.c_int, _P, _P, _I64, _I64, _U64, _P, _P, _P, _P]),
    "ddqst_histogram": (C.c_int, [_P, C.c_int, _I64, _I32, _P, _P]),
    "ddqst_pack_bits": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "ddqst_unpack_bits": (C.c_int, [_P, C.c_int, _I64, _I32, _P, _P]),
    "ddqst_linear_inversion": (C.c_int, [_P, _P, _I32, _I32, _P, C.c_int, _P, _P, _I64, _P]),
    "ddqst_psd_project": (C.c_int, [_P, _I32, _P, _P, _I64, _P]),
    "ddqst_fidelity_pure": (C.c_int, [_P, _P, _I32, _P, _P]),
    "ddqst_fidelity_mixed": (C.c_int, [_P, _P, _I32, _P, _P, _I64, _P]),
    "ddqst_metrics": (C.c_int, [_P, _I32, _P, _P, _I64, _P]),
    "ddqst_recon_report": (C.c_int, [_P, _I32, _P, C.c_int, _P, _P, _P, _I64, _P]),
    "ddqst_train_forward_backward": (C.c_int, [_DP, _P, _P, _P, _P, _P, _I64, _F, _P, _P, _P, _I64, _P]),
    "ddqst_forward_saved": (C.c_int, [_DP, _P, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "ddqst_backward_saved": (C.c_int, [_DP, _P, _P, _P, _P, _I64, _P, _P, _P, _I64, _P]),
    "ddqst_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _F, _F, _F, _F, _F, C.c_int, _F, _P]),
    "ddqst_cast_bf16": (C.c_int, [_P, _P, _I64, _P]),
    "ddqst_train_forward_backward_tc": (C.c_int, [_DP, _P, _P, _P, _P, _P, _P, _I64, _F, _P, _P, _P, _I64, _P]),
    "ddqst_adam_step_dev": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _F, C.c_int, _F, _P]),
    "ddqst_selftest_gemm_tc": (C.c_int, [_P, _P, C.c_int, C.c_int, _I32, _I32, _I32, _I32, _P, _P]),
    "ddqst_counts_scan": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "ddqst_counts_gather": (C.c_int, [_P, _P, _P, _I64, _I32, _I64, C.c_int, _U64, _U64, _I64, _I64, _P, _P, _P, _P]),
    "ddqst_synth_state": (C.c_int, [_I32, C.c_int, _I32, _U64, _P, _P]),
    "ddqst_synth_born_histograms": (C.c_int, [_P, _I32, _P, _I32, _I64, _U64, C.c_double, C.c_double, _P, _P, _P]),
    "ddqst_workspace_bytes": (_I64, [C.c_int, _DP, _I64, C.c_int]),
    "ddqst_sample_host": (C.c_int, [_DP, _P, _P, C.c_int, C.c_int, _P, _I32, _I64, _I64, _U64, _P, _P, _P, _I64, _P]),
    "ddqst_mlp_param_count": (_I64, [_MP, _P]),
    "ddqst_mlp_workspace_bytes": (_I64, [_MP, _I64]),
    "ddqst_mlp_forward_saved": (C.c_int, [_MP, _P, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "ddqst_mlp_backward_saved": (C.c_int, [_MP, _P, _P, _P, _I64, _P, _P, _P, _I64, _P]),
    "ddqst_mlp_sample": (C.c_int, [_MP, _P, _P, _I32, _I64, _I64, _U64, _P, _P, _P, _I64, _P]),
    "ddqst_selftest_philox": (C.c_int, [_P, _I64, _P, _P]),
    "ddqst_selftest_umma": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "ddqst_selftest_umma2": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "ddqst_selftest_gemm_tc_dbg": (C.c_int, [_P, _P, C.c_int, C.c_int, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "ddqst_debug_tc_trace": (C.c_int, [_P, _I32]),
    "ddqst_debug_tc_status": (C.c_int, []),
    "ddqst_debug_ft_stamps": (C.c_int, [_P]),
    "ddqst_debug_train_path": (C.c_int, [C.c_int]),
}

_lib = None


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if the .so is absent and nvcc exists).  Raises if neither is possible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.have_nvcc():
        _build.build()                   # incremental (content hash of csrc/ + include/ + flags): a stale .so is rebuilt
    elif not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and nvcc is not available: run `python __graft_entry__.py` / build() "
                           "where nvcc exists (no CPU fallback)")
    elif not _build.stamp_matches():
        raise RuntimeError(f"{path} was built from different sources than csrc/ + include/ddqst.h now hold and nvcc is "
                           "not available to rebuild it")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().ddqst_last_error().decode(errors="replace")
        raise RuntimeError(f"libddqst error {status}: {msg}")


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libddqst has no CPU path: tensors must live on a CUDA device")
    if not t.is_contiguous():
        raise RuntimeError("libddqst needs contiguous tensors")
    return C.c_void_p(t.data_ptr())


def check_index(idx, hi: int, what: str, lo: int = 0) -> None:
    """Raise IndexError unless every element of ``idx`` (tensor, list or int) lies in [lo, hi).

    The kernels index their embedding / FiLM / schedule tables with these values unchecked (include/ddqst.h,
    "index contract"), as ``nn.Embedding`` and ``Q_bar[t]`` raise IndexError in the reference (RQC/model.py:59-61,
    RQC/diffusion.py:48).  One ``aminmax`` (a device sync for CUDA tensors); callers inside a CUDA-graph capture
    skip it and own the contract."""
    if torch.is_tensor(idx):
        if idx.numel() == 0:
            return
        mn, mx = (int(v) for v in torch.aminmax(idx))
    else:
        seq = [int(idx)] if isinstance(idx, int) else [int(v) for v in idx]
        if not seq:
            return
        mn, mx = min(seq), max(seq)
    if mn < lo or mx >= hi:
        raise IndexError(f"{what} out of range: values span [{mn}, {mx}], valid range is [{lo}, {hi})")


def capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Workspace:
    """Grow-only scratch buffer per device; the library never allocates."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes: int, device) -> torch.Tensor:
        key = torch.device(device).index or 0
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


workspace = Workspace()
