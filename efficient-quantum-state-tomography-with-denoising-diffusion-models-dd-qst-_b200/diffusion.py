"""DiscreteDiffusion: the reference's diffusion call surface on top of libddqst.

``DiscreteDiffusion(model, num_timesteps, device)`` with ``.betas``, ``.Q_bar`` / ``.Q``, ``.q_sample(x_0, t)``
and ``.p_sample(num_samples, basis_idx, num_qubits)`` as in RQC/diffusion.py:5-80 (schedule="cosine",
true D3PM posterior sampler) and SS/diffusion.py:6-82 (schedule="linear", "predict x0 then re-noise" sampler),
plus the batched multi-basis ``sample(bases, n_shots)`` and the fused ``train_step`` that north_star adds.

Randomness is the counter-based Philox stream documented in include/ddqst.h (the reference is unseeded):
results depend only on (seed, basis, shot index, t), never on batch split or rank count.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .model import ConditionalD3PM, pack_bits, unpack_bits


def cosine_schedule(num_timesteps: int):
    """betas[T+1] fp32 and Q_bar[T+1,2,2] fp32 exactly as RQC/diffusion.py:15-43 builds them
    (float64 alpha-bar, clip 0.999, then a sequential fp32 matmul chain Q_t @ Q_bar[t-1])."""
    s = np.arange(num_timesteps + 1, dtype=np.float64) / num_timesteps
    abar = np.cos((s + 0.008) / 1.008 * np.pi / 2) ** 2
    abar = abar / abar[0]
    b = [0.0] + [min(1 - abar[i] / abar[i - 1], 0.999) for i in range(1, num_timesteps + 1)]
    betas = torch.tensor(b, dtype=torch.float32)
    q_bar = torch.zeros(num_timesteps + 1, 2, 2)
    cur = torch.eye(2)
    q_bar[0] = cur
    for t in range(1, num_timesteps + 1):
        bt = betas[t]
        cur = torch.tensor([[1 - bt, bt], [bt, 1 - bt]]) @ cur
        q_bar[t] = cur
    return betas, q_bar


def linear_schedule(num_timesteps: int):
    """SS/diffusion.py:14-25: beta = linspace(0.001, 0.5, T+1), Q[t] = [[1-b, b], [b, 1-b]] used as the marginal channel."""
    betas = torch.linspace(0.001, 0.5, num_timesteps + 1)
    q = torch.zeros(num_timesteps + 1, 2, 2)
    for t in range(num_timesteps + 1):
        bt = betas[t]
        q[t] = torch.tensor([[1 - bt, bt], [bt, 1 - bt]])
    return betas, q


class NativeAdam:
    """State of the fused Adam / AdamW kernel (ddqst_adam_step) over the model's flat parameter buffer.
    Defaults follow RQC/main.py:98 (Adam, lr 1e-3); ``decoupled=True, weight_decay=0.01`` gives SS/main.py:77 (AdamW)."""

    def __init__(self, model: ConditionalD3PM, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.weight_decay, self.decoupled = weight_decay, decoupled
        self.exp_avg = torch.zeros_like(model.flat_params)
        self.exp_avg_sq = torch.zeros_like(model.flat_params)
        self.step_count = 0
        self._step_dev = None

    @property
    def step_dev(self) -> torch.Tensor:
        """int64[2] in device memory: [0] = count of completed steps (read by the graph-replayable kernels), [1] = scratch."""
        p = self.model.flat_params
        if self._step_dev is None or self._step_dev.device != p.device:
            self._step_dev = torch.tensor([self.step_count, 0], dtype=torch.int64, device=p.device)
        return self._step_dev

    def step_device(self, grads: torch.Tensor, grad_scale: float = 1.0, shadow: torch.Tensor | None = None):
        """The same update with the step count read from (and incremented in) device memory and an optional bf16
        shadow of the new parameters: replayable from a CUDA graph (ddqst_adam_step_dev)."""
        lib = _lib.load()
        p = self.model.flat_params
        sd = self.step_dev
        _lib.check(lib.ddqst_adam_step_dev(_lib.ptr(p), _lib.ptr(shadow), _lib.ptr(grads), _lib.ptr(self.exp_avg),
                                           _lib.ptr(self.exp_avg_sq), p.numel(), _lib.ptr(sd), self.lr, self.betas[0],
                                           self.betas[1], self.eps, self.weight_decay, int(self.decoupled), grad_scale,
                                           _lib.stream_ptr()))
        self.step_count += 1
        self.model.native_version += 1
        if shadow is not None:
            self.model.mark_shadow_current()

    def step(self, grads: torch.Tensor, grad_scale: float = 1.0):
        lib = _lib.load()
        if self._step_dev is not None:
            return self.step_device(grads, grad_scale)       # keep the device-side counter authoritative once it exists
        self.step_count += 1
        p = self.model.flat_params
        _lib.check(lib.ddqst_adam_step(_lib.ptr(p), _lib.ptr(grads), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                       p.numel(), self.step_count, self.lr, self.betas[0], self.betas[1], self.eps,
                                       self.weight_decay, int(self.decoupled), grad_scale, _lib.stream_ptr()))
        self.model.native_version += 1


class TrainGraph:
    """A captured tensor-core training step (DiscreteDiffusion.make_train_graph).  ``replay()`` runs one step on whatever
    the static input buffers hold now and returns the loss (device scalar, overwritten by the next replay)."""

    def __init__(self, graph, loss, diffusion, optimizer):
        self.graph, self.loss, self.diffusion, self.optimizer = graph, loss, diffusion, optimizer

    def replay(self):
        self.graph.replay()
        self.optimizer.step_count += 1
        self.diffusion._train_steps += 1
        self.diffusion.model.native_version += 1      # packed inference tables are stale now
        self.diffusion.model.mark_shadow_current()    # ... but the bf16 shadow was updated by the captured Adam
        return self.loss


class DiscreteDiffusion:
    def __init__(self, model: ConditionalD3PM, num_timesteps: int, device, schedule: str = "cosine",
                 seed: int = 1234, precision: str = "bf16"):
        if schedule not in ("cosine", "linear"):
            raise ValueError("schedule must be 'cosine' (RQC) or 'linear' (SS)")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (tcgen05) or 'fp32' (exact)")
        self.model, self.num_timesteps, self.device = model, num_timesteps, torch.device(device)
        self.schedule, self.seed, self.precision = schedule, int(seed), precision
        if schedule == "cosine":
            betas, q = cosine_schedule(num_timesteps)
            self.betas, self.Q_bar = betas.to(self.device), q.to(self.device)
            self.mode, self.cumulative = _lib.MODE_POSTERIOR, 1
        else:
            betas, q = linear_schedule(num_timesteps)
            self.betas, self.Q = betas.to(self.device), q.to(self.device)
            self.mode, self.cumulative = _lib.MODE_RENOISE, 0
        self._q = q.to(self.device).contiguous()
        self._sched = torch.cat([betas.reshape(-1), q.reshape(-1)]).to(self.device).contiguous()
        self._q_calls = 0            # stream id of standalone q_sample calls
        self._train_steps = 0        # stream id of train steps

    # ------------------------------------------------------------------ helpers
    def _prec(self):
        return _lib.PRECISION_BF16 if self.precision == "bf16" else _lib.PRECISION_FP32

    def _require_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("DiscreteDiffusion has no CPU path: construct it with a cuda device")

    # ------------------------------------------------------------------ forward noising (D2 / D2')
    def q_sample(self, x_0: torch.Tensor, t: torch.Tensor, row_offset: int = 0, stream_id: int | None = None):
        """x_0[B,N] int64, t[B] int64 -> x_t[B,N] int64 (RQC/diffusion.py:45-51 / SS/diffusion.py:27-52)."""
        self._require_cuda()
        lib = _lib.load()
        N = x_0.shape[1]
        if stream_id is None:
            stream_id = self._q_calls
            self._q_calls += 1
        _lib.check_index(t, self.num_timesteps + 1, "t")          # Q_bar[t] raises IndexError in the reference
        x0p = pack_bits(x_0.to(self.device), N)
        t32 = t.to(self.device).to(torch.int32).contiguous()
        xtp = torch.empty_like(x0p)
        _lib.check(lib.ddqst_q_sample(_lib.ptr(self._q), self.num_timesteps, N, self.cumulative, _lib.ptr(x0p), _lib.ptr(t32),
                                      x0p.shape[0], row_offset, self.seed, stream_id, _lib.ptr(xtp), None, _lib.stream_ptr()))
        return unpack_bits(xtp, N)

    # ------------------------------------------------------------------ reverse sampling (D3 / D3')
    def sample(self, bases, n_shots: int, shot_offset: int = 0, return_bits: bool = False, return_hist: bool = True,
               hist_out: torch.Tensor | None = None):
        """Generate ``n_shots`` bitstrings for every basis index in ``bases`` in one launch.

        Returns (hist uint32[len(bases), 2^N] or None, packed uint8/uint16[len(bases), n_shots] or None)."""
        self._require_cuda()
        lib = _lib.load()
        m = self.model
        N = m.num_qubits
        if not _lib.capturing():
            _lib.check_index(bases, m.num_bases, "basis index")       # the FiLM table Tb has num_bases rows
        if torch.is_tensor(bases):
            ids = bases.to(device=self.device, dtype=torch.int32).contiguous()
        else:
            ids = torch.tensor(list(bases), dtype=torch.int32, device=self.device)
        nb = ids.numel()
        hist = None
        if return_hist:
            hist = hist_out if hist_out is not None else torch.zeros(nb, 1 << N, dtype=torch.uint32, device=self.device)
        packed = None
        if return_bits:
            packed = torch.empty(nb, n_shots, dtype=torch.uint8 if N <= 8 else torch.uint16, device=self.device)
        prec = self._prec()
        nbytes = lib.ddqst_workspace_bytes(_lib.OP_SAMPLE, C.byref(m.dims), nb * n_shots, prec)
        ws = _lib.workspace.get(nbytes, self.device)
        _lib.check(lib.ddqst_sample(C.byref(m.dims), _lib.ptr(m.packed()), _lib.ptr(self._sched), self.mode, prec,
                                    _lib.ptr(ids), nb, n_shots, shot_offset, self.seed, _lib.ptr(packed), _lib.ptr(hist),
                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return hist, packed

    def sample_to_host(self, basis_ids_host: torch.Tensor, n_shots: int, out_packed_host: torch.Tensor | None,
                       out_hist_host: torch.Tensor | None, shot_offset: int = 0):
        """End-to-end form with HOST buffers (ddqst_sample_host): basis ids are copied in, the packed bitstrings
        and/or the per-basis counts are copied out (pinned memory recommended) and the stream is synchronised --
        what the reference's ``p_sample(...).cpu().numpy()`` loop does per basis (RQC/evaluate.py:82-84)."""
        self._require_cuda()
        lib = _lib.load()
        m = self.model
        N = m.num_qubits
        nb = basis_ids_host.numel()
        if basis_ids_host.dtype != torch.int32 or basis_ids_host.is_cuda:
            raise ValueError("basis_ids_host must be a host int32 tensor")
        _lib.check_index(basis_ids_host, m.num_bases, "basis index")
        need = 4 * nb + (nb << N) * 4 + nb * n_shots * (1 if N <= 8 else 2) + 4096 + \
            lib.ddqst_workspace_bytes(_lib.OP_SAMPLE, C.byref(m.dims), nb * n_shots, self._prec())
        scratch = _lib.workspace.get(need, self.device)
        hp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        _lib.check(lib.ddqst_sample_host(C.byref(m.dims), _lib.ptr(m.packed()), _lib.ptr(self._sched), self.mode, self._prec(),
                                         hp(basis_ids_host), nb, n_shots, shot_offset, self.seed, hp(out_packed_host),
                                         hp(out_hist_host), _lib.ptr(scratch), scratch.numel(), _lib.stream_ptr()))

    @torch.no_grad()
    def p_sample(self, num_samples: int, basis_idx: int, num_qubits: int, shot_offset: int = 0):
        """-> x_0[num_samples, N] int64 on ``device`` (RQC/diffusion.py:53-80 / SS/diffusion.py:54-82)."""
        if num_qubits != self.model.num_qubits:
            raise ValueError(f"num_qubits={num_qubits} does not match the model ({self.model.num_qubits})")
        _, packed = self.sample([int(basis_idx)], int(num_samples), shot_offset, return_bits=True, return_hist=False)
        return unpack_bits(packed.view(-1), num_qubits)

    def sample_step(self, x_t: torch.Tensor, basis_idx: int, t: int, shot_offset: int = 0, precision: str | None = None):
        """Teacher-forced single reverse step: x_t[B,N] int64 -> (x_{t-1}[B,N] int64, logits[B,N,2])."""
        self._require_cuda()
        lib = _lib.load()
        m = self.model
        N = m.num_qubits
        prec = self._prec() if precision is None else (_lib.PRECISION_BF16 if precision == "bf16" else _lib.PRECISION_FP32)
        xp = pack_bits(x_t.to(self.device), N)
        B = xp.shape[0]
        out = torch.empty_like(xp)
        logits = torch.empty(B, N, 2, dtype=torch.float32, device=self.device)
        nbytes = max(lib.ddqst_workspace_bytes(_lib.OP_SAMPLE, C.byref(m.dims), B, _lib.PRECISION_FP32),
                     B * (m.hidden_dim * 12 + N * 8 + 8) + 8192)
        ws = _lib.workspace.get(nbytes, self.device)
        _lib.check(lib.ddqst_sample_step(C.byref(m.dims), _lib.ptr(m.packed()), _lib.ptr(self._sched), self.mode, prec,
                                         int(basis_idx), int(t), B, shot_offset, self.seed, _lib.ptr(xp), _lib.ptr(out),
                                         _lib.ptr(logits), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return unpack_bits(out, N), logits

    # ------------------------------------------------------------------ training step (T1)
    def train_precision(self) -> str:
        """'bf16' (tcgen05 GEMMs, ddqst_train_forward_backward_tc) when this object samples in bf16 and the model has
        tensor-core-friendly dims, else 'fp32' (CUDA-core exact path)."""
        m = self.model
        ok = (self.precision == "bf16" and m.hidden_dim % 64 == 0 and m.embed_dim % 16 == 0
              and m.num_qubits <= 15 and m.num_blocks <= 16)
        return "bf16" if ok else "fp32"

    def _dp(self, process_group, data_parallel):
        """-> (world, rank) of the data-parallel group, (1, 0) unless the caller asked for data parallelism.  An initialised
        default process group alone does NOT turn the all-reduce on: a step run by a subset of the ranks would block forever."""
        if process_group is None and not data_parallel:
            return 1, 0
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            raise RuntimeError("data-parallel train_step needs an initialised torch.distributed process group")
        return torch.distributed.get_world_size(process_group), torch.distributed.get_rank(process_group)

    def train_step(self, x_0: torch.Tensor, basis: torch.Tensor, optimizer: NativeAdam, row_offset: int | None = None,
                   process_group=None, precision: str | None = None, data_parallel: bool = False, validate: bool = True):
        """One step of RQC/main.py:105-115 fused: t ~ U{1..T}, x_t = q_sample(x_0, t), logits, mean cross-entropy,
        backward, Adam.  Returns the loss (device scalar; the local batch's mean).
        ``precision`` 'bf16' runs every GEMM on the tensor cores, 'fp32' the exact CUDA-core path (default: train_precision()).
        Data parallel (``process_group=...`` or ``data_parallel=True`` for the default group; EVERY rank of the group must
        call): the flat gradient is summed by one NCCL all-reduce and scaled by 1/world inside the Adam kernel;
        ``row_offset`` (the global row index of local row 0, which keys the t / noise draws) then defaults to rank * B so
        the ranks draw disjoint streams.  ``validate`` range-checks ``basis`` (one device sync; IndexError as nn.Embedding
        raises in the reference) -- skipped during CUDA-graph capture."""
        self._require_cuda()
        if validate and not _lib.capturing():
            _lib.check_index(basis, self.model.num_bases, "basis")
        world, rank = self._dp(process_group, data_parallel)
        if row_offset is None:
            row_offset = rank * x_0.shape[0] if world > 1 else 0
        if (precision or self.train_precision()) == "bf16":
            return self._train_step_tc(x_0, basis, optimizer, row_offset, process_group, world)
        lib = _lib.load()
        m = self.model
        N = m.num_qubits
        step = self._train_steps
        self._train_steps += 1
        x0p = x_0 if x_0.dtype == torch.uint16 else pack_bits(x_0.to(self.device), N)
        b32 = basis.to(self.device).to(torch.int32).contiguous()
        B = x0p.shape[0]
        xtp = torch.empty_like(x0p)
        t32 = torch.empty(B, dtype=torch.int32, device=self.device)
        _lib.check(lib.ddqst_q_sample(_lib.ptr(self._q), self.num_timesteps, N, self.cumulative, _lib.ptr(x0p), None, B,
                                      row_offset, self.seed, step, _lib.ptr(xtp), _lib.ptr(t32), _lib.stream_ptr()))
        grads = getattr(self, "_grads", None)
        if grads is None or grads.numel() != m.flat_params.numel() or grads.device != m.flat_params.device:
            grads = self._grads = torch.empty_like(m.flat_params)
            self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        nbytes = lib.ddqst_workspace_bytes(_lib.OP_TRAIN, C.byref(m.dims), B, _lib.PRECISION_FP32)
        ws = _lib.workspace.get(nbytes, self.device)
        _lib.check(lib.ddqst_train_forward_backward(C.byref(m.dims), _lib.ptr(m.flat_params), _lib.ptr(xtp), _lib.ptr(x0p),
                                                    _lib.ptr(t32), _lib.ptr(b32), B, 1.0, _lib.ptr(grads), _lib.ptr(self._loss),
                                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        scale = 1.0
        if world > 1:
            torch.distributed.all_reduce(grads, group=process_group)       # NCCL over NVLink: the one exchange of the step
            scale = 1.0 / world
        optimizer.step(grads, grad_scale=scale)
        return self._loss

    def _train_step_tc(self, x_0, basis, optimizer: NativeAdam, row_offset: int = 0, process_group=None, world: int = 1):
        """Tensor-core form of train_step.  Nothing here depends on host-side counters (the step count that keys the
        noising stream and Adam's bias correction lives in ``optimizer.step_dev``), so the call can be captured in a
        CUDA graph (see ``make_train_graph``)."""
        lib = _lib.load()
        m = self.model
        N = m.num_qubits
        x0p = x_0 if x_0.dtype == torch.uint16 else pack_bits(x_0.to(self.device), N)
        b32 = basis if basis.dtype == torch.int32 and basis.is_cuda else basis.to(self.device).to(torch.int32).contiguous()
        B = x0p.shape[0]
        st = getattr(self, "_tc_state", None)
        if st is None or st["B"] != B or st["grads"].numel() != m.flat_params.numel() or st["grads"].device != m.flat_params.device:
            nbytes = lib.ddqst_workspace_bytes(_lib.OP_TRAIN, C.byref(m.dims), B, _lib.PRECISION_BF16)
            if nbytes < 0:
                _lib.check(-1)
            st = self._tc_state = dict(
                B=B, xt=torch.empty(B, dtype=torch.uint16, device=self.device), t=torch.empty(B, dtype=torch.int32, device=self.device),
                grads=torch.empty_like(m.flat_params), loss=torch.zeros(1, dtype=torch.float32, device=self.device),
                ws=torch.empty(nbytes, dtype=torch.uint8, device=self.device))
        shadow = m.bf16_shadow()
        _lib.check(lib.ddqst_q_sample_dev(_lib.ptr(self._q), self.num_timesteps, N, self.cumulative, _lib.ptr(x0p), None, B,
                                          row_offset, self.seed, _lib.ptr(optimizer.step_dev), _lib.ptr(st["xt"]),
                                          _lib.ptr(st["t"]), _lib.stream_ptr()))
        _lib.check(lib.ddqst_train_forward_backward_tc(C.byref(m.dims), _lib.ptr(m.flat_params), _lib.ptr(shadow), _lib.ptr(st["xt"]),
                                                       _lib.ptr(x0p), _lib.ptr(st["t"]), _lib.ptr(b32), B, 1.0, _lib.ptr(st["grads"]),
                                                       _lib.ptr(st["loss"]), _lib.ptr(st["ws"]), st["ws"].numel(), _lib.stream_ptr()))
        scale = 1.0
        if world > 1:
            torch.distributed.all_reduce(st["grads"], group=process_group)       # NCCL; capturable in a CUDA graph
            scale = 1.0 / world
        optimizer.step_device(st["grads"], grad_scale=scale, shadow=shadow)
        self._train_steps += 1
        return st["loss"]

    def make_train_graph(self, x0_packed: torch.Tensor, basis_i32: torch.Tensor, optimizer: NativeAdam,
                         row_offset: int | None = None, process_group=None, data_parallel: bool = False):
        """Capture one tensor-core training step on the STATIC device buffers ``x0_packed`` (uint16[B]) and ``basis_i32``
        (int32[B]) in a CUDA graph.  Returns a ``TrainGraph``: refill the two buffers, call ``.replay()``.
        Data parallel (``process_group`` / ``data_parallel=True``, every rank must call and replay in lockstep): the NCCL
        gradient all-reduce is captured inside the graph, so a replay is still one ``cudaGraphLaunch`` per rank.
        The caller owns the index contract for ``basis_i32`` (values in [0, num_bases)) on every replay."""
        self._require_cuda()
        if x0_packed.dtype != torch.uint16 or basis_i32.dtype != torch.int32 or not x0_packed.is_cuda or not basis_i32.is_cuda:
            raise ValueError("make_train_graph needs device tensors: x0_packed uint16[B], basis_i32 int32[B]")
        _lib.check_index(basis_i32, self.model.num_bases, "basis")
        world, rank = self._dp(process_group, data_parallel)
        if row_offset is None:
            row_offset = rank * x0_packed.shape[0] if world > 1 else 0
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2 if world > 1 else 1):     # warm-up: allocations, kernel attributes, NCCL communicator set-up
                self._train_step_tc(x0_packed, basis_i32, optimizer, row_offset, process_group, world)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        saved = (optimizer.step_count, self._train_steps, self.model.native_version)
        with torch.cuda.graph(graph):
            loss = self._train_step_tc(x0_packed, basis_i32, optimizer, row_offset, process_group, world)
        # capturing launched nothing: put the host-side counters back where the device-side state is
        optimizer.step_count, self._train_steps, self.model.native_version = saved
        self.model.mark_shadow_current()
        return TrainGraph(graph, loss, self, optimizer)
