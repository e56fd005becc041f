"""Single-qubit phase (BASELINE config C1): the notebook's call surface on top of libddqst.

``SimpleMLP`` (NB c6:65-102), ``UpgradedMLP`` (NB c12:58-94) and ``BitstringDDM`` (NB c6:106-221) with the
notebook's names, constructor arguments, ``state_dict`` keys (``time_emb.weight, basis_emb.weight, net.{0,2,..}.*``)
and methods ``forward_diffusion / train_step / sample``.  The arithmetic runs in the native fp32 kernels
(``csrc/mlp.cu``); ``train_step`` returns a loss tensor wired into autograd so the notebook's
``loss.backward(); optimizer.step()`` loop (NB c6:279-284) works unchanged.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, xp, t32, b32, *params):
        lib = _lib.load()
        B = xp.shape[0]
        ws = torch.empty(lib.ddqst_mlp_workspace_bytes(C.byref(model.dims), B), dtype=torch.uint8, device=xp.device)
        logits = torch.empty(B, 2, dtype=torch.float32, device=xp.device)
        _lib.check(lib.ddqst_mlp_forward_saved(C.byref(model.dims), _lib.ptr(model.flat_params), _lib.ptr(xp), _lib.ptr(t32),
                                               _lib.ptr(b32), B, _lib.ptr(logits), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        ctx.model, ctx.saved = model, (t32, b32, ws, B)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        model = ctx.model
        t32, b32, ws, B = ctx.saved
        grads = torch.empty_like(model.flat_params)
        _lib.check(lib.ddqst_mlp_backward_saved(C.byref(model.dims), _lib.ptr(model.flat_params), _lib.ptr(t32), _lib.ptr(b32), B,
                                                _lib.ptr(dlogits.contiguous().float()), _lib.ptr(grads), _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr()))
        return (None, None, None, None) + tuple(model.views_of(grads))


class _NotebookMLP(nn.Module):
    EMBED, HIDDEN, NUM_HIDDEN = 32, 128, 2

    def __init__(self, num_timesteps=100, num_bases=3):
        super().__init__()
        E, H = self.EMBED, self.HIDDEN
        self.num_timesteps, self.num_bases = num_timesteps, num_bases
        self.time_emb = nn.Embedding(num_timesteps + 1, E)       # same construction order as the notebook
        self.basis_emb = nn.Embedding(num_bases, E)
        layers, k = [], 1 + 2 * E
        for _ in range(self.NUM_HIDDEN):
            layers += [nn.Linear(k, H), nn.ReLU()]
            k = H
        layers.append(nn.Linear(H, 2))
        self.net = nn.Sequential(*layers)
        self.dims = _lib.Mlpdims(num_bases, num_timesteps, E, H, self.NUM_HIDDEN)
        lib = _lib.load()
        offs = (C.c_int64 * (2 + 2 * (self.NUM_HIDDEN + 1)))()
        self._total = int(lib.ddqst_mlp_param_count(C.byref(self.dims), offs))
        ordered = [self.time_emb.weight, self.basis_emb.weight]
        for m in self.net:
            if isinstance(m, nn.Linear):
                ordered += [m.weight, m.bias]
        self._layout = list(zip(ordered, [int(o) for o in offs]))
        self._flatten()

    def _flatten(self):
        dev = self.time_emb.weight.device
        flat = torch.zeros(self._total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in self._layout:
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.flat_params = flat

    def views_of(self, flat):
        by_id = {id(p): off for p, off in self._layout}
        return [flat[by_id[id(p)]:by_id[id(p)] + p.numel()].view(p.shape) for p in self.parameters()]

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flatten()
        return out

    def forward(self, noisy_x, t, basis_id):
        """noisy_x[B] int64 in {0,1}, t[B], basis_id[B] -> logits[B,2] (NB c6:86-102)."""
        if not noisy_x.is_cuda or not self.flat_params.is_cuda:
            raise RuntimeError("the notebook MLPs have no CPU path here: move the model and inputs to a B200 (cuda)")
        if not _lib.capturing():
            _lib.check_index(t, self.num_timesteps + 1, "t")
            _lib.check_index(basis_id, self.num_bases, "basis_id")
        xp = noisy_x.reshape(-1).to(torch.int32).to(torch.uint16).contiguous()
        t32 = t.to(torch.int32).contiguous()
        b32 = basis_id.to(torch.int32).contiguous()
        return _MlpFn.apply(self, xp, t32, b32, *self.parameters())


class SimpleMLP(_NotebookMLP):
    EMBED, HIDDEN, NUM_HIDDEN = 32, 128, 2


class UpgradedMLP(_NotebookMLP):
    EMBED, HIDDEN, NUM_HIDDEN = 128, 256, 3


class BitstringDDM:
    """NB c6:106-221.  ``seed`` keys the injected Philox stream (the notebook is unseeded)."""

    def __init__(self, model, num_timesteps=100, device="cuda", seed: int = 1234):
        self.model = model.to(device)
        self.num_timesteps, self.device, self.seed = num_timesteps, torch.device(device), int(seed)
        p_stay = torch.linspace(1.0, 0.5, num_timesteps + 1)
        q = torch.zeros(num_timesteps + 1, 2, 2)
        for t in range(1, num_timesteps + 1):
            p = p_stay[t]
            q[t] = torch.tensor([[p, 1 - p], [1 - p, p]])
        self.Q = q.to(self.device)
        self._q = self.Q.contiguous()
        self._sched = torch.cat([torch.zeros(num_timesteps + 1), q.reshape(-1)]).to(self.device).contiguous()
        self._calls = 0

    def forward_diffusion(self, x_0, t, stream_id: int | None = None, row_offset: int = 0):
        """x_0[B] int64, t[B] -> x_t[B] int64: one draw from column x_0 of Q[t] (NB c6:132-168)."""
        lib = _lib.load()
        if stream_id is None:
            stream_id = self._calls
            self._calls += 1
        _lib.check_index(t, self.num_timesteps + 1, "t")
        x0p = x_0.to(self.device).reshape(-1).to(torch.int32).to(torch.uint16).contiguous()
        t32 = t.to(self.device).to(torch.int32).contiguous()
        xtp = torch.empty_like(x0p)
        _lib.check(lib.ddqst_q_sample(_lib.ptr(self._q), self.num_timesteps, 1, 0, _lib.ptr(x0p), _lib.ptr(t32), x0p.shape[0],
                                      row_offset, self.seed, stream_id, _lib.ptr(xtp), None, _lib.stream_ptr()))
        return xtp.to(torch.int32).to(torch.int64)

    def train_step(self, x_0, basis_id, t=None):
        """loss = CE(model(x_t, t, basis), x_0) with t ~ U{1..T} (NB c6:170-187); differentiable."""
        x_0 = x_0.to(self.device)
        basis_id = basis_id.to(self.device)
        if t is None:
            t = torch.randint(1, self.num_timesteps + 1, (x_0.shape[0],), device=self.device)
        x_t = self.forward_diffusion(x_0, t)
        return F.cross_entropy(self.model(x_t, t, basis_id), x_0)

    def sample(self, num_samples, basis_id, shot_offset: int = 0, as_numpy: bool = True):
        """T reverse steps of "sample x0-hat, re-noise to t-1" (NB c6:189-221) -> bits[num_samples]."""
        lib = _lib.load()
        m = self.model
        _lib.check_index(int(basis_id), m.num_bases, "basis_id")
        out = torch.empty(num_samples, dtype=torch.uint8, device=self.device)
        # scratch for the 2T-row logit table (the chain itself runs in registers: one launch for all samples and steps)
        ws = _lib.workspace.get(lib.ddqst_mlp_workspace_bytes(C.byref(m.dims), 2 * self.num_timesteps) + 16 * self.num_timesteps + 1024, self.device)
        _lib.check(lib.ddqst_mlp_sample(C.byref(m.dims), _lib.ptr(m.flat_params), _lib.ptr(self._sched), int(basis_id), num_samples,
                                        shot_offset, self.seed, _lib.ptr(out), None, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        res = out.to(torch.int64)
        return res.cpu().numpy() if as_numpy else res
