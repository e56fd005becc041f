"""ConditionalD3PM: the reference's denoiser call surface on top of libddqst.

Mirrors ``ConditionalD3PM(num_qubits, num_bases, num_timesteps, embed_dim, hidden_dim, num_blocks)``
(RQC/model.py:26-70 = variant "B"; SS/model.py:42-85 = variant "A"): same constructor, same module tree and
therefore the same ``state_dict`` keys (``x_emb.weight, input_proj.*, time_emb.weight, basis_emb.weight,
blocks.{i}.film.net.*, blocks.{i}.net.{0,2}.*, output_head.*``), same default initialisation order, so a
reference checkpoint loads unchanged and ``torch.optim.Adam(model.parameters())`` keeps working.

What differs is where the arithmetic happens: every parameter is a view into ONE flat fp32 buffer whose
layout the library defines (``ddqst_param_count``), ``forward`` runs the native kernels, and autograd is
served by ``ddqst_forward_saved`` / ``ddqst_backward_saved``.  There is no CPU forward.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


class FiLM(nn.Module):
    """Parameter holder for RQC/model.py:4-11 (gamma, beta = net(cond).chunk(2))."""

    def __init__(self, cond_dim: int, feat_dim: int):
        super().__init__()
        self.net = nn.Linear(cond_dim, feat_dim * 2)


class ResBlock(nn.Module):
    """Parameter holder for RQC/model.py:13-24."""

    def __init__(self, dim: int, cond_dim: int):
        super().__init__()
        self.film = FiLM(cond_dim, dim)
        self.net = nn.Sequential(nn.Linear(dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.act = nn.SiLU()


def pack_bits(x: torch.Tensor, num_qubits: int) -> torch.Tensor:
    """int64[B,N] {0,1} (column q = qubit q) -> uint16[B] with bit q = qubit q, on the device."""
    lib = _lib.load()
    x = x.contiguous()
    if x.dtype != torch.int64:
        x = x.to(torch.int64)
    if x.dim() == 1:
        x = x.view(-1, 1)
    out = torch.empty(x.shape[0], dtype=torch.uint16, device=x.device)
    _lib.check(lib.ddqst_pack_bits(_lib.ptr(x), x.shape[0], num_qubits, _lib.ptr(out), _lib.stream_ptr()))
    return out


def unpack_bits(packed: torch.Tensor, num_qubits: int) -> torch.Tensor:
    """uint8/uint16[B] -> int64[B,N], the reference's sample layout (RQC/diffusion.py:80)."""
    lib = _lib.load()
    out = torch.empty(packed.shape[0], num_qubits, dtype=torch.int64, device=packed.device)
    _lib.check(lib.ddqst_unpack_bits(_lib.ptr(packed), packed.element_size(), packed.shape[0], num_qubits,
                                     _lib.ptr(out), _lib.stream_ptr()))
    return out


class _DenoiserFn(torch.autograd.Function):
    """logits = model(x_t, t, basis) with the native forward / backward (RQC/main.py:109-113)."""

    @staticmethod
    def forward(ctx, model, xp, t32, b32, *params):
        lib = _lib.load()
        B, N = xp.shape[0], model.num_qubits
        nbytes = lib.ddqst_workspace_bytes(_lib.OP_TRAIN, C.byref(model.dims), B, _lib.PRECISION_FP32)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xp.device)      # holds the activations until backward
        logits = torch.empty(B, N, 2, dtype=torch.float32, device=xp.device)
        _lib.check(lib.ddqst_forward_saved(C.byref(model.dims), _lib.ptr(model.flat_params), _lib.ptr(xp), _lib.ptr(t32),
                                           _lib.ptr(b32), B, _lib.ptr(logits), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        ctx.model, ctx.saved = model, (xp, t32, b32, ws)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        model = ctx.model
        xp, t32, b32, ws = ctx.saved
        grads = torch.empty_like(model.flat_params)
        _lib.check(lib.ddqst_backward_saved(C.byref(model.dims), _lib.ptr(model.flat_params), _lib.ptr(xp), _lib.ptr(t32),
                                            _lib.ptr(b32), xp.shape[0], _lib.ptr(dlogits.contiguous().float()),
                                            _lib.ptr(grads), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return (None, None, None, None) + tuple(model.views_of(grads))


class ConditionalD3PM(nn.Module):
    def __init__(self, num_qubits, num_bases, num_timesteps, embed_dim, hidden_dim, num_blocks, variant: str = "B"):
        super().__init__()
        if variant not in ("A", "B"):
            raise ValueError("variant must be 'A' (SS/model.py) or 'B' (RQC/model.py)")
        self.num_qubits, self.num_bases, self.num_timesteps = num_qubits, num_bases, num_timesteps
        self.embed_dim, self.hidden_dim, self.num_blocks, self.variant = embed_dim, hidden_dim, num_blocks, variant
        # construction order = the reference's, so torch.manual_seed(s) gives the reference's initial weights
        if variant == "B":
            self.x_emb = nn.Embedding(2, embed_dim)
            self.input_proj = nn.Linear(num_qubits * embed_dim, hidden_dim)
            self.time_emb = nn.Embedding(num_timesteps + 1, embed_dim)
            self.basis_emb = nn.Embedding(num_bases, embed_dim)
        else:
            self.time_emb = nn.Embedding(num_timesteps + 1, embed_dim)
            self.basis_emb = nn.Embedding(num_bases, embed_dim)
            self.input_proj = nn.Linear(num_qubits, hidden_dim)
        self.blocks = nn.ModuleList([ResBlock(hidden_dim, embed_dim * 2) for _ in range(num_blocks)])
        self.output_head = nn.Linear(hidden_dim, num_qubits * 2)

        self.dims = _lib.Dims(num_qubits, num_bases, num_timesteps, embed_dim, hidden_dim, num_blocks,
                              _lib.VARIANT_B if variant == "B" else _lib.VARIANT_A)
        lib = _lib.load()
        n_off = 5 + 6 * num_blocks + 2
        offs = (C.c_int64 * n_off)()
        total = lib.ddqst_param_count(C.byref(self.dims), offs)
        if total < 0:
            _lib.check(-1)
        self._total = int(total)
        self._layout = list(zip(self._ordered_params(), [int(o) for o in offs if o >= 0]))
        self._pack = None
        self._pack_key = None
        self.native_version = 0           # bumped by native in-place updates (fused Adam)
        self._flatten()

    # -- flat parameter storage ------------------------------------------------------------------
    def _ordered_params(self):
        ps = []
        if self.variant == "B":
            ps.append(self.x_emb.weight)
        ps += [self.input_proj.weight, self.input_proj.bias, self.time_emb.weight, self.basis_emb.weight]
        for blk in self.blocks:
            ps += [blk.film.net.weight, blk.film.net.bias, blk.net[0].weight, blk.net[0].bias, blk.net[2].weight, blk.net[2].bias]
        ps += [self.output_head.weight, self.output_head.bias]
        return ps

    def _flatten(self):
        dev = self.output_head.weight.device
        flat = torch.zeros(self._total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in self._layout:
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.flat_params = flat
        self._pack, self._pack_key = None, None

    def views_of(self, flat: torch.Tensor):
        """Per-parameter views of a flat buffer, in ``parameters()`` order."""
        by_id = {id(p): (off, p) for p, off in self._layout}
        return [flat[by_id[id(p)][0]:by_id[id(p)][0] + p.numel()].view(p.shape) for p in self.parameters()]

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flatten()                   # .to(device) moves parameters one by one; restore the single buffer
        return out

    # -- packed inference state ------------------------------------------------------------------
    def packed(self) -> torch.Tensor:
        """Collapsed input table, FiLM tables and bf16 weights (ddqst_pack_weights); rebuilt when parameters changed."""
        flat = self.flat_params
        key = (flat.data_ptr(), self.native_version) + tuple(p._version for p, _ in self._layout)
        if self._pack is None or self._pack_key != key:
            lib = _lib.load()
            nbytes = lib.ddqst_pack_bytes(C.byref(self.dims))
            if self._pack is None or self._pack.numel() != nbytes or self._pack.device != flat.device:
                self._pack = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
            _lib.check(lib.ddqst_pack_weights(C.byref(self.dims), _lib.ptr(flat), _lib.ptr(self._pack), _lib.stream_ptr()))
            self._pack_key = key
        return self._pack

    # -- bf16 shadow of the flat parameters (operands of the tensor-core training step) ---------
    def _param_key(self):
        return (self.flat_params.data_ptr(), self.native_version) + tuple(p._version for p, _ in self._layout)

    def bf16_shadow(self) -> torch.Tensor:
        """bf16 copy of ``flat_params`` at the same element offsets; recast when the parameters changed behind its back
        (the fused Adam keeps it current and calls ``mark_shadow_current``)."""
        flat = self.flat_params
        sh = getattr(self, "_shadow", None)
        if sh is None or sh.numel() != flat.numel() or sh.device != flat.device:
            sh = self._shadow = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device)
            self._shadow_key = None
        key = self._param_key()
        if self._shadow_key != key:
            _lib.check(_lib.load().ddqst_cast_bf16(_lib.ptr(flat), _lib.ptr(sh), flat.numel(), _lib.stream_ptr()))
            self._shadow_key = key
        return sh

    def mark_shadow_current(self):
        self._shadow_key = self._param_key()

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, x, t, basis_idx):
        """x[B,N] int64 in {0,1}, t[B] int64, basis_idx[B] int64 -> logits[B,N,2] fp32 (RQC/model.py:51-70)."""
        if not x.is_cuda or not self.flat_params.is_cuda:
            raise RuntimeError("ConditionalD3PM.forward has no CPU path: move the model and inputs to a B200 (cuda)")
        lib = _lib.load()
        if not _lib.capturing():          # nn.Embedding raises IndexError in the reference (RQC/model.py:59-61)
            _lib.check_index(t, self.num_timesteps + 1, "t")
            _lib.check_index(basis_idx, self.num_bases, "basis_idx")
        xp = pack_bits(x, self.num_qubits)
        t32 = t.to(torch.int32).contiguous()
        b32 = basis_idx.to(torch.int32).contiguous()
        B = xp.shape[0]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _DenoiserFn.apply(self, xp, t32, b32, *self.parameters())
        logits = torch.empty(B, self.num_qubits, 2, dtype=torch.float32, device=x.device)
        nbytes = lib.ddqst_workspace_bytes(_lib.OP_FORWARD, C.byref(self.dims), B, _lib.PRECISION_FP32)
        ws = _lib.workspace.get(nbytes, x.device)
        _lib.check(lib.ddqst_denoiser_forward(C.byref(self.dims), _lib.ptr(self.packed()), _lib.PRECISION_FP32, _lib.ptr(xp),
                                              _lib.ptr(t32), _lib.ptr(b32), B, _lib.ptr(logits), _lib.ptr(ws), ws.numel(),
                                              _lib.stream_ptr()))
        return logits
