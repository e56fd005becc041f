"""Multi-GPU sharding of the hot path: one process per GPU, bases (or shot blocks when there are
fewer bases than ranks) split across ranks, NO collective inside sampling; one exact integer all-reduce of the
histogram counts afterwards.  Because every draw is keyed by the global (basis, shot) index, the counts are
bit-identical for any rank count."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous balanced split of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def plan(n_bases: int, n_shots: int, rank: int, world: int):
    """-> (basis_lo, basis_hi, shot_lo, shot_hi) owned by ``rank``."""
    if n_bases >= world:
        lo, hi = shard_range(n_bases, rank, world)
        return lo, hi, 0, n_shots
    # fewer bases than ranks: split ranks into groups per basis, then shots inside the group
    per = world // n_bases
    b = min(rank // per, n_bases - 1) if rank < per * n_bases else None
    if b is None:
        return 0, 0, 0, 0
    s_lo, s_hi = shard_range(n_shots, rank - b * per, per)
    return b, b + 1, s_lo, s_hi


def all_reduce_histograms(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Exact integer sum of uint32 counts over ranks (NCCL on GPU, gloo on CPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        view = hist.view(torch.int32)
        dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
    return hist


def sample_sharded(diffusion, bases, n_shots: int, group=None, reduce: bool = True, rank: int | None = None,
                   world: int | None = None):
    """Every rank samples its shard of (bases x shots) and contributes to the full uint32[len(bases), 2^N] table.
    ``rank`` / ``world`` default to the process group's; passing them explicitly (with ``reduce=False``) produces the
    contribution of any rank of any job size on this GPU."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
    elif rank is None or not 0 <= rank < world:
        raise ValueError("rank must be given with world, 0 <= rank < world")
    bases = list(bases)
    N = diffusion.model.num_qubits
    hist = torch.zeros(len(bases), 1 << N, dtype=torch.uint32, device=diffusion.device)
    b_lo, b_hi, s_lo, s_hi = plan(len(bases), n_shots, rank, world)
    if b_hi > b_lo and s_hi > s_lo:
        diffusion.sample(bases[b_lo:b_hi], s_hi - s_lo, shot_offset=s_lo, return_hist=True, hist_out=hist[b_lo:b_hi])
    return all_reduce_histograms(hist, group) if reduce else hist
