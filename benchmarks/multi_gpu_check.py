#!/usr/bin/env python
"""Multi-GPU invariance check, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        benchmarks/multi_gpu_check.py

1. histograms sharded by basis (and by shot block when there are fewer bases than ranks) and combined with ONE exact
   integer all-reduce are bit-identical to the single-GPU result, for both samplers' precisions;
2. a data-parallel train step (each rank half the batch, flat-gradient all-reduce, fused Adam) reproduces the
   single-process step on the full batch.
Prints one JSON line on rank 0; exits non-zero on mismatch.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    self_group = [dist.new_group([r]) for r in range(world)][rank]      # size-1 group: the single-process reference leg
    N, T = 5, 12
    torch.manual_seed(0)
    model = dq.ConditionalD3PM(N, 3 ** N, T, 32, 128, 2).to(dev)
    out = {"world": world}
    ok = True
    for prec in ("fp32", "bf16"):
        diff = dq.DiscreteDiffusion(model, T, dev, seed=11, precision=prec)
        for bases, shots in ((list(range(40)), 300), ([7], 1001), ([3, 200, 9], 257)):
            sharded = dq.sample_sharded(diff, bases, shots)                       # every rank ends with the full table
            single, _ = diff.sample(bases, shots)                                 # the same job on one GPU
            same = torch.equal(sharded.view(torch.int32), single.view(torch.int32))
            out[f"hist_{prec}_{len(bases)}x{shots}"] = bool(same)
            ok &= same
    # ---- data-parallel training == single-process training on the concatenated batch
    g = torch.Generator().manual_seed(5)
    B = 256
    x0 = torch.randint(0, 2, (B, N), generator=g)
    basis = torch.randint(0, 3 ** N, (B,), generator=g)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref = dq.ConditionalD3PM(N, 3 ** N, T, 32, 128, 2).to(dev)
    ref.load_state_dict(sd0)
    ref_diff = dq.DiscreteDiffusion(ref, T, dev, seed=21)
    ref_opt = dq.NativeAdam(ref, lr=1e-3)
    dp_diff = dq.DiscreteDiffusion(model, T, dev, seed=21)
    dp_opt = dq.NativeAdam(model, lr=1e-3)
    lo, hi = dq.shard_range(B, rank, world)
    for step in range(3):
        # single process: the whole batch through a private (size-1) group so no exchange happens
        ref_loss = ref_diff.train_step(x0.to(dev), basis.to(dev), ref_opt)
        dp_loss = dp_diff.train_step(x0[lo:hi].to(dev), basis[lo:hi].to(dev), dp_opt, row_offset=lo, data_parallel=True)
    diffs = [(a - b).abs().max().item() for a, b in zip(ref.state_dict().values(), model.state_dict().values())]
    out["dp_train_max_param_diff"] = max(diffs)
    ok &= max(diffs) < 1e-5
    agree = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    out["ok"] = bool(agree.item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
