#!/usr/bin/env python
"""The whole generative-tomography pass on the native path, timed phase by phase (what RQC/main.py + evaluate.py do for one
state): synthetic random-circuit state -> measured counts in all 3^N bases -> QuantumStateDataset -> K training steps
(tensor-core step replayed from a CUDA graph, batches gathered on the device) -> sample(all bases, shots) -> linear
inversion + PSD -> fidelity against the clean state, next to linear inversion of the raw counts.

    python benchmarks/pipeline.py --n 3 --steps 3000            # small: the model visibly learns the state
    python benchmarks/pipeline.py --n 8 --steps 2000 --shots-infer 4096 --embed 128 --hidden 512 --blocks 4   # C4 shapes
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3)
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--shots-train", type=int, default=100_000)
    ap.add_argument("--shots-infer", type=int, default=10_000)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--embed", type=int, default=64)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--blocks", type=int, default=3)
    ap.add_argument("--state", default="rqc", help="rqc | ghz | bell | plus (SS/data_gen.py state types)")
    ap.add_argument("--variant", default="B", help="B = RQC model (x_emb front end, cosine schedule, posterior sampler); "
                                                   "A = SS model (Linear(N,H) front end, linear schedule, re-noise sampler, AdamW)")
    ap.add_argument("--noise", default="readout")
    ap.add_argument("--error-rate", type=float, default=0.02)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--train-precision", default=None, help="bf16 (tensor cores, graph replay) or fp32 (exact CUDA-core step)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda")
    N, NB = args.n, 3 ** args.n
    res = {"config": vars(args)}

    def timed(label, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        res[label + "_s"] = time.perf_counter() - t0
        return out

    # 1. data: clean state + noisy measurement counts in every basis (stands in for the Aer generation)
    psi = dq.synth_state(N, args.state, args.depth, args.seed, dev)
    hist = timed("generate_counts", lambda: dq.born_histograms(psi, N, args.shots_train, args.seed, None, args.noise, args.error_rate))
    clean = dq.born_histograms(psi, N, args.shots_train, args.seed + 1)
    # 2. dataset (the reference's record format, then the device-side table)
    records = [{"depth": args.depth, "measurements": dq.counts_records(hist, N)}] if NB <= 2187 else None
    ds = timed("dataset_build", lambda: dq.QuantumStateDataset(records, N, device=dev, seed=args.seed) if records is not None
               else dq.QuantumStateDataset.from_counts_table(hist, N, device=dev, seed=args.seed))
    res["train_shots"] = len(ds)
    # 3. training
    torch.manual_seed(args.seed)
    model = dq.ConditionalD3PM(N, NB, args.T, args.embed, args.hidden, args.blocks, variant=args.variant).to(dev)
    diff = dq.DiscreteDiffusion(model, args.T, dev, schedule="cosine" if args.variant == "B" else "linear", seed=args.seed,
                                precision="bf16")
    opt = dq.NativeAdam(model, lr=1e-3) if args.variant == "B" else dq.NativeAdam(model, lr=1e-4, weight_decay=0.01, decoupled=True)
    x0_s, b_s = ds.batch(0, args.batch)
    prec = args.train_precision or diff.train_precision()
    graph = diff.make_train_graph(x0_s, b_s, opt) if prec == "bf16" else None
    losses = []

    def train():
        for step in range(1, args.steps + 1):
            x0, basis = ds.batch(step, args.batch)
            if graph is not None:
                x0_s.copy_(x0); b_s.copy_(basis)
                loss = graph.replay()
            else:
                loss = diff.train_step(x0, basis, opt, precision="fp32")
            if step % max(1, args.steps // 10) == 0:
                losses.append(float(loss.item()))
    timed("train", train)
    res["train_step_ms"] = 1e3 * res["train_s"] / max(1, args.steps)
    res["losses"] = losses
    # 4. generation + reconstruction
    syn = timed("sample", lambda: diff.sample(list(range(NB)), args.shots_infer)[0])
    res["bitstrings_per_s"] = NB * args.shots_infer / res["sample_s"]
    rho = timed("reconstruct", lambda: dq.linear_inversion(syn, N))
    res["fidelity_d3pm"] = timed("fidelity", lambda: dq.state_fidelity(psi, rho))
    res["fidelity_raw_noisy_counts"] = dq.state_fidelity(psi, dq.linear_inversion(hist, N))
    res["fidelity_raw_clean_counts"] = dq.state_fidelity(psi, dq.linear_inversion(clean, N))
    res["z_bias"] = dq.calculate_z_bias(syn, N)
    assert dq._lib.load().ddqst_debug_tc_status() == 0
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
