#!/usr/bin/env python
"""Timings for the other two pieces of the path next to the reference algorithm on the host cores:

  * training step (RQC/main.py:105-115) at the C4 architecture, batch 1024: fused native step vs the oracle port (torch CPU);
  * reconstruction + fidelity (RQC/reconstruct.py:56-67 + state_fidelity): native hist -> rho -> PSD -> F vs the reference's
    literal 4^N loop at the reference's own scale (10 000 shots per basis), N = 5 and 6 (the literal loop is
    O(4^N (shots N + 4^N)); the survey measured 146 s at N = 8).

    python benchmarks/path_extras.py [--out profiles/r1_path_extras.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq                       # noqa: E402
from oracle import ddqst_oracle as orc        # noqa: E402  (CPU baseline leg only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cpu-steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda")
    rows = []
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)

    # ---------------- training step, C4 architecture ----------------
    N, NB, T, E, H, L, B = 8, 6561, 100, 128, 512, 4, 1024
    torch.manual_seed(0)
    model = dq.ConditionalD3PM(N, NB, T, E, H, L)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    diff = dq.DiscreteDiffusion(model, T, dev, seed=3)
    opt = dq.NativeAdam(model, lr=1e-3)
    g = torch.Generator().manual_seed(1)
    x0 = torch.randint(0, 2, (B, N), generator=g)
    basis = torch.randint(0, NB, (B,), generator=g)
    x0d, bd = dq.pack_bits(x0.to(dev), N), basis.to(dev)
    for _ in range(3):
        diff.train_step(x0d, bd, opt)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        loss = diff.train_step(x0d, bd, opt)
    b.record()
    torch.cuda.synchronize()
    gpu_ms = a.elapsed_time(b) / 20
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    copt = torch.optim.Adam(list(params.values()), lr=1e-3)
    _, q_bar = orc.cosine_schedule(T)
    orc.train_step(params, copt, q_bar, x0, basis, N, T, 3, 0)
    t0 = time.perf_counter()
    for s in range(args.cpu_steps):
        orc.train_step(params, copt, q_bar, x0, basis, N, T, 3, s + 1)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / args.cpu_steps
    rows.append({"what": "train step C4 (N=8,E=128,H=512,L=4), batch 1024, fp32", "gpu_ms": gpu_ms, "cpu_ms": cpu_ms,
                 "cpu_cores": threads, "cpu_kind": "port", "speedup": cpu_ms / gpu_ms, "loss": float(loss.item())})
    print(rows[-1])

    # ---------------- reconstruction + fidelity at the reference's scale ----------------
    rng = np.random.default_rng(0)
    for n in (5, 6):
        psi = orc.haar_state(n, n)
        probs = orc.born_probabilities_all(psi, n)
        shots = 10_000
        hist = rng.multinomial(shots, probs).astype(np.int32)
        h_d = torch.from_numpy(hist).to(dev)
        psi_d = torch.from_numpy(psi).to(dev)
        for _ in range(2):
            f_gpu = dq.state_fidelity(psi_d, dq.linear_inversion(h_d, n))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f_gpu = dq.state_fidelity(psi_d, dq.linear_inversion(h_d, n))
        gpu_ms = 1e3 * (time.perf_counter() - t0)
        names = orc.basis_strings(n)
        data = {name: ((np.repeat(np.arange(1 << n), hist[i])[:, None] >> np.arange(n)) & 1) for i, name in enumerate(names)}
        t0 = time.perf_counter()
        rho_cpu = orc.linear_inversion_literal(data, n)
        f_cpu = orc.state_fidelity(psi, rho_cpu)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        rows.append({"what": f"recon+fidelity N={n}, {shots} shots/basis (hist -> rho -> PSD -> F)", "gpu_ms": gpu_ms, "cpu_ms": cpu_ms,
                     "cpu_kind": "port (literal 4^N loop)", "speedup": cpu_ms / gpu_ms, "fidelity_gpu": f_gpu, "fidelity_cpu": f_cpu,
                     "abs_diff": abs(f_gpu - f_cpu)})
        print(rows[-1])
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
