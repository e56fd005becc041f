#!/usr/bin/env python
"""Device-time breakdown of reconstruction + fidelity at N qubits: hist -> rho (two WHT passes) -> PSD projection
(Jacobi eigensolver + clip + rebuild) -> <psi|rho|psi>, each timed alone with CUDA events.

    python benchmarks/recon_breakdown.py [--n 8] [--out profiles/...json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq                       # noqa: E402
from oracle import ddqst_oracle as orc        # noqa: E402  (synthetic state only)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="4,6,8")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rows = []
    for n in [int(v) for v in args.n.split(",")]:
        rng = np.random.default_rng(n)
        psi = orc.haar_state(n, n)
        probs = orc.born_probabilities_all(psi, n)
        hist = torch.from_numpy(rng.multinomial(100_000, probs).astype(np.int32)).cuda()
        psi_d = torch.from_numpy(psi).cuda()
        t_li, raw = timed(lambda: dq.linear_inversion_raw(hist, n))
        t_psd, rho = timed(lambda: dq.make_positive_semidefinite(raw))
        t_f, f = timed(lambda: dq.state_fidelity(psi_d, rho))
        t_all, f2 = timed(lambda: dq.state_fidelity(psi_d, dq.linear_inversion(hist, n)))
        t_met, met = timed(lambda: dq.get_metrics(rho, n))
        rows.append({"n_qubits": n, "linear_inversion_ms": t_li, "psd_project_ms": t_psd, "fidelity_pure_ms": t_f,
                     "chain_ms": t_all, "get_metrics_ms": t_met, "fidelity": float(f2)})
        print(json.dumps(rows[-1]))
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
