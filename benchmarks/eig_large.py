#!/usr/bin/env python
"""Hermitian eigensolver at n = 256 / 512 / 1024 (N = 8 / 9 / 10 qubits): PSD projection and mixed-state fidelity against
numpy (LAPACK zheevd) on tomography-like matrices, with CUDA-event timings.

    python benchmarks/eig_large.py [--dims 256 512 1024] [--reps 5] [--out profiles/r2_eig_large.json]

DDQST_JACOBI_LINE=0 keeps the round-1 cooperative kernel for n > 256 (the "before" column of the table).
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


def tomography_like(dim, seed):
    """rank-one signal + white Hermitian noise of the size linear inversion leaves at ~1e4 shots per basis"""
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    psi /= np.linalg.norm(psi)
    h = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    h = (h + h.conj().T) / 2
    rho = 0.9 * np.outer(psi, psi.conj()) + 0.1 * np.eye(dim) / dim + 0.02 * h / np.sqrt(dim)
    rho /= np.trace(rho).real
    return psi, rho


def psd_numpy(rho):
    w, v = np.linalg.eigh(rho)
    w = np.clip(w, 0, None)
    if w.sum() > 0:
        w = w / w.sum()
    return (v * w) @ v.conj().T


def fidelity_numpy(r1, r2):
    w, v = np.linalg.eigh(r1)
    s = (v * np.sqrt(np.clip(w, 0, None))) @ v.conj().T
    m = s @ r2 @ s
    ev = np.linalg.eigvalsh((m + m.conj().T) / 2)
    return float(np.sqrt(np.clip(ev, 0, None)).sum() ** 2)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs="+", default=[256, 512, 1024])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rows = []
    for dim in args.dims:
        psi, rho = tomography_like(dim, 1)
        _, rho2 = tomography_like(dim, 2)
        t0 = time.perf_counter()
        want = psd_numpy(rho)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        d_rho = dq.DensityMatrix(torch.from_numpy(rho).cuda())
        got = dq.make_positive_semidefinite(d_rho)
        psd_err = float(np.abs(got.data - want).max())
        got2 = dq.make_positive_semidefinite(dq.DensityMatrix(torch.from_numpy(rho2).cuda()))
        want2 = psd_numpy(rho2)
        f_got = dq.state_fidelity(got, got2)
        f_want = fidelity_numpy(want, want2)
        status = dq._lib.load().ddqst_debug_tc_status()
        row = {"dim": dim, "psd_max_abs_err": psd_err, "mixed_fidelity": f_got, "mixed_fidelity_numpy": f_want,
               "mixed_fidelity_err": abs(f_got - f_want), "watchdog": status,
               "psd_ms": timed(lambda: dq.make_positive_semidefinite(d_rho), args.reps),
               "mixed_fidelity_ms": timed(lambda: dq.state_fidelity(got, got2), max(1, args.reps // 2)),
               "numpy_psd_ms": cpu_ms, "line_kernel": os.environ.get("DDQST_JACOBI_LINE", "1") != "0"}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
