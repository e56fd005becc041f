#!/usr/bin/env python
"""Roofline numbers for the HBM-bound kernels of the path (north_star: "achieved HBM GB/s against the B200 peak for
the histogram, reconstruction and noising kernels").  Each kernel is called through the C ABI on inputs larger than
L2 (or, for the small reconstruction sizes, its natural size), timed with CUDA events after warm-up, and reported as
algorithmic bytes / time against MEASURED_PEAKS.json hbm_gbs.

    python benchmarks/hbm_kernels.py [--out profiles/r1_hbm_kernels.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq  # noqa: E402


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only-hist", action="store_true")
    args = ap.parse_args()
    lib = dq._lib.load()
    P, S = dq._lib.ptr, dq._lib.stream_ptr
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except OSError:
        peak, src = 6650.0, "fallback"
    dev = torch.device("cuda")
    rows = []

    def report(name, bytes_algo, ms, note=""):
        gbs = bytes_algo / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "algorithmic_bytes": int(bytes_algo), "ms": ms, "achieved_gbs": gbs, "peak_gbs": peak,
                     "frac": gbs / peak, "peak_source": src, "note": note})
        print(f"{name:34s} {bytes_algo / 1e6:10.1f} MB {ms:9.4f} ms {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of {src} HBM peak  {note}")

    # ---- histogram (H0): N=8, 1 B/shot
    n = 1 << 28
    g = torch.Generator(device=dev).manual_seed(0)
    data = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev, generator=g)
    hist = torch.zeros(256, dtype=torch.uint32, device=dev)
    ms = timed(lambda: dq._lib.check(lib.ddqst_histogram(P(data), 1, n, 8, P(hist), S())))
    report("histogram N=8 uniform", n + 1024, ms, "268M shots, uint8")
    peaked = torch.where(torch.rand(n, device=dev) < 0.9, torch.tensor(37, dtype=torch.uint8, device=dev), data)
    ms = timed(lambda: dq._lib.check(lib.ddqst_histogram(P(peaked), 1, n, 8, P(hist), S())))
    report("histogram N=8 peaked (90% one bin)", n + 1024, ms, "warp-aggregated atomics")
    n16 = 1 << 27
    d16 = torch.randint(0, 1024, (n16,), dtype=torch.int32, device=dev, generator=g).to(torch.uint16)
    h10 = torch.zeros(1024, dtype=torch.uint32, device=dev)
    ms = timed(lambda: dq._lib.check(lib.ddqst_histogram(P(d16), 2, n16, 10, P(h10), S())))
    report("histogram N=10 uniform", 2 * n16 + 4096, ms, "134M shots, uint16")
    del data, peaked, d16
    if args.only_hist:
        return

    # ---- noising (D2): x0 uint16 in, t drawn in-kernel (int32 out), x_t uint16 out
    B = 1 << 27
    x0 = torch.zeros(B, dtype=torch.uint16, device=dev)
    xt = torch.empty_like(x0)
    tt = torch.empty(B, dtype=torch.int32, device=dev)
    _, q = dq.cosine_schedule(100)
    q = q.to(dev).contiguous()
    ms = timed(lambda: dq._lib.check(lib.ddqst_q_sample(P(q), 100, 8, 1, P(x0), None, B, 0, 1234, 7, P(xt), P(tt), S())))
    report("q_sample N=8 (+t draw)", B * (2 + 2 + 4), ms, "134M samples")
    # ---- unpack to the reference's int64[B,N] layout
    Bq = 1 << 25
    pk = torch.zeros(Bq, dtype=torch.uint8, device=dev)
    out = torch.empty(Bq, 8, dtype=torch.int64, device=dev)
    ms = timed(lambda: dq._lib.check(lib.ddqst_unpack_bits(P(pk), 1, Bq, 8, P(out), S())))
    report("unpack_bits N=8 -> int64[B,8]", Bq * (1 + 64), ms, "33M shots")
    del x0, xt, tt, pk, out

    # ---- Adam: 28 B/param
    npar = 1 << 26
    p = torch.zeros(npar, device=dev)
    gr = torch.ones(npar, device=dev)
    m = torch.zeros(npar, device=dev)
    v = torch.zeros(npar, device=dev)
    ms = timed(lambda: dq._lib.check(lib.ddqst_adam_step(P(p), P(gr), P(m), P(v), npar, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 1.0, S())))
    report("adam_step", 28 * npar, ms, "67M params")
    del p, gr, m, v

    # ---- linear inversion (R1-R3) and fidelity at N=8 and N=10
    for N in (8, 10):
        nb, dim = 3 ** N, 1 << N
        h = torch.randint(0, 1000, (nb, dim), dtype=torch.int32, device=dev, generator=g).view(torch.uint32)
        shots = h.view(torch.int32).to(torch.int64).sum(dim=1)
        rho = torch.empty(dim, dim, dtype=torch.complex128, device=dev)
        ws = torch.empty(max(nb * dim * 4, 8 * dim * dim) + 256, dtype=torch.uint8, device=dev)
        ms = timed(lambda: dq._lib.check(lib.ddqst_linear_inversion(P(h), P(shots), nb, N, None, 0, P(rho), P(ws), ws.numel(), S())),
                   iters=5)
        report(f"linear_inversion N={N}", 4 * nb * dim + 16 * dim * dim, ms, "hist -> rho (two WHT passes)")
        psi = torch.randn(dim, dtype=torch.complex128, device=dev)
        o = torch.zeros(1, dtype=torch.float64, device=dev)
        ms = timed(lambda: dq._lib.check(lib.ddqst_fidelity_pure(P(psi), P(rho), dim, P(o), S())))
        report(f"fidelity_pure N={N}", 16 * dim * dim, ms)
        r2 = torch.eye(dim, dtype=torch.complex128, device=dev) / dim + 1e-3 * torch.randn(dim, dim, dtype=torch.complex128, device=dev)
        r2 = (r2 + r2.conj().T) / 2
        wsp = torch.empty(2 * 16 * dim * dim + 8 * dim + 1024 + 24 * dim * dim + 4096, dtype=torch.uint8, device=dev)   # + the mixed-precision scratch
        work = r2.clone()

        def psd():
            work.copy_(r2)
            dq._lib.check(lib.ddqst_psd_project(P(work), dim, None, P(wsp), wsp.numel(), S()))
        ms = timed(psd, iters=3, warmup=1)
        report(f"psd_project N={N} (Jacobi)", 32 * dim * dim, ms, "latency-bound eigensolver; bytes = rho in + out")
        del h, ws

    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
