#!/usr/bin/env python
"""How many sweeps the Jacobi eigensolver takes on a typical linear-inversion rho (reads JacobiCtl out of the workspace)."""
import os, sys, struct
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
lib = dq._lib.load()
for N, shots in ((6, 100000), (8, 1000000), (8, 10000)):
    psi = dq.synth_state(N, "rqc", depth=16, seed=1)
    h = dq.born_histograms(psi, N, shots, seed=1)
    raw = dq.linear_inversion_raw(h, N)
    dim = 1 << N
    extra = 0 if os.environ.get('NO_EXTRA') else 24 * dim * dim
    ws = torch.zeros(2 * 16 * dim * dim + 8 * dim + 1024 + extra, dtype=torch.uint8, device="cuda")
    t = raw.clone()
    dq._lib.check(lib.ddqst_psd_project(dq._lib.ptr(t), dim, None, dq._lib.ptr(ws), ws.numel(), dq._lib.stream_ptr()))
    torch.cuda.synchronize()
    ctl = bytes(ws[32 * dim * dim: 32 * dim * dim + 512].cpu().numpy())
    sigma = struct.unpack_from("d", ctl, 0)[0]
    rot = struct.unpack_from("64i", ctl, 8)
    sweeps = struct.unpack_from("i", ctl, 8 + 256)[0]
    ratio = struct.unpack_from("48f", ctl, 8 + 256 + 4)
    want = np.linalg.eigvalsh(raw.cpu().numpy())
    ev = torch.empty(dim, dtype=torch.float64, device="cuda")
    t2 = raw.clone()
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        t2.copy_(raw); dq._lib.check(lib.ddqst_psd_project(dq._lib.ptr(t2), dim, dq._lib.ptr(ev), dq._lib.ptr(ws), ws.numel(), dq._lib.stream_ptr()))
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    pos = np.maximum(want, 0); pos = pos / pos.sum()
    err = np.abs(np.sort(ev.cpu().numpy()) - np.sort(pos)).max()
    print(f"   lambda_min={want.min():.4f} lambda_max={want.max():.4f}  psd ms={ms:.3f}  max|eval err|={err:.2e}")
    print(f"N={N} shots={shots}: sigma={sigma:.3f} sweeps={sweeps} rotations={rot[:sweeps]} sqrt(max_ratio2)={[float(np.sqrt(max(r, 0))) for r in ratio[:sweeps]]}")
