#!/usr/bin/env python
"""Sweep counts of the LAST eigensolve of a mixed-state fidelity (the strict-threshold one on sqrt(a) b sqrt(a)) at n = 512 / 1024,
read out of the workspace's JacobiCtl, plus the time of the whole fidelity call."""
import os, sys, struct
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
from benchmarks.eig_large import tomography_like, psd_numpy
for dim in (512, 1024):
    _, r1 = tomography_like(dim, 1)
    _, r2 = tomography_like(dim, 2)
    a = dq.DensityMatrix(torch.from_numpy(psd_numpy(r1)).cuda())
    b = dq.DensityMatrix(torch.from_numpy(psd_numpy(r2)).cuda())
    f = dq.state_fidelity(a, b)
    torch.cuda.synchronize()
    ws = dq._lib.workspace.get(1, "cuda")
    ctl = bytes(ws[64 * dim * dim: 64 * dim * dim + 512].cpu().numpy())
    rot = struct.unpack_from("64i", ctl, 8)
    sweeps = struct.unpack_from("i", ctl, 8 + 256)[0]
    ratio = struct.unpack_from("48f", ctl, 8 + 256 + 4)
    f32_sweeps, f32_last = struct.unpack_from("if", ctl, 8 + 256 + 4 + 192 + 4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dq.state_fidelity(a, b)
    e1.record(); torch.cuda.synchronize()
    print(f"n={dim}: F={f:.9f} fidelity {e0.elapsed_time(e1) / 3:.2f} ms; last solve: fp32 sweeps={f32_sweeps} (last started at {np.sqrt(max(f32_last, 0)):.2e}); fp64 sweeps={sweeps} "
          f"sqrt(max_ratio2)={[float(f'{np.sqrt(max(r, 0)):.2e}') for r in ratio[:sweeps]]}", flush=True)
