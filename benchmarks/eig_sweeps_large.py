#!/usr/bin/env python
"""Sweep counts and per-step time of the large-n eigensolver (reads JacobiCtl out of the workspace): one PSD projection at n = 512 / 1024.
DDQST_JACOBI_MIXED=0 shows the fp64-only line kernel, DDQST_JACOBI_LINE=0 the cooperative fallback."""
import os, sys, struct
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
from benchmarks.eig_large import tomography_like
lib = dq._lib.load()
for dim in (512, 1024):
    _, rho = tomography_like(dim, 1)
    raw = torch.from_numpy(rho).cuda()
    ws = torch.zeros(2 * 16 * dim * dim + 8 * dim + 1024 + 24 * dim * dim + 4096, dtype=torch.uint8, device="cuda")
    t = raw.clone()
    dq._lib.check(lib.ddqst_psd_project(dq._lib.ptr(t), dim, None, dq._lib.ptr(ws), ws.numel(), dq._lib.stream_ptr()))
    torch.cuda.synchronize()
    ctl = bytes(ws[32 * dim * dim: 32 * dim * dim + 512].cpu().numpy())
    rot = struct.unpack_from("64i", ctl, 8)
    sweeps = struct.unpack_from("i", ctl, 8 + 256)[0]
    ratio = struct.unpack_from("48f", ctl, 8 + 256 + 4)
    f32_sweeps, f32_last = struct.unpack_from("if", ctl, 8 + 256 + 4 + 192 + 4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        t.copy_(raw); dq._lib.check(lib.ddqst_psd_project(dq._lib.ptr(t), dim, None, dq._lib.ptr(ws), ws.numel(), dq._lib.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    print(f"n={dim}: psd {e0.elapsed_time(e1) / 3:.2f} ms; fp32 sweeps={f32_sweeps} (last started at ratio {np.sqrt(max(f32_last, 0)):.2e}); last phase sweeps={sweeps} rotations={rot[:sweeps]} sqrt(max_ratio2)={[float(np.sqrt(max(r, 0))) for r in ratio[:sweeps]]}", flush=True)
