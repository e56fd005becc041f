#!/usr/bin/env python
"""Repeatability stress test of the multi-CTA eigensolver: the rotation order and the arithmetic are deterministic, so repeated PSD
projections of the same matrix must be BITWISE identical -- a lost or torn hand-over (mailboxes, inboxes, barriers) would show up as a
differing result or a watchdog code.  Also alternates sizes so that stale mailbox contents from another size are exercised.

    python benchmarks/eig_stress.py [--reps 60]
"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
from benchmarks.eig_large import tomography_like

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=60)
args = ap.parse_args()
lib = dq._lib.load()
mats = {d: dq.DensityMatrix(torch.from_numpy(tomography_like(d, 7)[1]).cuda()) for d in (128, 256, 512, 1024)}
first, bad = {}, 0
for it in range(args.reps):
    for d, m in mats.items():
        if d == 1024 and it % 4:
            continue
        out = dq.make_positive_semidefinite(m).device_tensor()
        if d not in first:
            first[d] = out.clone()
        elif not torch.equal(out, first[d]):
            bad += 1
            print("MISMATCH", d, it, (out - first[d]).abs().max().item(), flush=True)
    st = lib.ddqst_debug_tc_status()
    if st != 0:
        print("WATCHDOG", st, it, flush=True)
        bad += 1
torch.cuda.synchronize()
print("reps", args.reps, "mismatches/watchdog", bad)
sys.exit(1 if bad else 0)
