#!/usr/bin/env python
"""Where a step of the multi-CTA line eigensolver goes (developer build: DDQST_NVCC_DEFINES=DDQST_JL_PROFILE).  clock64 stamps of
thread 0 of an interior pair group and of the mailbox-receiving group of CTA 5, summed over one PSD projection at n = 1024."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
from benchmarks.eig_large import tomography_like
lib = dq._lib.load()
raw_lib = lib
names = ["dots", "warp+group reduce", "angle", "rotate out + send", "rotate stay", "-", "receive", "loop"]
if os.environ.get("DDQST_JACOBI_LINE", "2") != "1":     # block kernel (default): its own phase list
    names = ["Gram sums + reduction + angle", "local: rotation (out to the ring, stay)", "local: __syncthreads + inbox read",
             "global step: rotation (out to the mailbox, stay)", "global step: waiting for the neighbour CTA's block", "pairs inside the blocks (per sweep)",
             "sweep barrier", "-"]
for _once in (0,):
    dim = int(os.environ.get("DIM", "1024"))
    _, rho = tomography_like(dim, 1)
    d_rho = dq.DensityMatrix(torch.from_numpy(rho).cuda())
    out = (ctypes.c_longlong * 32)()
    raw_lib.ddqst_debug_jl_profile(out)
    dq.make_positive_semidefinite(d_rho)
    raw_lib.ddqst_debug_jl_profile(out)
    for w, label in ((0, "interior group (CTA 5, group 1)"), (1, "mailbox receiver (CTA 5, group 3)")):
        v = [out[w * 16 + i] for i in range(8)]
        tot = sum(v) or 1
        print(label, "total cycles", tot)
        for nme, x in zip(names, v):
            print(f"   {nme:20s} {x:12d}  {100 * x / tot:5.1f} %")
    break
