#!/usr/bin/env python
"""Config C1 (single-qubit notebook model): BitstringDDM.sample throughput (one launch: 2T-row logit table + register-resident chain)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
out = {}
for cls in ("SimpleMLP", "UpgradedMLP"):
    torch.manual_seed(0)
    ddm = dq.BitstringDDM(getattr(dq, cls)(100, 3), 100, "cuda", seed=1)
    for n in (10_000, 10_000_000):
        ddm.sample(n, 1, as_numpy=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ddm.sample(n, 1, as_numpy=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        out[f"{cls}_{n}"] = {"ms": 1e3 * dt, "bitstrings_per_s": n / dt}
print(json.dumps(out))
