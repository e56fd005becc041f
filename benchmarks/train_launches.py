#!/usr/bin/env python
"""One eager tensor-core training step per batch size after warm-up, for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv python benchmarks/train_launches.py 1024 8192"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
for B in [int(a) for a in sys.argv[1:]] or [1024]:
    torch.manual_seed(0)
    m = dq.ConditionalD3PM(8, 6561, 100, 128, 512, 4).cuda()
    g = torch.Generator().manual_seed(1)
    x0 = torch.randint(0, 256, (B,), generator=g).to(torch.int32).to(torch.uint16).cuda()
    b = torch.randint(0, 6561, (B,), generator=g).to(torch.int32).cuda()
    diff = dq.DiscreteDiffusion(m, 100, "cuda", seed=3, precision="bf16")
    opt = dq.NativeAdam(m, lr=1e-3)
    for _ in range(3):
        diff.train_step(x0, b, opt, validate=False)
    torch.cuda.synchronize()
