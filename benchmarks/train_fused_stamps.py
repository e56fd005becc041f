#!/usr/bin/env python
"""Per-GEMM clock stamps of the fused training kernel's first tile (run with DDQST_FT_DEBUG=1)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DDQST_FT_DEBUG", "1")
import ddqst_b200 as dq
dq._lib.load().ddqst_debug_train_path(1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = dq.ConditionalD3PM(8, 6561, 100, 128, 512, 4).cuda()
g = torch.Generator().manual_seed(1)
x0 = torch.randint(0, 256, (B,), generator=g).to(torch.int32).to(torch.uint16).cuda()
b = torch.randint(0, 6561, (B,), generator=g).to(torch.int32).cuda()
diff = dq.DiscreteDiffusion(m, 100, "cuda", seed=3, precision="bf16")
opt = dq.NativeAdam(m, lr=1e-3)
for _ in range(3):
    diff.train_step(x0, b, opt, validate=False)
buf = (C.c_longlong * 256)()
dq._lib.check(dq._lib.load().ddqst_debug_ft_stamps(buf))
v = list(buf)
t0 = min(x for x in v if x > 0)
n = 19
dur=[v[2*g+1]-v[2*g] for g in range(19)]
print("sweep durations", dur, "total", max(v[:40])-t0)
print("g   epi_first epi_end | mma_first mma_end   (cycles since first stamp)")
for gi in range(n):
    print(gi, v[2 * gi] - t0, v[2 * gi + 1] - t0, "|", v[100 + 2 * gi] - t0 if v[100 + 2 * gi] else None, v[101 + 2 * gi] - t0 if v[101 + 2 * gi] else None)
