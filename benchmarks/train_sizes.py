#!/usr/bin/env python
"""Training-step time (CUDA-graph replay) of the C4 model over batch sizes; run with DDQST_TRAIN_FUSED=0/1."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq
dq._lib.load().ddqst_debug_train_path(int(os.environ.get("DDQST_TRAIN_FUSED", "1")))
out = {"fused": os.environ.get("DDQST_TRAIN_FUSED", "1"), "bn128_batch": os.environ.get("DDQST_TC_GROUP_BN128_BATCH", "default")}
dims = (8, 6561, 100, 128, 512, 4)
for B in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192]:
    torch.manual_seed(0)
    m = dq.ConditionalD3PM(*dims).cuda()
    g = torch.Generator().manual_seed(1)
    x0p = torch.randint(0, 256, (B,), generator=g).to(torch.int32).to(torch.uint16).cuda()
    b32 = torch.randint(0, dims[1], (B,), generator=g).to(torch.int32).cuda()
    diff = dq.DiscreteDiffusion(m, 100, "cuda", seed=3, precision="bf16")
    tg = diff.make_train_graph(x0p, b32, dq.NativeAdam(m, lr=1e-3))
    for _ in range(5):
        tg.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        tg.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    flop = 3 * (dims[5] * 4 * dims[4] ** 2 + 4 * dims[4] * dims[0]) * B
    out[str(B)] = {"ms": round(ms, 4), "tflops": round(flop / ms / 1e9, 1), "loss": round(tg.loss.item(), 4), "status": dq._lib.load().ddqst_debug_tc_status()}
print(json.dumps(out))
