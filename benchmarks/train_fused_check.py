#!/usr/bin/env python
"""Fused forward + data-gradient training kernel (csrc/train_fused.cuh) against the fp32 CUDA-core step and the per-layer
tensor-core path: per-parameter relative L2 gradient error, loss, and step time (CUDA-graph replay).
    python benchmarks/train_fused_check.py [--time]"""
import ctypes as C
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def grads(dq, m, x0p, xtp, t32, b32, tc):
    lib = dq._lib.load()
    B = x0p.shape[0]
    g = torch.full_like(m.flat_params, float("nan"))
    loss = torch.zeros(1, device="cuda")
    prec = dq._lib.PRECISION_BF16 if tc else dq._lib.PRECISION_FP32
    ws = torch.empty(lib.ddqst_workspace_bytes(dq._lib.OP_TRAIN, C.byref(m.dims), B, prec), dtype=torch.uint8, device="cuda")
    P = dq._lib.ptr
    if tc:
        dq._lib.check(lib.ddqst_train_forward_backward_tc(C.byref(m.dims), P(m.flat_params), P(m.bf16_shadow()), P(xtp), P(x0p), P(t32),
                                                          P(b32), B, 1.0, P(g), P(loss), P(ws), ws.numel(), dq._lib.stream_ptr()))
    else:
        dq._lib.check(lib.ddqst_train_forward_backward(C.byref(m.dims), P(m.flat_params), P(xtp), P(x0p), P(t32), P(b32), B, 1.0, P(g),
                                                       P(loss), P(ws), ws.numel(), dq._lib.stream_ptr()))
    torch.cuda.synchronize()
    return loss.item(), g, lib.ddqst_debug_tc_status()


def main():
    import ddqst_b200 as dq
    mode = int(os.environ.get("DDQST_TRAIN_FUSED", "1"))
    dq._lib.load().ddqst_debug_train_path(mode)
    out = {"fused": mode}
    for name, dims, B in (("h128", (4, 81, 50, 32, 128, 2), 300), ("h256_ragged", (4, 81, 50, 32, 256, 2), 6200),
                          ("c4", (8, 6561, 100, 128, 512, 4), 1024), ("c4_b4200", (8, 6561, 100, 128, 512, 4), 4200), ("c5", (10, 59049, 100, 128, 512, 4), 1000),
                          ("variantA", (3, 27, 100, 64, 512, 4), 256)):
        torch.manual_seed(0)
        variant = "A" if name == "variantA" else "B"
        m = dq.ConditionalD3PM(*dims, variant=variant).cuda()
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.02 * torch.randn_like(p))
        m.native_version += 1
        N, NB, T = dims[0], dims[1], dims[2]
        g = torch.Generator().manual_seed(1)
        x0p = torch.randint(0, 1 << N, (B,), generator=g).to(torch.int32).to(torch.uint16).cuda()
        xtp = torch.randint(0, 1 << N, (B,), generator=g).to(torch.int32).to(torch.uint16).cuda()
        t32 = torch.randint(1, T + 1, (B,), generator=g).to(torch.int32).cuda()
        b32 = torch.randint(0, NB, (B,), generator=g).to(torch.int32).cuda()
        l32, g32, _ = grads(dq, m, x0p, xtp, t32, b32, False)
        ltc, gtc, st = grads(dq, m, x0p, xtp, t32, b32, True)
        rel = {}
        for (pn, _), a, b in zip(m.named_parameters(), m.views_of(g32), m.views_of(gtc)):
            rel[pn] = round((a - b).norm().item() / max(a.norm().item(), 1e-12), 5)
        worst = max(rel.values())
        out[name] = {"loss_fp32": l32, "loss_tc": ltc, "tc_status": st, "worst_rel_grad": worst,
                     "bad": {k: v for k, v in rel.items() if v > 3e-2 or v != v}}
        if "--verbose" in sys.argv:
            out[name]["rel"] = rel
    if "--time" in sys.argv:
        for name, dims, B in (("c4_b1024", (8, 6561, 100, 128, 512, 4), 1024), ("c4_b8192", (8, 6561, 100, 128, 512, 4), 8192),
                              ("c5_b1024", (10, 59049, 100, 128, 512, 4), 1024)):
            torch.manual_seed(0)
            m = dq.ConditionalD3PM(*dims).cuda()
            N, NB, T = dims[0], dims[1], dims[2]
            g = torch.Generator().manual_seed(1)
            x0p = torch.randint(0, 1 << N, (B,), generator=g).to(torch.int32).to(torch.uint16).cuda()
            b32 = torch.randint(0, NB, (B,), generator=g).to(torch.int32).cuda()
            diff = dq.DiscreteDiffusion(m, T, "cuda", seed=3, precision="bf16")
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
            flop = 3 * (dims[5] * 4 * dims[4] ** 2 + 4 * dims[4] * N) * B
            out[name] = {"ms": ms, "tflops": flop / ms / 1e9, "loss": tg.loss.item(), "tc_status": dq._lib.load().ddqst_debug_tc_status()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
