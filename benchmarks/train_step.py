#!/usr/bin/env python
"""Training-step timings (T1, RQC/main.py:105-115) at the C4 architecture: the fp32 CUDA-core step, the tensor-core
(tcgen05 bf16) step launched eagerly, and the same step replayed from a CUDA graph, at per-GPU batch 1024
(RQC/config.py:15) and 8192, next to the reference algorithm (oracle port, torch CPU) on the host cores.

Roofline: tensor.  Algorithmic work = 3 x 4 210 688 FLOP per sample (forward + dgrad + wgrad of the irreducible
square GEMMs + head); literal work (FiLM and input GEMMs included) = 3 x 7 356 416.

    python benchmarks/train_step.py [--out profiles/r1_train_step.json] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq                       # noqa: E402

ALG_FLOP = 3 * 4_210_688
LIT_FLOP = 3 * 7_356_416


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--batches", default="1024,8192")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--modes", default="fp32,bf16,bf16_graph")
    args = ap.parse_args()
    dev = torch.device("cuda")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    N, NB, T, E, H, L = 8, 6561, 100, 128, 512, 4
    rows = []
    for B in [int(v) for v in args.batches.split(",")]:
        g = torch.Generator().manual_seed(1)
        x0 = torch.randint(0, 2, (B, N), generator=g)
        basis = torch.randint(0, NB, (B,), generator=g)
        x0p = dq.pack_bits(x0.to(dev), N)
        b32 = basis.to(dev).to(torch.int32)
        res = {"batch": B}
        for mode in args.modes.split(","):
            torch.manual_seed(0)
            model = dq.ConditionalD3PM(N, NB, T, E, H, L).to(dev)
            diff = dq.DiscreteDiffusion(model, T, dev, seed=3, precision="bf16")
            opt = dq.NativeAdam(model, lr=1e-3)
            if mode == "bf16_graph":
                tg = diff.make_train_graph(x0p, b32, opt)
                ms = timed(tg.replay, args.iters)
                loss = tg.loss.item()
            else:
                prec = "fp32" if mode == "fp32" else "bf16"
                ms = timed(lambda: diff.train_step(x0p, b32, opt, precision=prec), args.iters)
                loss = diff.train_step(x0p, b32, opt, precision=prec).item()
            assert dq._lib.load().ddqst_debug_tc_status() == 0
            res[mode] = {"ms": ms, "samples_per_s": B / ms * 1e3, "alg_tflops": ALG_FLOP * B / ms / 1e9,
                         "literal_tflops": LIT_FLOP * B / ms / 1e9, "frac_of_bf16_peak": ALG_FLOP * B / ms / 1e9 / peak,
                         "loss_after": loss}
        if not args.no_cpu and B == 1024:
            from oracle import ddqst_oracle as orc        # CPU baseline leg only
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            torch.manual_seed(0)
            ref = dq.ConditionalD3PM(N, NB, T, E, H, L)
            params = {k: v.detach().clone().requires_grad_(True) for k, v in ref.state_dict().items()}
            copt = torch.optim.Adam(list(params.values()), lr=1e-3)
            _, q_bar = orc.cosine_schedule(T)
            orc.train_step(params, copt, q_bar, x0, basis, N, T, 3, 0)
            t0 = time.perf_counter()
            for s in range(5):
                orc.train_step(params, copt, q_bar, x0, basis, N, T, 3, s + 1)
            res["cpu_port"] = {"ms": 1e3 * (time.perf_counter() - t0) / 5, "cores": threads}
        rows.append(res)
        print(json.dumps(res))
    out = {"what": "train step, C4 architecture (N=8 E=128 H=512 L=4, 6561 bases), Adam", "peak_tflops": peak,
           "peak_source": "measured bf16_tflops (burst)" if peaks else "fallback", "rows": rows}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
