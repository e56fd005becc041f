#!/usr/bin/env python
"""Timeline of one tensor-core training step replayed from a CUDA graph: %globaltimer stamps of every GEMM launch
(CTA 0: kernel entry, after griddepcontrol.wait, after its epilogue) -> per-kernel busy time and the gaps between kernels.

    python benchmarks/train_trace.py [--batch 1024] [--pdl 0|1]
"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--pdl", default=None)
    args = ap.parse_args()
    if args.pdl is not None:
        os.environ["DDQST_TC_PDL"] = args.pdl
    import ddqst_b200 as dq
    lib = dq._lib.load()
    lib.ddqst_debug_tc_trace.argtypes = [C.c_void_p, C.c_int32]
    dev = torch.device("cuda")
    N, NB, T = 8, 6561, 100
    torch.manual_seed(0)
    model = dq.ConditionalD3PM(N, NB, T, 128, 512, 4).to(dev)
    diff = dq.DiscreteDiffusion(model, T, dev, seed=3, precision="bf16")
    opt = dq.NativeAdam(model, lr=1e-3)
    g = torch.Generator().manual_seed(1)
    x0 = torch.randint(0, 256, (args.batch,), generator=g).to(torch.uint16).to(dev)
    b = torch.randint(0, NB, (args.batch,), generator=g).to(torch.int32).to(dev)
    diff.train_step(x0, b, opt)                                  # warm-up outside the trace
    buf = torch.zeros(4 * 64, dtype=torch.int64, device=dev)
    lib.ddqst_debug_tc_trace(buf.data_ptr(), 64)                 # capture below bakes the trace pointers into the graph
    tg = diff.make_train_graph(x0, b, opt)
    lib.ddqst_debug_tc_trace(None, 0)
    for _ in range(3):
        tg.replay()
    torch.cuda.synchronize()
    tr = buf.cpu().view(-1, 4).tolist()
    # make_train_graph ran the step twice (warm-up + capture): the captured launches are the second half of the used rows
    used = [r for r in tr if r[0] > 0]
    half = used[len(used) // 2:] if len(used) > 40 else used
    t0 = half[0][0]
    names = {0: "STORE", 1: "IN", 2: "W1", 3: "W2", 4: "HEAD", 5: "BHEAD", 6: "BW2", 7: "BW1", 8: "DCOND"}
    prev_end = None
    print(f"{'#':>3} {'epi':>6} {'entry':>9} {'wait':>7} {'busy':>7} {'gap_prev_end->ready':>20}   (us)")
    for i, (e, w, d, k) in enumerate(half):
        gap = (w - prev_end) / 1e3 if prev_end else 0.0
        print(f"{i:3d} {names.get(k, k):>6} {(e - t0) / 1e3:9.2f} {(w - e) / 1e3:7.2f} {(d - w) / 1e3:7.2f} {gap:20.2f}")
        prev_end = d
    print("first GEMM entry -> last GEMM epilogue:", (half[-1][2] - t0) / 1e3, "us;  sum busy:", sum((d - w) for _, w, d, _ in half) / 1e3, "us")


if __name__ == "__main__":
    main()
