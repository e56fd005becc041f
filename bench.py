#!/usr/bin/env python
"""bench.py -- generated bitstrings/sec of the T-step D3PM reverse sampler at BASELINE.json's C4 shape
(N=8 qubits, 3^8 bases, T=100, E=128, H=512, L=4, 10^6 shots/basis), plus recon+fidelity milliseconds.

A step = one launch of the persistent tcgen05 sampler over `--bases-per-step` measurement bases x `--shots`
shots (a slice of the 6561-basis job; every step takes the next bases), histogram fused.  Weak scaling: every
rank samples its own bases, no collective in the timed region.  One JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W]                  # this repo's CUDA path
  python bench.py --impl reference ...                                  # the reference algorithm on host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_STEP = 4_210_688          # 8 square 512x512 GEMMs + 512x16 head, per sample per step
C4 = dict(N=8, NB=6561, T=100, E=128, H=512, L=4)
# dram__bytes_read + dram__bytes_write of one sampler launch (ncu --set full)
SAMPLER_DRAM_BYTES_PER_LAUNCH = 6_276_096   # profiles/r1_v3c_sampler_ncu_summary.txt (weights + tables, read once per launch)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--bases-per-step", type=int, default=8)
    ap.add_argument("--shots", type=int, default=1_000_000)
    ap.add_argument("--cpu-shots", type=int, default=4000, help="shots per basis of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except OSError:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


def reference_state_dict(seed=0):
    """Random-init weights of the C4 architecture (torch default init, same order as RQC/model.py:27-49)."""
    import torch
    import ddqst_b200 as dq
    torch.manual_seed(seed)
    m = dq.ConditionalD3PM(C4["N"], C4["NB"], C4["T"], C4["E"], C4["H"], C4["L"])
    return m


def cpu_port_rate(sd, shots, bases, seed, threads):
    """The reference algorithm (oracle port of RQC/diffusion.py:53-80, torch fp32 on the host cores)."""
    import torch
    from oracle import ddqst_oracle as orc
    torch.set_num_threads(threads)
    betas, q_bar = orc.cosine_schedule(C4["T"])
    t0 = time.perf_counter()
    for b in bases:
        orc.p_sample_posterior(sd, betas, q_bar, shots, b, C4["N"], seed)
    dt = time.perf_counter() - t0
    return shots * len(bases) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    sd = {k: v.detach() for k, v in reference_state_dict().state_dict().items()}
    shots = min(args.cpu_shots, 10_000)
    for w in range(min(args.warmup, 1)):
        cpu_port_rate(sd, 256, [0], args.seed, threads)
    times = []
    for k in range(args.steps):
        rate, dt = cpu_port_rate(sd, shots, [k % C4["NB"]], args.seed, threads)
        times.append(dt)
    total = sum(times)
    value = shots * args.steps / total
    line = {
        "impl": "reference", "metric": "generated bitstrings/sec (T-step D3PM, N=8)", "value": value, "unit": "bitstrings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C4 slice: N=8 T=100 E=128 H=512 L=4, 1 of 6561 bases x {shots} shots per step "
                               "(bounded sample of the same job; reference algorithm restated in oracle/, torch CPU)"},
        "cpu_baseline": {"value": value, "unit": "bitstrings/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x p_sample({shots}, basis, 8), T=100"},
        "e2e": {"value": value, "unit": "bitstrings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import ddqst_b200 as dq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = dq._lib.load()
    N, NB, T = C4["N"], C4["NB"], C4["T"]

    model = reference_state_dict().to(dev)
    sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    diff = dq.DiscreteDiffusion(model, T, dev, seed=args.seed, precision="bf16")
    model.packed()
    bps, shots = args.bases_per_step, args.shots
    units_per_step = bps * shots                                      # per rank

    def step_bases(i):
        start = ((i * world + rank) * bps) % NB
        return [(start + j) % NB for j in range(bps)]

    hist = torch.zeros(bps, 1 << N, dtype=torch.uint32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def device_step(i):
        hist.zero_()
        diff.sample(step_bases(i), shots, return_hist=True, hist_out=hist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    for i in range(args.warmup):
        device_step(i)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    evs = []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                                         # L2 flush between timed steps (untimed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        device_step(args.warmup + i)
        b.record()
        evs.append((a, b))
    barrier()
    clock_info = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = total_ms.item()
    assert lib.ddqst_debug_tc_status() == 0, "tcgen05 pipeline timed out"
    counts = hist.view(torch.int32).sum(dim=1).cpu().tolist()
    assert all(c == shots for c in counts), counts
    value = world * units_per_step * args.steps / (total_ms / 1e3)

    # ---------------- end to end through the host-buffer C ABI (e2e) ----------------
    ids_host = torch.empty(bps, dtype=torch.int32).pin_memory()
    out_host = torch.empty(bps * shots, dtype=torch.uint8).pin_memory()
    hist_host = torch.empty(bps, 1 << N, dtype=torch.int32).pin_memory()

    def e2e_step(i):
        ids_host.copy_(torch.tensor(step_bases(i), dtype=torch.int32))
        diff.sample_to_host(ids_host, shots, out_host, hist_host)       # H2D ids, kernel, D2H bitstrings+counts, sync

    for i in range(min(args.warmup, 2)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(args.warmup + i)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * units_per_step * args.steps / e2e_s.item()
    assert int(hist_host.sum()) == units_per_step

    # ---------------- recon + fidelity milliseconds (rank 0) ----------------
    recon, train = None, None
    if rank == 0:
        # synthetic random-circuit state measured in all 3^8 bases with 10^6 shots each, generated on the
        # device by the native generator (csrc/synth.cu): brick-wall random circuit, Born sampling from the Philox stream
        psi_d = dq.synth_state(N, "rqc", depth=16, seed=args.seed, device=dev)
        h = dq.born_histograms(psi_d, N, 1_000_000, seed=args.seed)
        time.sleep(0.5)                     # let the clocks settle after the power-capped sampler run (the eigensolver is latency-bound)
        for _ in range(3):
            rho = dq.linear_inversion(h, N)
            dq.state_fidelity(psi_d, rho)
        torch.cuda.synchronize()
        times = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rho = dq.linear_inversion(h, N)
            f = dq.state_fidelity(psi_d, rho)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        recon_ms = statistics.median(times)
        recon = {"ms": recon_ms, "ms_runs": times, "what": "hist[6561,256] -> WHT -> rho[256,256] -> Jacobi PSD -> <psi|rho|psi>", "fidelity": f,
                 "input": "native generator: RQC depth 16, 6561 bases x 1e6 shots"}
    if rank == 0 and world == 1:
        # training step (T1) at the C4 architecture, batch 1024, tensor-core path replayed from a CUDA graph (single-GPU
        # runs only: inside a multi-rank job the step would all-reduce its gradients and wait for the other ranks)
        g = torch.Generator().manual_seed(1)
        x0p = torch.randint(0, 1 << N, (1024,), generator=g).to(torch.uint16).to(dev)
        b32 = torch.randint(0, NB, (1024,), generator=g).to(torch.int32).to(dev)
        tmodel = reference_state_dict().to(dev)
        tdiff = dq.DiscreteDiffusion(tmodel, T, dev, seed=args.seed, precision="bf16")
        tg = tdiff.make_train_graph(x0p, b32, dq.NativeAdam(tmodel, lr=1e-3))
        for _ in range(5):
            tg.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            tg.replay()
        b.record()
        torch.cuda.synchronize()
        train = {"ms": a.elapsed_time(b) / 20, "samples_per_s": 1024 / (a.elapsed_time(b) / 20) * 1e3,
                 "what": "C4 model, batch 1024: t draw + noising + tcgen05 bf16 fwd/bwd + Adam, CUDA-graph replay"}
        assert lib.ddqst_debug_tc_status() == 0, "tcgen05 pipeline timed out (train step)"

    # ---------------- roofline of the dominant kernel ----------------
    peaks, peak_src = measured_peaks()
    kern_ms = total_ms / args.steps                                    # one sampler launch per step
    flops = FLOP_PER_SAMPLE_STEP * T * units_per_step
    achieved = flops / (kern_ms / 1e3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": SAMPLER_DRAM_BYTES_PER_LAUNCH, "kernel": "sampler_pair_kernel<512>", "peak_source": f"{peak_src} bf16_tflops_sustained",
                "algorithmic_flop_per_launch": flops}

    # ---------------- CPU baseline (rank 0, N=1 only, bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, dt = cpu_port_rate(sd_cpu, args.cpu_shots, [0, 1], args.seed, threads)
        cpu = {"value": rate, "unit": "bitstrings/s", "cores": threads, "kind": "port",
               "sample": f"2 x p_sample({args.cpu_shots}, basis, 8), T=100, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": "generated bitstrings/sec (T-step D3PM, N=8)", "value": value, "unit": "bitstrings/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C4 slice: N=8 T=100 E=128 H=512 L=4, {bps} of 6561 bases x {shots} shots per step per GPU, "
                                   "posterior sampler, histogram fused", "weights": "torch default init, seed 0",
                       "l2": "flushed between steps (256 MiB write, untimed)", "timing": "CUDA events per step, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "bitstrings/s", "h2d_bytes_per_step": 4 * bps,
                    "d2h_bytes_per_step": bps * shots + bps * (4 << N)},
            "gpu_launches": args.steps,
            "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu, "recon_fidelity": recon, "train_step": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
