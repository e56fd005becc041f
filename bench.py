#!/usr/bin/env python
"""bench.py -- generated bitstrings/sec of the T-step D3PM reverse sampler at BASELINE.json's C4 shape
(N=8 qubits, 3^8 bases, T=100, E=128, H=512, L=4, 10^6 shots/basis), plus recon+fidelity milliseconds.

  python bench.py [--gpus N] [--steps K] [--warmup W]         # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]   # the UNMODIFIED reference (oracle/_ref) on the host cores

Native arm.  A step = one launch of the persistent tcgen05 sampler over `--bases-per-step` measurement bases x `--shots`
shots (a slice of the 6561-basis job; every step takes the next bases), histogram fused.  `value` is device-resident and
weak-scaled (every rank samples its own bases, no collective: SURVEY 8e); `e2e` goes through the host-buffer C ABI call
(ddqst_sample_host: H2D basis ids, kernel, D2H bitstrings + counts, sync).  Beyond the contract's keys the line carries
  full_job      all 6561 bases x 10 000 shots (the reference's shots_infer) sharded over the ranks -> NCCL all-reduce of the
                counts -> linear inversion -> PSD -> fidelity, timed end to end on every N (strong scaling, collective inside)
  train_step    C4 and C5 models, 1024 samples per GPU, CUDA-graph replay; for N > 1 data parallel with the NCCL gradient
                all-reduce captured inside the graph, every rank stepping
  recon_fidelity, c5, eager_b200, cpu_baseline.

Reference arm.  Imports nothing of this repo's package: the reference's own `ConditionalD3PM` + `DiscreteDiffusion.p_sample`
(RQC/model.py:27, RQC/diffusion.py:53) from the verbatim copy in oracle/_ref (oracle/make_ref.py), torch CPU, all host threads;
a step = `p_sample(10 000, basis, 8)` (10 000 = the reference's shots_infer, SS/config.py:23), same --steps/--warmup semantics.
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
C5 = dict(N=10, NB=59049, T=100, E=128, H=512, L=4)
REF_SHOTS = 10_000                         # the reference's shots_infer (SS/config.py:23)
METRIC = "generated bitstrings/sec (T-step D3PM, N=8)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--bases-per-step", type=int, default=8)
    ap.add_argument("--shots", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e only (skips full_job/train/recon/c5/eager legs)")
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
        num = lambda s: s.replace(".", "", 1).isdigit()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and num(r[1])]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and num(r[2])]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and num(r[3])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w": statistics.median(pw) if pw else None}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except OSError:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


# ================================================================================================ reference / CPU legs
# Everything in this section runs the reference (oracle/_ref, unmodified) or, where that copy is absent, the oracle port.
# None of it imports the product package.
def load_reference(phase="RQC", names=("model", "diffusion")):
    from oracle import ref_harness as rh
    if rh.path_root() is None:
        return None
    return rh.load_phase(phase, names)


def reference_sampler(device, threads=None):
    """-> (step(basis) -> seconds, kind).  The stock RQC model + DiscreteDiffusion.p_sample at the C4 architecture,
    torch default init under seed 0 (identical to the native arm's weights)."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    mods = load_reference()
    N, NB, T = C4["N"], C4["NB"], C4["T"]
    if mods is not None:
        torch.manual_seed(0)
        model = mods["model"].ConditionalD3PM(N, NB, T, C4["E"], C4["H"], C4["L"]).to(device).eval()
        diff = mods["diffusion"].DiscreteDiffusion(model, T, device)

        def step(basis, shots=REF_SHOTS):
            t0 = time.perf_counter()
            out = diff.p_sample(shots, basis, N)            # RQC/diffusion.py:53-80, unmodified
            if out.is_cuda:
                out = out.cpu()                            # what RQC/evaluate.py:83 does with every basis' samples
            assert out.shape == (shots, N)
            return time.perf_counter() - t0
        return step, "reference"
    from oracle import ddqst_oracle as orc
    sd = orc.default_init_state_dict(N, NB, T, C4["E"], C4["H"], C4["L"], seed=0)
    betas, q_bar = orc.cosine_schedule(T)

    def step(basis, shots=REF_SHOTS):
        t0 = time.perf_counter()
        orc.p_sample_posterior(sd, betas, q_bar, shots, basis, N, 1234)
        return time.perf_counter() - t0
    return step, "port"


def reference_recon_cpu(max_seconds=25.0):
    """recon + fidelity on the host cores with the reference's own loop body (RQC/reconstruct.py:56-67), 10 000 shots per
    basis (the reference's scale).  N=8 needs 65 536 iterations (146 s in the survey), so the timed
    sample is a seeded random subset of the Pauli strings pushed through the unmodified get_coefficient / get_pauli_matrix
    / accumulate statements, scaled by 65 536 / subset, plus the unmodified make_positive_semidefinite and <psi|rho|psi>.
    The unmodified linear_inversion is also run in full at N=6 as a cross-check of the per-iteration cost model."""
    import numpy as np
    mods = load_reference("RQC", ("reconstruct",))
    if mods is None:
        return None
    rec = mods["reconstruct"]
    from itertools import product
    rng = np.random.default_rng(0)
    out = {"kind": "reference", "cores": os.cpu_count(), "shots_per_basis": REF_SHOTS}

    def data_for(n):
        pool = [rng.integers(0, 2, size=(REF_SHOTS, n)).astype(np.int64) for _ in range(16)]      # distinct sample matrices, cycled:
        return {"".join(b): pool[i % 16] for i, b in enumerate(product("XYZ", repeat=n))}         # cost is value-independent
    # full unmodified call at N=6
    d6 = data_for(6)
    t0 = time.perf_counter()
    rho6 = rec.linear_inversion(d6, 6)
    out["n6_full_ms"] = 1e3 * (time.perf_counter() - t0)
    assert abs(np.trace(rho6.data).real - 1) < 1e-9
    # N=8: bounded subset of the loop
    n, dim = 8, 256
    d8 = data_for(n)
    all_paulis = ["".join(p) for p in product("IXYZ", repeat=n)]
    subset = rng.choice(len(all_paulis), size=768, replace=False)
    rho = np.zeros((dim, dim), dtype=complex)
    t0 = time.perf_counter()
    done = 0
    for idx in subset:
        pauli_str = all_paulis[idx]
        coeff = rec.get_coefficient(pauli_str, d8)       # RQC/reconstruct.py:62
        mat = rec.get_pauli_matrix(pauli_str)            # :63
        rho += coeff * mat                               # :64
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    loop_s = time.perf_counter() - t0
    herm = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    herm = (herm + herm.conj().T) / 2
    t0 = time.perf_counter()
    psd = rec.make_positive_semidefinite(herm / dim)     # :48-54 (LAPACK zheevd)
    psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    psi /= np.linalg.norm(psi)
    float(np.real(np.vdot(psi, psd.data @ psi)))          # state_fidelity for a pure target
    tail_s = time.perf_counter() - t0
    est = loop_s / done * len(all_paulis) + tail_s
    out.update({"value_ms": 1e3 * est, "unit": "ms", "sample": f"{done} of {len(all_paulis)} Pauli strings (seeded random subset) through the "
                "reference's get_coefficient/get_pauli_matrix/accumulate at N=8, x{:.1f}; + its PSD projection and <psi|rho|psi> "
                "({:.1f} ms)".format(len(all_paulis) / done, 1e3 * tail_s), "sampled_s": loop_s + tail_s})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, kind = reference_sampler("cpu", threads)
    for w in range(args.warmup):
        step(w % C4["NB"], shots=512)                     # warm-up: thread pool, allocator, MKL/oneDNN kernels
    times = [step((args.warmup + k) % C4["NB"]) for k in range(args.steps)]
    total = sum(times)
    value = REF_SHOTS * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "bitstrings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C4 slice: N=8 T=100 E=128 H=512 L=4, 1 of 6561 bases x {REF_SHOTS} shots per step: the reference's own "
                               "p_sample(shots_infer=10 000, basis, 8) call (RQC/diffusion.py:53), the unit RQC/evaluate.py:82-84 loops over; "
                               "the native arm runs the same architecture and weights on 8 bases x 10^6 shots per step",
                   "weights": "torch default init, seed 0", "code": "oracle/_ref (verbatim reference modules)" if kind == "reference" else "oracle port",
                   "warmup_steps": "512-shot calls (thread pool / kernel selection), untimed"},
        "cpu_baseline": {"value": value, "unit": "bitstrings/s", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} x p_sample({REF_SHOTS}, basis, 8), T=100"},
        "e2e": {"value": value, "unit": "bitstrings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ================================================================================================ native arm
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
    peaks, peak_src = measured_peaks()

    def make_model(cfg=C4, seed=0):
        torch.manual_seed(seed)
        return dq.ConditionalD3PM(cfg["N"], cfg["NB"], cfg["T"], cfg["E"], cfg["H"], cfg["L"]).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def tc_ok(what):
        assert lib.ddqst_debug_tc_status() == 0, f"tcgen05 pipeline timed out ({what})"

    model = make_model()
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
    total_ms = max_over_ranks(sum(step_ms))
    tc_ok("sampler")
    counts = hist.view(torch.int32).sum(dim=1).cpu().tolist()
    assert all(c == shots for c in counts), counts
    value = world * units_per_step * args.steps / (total_ms / 1e3)
    joules_per_bitstring = None
    if clock_info.get("power_w"):
        joules_per_bitstring = clock_info["power_w"] * (total_ms / 1e3) / (units_per_step * args.steps)

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
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * units_per_step * args.steps / e2e_s
    assert int(hist_host.sum()) == units_per_step
    assert int(np.bincount(out_host[:shots].numpy(), minlength=256).sum()) == shots

    extras = {}
    if not args.no_extras:
        extras = native_extras(args, dq, lib, dev, world, rank, model, diff, peaks, peak_src, barrier, max_over_ranks, tc_ok, make_model)

    # ---------------- roofline of the dominant kernel ----------------
    kern_ms = total_ms / args.steps                                    # one sampler launch per step
    flops = FLOP_PER_SAMPLE_STEP * T * units_per_step
    achieved = flops / (kern_ms / 1e3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    traffic, traffic_src = None, "profiles/: no --set full capture of this launch configuration recorded"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_sampler_traffic.json")))
        if tj.get("bases_per_step") == bps and tj.get("shots") == shots:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "sampler_pair_kernel<512>",
                "peak_source": f"{peak_src} bf16_tflops_sustained (the kernel runs for seconds under the power cap)",
                "frac_of_burst_peak": achieved / float(peaks.get("bf16_tflops", peak)), "algorithmic_flop_per_launch": flops}

    # ---------------- CPU baseline (rank 0, N=1 only, bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step, kind = reference_sampler("cpu", threads)
        step(0, shots=512)
        times = [step(b) for b in (1, 2)]
        cpu = {"value": REF_SHOTS * len(times) / sum(times), "unit": "bitstrings/s", "cores": threads, "kind": kind,
               "sample": f"{len(times)} x p_sample({REF_SHOTS}, basis, 8), T=100, {sum(times):.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "bitstrings/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C4 slice: N=8 T=100 E=128 H=512 L=4, {bps} of 6561 bases x {shots} shots per step per GPU, "
                                   "posterior sampler, histogram fused", "weights": "torch default init, seed 0",
                       "l2": "flushed between steps (256 MiB write, untimed)", "timing": "CUDA events per step, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "bitstrings/s", "h2d_bytes_per_step": 4 * bps,
                    "d2h_bytes_per_step": bps * shots + bps * (4 << N)},
            "gpu_launches": args.steps,
            "clocks": clock_info, "joules_per_bitstring": joules_per_bitstring, "roofline": roofline, "cpu_baseline": cpu,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def native_extras(args, dq, lib, dev, world, rank, model, diff, peaks, peak_src, barrier, max_over_ranks, tc_ok, make_model):
    """The legs beyond the headline: full tomography job with the histogram all-reduce inside the timed region, training
    steps (data parallel for world > 1), recon + fidelity, C5 and un-timed modes, PyTorch-eager secondary bar."""
    import numpy as np
    import torch
    import torch.distributed as dist
    N, NB, T = C4["N"], C4["NB"], C4["T"]
    out = {}
    hbm = float(peaks.get("hbm_gbs", 6550.0))
    burst = float(peaks.get("bf16_tflops", 1655.0))

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return statistics.median(ms), ms

    # ---------------- full job: every basis, sharded, counts all-reduced, rho, fidelity (strong scaling) ----------------
    psi_d = dq.synth_state(N, "rqc", depth=16, seed=args.seed, device=dev)
    all_bases = list(range(NB))
    dq.sample_sharded(diff, all_bases[:64], 256)                       # warm-up incl. the NCCL communicator
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    table = dq.sample_sharded(diff, all_bases, REF_SHOTS)              # sample own shard (fused histogram) + ncclAllReduce(int32 sum)
    rho = dq.linear_inversion(table, N)                                # every rank holds the full table; rho replicated
    fid = dq.state_fidelity(psi_d, rho)
    b.record()
    barrier()
    job_ms = max_over_ranks(a.elapsed_time(b))
    tc_ok("full job")
    assert int(table.view(torch.int32).to(torch.int64).sum().item()) == NB * REF_SHOTS
    ident = None
    if world > 1:
        sub = list(range(0, NB, NB // 16))[:16]
        sharded = dq.sample_sharded(diff, sub, 4096)
        single = diff.sample(sub, 4096)[0]
        ident = bool(torch.equal(sharded.view(torch.int32), single.view(torch.int32)))
        flag = torch.tensor([1 if ident else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ident = bool(flag.item())
    out["full_job"] = {"ms": job_ms, "bitstrings": NB * REF_SHOTS, "bitstrings_per_s": NB * REF_SHOTS / (job_ms / 1e3), "scaling": "strong",
                       "what": f"all {NB} bases x {REF_SHOTS} shots: sample_sharded (by basis) -> ncclAllReduce int32[{NB},256] (6.7 MB) -> "
                               "linear inversion -> PSD -> <psi|rho|psi>, one timed region, max over ranks",
                       "collective": "ncclAllReduce(int32, sum) of the counts table" if world > 1 else None,
                       "sharded_hist_bit_identical": ident, "fidelity_untrained_weights": fid}

    # ---------------- training step (T1): CUDA-graph replay; data parallel with the all-reduce captured for world > 1 ------
    def train_leg(cfg, batch=1024, reps=20):
        n = cfg["N"]
        g = torch.Generator().manual_seed(1 + rank)
        x0p = torch.randint(0, 1 << n, (batch,), generator=g).to(torch.int32).to(torch.uint16).to(dev)
        b32 = torch.randint(0, cfg["NB"], (batch,), generator=g).to(torch.int32).to(dev)
        tmodel = make_model(cfg)
        tdiff = dq.DiscreteDiffusion(tmodel, cfg["T"], dev, seed=args.seed, precision="bf16")
        topt = dq.NativeAdam(tmodel, lr=1e-3)
        how = "CUDA-graph replay"
        try:
            tg = tdiff.make_train_graph(x0p, b32, topt, data_parallel=world > 1)
            step = tg.replay
            ok = 1
        except Exception as e:                               # capture refused (e.g. NCCL inside a graph): every rank falls back together
            ok, err = 0, repr(e)[:160]
        flag = torch.tensor([ok], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            how = "eager launches (graph capture failed: " + (err if not ok else "on another rank") + ")"
            step = lambda: tdiff.train_step(x0p, b32, topt, data_parallel=world > 1, validate=False)
        for _ in range(5):
            loss_t = step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            loss_t = step()
        b.record()
        barrier()
        ms = max_over_ranks(a.elapsed_time(b) / reps)
        tc_ok("train step")
        loss = float(loss_t.item())
        assert loss == loss and 0.0 < loss < 2.0, loss
        params = tmodel.flat_params.numel()
        flop = 3 * (cfg["L"] * 2 * 2 * cfg["H"] * cfg["H"] + 2 * cfg["H"] * 2 * n) * batch
        ach = flop / (ms / 1e3) / 1e12
        del step, tdiff, tmodel
        return {"ms": ms, "how": how, "samples_per_s": world * batch / (ms / 1e3), "global_batch": world * batch, "params": params, "loss": loss,
                "roofline": {"bound": "tensor", "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst,
                             "peak_source": f"{peak_src} bf16_tflops (burst: a sub-millisecond step timed alone)",
                             "algorithmic_flop_per_step_per_gpu": flop},
                "collective": f"ncclAllReduce(fp32, sum) of the flat gradient, {4 * params / 1e6:.1f} MB, captured in the graph" if world > 1 else None,
                "what": f"t draw + noising + tcgen05 bf16 forward/backward + Adam, {batch} samples per GPU"
                        + (", data parallel" if world > 1 else "")}
    out["train_step"] = train_leg(C4)
    out["train_step_b8192"] = train_leg(C4, batch=8192)      # fused forward + data-gradient kernel (csrc/train_fused.cuh), 8192 samples per GPU
    out["train_step_c5"] = train_leg(C5)

    if rank != 0:
        return out
    # the legs below run on rank 0 only (no collectives): a failure in one is recorded in the line instead of taking it down
    try:
        rank0_legs(args, dq, lib, dev, world, peaks, tc_ok, make_model, timed, psi_d, out)
    except Exception as e:
        out["rank0_legs_error"] = repr(e)[:300]
    return out


def rank0_legs(args, dq, lib, dev, world, peaks, tc_ok, make_model, timed, psi_d, out):
    import numpy as np
    import torch
    N, NB, T = C4["N"], C4["NB"], C4["T"]
    hbm = float(peaks.get("hbm_gbs", 6550.0))
    burst = float(peaks.get("bf16_tflops", 1655.0))
    # ---------------- recon + fidelity milliseconds (rank 0), 10^6 shots per basis ----------------
    h = dq.born_histograms(psi_d, N, 1_000_000, seed=args.seed)
    time.sleep(0.5)                     # let the clocks settle after the power-capped sampler runs (the eigensolver is latency-bound)
    state = {}

    def recon_all():
        state["rho"] = dq.linear_inversion(h, N)
        state["f"] = dq.state_fidelity(psi_d, state["rho"])
    recon_ms, runs = timed(recon_all, 5, warm=3)
    li_ms, _ = timed(lambda: dq.linear_inversion_raw(h, N), 5)
    raw = dq.linear_inversion_raw(h, N)
    psd_ms, _ = timed(lambda: dq.make_positive_semidefinite(raw), 5)
    fid_ms, _ = timed(lambda: dq.state_fidelity(psi_d, state["rho"]), 5)
    li_bytes = 4 * NB * (1 << N) + 16 * (1 << 2 * N)
    recon = {"ms": recon_ms, "ms_runs": runs, "fidelity": state["f"],
             "what": "hist[6561,256] -> WHT -> rho[256,256] -> PSD (eigensolver) -> <psi|rho|psi>",
             "input": "native generator: RQC depth 16, 6561 bases x 1e6 shots",
             "parts_ms": {"linear_inversion": li_ms, "psd_projection": psd_ms, "fidelity_pure": fid_ms},
             "roofline": {"bound": "hbm", "achieved": li_bytes / (li_ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                          "frac": li_bytes / (li_ms / 1e3) / 1e9 / hbm, "kernel": "coeff_table + rho_assemble (linear inversion)",
                          "algorithmic_bytes": li_bytes,
                          "note": "7.8 MB of traffic: latency-bound at N=8; the PSD eigensolver (FP64, no HBM traffic) is the rest of the time"}}
    if world == 1 and not args.no_cpu_baseline:
        recon["cpu_baseline"] = reference_recon_cpu()
        # like for like: the GPU path on a 10 000-shot table (time is independent of the shot count)
        h10k = dq.born_histograms(psi_d, N, REF_SHOTS, seed=args.seed)
        recon["ms_at_10k_shots"] = timed(lambda: dq.state_fidelity(psi_d, dq.linear_inversion(h10k, N)), 3)[0]
    out["recon_fidelity"] = recon

    if world > 1:
        return out

    # ---------------- C5 (N=10, mixed state) and the un-timed modes ----------------
    c5 = {}
    n5 = C5["N"]
    m5 = make_model(C5)
    d5 = dq.DiscreteDiffusion(m5, C5["T"], dev, seed=args.seed, precision="bf16")
    hist5 = torch.zeros(8, 1 << n5, dtype=torch.uint32, device=dev)
    s5 = 200_000
    ms5, _ = timed(lambda: (hist5.zero_(), d5.sample(list(range(8)), s5, hist_out=hist5)), 2, warm=1)
    tc_ok("C5 sampler")
    flop5 = (C5["L"] * 2 * 2 * 512 * 512 + 2 * 512 * 2 * n5) * C5["T"] * 8 * s5
    sust = float(peaks.get("bf16_tflops_sustained", burst))
    c5["sampler"] = {"bitstrings_per_s": 8 * s5 / (ms5 / 1e3), "ms": ms5, "what": f"N=10, 59 049-row FiLM table, 8 bases x {s5} shots, uint16 outcomes, histogram fused",
                     "roofline": {"bound": "tensor", "achieved": flop5 / (ms5 / 1e3) / 1e12, "peak": sust, "unit": "TFLOP/s",
                                  "frac": flop5 / (ms5 / 1e3) / 1e12 / sust}}
    del d5, m5
    psi5 = dq.synth_state(n5, "rqc", depth=16, seed=args.seed, device=dev)
    h5 = dq.born_histograms(psi5, n5, 100_000, seed=args.seed, noise_type="depolarizing", error_rate=0.1)
    dim5 = 1 << n5
    target5 = dq.DensityMatrix(0.9 * torch.outer(psi5, psi5.conj()) + 0.1 * torch.eye(dim5, dtype=torch.complex128, device=dev) / dim5)
    li5, _ = timed(lambda: dq.linear_inversion_raw(h5, n5), 3, warm=1)
    raw5 = dq.linear_inversion_raw(h5, n5)
    psd5, _ = timed(lambda: dq.make_positive_semidefinite(raw5), 2, warm=1)
    rho5 = dq.make_positive_semidefinite(raw5)
    st5 = {}

    def mixed():
        st5["f"] = dq.state_fidelity(target5, rho5)
    fm5, _ = timed(mixed, 2, warm=1)
    bytes5 = 4 * C5["NB"] * dim5 + 16 * dim5 * dim5
    c5["recon"] = {"linear_inversion_ms": li5, "psd_ms": psd5, "mixed_fidelity_ms": fm5, "fidelity": st5["f"],
                   "what": "N=10 depolarized (p=0.1) RQC state, 59 049 bases x 1e5 shots: hist -> rho[1024,1024] -> PSD -> Uhlmann fidelity vs the mixed target",
                   "roofline": {"bound": "hbm", "achieved": bytes5 / (li5 / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                                "frac": bytes5 / (li5 / 1e3) / 1e9 / hbm, "kernel": "linear inversion N=10", "algorithmic_bytes": bytes5}}
    del h5, raw5, rho5, target5
    # renoise sampler + variant A (SS/model.py, SS/diffusion.py:54-82) at N=3, E=64
    torch.manual_seed(0)
    mA = dq.ConditionalD3PM(3, 27, 100, 64, 512, 4, variant="A").to(dev)
    dA = dq.DiscreteDiffusion(mA, 100, dev, schedule="linear", seed=args.seed, precision="bf16")
    sA = 250_000
    msA, _ = timed(lambda: dA.sample(list(range(27)), sA), 2, warm=1)
    tc_ok("variant A sampler")
    flopA = (4 * 2 * 2 * 512 * 512 + 2 * 512 * 6) * 100 * 27 * sA
    c5["renoise_variant_A"] = {"bitstrings_per_s": 27 * sA / (msA / 1e3), "ms": msA, "what": f"SS variant (E=64, H=512, L=4), linear schedule, "
                               f"x0-hat + re-noise sampler, N=3, 27 bases x {sA} shots",
                               "roofline": {"bound": "tensor", "achieved": flopA / (msA / 1e3) / 1e12, "peak": sust, "unit": "TFLOP/s",
                                            "frac": flopA / (msA / 1e3) / 1e12 / sust}}
    del dA, mA
    out["c5"] = c5

    # ---------------- secondary bar: the reference code itself with device='cuda' (PyTorch eager on this B200) ----------------
    try:
        torch.cuda.empty_cache()
        step, kind = reference_sampler(dev)
        if kind == "reference":
            step(0, shots=512)
            t_small = statistics.median([step(b) for b in (1, 2, 3)])
            eager = {"kind": "reference", "p_sample_10k": {"bitstrings_per_s": REF_SHOTS / t_small, "ms": 1e3 * t_small,
                                                            "what": "unmodified RQC p_sample(10 000, basis, 8) with device='cuda', incl. .cpu() of the samples"}}
            try:
                big = 1_000_000
                step(4, shots=100_000)
                t_big = step(5, shots=big)
                eager["p_sample_1M"] = {"bitstrings_per_s": big / t_big, "ms": 1e3 * t_big,
                                        "what": "same call with 10^6 shots (what the native arm runs per basis)"}
            except RuntimeError as e:                                  # out of memory at 10^6 rows
                eager["p_sample_1M"] = {"error": str(e)[:120]}
            out["eager_b200"] = eager
        else:
            out["eager_b200"] = {"unavailable": "oracle/_ref is not populated (python oracle/make_ref.py)"}
    except Exception as e:                                             # the secondary bar must never take the headline down
        out["eager_b200"] = {"error": repr(e)[:200]}
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
