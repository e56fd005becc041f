#!/usr/bin/env python
"""'Matching fidelity' on the reference's own published configurations (notes.pdf p.10 Table 3):

    Bell N=2, 9 bases, 5 000 train shots/basis -> 10 000 generated/basis, linear inversion + PSD   F = 0.95565
    GHZ  N=3, 27 bases, same protocol                                                              F = 0.87092
    SS/main.py:131 pass threshold                                                                  F > 0.9

The SS protocol (SS/config.py:3-24, SS/main.py:67-131) is run unchanged on the native path: variant-A model (E=64, H=512,
4 blocks), linear schedule, T=100, AdamW lr 1e-4 (weight decay 0.01), batch 256, 300 epochs over the unrolled shots with a
fresh shuffle per epoch, x0-hat + re-noise sampler, 10 000 generated shots per basis, linear inversion + PSD projection,
fidelity against (|0..0> + |1..1>)/sqrt(2).  Data: ideal Born sampling of the SS/data_gen.py circuit (the reference uses
the noiseless AerSimulator).  Several seeds show the run-to-run spread.

--oracle additionally trains the CPU oracle (torch fp32, the reference's arithmetic) on the IDENTICAL batches / timestep /
noise stream for the same number of steps and samples with the same stream, for Bell N=2 at 1 000 train shots/basis (the
SS default, 10 800 steps): native and oracle must agree to within sampling noise at equal budget.

    python benchmarks/fidelity_check.py --out profiles/r2_fidelity_check.json [--oracle] [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddqst_b200 as dq                       # noqa: E402

SS = dict(T=100, E=64, H=512, L=4, batch=256, lr=1e-4, epochs=300, shots_infer=10_000)
PUBLISHED = {("bell", 2, 5000): 0.95565, ("ghz", 3, 5000): 0.87092}


def ghz_target(n):
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = psi[-1] = 1 / np.sqrt(2)
    return psi


def native_run(state, n, shots_train, seed, epochs, precision="bf16"):
    dev = torch.device("cuda")
    nb = 3 ** n
    hist, _, _psi = dq.generate_synthetic_data(n, state, shots_train, noise_type="ideal", seed=seed)
    ds = dq.QuantumStateDataset.from_counts_table(hist, n, seed=seed)
    torch.manual_seed(seed)
    model = dq.ConditionalD3PM(n, nb, SS["T"], SS["E"], SS["H"], SS["L"], variant="A").to(dev)
    diff = dq.DiscreteDiffusion(model, SS["T"], dev, schedule="linear", seed=seed, precision="bf16")
    opt = dq.NativeAdam(model, lr=SS["lr"], weight_decay=0.01, decoupled=True)
    steps = epochs * ds.batches_per_epoch(SS["batch"])
    x0_s, b_s = ds.batch(0, SS["batch"])
    graph = diff.make_train_graph(x0_s, b_s, opt) if precision == "bf16" else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = []
    for step in range(steps):
        x0, basis = ds.batch(step, SS["batch"])
        if graph is not None:
            x0_s.copy_(x0); b_s.copy_(basis)
            loss = graph.replay()
        else:
            loss = diff.train_step(x0, basis, opt, precision="fp32", validate=False)
        if (step + 1) % max(1, steps // 6) == 0:
            losses.append(round(float(loss.item()), 4))
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    syn = diff.sample(list(range(nb)), SS["shots_infer"])[0]
    rep = dq.recon_report(syn, n, dq.Statevector(ghz_target(n)), convention="unreversed")
    torch.cuda.synchronize()
    infer_s = time.perf_counter() - t0
    raw = dq.recon_report(hist, n, dq.Statevector(ghz_target(n)), convention="unreversed")
    assert dq._lib.load().ddqst_debug_tc_status() == 0
    return {"fidelity": rep.fidelity, "fidelity_raw_train_counts": raw.fidelity, "purity": rep.purity, "steps": steps, "losses": losses,
            "train_s": train_s, "train_step_ms": 1e3 * train_s / steps, "sample_recon_s": infer_s, "seed": seed, "precision": precision}


def oracle_run(state, n, shots_train, seed, epochs):
    """The reference's arithmetic (torch fp32 CPU, oracle/ddqst_oracle.py) on the same batches and the same injected stream."""
    from oracle import ddqst_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    nb = 3 ** n
    hist, _, _psi = dq.generate_synthetic_data(n, state, shots_train, noise_type="ideal", seed=seed)
    h = hist.view(torch.int32).cpu().numpy().astype(np.int64)
    row_basis = np.arange(nb)
    sd = orc.default_init_state_dict(n, nb, SS["T"], SS["E"], SS["H"], SS["L"], variant="A", seed=seed)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=SS["lr"])                     # SS/main.py:77 (weight_decay 0.01 default)
    betas, Q = orc.linear_schedule(SS["T"])
    total = int(h.sum())
    per_epoch = (total + SS["batch"] - 1) // SS["batch"]
    steps = epochs * per_epoch
    t0 = time.perf_counter()
    losses = []
    for step in range(steps):
        start = step * SS["batch"]
        epoch, off = divmod(start, total)
        x0p, basis = orc.counts_batch(h, row_basis, n, off, SS["batch"], seed, epoch)[:2]
        x0 = torch.from_numpy(((np.asarray(x0p)[:, None] >> np.arange(n)) & 1).astype(np.int64))
        loss, _, _ = orc.train_step(params, opt, Q, x0, torch.from_numpy(np.asarray(basis).astype(np.int64)), n, SS["T"], seed, step,
                                    cumulative=False)
        if (step + 1) % max(1, steps // 6) == 0:
            losses.append(round(float(loss), 4))
    train_s = time.perf_counter() - t0
    final = {k: v.detach() for k, v in params.items()}
    t0 = time.perf_counter()
    table = np.zeros((nb, 1 << n), dtype=np.int64)
    for b in range(nb):
        x = orc.p_sample_renoise(final, Q, SS["shots_infer"], b, n, seed).numpy()
        table[b] = orc.histogram(x, n)
    rho = orc.linear_inversion_hist(table, n, reversed_kron=False)
    fid = orc.state_fidelity(ghz_target(n), rho)
    return {"fidelity": fid, "steps": steps, "losses": losses, "train_s": train_s, "sample_recon_s": time.perf_counter() - t0, "seed": seed,
            "cores": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--quick", action="store_true", help="30 epochs instead of 300 (smoke run)")
    ap.add_argument("--seeds", type=int, default=3)
    args = ap.parse_args()
    epochs = 30 if args.quick else SS["epochs"]
    out = {"protocol": dict(SS, epochs=epochs, optimizer="AdamW(lr=1e-4, weight_decay=0.01)", sampler="x0-hat + re-noise (SS/diffusion.py:54-82)",
                            schedule="linear", model="variant A (SS/model.py)"), "published": {f"{k[0]} N={k[1]} shots_train={k[2]}": v for k, v in PUBLISHED.items()},
           "threshold": "fidelity > 0.9 (SS/main.py:131)", "runs": []}
    for state, n in (("bell", 2), ("ghz", 3)):
        for shots in (1000, 5000):
            runs = [native_run(state, n, shots, seed, epochs) for seed in range(args.seeds)]
            fids = [r["fidelity"] for r in runs]
            entry = {"state": state, "num_qubits": n, "shots_train": shots, "native_bf16": runs, "native_fidelity_mean": float(np.mean(fids)),
                     "native_fidelity_min": float(np.min(fids)), "native_fidelity_max": float(np.max(fids)),
                     "published": PUBLISHED.get((state, n, shots))}
            if args.oracle and (state, n, shots) == ("bell", 2, 1000):
                entry["native_fp32_seed0"] = native_run(state, n, shots, 0, epochs, precision="fp32")
                entry["oracle_cpu_seed0"] = oracle_run(state, n, shots, 0, epochs)
            out["runs"].append(entry)
            print(json.dumps({k: v for k, v in entry.items() if k != "native_bf16"}), flush=True)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
