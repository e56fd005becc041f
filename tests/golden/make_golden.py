"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE.

Run once in the build container (needs /root/reference):
    python tests/golden/make_golden.py
The reference modules are imported read-only through oracle/ref_harness.py; the
only intervention is the injected Philox stream replacing torch.randint /
torch.multinomial (the reference is unseeded).  Outputs: tests/golden/*.npz.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ddqst_oracle as orc  # noqa: E402
from oracle import ref_harness as rh    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sd_np(model, prefix="sd."):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, sum(np.asarray(a).nbytes for a in arrs.values()), "bytes raw")


def main():
    assert rh.available(), "needs /root/reference"
    torch.set_num_threads(1)
    R = rh.load_phase("RQC")
    S = rh.load_phase("SS")

    # ---- schedules (D1, D1', NB) at the reference's T=100
    dR = R["diffusion"].DiscreteDiffusion(None, 100, "cpu")
    dS = S["diffusion"].DiscreteDiffusion(None, 100, "cpu")
    nb6 = rh.load_notebook_classes(6)
    nb6["BASIS_LABELS"] = ["X", "Y", "Z"]
    torch.manual_seed(0)
    ddm100 = nb6["BitstringDDM"](nb6["SimpleMLP"](100, 3), 100, "cpu")
    save("schedules.npz", cos_betas=dR.betas.numpy(), cos_Q_bar=dR.Q_bar.numpy(), lin_Q=dS.Q.numpy(),
         nb_Q=ddm100.Q.numpy())

    # ---- small models: forward, sampling, noising, training
    N, NB, T, E, H, L = 3, 27, 20, 16, 64, 2
    g = torch.Generator().manual_seed(7)
    x = torch.randint(0, 2, (64, N), generator=g)
    t = torch.randint(1, T + 1, (64,), generator=g)
    bs = torch.randint(0, NB, (64,), generator=g)
    for tag, mods, mode, qmode in (("B", R, "posterior", "q_cumulative"), ("A", S, "renoise", "q_marginal")):
        torch.manual_seed(11 if tag == "B" else 12)
        m = mods["model"].ConditionalD3PM(N, NB, T, E, H, L)
        # move the weights off the init distribution a little so biases/embeddings matter
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.05 * torch.randn_like(p))
        d = mods["diffusion"].DiscreteDiffusion(m, T, "cpu")
        logits = m(x, t, bs).detach()
        shots, basis, seed, off = 256, 5, 1234, 1000
        st = rh.InjectedStream(mode, seed, basis, N, T, shots, offset=off)
        with rh.injected(st):
            samples = d.p_sample(shots, basis, N)
        stq = rh.InjectedStream(qmode, 99, 3, N, T, 64, offset=17)
        with rh.injected(stq):
            xq = d.q_sample(x, t)
        # training: 3 steps of the reference's inline loop body (RQC/main.py:105-115, SS/main.py:84-99)
        if tag == "B":
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        else:
            opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
        init_sd = sd_np(m)
        x0 = torch.randint(0, 2, (128, N), generator=g)
        b0 = torch.randint(0, NB, (128,), generator=g)
        losses = []
        for step in range(3):
            stt = rh.InjectedStream("train", 4321, step, N, T, 128, cumulative=(tag == "B"))
            with rh.injected(stt):
                tt = torch.randint(1, T + 1, (128,))
                x_t = d.q_sample(x0, tt)
            lg = m(x_t, tt, b0)
            loss = torch.nn.functional.cross_entropy(lg.permute(0, 2, 1), x0)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
        final_sd = sd_np(m, "trained.")
        save(f"model_{tag}_small.npz", dims=np.array([N, NB, T, E, H, L]), x=x.numpy(), t=t.numpy(), basis=bs.numpy(),
             logits=logits.numpy(), sample_args=np.array([shots, basis, seed, off]), samples=samples.numpy(),
             q_args=np.array([99, 3, 17]), q_out=xq.numpy(), train_x0=x0.numpy(), train_basis=b0.numpy(),
             train_seed=np.array([4321]), train_losses=np.array(losses), **init_sd, **final_sd)

    # ---- notebook MLPs (M5) + their renoise sampler
    out = {}
    for cell, cls in ((6, "SimpleMLP"), (12, "UpgradedMLP")):
        ns = rh.load_notebook_classes(cell)
        ns["BASIS_LABELS"] = ["X", "Y", "Z"]
        torch.manual_seed(21 + cell)
        m = ns[cls](T, 3)
        ddm = ns["BitstringDDM"](m, T, "cpu")
        g2 = torch.Generator().manual_seed(5)
        x1 = torch.randint(0, 2, (32,), generator=g2)
        t1 = torch.randint(1, T + 1, (32,), generator=g2)
        b1 = torch.randint(0, 3, (32,), generator=g2)
        out[f"{cls}.x"], out[f"{cls}.t"], out[f"{cls}.basis"] = x1.numpy(), t1.numpy(), b1.numpy()
        out[f"{cls}.logits"] = m(x1, t1, b1).detach().numpy()
        st = rh.InjectedStream("renoise_nb", 31, 2, 1, T, 200, offset=3)
        with rh.injected(st), contextlib.redirect_stdout(io.StringIO()):
            out[f"{cls}.samples"] = ddm.sample(200, 2)
        out.update(sd_np(m, f"{cls}.sd."))
    save("nb_mlp.npz", T=np.array([T]), sample_args=np.array([200, 2, 31, 3]), **out)

    # ---- reconstruction (R1-R4) on synthetic Haar states, both Kronecker conventions
    rng = np.random.default_rng(0)
    rec = {}
    for n in (1, 2, 3):
        names = orc.basis_strings(n)
        psi = orc.haar_state(n, seed=n)
        data, hist = {}, np.zeros((3 ** n, 2 ** n), np.int64)
        for b, name in enumerate(names):
            s = rng.choice(2 ** n, size=400 + 3 * b, p=orc.born_probabilities(psi, n, name))
            samp = ((s[:, None] >> np.arange(n)) & 1).astype(np.int64)
            data[name] = samp
            hist[b] = orc.histogram(samp, n)
        rec[f"N{n}.psi"], rec[f"N{n}.hist"] = psi, hist
        rec[f"N{n}.rho_rqc"] = R["reconstruct"].linear_inversion(data, n).data
        rec[f"N{n}.rho_ss"] = S["reconstruct"].linear_inversion(data, n).data
        rec[f"N{n}.coeff_first"] = np.array([R["reconstruct"].get_coefficient(p, data)
                                             for p in ("X" + "I" * (n - 1), "Z" * n, "I" * n)])
    save("recon_small.npz", **rec)

    # ---- shipped Datapoints (N=3 RQC records): counts -> rho -> fidelity/metrics through the reference
    dp = {}
    recs = rh.load_datapoints(os.path.join(rh.REF_ROOT, "Datapoints/rqc_N3_data/part_0.pt"))[:4]
    recs += rh.load_datapoints(os.path.join(rh.REF_ROOT, "Datapoints/rqc_N3_data/part_20.pt"))[:1]
    for i, r in enumerate(recs):
        psi, hist = rh.record_to_arrays(r, 3)
        data = {n: rh.expand_hist_to_samples(hist[b], 3) for b, n in enumerate(orc.basis_strings(3))}
        rho = R["reconstruct"].linear_inversion(data, 3)
        dp[f"r{i}.psi"], dp[f"r{i}.hist"], dp[f"r{i}.rho"] = psi, hist, rho.data
        dp[f"r{i}.metrics"] = np.array(R["reconstruct"].get_metrics(rho, 3))
        dp[f"r{i}.id_depth"] = np.array([r["id"], r["depth"]])
    save("datapoints_N3.npz", n=np.array([len(recs)]), **dp)


if __name__ == "__main__":
    main()
