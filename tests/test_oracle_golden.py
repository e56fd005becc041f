"""The oracle must reproduce the fixtures the UNMODIFIED reference produced (tests/golden/make_golden.py)
and the known-answer numbers hard-coded in the reference notebook / report.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden
from oracle import ddqst_oracle as orc


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    z = orc.philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    assert [int(v) for v in z] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = orc.philox4x32_10(np.full(4, 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))
    assert [int(v) for v in f] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    p = orc.philox4x32_10(np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], np.uint32),
                          np.array([0xA4093822, 0x299F31D0], np.uint32))
    assert [int(v) for v in p] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_schedules_match_reference():
    z = load_golden("schedules.npz")
    betas, Q_bar = orc.cosine_schedule(100)
    assert np.array_equal(betas.numpy(), z["cos_betas"]) and np.array_equal(Q_bar.numpy(), z["cos_Q_bar"])
    assert np.array_equal(orc.linear_schedule(100)[1].numpy(), z["lin_Q"])
    assert np.array_equal(orc.notebook_schedule(100)[1].numpy(), z["nb_Q"])
    # values quoted in SURVEY section 4 (RQC/diffusion.py:20-31)
    assert betas[0] == 0 and abs(betas[1].item() - 0.00063128158217296) < 1e-12
    assert abs(betas[100].item() - 0.999) < 1e-7
    assert abs(Q_bar[50, 0, 0].item() - 0.620235) < 1e-5


@pytest.mark.parametrize("tag", ["A", "B"])
def test_forward_sampling_noising_match_reference(tag):
    z = load_golden(f"model_{tag}_small.npz")
    N, NB, T, E, H, L = (int(v) for v in z["dims"])
    sd = golden_state_dict(z)
    x, t, b = (torch.from_numpy(z[k]) for k in ("x", "t", "basis"))
    assert torch.equal(orc.denoiser_forward(sd, x, t, b, N), torch.from_numpy(z["logits"]))
    shots, basis, seed, off = (int(v) for v in z["sample_args"])
    qseed, qstream, qoff = (int(v) for v in z["q_args"])
    if tag == "B":
        betas, Q = orc.cosine_schedule(T)
        got = orc.p_sample_posterior(sd, betas, Q, shots, basis, N, seed, shot_offset=off)
        q = orc.q_sample_cumulative(Q, x, t, qseed, qstream, qoff)
    else:
        _, Q = orc.linear_schedule(T)
        got = orc.p_sample_renoise(sd, Q, shots, basis, N, seed, shot_offset=off)
        q = orc.q_sample_marginal(Q, x, t, qseed, qstream, row_offset=qoff)
    assert np.array_equal(got.numpy(), z["samples"])
    assert np.array_equal(q.numpy(), z["q_out"])


@pytest.mark.parametrize("tag", ["A", "B"])
def test_train_steps_match_reference(tag):
    z = load_golden(f"model_{tag}_small.npz")
    N, NB, T, E, H, L = (int(v) for v in z["dims"])
    params = {k: v.clone().requires_grad_(True) for k, v in golden_state_dict(z).items()}
    if tag == "B":
        opt = torch.optim.Adam(list(params.values()), lr=1e-3)
        sched = orc.cosine_schedule(T)[1]
    else:
        opt = torch.optim.AdamW(list(params.values()), lr=1e-4)
        sched = orc.linear_schedule(T)[1]
    x0, b0 = torch.from_numpy(z["train_x0"]), torch.from_numpy(z["train_basis"])
    losses = [orc.train_step(params, opt, sched, x0, b0, N, T, int(z["train_seed"][0]), s, cumulative=(tag == "B"))[0].item()
              for s in range(3)]
    assert np.allclose(losses, z["train_losses"], rtol=0, atol=1e-6)
    for k, v in golden_state_dict(z, "trained.").items():
        assert torch.allclose(params[k].detach(), v, rtol=0, atol=1e-6), k


def test_notebook_mlps_match_reference():
    z = load_golden("nb_mlp.npz")
    T = int(z["T"][0])
    shots, basis, seed, off = (int(v) for v in z["sample_args"])
    _, Q = orc.notebook_schedule(T)
    for cls in ("SimpleMLP", "UpgradedMLP"):
        sd = golden_state_dict(z, f"{cls}.sd.")
        x, t, b = (torch.from_numpy(z[f"{cls}.{k}"]) for k in ("x", "t", "basis"))
        assert torch.equal(orc.notebook_mlp_forward(sd, x, t, b), torch.from_numpy(z[f"{cls}.logits"]))
        got = orc.p_sample_renoise(None, Q, shots, basis, 1, seed, shot_offset=off,
                                   forward=lambda xx, tt, bb: orc.notebook_mlp_forward(sd, xx[:, 0], tt, bb).view(-1, 1, 2))
        assert np.array_equal(got[:, 0].numpy(), z[f"{cls}.samples"])


def test_linear_inversion_matches_reference():
    z = load_golden("recon_small.npz")
    for n in (1, 2, 3):
        hist = z[f"N{n}.hist"]
        assert np.abs(orc.linear_inversion_hist(hist, n, True) - z[f"N{n}.rho_rqc"]).max() < 1e-12
        assert np.abs(orc.linear_inversion_hist(hist, n, False) - z[f"N{n}.rho_ss"]).max() < 1e-12
        data = {name: ((np.repeat(np.arange(1 << n), hist[b])[:, None] >> np.arange(n)) & 1)
                for b, name in enumerate(orc.basis_strings(n))}
        assert np.abs(orc.linear_inversion_literal(data, n, True) - z[f"N{n}.rho_rqc"]).max() < 1e-12
        c = [orc.pauli_coefficient(p, data) for p in ("X" + "I" * (n - 1), "Z" * n, "I" * n)]
        assert np.allclose(c, z[f"N{n}.coeff_first"], atol=1e-15)


def test_datapoints_fixture_matches_reference():
    z = load_golden("datapoints_N3.npz")
    for i in range(int(z["n"][0])):
        rho = orc.linear_inversion_hist(z[f"r{i}.hist"], 3)
        assert np.abs(rho - z[f"r{i}.rho"]).max() < 1e-12
        assert np.allclose(orc.get_metrics(rho, 3), z[f"r{i}.metrics"], atol=1e-10)
        psi = z[f"r{i}.psi"]
        assert abs(orc.state_fidelity(psi, rho) - orc.state_fidelity(np.outer(psi, psi.conj()), rho)) < 1e-7


@pytest.mark.parametrize("counts,expected", [
    (({'0': 940, '1': 84}, {'0': 545, '1': 479}, {'0': 527, '1': 497}), 0.917969),     # NB c6:23-25 -> c9:38
    (({'0': 3785, '1': 311}, {'0': 2152, '1': 1944}, {'0': 2029, '1': 2067}), 0.924072),  # NB c7:26-28 -> c10:38
    (({'0': 931, '1': 93}, {'0': 564, '1': 460}, {'0': 445, '1': 579}), 0.909180),     # NB c9:32-34 -> c16:74
    (({'0': 952, '1': 84}, {'0': 525, '1': 499}, {'0': 454, '1': 570}), 0.918919),     # NB c13:34-36 -> notes.pdf p.5
])
def test_notebook_fidelity_known_answers(counts, expected):
    rho = orc.rho_from_single_qubit_counts(*counts)
    plus = np.array([1, 1]) / np.sqrt(2)
    assert abs(orc.state_fidelity(plus, rho) - expected) < 5e-7
    # the multi-qubit restatement agrees with the 1-qubit closed form (no PSD clipping needed here)
    hist = np.array([[c['0'], c['1']] for c in counts])
    assert np.abs(orc.linear_inversion_hist(hist, 1, psd=False) - rho).max() < 1e-15


def test_mixed_state_fidelity_properties():
    rng = np.random.default_rng(3)
    a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    r1 = a @ a.conj().T
    r1 /= np.trace(r1)
    b = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    r2 = b @ b.conj().T
    r2 /= np.trace(r2)
    assert abs(orc.state_fidelity(r1, r1) - 1) < 1e-9
    assert abs(orc.state_fidelity(r1, r2) - orc.state_fidelity(r2, r1)) < 1e-9
    psi = orc.haar_state(2, 1)
    assert abs(orc.state_fidelity(np.outer(psi, psi.conj()), r2) - orc.state_fidelity(psi, r2)) < 1e-7
