"""Synthetic data generator.  CPU: the oracle restatement against the independent Born-rule code the
recon fixtures were made with (oracle.born_probabilities, pinned through the reference's linear_inversion) and against
known answers (GHZ/Bell statistics of SS/data_gen.py's circuit).  GPU: kernels vs the oracle -- state and probabilities to
1e-12, histograms bit-exact under the same Philox stream."""
import numpy as np
import pytest
import torch

from oracle import ddqst_oracle as orc


def test_oracle_states_and_probabilities():
    for n in (1, 2, 3, 5):
        psi = orc.synth_state(n, "rqc", depth=6, seed=11)
        assert abs(np.vdot(psi, psi).real - 1) < 1e-12
        names = orc.basis_strings(n)
        ids = list(range(0, len(names), max(1, len(names) // 7)))
        p = orc.synth_probabilities(psi, n, ids)
        for k, b in enumerate(ids):
            assert np.abs(p[k] / p[k].sum() - orc.born_probabilities(psi, n, names[b])).max() < 1e-12
    # Bell state of SS/data_gen.py:22-26: perfectly correlated in ZZ and XX, anti-correlated in YY
    bell = orc.synth_state(2, "bell")
    names = orc.basis_strings(2)
    p = orc.synth_probabilities(bell, 2, [names.index("ZZ"), names.index("XX"), names.index("YY")])
    assert np.allclose(p[0], [0.5, 0, 0, 0.5]) and np.allclose(p[1], [0.5, 0, 0, 0.5]) and np.allclose(p[2], [0, 0.5, 0.5, 0])
    # depolarising mix and read-out flips keep the distribution normalised; full read-out error flips every bit
    q = orc.synth_probabilities(bell, 2, [names.index("ZZ")], p_depol=0.2)
    assert np.allclose(q[0], [0.45, 0.05, 0.05, 0.45])
    plus = orc.synth_state(3, "plus")
    zp = orc.synth_probabilities(plus, 3, [0], p_readout=1.0)          # XXX on |+++> is deterministic 000 -> flipped to 111
    assert np.allclose(zp[0], np.eye(8)[7])


def test_oracle_histograms_are_multinomial_draws():
    psi = orc.synth_state(3, "rqc", depth=4, seed=2)
    p = orc.synth_probabilities(psi, 3, list(range(27)))
    h = orc.synth_histograms(p, list(range(27)), 20001, seed=9)
    assert (h.sum(axis=1) == 20001).all()
    assert np.abs(h / 20001 - p / p.sum(axis=1, keepdims=True)).max() < 0.02
    h2 = orc.synth_histograms(p, list(range(27)), 20001, seed=10)
    assert not np.array_equal(h, h2)


@pytest.fixture(scope="module")
def dq():
    import ddqst_b200
    assert torch.cuda.is_available()
    return ddqst_b200


@pytest.mark.gpu
@pytest.mark.parametrize("n,kind,depth", [(1, "plus", 0), (2, "bell", 0), (3, "ghz", 0), (3, "rqc", 5), (5, "rqc", 8), (8, "rqc", 10),
                                          (10, "rqc", 3)])
def test_synth_state_matches_oracle(dq, n, kind, depth):
    got = dq.synth_state(n, kind, depth, seed=1234567890123).cpu().numpy()
    want = orc.synth_state(n, kind, depth, seed=1234567890123)
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("n,shots,noise,rate", [(1, 1, "ideal", 0.0), (2, 1001, "ideal", 0.0), (3, 50_000, "readout", 0.03),
                                                (3, 4096, "depolarizing", 0.1), (5, 10_000, "ideal", 0.0)])
def test_born_histograms_match_oracle(dq, n, shots, noise, rate):
    psi = orc.synth_state(n, "rqc", depth=6, seed=5)
    psi_d = torch.from_numpy(psi).cuda()
    hist, probs = dq.born_histograms(psi_d, n, shots, seed=77, noise_type=noise, error_rate=rate, return_probs=True)
    ids = list(range(3 ** n))
    want_p = orc.synth_probabilities(psi, n, ids, p_depol=rate if noise == "depolarizing" else 0.0,
                                     p_readout=rate if noise == "readout" else 0.0)
    got_p = probs.cpu().numpy()
    assert np.abs(got_p - want_p).max() < 1e-12
    h = hist.view(torch.int32).cpu().numpy()
    assert (h.sum(axis=1) == shots).all()
    # draws: the kernel's CDF comes from a blocked scan, numpy's from a sequential cumsum; a draw can only differ if u lands
    # within ~1e-16 of a CDF step, which does not happen at these sizes -> exact equality
    assert np.array_equal(h, orc.synth_histograms(got_p, ids, shots, seed=77))


@pytest.mark.gpu
def test_generate_synthetic_data_surface_and_full_size_properties(dq):
    # reference return convention (SS/data_gen.py:55-63)
    data, bases, psi = dq.generate_synthetic_data(2, "bell", 2000, as_counts=True, seed=3)
    assert bases == orc.basis_strings(2) and len(data) == 9
    assert data[8]["basis_str"] == "ZZ" and data[8]["basis_idx"] == 8 and set(data[8]["counts"]) <= {"00", "11"}
    assert sum(data[8]["counts"].values()) == 2000
    ds = dq.QuantumStateDataset(data, 2, device="cuda")           # SS phase: flat list of measurement dicts
    assert len(ds) == 9 * 2000
    # subset of bases + shot-chunked launch agree with the full call (draws are keyed by basis and draw index only)
    psi8 = dq.synth_state(8, "rqc", 12, seed=4)
    full = dq.born_histograms(psi8, 8, 300_000, seed=6)
    sub = dq.born_histograms(psi8, 8, 300_000, seed=6, bases=[0, 4000, 6560])
    assert torch.equal(sub.view(torch.int32), full.view(torch.int32)[[0, 4000, 6560]])
    assert int(full.view(torch.int32).to(torch.int64).sum()) == 6561 * 300_000
    # linear inversion of the generated data recovers the state (fidelity -> 1 with shots)
    f = dq.state_fidelity(psi8, dq.linear_inversion(full, 8))
    assert f > 0.75, f          # first-compatible-basis linear inversion at 3e5 shots/basis (0.90 at 1e6)
    with pytest.raises(ValueError):
        dq.born_histograms(psi8, 8, 10, noise_type="thermal")
