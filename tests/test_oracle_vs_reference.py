"""Pins the oracle against the UNMODIFIED reference imported from /root/reference (build container only;
skipped on the GPU box, where the committed golden fixtures take over)."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import ddqst_oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def phases():
    return {"RQC": rh.load_phase("RQC"), "SS": rh.load_phase("SS")}


def test_rqc_path_bit_identical(phases):
    R = phases["RQC"]
    N, T = 4, 12
    torch.manual_seed(3)
    m = R["model"].ConditionalD3PM(N, 81, T, 8, 32, 3)
    d = R["diffusion"].DiscreteDiffusion(m, T, "cpu")
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    betas, Q = orc.cosine_schedule(T)
    assert torch.equal(betas, d.betas) and torch.equal(Q, d.Q_bar)
    for basis, off in ((0, 0), (80, 12345)):
        st = rh.InjectedStream("posterior", 42, basis, N, T, 96, offset=off)
        with rh.injected(st):
            ref = d.p_sample(96, basis, N)
        assert torch.equal(ref, orc.p_sample_posterior(sd, betas, Q, 96, basis, N, 42, shot_offset=off))
    x0 = torch.randint(0, 2, (40, N))
    t = torch.randint(1, T + 1, (40,))
    st = rh.InjectedStream("q_cumulative", 9, 2, N, T, 40)
    with rh.injected(st):
        ref = d.q_sample(x0, t)
    assert torch.equal(ref, orc.q_sample_cumulative(Q, x0, t, 9, 2))


def test_ss_path_bit_identical(phases):
    S = phases["SS"]
    N, T = 2, 15
    torch.manual_seed(4)
    m = S["model"].ConditionalD3PM(N, 9, T, 8, 32, 2)
    d = S["diffusion"].DiscreteDiffusion(m, T, "cpu")
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    _, Q = orc.linear_schedule(T)
    st = rh.InjectedStream("renoise", 7, 4, N, T, 128, offset=5)
    with rh.injected(st):
        ref = d.p_sample(128, 4, N)
    assert torch.equal(ref, orc.p_sample_renoise(sd, Q, 128, 4, N, 7, shot_offset=5))


def test_reconstruction_identical(phases):
    rng = np.random.default_rng(1)
    N = 3
    psi = orc.haar_state(N, 5)
    data, hist = {}, np.zeros((27, 8), np.int64)
    for b, name in enumerate(orc.basis_strings(N)):
        s = rng.choice(8, size=300, p=orc.born_probabilities(psi, N, name))
        data[name] = ((s[:, None] >> np.arange(N)) & 1).astype(np.int64)
        hist[b] = orc.histogram(data[name], N)
    for tag, rev in (("RQC", True), ("SS", False)):
        ref = phases[tag]["reconstruct"].linear_inversion(data, N).data
        assert np.abs(ref - orc.linear_inversion_hist(hist, N, rev)).max() < 1e-13
    # a basis dict that is NOT in product order changes which basis feeds a Pauli (first-compatible rule)
    shuffled = dict(reversed(list(data.items())))
    ref = phases["RQC"]["reconstruct"].linear_inversion(shuffled, N).data
    assert np.abs(ref - orc.linear_inversion_literal(shuffled, N, True)).max() < 1e-13


def test_shipped_datapoints_directory_loads_through_the_dataset_surface():
    """Config C3's data: the whole Datapoints/rqc_N3_data directory (21 .pt shards that reference qiskit classes) through
    QuantumStateDataset's own loader (no GPU needed for the counts table): 363 circuits x 27 bases x 1024 shots."""
    import os
    import numpy as np
    import ddqst_b200 as dq
    from oracle import ref_harness as rh
    path = os.path.join(rh.REF_ROOT, "Datapoints", "rqc_N3_data")
    if not os.path.isdir(path):
        pytest.skip("reference Datapoints not mounted")
    ds = dq.QuantumStateDataset(path, 3, device="cpu")
    assert len(ds) == 363 * 27 * 1024 == 10_036_224
    assert ds.n_rows == 363 * 27
    h = ds.hist.view(torch.int32).numpy()
    assert (h.sum(axis=1) == 1024).all()
    assert np.array_equal(ds.row_basis.numpy().reshape(363, 27), np.tile(np.arange(27), (363, 1)))
    # the same records through the oracle's restatement of RQC/dataset.py
    recs = rh.load_datapoints(os.path.join(path, "part_0.pt"))
    want, rb, _ = orc.counts_rows_from_records(recs, 3)
    assert np.array_equal(h[: want.shape[0]], want)
