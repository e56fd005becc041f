"""Evaluation driver: the loop of RQC/evaluate.py:70-97 on the native path.  The raw-data columns are
pinned by what the unmodified reference computed for five shipped N=3 records (tests/golden/datapoints_N3.npz: rho and
get_metrics through the reference's reconstruct.py); the D3PM columns are checked through the oracle on the very
histograms the sampler produced (the sampler has its own parity tests)."""
import csv
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ddqst_oracle as orc


def fixture_records(z, n=3):
    names = orc.basis_strings(n)
    recs = []
    for i in range(int(z["n"][0])):
        hist = z[f"r{i}.hist"]
        meas = [{"basis": names[b], "counts": {format(s, f"0{n}b"): int(hist[b, s]) for s in range(1 << n) if hist[b, s]}}
                for b in range(len(names))]
        recs.append({"id": int(z[f"r{i}.id_depth"][0]), "depth": int(z[f"r{i}.id_depth"][1]), "clean_state_vec": z[f"r{i}.psi"],
                     "measurements": meas})
    return recs


def test_metrics_csv_schema(tmp_path):
    import ddqst_b200 as dq
    rows = [{"ID": 0, "Depth": 4, "Raw_Fidelity": 0.9, "D3PM_Fidelity": 0.95, "Raw_Entropy": 0.1, "D3PM_Entropy": 0.05, "Bias": 0.5}]
    path = dq.write_metrics_csv(rows, str(tmp_path / "results"))
    assert os.path.basename(path) == "metrics.csv"
    got = list(csv.reader(open(path)))
    assert got[0] == ["ID", "Depth", "Raw_Fidelity", "D3PM_Fidelity", "Raw_Entropy", "D3PM_Entropy", "Bias"]     # RQC/evaluate.py:93-97
    assert got[1][:2] == ["0", "4"] and float(got[1][3]) == 0.95
    with pytest.raises(FileNotFoundError):
        dq.evaluate(None, None, str(tmp_path / "missing.pt"), 3)


@pytest.mark.gpu
def test_evaluate_records_match_reference_and_oracle():
    import ddqst_b200 as dq
    z = load_golden("datapoints_N3.npz")
    n, T, shots = 3, 20, 3000
    recs = fixture_records(z, n)
    torch.manual_seed(0)
    model = dq.ConditionalD3PM(n, 27, T, 16, 64, 2).cuda()
    diff = dq.DiscreteDiffusion(model, T, "cuda", seed=21, precision="bf16")
    rows = dq.evaluate_records(diff, recs, n, shots)
    assert dq._lib.load().ddqst_debug_tc_status() == 0
    assert [r["ID"] for r in rows] == list(range(len(recs)))
    for i, r in enumerate(rows):
        psi = z[f"r{i}.psi"]
        # raw columns: against the reference's own rho / metrics for this record
        assert r["Depth"] == int(z[f"r{i}.id_depth"][1])
        assert abs(r["Raw_Fidelity"] - orc.state_fidelity(psi, z[f"r{i}.rho"])) < 1e-5
        assert abs(r["Raw_Entropy"] - z[f"r{i}.metrics"][1]) < 1e-5
        # D3PM columns: oracle recon on the histograms the sampler produced for this record (shot_offset = i * shots)
        hist = diff.sample(list(range(27)), shots, shot_offset=i * shots)[0].view(torch.int32).cpu().numpy()
        rho = orc.linear_inversion_hist(hist, n)
        assert abs(r["D3PM_Fidelity"] - orc.state_fidelity(psi, rho)) < 1e-5
        assert abs(r["D3PM_Entropy"] - orc.get_metrics(rho, n)[1]) < 1e-5
        zrow = hist[-1]
        zeros = sum(int(zrow[s]) * (n - bin(s).count("1")) for s in range(1 << n))
        assert abs(r["Bias"] - zeros / (shots * n)) < 1e-12
    # the reference's sample-matrix form gives the same bias
    samples = {"ZZZ": ((np.repeat(np.arange(8), hist[-1])[:, None] >> np.arange(n)) & 1)}
    assert abs(dq.calculate_z_bias(samples, n) - rows[-1]["Bias"]) < 1e-12
    assert dq.calculate_z_bias({"XXX": samples["ZZZ"]}, n) == 0.5                                    # RQC/evaluate.py:38


@pytest.mark.gpu
def test_raw_counts_keep_first_compatible_basis_order():
    """An incomplete, shuffled measurement list: linear_inversion must pick, for every Pauli string, the FIRST compatible
    basis in the list's order (RQC/reconstruct.py:32-38) and 0.0 when none is -- checked against the literal 4^N loop."""
    import ddqst_b200 as dq
    n = 3
    rng = np.random.default_rng(3)
    psi = orc.haar_state(n, 5)
    names = orc.basis_strings(n)
    order = rng.permutation(len(names))[:17]                  # 10 bases missing, the rest shuffled
    meas, data = [], {}
    for b in order:
        p = orc.born_probabilities(psi, n, names[b])
        h = rng.multinomial(700 + int(b), p)
        meas.append({"basis": names[b], "counts": {format(s, f"0{n}b"): int(h[s]) for s in range(1 << n) if h[s]}})
        data[names[b]] = ((np.repeat(np.arange(1 << n), h)[:, None] >> np.arange(n)) & 1).astype(np.int64)
    got = dq.linear_inversion(dq.format_raw_counts_for_inversion(meas, n), n).data
    want = orc.linear_inversion_literal(data, n)
    assert np.abs(got - want).max() < 1e-5
