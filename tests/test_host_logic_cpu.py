"""Host-side logic that needs no GPU: the counts-table builder behind QuantumStateDataset, the evaluation helpers'
pure-Python branches, the synthetic generator's record formatting, and the sharding plan."""
import numpy as np
import pytest
import torch

import ddqst_b200 as dq
from ddqst_b200.dataset import counts_table
from oracle import ddqst_oracle as orc


def test_counts_table_matches_oracle_rows_and_reference_conventions():
    recs = [{"measurements": [{"basis": "XZ", "counts": {"10": 3, "01": 2}}, {"basis": "QQ", "counts": {"00": 9}},
                              {"basis": "ZZ", "counts": {"11": 5}}]},
            {"measurements": []},
            {"basis_str": "YY", "basis_idx": 4, "counts": {"00": 7}}]                 # SS phase: flat measurement dict
    hist, row_basis, b2i = counts_table(recs, 2)
    assert b2i["XX"] == 0 and b2i["ZZ"] == 8 and list(b2i)[:4] == ["XX", "XY", "XZ", "YX"]    # product order (RQC/dataset.py:43)
    assert row_basis == [2, 8, 4]                                                    # unknown basis "QQ" skipped (RQC/dataset.py:54)
    # key 'q1 q0' = '10' -> bits reversed -> qubit0 = 0, qubit1 = 1 -> outcome index 2 (bit i = qubit i)
    assert hist[0].tolist() == [0, 2, 3, 0] and hist[1].tolist() == [0, 0, 0, 5] and hist[2].tolist() == [7, 0, 0, 0]
    want, rb, _ = orc.counts_rows_from_records([recs[0]], 2)
    assert np.array_equal(want[0], hist[0]) and list(rb) == [2, 8]
    ds = dq.QuantumStateDataset(recs, 2, device="cpu")
    assert len(ds) == 5 + 5 + 7 and ds.n_rows == 3
    with pytest.raises(RuntimeError):
        ds[0]                                                                        # no CPU path for the unroll itself
    with pytest.raises(TypeError):
        dq.QuantumStateDataset(123, 2)


def test_format_raw_counts_and_z_bias_host_branches():
    meas = [{"basis": "ZZ", "counts": {"00": 6, "11": 2}}, {"basis": "XX", "counts": {"01": 4}},
            {"basis": "ZZ", "counts": {"10": 8}}]                                      # repeated basis replaces the earlier one
    out = dq.format_raw_counts_for_inversion(meas, 2, device="cpu")
    assert list(out) == ["ZZ", "XX"]                                                   # dict order = list order
    assert out["ZZ"].view(torch.int32).tolist() == [0, 0, 8, 0]
    assert out["XX"].view(torch.int32).tolist() == [0, 4, 0, 0]
    # reference form (sample matrices): fraction of zeros over all bits of the Z..Z samples (RQC/evaluate.py:32-38)
    samples = {"ZZ": np.array([[0, 0], [0, 1], [1, 1], [0, 0]])}
    assert dq.calculate_z_bias(samples, 2) == 5 / 8
    assert dq.calculate_z_bias({"XX": samples["ZZ"]}, 2) == 0.5
    # counts-row form gives the same number: outcomes 00, 10 (qubit1=1 -> index 2), 11, 00
    row = torch.tensor([2, 0, 1, 1], dtype=torch.int32)
    assert dq.calculate_z_bias({"ZZ": row}, 2) == 5 / 8


def test_counts_records_format_and_basis_list():
    assert dq.get_basis_combinations(2) == ["XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"]      # SS/data_gen.py:9-12
    hist = torch.tensor([[1, 0, 2, 0], [0, 0, 0, 5]], dtype=torch.int32).view(torch.uint32)
    recs = dq.counts_records(hist, 2, bases=[0, 8])
    assert recs[0] == {"basis_str": "XX", "basis": "XX", "basis_idx": 0, "counts": {"00": 1, "10": 2}}   # keys 'q1q0'
    assert recs[1]["basis_str"] == "ZZ" and recs[1]["counts"] == {"11": 5}
    # round trip through the dataset's own parser
    h2, rb, _ = counts_table(recs, 2)
    assert np.array_equal(h2, hist.view(torch.int32).numpy()) and rb == [0, 8]
    with pytest.raises(ValueError):
        dq.synth_state(2, "cat")


def test_sharding_plan_covers_everything_once():
    from ddqst_b200.distributed import plan, shard_range
    for n_bases, shots, world in [(6561, 1000, 8), (27, 100, 4), (3, 1001, 8), (1, 10, 2), (5, 7, 8)]:
        seen = np.zeros((n_bases, shots), dtype=np.int64)
        for r in range(world):
            b0, b1, s0, s1 = plan(n_bases, shots, r, world)
            seen[b0:b1, s0:s1] += 1
        assert (seen == 1).all(), (n_bases, shots, world)
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_guard_band_checker_accepts_rounding_and_rejects_real_errors():
    """tests/guard.py itself: bits drawn from logits that differ from the oracle's by a measured error pass; a wrong bit
    where the draw is not marginal, a wrong qubit lane or a wrong Philox counter fail."""
    import guard
    torch.manual_seed(0)
    N, T, B, seed, basis, off = 5, 20, 512, 3, 4, 17
    betas, Q = orc.cosine_schedule(T)
    lin_b, lin_Q = orc.linear_schedule(T)
    g = torch.Generator().manual_seed(1)
    for mode, sched in (("posterior", (betas, Q)), ("renoise", (lin_b, lin_Q))):
        for t in (T, 7, 1):
            x_t = torch.randint(0, 2, (B, N), generator=g)
            want = 2.0 * torch.randn(B, N, 2, generator=g)
            got_logits = want + 2e-2 * torch.randn(B, N, 2, generator=g)            # a bf16-sized logit error
            got = guard.oracle_reverse_step(mode, got_logits, x_t, t, sched[0], sched[1], seed, basis, off, N)
            mm, band, total = guard.assert_draws_in_guard_band(got, got_logits, want, mode, x_t, t, sched[0], sched[1], seed,
                                                               basis, off, N)
            assert total == B * N and mm <= band < 0.05 * total
            # (a) flip the bit whose draw is farthest from marginal -> outside the band
            nominal = guard.oracle_reverse_step(mode, want, x_t, t, sched[0], sched[1], seed, basis, off, N)
            lo = guard.oracle_reverse_step(mode, want - torch.tensor([0.5, -0.5]), x_t, t, sched[0], sched[1], seed, basis, off, N)
            hi = guard.oracle_reverse_step(mode, want + torch.tensor([0.5, -0.5]), x_t, t, sched[0], sched[1], seed, basis, off, N)
            solid = torch.nonzero((lo == hi) & (lo == nominal))
            bad = got.clone()
            r, q = (int(v) for v in solid[0])
            bad[r, q] ^= 1
            with pytest.raises(AssertionError, match="OUTSIDE the guard band"):
                guard.assert_draws_in_guard_band(bad, got_logits, want, mode, x_t, t, sched[0], sched[1], seed, basis, off, N)
            # (b) qubit lanes swapped, (c) wrong shot offset in the Philox counter
            with pytest.raises(AssertionError):
                guard.assert_draws_in_guard_band(got.flip(1), got_logits, want, mode, x_t, t, sched[0], sched[1], seed, basis, off, N)
            shifted = guard.oracle_reverse_step(mode, got_logits, x_t, t, sched[0], sched[1], seed, basis, off + 1, N)
            with pytest.raises(AssertionError):
                guard.assert_draws_in_guard_band(shifted, got_logits, want, mode, x_t, t, sched[0], sched[1], seed, basis, off, N)


def test_check_index_contract():
    from ddqst_b200 import _lib
    _lib.check_index(torch.tensor([0, 5, 8]), 9, "basis")
    _lib.check_index([], 9, "basis")
    _lib.check_index(torch.zeros(0, dtype=torch.long), 9, "basis")
    for bad in (torch.tensor([0, 9]), [-1, 2], 9, torch.tensor([-3])):
        with pytest.raises(IndexError):
            _lib.check_index(bad, 9, "basis")
    with pytest.raises(IndexError):
        _lib.check_index(torch.tensor([0]), 101, "t", lo=1)


def test_build_stamp_is_content_keyed():
    """_build: the library's stamp is a hash of csrc/ + include/ddqst.h + flags, so a library built from other sources is
    detected without relying on file times."""
    from ddqst_b200 import _build, _lib
    _lib.load()                                        # builds (incrementally) when nvcc is available, else verifies the stamp
    assert _build.stamp_matches()
    assert len(_build.expected_stamp()) == 64
