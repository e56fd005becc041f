"""Dataset unrolling.  CPU: the oracle restatement against what the unmodified reference's
QuantumStateDataset produced (tests/golden/dataset_N3.npz).  GPU: the device unroll / batch sampler through the C ABI
against the oracle, bit-exact (integer work)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ddqst_oracle as orc


def records_from_fixture(z):
    """Rebuild circuit records (the reference's input format) from the stored rows, keeping each counts dict's key order."""
    n = int(z["n_qubits"][0])
    names = orc.basis_strings(n)
    recs = []
    for r in range(z["row_hist"].shape[0]):
        counts = {format(int(s), f"0{n}b"): int(z["row_hist"][r, s]) for s in z["key_order"][r] if s >= 0}
        if r % len(names) == 0:
            recs.append({"measurements": []})
        recs[-1]["measurements"].append({"basis": names[int(z["row_basis"][r])], "counts": counts})
    return recs, n


def test_oracle_unroll_matches_reference_dataset():
    z = load_golden("dataset_N3.npz")
    recs, n = records_from_fixture(z)
    hist, row_basis, order = orc.counts_rows_from_records(recs, n)
    assert np.array_equal(hist, z["row_hist"]) and np.array_equal(row_basis, z["row_basis"])
    bits, basis = orc.counts_unroll(hist, row_basis, n)
    assert bits.shape[0] == int(z["ref_len"][0]) == 138240
    packed = (bits << np.arange(n)).sum(axis=1)
    # same basis sequence (rows are contiguous in both orders); inside a row the reference follows the dict's key order
    assert np.array_equal(basis, z["ref_basis"].astype(np.int64))
    starts = np.concatenate([[0], np.cumsum(hist.sum(axis=1))])
    for r in range(hist.shape[0]):
        ref_row = z["ref_packed"][starts[r]:starts[r + 1]].astype(np.int64)
        assert np.array_equal(np.sort(ref_row), packed[starts[r]:starts[r + 1]])          # equal as multisets, ours sorted
        want = np.repeat([s for s in z["key_order"][r] if s >= 0], [hist[r, s] for s in z["key_order"][r] if s >= 0])
        assert np.array_equal(ref_row, want)                                               # the reference's exact order
    # endianness: column i of the reference's data_tensor is qubit i = bit i of the outcome index
    assert list(z["ref_item_17"][:n]) == [(int(z["ref_packed"][17]) >> i) & 1 for i in range(n)]


@pytest.mark.parametrize("total", [1, 2, 5, 16, 17, 1000, 138240, (1 << 20) + 3])
def test_feistel_is_a_bijection(total):
    n = min(total, 200_000)
    if n == total:
        p = orc.feistel_perm(np.arange(total), total, seed=99, epoch=3)
        assert np.array_equal(np.sort(p), np.arange(total))
    else:
        p = orc.feistel_perm(np.arange(n), total, seed=99, epoch=3)
        assert p.min() >= 0 and p.max() < total and np.unique(p).size == n
    if total > 16:
        q = orc.feistel_perm(np.arange(min(total, 1000)), total, seed=99, epoch=4)
        assert not np.array_equal(p[:q.size], q)                                           # epochs differ


def test_oracle_batches_cover_an_epoch_once():
    z = load_golden("dataset_N3.npz")
    hist, row_basis = z["row_hist"][:27], z["row_basis"][:27]
    total, B = int(hist.sum()), 4096
    seen = np.zeros((27, 8), dtype=np.int64)
    for step in range((total + B - 1) // B):
        cnt = min(B, total - step * B)
        s, b = orc.counts_batch(hist, row_basis, 3, step * B, cnt, seed=5, epoch=0)
        np.add.at(seen, (b, s), 1)
    assert np.array_equal(seen, hist)        # rows 0..26 are bases 0..26 of one circuit


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dq():
    import ddqst_b200
    assert torch.cuda.is_available()
    return ddqst_b200


@pytest.mark.gpu
def test_device_dataset_matches_oracle(dq):
    z = load_golden("dataset_N3.npz")
    recs, n = records_from_fixture(z)
    ds = dq.QuantumStateDataset(recs, n, device="cuda", seed=5)
    assert len(ds) == int(z["ref_len"][0])
    bits, basis = orc.counts_unroll(z["row_hist"], z["row_basis"], n)
    assert np.array_equal(ds.data_tensor.cpu().numpy(), bits)
    assert np.array_equal(ds.basis_tensor.cpu().numpy(), basis)
    b17, k17 = ds[17]
    assert np.array_equal(b17.cpu().numpy(), bits[17]) and int(k17) == int(basis[17])
    with pytest.raises(IndexError):
        ds[len(ds)]
    # shuffled batches: bit-exact against the oracle's Feistel permutation, including a batch that wraps the epoch end
    total, B = len(ds), 4096
    for step in (0, 1, 33, total // B, total // B + 1):
        x0, kb = ds.batch(step, B)
        epoch, off = divmod(step * B, total)
        s, b = orc.counts_batch(z["row_hist"], z["row_basis"], n, off, B, seed=5, epoch=epoch)
        assert np.array_equal(x0.cpu().numpy().astype(np.int64), s)
        assert np.array_equal(kb.cpu().numpy().astype(np.int64), b)
    # one epoch of batches visits every shot exactly once
    seen = torch.zeros(z["row_hist"].shape[0] * 0 + 27, 8, dtype=torch.int64, device="cuda")
    per_basis = np.zeros((27, 8), dtype=np.int64)
    np.add.at(per_basis, (np.repeat(z["row_basis"], 8), np.tile(np.arange(8), z["row_hist"].shape[0])), z["row_hist"].reshape(-1))
    for step in range(ds.batches_per_epoch(B)):
        cnt = min(B, total - step * B)
        x0, kb, _ = ds._gather(step * B, cnt, True, 0, False)
        seen.index_put_((kb.long(), x0.long()), torch.ones(cnt, dtype=torch.int64, device="cuda"), accumulate=True)
    assert np.array_equal(seen.cpu().numpy(), per_basis)


@pytest.mark.gpu
def test_device_dataset_edge_cases(dq):
    """Empty rows, a single shot, N=10 (1024 outcomes per row, multi-chunk scan), ragged totals."""
    rng = np.random.default_rng(0)
    n = 10
    names = orc.basis_strings(n)
    recs = [{"measurements": []}]
    hist = np.zeros((40, 1 << n), dtype=np.int64)
    for r in range(40):
        if r % 7 == 3:
            counts = {}                                      # an empty measurement record
        else:
            ks = rng.choice(1 << n, size=rng.integers(1, 300), replace=False)
            counts = {format(int(k), f"0{n}b"): int(rng.integers(1, 50)) for k in ks}
        for k, c in counts.items():
            hist[r, int(k, 2)] = c
        recs[0]["measurements"].append({"basis": names[r * 1000], "counts": counts})
    ds = dq.QuantumStateDataset(recs, n, device="cuda", seed=1)
    row_basis = np.array([r * 1000 for r in range(40)])
    bits, basis = orc.counts_unroll(hist, row_basis, n)
    assert len(ds) == bits.shape[0]
    assert np.array_equal(ds.data_tensor.cpu().numpy(), bits) and np.array_equal(ds.basis_tensor.cpu().numpy(), basis)
    x0, kb = ds.batch(3, 777)
    s, b = orc.counts_batch(hist, row_basis, n, (3 * 777) % len(ds), 777, seed=1, epoch=(3 * 777) // len(ds))
    assert np.array_equal(x0.cpu().numpy().astype(np.int64), s) and np.array_equal(kb.cpu().numpy().astype(np.int64), b)
    one = dq.QuantumStateDataset([{"measurements": [{"basis": "ZZ", "counts": {"10": 1}}]}], 2, device="cuda")
    assert len(one) == 1 and one[0][0].tolist() == [0, 1] and int(one[0][1]) == 8
    x0, kb = one.batch(5, 4)
    assert x0.tolist() == [2, 2, 2, 2] and kb.tolist() == [8, 8, 8, 8]
    with pytest.raises(FileNotFoundError):
        dq.QuantumStateDataset("/nonexistent/path.pt", 2)
    # straight from a device counts table (no dict round trip)
    tab = torch.from_numpy(hist.astype(np.int32)).cuda()
    ds2 = dq.QuantumStateDataset.from_counts_table(tab, n, row_basis=row_basis, seed=1)
    assert len(ds2) == len(ds)
    x2, k2 = ds2.batch(3, 777)
    x1, k1 = ds.batch(3, 777)
    assert torch.equal(x2, x1) and torch.equal(k2, k1)
