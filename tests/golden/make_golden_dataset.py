"""Golden fixture for the dataset row, produced by RUNNING THE UNMODIFIED REFERENCE
RQC_dataset_building_phase/dataset.py on five records of the shipped Datapoints/rqc_N3_data.

    python tests/golden/make_golden_dataset.py        (needs /root/reference)

Stored (tests/golden/dataset_N3.npz): the records' counts as rows (in the reference's iteration order, with each counts
dict's own key order), and what the reference's QuantumStateDataset made of them: len, the packed data_tensor rows
(outcome index with bit i = column i) and basis_tensor.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ddqst_oracle as orc  # noqa: E402
from oracle import ref_harness as rh    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert rh.available(), "needs /root/reference"
    D = rh.load_phase("RQC", names=("dataset",))["dataset"]
    recs = rh.load_datapoints(os.path.join(rh.REF_ROOT, "Datapoints/rqc_N3_data/part_0.pt"))[:4]
    recs += rh.load_datapoints(os.path.join(rh.REF_ROOT, "Datapoints/rqc_N3_data/part_20.pt"))[:1]
    n = 3
    with contextlib.redirect_stdout(io.StringIO()):
        ds = D.QuantumStateDataset(recs, n)                      # the reference itself
    data = ds.data_tensor.numpy()
    packed = (data << np.arange(n)).sum(axis=1).astype(np.uint8)
    hist, row_basis, order = orc.counts_rows_from_records(recs, n)
    key_order = np.full((len(order), 1 << n), -1, dtype=np.int64)
    for i, ko in enumerate(order):
        key_order[i, :len(ko)] = ko
    np.savez_compressed(os.path.join(OUT, "dataset_N3.npz"), n_qubits=np.array([n]), ref_len=np.array([len(ds)]),
                        ref_packed=packed, ref_basis=ds.basis_tensor.numpy().astype(np.uint8),
                        ref_item_17=np.concatenate([ds[17][0].numpy(), ds[17][1].numpy().reshape(1)]),
                        row_hist=hist, row_basis=row_basis, key_order=key_order)
    print("wrote dataset_N3.npz:", len(ds), "shots,", hist.shape[0], "rows")


if __name__ == "__main__":
    main()
