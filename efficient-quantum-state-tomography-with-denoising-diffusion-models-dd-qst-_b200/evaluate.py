"""The reference's evaluation loop (RQC/evaluate.py:20-102) on the native path: per circuit record, linear inversion of
the raw counts vs of D3PM-generated shots, fidelity against the clean state, entropies and the Z-basis bias, collected
into the ``metrics.csv`` schema of the reference (ID, Depth, Raw_Fidelity, D3PM_Fidelity, Raw_Entropy, D3PM_Entropy,
Bias).  Plots (RQC/evaluate.py:104-118) are out of scope.

What changes under the hood: raw counts stay counts (``format_raw_counts_for_inversion`` returns ``dict basis -> uint32[2^N]``
rows instead of expanded sample matrices), and the 3^N ``p_sample`` calls of RQC/evaluate.py:82-84 are ONE
``diffusion.sample(all_bases, shots)`` launch with the histogram fused.
"""
from __future__ import annotations

import csv
import os
from itertools import product

import numpy as np
import torch

from .dataset import load_circuit_records, state_vector_of
from .reconstruct import DensityMatrix, recon_report

CSV_COLUMNS = ["ID", "Depth", "Raw_Fidelity", "D3PM_Fidelity", "Raw_Entropy", "D3PM_Entropy", "Bias"]


def format_raw_counts_for_inversion(measurements_list, num_qubits: int, device="cuda") -> dict:
    """RQC/evaluate.py:20-30 without the expansion: -> dict basis -> counts row uint32[2^N] (outcome index bit i = qubit i:
    the reference flips each key so that column i is qubit i), in the list's own order -- linear_inversion keeps the
    reference's first-compatible-basis-in-dict-order rule (RQC/reconstruct.py:32-38); a repeated basis replaces the
    earlier entry, as the reference's dict assignment does."""
    formatted = {}
    for m in measurements_list:
        row = np.zeros(1 << num_qubits, dtype=np.int64)
        for key, cnt in m["counts"].items():
            row[int(key.replace(" ", ""), 2)] += int(cnt)
        formatted[m["basis"]] = torch.from_numpy(row.astype(np.uint32).view(np.int32)).view(torch.uint32).to(device)
    return formatted


def calculate_z_bias(counts, num_qubits: int) -> float:
    """RQC/evaluate.py:32-38: fraction of 0 bits among all bits of the Z...Z-basis samples; 0.5 when that basis is
    absent.  ``counts``: uint32[3^N, 2^N] in product order (all-Z = last row), or a dict basis -> counts row / sample matrix."""
    if isinstance(counts, dict):
        key = "Z" * num_qubits
        if key not in counts:
            return 0.5
        row = counts[key]
        if getattr(row, "ndim", 1) == 2:                      # the reference's sample-matrix form
            s = row.cpu().numpy() if torch.is_tensor(row) else np.asarray(row)
            return float(np.sum(s == 0) / s.size)
    else:
        row = counts[-1]
    row = (row.view(torch.int32) if row.dtype == torch.uint32 else row).to(torch.int64)
    shots = int(row.sum().item())
    if shots == 0:
        return float("nan")                                   # the reference divides 0 by samples.size == 0
    pop = torch.tensor([bin(s).count("1") for s in range(1 << num_qubits)], dtype=torch.int64, device=row.device)
    return int((row * (num_qubits - pop)).sum().item()) / (shots * num_qubits)


@torch.no_grad()
def evaluate_records(diffusion, records, num_qubits: int, shots_infer: int, verbose: bool = False):
    """The loop of RQC/evaluate.py:70-97 -> list of row dicts (CSV_COLUMNS)."""
    device = diffusion.device
    all_bases = list(range(3 ** num_qubits))
    rows = []
    for i, state_data in enumerate(records):
        target_dm = DensityMatrix(np.outer(state_vector_of(state_data), state_vector_of(state_data).conj()))   # :71
        depth = state_data.get("depth", 0)
        raw_input = format_raw_counts_for_inversion(state_data["measurements"], num_qubits, device)
        # :75-78 and :85-88 -- linear_inversion + state_fidelity + get_metrics share one eigendecomposition (recon_report)
        raw = recon_report(raw_input, num_qubits, target_dm)
        fid_raw, s_raw = raw.fidelity, raw.entropy
        syn_hist, _ = diffusion.sample(all_bases, shots_infer, shot_offset=i * shots_infer)                    # :80-84
        d3pm = recon_report(syn_hist, num_qubits, target_dm)
        fid_d3pm, s_d3pm = d3pm.fidelity, d3pm.entropy
        bias = calculate_z_bias(syn_hist, num_qubits)
        if verbose:
            print(f"State {i} (D={depth}): Raw={fid_raw:.3f} -> D3PM={fid_d3pm:.3f}")
        rows.append({"ID": i, "Depth": depth, "Raw_Fidelity": fid_raw, "D3PM_Fidelity": fid_d3pm, "Raw_Entropy": s_raw,
                     "D3PM_Entropy": s_d3pm, "Bias": bias})
    return rows


def write_metrics_csv(rows, out_dir: str) -> str:
    """``{out_dir}/metrics.csv`` with the reference's columns (RQC/evaluate.py:100-102)."""
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "metrics.csv")
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=CSV_COLUMNS)
        w.writeheader()
        w.writerows(rows)
    return path


def evaluate(model, diffusion, eval_data_path: str, num_qubits: int, shots_infer: int = 10000, out_dir: str = "results",
             verbose: bool = True):
    """RQC/evaluate.py:40-102 minus argparse and plots: load the eval subset, run the loop, write metrics.csv."""
    if not os.path.exists(eval_data_path):
        raise FileNotFoundError(f"Eval file not found: {eval_data_path}")                                     # :46-47
    records = load_circuit_records(eval_data_path)
    model.eval()
    rows = evaluate_records(diffusion, records, num_qubits, shots_infer, verbose)
    return rows, write_metrics_csv(rows, out_dir)
