"""B200-native implementation of the DD-QST generative-tomography hot path.

Import name: ``ddqst_b200`` (see ddqst_b200.py at the repo root; this directory carries the long project name).
The public names are the reference's own: ConditionalD3PM, DiscreteDiffusion (q_sample / p_sample),
linear_inversion, get_coefficient, get_pauli_matrix, make_positive_semidefinite, get_metrics, state_fidelity,
plus the batched ``DiscreteDiffusion.sample(bases, n_shots)``, ``train_step`` and the multi-GPU helpers.
"""
from . import _lib
from ._build import build
from .dataset import QuantumStateDataset, load_circuit_records, state_vector_of
from .diffusion import DiscreteDiffusion, NativeAdam, TrainGraph, cosine_schedule, linear_schedule
from .distributed import all_reduce_histograms, sample_sharded, shard_range
from .evaluate import calculate_z_bias, evaluate, evaluate_records, format_raw_counts_for_inversion, write_metrics_csv
from .synthetic import born_histograms, counts_records, generate_synthetic_data, get_basis_combinations, synth_state
from .model import ConditionalD3PM, pack_bits, unpack_bits
from .notebook import BitstringDDM, SimpleMLP, UpgradedMLP
from .reconstruct import (DensityMatrix, Statevector, basis_strings, get_coefficient, get_metrics, get_pauli_matrix,
                          histogram_samples, linear_inversion, linear_inversion_raw, make_positive_semidefinite,
                          ReconReport, recon_report, state_fidelity)

__all__ = [
    "ConditionalD3PM", "DiscreteDiffusion", "NativeAdam", "TrainGraph", "cosine_schedule", "linear_schedule", "pack_bits", "unpack_bits",
    "DensityMatrix", "Statevector", "basis_strings", "get_coefficient", "get_metrics", "get_pauli_matrix",
    "histogram_samples", "linear_inversion", "linear_inversion_raw", "make_positive_semidefinite", "state_fidelity", "ReconReport", "recon_report",
    "calculate_z_bias", "evaluate", "evaluate_records", "format_raw_counts_for_inversion", "write_metrics_csv", "born_histograms", "counts_records", "generate_synthetic_data", "get_basis_combinations", "synth_state", "QuantumStateDataset", "load_circuit_records", "state_vector_of", "all_reduce_histograms", "sample_sharded", "shard_range", "build",
]
