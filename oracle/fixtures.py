"""Deterministic weights / tensors shared by the golden-vector generator and the tests -- TEST INFRASTRUCTURE.

Golden fixtures store inputs and expected outputs but not multi-megabyte weight tensors: weights are
regenerated from a seed by :func:`det_state`, identically in ``tests/golden/make_golden.py`` (which feeds
them to the UNMODIFIED reference modules) and in the tests (which feed them to the oracle and to the CUDA
modules).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch

FEATURE_SIZES = {"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}   # main/runner.py:665-670


def det_tensor(shape: Tuple[int, ...], rng: np.random.Generator, name: str = "") -> torch.Tensor:
    if len(shape) == 0:
        return torch.tensor(0.8 + 0.1 * float(rng.random()), dtype=torch.float32)
    if len(shape) == 1:
        return torch.from_numpy(rng.uniform(-0.1, 0.1, size=shape).astype(np.float32))
    if "embedding.weight" in name:
        return torch.from_numpy(rng.uniform(-0.7, 0.7, size=shape).astype(np.float32))
    a = 1.6 / np.sqrt(shape[1])
    return torch.from_numpy(rng.uniform(-a, a, size=shape).astype(np.float32))


def det_state(shapes: "OrderedDict[str, Tuple[int, ...]]", seed: int) -> "OrderedDict[str, torch.Tensor]":
    rng = np.random.Generator(np.random.PCG64(seed))
    return OrderedDict((k, det_tensor(tuple(s), rng, k)) for k, s in shapes.items())


def batch_to_torch(arrs: Dict[str, np.ndarray]) -> Dict[str, object]:
    """npz arrays (flat keys) -> the batch dict consumed by ``oracle.model_port.gnn_forward``."""
    t = lambda k, dt: torch.from_numpy(np.asarray(arrs[k])).to(dt)
    return dict(
        atom_features_map={k: t("feat_" + k, torch.long) for k in FEATURE_SIZES},
        multi_hop_edge_indices=t("edges", torch.long), batch_indices=t("batch_indices", torch.long),
        total_charges=t("total_charges", torch.float32),
        final_tetrahedral_chiral_tensor=t("tetra", torch.long).reshape(-1, 4),
        final_cis_tensor=t("cis", torch.long).reshape(-1, 2), final_trans_tensor=t("trans", torch.long).reshape(-1, 2),
        targets=t("targets", torch.float32))


def batch_to_arrays(batch) -> Dict[str, np.ndarray]:
    """A collated batch object (reference field names) -> flat numpy arrays for an npz."""
    out = {"edges": np.ascontiguousarray(batch.multi_hop_edge_indices.numpy()),
           "batch_indices": batch.batch_indices.numpy(), "total_charges": batch.total_charges.numpy(),
           "tetra": batch.final_tetrahedral_chiral_tensor.numpy(), "cis": batch.final_cis_tensor.numpy(),
           "trans": batch.final_trans_tensor.numpy(), "targets": batch.targets.numpy()}
    for k in FEATURE_SIZES:
        out["feat_" + k] = batch.atom_features_map[k].numpy()
    return out
