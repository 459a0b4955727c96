"""Seeded synthetic molecular graphs of the shapes named in BASELINE.json.

There is no network and no RDKit here, so benchmark / test inputs are generated: random valence-limited
heavy-atom trees + ring closures + explicit hydrogens, then the reference's shell-edge BFS
(``ax2d_host_shell_edges`` == datasets/features.py:97-150) and the reference collation layout.
Everything derives from ``numpy.random.Generator(PCG64(seed))``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import numpy as np
import torch

from . import _lib
from .collate import MolBatch, MolData
from .molgen import FEATURE_SIZES, graph_stats, make_molecule  # noqa: F401  (re-exported)

def shell_edges_batch(mols: List[Dict], num_hops: int):
    """Reference-order shell edges for a list of molecules; fills ``mol['hops']`` (list of (2,E_h) int64)."""
    lib = _lib.load()
    B = len(mols)
    atom_ptr = np.zeros(B + 1, dtype=np.int64)
    bond_ptr = np.zeros(B + 1, dtype=np.int64)
    atom_ptr[1:] = np.cumsum([m["num_atoms"] for m in mols])
    bond_ptr[1:] = np.cumsum([len(m["bonds"]) for m in mols])
    bonds = np.ascontiguousarray(np.concatenate([m["bonds"].reshape(-1, 2) for m in mols], 0).astype(np.int32)) \
        if B else np.zeros((0, 2), np.int32)
    counts = np.zeros(B * num_hops, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    total = lib.ax2d_host_shell_edges(B, p(atom_ptr), p(bond_ptr), p(bonds), num_hops, p(counts), None, 0)
    if total < 0:
        _lib.check(int(total), "ax2d_host_shell_edges")
    edges = np.zeros((max(total, 1), 2), dtype=np.int64)
    total2 = lib.ax2d_host_shell_edges(B, p(atom_ptr), p(bond_ptr), p(bonds), num_hops, p(counts), p(edges), total)
    assert total2 == total
    edges = edges[:total]
    counts = counts.reshape(B, num_hops)
    pos = 0
    for g, m in enumerate(mols):
        hops = []
        for h in range(num_hops):
            c = int(counts[g, h])
            hops.append((edges[pos:pos + c] - atom_ptr[g]).T.copy())
            pos += c
        m["hops"] = hops
    return edges, counts


def to_data(mol: Dict) -> MolData:
    n = mol["num_atoms"]
    return MolData(x=torch.zeros((n, 1)), multi_hop_edges=[torch.from_numpy(h) for h in mol["hops"]],
                   atom_features_map={k: torch.from_numpy(v) for k, v in mol["features"].items()},
                   target=torch.from_numpy(mol["target"]), total_charge=torch.tensor([mol["total_charge"]]),
                   chiral_tensors=[torch.from_numpy(c) for c in mol["chiral"]],
                   cis_bonds_tensors=[torch.from_numpy(c) for c in mol["cis"]],
                   trans_bonds_tensors=[torch.from_numpy(c) for c in mol["trans"]],
                   smiles="", atomic_numbers=torch.from_numpy(mol["atomic_numbers"]))


def make_molecules(seed: int, num_graphs: int, num_hops: int = 3, kind: str = "qm9", num_targets: int = 12,
                   stereo: bool = False) -> List[Dict]:
    rng = np.random.Generator(np.random.PCG64(seed))
    mols = [make_molecule(rng, kind, num_targets, stereo) for _ in range(num_graphs)]
    shell_edges_batch(mols, num_hops)
    return mols


def make_batch(seed: int, num_graphs: int, num_hops: int = 3, kind: str = "qm9", num_targets: int = 12,
               stereo: bool = False, tile_rows: int = 32) -> MolBatch:
    mols = make_molecules(seed, num_graphs, num_hops, kind, num_targets, stereo)
    return MolBatch.from_data_list([to_data(m) for m in mols], FEATURE_SIZES, tile_rows)
