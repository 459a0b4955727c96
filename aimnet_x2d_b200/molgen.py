"""Seeded synthetic molecules (numpy only): random valence-limited heavy-atom trees + ring closures + explicit hydrogens,
of the shapes named in BASELINE.json.  Everything derives from ``numpy.random.Generator(PCG64(seed))``.

This module has NO imports from the package (and therefore does not load ``libax2d.so``): ``bench.py --impl reference``
loads it by file path so that the reference arm generates the very same molecules without mapping the CUDA library.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

FEATURE_SIZES = {"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}   # main/runner.py:665-670
_VALENCE = {5: 4, 6: 3, 7: 2, 8: 1}          # atom_type index (Z-1): C, N, O, F


def _heavy_skeleton(rng, n_heavy: int, ring_mean: float, ring_dist, elem_p, pref: float = 1.0):
    elems = rng.choice([5, 6, 7], size=n_heavy, p=elem_p)
    if n_heavy > 2:
        elems[0] = 5
    cap = np.array([_VALENCE[int(e)] for e in elems])
    deg = np.zeros(n_heavy, dtype=np.int64)
    bonds = []
    adj = [[] for _ in range(n_heavy)]
    for i in range(1, n_heavy):
        free = np.nonzero(deg[:i] < cap[:i])[0]
        if free.size == 0:                    # everything saturated: promote an earlier atom to carbon
            j = int(rng.integers(0, i))
            elems[j], cap[j] = 5, 4
            free = np.array([j]) if deg[j] < 4 else np.nonzero(deg[:i] < 4)[0]
            if free.size == 0:
                break
        w = (deg[free] + 1.0) ** pref           # mild preference for already-branched atoms
        j = int(free[rng.choice(free.size, p=w / w.sum())])
        bonds.append((j, i)); adj[j].append(i); adj[i].append(j)
        deg[j] += 1; deg[i] += 1
    n_heavy = max(len(adj) if bonds else 1, 1)
    for _ in range(int(rng.poisson(ring_mean))):
        # BFS distances from a random atom with free valence
        free = np.nonzero(deg < cap)[0]
        if free.size < 2:
            break
        a = int(free[rng.integers(0, free.size)])
        dist = {a: 0}
        queue = [a]
        for u in queue:
            if dist[u] >= ring_dist[1]:
                continue
            for w in adj[u]:
                if w not in dist:
                    dist[w] = dist[u] + 1
                    queue.append(w)
        cands = [w for w, dd in dist.items() if ring_dist[0] <= dd <= ring_dist[1] and deg[w] < cap[w]]
        if not cands:
            continue
        b = int(cands[rng.integers(0, len(cands))])
        bonds.append((min(a, b), max(a, b))); adj[a].append(b); adj[b].append(a)
        deg[a] += 1; deg[b] += 1
    return elems, cap, deg, bonds


def make_molecule(rng, kind: str = "qm9", num_targets: int = 12, stereo: bool = False) -> Dict:
    if kind == "qm9":
        n_heavy = int(np.clip(np.rint(rng.normal(8.8, 0.6)), 1, 9))
        # parameters tuned so that atoms/mol and directed shell edges/atom (hops 1-3) land within 5 % of the
        # QM9 statistics of SURVEY.md section 8 (17.96; 2.07 / 3.58 / 4.21): measured 18.3; 2.15 / 3.58 / 4.11
        elems, cap, deg, bonds = _heavy_skeleton(rng, n_heavy, 2.4, (2, 4), [0.90, 0.05, 0.05], 1.0)
        max_atoms, p_lose1, p_lose2 = 29, 0.40, 0.10
    else:                                     # drug-like: 20-70 heavy atoms, 5/6-rings
        n_heavy = int(rng.integers(20, 71))
        elems, cap, deg, bonds = _heavy_skeleton(rng, n_heavy, 3.5, (4, 5), [0.74, 0.12, 0.14], 0.0)
        max_atoms, p_lose1, p_lose2 = 10 ** 9, 0.45, 0.20
    n_heavy = len(elems)
    atom_type = [int(e) for e in elems]
    h_count = np.zeros(n_heavy, dtype=np.int64)
    bonds = list(bonds)
    n = n_heavy
    for i in range(n_heavy):
        # unsaturation (double bonds / aromaticity in real molecules) removes one or two hydrogens
        r = rng.random()
        free = max(int(cap[i] - deg[i]) - (2 if r < p_lose2 else (1 if r < p_lose2 + p_lose1 else 0)), 0)
        for _ in range(free):
            if n >= max_atoms:
                break
            bonds.append((i, n)); atom_type.append(0); n += 1
            h_count[i] += 1
    degree = np.zeros(n, dtype=np.int64)
    for a, b in bonds:
        degree[a] += 1; degree[b] += 1
    hyb = np.where(np.array(atom_type) == 0, 0, np.clip(degree, 1, 4) - 1 + 1)   # 0 for H, 1..4 by degree
    mol = dict(num_atoms=n, bonds=np.array(bonds, dtype=np.int32).reshape(-1, 2),
               features=dict(atom_type=np.array(atom_type, dtype=np.int64),
                             hydrogen_count=np.concatenate([h_count, np.zeros(n - n_heavy, np.int64)]),
                             degree=np.clip(degree, 0, 6), hybridization=np.clip(hyb, 0, 6).astype(np.int64)),
               target=rng.normal(0.0, 1.0, size=num_targets).astype(np.float32),
               total_charge=float(rng.choice([-1, 0, 0, 0, 0, 1])) if kind != "qm9" else 0.0,
               atomic_numbers=np.array(atom_type, dtype=np.int64) + 1, chiral=[], cis=[], trans=[])
    if stereo:
        nbrs = [[] for _ in range(n)]
        for a, b in bonds:
            nbrs[a].append(b); nbrs[b].append(a)
        centres = [i for i in range(n_heavy) if len(nbrs[i]) == 4]
        for _ in range(min(int(rng.poisson(1.5)), len(centres))):
            c = centres.pop(int(rng.integers(0, len(centres))))
            mol["chiral"].append(np.array(nbrs[c], dtype=np.int64))       # rows = the 4 neighbours (quirk Q4)
        heavy_bonds = [(a, b) for a, b in bonds if a < n_heavy and b < n_heavy]
        for _ in range(int(rng.poisson(0.3))):
            if not heavy_bonds:
                break
            a, b = heavy_bonds[int(rng.integers(0, len(heavy_bonds)))]
            na = [w for w in nbrs[a] if w != b]
            nb = [w for w in nbrs[b] if w != a]
            if na and nb:
                p, q = na[0], nb[0]
                (mol["cis"] if rng.random() < 0.5 else mol["trans"]).append(np.array([p, q], dtype=np.int64))
    return mol


def graph_stats(mols: List[Dict]) -> Dict[str, float]:
    n = sum(m["num_atoms"] for m in mols)
    per_hop = np.sum([[h.shape[1] for h in m["hops"]] for m in mols], axis=0)
    return dict(atoms_per_mol=n / len(mols), edges_per_atom_by_hop=[float(c) / n for c in per_hop],
                edges_per_atom=float(per_hop.sum()) / n, max_atoms=max(m["num_atoms"] for m in mols))
