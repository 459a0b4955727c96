"""Integer side of the oracle (numpy / pure Python) -- TEST INFRASTRUCTURE.

Restates, for checking only:

* shell-edge BFS           -- ``src/datasets/features.py:82-150``
* batch collation          -- ``src/datasets/molecular.py:339-458`` (``MyBatch.from_data_list``)
* CSR / segment artefacts  -- what ``layers.py:154-163`` (gather + scatter_add on CPU) is
  equivalent to: stable sort of the edge list by ``target`` (SURVEY.md section 8c determinism note).

All results are exact integers; the CUDA/host product path must match them bit for bit.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- BFS
def shell_edges_bfs(num_atoms: int, bonds, max_hops: int):
    """Hop-by-hop ordered atom pairs at shortest-path distance exactly k.

    Follows ``features.py:97-150``: hop 1 enumerates (v, w) for v ascending and w over the
    ascending neighbour list of v; hop k+1 walks the previous frontier IN ORDER and, for each
    (u, v), appends (u, w) for every neighbour w of v with w != u that has not been reached
    from u yet.  If a hop comes out empty the remaining hops are empty too.
    Returns a list of ``max_hops`` int32 arrays of shape (2, E_h): row 0 = origin u, row 1 = w.
    """
    nbrs = [[] for _ in range(num_atoms)]
    for a, b in bonds:
        a = int(a); b = int(b)
        if a == b:
            continue
        nbrs[a].append(b)
        nbrs[b].append(a)
    nbrs = [sorted(set(l)) for l in nbrs]           # np.where(adj[v] > 0) is ascending, no self loop
    seen = np.zeros((num_atoms, num_atoms), dtype=bool)
    frontier = []
    for v in range(num_atoms):
        for w in nbrs[v]:
            if not seen[v, w]:
                seen[v, w] = True
                frontier.append((v, w))
    hops = [np.array(frontier, dtype=np.int32).reshape(-1, 2).T.copy()]
    for _ in range(1, max_hops):
        nxt = []
        for (u, v) in frontier:
            for w in nbrs[v]:
                if w != u and not seen[u, w]:
                    seen[u, w] = True
                    nxt.append((u, w))
        if not nxt:
            hops.append(np.empty((2, 0), dtype=np.int32))
            break
        hops.append(np.array(nxt, dtype=np.int32).reshape(-1, 2).T.copy())
        frontier = nxt
    while len(hops) < max_hops:
        hops.append(np.empty((2, 0), dtype=np.int32))
    return hops


# --------------------------------------------------------------------------- collation
def collate(mols):
    """Concatenate molecules the way ``MyBatch.from_data_list`` does (``molecular.py:339-458``).

    ``mols`` is a list of dicts with keys
      ``num_atoms``, ``hops`` (list of (2,E_h) int arrays), ``features`` (dict name -> int array [n]),
      ``target`` (float array [T]), ``total_charge`` (float), ``chiral`` (list of int arrays, any length),
      ``cis`` / ``trans`` (lists of length-2 int arrays), optional ``atomic_numbers``.
    Returns a dict of numpy arrays with the reference's Batch field names.
    """
    B = len(mols)
    n_atoms = np.array([m["num_atoms"] for m in mols], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(n_atoms[:-1])]).astype(np.int64) if B else np.zeros(0, np.int64)
    num_hops = len(mols[0]["hops"]) if B else 0

    chir, cis, trans = [], [], []
    for m, off in zip(mols, offsets):
        chir.extend(np.asarray(c, dtype=np.int64) + off for c in m.get("chiral", []) if len(c) == 4)   # :365
        cis.extend(np.asarray(c, dtype=np.int64) + off for c in m.get("cis", []))
        trans.extend(np.asarray(c, dtype=np.int64) + off for c in m.get("trans", []))
    tetra = np.stack(chir, 0) if chir else np.empty((0, 4), np.int64)
    cis_t = np.stack(cis, 0) if cis else np.empty((0, 2), np.int64)
    trans_t = np.stack(trans, 0) if trans else np.empty((0, 2), np.int64)
    # reversed direction appended (:387-397)
    final_cis = np.concatenate([cis_t, cis_t[:, ::-1]], 0) if cis_t.size else np.empty((0, 2), np.int64)
    final_trans = np.concatenate([trans_t, trans_t[:, ::-1]], 0) if trans_t.size else np.empty((0, 2), np.int64)

    keys = list(mols[0]["features"].keys()) if B else []
    feats = {k: np.concatenate([np.asarray(m["features"][k], dtype=np.int64) for m in mols]) for k in keys}
    batch_indices = np.repeat(np.arange(B, dtype=np.int64), n_atoms)
    targets = np.stack([np.asarray(m["target"], dtype=np.float32) for m in mols], 0) if B else np.zeros((0, 0), np.float32)
    total_charges = np.array([m["total_charge"] for m in mols], dtype=np.float32)

    pieces = []
    for m, off in zip(mols, offsets):                      # :427-433  atom offset ONLY (quirk Q1)
        for h in range(num_hops):
            e = np.asarray(m["hops"][h], dtype=np.int64)
            if e.size:
                pieces.append(e + off)
    edges = np.concatenate(pieces, 1).T if pieces else np.empty((0, 2), np.int64)   # [E,2] (target, src)
    out = dict(multi_hop_edge_indices=np.ascontiguousarray(edges), batch_indices=batch_indices,
               atom_features_map=feats, targets=targets, total_charges=total_charges,
               final_tetrahedral_chiral_tensor=tetra, final_cis_tensor=final_cis,
               final_trans_tensor=final_trans, num_atoms=n_atoms)
    if B and "atomic_numbers" in mols[0]:
        out["atomic_numbers"] = np.concatenate([np.asarray(m["atomic_numbers"], np.int64) for m in mols])
    return out


# --------------------------------------------------------------------------- CSR artefacts
def csr_by_key(key: np.ndarray, val: np.ndarray, num_rows: int):
    """rowptr/col/perm of a STABLE sort of the edge list by ``key`` (int32 outputs)."""
    key = np.asarray(key, dtype=np.int64)
    perm = np.argsort(key, kind="stable")
    rowptr = np.zeros(num_rows + 1, dtype=np.int64)
    np.add.at(rowptr, key + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), np.asarray(val, dtype=np.int64)[perm].astype(np.int32), perm.astype(np.int32)


def csr_artefacts(edges: np.ndarray, num_atoms: int, num_hops: int):
    """Forward CSR (rows = target in [0, H*N), cols = src mod N) and transposed CSR (rows = src mod N,
    cols = target) -- the accumulation orders of ``layers.py:155-163`` forward and of its CPU backward."""
    tgt = edges[:, 0].astype(np.int64)
    src = edges[:, 1].astype(np.int64) % max(num_atoms, 1)          # layers.py:154
    R = num_hops * num_atoms
    rowptr, col, perm = csr_by_key(tgt, src, R)
    rowptr_t, col_t, perm_t = csr_by_key(src, tgt, num_atoms)
    return dict(rowptr=rowptr, col=col, perm=perm, rowptr_t=rowptr_t, col_t=col_t, perm_t=perm_t)


def segment_ptr(batch_indices: np.ndarray, num_graphs: int) -> np.ndarray:
    """seg_ptr[g] = first atom of molecule g (batch_indices is sorted, ``molecular.py:406-410``)."""
    counts = np.bincount(np.asarray(batch_indices, dtype=np.int64), minlength=num_graphs)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)


def aggregate_rows(x: np.ndarray, rowptr: np.ndarray, col: np.ndarray) -> np.ndarray:
    """Sequential fp32 accumulation of each output row in CSR order starting from 0 -- bit-identical to
    the reference CPU ``scatter_add`` (SURVEY.md section 8c determinism note)."""
    R = len(rowptr) - 1
    out = np.zeros((R, x.shape[1]), dtype=x.dtype)
    for r in range(R):
        acc = np.zeros(x.shape[1], dtype=x.dtype)
        for k in range(rowptr[r], rowptr[r + 1]):
            acc = acc + x[col[k]]
        out[r] = acc
    return out
