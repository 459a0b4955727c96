"""Batch collation that emits CSR / segment offsets.

Mirrors ``MyBatch.from_data_list`` (reference ``src/datasets/molecular.py:339-458``): same input
objects (anything with the reference ``Data`` attributes), same output field names, dtypes and layouts,
plus a :class:`GraphIndex` with the int32 artefacts the CUDA kernels consume.  The Python loops of the
reference (``molecular.py:352-438``) are replaced by vectorised numpy and the C functions
``ax2d_host_csr_build`` / ``ax2d_host_tile_plan`` (include/ax2d.h).  All integer results are exact.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

DEFAULT_TILE_ROWS = 32          # rows of x staged per CTA by the aggregation kernel (QM9: <= 29 atoms/molecule)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _np(t, dtype=None) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    a = np.asarray(t)
    return a if dtype is None else a.astype(dtype, copy=False)


def shell_csr(atom_counts, bonds, num_hops: int, device=None):
    """Shell-edge BFS (features.py:82-150) of a batch of molecules straight to the CSR pair of the aggregation kernels.
    ``atom_counts`` [B]; ``bonds``: per molecule an [nb, 2] array of molecule-local atom indices (undirected, each bond
    once).  Returns int32 tensors (rowptr [N + 1 + slack], col [E + slack], col_t [E + slack]): rows of ``col`` = sources
    of a target in the reference's stable edge order (hop-major, BFS discovery order), rows of ``col_t`` = targets of a
    source (hop-major, ascending); both share ``rowptr``.  ``device``: a CUDA device -> built on the GPU and returned there
    (molecules of at most 256 atoms); None -> host."""
    lib = _lib.load()
    counts = _np(atom_counts, np.int64).reshape(-1)
    B = int(counts.shape[0])
    if len(bonds) != B:
        raise ValueError("one bond array per molecule")
    atom_ptr = np.zeros(B + 1, dtype=np.int64)
    atom_ptr[1:] = np.cumsum(counts)
    bond_ptr = np.zeros(B + 1, dtype=np.int64)
    bond_ptr[1:] = np.cumsum([int(np.asarray(b).reshape(-1, 2).shape[0]) for b in bonds])
    flat = (np.ascontiguousarray(np.concatenate([np.asarray(b).reshape(-1, 2) for b in bonds], 0).astype(np.int32))
            if B and bond_ptr[-1] else np.zeros((0, 2), np.int32))
    N = int(atom_ptr[-1])
    slack = GraphIndex.INDEX_SLACK
    if device is None:
        rowptr = np.zeros(N + 1 + slack, dtype=np.int32)
        E = int(lib.ax2d_host_shell_csr(B, _ptr(atom_ptr), _ptr(bond_ptr), _ptr(flat), int(num_hops), _ptr(rowptr), None, None, 0))
        if E < 0:
            _lib.check(E, "ax2d_host_shell_csr")
        col = np.zeros(max(E, 1) + slack, dtype=np.int32)
        col_t = np.zeros(max(E, 1) + slack, dtype=np.int32)
        E2 = int(lib.ax2d_host_shell_csr(B, _ptr(atom_ptr), _ptr(bond_ptr), _ptr(flat), int(num_hops), _ptr(rowptr), _ptr(col),
                                         _ptr(col_t), E))
        if E2 != E:
            _lib.check(E2 if E2 < 0 else -1, "ax2d_host_shell_csr")
        rowptr[N + 1:] = E
        return torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(col_t)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("shell_csr: device must be a CUDA device (or None for the host path)")
    max_atoms = int(counts.max()) if B else 0
    if max_atoms > 256:
        raise RuntimeError(f"shell_csr: a molecule of {max_atoms} atoms exceeds the device kernel's 256; use device=None")
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        d_atom = torch.from_numpy(atom_ptr.astype(np.int32)).to(dev)
        d_bond = torch.from_numpy(bond_ptr.astype(np.int32)).to(dev)
        d_flat = torch.from_numpy(flat).to(dev) if flat.size else torch.zeros((1, 2), dtype=torch.int32, device=dev)
        rowptr = torch.zeros(N + 1 + slack, dtype=torch.int32, device=dev)
        work = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.ax2d_shell_csr_count(p(d_atom), p(d_bond), p(d_flat), B, N, max_atoms, int(num_hops), p(rowptr), p(work),
                                            p(err), stream), "ax2d_shell_csr_count")
        E, bad = int(rowptr[N].item()), int(err.item())
        if bad:
            raise RuntimeError(f"ax2d_shell_csr_count: inconsistent input (code {bad})")
        col = torch.zeros(max(E, 1) + slack, dtype=torch.int32, device=dev)
        col_t = torch.zeros(max(E, 1) + slack, dtype=torch.int32, device=dev)
        _lib.check(lib.ax2d_shell_csr_fill(p(d_atom), p(d_bond), p(d_flat), B, max_atoms, int(num_hops), p(rowptr), p(col),
                                           p(col_t), p(err), stream), "ax2d_shell_csr_fill")
        if int(err.item()):
            raise RuntimeError(f"ax2d_shell_csr_fill: inconsistent input (code {int(err.item())})")
        rowptr[N + 1:] = E
    return rowptr, col, col_t


class GraphIndex:
    """Integer artefacts of one batch (all int32, built once at collation):

    ``rowptr/col``      forward CSR: rows = target, stable in edge order  (layers.py:154-163)
    ``rowptr_t/col_t``  transposed CSR: rows = src mod N                 (its CPU backward order)
    ``seg_ptr``         molecule row offsets from the sorted batch_indices (molecular.py:406-410)
    ``tile_ptr``        whole-molecule row tiles for the shared-memory aggregation kernel
    ``collapsed``       True when every target < N (quirk Q1: the shipped collation never offsets hops)
    ``tile_local``      True when no edge leaves its tile (no edge leaves its molecule, molecular.py:429-433)
    """

    _TENSORS = ("rowptr", "col", "rowptr_t", "col_t", "seg_ptr", "tile_ptr", "tile_info", "tile_info_t")
    INDEX_SLACK = 8     # readable entries past the logical end of rowptr / col (16-byte copy windows of the tiled kernel)

    def __init__(self):
        self.num_atoms = 0
        self.num_graphs = 0
        self.num_edges = 0
        self.num_hops = 0
        self.num_rows = 0              # rows of the forward CSR: N if collapsed else H*N
        self.collapsed = True
        self.tile_local = False
        self.unique_edges = False      # no (target, source) pair occurs twice (the dense tile product cannot count)
        self.n_tiles = 0
        self.max_tile_rows = 0
        self.max_tile_edges = 0
        self.max_seg = 0
        self.rowptr = self.col = self.rowptr_t = self.col_t = self.seg_ptr = self.tile_ptr = None
        self.tile_info = self.tile_info_t = None     # [n_tiles, 4] int32: row0, row1, rowptr[row0], rowptr[row1]
        self.embed: Dict[str, tuple] = {}      # name -> (order int32 [N], ptr int32 [vocab+1], vocab)
        self.tetra = None                      # (idx int32 [M,4], slot_ptr int32 [N+1], slot_idx int32 [4M], M)
        self.cistrans = None                   # (src int32, tgt int32, sign f32, n_upd)

    # ------------------------------------------------------------------ construction (host, exact integers)
    @staticmethod
    def from_bonds(atom_counts, bonds, num_hops: int, atom_features: Optional[Dict[str, torch.Tensor]] = None,
                   feature_sizes: Optional[Dict[str, int]] = None, tetra=None, cis=None, trans=None,
                   tile_rows: int = DEFAULT_TILE_ROWS, device=None) -> "GraphIndex":
        """Featurisation + collation of the graph structure in one step (SURVEY f-1): the shell-edge BFS
        (features.py:82-150) emits the CSR pair directly from the molecules' bond lists -- on the host
        (``ax2d_host_shell_csr``) or, with ``device``, on the GPU (``ax2d_shell_csr_count`` / ``_fill``) -- so no edge
        list is built and nothing is sorted.  ``atom_counts`` [B]; ``bonds``: one [nb, 2] array of molecule-local atom
        indices per molecule.  Bit-identical to ``build`` over the reference's collated edge list."""
        counts = _np(atom_counts, np.int64).reshape(-1)
        B = int(counts.shape[0])
        rowptr, col, col_t = shell_csr(counts, bonds, num_hops, device=device)
        bi = np.repeat(np.arange(B, dtype=np.int64), counts)
        return GraphIndex.build(None, bi, B, num_hops, atom_features, feature_sizes, tetra, cis, trans, tile_rows,
                                csr=(rowptr.cpu().numpy(), col.cpu().numpy(), col_t.cpu().numpy()))

    @staticmethod
    def build(edges, batch_indices, num_graphs: int, num_hops: int,
              atom_features: Optional[Dict[str, torch.Tensor]] = None,
              feature_sizes: Optional[Dict[str, int]] = None,
              tetra=None, cis=None, trans=None, tile_rows: int = DEFAULT_TILE_ROWS, csr=None) -> "GraphIndex":
        """``csr``: (rowptr [N+1+slack], col [E+slack], col_t [E+slack]) int32 arrays from ``shell_csr`` instead of
        ``edges`` (shell relations are symmetric: both CSRs share rowptr; the edge list is duplicate-free and never
        leaves a molecule by construction)."""
        lib = _lib.load()
        gi = GraphIndex()
        bi = _np(batch_indices, np.int64)
        N = int(bi.shape[0])
        B = int(num_graphs)
        gi.num_atoms, gi.num_graphs, gi.num_hops = N, B, int(num_hops)
        if N and (np.any(np.diff(bi) < 0) or bi[0] < 0 or bi[-1] >= B):
            raise ValueError("batch_indices must be sorted and within [0, num_graphs)")
        seg = np.zeros(B + 1, dtype=np.int32)
        if N:
            seg[1:] = np.cumsum(np.bincount(bi, minlength=B)).astype(np.int32)
        gi.max_seg = int(np.max(np.diff(seg))) if B else 0

        if csr is not None and edges is not None:
            raise ValueError("GraphIndex.build: pass either edges or csr")
        if edges is None:
            e_np = np.empty((0, 2), dtype=np.int64)
        elif isinstance(edges, torch.Tensor):
            e_np = edges.detach().cpu().numpy()          # keeps the (possibly transposed) strides
        else:
            e_np = np.asarray(edges)
        if e_np.dtype != np.int64:
            e_np = e_np.astype(np.int64)
        E = int(e_np.shape[0]) if e_np.ndim == 2 else 0
        if csr is not None:
            E = int(csr[0][N])
        gi.num_edges = E
        if E and csr is None:
            tmax = int(e_np[:, 0].max())
            if tmax >= max(num_hops, 1) * N or int(e_np.min()) < 0:
                raise ValueError("edge index out of range for message_passing (target must be < num_hops * N)")
            gi.collapsed = tmax < N
        R = N if gi.collapsed else num_hops * N
        gi.num_rows = R
        gi.unique_edges = True                   # edge lists: checked on the CSR below (shell CSRs are sets by construction)
        slack = GraphIndex.INDEX_SLACK
        rowptr = np.zeros(R + 1 + slack, dtype=np.int32)
        col = np.zeros(max(E, 1) + slack, dtype=np.int32)
        rowptr_t = np.zeros(N + 1 + slack, dtype=np.int32)
        col_t = np.zeros(max(E, 1) + slack, dtype=np.int32)
        if csr is not None:
            rowptr[: N + 1] = csr[0][: N + 1]
            rowptr_t[: N + 1] = csr[0][: N + 1]
            col[:E] = csr[1][:E]
            col_t[:E] = csr[2][:E]
        elif E:
            se, sc = (s // 8 for s in e_np.strides)
            _lib.check(lib.ax2d_host_csr_build(_ptr(e_np), E, se, sc, N, R, 0, _ptr(rowptr), _ptr(col), None),
                       "ax2d_host_csr_build")
            _lib.check(lib.ax2d_host_csr_build(_ptr(e_np), E, se, sc, N, N, 1, _ptr(rowptr_t), _ptr(col_t), None),
                       "ax2d_host_csr_build(transposed)")
            uniq = C.c_int32(1)                   # (np.unique over E 64-bit keys cost 1.1 s of a 1.3 s collation here)
            _lib.check(lib.ax2d_host_csr_rows_unique(_ptr(rowptr), _ptr(col), R, N, C.byref(uniq)), "ax2d_host_csr_rows_unique")
            gi.unique_edges = bool(uniq.value)
        tile_ptr = np.zeros(B + 1, dtype=np.int32)
        n_tiles, max_rows, local = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        cap = max(int(tile_rows), gi.max_seg)
        use_csr = gi.collapsed and E > 0
        _lib.check(lib.ax2d_host_tile_plan(_ptr(seg), B, cap, _ptr(rowptr) if use_csr else None,
                                           _ptr(col) if use_csr else None, _ptr(tile_ptr),
                                           C.byref(n_tiles), C.byref(max_rows), C.byref(local)), "ax2d_host_tile_plan")
        gi.n_tiles, gi.max_tile_rows = int(n_tiles.value), int(max_rows.value)
        gi.tile_local = bool(local.value) and gi.collapsed
        if gi.tile_local and E:
            # the transposed CSR must be tile-local too (true whenever the forward one is: same edge set)
            t2, m2, l2 = C.c_int64(0), C.c_int64(0), C.c_int32(0)
            tp2 = np.zeros(B + 1, dtype=np.int32)
            _lib.check(lib.ax2d_host_tile_plan(_ptr(seg), B, cap, _ptr(rowptr_t), _ptr(col_t), _ptr(tp2),
                                               C.byref(t2), C.byref(m2), C.byref(l2)), "ax2d_host_tile_plan")
            gi.tile_local = bool(l2.value)
        rowptr[R + 1:] = E                       # slack entries: harmless values
        rowptr_t[N + 1:] = E
        gi.rowptr, gi.col = torch.from_numpy(rowptr), torch.from_numpy(col)
        gi.rowptr_t, gi.col_t = torch.from_numpy(rowptr_t), torch.from_numpy(col_t)
        gi.seg_ptr = torch.from_numpy(seg)
        gi.tile_ptr = torch.from_numpy(tile_ptr[: gi.n_tiles + 1].copy())
        gi.refresh_tile_info()

        if atom_features is not None:
            sizes = feature_sizes or {}
            for name, idx in atom_features.items():
                gi.add_embedding_index(name, idx, sizes.get(name))
        gi.set_stereo(tetra, cis, trans)
        return gi

    def refresh_tile_info(self) -> None:
        """Per-tile descriptors of the persistent aggregation kernel, for the forward and the transposed CSR."""
        tp = self.tile_ptr.numpy().astype(np.int64)
        nt = self.n_tiles
        if nt == 0 or not self.collapsed:
            self.tile_info = self.tile_info_t = None
            self.max_tile_edges = 0
            return
        mx = 0
        for name, rp in (("tile_info", self.rowptr), ("tile_info_t", self.rowptr_t)):
            r = rp.numpy()
            info = np.stack([tp[:nt], tp[1:nt + 1], r[tp[:nt]], r[tp[1:nt + 1]]], axis=1).astype(np.int32)
            mx = max(mx, int((info[:, 3] - info[:, 2]).max()))
            # largest tiles first: the kernel deals this list to its persistent CTAs in serpentine order, which balances
            # the edges per CTA (tiles are independent, any order gives the same result)
            info = info[np.argsort(-(info[:, 3] - info[:, 2]).astype(np.int64), kind="stable")]
            setattr(self, name, torch.from_numpy(np.ascontiguousarray(info)))
        self.max_tile_edges = mx

    def add_embedding_index(self, name: str, idx, vocab: Optional[int] = None) -> None:
        """Atoms sorted (stably) by table row: the CPU ``index_add`` order of the embedding backward."""
        a = _np(idx, np.int64)
        v = int(vocab) if vocab is not None else (int(a.max()) + 1 if a.size else 1)
        if a.size and (a.min() < 0 or a.max() >= v):
            raise ValueError(f"feature index '{name}' out of range [0, {v})")
        order = np.argsort(a, kind="stable").astype(np.int32)
        ptr = np.zeros(v + 1, dtype=np.int32)
        if a.size:
            ptr[1:] = np.cumsum(np.bincount(a, minlength=v)).astype(np.int32)
        self.embed[name] = (torch.from_numpy(order), torch.from_numpy(ptr), v)

    def set_stereo(self, tetra, cis, trans) -> None:
        N = self.num_atoms
        if tetra is not None:
            t = _np(tetra, np.int64).reshape(-1, 4)
            M = int(t.shape[0])
            if M:
                if t.min() < 0 or t.max() >= N:
                    raise ValueError("tetrahedral index out of range")
                flat = t.reshape(-1)
                slot_idx = np.argsort(flat, kind="stable").astype(np.int32)       # index_add_ order, gnn.py:449-453
                slot_ptr = np.zeros(N + 1, dtype=np.int32)
                slot_ptr[1:] = np.cumsum(np.bincount(flat, minlength=N)).astype(np.int32)
                self.tetra = (torch.from_numpy(t.astype(np.int32)), torch.from_numpy(slot_ptr),
                              torch.from_numpy(slot_idx), M)
        src, tgt, sign = [], [], []
        for arr, s in ((cis, -1.0), (trans, 1.0)):                                # gnn.py:478-497 (quirk Q3)
            if arr is None:
                continue
            a = _np(arr, np.int64)
            if a.size == 0:
                continue
            if a.ndim != 2 or a.shape[0] < 2:
                raise ValueError("cis/trans index tensors must be [2K,2] with K >= 1")
            src.extend(int(v) for v in a[0])
            tgt.extend(int(v) for v in a[1])
            sign.extend([s] * a.shape[1])
        if src:
            if min(src + tgt) < 0 or max(src + tgt) >= N:
                raise ValueError("cis/trans index out of range")
            self.cistrans = (torch.tensor(src, dtype=torch.int32), torch.tensor(tgt, dtype=torch.int32),
                             torch.tensor(sign, dtype=torch.float32), len(src))

    # ------------------------------------------------------------------ movement
    def to(self, device, non_blocking: bool = False) -> "GraphIndex":
        out = GraphIndex()
        out.__dict__.update(self.__dict__)
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        for k in self._TENSORS:
            setattr(out, k, None if getattr(self, k) is None else mv(getattr(self, k)))
        out.embed = {k: (mv(o), mv(p), v) for k, (o, p, v) in self.embed.items()}
        if self.tetra is not None:
            i, sp, si, M = self.tetra
            out.tetra = (mv(i), mv(sp), mv(si), M)
        if self.cistrans is not None:
            s, t, g, n = self.cistrans
            out.cistrans = (mv(s), mv(t), mv(g), n)
        return out

    def pin_memory(self) -> "GraphIndex":
        out = GraphIndex()
        out.__dict__.update(self.__dict__)
        for k in self._TENSORS:
            setattr(out, k, None if getattr(self, k) is None else getattr(self, k).pin_memory())
        out.embed = {k: (o.pin_memory(), p.pin_memory(), v) for k, (o, p, v) in self.embed.items()}
        if self.tetra is not None:
            i, sp, si, M = self.tetra
            out.tetra = (i.pin_memory(), sp.pin_memory(), si.pin_memory(), M)
        if self.cistrans is not None:
            s, t, g, n = self.cistrans
            out.cistrans = (s.pin_memory(), t.pin_memory(), g.pin_memory(), n)
        return out

    def nbytes(self) -> int:
        n = sum(getattr(self, k).numel() * 4 for k in self._TENSORS if getattr(self, k) is not None)
        n += sum((o.numel() + p.numel()) * 4 for o, p, _ in self.embed.values())
        if self.tetra is not None:
            n += sum(t.numel() * 4 for t in self.tetra[:3])
        if self.cistrans is not None:
            n += sum(t.numel() * 4 for t in self.cistrans[:3])
        return n


class MolData:
    """Attribute bag with the reference's per-molecule ``Data`` fields (molecular.py:349-433)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class MolBatch:
    """Output of collation: the reference ``Batch`` fields (molecular.py:441-456) + ``graph_index``."""

    FIELDS = ("x", "multi_hop_edge_indices", "batch_indices", "targets", "total_charges",
              "final_tetrahedral_chiral_tensor", "final_cis_tensor", "final_trans_tensor", "atomic_numbers")

    def __init__(self):
        self.x = None
        self.multi_hop_edge_indices = None
        self.batch_indices = None
        self.batch = None
        self.atom_features_map: Dict[str, torch.Tensor] = {}
        self.targets = None
        self.total_charges = None
        self.final_tetrahedral_chiral_tensor = None
        self.final_cis_tensor = None
        self.final_trans_tensor = None
        self.smiles_list: List[str] = []
        self.atomic_numbers = None
        self.graph_index: Optional[GraphIndex] = None

    @staticmethod
    def from_data_list(data_list: Sequence, feature_sizes: Optional[Dict[str, int]] = None,
                       tile_rows: int = DEFAULT_TILE_ROWS) -> "MolBatch":
        batch = MolBatch()
        B = len(data_list)
        if B == 0:                                                         # molecular.py:346-347
            return batch
        num_hops = len(data_list[0].multi_hop_edges)
        n_atoms = np.array([int(d.x.shape[0]) for d in data_list], dtype=np.int64)
        offsets = np.zeros(B, dtype=np.int64)
        np.cumsum(n_atoms[:-1], out=offsets[1:])

        def shifted(attr, width, keep=None):
            rows = []
            for d, off in zip(data_list, offsets):
                for t in getattr(d, attr, []) or []:
                    a = _np(t, np.int64)
                    if keep is None or a.shape[0] == keep:                 # 4-neighbour filter, molecular.py:365
                        rows.append(a + off)
            return np.stack(rows, 0) if rows else np.empty((0, width), np.int64)

        tetra = shifted("chiral_tensors", 4, keep=4)
        cis = shifted("cis_bonds_tensors", 2)
        trans = shifted("trans_bonds_tensors", 2)
        final_cis = np.concatenate([cis, cis[:, ::-1]], 0) if cis.size else np.empty((0, 2), np.int64)
        final_trans = np.concatenate([trans, trans[:, ::-1]], 0) if trans.size else np.empty((0, 2), np.int64)

        # Many small per-molecule tensors: one torch.cat per field (a C++ loop) instead of a numpy round trip per tensor
        # (18 k conversions = 2/3 of the collation time of a 2048-molecule batch).
        def cat64(items, dim=0):
            ts = [t if type(t) is torch.Tensor else torch.as_tensor(t) for t in items]
            out = torch.cat(ts, dim)
            return out if out.dtype == torch.int64 else out.to(torch.int64)

        keys = list(data_list[0].atom_features_map.keys())
        feats = {k: cat64([d.atom_features_map[k] for d in data_list]) for k in keys}
        batch_indices = np.repeat(np.arange(B, dtype=np.int64), n_atoms)

        tdim = data_list[0].target.shape[0]
        if all(d.target.shape[0] == tdim for d in data_list):              # molecular.py:413-418
            targets = torch.stack([torch.as_tensor(d.target) for d in data_list], 0).to(torch.float32)
        else:
            targets = [torch.as_tensor(d.target) for d in data_list]
        total_charges = torch.cat([torch.as_tensor(d.total_charge).reshape(-1) for d in data_list]).to(torch.float32)

        # [2, E_total] block written molecule by molecule, hop by hop, then handed out transposed -- the
        # same memory layout (and strides) as ``torch.cat(..., dim=1).t()`` at molecular.py:435-436
        pieces, piece_off = [], []
        for d, off in zip(data_list, offsets):
            for h in range(num_hops):
                e = d.multi_hop_edges[h]
                if type(e) is not torch.Tensor:
                    e = torch.as_tensor(e)
                if e.numel():
                    pieces.append(e)
                    piece_off.append(off)
        if pieces:
            block = cat64(pieces, 1)                                       # fresh [2, E_total] tensor
            counts = np.fromiter((int(e.shape[1]) for e in pieces), dtype=np.int64, count=len(pieces))
            block += torch.from_numpy(np.repeat(np.asarray(piece_off, dtype=np.int64), counts))
            edges = block.t()
        else:
            edges = torch.empty((0, 2), dtype=torch.long)

        batch.x = torch.cat([torch.as_tensor(d.x) for d in data_list], 0)
        batch.multi_hop_edge_indices = edges
        batch.batch_indices = torch.from_numpy(batch_indices)
        batch.batch = batch.batch_indices
        batch.atom_features_map = feats
        batch.targets = targets
        batch.total_charges = total_charges
        batch.final_tetrahedral_chiral_tensor = torch.from_numpy(tetra)
        batch.final_cis_tensor = torch.from_numpy(np.ascontiguousarray(final_cis))
        batch.final_trans_tensor = torch.from_numpy(np.ascontiguousarray(final_trans))
        batch.smiles_list = [getattr(d, "smiles", "") for d in data_list]
        if hasattr(data_list[0], "atomic_numbers"):
            batch.atomic_numbers = torch.cat([torch.as_tensor(d.atomic_numbers) for d in data_list], 0)
        batch.graph_index = GraphIndex.build(edges, batch_indices, B, num_hops, feats, feature_sizes,
                                             tetra, final_cis, final_trans, tile_rows)
        return batch

    # ------------------------------------------------------------------ movement
    def _apply(self, fn_t, fn_g) -> "MolBatch":
        out = MolBatch()
        out.__dict__.update(self.__dict__)
        for k in self.FIELDS:
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                setattr(out, k, fn_t(v))
        out.batch = out.batch_indices
        out.atom_features_map = {k: fn_t(v) for k, v in self.atom_features_map.items()}
        if self.graph_index is not None:
            out.graph_index = fn_g(self.graph_index)
        return out

    def to(self, device, non_blocking: bool = False) -> "MolBatch":
        return self._apply(lambda t: t.to(device, non_blocking=non_blocking),
                           lambda g: g.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "MolBatch":
        return self._apply(lambda t: t.contiguous().pin_memory() if t.numel() else t, lambda g: g.pin_memory())

    def nbytes(self) -> int:
        n = 0
        for k in self.FIELDS:
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                n += v.numel() * v.element_size()
        n += sum(v.numel() * v.element_size() for v in self.atom_features_map.values())
        if self.graph_index is not None:
            n += self.graph_index.nbytes()
        return n


def pad_batch(batch: MolBatch, num_atoms: int, num_edges: int, num_dummy: int = 64, num_tiles: Optional[int] = None,
              tile_rows: int = DEFAULT_TILE_ROWS, feature_sizes: Optional[Dict[str, int]] = None,
              max_tile_edges: int = 0, num_tetra: Optional[int] = None, pad_cistrans: bool = False) -> MolBatch:
    """Static-shape version of a collated batch (what a CUDA-graph-captured training step needs: every launch
    configuration and scalar kernel argument must be the same from batch to batch).

    ``num_dummy`` edge-less dummy molecules are appended; together they hold the ``num_atoms - N`` padding atoms
    (at most 29 each, like a QM9 molecule, so that they pack into the same whole-molecule row tiles).  Their feature
    indices are 0, their targets are never read: the loss is taken over the first ``B`` (real) rows only, so no
    gradient flows through them.  ``col`` arrays are padded to ``num_edges`` entries and the tile list to
    ``num_tiles`` (trailing tiles are empty).  Real molecules keep their rows, CSR rows and results bit-for-bit.

    Stereo batches: ``num_tetra`` pads the list of tetrahedral centres with dummy centres whose four "neighbours" are
    dummy atoms (real atoms are listed / zeroed exactly as before, gnn.py:456-460; a batch WITHOUT real centres stays
    un-padded because the reference then skips the term altogether, gnn.py:402-403).  ``pad_cistrans`` fills the cis /
    trans update list (at most four rows, quirk Q3) up to four with zero-weight updates on a dummy atom."""
    gi0 = batch.graph_index
    N, B = gi0.num_atoms, gi0.num_graphs
    pad = num_atoms - N
    if pad < 0 or pad > 29 * num_dummy:
        raise ValueError(f"cannot pad {N} atoms to {num_atoms} with {num_dummy} dummy molecules of <= 29 atoms")
    if gi0.num_edges > num_edges:
        raise ValueError(f"batch has {gi0.num_edges} edges, capacity {num_edges}")
    sizes = np.full(num_dummy, pad // num_dummy, dtype=np.int64)
    sizes[: pad % num_dummy] += 1
    out = MolBatch()
    out.__dict__.update(batch.__dict__)
    zl = lambda n: torch.zeros(n, dtype=torch.long)
    out.atom_features_map = {k: torch.cat([v, zl(pad)]) for k, v in batch.atom_features_map.items()}
    out.batch_indices = torch.cat([batch.batch_indices, torch.from_numpy(np.repeat(np.arange(B, B + num_dummy), sizes))])
    out.batch = out.batch_indices
    out.x = torch.cat([batch.x, torch.zeros((pad,) + tuple(batch.x.shape[1:]), dtype=batch.x.dtype)], 0)
    if isinstance(batch.targets, torch.Tensor):
        out.targets = torch.cat([batch.targets, torch.zeros((num_dummy,) + tuple(batch.targets.shape[1:]))], 0)
    out.total_charges = torch.cat([batch.total_charges, torch.zeros(num_dummy)])
    if batch.atomic_numbers is not None:
        out.atomic_numbers = torch.cat([batch.atomic_numbers, zl(pad)])
    out.smiles_list = list(batch.smiles_list) + [""] * num_dummy
    tetra = batch.final_tetrahedral_chiral_tensor
    if num_tetra is not None and tetra is not None and tetra.numel() > 0:
        M = int(tetra.shape[0])
        if M > num_tetra:
            raise ValueError(f"batch has {M} tetrahedral centres, capacity {num_tetra}")
        if pad < 4:
            raise ValueError("padding tetrahedral centres needs at least 4 padding atoms")
        dummy = torch.arange(N, N + 4, dtype=torch.long).repeat(num_tetra - M, 1)
        tetra = torch.cat([tetra.reshape(-1, 4), dummy], 0)
    gi = GraphIndex.build(batch.multi_hop_edge_indices, out.batch_indices, B + num_dummy, gi0.num_hops,
                          out.atom_features_map, feature_sizes or {k: v[2] for k, v in gi0.embed.items()},
                          tetra, batch.final_cis_tensor, batch.final_trans_tensor, tile_rows)
    if pad_cistrans:
        if pad < 1:
            raise ValueError("padding the cis/trans updates needs at least one padding atom")
        src, tgt, sign, n = gi.cistrans if gi.cistrans is not None else (torch.zeros(0, dtype=torch.int32),) * 2 + (
            torch.zeros(0), 0)
        fill = 4 - n
        gi.cistrans = (torch.cat([src, torch.full((fill,), N, dtype=torch.int32)]),
                       torch.cat([tgt, torch.full((fill,), N, dtype=torch.int32)]),
                       torch.cat([sign.float(), torch.zeros(fill)]), 4)
    if not (gi.collapsed and gi.tile_local):
        raise ValueError("static padding supports the shipped collation only (no edge leaves its molecule)")

    def grow(t: torch.Tensor, n: int, fill: int) -> torch.Tensor:
        return t if t.numel() >= n else torch.cat([t, torch.full((n - t.numel(),), fill, dtype=t.dtype)])

    gi.col = grow(gi.col, num_edges + GraphIndex.INDEX_SLACK, 0)
    gi.col_t = grow(gi.col_t, num_edges + GraphIndex.INDEX_SLACK, 0)
    nt = gi.n_tiles if num_tiles is None else int(num_tiles)
    if gi.n_tiles > nt:
        raise ValueError(f"batch needs {gi.n_tiles} row tiles, capacity {nt}")
    gi.tile_ptr = grow(gi.tile_ptr, nt + 1, num_atoms)
    gi.n_tiles = nt
    gi.refresh_tile_info()                      # trailing tiles are empty: {N, N, E, E}
    # capacities, not the batch's own maxima (pass tile_rows >= the largest molecule for a batch-independent signature)
    gi.max_tile_rows = max(int(tile_rows), 29, gi.max_tile_rows)
    gi.max_tile_edges = max(gi.max_tile_edges, int(max_tile_edges))
    gi.max_seg = max(int(tile_rows), 29, gi.max_seg)
    out.graph_index = gi
    out.num_real_graphs = B
    return out


def static_signature(batch: MolBatch) -> tuple:
    """Everything a captured step bakes in: two padded batches are interchangeable iff their signatures are equal."""
    gi = batch.graph_index
    return (gi.num_atoms, gi.num_graphs, gi.num_rows, gi.n_tiles, gi.max_tile_rows, gi.max_tile_edges, gi.max_seg,
            int(gi.col.numel()),
            gi.collapsed, gi.tile_local, gi.unique_edges, None if gi.tetra is None else gi.tetra[3],
            None if gi.cistrans is None else gi.cistrans[3], getattr(batch, "num_real_graphs", gi.num_graphs),
            tuple(sorted((k, v[2]) for k, v in gi.embed.items())))


def collate_fn(data_list):
    """Drop-in for ``iterable_collate_fn`` (datasets/loaders.py:10-15)."""
    kept = [d for d in data_list if d is not None]
    return MolBatch.from_data_list(kept) if kept else None
