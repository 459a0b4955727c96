"""Pre-collated binary CSR shards: the on-disk / wire format that feeds the captured training step.

The reference streams per-molecule pickled dicts out of an HDF5 vlen-uint8 dataset (``datasets/features.py:416-572``),
decodes and collates them in Python per sample (``datasets/molecular.py:258-329,339-458``: ~0.1-0.2 ms per molecule), which
cannot feed >= 10^5 molecules/s per GPU.  Here the unit on disk is a whole *padded, collated batch* in exactly the byte
layout of the captured step's device slot (``trainer.StaticSlot``: every batch tensor and every CSR / segment / tile artefact
of its ``GraphIndex`` at a 256-byte aligned offset of ONE buffer).  Reading a batch is one ``memcpy`` of the record from the
(page-cached / memory-mapped) file into a pinned buffer and ONE host-to-device copy -- no per-sample Python at all.

File layout (little endian)::

    0    8   magic  b"AX2DSHRD"
    8    4   u32    version (1)
    12   4   u32    length L of the JSON header
    16   L   JSON   {"arena_bytes", "n_batches", "signature", "tensors": [{"name", "dtype", "shape", "offset"}...],
                     "graph_index": {scalars of GraphIndex}, "num_real_graphs", "meta": {...}}
    ...      zero padding to a multiple of 4096
    then n_batches records of ``record_bytes`` (= arena_bytes rounded up to 4096) each.

Sharding over ranks and DataLoader workers follows ``molecular.py:209-250`` with the BATCH as the unit: optional
rank-seeded shuffle of the batch order, contiguous chunks of ceil(n / world) per rank, then contiguous chunks of
ceil(n_rank / workers) per worker.
"""
from __future__ import annotations

import json
import math
import mmap
import os
import random
import struct
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

from .collate import GraphIndex, MolBatch, static_signature
from .trainer import HostBatch, _arena_layout, _batch_tensors

MAGIC = b"AX2DSHRD"
VERSION = 1
_ALIGN = 4096
_GI_SCALARS = ("num_atoms", "num_graphs", "num_edges", "num_hops", "num_rows", "collapsed", "tile_local", "unique_edges", "n_tiles",
               "max_tile_rows", "max_tile_edges", "max_seg")
_DT = {"int32": torch.int32, "int64": torch.int64, "float32": torch.float32}


def _tensor_names(b: MolBatch) -> List[str]:
    """Names of ``trainer._batch_tensors(b)``, in the same order."""
    gi = b.graph_index
    names = ["batch_indices", "targets", "total_charges"]
    names += [f"feat.{k}" for k in sorted(b.atom_features_map)]
    names += ["gi.rowptr", "gi.col", "gi.rowptr_t", "gi.col_t", "gi.seg_ptr", "gi.tile_ptr", "gi.tile_info", "gi.tile_info_t"]
    for k in sorted(gi.embed):
        names += [f"embed.{k}.order", f"embed.{k}.ptr"]
    if gi.tetra is not None:
        names += ["tetra.idx", "tetra.slot_ptr", "tetra.slot_idx"]
    if gi.cistrans is not None:
        names += ["cistrans.src", "cistrans.tgt", "cistrans.sign"]
    return names


def _signature_json(sig) -> list:
    return json.loads(json.dumps(sig))            # tuples -> lists, in the order static_signature defines


class ShardWriter:
    """``with ShardWriter(path) as w: w.add(padded_batch) ...`` -- every batch must have the static signature of the first."""

    def __init__(self, path: str, meta: Optional[Dict] = None):
        self.path = path
        self.meta = dict(meta or {})
        self._fh = None
        self._header = None
        self._offsets = None
        self._n = 0
        self._record = 0
        self._real: List[int] = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def _start(self, b: MolBatch) -> None:
        ts = _batch_tensors(b)
        names = _tensor_names(b)
        assert len(ts) == len(names)
        self._offsets, arena = _arena_layout(ts)
        gi = b.graph_index
        self._header = {
            "arena_bytes": int(arena), "n_batches": 0, "signature": _signature_json(static_signature(b)),
            "tensors": [{"name": n, "dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape), "offset": int(o)}
                        for n, t, o in zip(names, ts, self._offsets)],
            "graph_index": {**{k: (bool(getattr(gi, k)) if isinstance(getattr(gi, k), bool) else int(getattr(gi, k)))
                               for k in _GI_SCALARS},
                            "embed_vocab": {k: int(v[2]) for k, v in gi.embed.items()},
                            "tetra_M": None if gi.tetra is None else int(gi.tetra[3]),
                            "cistrans_n": None if gi.cistrans is None else int(gi.cistrans[3])},
            "num_real_graphs": [], "meta": self.meta}
        self._record = (arena + _ALIGN - 1) // _ALIGN * _ALIGN
        self._fh = open(self.path, "wb")
        self._sig = static_signature(b)
        # the header is rewritten at close(); reserve room for it (real-graph counts grow with the number of batches)
        self._data_start = 1 << 20
        self._fh.write(b"\0" * self._data_start)

    def add(self, padded: MolBatch) -> None:
        if self._fh is None:
            self._start(padded)
        if static_signature(padded) != self._sig:
            raise ValueError("all batches of a shard must share one static signature (pad them with the same capacities)")
        buf = np.zeros(self._record, dtype=np.uint8)
        for t, o in zip(_batch_tensors(padded), self._offsets):
            if t.numel():
                a = t.detach().cpu().contiguous().numpy()
                buf[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
        self._fh.write(buf.tobytes())
        self._real.append(int(getattr(padded, "num_real_graphs", padded.graph_index.num_graphs)))
        self._n += 1

    def close(self) -> None:
        if self._fh is None:
            return
        self._header["n_batches"] = self._n
        self._header["num_real_graphs"] = self._real
        self._header["record_bytes"] = self._record
        self._header["data_start"] = self._data_start
        js = json.dumps(self._header).encode()
        if 16 + len(js) > self._data_start:
            raise RuntimeError("shard header does not fit its reserved megabyte")
        self._fh.seek(0)
        self._fh.write(MAGIC + struct.pack("<II", VERSION, len(js)) + js)
        self._fh.close()
        self._fh = None


def shard_batch_indices(n: int, rank: int, world_size: int, worker_id: int = 0, num_workers: int = 1, shuffle: bool = False,
                        seed: int = 42, epoch_seed: int = 0) -> List[int]:
    """Which batches of an n-batch shard a (rank, worker) pair reads: ``molecular.py:209-250`` with the batch as the unit."""
    total = list(range(n))
    if shuffle:                                                    # molecular.py:211-226: one shuffle per rank-seeded RNG
        combined = (epoch_seed + seed + rank * 10000) % (2 ** 32 - 1)
        random.Random(combined).shuffle(total)
    if world_size > 1:                                             # molecular.py:229-237
        chunk = int(math.ceil(n / float(world_size)))
        total = total[rank * chunk: min(rank * chunk + chunk, n)]
    if num_workers > 1:                                            # molecular.py:240-250
        per = int(math.ceil(len(total) / float(num_workers)))
        total = total[worker_id * per: min(worker_id * per + per, len(total))]
    return total


class ShardDataset(torch.utils.data.IterableDataset):
    """Iterates ``HostBatch`` objects (pinned, in the captured step's slot layout) over this rank's / worker's batches.

    ``batch(i)`` rebuilds batch i as a padded ``MolBatch`` whose tensors are views of the record (what ``capture()`` and the
    parity tests need)."""

    def __init__(self, path: str, rank: int = 0, world_size: int = 1, shuffle: bool = False, seed: int = 42, pin: bool = True,
                 ring: int = 4, loop: bool = False):
        self.path, self.rank, self.world_size = path, int(rank), int(world_size)
        self.shuffle, self.seed, self.pin, self.ring, self.loop = shuffle, seed, pin, max(int(ring), 2), loop
        with open(path, "rb") as fh:
            head = fh.read(16)
            if head[:8] != MAGIC:
                raise ValueError(f"{path} is not an AX2D shard")
            version, n = struct.unpack("<II", head[8:16])
            if version != VERSION:
                raise ValueError(f"shard version {version}, expected {VERSION}")
            self.header = json.loads(fh.read(n).decode())
        self.n_batches = int(self.header["n_batches"])
        self.arena_bytes = int(self.header["arena_bytes"])
        self.record_bytes = int(self.header["record_bytes"])
        self.data_start = int(self.header["data_start"])
        self.signature = self._sig_tuple(self.header["signature"])
        self._mm = None
        self._pinned: List[torch.Tensor] = []
        self.epoch = 0

    @staticmethod
    def _sig_tuple(sig):
        return tuple(tuple(tuple(x) if isinstance(x, list) else x for x in s) if isinstance(s, list) else s for s in sig)

    def _map(self):
        if self._mm is None:
            self._fh = open(self.path, "rb")
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        return self._mm

    def __len__(self) -> int:
        return len(shard_batch_indices(self.n_batches, self.rank, self.world_size))

    def record(self, i: int) -> np.ndarray:
        if not 0 <= i < self.n_batches:
            raise IndexError(i)
        lo = self.data_start + i * self.record_bytes
        return np.frombuffer(self._map(), dtype=np.uint8, count=self.arena_bytes, offset=lo)

    def host_batch(self, i: int, slot: int = 0) -> HostBatch:
        """Record i copied into the pinned ring buffer ``slot`` (the buffer is reused: consume it before it comes round)."""
        while len(self._pinned) <= slot:
            t = torch.empty(self.arena_bytes, dtype=torch.uint8)
            self._pinned.append(t.pin_memory() if (self.pin and torch.cuda.is_available()) else t)
        dst = self._pinned[slot]
        dst.numpy()[:] = self.record(i)
        return HostBatch(dst, self.signature)

    def batch(self, i: int) -> MolBatch:
        """The padded MolBatch of record i (tensors are copies of the record; index layout as written)."""
        rec = torch.from_numpy(self.record(i).copy())
        ts = {}
        for d in self.header["tensors"]:
            dt = _DT[d["dtype"]]
            n = int(np.prod(d["shape"])) if d["shape"] else 1
            nbytes = n * torch.empty((), dtype=dt).element_size()
            ts[d["name"]] = rec[d["offset"]: d["offset"] + nbytes].view(dt).view(d["shape"]) if n else torch.empty(d["shape"], dtype=dt)
        h = self.header["graph_index"]
        gi = GraphIndex()
        for k in _GI_SCALARS:
            setattr(gi, k, h[k])
        for k in ("rowptr", "col", "rowptr_t", "col_t", "seg_ptr", "tile_ptr", "tile_info", "tile_info_t"):
            setattr(gi, k, ts["gi." + k])
        gi.embed = {k: (ts[f"embed.{k}.order"], ts[f"embed.{k}.ptr"], int(v)) for k, v in h["embed_vocab"].items()}
        if h["tetra_M"] is not None:
            gi.tetra = (ts["tetra.idx"], ts["tetra.slot_ptr"], ts["tetra.slot_idx"], int(h["tetra_M"]))
        if h["cistrans_n"] is not None:
            gi.cistrans = (ts["cistrans.src"], ts["cistrans.tgt"], ts["cistrans.sign"], int(h["cistrans_n"]))
        b = MolBatch()
        b.batch_indices = b.batch = ts["batch_indices"]
        b.targets, b.total_charges = ts["targets"], ts["total_charges"]
        b.atom_features_map = {n[5:]: t for n, t in ts.items() if n.startswith("feat.")}
        # the kernels read the CSR, not the [E, 2] edge list; a one-row placeholder keeps ``edges.numel() > 0`` (gnn.py:287)
        b.multi_hop_edge_indices = torch.zeros((1 if gi.num_edges else 0, 2), dtype=torch.long)
        b.final_tetrahedral_chiral_tensor = torch.zeros((0, 4), dtype=torch.long)
        b.final_cis_tensor = torch.zeros((0, 2), dtype=torch.long)
        b.final_trans_tensor = torch.zeros((0, 2), dtype=torch.long)
        b.x = torch.zeros((gi.num_atoms, 1))
        b.graph_index = gi
        b.num_real_graphs = int(self.header["num_real_graphs"][i])
        return b

    def indices(self) -> List[int]:
        info = torch.utils.data.get_worker_info()
        wid, nw = (info.id, info.num_workers) if info is not None else (0, 1)
        return shard_batch_indices(self.n_batches, self.rank, self.world_size, wid, nw, self.shuffle, self.seed,
                                   torch.initial_seed() % (2 ** 32 - 1) + self.epoch if self.shuffle else 0)

    def __iter__(self) -> Iterator[HostBatch]:
        ids = self.indices()
        k = 0
        while True:
            for i in ids:
                yield self.host_batch(i, k % self.ring)
                k += 1
            if not self.loop or not ids:
                return
            self.epoch += 1


def write_shard(path: str, padded_batches, meta: Optional[Dict] = None) -> int:
    """Write an iterable of padded ``MolBatch`` objects (one static signature); returns the number of batches."""
    with ShardWriter(path, meta) as w:
        for b in padded_batches:
            w.add(b)
        return w._n
