"""Forward-only use of the hot path: predictions, pooled molecule embeddings and partial charges
(reference callers: ``training/predictor.py:17-89``, ``training/extractors.py:16-70,278-465``,
``inference/embeddings.py:18-156``).  Batches are sharded over ranks with no communication
(``distributed.shard_indices`` == ``inference/pipeline.py:282-310``); every rank writes its own outputs.

``InferenceStep`` is the eager call; ``GraphedInferenceStep`` captures it once as a CUDA graph over static-shape
(padded) batches, like ``trainer.GraphedTrainStep``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .collate import MolBatch
from .trainer import HostBatch, StaticSlot


class InferenceStep:
    """One forward pass in eval mode; returns a dict with ``outputs [B, T]``, ``embeddings [B, hidden]`` (what the
    reference reads through a forward hook on ``model.pooling``), ``atom_embeddings [N, hidden]`` (hook on
    ``model.concat_self_other``), ``attention`` and ``partial_charges [N] | None``."""

    def __init__(self, model: torch.nn.Module, device=None, atom_embeddings: bool = False):
        self.model = model
        self.device = device if device is not None else next(model.parameters()).device
        self.want_atoms = atom_embeddings

    @torch.no_grad()
    def device_step(self, bd: MolBatch) -> Dict[str, Optional[torch.Tensor]]:
        model = self.model
        was_training = model.training
        model.eval()
        seen = {}
        h1 = model.pooling.register_forward_hook(lambda m, i, o: seen.__setitem__("pool", o[0]))
        h2 = model.concat_self_other.register_forward_hook(lambda m, i, o: seen.__setitem__("atom", o)) if self.want_atoms else None
        try:
            out, attn, q = model(bd.atom_features_map, bd.multi_hop_edge_indices, bd.batch_indices, bd.total_charges,
                                 bd.final_tetrahedral_chiral_tensor, bd.final_cis_tensor, bd.final_trans_tensor,
                                 graph_index=bd.graph_index)
        finally:
            h1.remove()
            if h2 is not None:
                h2.remove()
            model.train(was_training)
        b = getattr(bd, "num_real_graphs", out.shape[0])
        res = {"outputs": out[:b], "embeddings": seen["pool"][:b], "attention": attn, "partial_charges": q,
               "atom_embeddings": seen.get("atom")}
        return res

    def __call__(self, batch: MolBatch) -> Dict[str, Optional[torch.Tensor]]:
        return self.device_step(batch.to(self.device, non_blocking=True))


class GraphedInferenceStep(StaticSlot):
    """CUDA-graph replay of ``InferenceStep.device_step`` for padded batches of one static signature.  The returned
    tensors are static device buffers (valid until the next call)."""

    def __init__(self, model: torch.nn.Module, device=None, atom_embeddings: bool = False):
        self.eager = InferenceStep(model, device, atom_embeddings)
        self.device = self.eager.device
        self.slot: Optional[MolBatch] = None
        self.signature = None
        self.graph = None
        self.result = None

    def capture(self, padded: MolBatch, warmup: int = 2) -> None:
        self._make_slot(padded)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.eager.device_step(self.slot)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = self.eager.device_step(self.slot)
        torch.cuda.synchronize()

    def replay(self):
        self.graph.replay()
        return self.result

    def __call__(self, padded, prefetch: Optional[HostBatch] = None):
        """``padded``: a padded ``MolBatch`` or a ``HostBatch`` (``pack``); ``prefetch``: the ``HostBatch`` of a later call."""
        if isinstance(padded, HostBatch):
            if self.graph is None:
                raise RuntimeError("capture() with a padded MolBatch before passing HostBatch objects")
            self._load_host(padded)
        else:
            if self.graph is None:
                self.capture(padded)
            self.load(padded)
        res = self.replay()
        if prefetch is not None:
            self.prefetch(prefetch)
        return res


# ------------------------------------------------------------------------------------------------ output plumbing
class ShardedOutputWriter:
    """Per-rank output files for inference / embedding extraction -- every rank writes what it computed, nothing is
    gathered while the job runs (reference: ranks pickle partial results and rank 0 merges them,
    ``inference/pipeline.py:637-701``, ``inference/embeddings.py:376-498``; the evaluation path pads and all_gathers every
    array, ``utils/distributed.py:49-95``).

    Layout under ``directory``: ``<field>.rank<r>.bin`` (raw little-endian rows appended batch by batch) and
    ``index.rank<r>.json`` ({field: {dtype, row_shape, rows}}; ragged per-atom fields additionally get ``<field>_ptr`` = the
    molecule offsets).  ``merge_rank_outputs`` concatenates the rank files in rank order into ``.npy`` arrays."""

    def __init__(self, directory: str, rank: int = 0, world_size: int = 1):
        import os
        self.dir, self.rank, self.world = directory, int(rank), int(world_size)
        os.makedirs(directory, exist_ok=True)
        self._fh = {}
        self._index = {}
        self._atoms = 0

    def _file(self, field: str, arr):
        import os
        if field not in self._fh:
            self._fh[field] = open(os.path.join(self.dir, f"{field}.rank{self.rank}.bin"), "wb")
            self._index[field] = {"dtype": str(arr.dtype), "row_shape": list(arr.shape[1:]), "rows": 0}
        meta = self._index[field]
        if str(arr.dtype) != meta["dtype"] or list(arr.shape[1:]) != meta["row_shape"]:
            raise ValueError(f"field '{field}': rows of {arr.dtype}{list(arr.shape[1:])} after {meta['dtype']}{meta['row_shape']}")
        return self._fh[field]

    def append(self, field: str, rows) -> None:
        """Rows of a per-molecule field ([B, ...]) -- a host array or a (pinned) CPU tensor."""
        import numpy as np
        arr = np.ascontiguousarray(rows.numpy() if isinstance(rows, torch.Tensor) else rows)
        self._file(field, arr).write(arr.tobytes())
        self._index[field]["rows"] += int(arr.shape[0])

    def append_ragged(self, field: str, values, counts) -> None:
        """A per-atom field (e.g. partial charges [N]) with the number of atoms of each molecule of the batch."""
        import numpy as np
        counts = np.asarray(counts, dtype=np.int64)
        self.append(field, values)
        ptr = self._atoms + np.concatenate([[0], np.cumsum(counts)])[:-1]
        self.append(field + "_ptr", ptr.astype(np.int64))
        self._atoms += int(counts.sum())

    def close(self) -> None:
        import json
        import os
        for fh in self._fh.values():
            fh.close()
        self._fh = {}
        with open(os.path.join(self.dir, f"index.rank{self.rank}.json"), "w") as fh:
            json.dump({"world_size": self.world, "fields": self._index, "atoms": self._atoms}, fh)


def merge_rank_outputs(directory: str, world_size: int) -> Dict[str, "np.ndarray"]:
    """Rank 0, after a barrier: concatenate the per-rank files in rank order (contiguous rank shards of the input,
    ``distributed.shard_indices``, therefore come back in input order).  Writes ``<field>.npy`` and returns the arrays;
    ``*_ptr`` offsets are rebased onto the merged per-atom array."""
    import json
    import os

    import numpy as np
    idx = []
    for r in range(world_size):
        with open(os.path.join(directory, f"index.rank{r}.json")) as fh:
            idx.append(json.load(fh))
    out = {}
    atom_base = np.concatenate([[0], np.cumsum([i["atoms"] for i in idx])])
    for field in idx[0]["fields"]:
        parts = []
        for r in range(world_size):
            meta = idx[r]["fields"].get(field)
            if meta is None or meta["rows"] == 0:
                continue
            a = np.fromfile(os.path.join(directory, f"{field}.rank{r}.bin"), dtype=np.dtype(meta["dtype"]))
            a = a.reshape([meta["rows"]] + meta["row_shape"])
            if field.endswith("_ptr"):
                a = a + atom_base[r]
            parts.append(a)
        if parts:
            out[field] = np.concatenate(parts, 0)
            np.save(os.path.join(directory, field + ".npy"), out[field])
    return out


class OutputFetcher:
    """Device -> host copies of a step's result tensors that OVERLAP the next forward pass: the static result buffers of a
    graph-replayed step are snapshotted into one of two device staging sets on the compute stream (a ~3 us copy), the
    device-to-host copy of the snapshot runs on a side stream into pinned memory, and the compute stream only waits (on the
    device, never the host) before it overwrites a staging set whose copy is still in flight two steps later."""

    def __init__(self, device, fields=("outputs", "embeddings", "partial_charges"), depth: int = 2):
        self.device = torch.device(device)
        self.fields = tuple(fields)
        self.depth = int(depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self._stage = [dict() for _ in range(self.depth)]
        self._host = [dict() for _ in range(self.depth)]
        self._done = [None] * self.depth
        self._i = 0

    def fetch(self, res: Dict[str, Optional[torch.Tensor]]) -> int:
        """Start the copy of ``res`` (result dict of the step just enqueued); returns a handle for ``wait``."""
        k = self._i % self.depth
        self._i += 1
        cur = torch.cuda.current_stream(self.device)
        if self._done[k] is not None:
            cur.wait_event(self._done[k])                   # the staging set is free again (device-side wait)
        for f in self.fields:
            t = res.get(f)
            if t is None:
                continue
            st = self._stage[k].get(f)
            if st is None or st.shape != t.shape or st.dtype != t.dtype:
                st = torch.empty_like(t)
                self._stage[k][f] = st
                self._host[k][f] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            st.copy_(t, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for f, st in self._stage[k].items():
                self._host[k][f].copy_(st, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._done[k] = done
        return k

    def wait(self, handle: int) -> Dict[str, torch.Tensor]:
        """Host tensors (pinned, valid until ``depth`` more ``fetch`` calls) of the copy started by ``fetch``."""
        self._done[handle].synchronize()
        return self._host[handle]


class EmbeddingExtractor:
    """Forward-only pass over a sequence of (padded) batches with results streamed to a ``ShardedOutputWriter``: ONE forward
    per batch yields predictions, pooled embeddings and partial charges (the reference runs the model twice per batch for
    embeddings, ``inference/embeddings.py:109-119``), device-to-host copies run on a side stream from a snapshot of the result
    (``OutputFetcher``) so that the copy of batch i overlaps the forward of batch i + 1, and nothing is gathered across ranks."""

    def __init__(self, step, writer: ShardedOutputWriter, fields=("outputs", "embeddings", "partial_charges")):
        self.step, self.writer, self.fields = step, writer, tuple(fields)
        self._fetcher = None
        self._pending = [None, None]

    def _flush(self, k: int) -> None:
        if self._pending[k] is None:
            return
        handle, counts, n_real, n_atoms = self._pending[k]
        bufs = self._fetcher.wait(handle)
        for f, t in bufs.items():
            if f == "partial_charges":
                self.writer.append_ragged(f, t[:n_atoms].clone(), counts)
            else:
                self.writer.append(f, t[:n_real].clone())
        self._pending[k] = None

    def run(self, batches) -> int:
        """``batches``: iterable of padded ``MolBatch`` or ``HostBatch`` objects accepted by the step.  Returns the number of
        molecules written."""
        import numpy as np
        total = 0
        k = 1
        for i, b in enumerate(batches):
            k = i & 1
            self._flush(k)                                  # the pinned buffers of two batches ago are free again
            res = self.step(b)
            slot = self.step.slot if hasattr(self.step, "slot") and self.step.slot is not None else b
            gi = slot.graph_index
            n_real = int(getattr(slot, "num_real_graphs", gi.num_graphs))
            seg = gi.seg_ptr[: n_real + 1].cpu().numpy() if gi.seg_ptr.is_cuda else gi.seg_ptr[: n_real + 1].numpy()
            if self._fetcher is None:
                dev = next(t for t in res.values() if t is not None).device
                self._fetcher = OutputFetcher(dev, self.fields, depth=2)
            handle = self._fetcher.fetch(res)               # snapshot + device-to-host copy on a side stream
            self._pending[k] = (handle, np.diff(seg), n_real, int(seg[-1]))
            total += n_real
        self._flush(1 - k)                                  # the older of the two outstanding batches first: input order
        self._flush(k)
        return total
