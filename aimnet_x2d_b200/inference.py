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
