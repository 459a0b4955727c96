"""The training step of the hot path: what one iteration of ``_training_epoch`` does
(reference ``src/training/trainer.py:116-170``) -- host batch -> device, ``zero_grad``, forward,
criterion, backward, (DDP) gradient all-reduce, ``clip_grad_norm_(1.0)``, ``Adam.step()``.

The reference's loop, logging, schedulers and early stopping stay with the reference; this class is the
call a user (or the reference's trainer) makes per batch.
"""
from __future__ import annotations

from typing import Optional

import torch

from .collate import MolBatch
from .optim import FlatAdam


class TrainStep:
    def __init__(self, model: torch.nn.Module, criterion: torch.nn.Module, optimizer: FlatAdam, device=None):
        self.model = model
        self.criterion = criterion
        self.optimizer = optimizer
        self.device = device if device is not None else next(model.parameters()).device

    def to_device(self, batch: MolBatch) -> MolBatch:
        """trainer.py:116-128: the 7 batch fields (+ targets) to the device; here also the CSR / segment
        artefacts emitted at collation.  ``non_blocking`` pays off when ``batch`` is pinned."""
        return batch.to(self.device, non_blocking=True)

    def device_step(self, bd: MolBatch) -> torch.Tensor:
        """One optimisation step on a device-resident batch; returns the loss as a device scalar (no sync)."""
        opt = self.optimizer
        opt.zero_grad()                                                     # trainer.py:130
        out, _, _ = self.model(bd.atom_features_map, bd.multi_hop_edge_indices, bd.batch_indices, bd.total_charges,
                               bd.final_tetrahedral_chiral_tensor, bd.final_cis_tensor, bd.final_trans_tensor,
                               graph_index=bd.graph_index)                  # trainer.py:151-159
        loss = self.criterion(out, bd.targets)                              # trainer.py:162
        loss.backward()                                                     # trainer.py:163
        opt.all_reduce_grads()                                              # DDP reducer (runner.py:703-707)
        opt.step()                                                          # trainer.py:164-165
        return loss.detach()

    def __call__(self, batch: MolBatch, return_float: bool = True):
        """Host batch in, loss out: includes the H2D copies and (``return_float``) the ``loss.item()``
        device->host read the reference does every step (trainer.py:169)."""
        loss = self.device_step(self.to_device(batch))
        return float(loss.item()) if return_float else loss
