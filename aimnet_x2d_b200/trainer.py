"""The training step of the hot path: what one iteration of ``_training_epoch`` does
(reference ``src/training/trainer.py:116-170``) -- host batch -> device, ``zero_grad``, forward,
criterion, backward, (DDP) gradient all-reduce, ``clip_grad_norm_(1.0)``, ``Adam.step()``.

The reference's loop, logging, schedulers and early stopping stay with the reference; this class is the
call a user (or the reference's trainer) makes per batch.
"""
from __future__ import annotations

from typing import Optional

import torch

import torch.distributed as dist

from .collate import MolBatch, static_signature
from .optim import FlatAdam


NVTX = bool(int(__import__("os").environ.get("AX2D_NVTX", "0")))      # AX2D_NVTX=1: NVTX ranges around the phases of a step


class _Range:
    """``with _Range("name"):`` -- an NVTX range when AX2D_NVTX=1 (for nsys / ncu --nvtx timelines), nothing otherwise."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


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
        with _Range("ax2d.forward"):
            opt.zero_grad()                                                 # trainer.py:130
            out, _, _ = self.model(bd.atom_features_map, bd.multi_hop_edge_indices, bd.batch_indices, bd.total_charges,
                                   bd.final_tetrahedral_chiral_tensor, bd.final_cis_tensor, bd.final_trans_tensor,
                                   graph_index=bd.graph_index)              # trainer.py:151-159
            loss = self.criterion(out, bd.targets)                          # trainer.py:162
        with _Range("ax2d.backward"):
            loss.backward()                                                 # trainer.py:163
        with _Range("ax2d.allreduce"):
            opt.all_reduce_grads()                                          # DDP reducer (runner.py:703-707)
        with _Range("ax2d.clip_adam"):
            opt.step()                                                      # trainer.py:164-165
        return loss.detach()

    def __call__(self, batch: MolBatch, return_float: bool = True):
        """Host batch in, loss out: includes the H2D copies and (``return_float``) the ``loss.item()``
        device->host read the reference does every step (trainer.py:169)."""
        loss = self.device_step(self.to_device(batch))
        return float(loss.item()) if return_float else loss


def _batch_tensors(b: MolBatch):
    """Every tensor of a (padded) batch that the step reads on the device, in a fixed order."""
    gi = b.graph_index
    ts = [b.batch_indices, b.targets, b.total_charges]
    ts += [b.atom_features_map[k] for k in sorted(b.atom_features_map)]
    ts += [gi.rowptr, gi.col, gi.rowptr_t, gi.col_t, gi.seg_ptr, gi.tile_ptr, gi.tile_info, gi.tile_info_t]
    for k in sorted(gi.embed):
        ts += [gi.embed[k][0], gi.embed[k][1]]
    # stereo: the kernels read the int32 artefacts of the GraphIndex, not the reference's index tensors
    if gi.tetra is not None:
        ts += list(gi.tetra[:3])
    if gi.cistrans is not None:
        ts += list(gi.cistrans[:3])
    return ts


_ARENA_ALIGN = 256


def _arena_layout(tensors):
    """Byte offsets of the batch tensors inside one contiguous buffer (256-byte aligned) and its size."""
    offs, n = [], 0
    for t in tensors:
        offs.append(n)
        n += (t.numel() * t.element_size() + _ARENA_ALIGN - 1) // _ARENA_ALIGN * _ARENA_ALIGN
    return offs, max(n, _ARENA_ALIGN)


class HostBatch:
    """A padded batch packed into ONE pinned host buffer with the layout of a captured step's device slot
    (``GraphedTrainStep.pack``): the whole batch crosses PCIe as a single copy, which can be issued ahead of time
    (``prefetch``) so that it overlaps the previous step."""

    def __init__(self, arena: torch.Tensor, signature):
        self.arena, self.signature = arena, signature

    def nbytes(self) -> int:
        return int(self.arena.numel())


class StaticSlot:
    """The static device copy of a padded batch that captured graphs read, kept as views of ONE device buffer, and the
    ways a batch gets into it: per tensor (``load``), as one packed buffer from the host (``HostBatch``, optionally
    prefetched on a side stream) or from the device (``load_arena``)."""

    slot: Optional[MolBatch] = None
    signature = None
    slot_arena = stage_arena = None
    _offsets = None
    _staged = None
    _copy_stream = None
    _stage_ready = _stage_free = None

    def _make_slot(self, padded: MolBatch) -> None:
        self.slot = padded.to(self.device)
        self.signature = static_signature(padded)
        # every tensor of the slot becomes a view of one device buffer (in place: all references stay valid), so that a
        # HostBatch arrives with a single copy
        ts = _batch_tensors(self.slot)
        self._offsets, total = _arena_layout(ts)
        self.slot_arena = torch.zeros(total, dtype=torch.uint8, device=self.device)
        self.stage_arena = torch.zeros(total, dtype=torch.uint8, device=self.device)
        for t, o in zip(ts, self._offsets):
            if t.numel():
                view = self.slot_arena[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
                view.copy_(t)
                t.set_(view)
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._stage_ready, self._stage_free = torch.cuda.Event(), torch.cuda.Event()

    def load(self, padded: MolBatch) -> None:
        """H2D copies of a padded host batch into the static slot, tensor by tensor (pinned source => asynchronous)."""
        if static_signature(padded) != self.signature:
            raise RuntimeError("batch does not match the static signature the step was captured for; pad it with the "
                               "same capacities (collate.pad_batch) or capture a new step")
        for dst, src in zip(_batch_tensors(self.slot), _batch_tensors(padded)):
            if dst.shape != src.shape:
                raise RuntimeError(f"static slot tensor {tuple(dst.shape)} vs batch tensor {tuple(src.shape)}")
            if dst.numel():
                dst.copy_(src, non_blocking=True)

    def pack(self, padded: MolBatch) -> HostBatch:
        """Pack a padded host batch into one pinned buffer with the slot's layout (host-side work of the data loader)."""
        if self.slot_arena is None:
            raise RuntimeError("capture() first: the layout of the host buffer is the captured slot's")
        if static_signature(padded) != self.signature:
            raise RuntimeError("batch does not match the static signature the step was captured for")
        arena = torch.zeros(self.slot_arena.numel(), dtype=torch.uint8).pin_memory()
        for src, dst, o in zip(_batch_tensors(padded), _batch_tensors(self.slot), self._offsets):
            if tuple(src.shape) != tuple(dst.shape) or src.dtype != dst.dtype:
                raise RuntimeError(f"static slot tensor {tuple(dst.shape)} vs batch tensor {tuple(src.shape)}")
            if src.numel():
                arena[o:o + src.numel() * src.element_size()].view(src.dtype).view(src.shape).copy_(src)
        return HostBatch(arena, self.signature)

    def prefetch(self, hb: HostBatch) -> None:
        """Start the host-to-device copy of a later batch on a side stream; it overlaps whatever the main stream runs."""
        if hb.signature != self.signature:
            raise RuntimeError("batch does not match the static signature the step was captured for")
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free)       # the staging buffer has been consumed
            self.stage_arena.copy_(hb.arena, non_blocking=True)
            self._stage_ready.record(self._copy_stream)
        self._staged = hb

    def _load_host(self, hb: HostBatch) -> None:
        if hb.signature != self.signature:
            raise RuntimeError("batch does not match the static signature the step was captured for")
        main = torch.cuda.current_stream(self.device)
        if self._staged is hb:                                    # prefetched: one device-to-device copy
            main.wait_event(self._stage_ready)
            self.slot_arena.copy_(self.stage_arena, non_blocking=True)
            self._stage_free.record(main)
            self._staged = None
        else:
            self.slot_arena.copy_(hb.arena, non_blocking=True)

    def load_arena(self, arena: torch.Tensor) -> None:
        """A batch that already lives on the device as one packed buffer (``HostBatch.arena.to(device)``)."""
        self.slot_arena.copy_(arena, non_blocking=True)


class GraphedTrainStep(StaticSlot):
    """The same training step captured ONCE as CUDA graphs and replayed for every batch of the same static
    signature (``collate.pad_batch``): no per-kernel host work is left in the loop, which is what lets a ~3 ms GPU
    step run at GPU speed (the eager path needs ~10 ms of Python / launch time for its ~250 launches).

    Per step: H2D copies of the batch into the static device slot -> graph A (zero_grad, forward, loss, backward)
    -> [world > 1: ONE NCCL all-reduce of the flat gradient arena] -> graph B (clip + Adam).  The loss is a static
    device scalar; ``__call__`` reads it back like ``trainer.py:169`` does.
    """

    def __init__(self, model: torch.nn.Module, criterion: torch.nn.Module, optimizer: FlatAdam, device=None,
                 single_graph: bool = False):
        """``single_graph``: capture forward + backward, the NCCL all-reduce and the update into ONE graph (one launch per
        step, no host round trip between backward and the collective).  Needs every rank to replay in lock step."""
        self.eager = TrainStep(model, criterion, optimizer, device)
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.device = self.eager.device
        self.slot: Optional[MolBatch] = None
        self.signature = None
        self.graph_fb = self.graph_opt = None
        self.loss = None
        self.single_graph = bool(single_graph)

    def _fwd_bwd(self) -> torch.Tensor:
        bd, opt = self.slot, self.optimizer
        opt.zero_grad()
        out, _, _ = self.model(bd.atom_features_map, bd.multi_hop_edge_indices, bd.batch_indices, bd.total_charges,
                               bd.final_tetrahedral_chiral_tensor, bd.final_cis_tensor, bd.final_trans_tensor,
                               graph_index=bd.graph_index)
        n = getattr(bd, "num_real_graphs", out.shape[0])          # dummy (padding) molecules carry no loss
        loss = self.criterion(out[:n], bd.targets[:n])
        loss.backward()
        return loss.detach()

    def capture(self, padded: MolBatch, warmup: int = 3, timer=None) -> None:
        """Allocate the static slot from ``padded`` (a host batch from ``pad_batch``), run ``warmup`` eager steps on a
        side stream (their effect on parameters / optimiser state is rolled back), then capture.  ``timer``: an
        ``ops.KernelTimer`` installed during the capture only -- its event pairs become nodes of the graphs, so every
        replay measures each launch inside the step."""
        import contextlib
        opt = self.optimizer
        self._make_slot(padded)
        saved = [t.clone() for t in (opt.flat_param, opt.exp_avg, opt.exp_avg_sq, opt.step_count)]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                opt.all_reduce_grads()
                opt.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph_fb = torch.cuda.CUDAGraph()
        with (timer if timer is not None else contextlib.nullcontext()):
            if self.single_graph:
                # thread-local capture mode: ProcessGroupNCCL's watchdog thread polls CUDA events while the capture runs
                with torch.cuda.graph(self.graph_fb, capture_error_mode="thread_local"):
                    self.loss = self._fwd_bwd()
                    opt.all_reduce_grads()
                    opt.step()
                self.graph_opt = None
            else:
                with torch.cuda.graph(self.graph_fb):
                    self.loss = self._fwd_bwd()
                self.graph_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_opt):
                    opt.step()
        with torch.no_grad():
            for dst, src in zip((opt.flat_param, opt.exp_avg, opt.exp_avg_sq, opt.step_count), saved):
                dst.copy_(src)
        torch.cuda.synchronize()

    def replay(self) -> torch.Tensor:
        if self.graph_opt is None:           # single graph: backward -> all-reduce -> update in one launch
            self.optimizer.sync_hyper()
            with _Range("ax2d.graph(forward+backward+allreduce+clip_adam)"):
                self.graph_fb.replay()
            return self.loss
        with _Range("ax2d.graph(forward+backward)"):
            self.graph_fb.replay()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            with _Range("ax2d.allreduce"):
                self.optimizer.all_reduce_grads()
        self.optimizer.sync_hyper()          # LR schedulers: the captured update reads its hyper-parameters from device memory
        with _Range("ax2d.graph(clip_adam)"):
            self.graph_opt.replay()
        return self.loss

    def __call__(self, padded, return_float: bool = True, prefetch: Optional[HostBatch] = None):
        """``padded``: a padded ``MolBatch`` (one copy per tensor) or a ``HostBatch`` from ``pack`` (one copy, possibly
        already under way).  ``prefetch``: the ``HostBatch`` of a later call, copied while this step runs."""
        with _Range("ax2d.h2d"):
            if isinstance(padded, HostBatch):
                if self.graph_fb is None:
                    raise RuntimeError("capture() with a padded MolBatch before passing HostBatch objects")
                self._load_host(padded)
            else:
                if self.graph_fb is None:
                    self.capture(padded)
                self.load(padded)
        loss = self.replay()
        if prefetch is not None:
            self.prefetch(prefetch)
        return float(loss.item()) if return_float else loss
