"""Flat-arena optimiser step: gradient all-reduce (NCCL) -> fused clip_grad_norm_ + Adam.

Replaces, for the training step of ``src/training/trainer.py:130-165``:
  * ``DistributedDataParallel(find_unused_parameters=True)`` (``main/runner.py:703-707``): parameters and
    gradients live in two flat 16-byte aligned fp32 arenas (parameters are views, state_dict keys/shapes are
    unchanged); one ``ncclAllReduce(SUM)`` over the gradient arena per step; parameters that never receive a
    gradient (``long_range_projection``, ``stereochemical_embedding``) simply stay zero in the arena.
  * ``clip_grad_norm_(params, 1.0)`` + ``Adam.step()``: ``ax2d_sqnorm`` + ``ax2d_clip_adam`` (two passes, fixed
    order, 1/world scaling folded in, bias corrections from a device-side step counter -> no host sync and
    CUDA-graph capturable).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib
from .ops import _p, _stream


class FlatAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 2.5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: Optional[float] = 1.0, process_group=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs CUDA parameters (no CPU fallback)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.group = process_group
        offs, n = [], 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise RuntimeError("FlatAdam keeps fp32 master parameters")
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4                       # keep every view 16-byte aligned
        self.numel = n
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat_param[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        self.offsets = offs
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.norm2 = torch.zeros(1, dtype=torch.float32, device=dev)
        lib = _lib.load()
        self._ws = torch.empty(lib.ax2d_sqnorm_workspace(n) // 4, dtype=torch.float32, device=dev)

    def zero_grad(self) -> None:
        """One memset of the arena (``optimizer.zero_grad(set_to_none=True)`` equivalent, trainer.py:130)."""
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):     # re-attach if autograd / the user dropped the view
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def all_reduce_grads(self, async_op: bool = False):
        """SUM over ranks of the whole gradient arena in ONE collective (the 1/world factor is applied in the
        fused update: DDP's mean-of-per-rank-means semantics)."""
        if self.world_size() == 1:
            return None
        return dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def step(self) -> None:
        lib = _lib.load()
        w = self.world_size()
        _lib.check(lib.ax2d_sqnorm(_p(self.flat_grad), self.numel, _p(self.norm2), _p(self.step_count), _p(self._ws),
                                   _stream()), "ax2d_sqnorm")
        _lib.check(lib.ax2d_clip_adam(_p(self.flat_param), _p(self.flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq),
                                      self.numel, _p(self.norm2), 1.0 / w, self.max_grad_norm, self.lr, self.betas[0],
                                      self.betas[1], self.eps, _p(self.step_count), _stream()), "ax2d_clip_adam")

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm of the last step (after the 1/world scaling), as a device scalar."""
        return self.norm2.sqrt() / self.world_size()
