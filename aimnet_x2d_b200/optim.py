"""Flat-arena optimiser step: gradient all-reduce (NCCL) -> fused clip_grad_norm_ + Adam.

Replaces, for the training step of ``src/training/trainer.py:130-165``:
  * ``DistributedDataParallel(find_unused_parameters=True)`` (``main/runner.py:703-707``): parameters and
    gradients live in two flat 16-byte aligned fp32 arenas (parameters are views, state_dict keys/shapes are
    unchanged); rank 0's parameters are broadcast when the optimiser is built (what DDP does when it wraps the
    module); the gradient arena is summed over ranks per step -- in one collective, or bucketed in backward order so
    that NCCL overlaps the rest of backward (``bucket_ranges`` / ``all_reduce_bucket``); parameters that never receive
    a gradient (``long_range_projection``, ``stereochemical_embedding``) simply stay zero in the arena.
  * ``clip_grad_norm_(params, 1.0)`` + ``Adam.step()``: ``ax2d_sqnorm`` + ``ax2d_clip_adam_dev`` (two passes, fixed
    order, 1/world scaling folded in, bias corrections from a device-side step counter, every hyper-parameter read
    from device memory -> no host sync, CUDA-graph capturable, and a captured step follows LR schedulers).

``FlatAdam`` is a ``torch.optim.Optimizer``: ``param_groups`` (one contiguous arena range per group, so the reference's
``layer_wise_lr_decay`` groups and ``ReduceLROnPlateau`` / ``StepLR`` ... drive it unchanged, ``training/trainer.py:60-93``),
``state_dict`` / ``load_state_dict`` and ``zero_grad`` behave as callers of ``torch.optim.Adam`` expect.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .ops import _p, _stream


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 2.5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: Optional[float] = 1.0, process_group=None, sync_params: bool = True):
        defaults = dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps))
        super().__init__(params, defaults)
        self.params: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"]]
        if not self.params:
            raise ValueError("FlatAdam got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs CUDA parameters (no CPU fallback)")
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.group = process_group
        offs, n = [], 0
        self.group_ranges: List[Tuple[int, int]] = []          # arena range [lo, hi) of every param group
        for g in self.param_groups:
            lo = n
            for p in g["params"]:
                if p.dtype != torch.float32:
                    raise RuntimeError("FlatAdam keeps fp32 master parameters")
                if p.device != dev:
                    raise RuntimeError("FlatAdam needs all parameters on one device")
                offs.append(n)
                n += (p.numel() + 3) // 4 * 4                   # keep every view 16-byte aligned
            self.group_ranges.append((lo, n))
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
        # {grad_scale, max_norm, lr, beta1, beta2, eps} per group, on the device (read by the kernel) and pinned on the host
        ng = len(self.param_groups)
        self.hyper = torch.zeros((ng, 8), dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros((ng, 8), dtype=torch.float32).pin_memory()
        self._hyper_sent = None
        self.sync_hyper()
        if sync_params:
            self.sync_parameters()

    # ------------------------------------------------------------------ reference-facing conveniences
    @property
    def lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    @lr.setter
    def lr(self, value: float) -> None:
        for g in self.param_groups:
            g["lr"] = float(value)

    @property
    def betas(self):
        return tuple(self.param_groups[0]["betas"])

    @property
    def eps(self) -> float:
        return float(self.param_groups[0]["eps"])

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def sync_parameters(self, src: int = 0) -> None:
        """Every rank takes rank ``src``'s parameters (and optimiser state): what ``DistributedDataParallel`` does when
        it wraps a module (runner.py:703-707).  Without it differently seeded replicas diverge silently, because only
        gradients are exchanged afterwards."""
        if self.world_size() == 1:
            return
        for t in (self.flat_param, self.exp_avg, self.exp_avg_sq, self.step_count):
            dist.broadcast(t, src=src, group=self.group)

    def zero_grad(self, set_to_none: bool = True) -> None:        # noqa: ARG002  (the arena is never dropped)
        """One memset of the arena (``optimizer.zero_grad(set_to_none=True)`` equivalent, trainer.py:130)."""
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):     # re-attach if autograd / the user dropped the view
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    # ------------------------------------------------------------------ the exchange step
    def all_reduce_grads(self, async_op: bool = False):
        """SUM over ranks of the whole gradient arena in ONE collective (the 1/world factor is applied in the
        fused update: DDP's mean-of-per-rank-means semantics)."""
        if self.world_size() == 1:
            return None
        return dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def bucket_ranges(self, boundaries: List[torch.nn.Parameter]) -> List[Tuple[int, int]]:
        """Arena ranges cut at the given parameters (each boundary parameter starts a new range), first range first.
        Backward produces gradients from the END of the parameter list (head) towards its start (embeddings), so the
        ranges are reduced last-to-first while backward still runs (``trainer.GraphedTrainStep``)."""
        cuts = sorted({self.offsets[[id(q) for q in self.params].index(id(p))] for p in boundaries} | {0})
        cuts = [c for c in cuts if c < self.numel]
        return [(lo, hi) for lo, hi in zip(cuts, cuts[1:] + [self.numel]) if hi > lo]

    def all_reduce_bucket(self, lo: int, hi: int) -> None:
        if self.world_size() > 1:
            dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    # ------------------------------------------------------------------ the update
    def sync_hyper(self) -> None:
        """Upload the hyper-parameters when (and only when) the host values changed -- a 32-byte-per-group async copy
        from pinned memory on the current stream, ordered before the next update kernel.  Called by ``step()`` and by
        ``GraphedTrainStep.replay`` (a captured graph reads them from device memory)."""
        w = self.world_size()
        vals = tuple((1.0 / w, self.max_grad_norm, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                      float(g["eps"])) for g in self.param_groups)
        if vals == self._hyper_sent:
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("optimizer hyper-parameters changed while a CUDA graph is being captured")
        for i, v in enumerate(vals):
            self._hyper_host[i, :6] = torch.tensor(v, dtype=torch.float32)
        self.hyper.copy_(self._hyper_host, non_blocking=True)
        self._hyper_sent = vals

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        self.sync_hyper()
        _lib.check(lib.ax2d_sqnorm(_p(self.flat_grad), self.numel, _p(self.norm2), _p(self.step_count), _p(self._ws),
                                   _stream()), "ax2d_sqnorm")
        for gi, (lo, hi) in enumerate(self.group_ranges):
            at = lambda t: C.c_void_p(t.data_ptr() + 4 * lo)
            _lib.check(lib.ax2d_clip_adam_dev(at(self.flat_param), at(self.flat_grad), at(self.exp_avg),
                                              at(self.exp_avg_sq), hi - lo, _p(self.norm2),
                                              C.c_void_p(self.hyper.data_ptr() + 32 * gi), _p(self.step_count),
                                              _stream()), "ax2d_clip_adam_dev")
        return loss

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm of the last step (after the 1/world scaling), as a device scalar."""
        return self.norm2.sqrt() / self.world_size()

    # ------------------------------------------------------------------ checkpoints
    def state_dict(self):
        groups = []
        start = 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = list(range(start, start + len(g["params"])))
            start += len(g["params"])
            groups.append(d)
        return {"state": {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                          "step": self.step_count.detach().clone(), "numel": self.numel},
                "param_groups": groups}

    def load_state_dict(self, state_dict) -> None:
        st = state_dict["state"]
        if int(st["numel"]) != self.numel or len(state_dict["param_groups"]) != len(self.param_groups):
            raise ValueError("optimizer state does not match this parameter arena")
        with torch.no_grad():
            self.exp_avg.copy_(st["exp_avg"])
            self.exp_avg_sq.copy_(st["exp_avg_sq"])
            self.step_count.copy_(st["step"])
        for g, saved in zip(self.param_groups, state_dict["param_groups"]):
            for k, v in saved.items():
                if k != "params":
                    g[k] = v
        self.sync_hyper()
