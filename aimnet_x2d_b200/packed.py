"""Packed projection weights: the padded / packed / TF32-split operands of every ``nn.Linear`` on the path, refreshed
from the reference-shaped parameters by ONE kernel per forward (``ax2d_pack_weights``), and the weight gradients
accumulated back into the parameters' ``.grad`` by ONE kernel per backward (``ax2d_unpack_grads``).

The parameters stay what the reference defines (names, shapes, ``state_dict``); everything here is a derived cache.
Without it every call pads / concatenates / splits its weights with ~190 small torch and split kernels per training
step and autograd adds ~60 gradient-accumulation kernels; with it the projections read ready operands and write
their weight gradients straight into the packed gradient buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_DESC = np.dtype([("src", "u8"), ("grad", "u8"), ("w", "u8"), ("hi", "u8"), ("lo", "u8"), ("hiT", "u8"), ("loT", "u8"),
                  ("g", "u8"), ("part", "u8"), ("rows", "i4"), ("cols", "i4"), ("src_ld", "i4"), ("dst_ld", "i4"),
                  ("dstT_ld", "i4"), ("split", "i4"), ("p_rows", "i4"), ("p_cols", "i4"), ("dr", "i4"), ("dc", "i4"),
                  ("wb", "u8"), ("wbT", "u8")])
assert _DESC.itemsize == 128


class PackInfo:
    """Attached to a packed weight tensor as ``t._ax2d``: its TF32 terms (plain and transposed) and the buffer its
    gradient is written to."""
    __slots__ = ("hi", "lo", "hiT", "loT", "grad", "owner", "written", "name", "partials", "wb", "wbT")

    def __init__(self, hi, lo, hiT, loT, grad, owner, name=""):
        self.hi, self.lo, self.hiT, self.loT, self.grad, self.owner, self.name = hi, lo, hiT, loT, grad, owner, name
        self.wb = self.wbT = None  # bf16 copies [rows, cols] / [cols, rows] (``PackedWeights.enable_bf16``)
        self.written = False      # the gradient buffer already holds a contribution of the running backward pass
        # split-K partials the weight-gradient kernel left for this matrix in the running backward pass:
        # (data_ptr, splits, rows, cols) -- summed by the gradient-collect kernel instead of a reduce launch per matrix
        self.partials = None


def _capturing(device) -> bool:
    return device.type == "cuda" and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def packed_info(t) -> Optional[PackInfo]:
    return None if t is None else getattr(t, "_ax2d", None)


class _Matrix:
    def __init__(self, name, rows, cols, blocks, vector):
        self.name, self.rows, self.cols, self.blocks, self.vector = name, rows, cols, blocks, vector
        self.off = 0


class PackedWeights:
    """``add(name, rows, cols, blocks)`` declares a packed matrix built from rectangular blocks of parameters:
    ``blocks = [(param, r0, r1, c0, c1, dst_r, dst_c), ...]`` (for a bias: rows = 1, r0 = 0, r1 = 1, columns = entries).
    ``finalize()`` allocates the (zero) arenas; ``refresh()`` / ``unpack_grads()`` launch the two kernels."""

    MAX_TABLES = 16          # descriptor tables kept (LRU); tables a captured CUDA graph reads are never dropped

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:       # "cuda" -> the current device, as tensors report it
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mats: Dict[str, _Matrix] = {}
        self.tensors: Dict[str, torch.Tensor] = {}
        self._table = None
        self._table_key = None
        self._n_blocks = 0
        self._max_elems = 1
        self.dirty = False
        self._workspaces: Dict[str, torch.Tensor] = {}     # persistent split-K workspaces, one per matrix
        # Buffers a captured CUDA graph may still read are never freed or rewritten: a second step captured on the same
        # model with other shapes gets its own descriptor table (one per configuration) and, if it needs a larger
        # workspace, the old one stays alive.
        self._tables: Dict[tuple, torch.Tensor] = {}
        self._pinned = set()
        self.bf16 = False
        self.wb = self.wbT = None
        self._retired: List[torch.Tensor] = []

    # a derived cache: copies / pickles of the owning module rebuild their own
    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (type(None), ())

    def add(self, name: str, rows: int, cols: int, blocks: Sequence[Tuple], vector: bool = False) -> None:
        for (p, r0, r1, c0, c1, dr, dc) in blocks:
            if p.dim() == 1:
                if not (r0 == 0 and r1 == 1 and 0 <= c0 <= c1 <= p.shape[0]):
                    raise ValueError(f"{name}: bad vector block")
            elif not (0 <= r0 <= r1 <= p.shape[0] and 0 <= c0 <= c1 <= p.shape[1]):
                raise ValueError(f"{name}: block outside its parameter")
            if dr + (r1 - r0) > rows or dc + (c1 - c0) > cols:
                raise ValueError(f"{name}: block outside the packed matrix")
        self.mats[name] = _Matrix(name, int(rows), int(cols), list(blocks), vector)

    def finalize(self) -> None:
        n = 0
        for m in self.mats.values():
            m.off = n
            n += (m.rows * m.cols + 3) // 4 * 4              # 16-byte aligned matrices
        z = lambda: torch.zeros(max(n, 4), dtype=torch.float32, device=self.device)
        self.w, self.hi, self.lo, self.hiT, self.loT, self.grad = z(), z(), z(), z(), z(), z()
        for m in self.mats.values():
            sl = slice(m.off, m.off + m.rows * m.cols)
            shape = (m.cols,) if m.vector else (m.rows, m.cols)
            t = self.w[sl].view(shape)
            gr = self.grad[sl].view(shape)
            if m.vector:
                info = PackInfo(None, None, None, None, gr, self, m.name)
            else:
                info = PackInfo(self.hi[sl].view(shape), self.lo[sl].view(shape), self.hiT[sl].view(m.cols, m.rows),
                                self.loT[sl].view(m.cols, m.rows), gr, self, m.name)
            t._ax2d = info
            self.tensors[m.name] = t
        self._n_blocks = sum(len(m.blocks) for m in self.mats.values())
        self._max_elems = max([1] + [(b[2] - b[1]) * (b[4] - b[3]) for m in self.mats.values() for b in m.blocks])

    def enable_bf16(self) -> None:
        """Also keep bf16 copies of every matrix (plain and transposed) for the bf16 configuration: written by the same
        pack launch, read by ``ax2d_gemm_bf16``."""
        if self.bf16:
            return
        n = self.w.numel()
        self.wb = torch.zeros(n, dtype=torch.bfloat16, device=self.device)
        self.wbT = torch.zeros(n, dtype=torch.bfloat16, device=self.device)
        for m in self.mats.values():
            if m.vector:
                continue
            sl = slice(m.off, m.off + m.rows * m.cols)
            info = self.tensors[m.name]._ax2d
            info.wb = self.wb[sl].view(m.rows, m.cols)
            info.wbT = self.wbT[sl].view(m.cols, m.rows)
        self.bf16 = True
        self._table = self._table_key = None          # descriptor tables carry the new pointers
        self._tables = {k: v for k, v in self._tables.items() if k in self._pinned}

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.tensors[name]

    def __contains__(self, name: str) -> bool:
        return name in self.tensors

    # ------------------------------------------------------------------ descriptor table
    def _params(self) -> List[torch.nn.Parameter]:
        return [b[0] for m in self.mats.values() for b in m.blocks]

    def _key(self, with_grads: bool):
        ps = self._params()
        parts = tuple((n, t._ax2d.partials) for n, t in self.tensors.items() if t._ax2d.partials is not None) if with_grads else None
        return (tuple(p.data_ptr() for p in ps) + (self.bf16,),
                tuple((p.grad.data_ptr() if p.grad is not None else 0) for p in ps) if with_grads else None, parts)

    def workspace(self, name: str, nbytes: int) -> torch.Tensor:
        """Persistent split-K workspace of one matrix (its partial tiles must outlive the backward pass)."""
        ws = self._workspaces.get(name)
        if ws is None or ws.numel() * 4 < nbytes:
            if ws is not None:
                self._retired.append(ws)
            ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=self.device)
            self._workspaces[name] = ws
        return ws

    def _build(self, with_grads: bool) -> None:
        rec = np.zeros(self._n_blocks, dtype=_DESC)
        i = 0
        for m in self.mats.values():
            base = 4 * m.off
            for (p, r0, r1, c0, c1, dr, dc) in m.blocks:
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != self.device:
                    raise RuntimeError(f"packed weights need contiguous fp32 parameters on {self.device} ({m.name})")
                ld = 1 if p.dim() == 1 else p.shape[1]
                src_off = 4 * (c0 if p.dim() == 1 else r0 * ld + c0)
                dst_off = base + 4 * (dr * m.cols + dc)
                dstT_off = base + 4 * (dc * m.rows + dr)
                d = rec[i]
                d["src"] = p.data_ptr() + src_off
                d["grad"] = (p.grad.data_ptr() + src_off) if (with_grads and p.grad is not None and p.requires_grad) else 0
                d["w"] = self.w.data_ptr() + dst_off
                d["g"] = self.grad.data_ptr() + dst_off
                if not m.vector:
                    d["hi"], d["lo"] = self.hi.data_ptr() + dst_off, self.lo.data_ptr() + dst_off
                    d["hiT"], d["loT"] = self.hiT.data_ptr() + dstT_off, self.loT.data_ptr() + dstT_off
                    if self.bf16:
                        d["wb"], d["wbT"] = self.wb.data_ptr() + dst_off // 2, self.wbT.data_ptr() + dstT_off // 2
                d["rows"], d["cols"] = r1 - r0, c1 - c0
                d["src_ld"] = (c1 - c0) if p.dim() == 1 else ld
                d["dst_ld"], d["dstT_ld"] = m.cols, m.rows
                part = self.tensors[m.name]._ax2d.partials if with_grads else None
                if part is not None:
                    d["part"], d["split"], d["p_rows"], d["p_cols"] = part
                    d["dr"], d["dc"] = dr, dc
                i += 1
        key = self._key(with_grads)
        if _capturing(self.device):
            # a pageable host-to-device copy is illegal inside a capture; the eager warm-up passes of capture() build every
            # table the captured pass needs, so getting here means the capture was started cold
            raise RuntimeError("packed-weight descriptor table would have to be built during CUDA graph capture; run "
                               "at least one eager step with the same parameters / gradients first")
        self._table = torch.from_numpy(rec.view(np.uint8).copy()).to(self.device)
        self._tables[key] = self._table
        self._table_key = key
        # bounded: a stock optimiser with zero_grad(set_to_none=True) hands out fresh .grad tensors every step, i.e. a new
        # key per step.  Tables a captured graph reads (``_pinned``) are never dropped.
        while len(self._tables) > self.MAX_TABLES:
            victim = next((k for k in self._tables if k not in self._pinned and k != key), None)
            if victim is None:
                break
            del self._tables[victim]

    def _ensure(self, with_grads: bool) -> None:
        if with_grads:
            for p in self._params():
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p)
        key = self._key(with_grads)
        if self._table is not None and self._table_key is not None and key[0] == self._table_key[0] and (
                not with_grads or (key[1] == self._table_key[1] and key[2] == self._table_key[2])):
            if _capturing(self.device):
                self._pinned.add(self._table_key)
            return                                       # the current table fits (a refresh only needs the parameters)
        cached = self._tables.get(key)
        if cached is not None:
            self._tables[key] = self._tables.pop(key)          # most recently used last
            self._table, self._table_key = cached, key
        else:
            self._build(with_grads)
        if _capturing(self.device):
            self._pinned.add(self._table_key)

    # ------------------------------------------------------------------ the two launches
    def _clear_grads(self) -> None:
        self.grad.zero_()
        for t in self.tensors.values():
            t._ax2d.written = False
            t._ax2d.partials = None
        self.dirty = False

    def refresh(self) -> None:
        from .ops import _stream
        self._ensure(False)
        if self.dirty:                       # gradients of an earlier backward that nobody collected
            self._clear_grads()
        _lib.check(_lib.load().ax2d_pack_weights(C.c_void_p(self._table.data_ptr()), self._n_blocks, self._max_elems,
                                                 _stream()), "ax2d_pack_weights")

    def unpack_grads(self) -> None:
        """p.grad += packed gradient, for every parameter block; the packed gradient buffers are cleared afterwards."""
        from .ops import _stream
        if not self.dirty:
            return
        self._ensure(True)
        _lib.check(_lib.load().ax2d_unpack_grads(C.c_void_p(self._table.data_ptr()), self._n_blocks, self._max_elems,
                                                 _stream()), "ax2d_unpack_grads")
        self._clear_grads()


class PackAnchorFn(torch.autograd.Function):
    """Identity on a leaf that every other operation of the forward pass depends on (an embedding table): its forward
    refreshes the packed weights, its backward -- the LAST node autograd executes -- adds the packed weight gradients
    into the parameters' ``.grad``."""

    @staticmethod
    def forward(ctx, pk: PackedWeights, t: torch.Tensor):
        pk.refresh()
        ctx.pk = pk
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        ctx.pk.unpack_grads()
        return None, g
