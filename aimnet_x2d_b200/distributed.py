"""Data-parallel plumbing: one process per GPU over NCCL / NVLink (reference ``src/utils/distributed.py:12-228``
and ``src/main/utils.py:24-76``).  The hot collective -- the per-step gradient all-reduce -- lives in
``optim.FlatAdam.all_reduce_grads``; this module holds process-group setup and the slow-path helpers with the
reference's names and degrade-to-no-op behaviour when torch.distributed is not initialised.
"""
from __future__ import annotations

import os
import pickle
from typing import Any, List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def setup_distributed_environment(backend: str = None) -> Tuple[torch.device, bool, int, int]:
    """``main/utils.py:24-76``: (device, is_ddp, local_rank, world_size) from LOCAL_RANK / WORLD_SIZE."""
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world_size = int(os.environ.get("WORLD_SIZE", 1))
    cuda = torch.cuda.is_available()
    if "LOCAL_RANK" in os.environ and world_size > 1:
        if backend is None:
            backend = "nccl" if cuda else "gloo"
        if cuda:
            torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device(f"cuda:{local_rank}")
            dist.init_process_group(backend=backend, **kw)
        device = torch.device(f"cuda:{local_rank}") if cuda else torch.device("cpu")
        return device, True, local_rank, world_size
    return (torch.device("cuda") if cuda else torch.device("cpu")), False, 0, 1


def is_main_process() -> bool:
    return (not dist.is_available()) or (not dist.is_initialized()) or dist.get_rank() == 0


def safe_get_rank() -> int:
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def get_world_size() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def shard_indices(n: int, rank: int, world_size: int) -> np.ndarray:
    """Contiguous shard of ``range(n)`` per rank, ceil(n / W) each (``datasets/molecular.py:228-237``,
    ``inference/pipeline.py:282-310``)."""
    per = (n + world_size - 1) // world_size
    lo = min(rank * per, n)
    return np.arange(lo, min(lo + per, n))


def gather_ndarray_to_rank0(arr: np.ndarray, device: str = "cpu") -> np.ndarray:
    """``distributed.py:49-95``: padded all_gather, concatenation on rank 0, empty array elsewhere."""
    if not (dist.is_available() and dist.is_initialized()):
        return arr
    local = torch.from_numpy(np.ascontiguousarray(arr)).float().to(device)
    size = torch.tensor([local.size(0)], dtype=torch.long, device=device)
    W = dist.get_world_size()
    sizes = [torch.zeros_like(size) for _ in range(W)]
    dist.all_gather(sizes, size)
    mx = max(int(s.item()) for s in sizes)
    if local.size(0) < mx:
        pad = torch.zeros((mx - local.size(0),) + tuple(local.shape[1:]), device=device)
        local = torch.cat([local, pad], dim=0)
    out = [torch.zeros_like(local) for _ in range(W)]
    dist.all_gather(out, local)
    if dist.get_rank() == 0:
        return np.concatenate([o[: int(s.item())].cpu().numpy() for o, s in zip(out, sizes)], axis=0)
    return np.array([], dtype=arr.dtype)


def gather_tensor_to_rank0(t: torch.Tensor) -> torch.Tensor:
    """Device-resident replacement of ``gather_ndarray_to_rank0`` for the evaluation path (``training/evaluator.py:158-187``):
    rows of ``t`` ([n_r, ...], different n_r per rank) are exchanged with ONE size all_gather + ONE padded all_gather on the
    tensor's own device (NCCL for CUDA tensors) -- no ``.cpu().numpy()`` round trip, no float cast, no pickling.  Rank 0 gets
    the concatenation in rank order, the other ranks an empty tensor."""
    if not (dist.is_available() and dist.is_initialized()):
        return t
    W = dist.get_world_size()
    size = torch.tensor([t.shape[0]], dtype=torch.long, device=t.device)
    sizes = torch.zeros(W, dtype=torch.long, device=t.device)
    dist.all_gather_into_tensor(sizes, size)
    sizes = [int(s) for s in sizes.tolist()]
    mx = max(sizes) if sizes else 0
    local = t.contiguous()
    if local.shape[0] < mx:
        local = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((W * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local)
    if dist.get_rank() != 0:
        return t.new_empty((0,) + tuple(t.shape[1:]))
    out = out.view((W, mx) + tuple(local.shape[1:]))
    return torch.cat([out[r, : sizes[r]] for r in range(W)], 0)


def gather_strings_to_rank0(local_list: List[str], device: str = "cpu") -> List[str]:
    """``distributed.py:98-144``."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_list
    W = dist.get_world_size()
    objs: List[Any] = [None] * W
    dist.all_gather_object(objs, list(local_list))
    if dist.get_rank() == 0:
        return [s for part in objs for s in part]
    return []


def broadcast_object(obj: Any, src_rank: int = 0, device: str = "cpu") -> Any:
    """``distributed.py:147-185``."""
    if not (dist.is_available() and dist.is_initialized()):
        return obj
    box = [obj if dist.get_rank() == src_rank else None]
    dist.broadcast_object_list(box, src=src_rank)
    return box[0]


def all_reduce_tensor(tensor: torch.Tensor, op: str = "sum") -> torch.Tensor:
    """``distributed.py:188-220``."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensor
    ops = {"sum": dist.ReduceOp.SUM, "mean": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
    if op not in ops:
        raise ValueError(f"Unsupported operation: {op}")
    dist.all_reduce(tensor, op=ops[op])
    if op == "mean":
        tensor /= dist.get_world_size()
    return tensor


def barrier() -> None:
    """``distributed.py:223-228``."""
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
