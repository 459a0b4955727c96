"""B200-native graph pooling layers -- same classes / signatures / state_dict keys as the reference
``src/models/pooling.py`` (15-273); computation = the fused segment kernels of ``csrc/pool.cu``.

``forward(x[N, F], batch_indices[N]) -> (pooled[B, F], attention_weights[heads, N] | None)``.
``batch_indices`` must be sorted (it always is: ``molecular.py:406-410``); the number of graphs is
``batch_indices.max() + 1`` exactly like torch_scatter's ``dim_size=None`` (``pooling.py:33,56,79,145,159``)
unless a collated :class:`GraphIndex` is supplied, which also removes that device synchronisation.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .collate import GraphIndex
from .layers import INDEX_CACHE, pad_cols


def _segments_for(x: torch.Tensor, batch_indices: Optional[torch.Tensor], graph_index: Optional[GraphIndex]) -> GraphIndex:
    if graph_index is not None:
        return graph_index
    N = x.shape[0]
    if batch_indices is None:                        # pooling.py:146-147 / 162-166: one graph over all nodes
        key = INDEX_CACHE.key(extra=("all", N, str(x.device)))
        return INDEX_CACHE.get(key, lambda: GraphIndex.build(None, torch.zeros(N, dtype=torch.long), 1, 1).to(x.device))
    key = INDEX_CACHE.key(batch_indices, extra=("seg",))

    def build():
        bi = batch_indices.detach().cpu()
        B = int(bi.max()) + 1 if bi.numel() else 0
        return GraphIndex.build(None, bi, B, 1).to(x.device)

    return INDEX_CACHE.get(key, build)


class _SegmentPooling(nn.Module):
    mode = "sum"

    def forward(self, x: torch.Tensor, batch_indices: torch.Tensor,
                graph_index: Optional[GraphIndex] = None) -> Tuple[torch.Tensor, None]:
        gi = _segments_for(x, batch_indices, graph_index)
        F_ = x.shape[1]
        dt = x.dtype
        xp = pad_cols(x, ops.pad_to(F_, 8 if dt != torch.float32 else 4)).contiguous()
        out = ops.cast(ops.SegPoolFn.apply(ops.cast(xp, torch.float32), self.mode, gi), dt)     # bf16 features: fp32 reduction
        return (out[:, :F_] if out.shape[1] != F_ else out), None


class MeanPoolingLayer(_SegmentPooling):
    """Reference ``pooling.py:15-34`` (scatter_mean: sum / max(count, 1))."""
    mode = "mean"


class MaxPoolingLayer(_SegmentPooling):
    """Reference ``pooling.py:37-57`` (scatter_max: empty segment -> 0, gradient to the first maximum)."""
    mode = "max"


class SumPoolingLayer(_SegmentPooling):
    """Reference ``pooling.py:60-80``."""
    mode = "sum"


class MultiHeadAttentionPoolingLayer(nn.Module):
    """Reference ``pooling.py:83-172``."""

    def __init__(self, input_dim: int, num_heads: int = 4, initial_temperature: float = 1.0,
                 learnable_temperature: bool = True, dropout_prob: float = 0.0):
        super().__init__()
        self.num_heads = num_heads
        self.input_dim = input_dim
        self.attention_weights = nn.ModuleList([nn.Linear(input_dim, 1) for _ in range(num_heads)])
        if learnable_temperature:
            self.temperature = nn.Parameter(torch.tensor(initial_temperature))
        else:
            self.register_buffer("temperature", torch.tensor(initial_temperature))
        self.dropout = nn.Dropout(dropout_prob)

    def declare_packed(self, pk, prefix: str) -> None:
        """The stacked head weights [heads, F] / biases [heads] in a ``packed.PackedWeights`` (one gather per forward
        and one gradient collection per backward instead of cat + per-parameter gradient accumulation)."""
        F_ = self.attention_weights[0].in_features
        heads = len(self.attention_weights)
        pk.add(f"{prefix}.W", heads, ops.pad_to(F_, 4), [(l.weight, 0, 1, 0, F_, h, 0) for h, l in enumerate(self.attention_weights)])
        pk.add(f"{prefix}.b", 1, heads, [(l.bias, 0, 1, 0, 1, 0, h) for h, l in enumerate(self.attention_weights)], vector=True)

    def forward(self, x: torch.Tensor, batch_indices: Optional[torch.Tensor],
                graph_index: Optional[GraphIndex] = None, packed=None) -> Tuple[torch.Tensor, torch.Tensor]:
        gi = _segments_for(x, batch_indices, graph_index)
        F_ = x.shape[1]
        Fp = ops.pad_to(F_, 4)
        if packed is not None:
            w, b = packed[0][f"{packed[1]}.W"], packed[0][f"{packed[1]}.b"]
        else:
            w = pad_cols(torch.cat([l.weight for l in self.attention_weights], dim=0), Fp).contiguous()   # [heads, F]
            b = torch.cat([l.bias for l in self.attention_weights], dim=0)                               # [heads]
        # bf16 features: scores, the segmented softmax and the weighted sum run in fp32 on up-cast values
        dt = x.dtype
        if dt != torch.float32 and Fp % 8:
            raise ValueError("bf16 attention pooling needs input_dim % 8 == 0")
        pooled, attn = ops.AttnPoolFn.apply(ops.cast(pad_cols(x, Fp).contiguous(), torch.float32), w, b, self.temperature, gi)
        pooled = ops.cast(pooled, dt)
        if Fp != F_:
            pooled = pooled[:, :F_]
        if self.dropout.p > 0:                                                    # pooling.py:169-170
            pooled = self.dropout(pooled)
        return pooled, attn


def create_pooling_layer(pooling_type: str, input_dim: int, **kwargs) -> nn.Module:
    """Reference ``pooling.py:246-273``.  ``set_attention`` is listed by the reference factory but cannot be
    constructed there either (the factory passes ``num_heads=`` which its constructor rejects)."""
    if pooling_type == "attention":
        return MultiHeadAttentionPoolingLayer(input_dim, **kwargs)
    if pooling_type == "mean":
        return MeanPoolingLayer()
    if pooling_type == "max":
        return MaxPoolingLayer()
    if pooling_type == "sum":
        return SumPoolingLayer()
    if pooling_type == "set_attention":
        raise TypeError("SetAttentionPoolingLayer.__init__() got an unexpected keyword argument 'num_heads'")
    supported = ["attention", "mean", "max", "sum", "set_attention"]
    raise ValueError(f"Unsupported pooling type: {pooling_type}. Supported: {supported}")
