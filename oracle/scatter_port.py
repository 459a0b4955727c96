"""Restatement of ``torch_scatter==2.1.2`` (absent third-party wheel) -- TEST INFRASTRUCTURE.

The reference pins ``torch_scatter==2.1.2+pt25cu121`` (``requirements.txt:6``) and
calls it at ``src/models/layers.py:11,158`` and
``src/models/pooling.py:11,33,56,79,145,159,230,235,241``.  The wheel is not
installed here and cannot be fetched, so this module restates the published
algorithm of the five functions the path uses (torch_scatter 2.1.x,
``torch_scatter/scatter.py`` and ``torch_scatter/composite/softmax.py``):

* ``broadcast``      -- 1-D index is unsqueezed up to ``dim`` and expanded to ``src``.
* ``scatter_sum``    -- ``zeros(dim_size or index.max()+1).scatter_add_(dim, index, src)``.
* ``scatter_mean``   -- sum / clamp(count, min=1) (true division for floats).
* ``scatter_max``    -- (values, argmax); strict ``>`` update so the FIRST occurrence of
  the maximum wins; untouched segments are filled with 0 and ``arg = src.size(dim)``;
  backward routes the gradient to ``arg`` only.
* ``scatter_softmax``-- ``m = scatter_max``; ``e = exp(src - m[index])``;
  ``e / scatter_sum(e)[index]`` (no epsilon).

It doubles as the ``sys.modules['torch_scatter']`` shim that lets the UNMODIFIED
reference modules run in the build container (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import torch


def broadcast(index: torch.Tensor, other: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim = other.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), other.dim()):
        index = index.unsqueeze(-1)
    return index.expand(other.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count = count.clamp(min=1)
    count = broadcast(count, out, dim)
    if out.is_floating_point():
        return out / count
    return out.div(count, rounding_mode="floor")


class _ScatterMax(torch.autograd.Function):
    """values/argmax with torch_scatter's first-occurrence tie rule and arg-only backward."""

    @staticmethod
    def forward(ctx, src, index, dim, dim_size):
        dim = dim if dim >= 0 else dim + src.dim()
        L = src.size(dim)
        s2 = src.movedim(dim, 0).reshape(L, -1)
        i2 = index.movedim(dim, 0).reshape(L, -1)
        M = s2.size(1)
        lowest = torch.finfo(src.dtype).min if src.is_floating_point() else torch.iinfo(src.dtype).min
        vals = torch.full((dim_size, M), lowest, dtype=src.dtype)
        vals.scatter_reduce_(0, i2, s2, "amax", include_self=True)
        pos = torch.arange(L).unsqueeze(1).expand(L, M)
        cand = torch.where(s2 == vals.gather(0, i2), pos, torch.full_like(pos, L))
        arg = torch.full((dim_size, M), L, dtype=torch.long)
        arg.scatter_reduce_(0, i2, cand, "amin", include_self=True)
        vals = vals.masked_fill(arg == L, 0)
        out_shape = list(src.movedim(dim, 0).shape)
        out_shape[0] = dim_size
        vals = vals.reshape(out_shape).movedim(0, dim).contiguous()
        arg = arg.reshape(out_shape).movedim(0, dim).contiguous()
        ctx.save_for_backward(arg)
        ctx.dim = dim
        ctx.src_shape = list(src.shape)
        ctx.mark_non_differentiable(arg)
        return vals, arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        shape = list(ctx.src_shape)
        shape[ctx.dim] += 1
        grad_in = torch.zeros(shape, dtype=grad_out.dtype)
        grad_in.scatter_(ctx.dim, arg, grad_out)
        return grad_in.narrow(ctx.dim, 0, shape[ctx.dim] - 1), None, None, None


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert out is None, "out= is not used on the path"
    index = broadcast(index, src, dim)
    if dim_size is None:
        dim_size = 0 if index.numel() == 0 else int(index.max()) + 1
    return _ScatterMax.apply(src, index, dim, dim_size)


def scatter_softmax(src, index, dim=-1, dim_size=None):
    index = broadcast(index, src, dim)
    max_value_per_index = scatter_max(src, index, dim=dim, dim_size=dim_size)[0]
    max_per_src_element = max_value_per_index.gather(dim, index)
    recentered = src - max_per_src_element
    recentered_exp = recentered.exp()
    sum_per_index = scatter_sum(recentered_exp, index, dim, dim_size=dim_size)
    return recentered_exp / sum_per_index.gather(dim, index)
