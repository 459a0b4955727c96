"""B200-native ``ShellConvolutionLayer`` / ``LinearBlock`` / ``MultiLayerPerceptron``.

Same constructor arguments, forward signatures, attribute names and ``state_dict`` keys/shapes as the
reference ``src/models/layers.py`` (17-267); the computation is the fused CUDA path of ``ops.py``.
Padded / packed weights are derived from the parameters -- per call with torch ops for stand-alone modules, by
one kernel per forward inside ``GNN`` (``packed.PackedWeights``) -- and are never part of the ``state_dict``, so
checkpoints round-trip with the reference.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .activation import get_activation_function
from .collate import GraphIndex

FEATURE_PAD = 32      # node-feature widths are padded to 128-byte rows (SURVEY.md hard part 1: D = 153 -> 160)


def pad_cols(t: torch.Tensor, width: int) -> torch.Tensor:
    return t if t.shape[-1] == width else F.pad(t, (0, width - t.shape[-1]))


def pad2d(w: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    if w.shape[0] == rows and w.shape[1] == cols:
        return w
    return F.pad(w, (0, cols - w.shape[1], 0, rows - w.shape[0]))


def pad1d(b: Optional[torch.Tensor], n: int) -> Optional[torch.Tensor]:
    if b is None or b.shape[0] == n:
        return b
    return F.pad(b, (0, n - b.shape[0]))


def pack_chunked(w: torch.Tensor, rows_pad: int, n_chunks: int, chunk: int, chunk_pad: int, used: int) -> torch.Tensor:
    """[R, n_chunks*chunk] -> [rows_pad, used*chunk_pad]: every column chunk padded on its own, first `used` kept."""
    R = w.shape[0]
    v = w.reshape(R, n_chunks, chunk)[:, :used]
    v = F.pad(v, (0, chunk_pad - chunk, 0, 0, 0, rows_pad - R))
    return v.reshape(rows_pad, used * chunk_pad)


class _IndexCache:
    """Small LRU of GraphIndex objects rebuilt on the fly when the caller passes raw index tensors.

    An entry is keyed on the IDENTITY of the index tensors (the Python objects, or the base tensor of a view plus the
    view's geometry) and their autograd version counters, and it holds weak references to those objects: it is only
    served while every one of them is still alive, is the same object and has not been written in place.  Addresses are
    never part of the key -- the caching allocator hands a freed block to the next batch, so a different batch of the same
    shape routinely reappears at the same ``data_ptr``."""

    def __init__(self, size: int = 8):
        self.size = size
        self.items: "OrderedDict[tuple, tuple]" = OrderedDict()

    @staticmethod
    def _root(t):
        base = t._base if t._base is not None else t
        return base

    @staticmethod
    def key(*tensors, extra=()):
        k = []
        for t in tensors:
            if t is None:
                k.append(None)
                continue
            k.append((id(_IndexCache._root(t)), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t._version,
                      str(t.device), str(t.dtype)))
        return (tuple(k) + tuple(extra), tuple(tensors))

    def get(self, key, builder):
        key, tensors = key
        hit = self.items.get(key)
        if hit is not None:
            gi, refs = hit
            roots = [None if t is None else self._root(t) for t in tensors]
            if len(refs) == len(roots) and all((r is None and o is None) or (r is not None and r() is o)
                                               for r, o in zip(refs, roots)):
                self.items.move_to_end(key)
                return gi
            del self.items[key]            # id() of a dead object was reused by another tensor
        gi = builder()
        refs = [None if t is None else weakref.ref(self._root(t)) for t in tensors]
        self.items[key] = (gi, refs)
        while len(self.items) > self.size:
            self.items.popitem(last=False)
        # drop entries whose tensors died (their ids may be reused)
        for k in [k for k, (_, rs) in self.items.items() if any(r is not None and r() is None for r in rs)]:
            del self.items[k]
        return gi


INDEX_CACHE = _IndexCache()


class DropClock(nn.Module):
    """Device-side tick for the counter-based dropout of the GEMM epilogues (CUDA-graph friendly).

    Every dropout site has its own seed; the tick only has to change from one forward pass to the next.  Inside
    ``GNN.forward`` all sites therefore share ONE tick per pass (``shared_tick``: one add + one snapshot instead of
    one pair per layer and block); a stand-alone module advances its own."""

    _shared = None          # snapshot of the enclosing model's tick for the running forward pass

    def __init__(self):
        super().__init__()
        self.register_buffer("tick", torch.zeros(1, dtype=torch.int64), persistent=False)
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item())

    def advance(self) -> torch.Tensor:
        if DropClock._shared is not None and DropClock._shared.device == self.tick.device:
            return DropClock._shared
        self.tick.add_(1)
        return self.tick.clone()      # snapshot: backward must see the value of ITS forward


class shared_tick:
    """``with shared_tick(clock):`` -- the dropout sites below use one tick of ``clock`` for this forward pass."""

    def __init__(self, clock: DropClock):
        self.clock = clock

    def __enter__(self):
        self.prev = DropClock._shared
        DropClock._shared = None
        DropClock._shared = self.clock.advance()
        return self

    def __exit__(self, *exc):
        DropClock._shared = self.prev
        return False


class ShellConvolutionLayer(nn.Module):
    """Shell-based graph convolution (reference ``layers.py:17-167``).

    ``forward(x[N, D], target[E], src[E]) -> [N, output_dim]`` with ``target in [0, num_hops * N)`` and ``src``
    taken modulo N (``layers.py:139-163``).  Under the shipped collation every ``target < N`` (quirk Q1), in
    which case only the first two column chunks of ``input_proj`` / ``global_skip_proj`` meet non-zero data
    and the kernel skips the rest (bit-identical: the skipped products are exact zeros).
    """

    def __init__(self, atom_input_dim: int, output_dim: int, num_hops: int = 3, dropout: float = 0.00,
                 activation_type: str = "silu", num_mlp_layers: int = 2):
        super().__init__()
        self.num_hops = num_hops
        self.atom_input_dim = atom_input_dim
        self.output_dim = output_dim
        self.activation_type = activation_type
        input_dim = atom_input_dim * (num_hops + 1)
        self.activation = get_activation_function(activation_type)
        self.input_proj = nn.Linear(input_dim, output_dim)
        self.mlp_blocks = nn.ModuleList()
        for _ in range(num_mlp_layers):
            self.mlp_blocks.append(nn.ModuleDict({
                "linear_1": nn.Linear(output_dim, output_dim),
                "activation": get_activation_function(activation_type),
                "dropout": nn.Dropout(dropout),
                "linear_2": nn.Linear(output_dim, output_dim),
            }))
        self.global_skip_proj = nn.Linear(input_dim, output_dim) if input_dim != output_dim else None
        self._clock = DropClock()
        self._seeds = [int(s) for s in torch.randint(0, 2 ** 62, (max(num_mlp_layers, 1),))]

    # ------------------------------------------------------------------ reference-facing API
    def forward(self, x: torch.Tensor, target: torch.Tensor, src: torch.Tensor,
                graph_index: Optional[GraphIndex] = None) -> torch.Tensor:
        N = x.shape[0]
        if graph_index is None:
            key = INDEX_CACHE.key(target, src, extra=(N, self.num_hops))
            graph_index = INDEX_CACHE.get(key, lambda: GraphIndex.build(
                torch.stack([target, src], dim=1), torch.zeros(N, dtype=torch.long), 1, self.num_hops).to(x.device))
        xp = pad_cols(x, ops.pad_to(self.atom_input_dim, FEATURE_PAD)).contiguous()
        out = self.forward_padded(xp, graph_index, add_input=False)
        return out[:, : self.output_dim]

    def message_passing(self, atom_features: torch.Tensor, target: torch.Tensor, src: torch.Tensor) -> List[torch.Tensor]:
        """Reference ``layers.py:133-167``: list of ``num_hops`` tensors ``[N, D]`` (no autograd through this helper)."""
        N, D = atom_features.shape
        if target.numel() == 0:
            return [torch.zeros_like(atom_features) for _ in range(self.num_hops)]
        gi = GraphIndex.build(torch.stack([target, src], dim=1), torch.zeros(N, dtype=torch.long), 1,
                              self.num_hops).to(atom_features.device)
        xp = pad_cols(atom_features.detach(), ops.pad_to(D, 4)).contiguous()
        ag = ops.agg(xp, gi)[:, :D]
        chunks = list(torch.split(ag, N, dim=0))
        while len(chunks) < self.num_hops:           # collapsed index: hops 2..H are exact zeros (quirk Q1)
            chunks.append(torch.zeros_like(atom_features))
        return chunks

    # ------------------------------------------------------------------ fused path on padded features
    def declare_packed(self, pk, prefix: str, collapsed: bool) -> None:
        """Declare this layer's packed operands in a ``packed.PackedWeights``: W_io = [W_in ; W_skip] over the column
        chunks that meet non-zero data, b_io, and the padded MLP weights."""
        Din, Dout, H = self.atom_input_dim, self.output_dim, self.num_hops
        Di, Do = ops.pad_to(Din, FEATURE_PAD), ops.pad_to(Dout, FEATURE_PAD)
        used = 1 + (1 if collapsed else H)
        projs = (self.input_proj, self.global_skip_proj)
        pk.add(f"{prefix}.W_io", 2 * Do, used * Di,
               [(pr.weight, 0, Dout, c * Din, (c + 1) * Din, j * Do, c * Di) for j, pr in enumerate(projs) for c in range(used)])
        pk.add(f"{prefix}.b_io", 1, 2 * Do, [(pr.bias, 0, 1, 0, Dout, 0, j * Do) for j, pr in enumerate(projs)], vector=True)
        for k, blk in enumerate(self.mlp_blocks):
            for name in ("linear_1", "linear_2"):
                pk.add(f"{prefix}.{k}.{name}.W", Do, Do, [(blk[name].weight, 0, Dout, 0, Dout, 0, 0)])
                pk.add(f"{prefix}.{k}.{name}.b", 1, Do, [(blk[name].bias, 0, 1, 0, Dout, 0, 0)], vector=True)

    def forward_padded(self, xp: torch.Tensor, gi: GraphIndex, add_input: bool, packed=None) -> torch.Tensor:
        """``packed = (PackedWeights, prefix)``: read the operands declared by ``declare_packed`` instead of deriving
        them from the parameters with torch ops."""
        if self.global_skip_proj is None:
            raise NotImplementedError("ShellConvolutionLayer without global_skip_proj "
                                      "(atom_input_dim * (num_hops + 1) == output_dim) is not supported")
        Din, Dout, H = self.atom_input_dim, self.output_dim, self.num_hops
        Di, Do = ops.pad_to(Din, FEATURE_PAD), ops.pad_to(Dout, FEATURE_PAD)
        used = 1 + (1 if gi.collapsed else H)
        if packed is not None:
            pk, prefix = packed
            w_io, b_io = pk[f"{prefix}.W_io"], pk[f"{prefix}.b_io"]
            if w_io.shape[1] != used * Di:
                raise RuntimeError("packed shell weights were declared for a different hop layout")
        else:
            w_io = torch.cat([pack_chunked(self.input_proj.weight, Do, H + 1, Din, Di, used),
                              pack_chunked(self.global_skip_proj.weight, Do, H + 1, Din, Di, used)], dim=0)
            b_io = torch.cat([pad1d(self.input_proj.bias, Do), pad1d(self.global_skip_proj.bias, Do)], dim=0)
        ps = []
        mlp = []
        for k, blk in enumerate(self.mlp_blocks):
            d = blk["dropout"]            # each nn.Dropout keeps its own .training flag (MC-dropout flips only those)
            ps.append(float(d.p) if (d.training and d.p > 0) else 0.0)
            if packed is not None:
                mlp += [pk[f"{prefix}.{k}.linear_1.W"], pk[f"{prefix}.{k}.linear_1.b"],
                        pk[f"{prefix}.{k}.linear_2.W"], pk[f"{prefix}.{k}.linear_2.b"]]
                continue
            mlp += [pad2d(blk["linear_1"].weight, Do, Do), pad1d(blk["linear_1"].bias, Do),
                    pad2d(blk["linear_2"].weight, Do, Do), pad1d(blk["linear_2"].bias, Do)]
        tick = self._clock.advance() if any(p > 0 for p in ps) else None
        opts = ops._Opts(act=self.activation_type, ps=ps, seeds=self._seeds, tick=tick, w_in=Di, w_out=Do,
                         n_mlp=len(self.mlp_blocks), add_input=add_input and Di == Do, gi=gi)
        out = ops.ShellConvFn.apply(opts, xp, w_io, b_io, *mlp)
        if add_input and Di != Do:
            raise RuntimeError("residual connection needs atom_input_dim == output_dim")
        return out


class _FusedLinear(nn.Linear):
    """``nn.Linear`` (same parameters / state_dict) whose forward is the libax2d GEMM.  It also accepts a
    list of already padded column segments so that the reference's ``torch.cat`` is never materialised."""

    def declare_packed(self, pk, prefix: str, true_widths, widths, n_out=None) -> None:
        """Packed operand for a call with column segments of ``true_widths`` padded to ``widths``."""
        n_out = ops.pad_to(self.out_features, 4) if n_out is None else n_out
        blocks, off, dst = [], 0, 0
        for tw, w in zip(true_widths, widths):
            blocks.append((self.weight, 0, self.out_features, off, off + tw, 0, dst))
            off += tw
            dst += w
        if off != self.in_features:
            raise ValueError(f"{prefix}: segments cover {off} of {self.in_features} input features")
        pk.add(f"{prefix}.W", n_out, dst, blocks)
        if self.bias is not None:
            pk.add(f"{prefix}.b", 1, n_out, [(self.bias, 0, 1, 0, self.out_features, 0, 0)], vector=True)

    def forward(self, x, packed=None, out_dtype=None):  # type: ignore[override]
        if isinstance(x, (list, tuple)):
            segs, true_widths = x[0], x[1]   # ([tensors], [true widths][, padded out width]) -- internal convention
            widths = [s.shape[1] for s in segs]
            n_out = x[2] if len(x) > 2 else ops.pad_to(self.out_features, 4)
            if packed is not None:
                pk, prefix = packed
                W, b = pk[f"{prefix}.W"], (pk[f"{prefix}.b"] if self.bias is not None else None)
                if tuple(W.shape) != (n_out, sum(widths)):
                    raise RuntimeError(f"{prefix}: packed weight {tuple(W.shape)} vs call ({n_out}, {sum(widths)})")
            else:
                cols = []
                off = 0
                for tw, w in zip(true_widths, widths):
                    cols.append(pad2d(self.weight[:, off:off + tw], n_out, w))
                    off += tw
                W = cols[0] if len(cols) == 1 else torch.cat(cols, dim=1)
                b = pad1d(self.bias, n_out)
            out = ops.LinearFn.apply(ops._Opts(widths=widths, n_out=n_out, out_dtype=out_dtype), W, b, *segs)
            return out if (n_out == self.out_features or len(x) > 2) else out[:, : self.out_features]
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        k = ops.pad_to(self.in_features, 4)
        out = self.forward(([pad_cols(x2, k).contiguous()], [self.in_features]), packed=packed)
        return out.reshape(*lead, self.out_features)


class LinearBlock(nn.Module):
    """Reference ``layers.py:170-219``: linear1 -> act -> dropout -> linear2 (+ identity skip)."""

    def __init__(self, input_dim: int, output_dim: int, activation_type: str = "silu", dropout: float = 0.0,
                 use_skip: bool = True):
        super().__init__()
        self.use_skip = use_skip and (input_dim == output_dim)
        self.activation_type = activation_type
        self.linear1 = nn.Linear(input_dim, output_dim)
        self.activation = get_activation_function(activation_type)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(output_dim, output_dim)
        self.skip_proj = None                      # unreachable in the reference as well (layers.py:190,198)
        self._clock = DropClock()

    def declare_packed(self, pk, prefix: str) -> None:
        din, dout = self.linear1.in_features, self.linear1.out_features
        wi, wo = ops.pad_to(din, 4), ops.pad_to(dout, 4)
        pk.add(f"{prefix}.W1", wo, wi, [(self.linear1.weight, 0, dout, 0, din, 0, 0)])
        pk.add(f"{prefix}.b1", 1, wo, [(self.linear1.bias, 0, 1, 0, dout, 0, 0)], vector=True)
        pk.add(f"{prefix}.W2", wo, wo, [(self.linear2.weight, 0, dout, 0, dout, 0, 0)])
        pk.add(f"{prefix}.b2", 1, wo, [(self.linear2.bias, 0, 1, 0, dout, 0, 0)], vector=True)

    def forward(self, x: torch.Tensor, packed=None) -> torch.Tensor:
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        din, dout = self.linear1.in_features, self.linear1.out_features
        wi, wo = ops.pad_to(din, 4), ops.pad_to(dout, 4)
        p = float(self.dropout.p) if (self.dropout.training and self.dropout.p > 0) else 0.0
        tick = self._clock.advance() if p > 0 else None
        opts = ops._Opts(act=self.activation_type, p=p, seed=self._clock.seed, tick=tick, skip=self.use_skip,
                         w_in=wi, w_out=wo)
        if packed is not None:
            pk, prefix = packed
            ws = [pk[f"{prefix}.{n}"] for n in ("W1", "b1", "W2", "b2")]
        else:
            ws = [pad2d(self.linear1.weight, wo, wi), pad1d(self.linear1.bias, wo), pad2d(self.linear2.weight, wo, wo),
                  pad1d(self.linear2.bias, wo)]
        out = ops.MLPBlockFn.apply(opts, pad_cols(x2, wi).contiguous(), *ws)
        return out[:, :dout].reshape(*lead, dout) if wo != dout else out.reshape(*lead, dout)


class MultiLayerPerceptron(nn.Module):
    """Reference ``layers.py:222-267``."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int = 2,
                 activation_type: str = "silu", dropout: float = 0.0, use_skip: bool = True):
        super().__init__()
        layers = []
        if num_layers == 1:
            layers.append(LinearBlock(input_dim, output_dim, activation_type, dropout, False))
        else:
            layers.append(LinearBlock(input_dim, hidden_dim, activation_type, dropout, False))
            for _ in range(num_layers - 2):
                layers.append(LinearBlock(hidden_dim, hidden_dim, activation_type, dropout, use_skip))
            layers.append(LinearBlock(hidden_dim, output_dim, activation_type, dropout, False))
        self.layers = nn.ModuleList(layers)

    def declare_packed(self, pk, prefix: str) -> None:
        for i, layer in enumerate(self.layers):
            layer.declare_packed(pk, f"{prefix}.{i}")

    def forward(self, x: torch.Tensor, packed=None) -> torch.Tensor:
        for i, layer in enumerate(self.layers):
            x = layer(x) if packed is None else layer(x, packed=(packed[0], f"{packed[1]}.{i}"))
        return x
