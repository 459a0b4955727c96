"""B200-native ``GNN`` -- drop-in for the reference ``src/models/gnn.py`` (19-779).

Identical constructor arguments, ``forward`` signature and return tuple, attribute names used by the
reference's callers (``pooling``, ``concat_self_other``, ``message_passing_layers``, ``output_layer``,
``loss_function``, ``task_type``, ``hidden_dim``, ``num_shells``, ``embedding_dim``, ``init_weights``,
``get_model_info``) and ``state_dict`` keys/shapes, so reference checkpoints load unchanged.

The forward wiring keeps node features in a padded layout (x_other: D -> multiple of 32 columns) between
the fused kernels; ``torch.cat`` calls of the reference become multi-segment GEMM operands.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .activation import get_activation_function
from .collate import GraphIndex
from .layers import (FEATURE_PAD, INDEX_CACHE, DropClock, MultiLayerPerceptron, ShellConvolutionLayer, _FusedLinear,
                     pad1d, pad2d, shared_tick)
from .packed import PackAnchorFn, PackedWeights
from .pooling import create_pooling_layer

FEATURE_ORDER = ("atom_type", "hydrogen_count", "degree", "hybridization")      # gnn.py:262-274


class GNN(nn.Module):
    def __init__(self, feature_sizes: Dict[str, int], hidden_dim: int, output_dim: int, num_shells: int = 3,
                 num_message_passing_layers: int = 3, dropout: float = 0.05, ffn_hidden_dim: Optional[int] = None,
                 ffn_num_layers: int = 3, pooling_type: str = "attention", task_type: str = "regression",
                 embedding_dim: int = 64, use_partial_charges: bool = False, use_stereochemistry: bool = False,
                 ffn_dropout: float = 0.05, activation_type: str = "silu", shell_conv_num_mlp_layers: int = 2,
                 shell_conv_dropout: float = 0.05, attention_num_heads: int = 4, attention_temperature: float = 1.0,
                 loss_function: str = "l1", verbose: bool = False):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.num_shells = num_shells
        self.task_type = task_type
        self.embedding_dim = embedding_dim
        self.use_partial_charges = use_partial_charges
        self.use_stereochemistry = use_stereochemistry
        self.loss_function = loss_function
        self.activation_type = activation_type
        self._verbose = verbose
        if ffn_hidden_dim is None:
            ffn_hidden_dim = hidden_dim

        # gnn.py:151-171
        self.atom_type_embedding = nn.Embedding(feature_sizes["atom_type"], embedding_dim)
        self.hydrogen_count_embedding = nn.Embedding(feature_sizes["hydrogen_count"], embedding_dim)
        self.degree_embedding = nn.Embedding(feature_sizes["degree"], embedding_dim)
        self.hybridization_embedding = nn.Embedding(feature_sizes["hybridization"], embedding_dim)

        self.embedding_projection = nn.Linear(embedding_dim * len(feature_sizes), hidden_dim)   # gnn.py:95-96
        self.activation = get_activation_function(activation_type)
        self.x_other_dim = int(0.3 * hidden_dim)                                               # gnn.py:100-101
        self.x_self_dim = hidden_dim - self.x_other_dim

        self.message_passing_layers = nn.ModuleList([                                           # gnn.py:173-186
            ShellConvolutionLayer(atom_input_dim=self.x_other_dim, output_dim=self.x_other_dim, num_hops=num_shells,
                                  activation_type=activation_type, dropout=shell_conv_dropout,
                                  num_mlp_layers=shell_conv_num_mlp_layers)
            for _ in range(num_message_passing_layers)])

        self.pooling = create_pooling_layer(pooling_type, hidden_dim, num_heads=attention_num_heads,
                                            initial_temperature=attention_temperature)          # gnn.py:110-115

        self.concat_self_other = _FusedLinear(hidden_dim, hidden_dim)                           # gnn.py:190
        if self.use_stereochemistry:                                                            # gnn.py:192-195
            self.stereochemical_embedding = nn.Linear(hidden_dim * 3, hidden_dim)               # dead parameter (Q6)
            self.stereochemical_embedding_2 = _FusedLinear(self.x_other_dim * 3, self.x_other_dim)

        self.post_pooling_projection = _FusedLinear(hidden_dim, ffn_hidden_dim)                 # gnn.py:121
        self.ffn = MultiLayerPerceptron(input_dim=ffn_hidden_dim, hidden_dim=ffn_hidden_dim, output_dim=ffn_hidden_dim,
                                        num_layers=ffn_num_layers, activation_type=activation_type,
                                        dropout=ffn_dropout, use_skip=True)
        self.skip_transform = _FusedLinear(ffn_hidden_dim, ffn_hidden_dim)                      # gnn.py:133
        final_output_dim = output_dim * 4 if loss_function == "evidential" else output_dim      # gnn.py:136-141
        self.output_layer = _FusedLinear(ffn_hidden_dim * 2, final_output_dim)                  # gnn.py:143
        self.long_range_projection = nn.Linear(hidden_dim, ffn_hidden_dim)                      # dead parameter (Q6)
        self._clock = DropClock()           # one dropout tick per forward pass for all layers (layers.shared_tick)
        self.use_packed_weights = True      # one pack kernel per forward instead of per-call torch pad / cat / split
        # torch.bfloat16: activations / saved tensors / packed weight copies in bf16, fp32 accumulation, fp32 master weights
        # and gradients (BASELINE configs[3]).  None: follow torch.autocast (the reference's mixed precision,
        # training/trainer.py:133-149) -- bf16 inside ``torch.autocast("cuda", dtype=torch.bfloat16)``, fp32 otherwise.
        self.compute_dtype = None
        self._packed: Dict[tuple, PackedWeights] = {}
        self.init_weights()

    # ------------------------------------------------------------------------------------------ forward
    def _index_for(self, atom_features, edges, batch_indices, total_charges, tetra, cis, trans) -> GraphIndex:
        dev = batch_indices.device
        key = INDEX_CACHE.key(edges, batch_indices, tetra, cis, trans, *[atom_features[k] for k in FEATURE_ORDER],
                              extra=(self.num_shells,))

        def build():
            sizes = {"atom_type": self.atom_type_embedding.num_embeddings,
                     "hydrogen_count": self.hydrogen_count_embedding.num_embeddings,
                     "degree": self.degree_embedding.num_embeddings,
                     "hybridization": self.hybridization_embedding.num_embeddings}
            return GraphIndex.build(edges, batch_indices, int(total_charges.shape[0]), self.num_shells,
                                    {k: atom_features[k] for k in FEATURE_ORDER}, sizes, tetra, cis, trans).to(dev)

        return INDEX_CACHE.get(key, build)

    def _packed_weights(self, gi: GraphIndex, bf16: bool = False) -> Optional[PackedWeights]:
        """The packed operands of every projection of this model for the hop layout of ``gi`` (built once per device
        and layout), or None when the per-call path has to be used: embedding widths that need padding, parameters
        that are not all trainable fp32 CUDA tensors (the gradients are collected behind the first embedding table)."""
        if not self.use_packed_weights or self.embedding_dim % 4 != 0:
            return None
        t0 = self.atom_type_embedding.weight
        if not t0.is_cuda or t0.dtype != torch.float32 or (torch.is_grad_enabled() and not t0.requires_grad):
            return None
        key = (str(t0.device), bool(gi.collapsed), bool(bf16))
        pk = self._packed.get(key)
        if pk is not None:
            return pk
        if any(p.dtype != torch.float32 or p.device != t0.device for p in self.parameters()):
            return None
        pk = self._declare_packed(t0.device, bool(gi.collapsed), bf16)
        if pk is not None:
            self._packed[key] = pk
        return pk

    def _bf16_active(self) -> bool:
        if self.compute_dtype is not None:
            if self.compute_dtype not in (torch.float32, torch.bfloat16):
                raise ValueError("compute_dtype must be torch.float32, torch.bfloat16 or None (follow autocast)")
            return self.compute_dtype == torch.bfloat16
        if torch.is_autocast_enabled():
            dt = torch.get_autocast_gpu_dtype()
            if dt != torch.bfloat16:
                raise RuntimeError(f"autocast dtype {dt} is not supported by the B200 path: use torch.bfloat16 (fp32 "
                                   "accumulation, no GradScaler needed) or disable autocast")
            return True
        return False

    def _declare_packed(self, device, collapsed: bool, bf16: bool = False) -> Optional[PackedWeights]:
        D, S = self.x_other_dim, self.x_self_dim
        Dp, Sp = ops.pad_to(D, FEATURE_PAD), ops.pad_to(S, FEATURE_PAD)
        ntE = self.embedding_dim * len(FEATURE_ORDER)
        H4 = ops.pad_to(self.hidden_dim, 4)
        pk = PackedWeights(device)
        Wp, bp = self.embedding_projection.weight, self.embedding_projection.bias
        pk.add("ep.W", Sp + Dp, ntE, [(Wp, 0, S, 0, ntE, 0, 0), (Wp, S, S + D, 0, ntE, Sp, 0)])
        pk.add("ep.b", 1, Sp + Dp, [(bp, 0, 1, 0, S, 0, 0), (bp, 0, 1, S, S + D, 0, Sp)], vector=True)
        for i, layer in enumerate(self.message_passing_layers):
            if layer.global_skip_proj is None:
                return None
            layer.declare_packed(pk, f"mp.{i}", collapsed)
        self.concat_self_other.declare_packed(pk, "cso", [S, D], [Sp, Dp])
        if self.use_stereochemistry:
            self.stereochemical_embedding_2.declare_packed(pk, "stereo2", [D, D, D], [Dp, Dp, Dp], n_out=Dp)
        F4 = ops.pad_to(self.post_pooling_projection.out_features, 4)
        self.post_pooling_projection.declare_packed(pk, "ppp", [self.hidden_dim], [H4])
        self.ffn.declare_packed(pk, "ffn")
        if F4 == self.post_pooling_projection.out_features:      # the segments of the output layer are unpadded
            self.skip_transform.declare_packed(pk, "skip", [F4], [F4])
            # bf16: the output width is padded to 32 (tensor-core operands of the backward products need 16-byte rows)
            self.output_layer.declare_packed(pk, "out", [F4, F4], [F4, F4],
                                             n_out=ops.pad_to(self.output_layer.out_features, 32) if bf16 else None)
        if hasattr(self.pooling, "declare_packed"):
            self.pooling.declare_packed(pk, "pool")
        pk.finalize()
        if bf16:
            pk.enable_bf16()
        return pk

    def forward(self, atom_features: Dict[str, torch.Tensor], multi_hop_edge_indices: torch.Tensor,
                batch_indices: torch.Tensor, total_charges: torch.Tensor, tetrahedral_indices: torch.Tensor,
                cis_indices: torch.Tensor, trans_indices: torch.Tensor, graph_index: Optional[GraphIndex] = None
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Reference ``gnn.py:197-260``.  ``graph_index`` (optional, from ``MolBatch``) carries the CSR / segment
        offsets emitted at collation; when absent it is rebuilt from the index tensors (and cached)."""
        dropping = any(m.training and m.p > 0 for m in self.modules() if isinstance(m, nn.Dropout))
        if not dropping:
            return self._forward(atom_features, multi_hop_edge_indices, batch_indices, total_charges,
                                 tetrahedral_indices, cis_indices, trans_indices, graph_index)
        with shared_tick(self._clock):
            return self._forward(atom_features, multi_hop_edge_indices, batch_indices, total_charges,
                                 tetrahedral_indices, cis_indices, trans_indices, graph_index)

    def _forward(self, atom_features, multi_hop_edge_indices, batch_indices, total_charges, tetrahedral_indices,
                 cis_indices, trans_indices, graph_index=None):
        gi = graph_index
        if gi is None:
            gi = self._index_for(atom_features, multi_hop_edge_indices, batch_indices, total_charges,
                                 tetrahedral_indices, cis_indices, trans_indices)
        bf16 = self._bf16_active()
        dt = torch.bfloat16 if bf16 else torch.float32
        if bf16 and (self.hidden_dim % 32 or self.post_pooling_projection.out_features % 32 or (self.embedding_dim * 4) % 32):
            raise ValueError("the bf16 configuration needs hidden_dim, ffn_hidden_dim and 4 * embedding_dim to be multiples "
                             "of 32 (16-byte rows for the tensor-core operands)")
        D, S = self.x_other_dim, self.x_self_dim
        Dp, Sp = ops.pad_to(D, FEATURE_PAD), ops.pad_to(S, FEATURE_PAD)
        E = self.embedding_dim
        Ep = ops.pad_to(E, 4)
        tables = [getattr(self, f"{k}_embedding").weight for k in FEATURE_ORDER]
        if Ep != E:
            tables = [pad2d(t, t.shape[0], Ep) for t in tables]
        nt = len(FEATURE_ORDER)
        pk = self._packed_weights(gi, bf16)
        if bf16 and pk is None:
            raise RuntimeError("the bf16 configuration needs the packed-weight path (trainable fp32 CUDA parameters, "
                               "embedding_dim % 4 == 0, use_packed_weights = True)")
        use = (lambda name: (pk, name) if f"{name}.W" in pk or f"{name}.W_io" in pk or f"{name}.0.W1" in pk else None) \
            if pk is not None else (lambda name: None)
        if pk is not None:
            # refreshes the packed operands now; its backward (the last node of the graph) collects their gradients
            tables[0] = PackAnchorFn.apply(pk, tables[0])
            W_ep, b_ep = pk["ep.W"], pk["ep.b"]
        else:
            Wp = self.embedding_projection.weight
            if Ep != E:
                Wp = torch.cat([pad2d(Wp[:, i * E:(i + 1) * E], Wp.shape[0], Ep) for i in range(nt)], dim=1)
            W_ep = torch.cat([pad2d(Wp[:S], Sp, nt * Ep), pad2d(Wp[S:], Dp, nt * Ep)], dim=0)
            bp = self.embedding_projection.bias
            b_ep = torch.cat([pad1d(bp[:S], Sp), pad1d(bp[S:], Dp)], dim=0)
        opts = ops._Opts(act=self.activation_type, names=FEATURE_ORDER, emb_dim=Ep, s_pad=Sp, d_pad=Dp, gi=gi, dtype=dt)
        x_self, x = ops.EmbedProjFn.apply(opts, W_ep, b_ep, *tables,
                                          *[atom_features[k].contiguous() for k in FEATURE_ORDER])   # gnn.py:220-231

        if multi_hop_edge_indices.numel() > 0:                                                     # gnn.py:287
            for i, layer in enumerate(self.message_passing_layers):
                if self.use_partial_charges:                                                       # gnn.py:290-293
                    # bf16: fp32 equilibration on up-cast values (the reference crashes here under autocast, quirk Q5)
                    x = ops.cast(ops.ChargeEqFn.apply(ops.cast(x, torch.float32), total_charges, gi), dt)
                if self.use_stereochemistry:                                                       # gnn.py:296-299
                    x = self._apply_stereochemistry(x, gi, D, Dp, use("stereo2"))
                x = layer.forward_padded(x, gi, add_input=True, packed=use(f"mp.{i}"))             # gnn.py:302-306

        partial_charges = None
        if self.use_partial_charges and D >= 2:                                                    # gnn.py:240-242
            partial_charges = x[:, 0].float().clone()

        atom_emb = self.concat_self_other(([x_self, x], [S, D]), packed=use("cso"))                # gnn.py:245-246
        if pk is not None and "pool.W" in pk:
            x_pooled, attention_weights = self.pooling(atom_emb, batch_indices, graph_index=gi, packed=(pk, "pool"))
        else:
            x_pooled, attention_weights = self.pooling(atom_emb, batch_indices, graph_index=gi)    # gnn.py:249
        v = self.post_pooling_projection(x_pooled, packed=use("ppp"))                              # gnn.py:252
        v = self.ffn(v, packed=use("ffn"))                                                         # gnn.py:253
        skip = self.skip_transform(v, packed=use("skip"))                                          # gnn.py:256
        Fh = v.shape[1]
        if bf16:        # fp32 outputs (they feed the loss), computed 32 columns wide and cut back to the real targets
            T = self.output_layer.out_features
            output = self.output_layer(([v, skip], [Fh, Fh], ops.pad_to(T, 32)), packed=use("out"),
                                       out_dtype=torch.float32)[:, :T]
        else:
            output = self.output_layer(([v, skip], [Fh, Fh]), packed=use("out"))                   # gnn.py:257-258
        return output, attention_weights, partial_charges

    def _apply_stereochemistry(self, x: torch.Tensor, gi: GraphIndex, D: int, Dp: int, packed=None) -> torch.Tensor:
        """Reference ``gnn.py:310-327``: Linear([x | cis_trans(x) | tetra(x)])."""
        dt = x.dtype
        x32 = ops.cast(x, torch.float32) if (gi.cistrans is not None or gi.tetra is not None) else x   # bf16: fp32 terms
        ct = ops.cast(ops.CisTransFn.apply(x32, gi), dt) if gi.cistrans is not None else x   # identity when empty (gnn.py:475-476)
        tt = ops.cast(ops.TetraFn.apply(x32, D, gi), dt) if gi.tetra is not None else x      # identity when empty (gnn.py:402-403)
        return self.stereochemical_embedding_2(([x, ct, tt], [D, D, D], Dp), packed=packed)  # padded [N, Dp], pads = 0

    # ------------------------------------------------------------------------------------------ misc API
    def init_weights(self) -> None:
        """Reference ``gnn.py:660-703`` (xavier-uniform on the listed layers and embeddings, zero biases)."""
        linear_layers = [self.embedding_projection, self.concat_self_other, self.post_pooling_projection,
                         self.skip_transform, self.output_layer, self.long_range_projection]
        if hasattr(self, "stereochemical_embedding"):
            linear_layers += [self.stereochemical_embedding, self.stereochemical_embedding_2]
        for layer in linear_layers:
            nn.init.xavier_uniform_(layer.weight)
            if layer.bias is not None:
                nn.init.zeros_(layer.bias)
        for emb in (self.atom_type_embedding, self.degree_embedding, self.hybridization_embedding,
                    self.hydrogen_count_embedding):
            nn.init.xavier_uniform_(emb.weight)
        if hasattr(self.pooling, "attention_weights"):
            for aw in self.pooling.attention_weights:
                nn.init.xavier_uniform_(aw.weight)
                if aw.bias is not None:
                    nn.init.zeros_(aw.bias)
        if self._verbose:
            print("[GNN] Model weights initialized")

    def get_model_info(self) -> Dict[str, object]:
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return {"total_parameters": total, "trainable_parameters": trainable, "hidden_dim": self.hidden_dim,
                "num_shells": self.num_shells, "embedding_dim": self.embedding_dim, "task_type": self.task_type,
                "use_partial_charges": self.use_partial_charges, "use_stereochemistry": self.use_stereochemistry,
                "loss_function": self.loss_function,
                "num_message_passing_layers": len(self.message_passing_layers),
                "pooling_type": type(self.pooling).__name__}

    def __repr__(self) -> str:
        info = self.get_model_info()
        return (f"GNN(\n  parameters={info['total_parameters']:,}\n  hidden_dim={info['hidden_dim']}\n"
                f"  num_shells={info['num_shells']}\n  task_type='{info['task_type']}'\n"
                f"  loss_function='{info['loss_function']}'\n"
                f"  features=[partial_charges={info['use_partial_charges']}, "
                f"stereochemistry={info['use_stereochemistry']}]\n)")


class GNNConfig:
    """Reference ``gnn.py:738-779``."""

    @staticmethod
    def from_args(args) -> Dict[str, object]:
        feature_sizes = {"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}
        return {"feature_sizes": feature_sizes, "hidden_dim": args.hidden_dim,
                "output_dim": getattr(args, "output_dim", 1), "num_shells": args.num_shells,
                "num_message_passing_layers": args.num_message_passing_layers, "ffn_hidden_dim": args.ffn_hidden_dim,
                "ffn_num_layers": args.ffn_num_layers, "pooling_type": args.pooling_type, "task_type": args.task_type,
                "embedding_dim": args.embedding_dim, "use_partial_charges": args.use_partial_charges,
                "use_stereochemistry": args.use_stereochemistry, "ffn_dropout": args.ffn_dropout,
                "activation_type": args.activation_type, "shell_conv_num_mlp_layers": args.shell_conv_num_mlp_layers,
                "shell_conv_dropout": args.shell_conv_dropout, "attention_num_heads": args.attention_num_heads,
                "attention_temperature": args.attention_temperature, "loss_function": args.loss_function}

    @staticmethod
    def create_model_from_args(args) -> GNN:
        return GNN(**GNNConfig.from_args(args))
