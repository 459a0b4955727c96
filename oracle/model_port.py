"""Floating-point side of the oracle (torch, CPU) -- TEST INFRASTRUCTURE.

A functional restatement of the reference hot path.  It keeps the reference's op sequence (gather ->
scatter_add, materialised ``[heads, N, F]`` product, ``cat`` of hop chunks, ...) so that its timing is a
fair "port" CPU baseline, and it keeps the reference's quirks Q1-Q6.
Parameters are passed as a ``state_dict``-style mapping with the reference's key names, so the very
same tensors can be loaded into the CUDA modules.  Gradients come from torch autograd on CPU.

Cited reference lines are relative to ``/root/reference/src``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import scatter_port as ts

ACTIVATIONS = {                                   # utils/activation.py:23-29
    "relu": F.relu,
    "leakyrelu": lambda v: F.leaky_relu(v, 0.01),
    "elu": F.elu,
    "gelu": F.gelu,
    "silu": F.silu,
}


def activation(name: str):
    if name not in ACTIVATIONS:                   # utils/activation.py:31-33
        raise ValueError(f"Invalid activation type: {name}. Supported: {', '.join(ACTIVATIONS)}")
    return ACTIVATIONS[name]


def linear(P, prefix, v):
    return F.linear(v, P[prefix + ".weight"], P.get(prefix + ".bias"))


# ----------------------------------------------------------------------------- a1
def message_passing(x, target, src, num_hops):
    """models/layers.py:133-167."""
    N = x.shape[0]
    if target.numel() == 0:                                      # :148-149
        return [torch.zeros_like(x) for _ in range(num_hops)]
    gathered = x[src % N]                                        # :154-155
    agg = ts.scatter_add(gathered, target, dim=0, dim_size=num_hops * N)   # :158-163
    return list(torch.split(agg, N, dim=0))                      # :166


# ----------------------------------------------------------------------------- a2
def shell_conv(P, prefix, x, target, src, num_hops, act="silu", num_mlp_layers=2,
               dropout_masks=None):
    """models/layers.py:63-108.  ``dropout_masks[k]`` (already scaled by 1/(1-p)) multiplies the hidden
    activation of MLP block k; ``None`` = eval / p=0."""
    f = activation(act)
    inp = torch.cat([x] + message_passing(x, target, src, num_hops), dim=-1)        # :76-79
    h = f(linear(P, prefix + ".input_proj", inp))                                   # :82-83
    if (prefix + ".global_skip_proj.weight") in P:                                  # :86-89
        g = linear(P, prefix + ".global_skip_proj", inp)
    else:
        g = h.clone()
    for k in range(num_mlp_layers):                                                 # :92-103
        t = f(linear(P, f"{prefix}.mlp_blocks.{k}.linear_1", h))
        if dropout_masks is not None:
            t = t * dropout_masks[k]
        h = linear(P, f"{prefix}.mlp_blocks.{k}.linear_2", t) + h
    return h + g                                                                    # :106


# ----------------------------------------------------------------------------- a4
def partial_charge(x, batch_indices, total_charges):
    """models/gnn.py:622-658."""
    q, f, rest = x.split([1, 1, x.shape[-1] - 2], dim=-1)
    f = torch.clamp(f, min=1e-6)
    B = total_charges.shape[0]
    idx = batch_indices.unsqueeze(1)
    Qu = torch.zeros((B, 1), dtype=x.dtype).scatter_add(0, idx, q)
    Fu = torch.zeros((B, 1), dtype=x.dtype).scatter_add(0, idx, f) + 1e-6
    Fu = torch.clamp(Fu, min=1e-6)
    dQ = total_charges.unsqueeze(-1) - Qu
    f_new = f / Fu[batch_indices]
    q_new = q + f_new * dQ[batch_indices]
    return torch.cat([q_new, f_new, rest], dim=-1)


# ----------------------------------------------------------------------------- a5
def tetrahedral(x, tetra):
    """models/gnn.py:387-462 (quirk Q4)."""
    if tetra.numel() == 0:
        return x
    upd = x.clone()
    raw = upd[tetra]                                              # [M,4,D]
    mag = torch.norm(raw, dim=-1, keepdim=True)
    e = F.normalize(raw, dim=-1, eps=1e-8)
    s = e ** 2
    s1, s2, s3 = (torch.roll(s, -k, 1) for k in (1, 2, 3))
    e1, e2, e3 = (torch.roll(e, -k, 1) for k in (1, 2, 3))
    chir = s1 * (e2 - e3) + s2 * (e3 - e1) + s3 * (e1 - e2)       # :429-433
    chir = chir * torch.tanh(mag.mean(dim=1, keepdim=True) / 3.0)  # :440-446
    idx = tetra.reshape(-1)
    upd.index_add_(0, idx, chir.reshape(-1, x.shape[-1]))          # :453
    mask = torch.zeros(x.shape[0], dtype=torch.bool)
    mask[torch.unique(idx)] = True
    upd[~mask] = 0.0                                               # :456-460
    return upd


def cis_trans(x, cis, trans):
    """models/gnn.py:465-509 (quirk Q3: rows 0 and 1 of the ``[2K,2]`` tensors are used as src/target)."""
    if cis.numel() == 0 and trans.numel() == 0:
        return x
    D = x.shape[1]
    if cis.numel() > 0:
        t_c, s_c = cis[1], x[cis[0]]
    else:
        t_c, s_c = torch.empty(0, dtype=torch.long), torch.empty(0, D, dtype=x.dtype)
    if trans.numel() > 0:
        t_t, s_t = trans[1], x[trans[0]]
    else:
        t_t, s_t = torch.empty(0, dtype=torch.long), torch.empty(0, D, dtype=x.dtype)
    tg = torch.cat([t_c, t_t], 0)
    sc = torch.cat([-s_c, s_t], 0)
    if tg.numel() == 0:
        return x
    return x.scatter_add(0, tg.unsqueeze(1).expand(-1, D), sc)


def stereo(P, x, tetra, cis, trans):
    """models/gnn.py:310-327."""
    cat = torch.cat([x, cis_trans(x, cis, trans), tetrahedral(x, tetra)], dim=-1)
    return linear(P, "stereochemical_embedding_2", cat)


# ----------------------------------------------------------------------------- a7 / a8
def attention_pool(P, prefix, x, batch_indices, num_heads):
    """models/pooling.py:122-172 (dropout_prob = 0)."""
    T = P[prefix + ".temperature"]
    scores = torch.stack([linear(P, f"{prefix}.attention_weights.{h}", x).squeeze(-1) / T
                          for h in range(num_heads)], dim=0)                         # :134-140
    idx = batch_indices.unsqueeze(0).expand(num_heads, -1)
    a = ts.scatter_softmax(scores, idx, dim=1)                                       # :145
    weighted = x.unsqueeze(0).expand(num_heads, -1, -1) * a.unsqueeze(-1)            # :150-154
    pooled = ts.scatter_sum(weighted, batch_indices, dim=1).mean(dim=0)              # :159-161
    return pooled, a


def simple_pool(kind, x, batch_indices):
    """models/pooling.py:15-80."""
    if kind == "mean":
        return ts.scatter_mean(x, batch_indices, dim=0), None
    if kind == "max":
        return ts.scatter_max(x, batch_indices, dim=0)[0], None
    if kind == "sum":
        return ts.scatter_add(x, batch_indices, dim=0), None
    raise ValueError(f"Unsupported pooling type: {kind}")


# ----------------------------------------------------------------------------- head pieces
def linear_block(P, prefix, v, act, use_skip, mask=None):
    """models/layers.py:203-219."""
    f = activation(act)
    t = f(linear(P, prefix + ".linear1", v))
    if mask is not None:
        t = t * mask
    out = linear(P, prefix + ".linear2", t)
    return out + v if use_skip else out


def ffn(P, prefix, v, num_layers, act, masks=None):
    """models/layers.py:236-267 with input=hidden=output width (gnn.py:122-130)."""
    for i in range(num_layers):
        skip = (num_layers > 1) and (0 < i < num_layers - 1)
        v = linear_block(P, f"{prefix}.layers.{i}", v, act, skip, None if masks is None else masks[i])
    return v


# ----------------------------------------------------------------------------- a3 / a6
def gnn_forward(P: Dict[str, torch.Tensor], cfg: dict, batch: dict, dropout: Optional[dict] = None):
    """models/gnn.py:197-308.  ``cfg`` keys: hidden_dim, num_shells, num_message_passing_layers,
    ffn_num_layers, pooling_type, use_partial_charges, use_stereochemistry, activation_type,
    shell_conv_num_mlp_layers, attention_num_heads.  ``batch`` holds torch tensors with the reference
    Batch field names.  Returns (output, attention_weights, partial_charges, extras)."""
    act = cfg.get("activation_type", "silu")
    f = activation(act)
    hidden = cfg["hidden_dim"]
    d_other = int(0.3 * hidden)                                   # gnn.py:100
    d_self = hidden - d_other
    fm = batch["atom_features_map"]
    emb = torch.cat([F.embedding(fm["atom_type"], P["atom_type_embedding.weight"]),
                     F.embedding(fm["hydrogen_count"], P["hydrogen_count_embedding.weight"]),
                     F.embedding(fm["degree"], P["degree_embedding.weight"]),
                     F.embedding(fm["hybridization"], P["hybridization_embedding.weight"])], dim=-1)   # :262-274
    h = f(linear(P, "embedding_projection", emb))                 # :224-225
    x_self, x = torch.split(h, [d_self, d_other], dim=-1)         # :227-231
    edges = batch["multi_hop_edge_indices"]
    bi = batch["batch_indices"]
    if edges.numel() > 0:                                         # :287
        for l in range(cfg.get("num_message_passing_layers", 3)):
            if cfg.get("use_partial_charges", False):
                x = partial_charge(x, bi, batch["total_charges"])
            if cfg.get("use_stereochemistry", False):
                x = stereo(P, x, batch["final_tetrahedral_chiral_tensor"],
                           batch["final_cis_tensor"], batch["final_trans_tensor"])
            masks = None if dropout is None else dropout["conv"][l]
            x = shell_conv(P, f"message_passing_layers.{l}", x, edges[:, 0], edges[:, 1],
                           cfg.get("num_shells", 3), act, cfg.get("shell_conv_num_mlp_layers", 2), masks) + x
    q = x[:, 0].clone() if (cfg.get("use_partial_charges", False) and x.shape[-1] >= 2) else None   # :240-242
    atom_emb = linear(P, "concat_self_other", torch.cat([x_self, x], dim=-1))       # :245-246
    kind = cfg.get("pooling_type", "attention")
    if kind == "attention":
        pooled, attn = attention_pool(P, "pooling", atom_emb, bi, cfg.get("attention_num_heads", 4))
    else:
        pooled, attn = simple_pool(kind, atom_emb, bi)
    v = linear(P, "post_pooling_projection", pooled)              # :252
    v = ffn(P, "ffn", v, cfg.get("ffn_num_layers", 3), act, None if dropout is None else dropout["ffn"])
    out = linear(P, "output_layer", torch.cat([v, linear(P, "skip_transform", v)], dim=-1))   # :256-258
    return out, attn, q, dict(atom_embeddings=atom_emb, pooled=pooled, x_other=x)


# ----------------------------------------------------------------------------- a11 / a12
def weighted_l1(pred, target, weights):
    """models/losses.py:29-49."""
    return (torch.abs(pred - target) * weights).sum(dim=1).mean()


def clip_and_adam(params, grads, state, step, lr=2.5e-4, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8):
    """``clip_grad_norm_(1.0)`` then ``Adam.step()`` (training/trainer.py:164-165; torch defaults).
    ``params``/``grads``/``state['m']``/``state['v']`` are lists of tensors updated in place."""
    total = torch.norm(torch.stack([torch.norm(g.detach(), 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, state["m"], state["v"]):
        g = g * coef
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / (1 - b2 ** step) ** 0.5).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return float(total)
