"""Shared helpers of the test-suite (CPU oracle side and GPU side)."""
from collections import OrderedDict

import numpy as np
import torch

from oracle import model_port as MP
from oracle.fixtures import FEATURE_SIZES, batch_to_torch, det_state

# fp32 parity bar of BASELINE.json: 1e-5 relative.  "Relative" is taken against the scale of the tensor
# (max |reference|), the usual reading for accumulated fp32 results; element-wise the check is
# |a - b| <= RTOL * |b| + RTOL * max|b|.
RTOL_F32 = 1e-5


def assert_close(a, b, rtol=RTOL_F32, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = float(np.max(np.abs(b))) if b.size else 0.0
    err = np.abs(a - b) - rtol * np.abs(b)
    worst = float(np.max(err)) if err.size else 0.0
    assert worst <= rtol * scale + 1e-30, (f"{what}: max |a-b| = {float(np.max(np.abs(a - b))):.3e}, scale {scale:.3e}, "
                                           f"allowed {rtol * scale:.3e}")


def cfg_from_golden(g):
    cfg = {}
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        v = str(v)
        cfg[str(k)] = True if v == "True" else False if v == "False" else (int(v) if v.lstrip("-").isdigit() else v)
    return cfg


def gnn_shapes(cfg, T):
    """state_dict keys/shapes of the reference GNN for a configuration (gnn.py:50-195), without the reference."""
    hid = cfg["hidden_dim"]
    E = cfg.get("embedding_dim", 64)
    D = int(0.3 * hid)
    H = cfg.get("num_shells", 3)
    s = OrderedDict()
    for k in ("atom_type", "hydrogen_count", "degree", "hybridization"):
        s[f"{k}_embedding.weight"] = (FEATURE_SIZES[k], E)
    s["embedding_projection.weight"] = (hid, 4 * E)
    s["embedding_projection.bias"] = (hid,)
    for l in range(cfg.get("num_message_passing_layers", 3)):
        p = f"message_passing_layers.{l}"
        s[p + ".input_proj.weight"] = (D, D * (H + 1)); s[p + ".input_proj.bias"] = (D,)
        for k in range(cfg.get("shell_conv_num_mlp_layers", 2)):
            for n in ("linear_1", "linear_2"):
                s[f"{p}.mlp_blocks.{k}.{n}.weight"] = (D, D); s[f"{p}.mlp_blocks.{k}.{n}.bias"] = (D,)
        s[p + ".global_skip_proj.weight"] = (D, D * (H + 1)); s[p + ".global_skip_proj.bias"] = (D,)
    if cfg.get("pooling_type", "attention") == "attention":
        s["pooling.temperature"] = ()
        for h in range(cfg.get("attention_num_heads", 4)):
            s[f"pooling.attention_weights.{h}.weight"] = (1, hid); s[f"pooling.attention_weights.{h}.bias"] = (1,)
    s["concat_self_other.weight"] = (hid, hid); s["concat_self_other.bias"] = (hid,)
    if cfg.get("use_stereochemistry", False):
        s["stereochemical_embedding.weight"] = (hid, 3 * hid); s["stereochemical_embedding.bias"] = (hid,)
        s["stereochemical_embedding_2.weight"] = (D, 3 * D); s["stereochemical_embedding_2.bias"] = (D,)
    s["post_pooling_projection.weight"] = (hid, hid); s["post_pooling_projection.bias"] = (hid,)
    for i in range(cfg.get("ffn_num_layers", 3)):
        for n in ("linear1", "linear2"):
            s[f"ffn.layers.{i}.{n}.weight"] = (hid, hid); s[f"ffn.layers.{i}.{n}.bias"] = (hid,)
    s["skip_transform.weight"] = (hid, hid); s["skip_transform.bias"] = (hid,)
    s["output_layer.weight"] = (T, 2 * hid); s["output_layer.bias"] = (T,)
    s["long_range_projection.weight"] = (hid, hid); s["long_range_projection.bias"] = (hid,)
    return s


def oracle_model_run(g, with_grads=True):
    """Run the oracle GNN on a golden model fixture; returns (out, attn, q, loss, grads dict, params dict)."""
    cfg = cfg_from_golden(g)
    T = int(g["T"])
    P = det_state(gnn_shapes(cfg, T), int(g["seed"]))
    if with_grads:
        for v in P.values():
            v.requires_grad_(True)
    batch = batch_to_torch(g)
    out, attn, q, extras = MP.gnn_forward(P, cfg, batch)
    loss = MP.weighted_l1(out, batch["targets"], torch.from_numpy(g["loss_weights"]))
    grads = {}
    if with_grads:
        loss.backward()
        grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32)) for k, v in P.items()}
    return out, attn, q, loss, grads, P, cfg, batch
