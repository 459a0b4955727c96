"""CPU: the oracle (oracle/*.py) must reproduce the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import assert_close, oracle_model_run
from oracle import graph_port as GP
from oracle import model_port as MP
from oracle import scatter_port as TS
from oracle.fixtures import det_state

torch.set_num_threads(1)


def test_bfs_kat_matches_reference():
    g = load_golden("kat_branch")
    hops = GP.shell_edges_bfs(6, g["bonds"], 3)
    for h in range(3):
        assert np.array_equal(hops[h], g[f"hop{h}"])
    # the hand-checked values of SURVEY.md section 8c
    assert hops[0].tolist() == [[0, 1, 1, 1, 2, 2, 3, 4, 4, 5], [1, 0, 2, 4, 1, 3, 2, 1, 5, 4]]
    assert hops[2].tolist() == [[0, 0, 2, 3, 3, 4, 5, 5], [3, 5, 5, 0, 4, 3, 0, 2]]


def test_bfs_random_matches_reference():
    g = load_golden("bfs_random")
    for i in range(int(g["count"])):
        hops = GP.shell_edges_bfs(int(g[f"n_{i}"]), g[f"bonds_{i}"], 4)
        for h in range(4):
            assert np.array_equal(hops[h], g[f"hop_{i}_{h}"].reshape(2, -1)), (i, h)


def test_collate_matches_reference():
    g = load_golden("kat_branch")
    hops = [g["hop0"], g["hop1"], g["hop2"]]
    feat = lambda n: {"atom_type": np.arange(n) % 119}
    mols = [dict(num_atoms=6, hops=hops, features=feat(6), target=[1.0, 2.0], total_charge=0.0,
                 chiral=[[0, 1, 2, 3], [1, 2, 4]], cis=[[0, 2]], trans=[[3, 5]]),
            dict(num_atoms=6, hops=hops, features=feat(6), target=[3.0, 4.0], total_charge=1.0,
                 chiral=[[0, 1, 2, 3]], cis=[], trans=[[1, 4], [2, 5]])]
    b = GP.collate(mols)
    assert np.array_equal(b["multi_hop_edge_indices"], g["edges"]) and b["multi_hop_edge_indices"].shape == (56, 2)
    assert np.array_equal(b["batch_indices"], g["batch_indices"])
    assert np.array_equal(b["final_tetrahedral_chiral_tensor"], g["tetra"])
    assert b["final_tetrahedral_chiral_tensor"].tolist() == [[0, 1, 2, 3], [6, 7, 8, 9]]
    assert np.array_equal(b["final_cis_tensor"], g["cis"])
    assert np.array_equal(b["final_trans_tensor"], g["trans"])
    assert np.array_equal(b["targets"], g["targets"]) and np.array_equal(b["total_charges"], g["total_charges"])
    assert np.array_equal(b["atom_features_map"]["atom_type"], g["feat_atom_type"])


def _layer_shapes(D, H, n_mlp=2):
    s = OrderedDict()
    s["input_proj.weight"] = (D, D * (H + 1)); s["input_proj.bias"] = (D,)
    for k in range(n_mlp):
        for n in ("linear_1", "linear_2"):
            s[f"mlp_blocks.{k}.{n}.weight"] = (D, D); s[f"mlp_blocks.{k}.{n}.bias"] = (D,)
    s["global_skip_proj.weight"] = (D, D * (H + 1)); s["global_skip_proj.bias"] = (D,)
    return s


@pytest.mark.parametrize("act", ["silu", "relu", "leakyrelu", "elu", "gelu", "hopoffset"])
def test_shell_conv_matches_reference(act):
    g = load_golden(f"layer_{act}")
    P = {"L." + k: v.requires_grad_(True) for k, v in det_state(_layer_shapes(19, 3), 11).items()}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    if act == "hopoffset":
        tgt, src, name = torch.from_numpy(g["target"]), torch.from_numpy(g["src"]), "silu"
    else:
        e = torch.from_numpy(g["edges"]); tgt, src, name = e[:, 0], e[:, 1], act
    out = MP.shell_conv(P, "L", x, tgt, src, 3, name, 2)
    (out * torch.from_numpy(g["R"])).sum().backward()
    assert np.array_equal(out.detach().numpy(), g["out"])          # same ops, same order -> bit identical
    assert np.array_equal(x.grad.numpy(), g["gx"])
    for k, v in P.items():
        assert np.array_equal(v.grad.numpy(), g["g_" + k[2:]]), k
    if act == "hopoffset":
        mp = MP.message_passing(torch.from_numpy(g["x"]), tgt, src, 3)
        assert np.array_equal(np.stack([c.numpy() for c in mp]), g["mp"])
    else:       # quirk Q1: hop chunks 1..H-1 are exactly zero under the shipped collation
        mp = MP.message_passing(torch.from_numpy(g["x"]), tgt, src, 3)
        assert float(mp[1].abs().max()) == 0.0 and float(mp[2].abs().max()) == 0.0
        csr = GP.csr_artefacts(g["edges"], g["x"].shape[0], 3)
        agg = GP.aggregate_rows(g["x"], csr["rowptr"][: g["x"].shape[0] + 1], csr["col"])
        assert np.array_equal(agg, mp[0].numpy())                  # stable-CSR sequential sum == scatter_add


def test_attention_pool_matches_reference():
    g = load_golden("pool_attention")
    s = OrderedDict([("temperature", ())])
    for h in range(4):
        s[f"attention_weights.{h}.weight"] = (1, 64); s[f"attention_weights.{h}.bias"] = (1,)
    P = {"pooling." + k: v.requires_grad_(True) for k, v in det_state(s, 12).items()}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    pooled, attn = MP.attention_pool(P, "pooling", x, torch.from_numpy(g["batch_indices"]), 4)
    ((pooled * torch.from_numpy(g["Rp"])).sum() + 0.3 * (attn * torch.from_numpy(g["Ra"])).sum()).backward()
    assert np.array_equal(pooled.detach().numpy(), g["pooled"]) and np.array_equal(attn.detach().numpy(), g["attn"])
    assert np.array_equal(x.grad.numpy(), g["gx"])
    for k, v in P.items():
        assert np.array_equal(v.grad.numpy(), g["g_" + k[len("pooling."):]]), k


@pytest.mark.parametrize("kind", ["mean", "max", "sum"])
def test_simple_pool_matches_reference(kind):
    g = load_golden(f"pool_{kind}")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    pooled, _ = MP.simple_pool(kind, x, torch.from_numpy(g["batch_indices"]))
    (pooled * torch.from_numpy(g["Rp"])).sum().backward()
    assert np.array_equal(pooled.detach().numpy(), g["pooled"])
    assert np.array_equal(x.grad.numpy(), g["gx"])


def test_scatter_max_first_maximum_rule():
    src = torch.tensor([[1.0, 5.0], [1.0, 2.0], [3.0, 5.0], [-1.0, -2.0]], requires_grad=True)
    idx = torch.tensor([0, 0, 0, 2])
    vals, arg = TS.scatter_max(src, idx, dim=0, dim_size=3)
    assert vals.tolist() == [[3.0, 5.0], [0.0, 0.0], [-1.0, -2.0]]          # empty segment -> 0
    assert arg.tolist() == [[2, 0], [4, 4], [3, 3]]                          # first maximum; sentinel = src.size(0)
    vals.sum().backward()
    assert src.grad.tolist() == [[0.0, 1.0], [0.0, 0.0], [1.0, 0.0], [1.0, 1.0]]


@pytest.mark.parametrize("name", ["gnn_small", "gnn_small_gelu_mean", "gnn_stereo_charges", "gnn_stereo_empty",
                                  "gnn_h4_l3", "gnn_default"])
def test_gnn_matches_reference(name):
    g = load_golden(name)
    out, attn, q, loss, grads, P, cfg, batch = oracle_model_run(g)
    assert_close(out.detach().numpy(), g["out"], 1e-6, "output")
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    if "attn" in g:
        assert_close(attn.detach().numpy(), g["attn"], 1e-6, "attention weights")
    if "q" in g:
        assert_close(q.detach().numpy(), g["q"], 1e-6, "partial charges")
    for k, v in grads.items():
        if "g_" + k in g:
            assert_close(v, g["g_" + k], 1e-6, "grad " + k)
        ref_norm = float(g["gn_" + k])
        assert abs(np.linalg.norm(v.astype(np.float64)) - ref_norm) <= 1e-6 * max(ref_norm, 1e-12), k
    if name.startswith("gnn_stereo") or name == "gnn_small":     # dead / zero-gradient parameters of the reference
        assert float(g["gn_long_range_projection.weight"]) == 0.0


def test_clip_and_adam_matches_reference():
    g = load_golden("gnn_small")
    out, attn, q, loss, grads, P, cfg, batch = oracle_model_run(g)
    keys = [k for k in P if not k.startswith("long_range") and P[k].grad is not None]
    params = [P[k].detach().clone() for k in keys]
    gl = [P[k].grad.clone() for k in keys]
    state = dict(m=[torch.zeros_like(p) for p in params], v=[torch.zeros_like(p) for p in params])
    MP.clip_and_adam(params, gl, state, step=1)
    for k, p in zip(keys, params):
        assert_close(p.numpy(), g["p1_" + k], 1e-6, "param after step " + k)
    # parameters without gradient are left untouched by torch.optim.Adam (grad is None)
    assert np.array_equal(g["p1_long_range_projection.weight"], P["long_range_projection.weight"].detach().numpy())
