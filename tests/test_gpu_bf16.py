"""GPU: the bf16 configuration (BASELINE configs[3]; north_star tolerance 2e-2 relative) against the fp32 golden vectors of
the UNMODIFIED reference: bf16 activations / saved tensors / packed weight copies, fp32 accumulation, fp32 master weights and
gradients.  "Relative" as everywhere in this suite: against the scale of the tensor (tests/helpers.py)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import assert_close, cfg_from_golden, gnn_shapes
from oracle import model_port as MP
from oracle.fixtures import FEATURE_SIZES, batch_to_torch, det_state

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL_BF16 = 2e-2


def _ax():
    import aimnet_x2d_b200 as ax
    return ax


def _model(g, **kw):
    ax = _ax()
    cfg = cfg_from_golden(g)
    T = int(g["T"])
    model = ax.GNN(FEATURE_SIZES, cfg["hidden_dim"], T, num_shells=cfg["num_shells"],
                   num_message_passing_layers=cfg["num_message_passing_layers"], dropout=0.0,
                   ffn_num_layers=cfg.get("ffn_num_layers", 3), pooling_type=cfg.get("pooling_type", "attention"),
                   task_type="multitask", embedding_dim=cfg.get("embedding_dim", 64),
                   use_partial_charges=cfg.get("use_partial_charges", False),
                   use_stereochemistry=cfg.get("use_stereochemistry", False), ffn_dropout=0.0,
                   activation_type=cfg.get("activation_type", "silu"),
                   shell_conv_num_mlp_layers=cfg.get("shell_conv_num_mlp_layers", 2), shell_conv_dropout=0.0,
                   attention_num_heads=cfg.get("attention_num_heads", 4))
    model.load_state_dict(det_state(gnn_shapes(cfg, T), int(g["seed"])), strict=True)
    return model.to(DEV).train(), cfg, T


def _run(model, g, autocast=False):
    ax = _ax()
    b = batch_to_torch(g)
    mv = lambda t: t.to(DEV)
    gi = ax.GraphIndex.build(b["multi_hop_edge_indices"], b["batch_indices"], int(b["total_charges"].shape[0]), model.num_shells,
                             b["atom_features_map"], FEATURE_SIZES, b["final_tetrahedral_chiral_tensor"], b["final_cis_tensor"],
                             b["final_trans_tensor"]).to(DEV)
    args = ({k: mv(v) for k, v in b["atom_features_map"].items()}, mv(b["multi_hop_edge_indices"]), mv(b["batch_indices"]),
            mv(b["total_charges"]), mv(b["final_tetrahedral_chiral_tensor"]), mv(b["final_cis_tensor"]), mv(b["final_trans_tensor"]))
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, attn, q = model(*args, graph_index=gi)
    else:
        out, attn, q = model(*args, graph_index=gi)
    loss = ax.WeightedL1Loss(torch.from_numpy(g["loss_weights"])).to(DEV)(out, mv(b["targets"]))
    return out, attn, q, loss


def _upstream(g):
    """d loss / d out of the fp32 golden run (WeightedL1Loss: w_t * sign(out - y) / B).  The bf16 gradients are taken
    for THIS upstream gradient: the L1 loss is not smooth, so a 1 % output error flips the sign of residuals that happen to
    be near zero and changes the gradient of every parameter by O(1 / B) -- which says nothing about the kernels."""
    B = g["out"].shape[0]
    return (np.sign(g["out"].astype(np.float64) - g["targets"]) * g["loss_weights"] / B).astype(np.float32)


_AUTOCAST = {}


def autocast_reference_errors(name, g):
    """Relative errors (to the tensor scale) of the reference's op sequence under torch.autocast('cpu', bfloat16) -- the
    oracle port run on the CPU -- against the same fp32 golden, with the same fixed upstream gradient.  SURVEY.md section 8c
    names it as the secondary reference for bf16: it is the error bf16 arithmetic itself brings on this fixture."""
    if name in _AUTOCAST:
        return _AUTOCAST[name]
    cfg = cfg_from_golden(g)
    T = int(g["T"])
    P = {k: v.requires_grad_(True) for k, v in det_state(gnn_shapes(cfg, T), int(g["seed"])).items()}
    b = batch_to_torch(g)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out, attn, q, _ = MP.gnn_forward(P, cfg, b)
    (out.float() * torch.from_numpy(_upstream(g))).sum().backward()
    err = {"out": _rel(out.detach().float().numpy(), g["out"])}
    if "attn" in g:
        err["attn"] = _rel(attn.detach().float().numpy(), g["attn"])
    for k, v in P.items():
        if "g_" + k in g:
            err[k] = _rel(v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape)), g["g_" + k])
        elif float(g["gn_" + k]) > 0:
            err[k] = abs(float(v.grad.double().norm()) - float(g["gn_" + k])) / float(g["gn_" + k]) if v.grad is not None else 1.0
    _AUTOCAST[name] = err
    return err


def _rel(a, ref):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(float(np.abs(ref).max()), 1e-30))


def _bar(ac_err, key=""):
    """2e-2 of the tensor scale (north_star), or -- where bf16 arithmetic itself is further than that from fp32 on the
    fixture -- 2.5 x the error of the reference's own op sequence under autocast(bf16) (two independent realisations of
    bf16 rounding noise of the same magnitude; the statistic is a maximum over the tensor).  The temperature gradient is a
    heavily cancelling scalar sum over all (head, atom) pairs: 0.1."""
    bar = max(RTOL_BF16, 2.5 * ac_err)
    return max(bar, 0.1) if key.endswith("pooling.temperature") else bar


@pytest.mark.timeout(300)
@pytest.mark.parametrize("name", ["gnn_small", "gnn_h4_l3", "gnn_default", "gnn_stereo_charges", "gnn_small_gelu_mean"])
def test_bf16_model_matches_fp32_reference_within_2e2(name):
    g = load_golden(name)
    model, cfg, T = _model(g)
    if cfg["hidden_dim"] % 32:
        with pytest.raises(ValueError):
            model.compute_dtype = torch.bfloat16
            _run(model, g)
        return
    model.compute_dtype = torch.bfloat16
    out, attn, q, loss = _run(model, g)
    assert out.dtype == torch.float32
    # outputs and loss: the stated tolerance, no allowance
    assert_close(out.detach().cpu().numpy(), g["out"], RTOL_BF16, "output (bf16)")
    assert abs(float(loss.detach()) - float(g["loss"])) <= RTOL_BF16 * abs(float(g["loss"]))
    (out * torch.from_numpy(_upstream(g)).to(DEV)).sum().backward()
    charges = cfg.get("use_partial_charges", False)
    ac = {} if charges else autocast_reference_errors(name, g)      # (the reference crashes under autocast + charges, Q5)
    if "attn" in g:
        e = _rel(attn.detach().float().cpu().numpy(), g["attn"])
        assert e <= (_bar(ac.get("attn", 0.0)) if not charges else 3 * RTOL_BF16), f"attention weights (bf16): {e:.4f} of scale, autocast reference {ac.get('attn', 0.0):.4f}"
    if "q" in g:
        assert_close(q.detach().float().cpu().numpy(), g["q"], RTOL_BF16, "partial charges (bf16)")
    bad = []
    for k, p in model.named_parameters():
        assert p.grad is None or p.grad.dtype == torch.float32          # fp32 master gradients
        if "attention_weights." in k and k.endswith(".bias"):
            continue                                                     # exactly cancelling sum (see test_gpu_parity.py)
        gr = np.zeros(tuple(p.shape), np.float32) if p.grad is None else p.grad.float().cpu().numpy()
        if "g_" + k in g:
            e = _rel(gr, g["g_" + k])
        elif float(g["gn_" + k]) > 0:
            e = abs(float(np.linalg.norm(gr.astype(np.float64))) - float(g["gn_" + k])) / float(g["gn_" + k])
        else:
            e = float(np.abs(gr).max())
        bar = _bar(ac.get(k, 0.0), k) if not charges else _bar(3 * RTOL_BF16 / 2.5, k)
        if e > bar:
            bad.append(f"grad (bf16) {k}: {e:.4f} of scale, bar {bar:.4f} (autocast reference {ac.get(k, 0.0):.4f})")
    assert not bad, "\n".join(bad)


@pytest.mark.timeout(300)
def test_bf16_follows_autocast_and_fp32_is_untouched():
    g = load_golden("gnn_default")
    model, cfg, T = _model(g)
    o32, _, _, l32 = _run(model, g)
    assert_close(o32.detach().cpu().numpy(), g["out"], 1e-5, "fp32 output")
    o16, _, _, l16 = _run(model, g, autocast=True)
    assert float((o16 - o32).abs().max()) > 0.0                          # really another code path
    assert_close(o16.detach().cpu().numpy(), g["out"], RTOL_BF16, "output under autocast(bf16)")
    with pytest.raises(RuntimeError):
        with torch.autocast("cuda", dtype=torch.float16):
            _run(model, g, autocast=False)
    o32b, _, _, _ = _run(model, g)
    assert torch.equal(o32, o32b)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("width", [160, 32, 64, 40])
@pytest.mark.parametrize("tiled", [True, False])
def test_bf16_aggregation_equals_rounded_fp32_aggregation(width, tiled):
    """Same CSR order and fp32 accumulators as the fp32 kernel: the bf16 result is its fp32 result rounded once."""
    from aimnet_x2d_b200 import ops, synthetic as S
    batch = S.make_batch(5, 300, 3, "qm9")
    gi = batch.graph_index.to(DEV)
    if not tiled:
        gi.tile_local = False
    N = gi.num_atoms
    x = torch.randn(N, width, device=DEV).to(torch.bfloat16)
    add = torch.randn(N, width, device=DEV).to(torch.bfloat16)
    ref = ops.agg(x.float(), gi)
    got = ops.agg(x, gi)
    assert got.dtype == torch.bfloat16 and torch.equal(got, ref.to(torch.bfloat16))
    ref_b = ops.agg(x.float(), gi, transpose=True, addend=add.float())
    got_b = ops.agg(x, gi, transpose=True, addend=add)
    assert torch.equal(got_b, ref_b.to(torch.bfloat16))
    # optional tensor-core tiles: other order of the fp32 additions -> at most the last bf16 bit of an output differs
    ops.AGG_TENSOR_CORES = True
    try:
        for got_tc, want in ((ops.agg(x, gi), ref), (ops.agg(x, gi, transpose=True, addend=add), ref_b)):
            assert got_tc.dtype == torch.bfloat16
            err = (got_tc.float() - want).abs()
            assert bool((err <= want.abs() * 2.0 ** -8 + 1e-6).all())
    finally:
        ops.AGG_TENSOR_CORES = False


@pytest.mark.timeout(300)
def test_bf16_train_step_full_size_c4_shape_vs_oracle():
    """Foundation shape (hidden 512, L = 6, H = 4) on 512 QM9-shaped molecules: loss and the gradient norms of the bf16 step
    against the fp32 CPU oracle at 2e-2; graph-captured bf16 step == eager bf16 step."""
    ax = _ax()
    from aimnet_x2d_b200 import synthetic as S
    cfg = dict(hidden_dim=512, num_shells=4, num_message_passing_layers=6)
    T = 12
    batch = S.make_batch(4321, 512, 4, "qm9", T)
    # the reference's own initialisation (xavier-uniform, gnn.py:660-703): the deterministic large-weight fixtures used
    # elsewhere make a 6-layer network chaotic in ANY 8-bit-mantissa arithmetic (tools/bf16_error_table.py --c4: the
    # reference's op sequence under autocast(bf16) is 50-100 % off per gradient element there, this path 30-50 %)
    torch.manual_seed(5)
    model = ax.GNN(FEATURE_SIZES, 512, T, num_shells=4, num_message_passing_layers=6, task_type="multitask",
                   ffn_dropout=0.0, shell_conv_dropout=0.0)
    model.init_weights()
    P = OrderedDict((k, v.detach().clone()) for k, v in model.state_dict().items())
    model.to(DEV).train()
    model.compute_dtype = torch.bfloat16
    bd = batch.to(DEV)
    out, _, _ = model(bd.atom_features_map, bd.multi_hop_edge_indices, bd.batch_indices, bd.total_charges,
                      bd.final_tetrahedral_chiral_tensor, bd.final_cis_tensor, bd.final_trans_tensor, graph_index=bd.graph_index)
    w = torch.ones(T)
    loss = ax.WeightedL1Loss(w).to(DEV)(out, bd.targets)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    ob = dict(atom_features_map=batch.atom_features_map, multi_hop_edge_indices=batch.multi_hop_edge_indices,
              batch_indices=batch.batch_indices, total_charges=batch.total_charges,
              final_tetrahedral_chiral_tensor=batch.final_tetrahedral_chiral_tensor, final_cis_tensor=batch.final_cis_tensor,
              final_trans_tensor=batch.final_trans_tensor)
    ro, _, _, _ = MP.gnn_forward(Pr, cfg, ob)
    rl = MP.weighted_l1(ro, batch.targets, w)
    rl.backward()
    up = torch.sign(ro.detach() - batch.targets) / ro.shape[0]            # fixed upstream gradient (see _upstream)
    (out * up.to(DEV)).sum().backward()
    Pa = {k: v.clone().requires_grad_(True) for k, v in P.items()}        # the reference's op sequence under autocast(bf16)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ao, _, _, _ = MP.gnn_forward(Pa, cfg, ob)
    (ao.float() * up).sum().backward()
    assert_close(out.detach().cpu().numpy(), ro.detach().numpy(), RTOL_BF16, "output (bf16, foundation shape)")
    assert abs(float(loss.detach()) - float(rl)) <= RTOL_BF16 * abs(float(rl))
    bad = []
    for k, p in model.named_parameters():
        if "attention_weights." in k and k.endswith(".bias"):
            continue
        ref = Pr[k].grad
        rn = 0.0 if ref is None else float(ref.double().norm())
        gn = 0.0 if p.grad is None else float(p.grad.double().norm())
        an = 0.0 if Pa[k].grad is None else float(Pa[k].grad.double().norm())
        bar = _bar(abs(an - rn) / max(rn, 1e-12), k)
        if abs(gn - rn) > bar * max(rn, 1e-12) + 1e-12:
            bad.append(f"gradient norm of {k}: {gn:.6e} vs {rn:.6e} (autocast reference {an:.6e}, bar {bar:.3f})")
    assert not bad, "\n".join(bad)
