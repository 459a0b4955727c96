"""GPU: the packed-weights path of GNN (packed.PackedWeights: one pack kernel per forward, one gradient-collect kernel
per backward) against the per-call path that derives every operand from the parameters with torch ops.  Same kernels,
same operands => the results must be bit-identical."""
import numpy as np
import pytest
import torch

from helpers import gnn_shapes
from oracle.fixtures import FEATURE_SIZES, det_state

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(stereo, charges, packed, seed=5):
    import aimnet_x2d_b200 as ax
    cfg = dict(hidden_dim=64, num_shells=3, num_message_passing_layers=2, use_stereochemistry=stereo,
               use_partial_charges=charges)
    m = ax.GNN(FEATURE_SIZES, 64, 3, num_shells=3, num_message_passing_layers=2, task_type="multitask",
               shell_conv_dropout=0.0, ffn_dropout=0.0, use_stereochemistry=stereo, use_partial_charges=charges)
    m.load_state_dict(det_state(gnn_shapes(cfg, 3), seed))
    m.use_packed_weights = packed
    return m.to(DEV).train()


def _run(model, b, steps=2):
    import aimnet_x2d_b200 as ax
    crit = ax.WeightedL1Loss(torch.linspace(0.5, 1.5, 3)).to(DEV)
    outs = []
    for _ in range(steps):                       # gradients accumulate over the steps, as autograd's do
        out, attn, _ = model(b.atom_features_map, b.multi_hop_edge_indices, b.batch_indices, b.total_charges,
                             b.final_tetrahedral_chiral_tensor, b.final_cis_tensor, b.final_trans_tensor,
                             graph_index=b.graph_index)
        crit(out, b.targets).backward()
        outs.append(out.detach().cpu().numpy())
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters() if p.grad is not None}
    return outs, grads


@pytest.fixture
def immediate_reduce():
    """The packed path normally leaves the split-K partials of the weight gradients to the gradient-collect kernel
    (a different, equally fixed summation order); with the reduce done per matrix the two paths are bit-identical."""
    from aimnet_x2d_b200 import ops
    old = ops.DEFER_SPLITK_REDUCE
    ops.DEFER_SPLITK_REDUCE = False
    yield
    ops.DEFER_SPLITK_REDUCE = old


@pytest.mark.parametrize("stereo,charges", [(False, False), (True, True)])
def test_packed_path_is_bit_identical_to_per_call_path(stereo, charges, immediate_reduce):
    from aimnet_x2d_b200 import synthetic as S
    b = S.make_batch(77, 48, 3, "druglike" if stereo else "qm9", num_targets=3, stereo=stereo).to(DEV)
    m_p, m_l = _model(stereo, charges, True), _model(stereo, charges, False)
    o_p, g_p = _run(m_p, b)
    o_l, g_l = _run(m_l, b)
    assert m_p._packed and not m_l._packed                       # the packed path really ran (and only there)
    for a, c in zip(o_p, o_l):
        assert np.array_equal(a, c)
    assert set(g_p) == set(g_l)
    for k in g_l:
        # step 1 is bit-identical; accumulating step 2 adds (g1 + g2) in one order or the other: 1 ulp
        np.testing.assert_allclose(g_p[k], g_l[k], rtol=3e-7, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("stereo,charges", [(False, False), (True, True)])
def test_deferred_splitk_reduce_matches_immediate_reduce(stereo, charges):
    """Default packed path (partials summed by ax2d_unpack_grads) against the per-call path: same values up to the
    rounding of a different summation order; and deterministic from run to run."""
    from aimnet_x2d_b200 import ops, synthetic as S
    assert ops.DEFER_SPLITK_REDUCE
    b = S.make_batch(81, 64, 3, "druglike" if stereo else "qm9", num_targets=3, stereo=stereo).to(DEV)
    o_p, g_p = _run(_model(stereo, charges, True), b)
    o_q, g_q = _run(_model(stereo, charges, True), b)
    o_l, g_l = _run(_model(stereo, charges, False), b)
    for a, c in zip(o_p, o_l):
        assert np.array_equal(a, c)
    for k in g_l:
        assert np.array_equal(g_p[k], g_q[k]), k                      # deterministic
        scale = float(np.abs(g_l[k]).max()) + 1e-30
        assert float(np.abs(g_p[k] - g_l[k]).max()) <= 2e-6 * scale, k


def test_packed_first_step_gradients_bitwise(immediate_reduce):
    from aimnet_x2d_b200 import synthetic as S
    b = S.make_batch(78, 32, 3, "qm9", num_targets=3).to(DEV)
    _, g_p = _run(_model(False, False, True), b, steps=1)
    _, g_l = _run(_model(False, False, False), b, steps=1)
    for k in g_l:
        assert np.array_equal(g_p[k], g_l[k]), k


def test_packed_follows_parameter_updates_and_state_dict(immediate_reduce):
    """The packed operands are a cache: an optimiser step / load_state_dict must show up in the next forward."""
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.trainer import TrainStep
    b = S.make_batch(79, 32, 3, "qm9", num_targets=3).to(DEV)
    m_p, m_l = _model(False, False, True), _model(False, False, False)
    o_p, o_l = ax.FlatAdam(m_p.parameters(), lr=1e-2), ax.FlatAdam(m_l.parameters(), lr=1e-2)
    crit = ax.WeightedL1Loss(torch.ones(3)).to(DEV)
    for _ in range(3):
        losses = []
        for m, o in ((m_p, o_p), (m_l, o_l)):
            losses.append(TrainStep(m, crit, o, DEV).device_step(b).item())
        assert losses[0] == losses[1]
    for (k, p), (_, q) in zip(m_p.named_parameters(), m_l.named_parameters()):
        assert torch.equal(p, q), k
    m_p.load_state_dict({k: v * 0.5 for k, v in m_l.state_dict().items()})
    m_l.load_state_dict(m_p.state_dict())
    with torch.no_grad():
        a = m_p.eval()(b.atom_features_map, b.multi_hop_edge_indices, b.batch_indices, b.total_charges,
                       b.final_tetrahedral_chiral_tensor, b.final_cis_tensor, b.final_trans_tensor, graph_index=b.graph_index)[0]
        c = m_l.eval()(b.atom_features_map, b.multi_hop_edge_indices, b.batch_indices, b.total_charges,
                       b.final_tetrahedral_chiral_tensor, b.final_cis_tensor, b.final_trans_tensor, graph_index=b.graph_index)[0]
    assert torch.equal(a, c)


def test_frozen_embeddings_fall_back_to_per_call_path():
    from aimnet_x2d_b200 import synthetic as S
    b = S.make_batch(80, 16, 3, "qm9", num_targets=3).to(DEV)
    m = _model(False, False, True)
    m.atom_type_embedding.weight.requires_grad_(False)
    _, g = _run(m, b, steps=1)
    assert not m._packed and "concat_self_other.weight" in g
