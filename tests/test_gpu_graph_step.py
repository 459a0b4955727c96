"""GPU: static-shape padding (collate.pad_batch) and the CUDA-graph captured training step (trainer.GraphedTrainStep)
against the eager step on the same batches."""
import numpy as np
import pytest
import torch

from helpers import RTOL_F32, assert_close, gnn_shapes
from oracle.fixtures import FEATURE_SIZES, det_state

pytestmark = pytest.mark.gpu
DEV = "cuda"
CFG = dict(hidden_dim=64, num_shells=3, num_message_passing_layers=2)


def _model(seed=9):
    import aimnet_x2d_b200 as ax
    m = ax.GNN(FEATURE_SIZES, 64, 3, num_shells=3, num_message_passing_layers=2, task_type="multitask",
               shell_conv_dropout=0.0, ffn_dropout=0.0)
    m.load_state_dict(det_state(gnn_shapes(CFG, 3), seed))
    return m.to(DEV).train()


def _fwd(model, b):
    return model(b.atom_features_map, b.multi_hop_edge_indices, b.batch_indices, b.total_charges,
                 b.final_tetrahedral_chiral_tensor, b.final_cis_tensor, b.final_trans_tensor, graph_index=b.graph_index)


def test_padding_leaves_real_molecules_untouched():
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch
    raw = S.make_batch(31, 24, 3, "qm9", num_targets=3)
    N = raw.graph_index.num_atoms
    pad = pad_batch(raw, N + 100, raw.graph_index.num_edges + 500, num_dummy=8)
    assert pad.graph_index.num_atoms == N + 100 and pad.graph_index.num_graphs == 32 and pad.num_real_graphs == 24
    assert np.array_equal(pad.graph_index.rowptr.numpy()[: N + 1], raw.graph_index.rowptr.numpy()[: N + 1])
    assert np.all(pad.graph_index.rowptr.numpy()[N:] == raw.graph_index.num_edges)        # dummy atoms have no edges
    crit = ax.WeightedL1Loss(torch.ones(3)).to(DEV)
    outs, grads = [], []
    for b, n in ((raw.to(DEV), 24), (pad.to(DEV), 24)):
        model = _model()
        out, attn, _ = _fwd(model, b)
        crit(out[:n], b.targets[:n]).backward()
        outs.append((out[:n].detach().cpu().numpy(), attn[:, :N].detach().cpu().numpy()))
        grads.append({k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters() if p.grad is not None})
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for k in grads[0]:
        if "attention_weights" in k and k.endswith(".bias"):
            continue                                           # rounding noise of an exactly cancelling sum
        assert_close(grads[1][k], grads[0][k], RTOL_F32, "padded vs unpadded grad " + k)


def test_graph_replay_equals_eager_step():
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch
    from aimnet_x2d_b200.trainer import GraphedTrainStep, TrainStep
    raws = [S.make_batch(40 + i, 24, 3, "qm9", num_targets=3) for i in range(3)]
    n_pad = max(b.graph_index.num_atoms for b in raws) + 64
    e_cap = max(b.graph_index.num_edges for b in raws) + 64
    first = [pad_batch(b, n_pad, e_cap, 8) for b in raws]
    t_cap = max(p.graph_index.n_tiles for p in first) + 2
    me_cap = max(p.graph_index.max_tile_edges for p in first)
    padded = [pad_batch(b, n_pad, e_cap, 8, t_cap, max_tile_edges=me_cap).pin_memory() for b in raws]
    crit = ax.WeightedL1Loss(torch.linspace(0.5, 1.5, 3)).to(DEV)
    m_e, m_g = _model(), _model()
    o_e = ax.FlatAdam(m_e.parameters(), lr=1e-3)
    o_g = ax.FlatAdam(m_g.parameters(), lr=1e-3)

    class PaddedEager(TrainStep):                              # eager step with the same "real rows only" loss
        def device_step(self, bd):
            self.optimizer.zero_grad()
            out, _, _ = _fwd(self.model, bd)
            loss = self.criterion(out[: bd.num_real_graphs], bd.targets[: bd.num_real_graphs])
            loss.backward()
            self.optimizer.step()
            return loss.detach()

    eager = PaddedEager(m_e, crit, o_e, DEV)
    graphed = GraphedTrainStep(m_g, crit, o_g, DEV)
    graphed.capture(padded[0], warmup=2)
    assert torch.equal(o_g.flat_param, o_e.flat_param)         # warm-up steps were rolled back
    assert int(o_g.step_count) == 0
    for it in range(5):
        b = padded[it % 3]
        le = eager(b)
        lg = graphed(b)
        assert le == lg, (it, le, lg)
    torch.cuda.synchronize()
    assert torch.equal(o_g.flat_param, o_e.flat_param) and int(o_g.step_count) == 5
    # a batch padded to other capacities is refused instead of silently replaying the wrong launch configuration
    other = pad_batch(raws[0], n_pad + 128, e_cap, 8, t_cap + 4, max_tile_edges=me_cap)
    with pytest.raises(RuntimeError):
        graphed(other)


def test_inference_step_matches_oracle_and_graph_replay():
    """Forward-only path (predictor.py / extractors.py): outputs, pooled embeddings (pooling hook) and partial charges."""
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch
    from oracle import model_port as MP
    cfg = dict(CFG, use_partial_charges=True)
    P = det_state(gnn_shapes(cfg, 3), 13)
    model = ax.GNN(FEATURE_SIZES, 64, 3, num_shells=3, num_message_passing_layers=2, task_type="multitask",
                   use_partial_charges=True)
    model.load_state_dict(P)
    model.to(DEV).train()                                       # the step switches to eval itself and restores the mode
    raw = S.make_batch(61, 20, 3, "drug", num_targets=3)
    r = ax.InferenceStep(model, DEV, atom_embeddings=True)(raw)
    assert model.training
    ob = dict(atom_features_map=raw.atom_features_map, multi_hop_edge_indices=raw.multi_hop_edge_indices,
              batch_indices=raw.batch_indices, total_charges=raw.total_charges,
              final_tetrahedral_chiral_tensor=raw.final_tetrahedral_chiral_tensor,
              final_cis_tensor=raw.final_cis_tensor, final_trans_tensor=raw.final_trans_tensor)
    ro, _, rq, extras = MP.gnn_forward(P, cfg, ob)
    assert_close(r["outputs"].cpu().numpy(), ro.numpy(), RTOL_F32, "inference outputs")
    assert_close(r["embeddings"].cpu().numpy(), extras["pooled"].numpy(), RTOL_F32, "pooled embeddings")
    assert_close(r["partial_charges"].cpu().numpy(), rq.numpy(), RTOL_F32, "partial charges")
    assert_close(r["atom_embeddings"].cpu().numpy()[:, :64], extras["atom_embeddings"].numpy(), RTOL_F32, "atom embeddings")
    # graph replay over padded batches == eager on the same padded batch, for a batch other than the captured one
    raws = [raw, S.make_batch(62, 20, 3, "drug", num_targets=3)]
    n_pad = max(b.graph_index.num_atoms for b in raws) + 64
    e_cap = max(b.graph_index.num_edges for b in raws) + 64
    first = [pad_batch(b, n_pad, e_cap, 32, tile_rows=192) for b in raws]
    caps = dict(num_tiles=max(p.graph_index.n_tiles for p in first) + 2,
                max_tile_edges=max(p.graph_index.max_tile_edges for p in first))
    padded = [pad_batch(b, n_pad, e_cap, 32, tile_rows=192, **caps) for b in raws]
    g = ax.GraphedInferenceStep(model, DEV)
    g.capture(padded[0])
    rg = g(padded[1])
    re_ = ax.InferenceStep(model, DEV)(padded[1])
    for k in ("outputs", "embeddings", "partial_charges"):
        assert torch.equal(rg[k], re_[k]), k
    assert_close(rg["outputs"].cpu().numpy(), MP.gnn_forward(P, cfg, dict(
        atom_features_map=raws[1].atom_features_map, multi_hop_edge_indices=raws[1].multi_hop_edge_indices,
        batch_indices=raws[1].batch_indices, total_charges=raws[1].total_charges,
        final_tetrahedral_chiral_tensor=raws[1].final_tetrahedral_chiral_tensor, final_cis_tensor=raws[1].final_cis_tensor,
        final_trans_tensor=raws[1].final_trans_tensor))[0].numpy(), RTOL_F32, "graphed inference outputs")


def test_stereo_batches_pad_to_a_static_shape_and_replay():
    """Drug-like batches with tetrahedral centres, cis/trans updates and partial charges: padding with dummy centres /
    zero-weight updates leaves the real molecules untouched, and the captured step replays batch after batch."""
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch, static_signature
    from aimnet_x2d_b200.trainer import GraphedTrainStep
    cfg = dict(CFG, use_partial_charges=True, use_stereochemistry=True)
    P = det_state(gnn_shapes(cfg, 3), 17)

    def model():
        m = ax.GNN(FEATURE_SIZES, 64, 3, num_shells=3, num_message_passing_layers=2, task_type="multitask",
                   use_partial_charges=True, use_stereochemistry=True, shell_conv_dropout=0.0, ffn_dropout=0.0)
        m.load_state_dict(P)
        return m.to(DEV).train()

    raws = [S.make_batch(80 + i, 24, 3, "drug", num_targets=3, stereo=True) for i in range(3)]
    assert all(b.graph_index.tetra is not None for b in raws)
    n_pad = max(b.graph_index.num_atoms for b in raws) + 96
    e_cap = max(b.graph_index.num_edges for b in raws) + 64
    kw = dict(tile_rows=192, num_tetra=max(b.graph_index.tetra[3] for b in raws) + 3, pad_cistrans=True)
    first = [pad_batch(b, n_pad, e_cap, 32, **kw) for b in raws]
    kw.update(num_tiles=max(p.graph_index.n_tiles for p in first) + 1,
              max_tile_edges=max(p.graph_index.max_tile_edges for p in first))
    padded = [pad_batch(b, n_pad, e_cap, 32, **kw).pin_memory() for b in raws]
    assert len({static_signature(p) for p in padded}) == 1
    # real molecules: identical outputs / charges with and without the padding
    m0 = model()
    o_raw, _, q_raw = _fwd(m0, raws[1].to(DEV))
    o_pad, _, q_pad = _fwd(m0, padded[1].to(DEV))
    n_real = raws[1].graph_index.num_atoms
    assert torch.equal(o_raw, o_pad[:24]) and torch.equal(q_raw, q_pad[:n_real])
    # graph replay == eager on the same padded batches
    crit = ax.WeightedL1Loss(torch.ones(3)).to(DEV)
    m_e, m_g = model(), model()
    o_e, o_g = ax.FlatAdam(m_e.parameters(), lr=1e-3), ax.FlatAdam(m_g.parameters(), lr=1e-3)
    graphed = GraphedTrainStep(m_g, crit, o_g, DEV)
    graphed.capture(padded[0], warmup=2)
    for it in range(4):
        b = padded[it % 3]
        bd = b.to(DEV)
        o_e.zero_grad()
        out, _, _ = _fwd(m_e, bd)
        le = crit(out[:24], bd.targets[:24])
        le.backward()
        o_e.step()
        assert float(le) == graphed(b), it
    torch.cuda.synchronize()
    assert torch.equal(o_g.flat_param, o_e.flat_param)


def test_host_batch_single_copy_and_prefetch_equal_per_tensor_loads():
    """GraphedTrainStep with HostBatch objects (one pinned buffer per batch, one copy, optionally prefetched on a side
    stream) must produce exactly the losses and parameters of the per-tensor load path."""
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch
    from aimnet_x2d_b200.trainer import GraphedTrainStep
    raws = [S.make_batch(60 + i, 24, 3, "qm9", num_targets=3) for i in range(4)]
    n_pad = max(b.graph_index.num_atoms for b in raws) + 64
    e_cap = max(b.graph_index.num_edges for b in raws) + 64
    first = [pad_batch(b, n_pad, e_cap, 8) for b in raws]
    t_cap = max(p.graph_index.n_tiles for p in first) + 2
    me_cap = max(p.graph_index.max_tile_edges for p in first)
    padded = [pad_batch(b, n_pad, e_cap, 8, t_cap, max_tile_edges=me_cap).pin_memory() for b in raws]
    crit = ax.WeightedL1Loss(torch.linspace(0.5, 1.5, 3)).to(DEV)
    m_a, m_b = _model(), _model()
    for m in (m_a, m_b):
        m.eval()                          # no dropout: the two runs must agree bit for bit
        m.train(False)
    s_a = GraphedTrainStep(m_a, crit, ax.FlatAdam(m_a.parameters(), lr=1e-3), DEV)
    s_b = GraphedTrainStep(m_b, crit, ax.FlatAdam(m_b.parameters(), lr=1e-3), DEV)
    s_a.capture(padded[0])
    s_b.capture(padded[0])
    packed = [s_b.pack(p) for p in padded]
    order = [0, 1, 2, 3, 1, 0, 3]
    la, lb = [], []
    for k, i in enumerate(order):
        la.append(s_a(padded[i]))
        nxt = packed[order[k + 1]] if k + 1 < len(order) and k % 2 == 0 else None     # prefetched and direct copies mixed
        lb.append(s_b(packed[i], prefetch=nxt))
    assert la == lb
    for (k, p), (_, q) in zip(m_a.named_parameters(), m_b.named_parameters()):
        assert torch.equal(p, q), k
    with pytest.raises(RuntimeError):
        s_b.pack(pad_batch(raws[0], n_pad + 32, e_cap, 8, t_cap, max_tile_edges=me_cap))


def test_two_captured_steps_of_different_shapes_share_one_model():
    """Two GraphedTrainSteps with different static signatures on the SAME model and optimiser, replayed alternately:
    the packed-weight descriptor tables and split-K workspaces a captured graph reads must survive the other capture.
    Checked against the eager step on a twin model."""
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch
    from aimnet_x2d_b200.trainer import GraphedTrainStep, TrainStep
    crit = ax.WeightedL1Loss(torch.linspace(0.5, 1.5, 3)).to(DEV)

    def padded(seed, graphs, extra):
        raw = S.make_batch(seed, graphs, 3, "qm9", num_targets=3)
        return pad_batch(raw, raw.graph_index.num_atoms + extra, raw.graph_index.num_edges + extra, 8).pin_memory()

    small, big = padded(90, 24, 40), padded(91, 96, 200)          # different atom / tile / molecule counts
    m_g, m_e = _model(), _model()
    for m in (m_g, m_e):
        m.train(False)
    o_g, o_e = ax.FlatAdam(m_g.parameters(), lr=1e-3), ax.FlatAdam(m_e.parameters(), lr=1e-3)
    s_small, s_big = GraphedTrainStep(m_g, crit, o_g, DEV), GraphedTrainStep(m_g, crit, o_g, DEV)
    s_small.capture(small)
    s_big.capture(big)                                             # larger workspaces, another descriptor table

    class PaddedEager(TrainStep):
        def device_step(self, bd):
            self.optimizer.zero_grad()
            out, _, _ = _fwd(self.model, bd)
            loss = self.criterion(out[: bd.num_real_graphs], bd.targets[: bd.num_real_graphs])
            loss.backward()
            self.optimizer.step()
            return loss.detach()

    eager = PaddedEager(m_e, crit, o_e, DEV)
    for which in (0, 1, 0, 0, 1, 1, 0):
        b, st = (small, s_small) if which == 0 else (big, s_big)
        lg = st(b)
        le = float(eager(b))
        assert abs(lg - le) <= 1e-6 * max(abs(le), 1.0), (which, lg, le)
    for (k, p), (_, q) in zip(m_g.named_parameters(), m_e.named_parameters()):
        assert float((p - q).abs().max()) <= 1e-6 * (float(q.abs().max()) + 1e-12), k


def _ring(n=3, B=24, seed0=60):
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch
    raws = [S.make_batch(seed0 + i, B, 3, "qm9", num_targets=3) for i in range(n)]
    n_pad = max(b.graph_index.num_atoms for b in raws) + 64
    e_cap = max(b.graph_index.num_edges for b in raws) + 64
    first = [pad_batch(b, n_pad, e_cap, 8) for b in raws]
    t_cap = max(p.graph_index.n_tiles for p in first) + 2
    me_cap = max(p.graph_index.max_tile_edges for p in first)
    return raws, [pad_batch(b, n_pad, e_cap, 8, t_cap, max_tile_edges=me_cap) for b in raws]


def test_step_fed_from_a_shard_file_equals_the_in_memory_path(tmp_path):
    """Pre-collated binary CSR shards (shards.py): capture from a batch READ BACK from the file, then train from the file's
    records (one memcpy into a pinned ring + one H2D copy per batch) -- the same losses, bit for bit, as feeding the padded
    batches directly; a scheduler-style change of the learning rate after capture is honoured by the replayed update."""
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200.trainer import GraphedTrainStep
    raws, padded = _ring()
    path = str(tmp_path / "ring.ax2d")
    ax.write_shard(path, padded)
    ds = ax.ShardDataset(path, ring=2)
    crit = ax.WeightedL1Loss(torch.linspace(0.5, 1.5, 3)).to(DEV)
    losses = []
    for mode in ("memory", "shard"):
        m = _model()
        o = ax.FlatAdam(m.parameters(), lr=1e-3)
        step = GraphedTrainStep(m, crit, o, DEV)
        step.capture(padded[0] if mode == "memory" else ds.batch(0))
        run = []
        src = [step.pack(b) for b in padded] if mode == "memory" else list(range(len(padded)))
        for i in range(6):
            hb = src[i % 3] if mode == "memory" else ds.host_batch(i % 3, i % 2)
            if i == 3:
                o.param_groups[0]["lr"] = 1e-4                  # what ReduceLROnPlateau does between steps
            run.append(step(hb))
        losses.append((run, m.output_layer.weight.detach().cpu().clone()))
    assert losses[0][0] == losses[1][0]
    assert torch.equal(losses[0][1], losses[1][1])
    # the lr change took effect: compare with a run that keeps lr = 1e-3
    m = _model()
    o = ax.FlatAdam(m.parameters(), lr=1e-3)
    step = GraphedTrainStep(m, crit, o, DEV)
    step.capture(padded[0])
    run = [step(step.pack(padded[i % 3])) for i in range(6)]
    assert run[:4] == losses[0][0][:4] and run[5] != losses[0][0][5]


def test_embedding_extractor_writes_what_the_eager_forward_returns(tmp_path):
    import aimnet_x2d_b200 as ax
    raws, padded = _ring(3, 20, 80)
    import copy
    m = ax.GNN(FEATURE_SIZES, 64, 3, num_shells=3, num_message_passing_layers=2, task_type="multitask",
               use_partial_charges=True).to(DEV).eval()
    step = ax.GraphedInferenceStep(m, DEV)
    step.capture(padded[0])
    w = ax.ShardedOutputWriter(str(tmp_path), 0, 1)
    n = ax.EmbeddingExtractor(step, w).run([p.pin_memory() for p in padded])
    w.close()
    assert n == 60
    merged = ax.merge_rank_outputs(str(tmp_path), 1)
    eager = ax.InferenceStep(m, DEV)
    outs, embs, qs = [], [], []
    for r in raws:
        res = eager(r)
        outs.append(res["outputs"].cpu().numpy())
        embs.append(res["embeddings"].cpu().numpy())
        qs.append(res["partial_charges"].cpu().numpy())
    assert_close(merged["outputs"], np.concatenate(outs), RTOL_F32, "extracted outputs")
    assert_close(merged["embeddings"], np.concatenate(embs), RTOL_F32, "extracted embeddings")
    assert_close(merged["partial_charges"], np.concatenate(qs), RTOL_F32, "extracted partial charges")
    n_atoms = np.concatenate([np.diff(r.graph_index.seg_ptr.numpy()) for r in raws])
    assert np.array_equal(merged["partial_charges_ptr"], np.concatenate([[0], np.cumsum(n_atoms)])[:-1])
