"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference modules.

Runs only in the build container (needs /root/reference, read-only).  The reference ships no tests or
golden vectors, so these fixtures ARE the parity pins: the oracle (oracle/*.py) must
reproduce them on CPU (tests/test_oracle_golden.py), and the CUDA path must reproduce the oracle / the
fixtures on the GPU (tests/test_gpu_parity.py).

    python tests/golden/make_golden.py

`torch_scatter`, `torch_geometric`, `rdkit`, `h5py` are absent here; `torch_scatter` is provided by
oracle/scatter_port.py (restatement of torch_scatter 2.1.2), the others by empty stub modules -- the
reference files models/{gnn,layers,pooling,losses}.py and datasets/{features,molecular}.py are imported
as they are from /root/reference/src.
"""
import importlib.util
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import scatter_port  # noqa: E402
from oracle.fixtures import FEATURE_SIZES, batch_to_arrays, det_state  # noqa: E402

sys.modules["torch_scatter"] = scatter_port
sys.path.insert(0, REF)
from models.gnn import GNN as RefGNN  # noqa: E402
from models.layers import ShellConvolutionLayer as RefShell  # noqa: E402
from models.losses import WeightedL1Loss as RefL1  # noqa: E402
from models.pooling import (MaxPoolingLayer, MeanPoolingLayer, MultiHeadAttentionPoolingLayer,  # noqa: E402
                            SumPoolingLayer)


def load_ref_datasets():
    """datasets/{features,molecular}.py without importing the package __init__ (which needs RDKit)."""
    for name in ("rdkit", "rdkit.Chem", "rdkit.Chem.rdBase", "rdkit.Chem.rdchem", "h5py", "torch_geometric",
                 "torch_geometric.data"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]
    sys.modules["rdkit.Chem"].rdBase = sys.modules["rdkit.Chem.rdBase"]
    sys.modules["rdkit.Chem"].rdchem = sys.modules["rdkit.Chem.rdchem"]
    sys.modules["rdkit.Chem.rdchem"].HybridizationType = types.SimpleNamespace(S=0, SP=1, SP2=2, SP3=3, SP3D=4, SP3D2=5)
    tgd = sys.modules["torch_geometric.data"]

    class Bag:
        def __init__(self, *a, **k):
            pass
    tgd.Data, tgd.Batch, tgd.InMemoryDataset = type("Data", (Bag,), {}), type("Batch", (Bag,), {}), type("IMD", (Bag,), {})
    pkg = types.ModuleType("refds")
    pkg.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["refds"] = pkg
    mods = {}
    for n in ("constants", "features", "molecular"):
        spec = importlib.util.spec_from_file_location(f"refds.{n}", os.path.join(REF, "datasets", f"{n}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refds.{n}"] = m
        spec.loader.exec_module(m)
        mods[n] = m
    return mods, tgd


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KiB)")


def shapes_of(module):
    return OrderedDict((k, tuple(v.shape)) for k, v in module.state_dict().items())


def grads_of(module):
    return {k: (p.grad.detach().numpy().copy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
            for k, p in module.named_parameters()}


def main():
    torch.set_num_threads(1)
    torch.manual_seed(0)
    from aimnet_x2d_b200 import synthetic as S          # input generator only (integers + seeds)

    # ---------------------------------------------------------------- 1. integer KAT: reference BFS + collation
    mods, tgd = load_ref_datasets()
    feats, mol = mods["features"], mods["molecular"]
    bonds = [(0, 1), (1, 2), (2, 3), (1, 4), (4, 5)]
    adj = np.zeros((6, 6), dtype=np.int32)
    for a, b in bonds:
        adj[a, b] = adj[b, a] = 1
    hops = feats.compute_multi_hop_edges_bfs_numba(feats.build_numba_adjacency_list(adj), 3)
    hops = [np.asarray(h) for h in hops]

    def data_obj(n, hops, chiral, cis, trans, target, charge):
        d = tgd.Data()
        d.x = torch.zeros((n, 1))
        d.multi_hop_edges = [torch.from_numpy(np.asarray(h)).long() for h in hops]
        d.atom_features_map = {k: torch.arange(n) % v for k, v in FEATURE_SIZES.items()}
        d.target = torch.tensor(target, dtype=torch.float32)
        d.total_charge = torch.tensor([charge], dtype=torch.float32)
        d.chiral_tensors = [torch.tensor(c) for c in chiral]
        d.cis_bonds_tensors = [torch.tensor(c) for c in cis]
        d.trans_bonds_tensors = [torch.tensor(c) for c in trans]
        d.smiles = "X"
        d.atomic_numbers = torch.ones(n, dtype=torch.long)
        return d
    dl = [data_obj(6, hops, [[0, 1, 2, 3], [1, 2, 4]], [[0, 2]], [[3, 5]], [1.0, 2.0], 0.0),
          data_obj(6, hops, [[0, 1, 2, 3]], [], [[1, 4], [2, 5]], [3.0, 4.0], 1.0)]
    b = mol.MyBatch.from_data_list(dl)
    save("kat_branch", bonds=np.array(bonds), hop0=hops[0], hop1=hops[1], hop2=hops[2],
         edges=b.multi_hop_edge_indices.numpy(), batch_indices=b.batch_indices.numpy(),
         tetra=b.final_tetrahedral_chiral_tensor.numpy(), cis=b.final_cis_tensor.numpy(),
         trans=b.final_trans_tensor.numpy(), targets=b.targets.numpy(), total_charges=b.total_charges.numpy(),
         feat_atom_type=b.atom_features_map["atom_type"].numpy())

    # reference BFS on random synthetic molecules (pins ax2d_host_shell_edges / oracle.graph_port.shell_edges_bfs)
    mols = S.make_molecules(77, 12, 4, "qm9")
    rec = {}
    for i, m in enumerate(mols):
        adj = np.zeros((m["num_atoms"], m["num_atoms"]), dtype=np.int32)
        for a, bb in m["bonds"]:
            adj[a, bb] = adj[bb, a] = 1
        ref_h = feats.compute_multi_hop_edges_bfs_numba(feats.build_numba_adjacency_list(adj), 4)
        rec[f"n_{i}"] = m["num_atoms"]
        rec[f"bonds_{i}"] = m["bonds"]
        for h in range(4):
            rec[f"hop_{i}_{h}"] = np.asarray(ref_h[h])
    save("bfs_random", count=len(mols), **rec)

    # ---------------------------------------------------------------- 2. ShellConvolutionLayer, 5 activations
    batch = S.make_batch(101, 6, 3, "qm9")
    edges = batch.multi_hop_edge_indices
    N = int(batch.batch_indices.shape[0])
    rng = np.random.Generator(np.random.PCG64(5))
    D = 19
    x0 = rng.normal(0, 1, size=(N, D)).astype(np.float32)
    R = rng.normal(0, 1, size=(N, D)).astype(np.float32)
    for act in ("silu", "relu", "leakyrelu", "elu", "gelu"):
        layer = RefShell(D, D, num_hops=3, dropout=0.0, activation_type=act, num_mlp_layers=2)
        layer.load_state_dict(det_state(shapes_of(layer), 11))
        x = torch.from_numpy(x0).requires_grad_(True)
        out = layer(x, edges[:, 0], edges[:, 1])
        (out * torch.from_numpy(R)).sum().backward()
        save(f"layer_{act}", x=x0, R=R, edges=np.ascontiguousarray(edges.numpy()), out=out.detach().numpy(),
             gx=x.grad.numpy(), **{"g_" + k: v for k, v in grads_of(layer).items()})
    # general contract: target = hop * N + atom (layers.py:139)
    pieces, hop_id = [], []
    mols = S.make_molecules(101, 6, 3, "qm9")
    off = 0
    for m in mols:
        for h, e in enumerate(m["hops"]):
            pieces.append(e + off)
            hop_id.append(np.full(e.shape[1], h))
        off += m["num_atoms"]
    e_all = np.concatenate(pieces, 1)
    hop_all = np.concatenate(hop_id)
    tgt = e_all[0] + hop_all * N
    src = e_all[1] + hop_all * N                     # "same indexing space as target" -> src % N recovers the atom
    layer = RefShell(D, D, num_hops=3, dropout=0.0, activation_type="silu", num_mlp_layers=2)
    layer.load_state_dict(det_state(shapes_of(layer), 11))
    x = torch.from_numpy(x0).requires_grad_(True)
    out = layer(x, torch.from_numpy(tgt), torch.from_numpy(src))
    (out * torch.from_numpy(R)).sum().backward()
    mp = layer.message_passing(torch.from_numpy(x0), torch.from_numpy(tgt), torch.from_numpy(src))
    save("layer_hopoffset", x=x0, R=R, target=tgt, src=src, out=out.detach().numpy(), gx=x.grad.numpy(),
         mp=np.stack([c.numpy() for c in mp]), **{"g_" + k: v for k, v in grads_of(layer).items()})

    # ---------------------------------------------------------------- 3. pooling layers
    Fdim = 64
    xp = rng.normal(0, 1, size=(N, Fdim)).astype(np.float32)
    Rp = rng.normal(0, 1, size=(6, Fdim)).astype(np.float32)
    Ra = rng.normal(0, 1, size=(4, N)).astype(np.float32)
    bi = batch.batch_indices
    pool = MultiHeadAttentionPoolingLayer(Fdim, num_heads=4, initial_temperature=1.0)
    pool.load_state_dict(det_state(shapes_of(pool), 12))
    x = torch.from_numpy(xp).requires_grad_(True)
    pooled, attn = pool(x, bi)
    ((pooled * torch.from_numpy(Rp)).sum() + 0.3 * (attn * torch.from_numpy(Ra)).sum()).backward()
    save("pool_attention", x=xp, Rp=Rp, Ra=Ra, batch_indices=bi.numpy(), pooled=pooled.detach().numpy(),
         attn=attn.detach().numpy(), gx=x.grad.numpy(), **{"g_" + k: v for k, v in grads_of(pool).items()})
    xq = np.round(xp * 2) / 2                          # coarse values -> ties exercise the first-maximum rule
    for name, cls in (("mean", MeanPoolingLayer), ("max", MaxPoolingLayer), ("sum", SumPoolingLayer)):
        x = torch.from_numpy(xq.astype(np.float32)).requires_grad_(True)
        pooled, none = cls()(x, bi)
        assert none is None
        (pooled * torch.from_numpy(Rp)).sum().backward()
        save(f"pool_{name}", x=xq.astype(np.float32), Rp=Rp, batch_indices=bi.numpy(), pooled=pooled.detach().numpy(),
             gx=x.grad.numpy())

    # ---------------------------------------------------------------- 4. whole model
    def run_model(name, cfg, batch, seed, store_full_grads=True, T=3, adam=False):
        model = RefGNN(FEATURE_SIZES, cfg["hidden_dim"], T, num_shells=cfg["num_shells"],
                       num_message_passing_layers=cfg["num_message_passing_layers"], dropout=0.0,
                       ffn_num_layers=cfg.get("ffn_num_layers", 3), pooling_type=cfg.get("pooling_type", "attention"),
                       task_type="multitask", embedding_dim=cfg.get("embedding_dim", 64),
                       use_partial_charges=cfg.get("use_partial_charges", False),
                       use_stereochemistry=cfg.get("use_stereochemistry", False), ffn_dropout=0.0,
                       activation_type=cfg.get("activation_type", "silu"),
                       shell_conv_num_mlp_layers=cfg.get("shell_conv_num_mlp_layers", 2), shell_conv_dropout=0.0,
                       attention_num_heads=cfg.get("attention_num_heads", 4))
        model.load_state_dict(det_state(shapes_of(model), seed))
        model.train()
        arrs = batch_to_arrays(batch)
        targets = batch.targets[:, :T].contiguous()
        arrs["targets"] = targets.numpy()
        weights = torch.linspace(0.5, 1.5, T)
        out, attn, q = model(batch.atom_features_map, batch.multi_hop_edge_indices, batch.batch_indices,
                             batch.total_charges, batch.final_tetrahedral_chiral_tensor, batch.final_cis_tensor,
                             batch.final_trans_tensor)
        loss = RefL1(weights)(out, targets)
        loss.backward()
        g = grads_of(model)
        extra = {}
        if store_full_grads:
            extra.update({"g_" + k: v for k, v in g.items()})
        extra.update({"gn_" + k: np.float64(np.linalg.norm(v.astype(np.float64))) for k, v in g.items()})
        if attn is not None:
            extra["attn"] = attn.detach().numpy()
        if q is not None:
            extra["q"] = q.detach().numpy()
        if adam:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
            opt = torch.optim.Adam(model.parameters(), lr=2.5e-4)
            opt.step()
            extra.update({"p1_" + k: v.detach().numpy() for k, v in model.state_dict().items()})
        save(name, seed=seed, T=T, loss_weights=weights.numpy(), out=out.detach().numpy(), loss=loss.item(),
             cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array([str(v) for v in cfg.values()]), **arrs, **extra)

    small = dict(hidden_dim=64, num_shells=3, num_message_passing_layers=2)
    run_model("gnn_small", small, S.make_batch(201, 8, 3, "qm9"), seed=21, adam=True)
    run_model("gnn_small_gelu_mean", dict(small, activation_type="gelu", pooling_type="mean"),
              S.make_batch(202, 5, 3, "qm9"), seed=22)
    run_model("gnn_stereo_charges", dict(small, use_partial_charges=True, use_stereochemistry=True),
              S.make_batch(203, 4, 3, "drug", stereo=True), seed=23)
    b_ns = S.make_batch(204, 3, 3, "qm9")              # stereo enabled but no stereo centres in the batch
    run_model("gnn_stereo_empty", dict(small, use_partial_charges=True, use_stereochemistry=True), b_ns, seed=24)
    run_model("gnn_h4_l3", dict(hidden_dim=96, num_shells=4, num_message_passing_layers=3, attention_num_heads=2),
              S.make_batch(205, 6, 4, "qm9"), seed=25)
    run_model("gnn_default", dict(hidden_dim=512, num_shells=3, num_message_passing_layers=3),
              S.make_batch(206, 4, 3, "qm9"), seed=26, store_full_grads=False, T=12)


if __name__ == "__main__":
    main()
