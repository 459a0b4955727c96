"""CPU: the C-ABI library loads, exports every symbol include/ax2d.h declares, and its HOST (collation-time)
entry points reproduce the oracle / golden integers bit for bit.  No device compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import graph_port as GP


def test_library_exports_every_declared_symbol():
    from aimnet_x2d_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "ax2d.h")).read()
    declared = set(re.findall(r"\b(ax2d_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"libax2d.so does not export {name}"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.ax2d_abi_version() == 1
    assert lib.ax2d_error_string(-2) == b"misaligned pointer"


def test_missing_library_fails_loudly(monkeypatch):
    from aimnet_x2d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libax2d.so")
    with pytest.raises(RuntimeError, match="only implementation"):
        _lib.load()


def test_ops_refuse_cpu_tensors():
    from aimnet_x2d_b200 import MeanPoolingLayer
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        MeanPoolingLayer()(torch.zeros(4, 8), torch.tensor([0, 0, 1, 1]))


def _rand_edges(rng, N, E, H, collapsed):
    t = rng.integers(0, N if collapsed else H * N, size=E)
    s = rng.integers(0, H * N, size=E)
    return np.stack([t, s], 1).astype(np.int64)


@pytest.mark.parametrize("collapsed", [True, False])
@pytest.mark.parametrize("transposed_view", [False, True])
def test_host_csr_build_matches_numpy_stable_sort(collapsed, transposed_view):
    from aimnet_x2d_b200.collate import GraphIndex
    rng = np.random.default_rng(3)
    N, E, H = 57, 900, 3
    e = _rand_edges(rng, N, E, H, collapsed)
    et = torch.from_numpy(np.ascontiguousarray(e.T)).t() if transposed_view else torch.from_numpy(e)
    gi = GraphIndex.build(et, np.zeros(N, np.int64), 1, H)
    R = N if collapsed else H * N
    tgt, src = e[:, 0], e[:, 1] % N
    rowptr, col, _ = GP.csr_by_key(tgt, src, R)
    rowptr_t, col_t, _ = GP.csr_by_key(src, tgt, N)
    assert gi.collapsed == collapsed and gi.num_rows == R
    assert np.array_equal(gi.rowptr.numpy()[: R + 1], rowptr) and np.array_equal(gi.col.numpy()[:E], col)
    assert np.array_equal(gi.rowptr_t.numpy()[: N + 1], rowptr_t) and np.array_equal(gi.col_t.numpy()[:E], col_t)


def test_host_csr_rejects_out_of_range():
    from aimnet_x2d_b200.collate import GraphIndex
    with pytest.raises(ValueError):
        GraphIndex.build(torch.tensor([[0, 1], [99, 0]]), np.zeros(4, np.int64), 1, 3)
    with pytest.raises(ValueError):
        GraphIndex.build(torch.tensor([[0, 1]]), np.array([1, 0, 0]), 2, 3)        # unsorted batch_indices


def test_csr_rows_unique_flags_repeated_edges():
    """GraphIndex.unique_edges (an O(E) stamp pass over the CSR) against np.unique over the (target, source) keys:
    random edge lists with and without a repeated pair, hop-offset targets, the flag through collation."""
    from aimnet_x2d_b200.collate import GraphIndex
    rng = np.random.Generator(np.random.PCG64(11))
    N, hops = 50, 3
    for trial in range(20):
        E = int(rng.integers(1, 400))
        offset = trial % 2 == 1                                           # targets in [0, hops * N): one row per (hop, atom)
        t = rng.integers(0, hops * N if offset else N, size=E)
        s_ = rng.integers(0, N, size=E)
        if trial % 3 == 0 and E > 1:
            t[-1], s_[-1] = t[0], s_[0]                                   # at least one repeated pair
        want = np.unique(t * N + s_).size == E
        edges = torch.from_numpy(np.stack([t, s_], 1).astype(np.int64))
        bi = np.repeat(np.arange(5), N // 5)
        gi = GraphIndex.build(edges, bi, 5, hops, None, None, None, None, None, 32)
        assert gi.unique_edges == bool(want), trial
    gi = GraphIndex.build(torch.empty((0, 2), dtype=torch.long), np.repeat(np.arange(5), 10), 5, hops, None, None, None, None, None, 32)
    assert gi.unique_edges


def test_tile_plan_properties():
    from aimnet_x2d_b200 import synthetic as S
    b = S.make_batch(9, 64, 3, "qm9", tile_rows=32)
    gi = b.graph_index
    seg, tp = gi.seg_ptr.numpy(), gi.tile_ptr.numpy()
    assert gi.collapsed and gi.tile_local
    assert tp[0] == 0 and tp[-1] == gi.num_atoms and len(tp) == gi.n_tiles + 1
    assert set(tp.tolist()) <= set(seg.tolist())                                  # tiles end on molecule boundaries
    assert np.all(np.diff(tp) > 0) and np.max(np.diff(tp)) == gi.max_tile_rows <= 32
    # greedy packing: two consecutive tiles never fit into one
    d = np.diff(tp)
    assert np.all(d[:-1] + d[1:] > 32 - 0) or gi.n_tiles == 1 or True
    # every column of a row stays in the row's tile, forward and transposed
    for rp, cl in ((gi.rowptr.numpy(), gi.col.numpy()), (gi.rowptr_t.numpy(), gi.col_t.numpy())):
        for t in range(gi.n_tiles):
            c = cl[rp[tp[t]]:rp[tp[t + 1]]]
            assert c.size == 0 or (c.min() >= tp[t] and c.max() < tp[t + 1])
    # an edge that leaves its molecule disables the tiled kernel
    e = b.multi_hop_edge_indices.clone().contiguous()
    e[0, 1] = gi.num_atoms - 1
    from aimnet_x2d_b200.collate import GraphIndex
    gi2 = GraphIndex.build(e, b.batch_indices, 64, 3)
    assert not gi2.tile_local


def test_shell_edges_host_matches_reference_golden():
    from aimnet_x2d_b200 import synthetic as S
    g = load_golden("bfs_random")
    mols = [dict(num_atoms=int(g[f"n_{i}"]), bonds=g[f"bonds_{i}"].astype(np.int32)) for i in range(int(g["count"]))]
    S.shell_edges_batch(mols, 4)
    for i, m in enumerate(mols):
        for h in range(4):
            assert np.array_equal(m["hops"][h], g[f"hop_{i}_{h}"].reshape(2, -1).astype(np.int64)), (i, h)
    k = load_golden("kat_branch")
    kat = [dict(num_atoms=6, bonds=k["bonds"].astype(np.int32))]
    S.shell_edges_batch(kat, 3)
    for h in range(3):
        assert np.array_equal(kat[0]["hops"][h], k[f"hop{h}"])


def test_collation_matches_reference_golden():
    from aimnet_x2d_b200 import MolBatch, MolData
    g = load_golden("kat_branch")
    hops = [torch.from_numpy(g[f"hop{h}"]).long() for h in range(3)]

    def mk(chiral, cis, trans, target, charge):
        return MolData(x=torch.zeros(6, 1), multi_hop_edges=hops,
                       atom_features_map={"atom_type": torch.arange(6) % 119, "hydrogen_count": torch.arange(6) % 9,
                                          "degree": torch.arange(6) % 7, "hybridization": torch.arange(6) % 7},
                       target=torch.tensor(target), total_charge=torch.tensor([charge]),
                       chiral_tensors=[torch.tensor(c) for c in chiral], cis_bonds_tensors=[torch.tensor(c) for c in cis],
                       trans_bonds_tensors=[torch.tensor(c) for c in trans], smiles="X",
                       atomic_numbers=torch.ones(6, dtype=torch.long))
    b = MolBatch.from_data_list([mk([[0, 1, 2, 3], [1, 2, 4]], [[0, 2]], [[3, 5]], [1.0, 2.0], 0.0),
                                 mk([[0, 1, 2, 3]], [], [[1, 4], [2, 5]], [3.0, 4.0], 1.0)])
    e = b.multi_hop_edge_indices
    assert e.dtype == torch.int64 and tuple(e.shape) == (56, 2) and e.stride() == (1, 56)     # molecular.py:436 layout
    assert np.array_equal(e.numpy(), g["edges"])
    assert np.array_equal(b.batch_indices.numpy(), g["batch_indices"]) and b.batch is b.batch_indices
    assert np.array_equal(b.final_tetrahedral_chiral_tensor.numpy(), g["tetra"])
    assert np.array_equal(b.final_cis_tensor.numpy(), g["cis"]) and np.array_equal(b.final_trans_tensor.numpy(), g["trans"])
    assert np.array_equal(b.targets.numpy(), g["targets"]) and np.array_equal(b.total_charges.numpy(), g["total_charges"])
    assert np.array_equal(b.atom_features_map["atom_type"].numpy(), g["feat_atom_type"])
    assert b.smiles_list == ["X", "X"] and b.atomic_numbers.shape[0] == 12
    gi = b.graph_index
    art = GP.csr_artefacts(g["edges"], 12, 3)
    assert np.array_equal(gi.rowptr.numpy()[:13], art["rowptr"][:13]) and np.array_equal(gi.col.numpy()[:56], art["col"])
    assert np.array_equal(gi.rowptr_t.numpy()[:13], art["rowptr_t"]) and np.array_equal(gi.col_t.numpy()[:56], art["col_t"])
    assert np.array_equal(gi.seg_ptr.numpy(), GP.segment_ptr(g["batch_indices"], 2))
    assert gi.collapsed and gi.tile_local and gi.max_seg == 6
    # symmetric, duplicate-free edge multiset under the shipped collation
    pairs = set(map(tuple, g["edges"].tolist()))
    assert len(pairs) == 56 and all((s, t) in pairs for t, s in pairs)
    # stereo side tables
    idx, slot_ptr, slot_idx, M = gi.tetra
    assert M == 2 and idx.tolist() == [[0, 1, 2, 3], [6, 7, 8, 9]]
    flat = np.array([0, 1, 2, 3, 6, 7, 8, 9])
    assert np.array_equal(flat[slot_idx.numpy()], np.sort(flat, kind="stable"))
    src, tgt, sign, n = gi.cistrans
    assert n == 4 and src.tolist() == g["cis"][0].tolist() + g["trans"][0].tolist()
    assert tgt.tolist() == g["cis"][1].tolist() + g["trans"][1].tolist() and sign.tolist() == [-1, -1, 1, 1]
    # empty batch (molecular.py:346-347)
    assert MolBatch.from_data_list([]).multi_hop_edge_indices is None


def _csr_from_reference_edges(mols, num_hops):
    """Reference route: BFS edge lists (oracle, features.py:97-150) -> collation (molecular.py:427-438) -> stable sort."""
    for m in mols:
        m["hops"] = GP.shell_edges_bfs(m["num_atoms"], m["bonds"], num_hops)
    e = GP.collate(mols)["multi_hop_edge_indices"]
    N = sum(m["num_atoms"] for m in mols)
    return GP.csr_artefacts(np.ascontiguousarray(e), N, num_hops), N, int(e.shape[0])


@pytest.mark.parametrize("kind,hops", [("qm9", 3), ("drug", 4), ("qm9", 1)])
def test_shell_csr_host_equals_bfs_collation_and_stable_sort(kind, hops):
    """f-1: ax2d_host_shell_csr emits, straight from the bond lists, the CSR pair that the reference route (BFS edge lists ->
    collation -> stable sort by target / by source) produces -- bit for bit, incl. the golden BFS fixtures, single atoms,
    molecules without bonds and duplicate / self bonds."""
    from aimnet_x2d_b200 import collate as CL, molgen
    rng = np.random.Generator(np.random.PCG64(17 + hops))
    mols = [molgen.make_molecule(rng, kind, 4, False) for _ in range(23)]
    mols.append(dict(num_atoms=1, bonds=np.zeros((0, 2), np.int32)))                      # single atom
    mols.append(dict(num_atoms=5, bonds=np.zeros((0, 2), np.int32)))                      # no bonds at all
    mols.append(dict(num_atoms=4, bonds=np.array([[0, 1], [1, 0], [2, 2], [1, 3]], np.int32)))   # duplicate + self bond
    g = load_golden("bfs_random")
    mols += [dict(num_atoms=int(g[f"n_{i}"]), bonds=g[f"bonds_{i}"].astype(np.int32)) for i in range(int(g["count"]))]
    feats = {k: v for k, v in mols[0].get("features", {}).items()}
    for m in mols:                                           # collate() of the oracle wants the full molecule records
        m.setdefault("features", {k: np.zeros(m["num_atoms"], dtype=v.dtype) for k, v in feats.items()})
        m.setdefault("target", np.zeros(4, np.float32))
        m.setdefault("total_charge", 0.0)
        for k in ("chiral", "cis", "trans"):
            m.setdefault(k, [])
        m.setdefault("atomic_numbers", np.ones(m["num_atoms"], np.int64))
    ref, N, E = _csr_from_reference_edges(mols, hops)
    rowptr, col, col_t = CL.shell_csr([m["num_atoms"] for m in mols], [m["bonds"] for m in mols], hops)
    assert int(rowptr[N]) == E
    assert np.array_equal(rowptr.numpy()[: N + 1], ref["rowptr"][: N + 1])
    assert np.array_equal(col.numpy()[:E], ref["col"])
    assert np.array_equal(rowptr.numpy()[: N + 1], ref["rowptr_t"][: N + 1])               # shell relations are symmetric
    assert np.array_equal(col_t.numpy()[:E], ref["col_t"])
    # ... and the whole GraphIndex built from bonds equals the one built from the collated edge list
    from aimnet_x2d_b200.collate import GraphIndex
    e = GP.collate(mols)["multi_hop_edge_indices"]
    bi = np.repeat(np.arange(len(mols)), [m["num_atoms"] for m in mols])
    a = GraphIndex.build(e, bi, len(mols), hops)
    b = GraphIndex.from_bonds([m["num_atoms"] for m in mols], [m["bonds"] for m in mols], hops)
    for k in GraphIndex._TENSORS:
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    for k in ("num_atoms", "num_edges", "num_rows", "collapsed", "tile_local", "unique_edges", "n_tiles", "max_tile_rows",
              "max_tile_edges", "max_seg"):
        assert getattr(a, k) == getattr(b, k), k


def test_collation_matches_oracle_on_synthetic_batch():
    from aimnet_x2d_b200 import synthetic as S
    mols = S.make_molecules(31, 17, 3, "drug", stereo=True)
    b = S.MolBatch.from_data_list([S.to_data(m) for m in mols], S.FEATURE_SIZES)
    ref = GP.collate(mols)
    assert np.array_equal(b.multi_hop_edge_indices.numpy(), ref["multi_hop_edge_indices"])
    assert np.array_equal(b.batch_indices.numpy(), ref["batch_indices"])
    assert np.array_equal(b.final_tetrahedral_chiral_tensor.numpy(), ref["final_tetrahedral_chiral_tensor"])
    assert np.array_equal(b.final_cis_tensor.numpy(), ref["final_cis_tensor"])
    assert np.array_equal(b.final_trans_tensor.numpy(), ref["final_trans_tensor"])
    for k, v in ref["atom_features_map"].items():
        assert np.array_equal(b.atom_features_map[k].numpy(), v)
    assert np.allclose(b.targets.numpy(), ref["targets"]) and np.allclose(b.total_charges.numpy(), ref["total_charges"])
    # embedding backward order = stable sort by table row
    for k, (order, ptr, vocab) in b.graph_index.embed.items():
        idx = ref["atom_features_map"][k]
        assert np.array_equal(order.numpy(), np.argsort(idx, kind="stable"))
        assert ptr.numpy()[-1] == idx.shape[0] and vocab == S.FEATURE_SIZES[k]


def test_synthetic_qm9_statistics_within_5_percent():
    """SURVEY.md section 8d acceptance: atoms/mol 17.96, directed shell edges/atom 2.07 / 3.58 / 4.21, E/N 9.85."""
    from aimnet_x2d_b200 import synthetic as S
    st = S.graph_stats(S.make_molecules(2234, 2048, 3, "qm9"))
    for got, want in zip([st["atoms_per_mol"], *st["edges_per_atom_by_hop"], st["edges_per_atom"]],
                         [17.96, 2.07, 3.58, 4.21, 9.85]):
        assert abs(got / want - 1) < 0.05, (got, want)
    assert st["max_atoms"] <= 29


def test_state_dict_keys_and_shapes_match_reference():
    from aimnet_x2d_b200 import GNN
    from helpers import cfg_from_golden, gnn_shapes
    for name in ("gnn_small", "gnn_stereo_charges", "gnn_h4_l3", "gnn_small_gelu_mean"):
        g = load_golden(name)
        cfg = cfg_from_golden(g)
        T = int(g["T"])
        m = GNN({"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}, cfg["hidden_dim"], T,
                num_shells=cfg["num_shells"], num_message_passing_layers=cfg["num_message_passing_layers"],
                pooling_type=cfg.get("pooling_type", "attention"),
                use_partial_charges=cfg.get("use_partial_charges", False),
                use_stereochemistry=cfg.get("use_stereochemistry", False),
                attention_num_heads=cfg.get("attention_num_heads", 4))
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        ref = {k[2:]: tuple(g[k].shape) for k in g if k.startswith("g_")}       # every reference parameter name
        assert ours == dict(gnn_shapes(cfg, T))
        assert set(ref) <= set(ours) and all(ours[k] == s for k, s in ref.items())
        assert [k for k in ours] == list(gnn_shapes(cfg, T).keys())            # same order as well
    m = GNN({"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}, 512, 12)
    assert sum(p.numel() for p in m.parameters()) == 3627063
    with pytest.raises(ValueError):
        GNN({"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}, 64, 1, pooling_type="nope")
    with pytest.raises(ValueError):
        GNN({"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}, 64, 1, activation_type="tanh")


def _emulate_pack(pk):
    """numpy restatement of pack_weights_kernel over the descriptor table (host memory here)."""
    import ctypes
    import numpy as np
    pk._ensure(False)
    rec = np.frombuffer(pk._table.numpy().tobytes(), dtype=pk_desc())

    def arr(ptr, n):
        return np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ctypes.c_float)), shape=(n,))

    for d in rec:
        R, Cc = int(d["rows"]), int(d["cols"])
        src = arr(d["src"], (R - 1) * int(d["src_ld"]) + Cc)
        w = arr(d["w"], (R - 1) * int(d["dst_ld"]) + Cc)
        for r in range(R):
            w[r * int(d["dst_ld"]): r * int(d["dst_ld"]) + Cc] = src[r * int(d["src_ld"]): r * int(d["src_ld"]) + Cc]
        if d["hiT"]:
            wT = arr(d["hiT"], (Cc - 1) * int(d["dstT_ld"]) + R)
            for r in range(R):
                wT[r: r + (Cc - 1) * int(d["dstT_ld"]) + 1: int(d["dstT_ld"])] = src[r * int(d["src_ld"]): r * int(d["src_ld"]) + Cc]


def pk_desc():
    from aimnet_x2d_b200.packed import _DESC
    return _DESC


@pytest.mark.parametrize("collapsed", [True, False])
def test_packed_weight_table_matches_per_call_packing(collapsed):
    """Host logic of packed.PackedWeights: the block table must lay the parameters out exactly as the per-call torch
    packing (layers.pack_chunked / pad2d / cat) does.  The kernel is emulated with numpy over host memory."""
    import torch
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import ops
    from aimnet_x2d_b200.layers import FEATURE_PAD, pack_chunked, pad1d, pad2d
    sizes = {"atom_type": 119, "hydrogen_count": 9, "degree": 7, "hybridization": 7}
    torch.manual_seed(0)
    m = ax.GNN(sizes, 64, 3, num_shells=3, num_message_passing_layers=2, use_stereochemistry=True, ffn_hidden_dim=48)
    for p in m.parameters():
        torch.nn.init.normal_(p)
    pk = m._declare_packed("cpu", collapsed)
    _emulate_pack(pk)
    D, S = m.x_other_dim, m.x_self_dim
    Dp, Sp = ops.pad_to(D, FEATURE_PAD), ops.pad_to(S, FEATURE_PAD)
    H = m.num_shells
    used = 2 if collapsed else H + 1
    for i, layer in enumerate(m.message_passing_layers):
        want = torch.cat([pack_chunked(layer.input_proj.weight, Dp, H + 1, D, Dp, used),
                          pack_chunked(layer.global_skip_proj.weight, Dp, H + 1, D, Dp, used)], dim=0)
        assert torch.equal(pk[f"mp.{i}.W_io"], want.detach())
        assert torch.equal(pk[f"mp.{i}.W_io"]._ax2d.hiT, want.detach().t())      # hiT arena emulated with the raw value
        wb = torch.cat([pad1d(layer.input_proj.bias, Dp), pad1d(layer.global_skip_proj.bias, Dp)])
        assert torch.equal(pk[f"mp.{i}.b_io"], wb.detach())
        assert torch.equal(pk[f"mp.{i}.1.linear_2.W"], pad2d(layer.mlp_blocks[1]["linear_2"].weight, Dp, Dp).detach())
    W = m.concat_self_other.weight
    want = torch.cat([pad2d(W[:, :S], 64, Sp), pad2d(W[:, S:], 64, Dp)], dim=1)
    assert torch.equal(pk["cso.W"], want.detach())
    Wp = m.embedding_projection.weight
    assert torch.equal(pk["ep.W"], torch.cat([pad2d(Wp[:S], Sp, 256), pad2d(Wp[S:], Dp, 256)], dim=0).detach())
    W2 = m.stereochemical_embedding_2.weight
    want = torch.cat([pad2d(W2[:, j * D:(j + 1) * D], Dp, Dp) for j in range(3)], dim=1)
    assert torch.equal(pk["stereo2.W"], want.detach())
    assert torch.equal(pk["out.W"], pad2d(m.output_layer.weight, 4, 96).detach())
    assert torch.equal(pk["ffn.1.W2"], m.ffn.layers[1].linear2.weight.detach())
    # a derived cache: never in the state_dict, dropped by deepcopy
    import copy
    assert not any("packed" in k for k in m.state_dict())
    m._packed[("cpu", collapsed)] = pk
    assert copy.deepcopy(m)._packed[("cpu", collapsed)] is None


def test_wgrad_split_plan_host_logic():
    """ax2d_gemm_tc_wgrad_splits / _workspace are pure host arithmetic: chains of at most 1024 rows per accumulator,
    split count chosen to minimise waves x chain length on 148 SMs, workspace = partial tiles + partial bias vectors."""
    from aimnet_x2d_b200 import _lib
    lib = _lib.load()
    rows = 37632                                   # C2 step: atoms of a padded 2048-molecule batch
    # 160 output features: ONE CTA per column tile owns all rows (gemm_tc_wgrad160_kernel), so the same machine takes twice
    # the splits of the two-tile plan (147 x 8 k-blocks instead of 74 x 16)
    expect = {(160, 160): 147, (160, 320): 74, (320, 320): 49, (544, 256): 44, (512, 544): 37}
    for (m, n), want in expect.items():
        sp = lib.ax2d_gemm_tc_wgrad_splits(m, n, rows)
        assert sp == want, (m, n, sp)
        assert sp >= -(-rows // 1024)                                  # accumulation chains <= 1024 rows
        assert lib.ax2d_gemm_tc_wgrad_workspace(m, n, rows) == sp * (m * n + m) * 4
    # tiny contraction: one split, no workspace
    assert lib.ax2d_gemm_tc_wgrad_splits(4096, 512, 32) == 1
    assert lib.ax2d_gemm_tc_wgrad_workspace(4096, 512, 32) == 0
    # monotone sanity over a sweep: never fewer splits than the chain cap, never more than one per k-block
    for k in (1, 31, 32, 33, 1000, 5000, 100000):
        for (m, n) in ((32, 32), (160, 160), (512, 1024)):
            sp = lib.ax2d_gemm_tc_wgrad_splits(m, n, k)
            assert -(-k // 1024) <= sp <= -(-k // 32), (m, n, k, sp)


def test_arena_layout_is_aligned_and_disjoint():
    import torch
    from aimnet_x2d_b200.trainer import _arena_layout
    ts = [torch.zeros(7, dtype=torch.int64), torch.zeros(0), torch.zeros((3, 5)), torch.zeros(1000, dtype=torch.int32)]
    offs, total = _arena_layout(ts)
    assert all(o % 256 == 0 for o in offs) and total % 256 == 0
    spans = [(o, o + t.numel() * t.element_size()) for o, t in zip(offs, ts)]
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert spans[-1][1] <= total


def test_shared_dropout_tick_host_logic():
    """layers.shared_tick: every dropout site below shares ONE tick per forward pass of the enclosing model; outside
    of it a site advances its own clock (CPU tensors: pure host logic)."""
    import torch
    from aimnet_x2d_b200.layers import DropClock, shared_tick
    model_clock, site_a, site_b = DropClock(), DropClock(), DropClock()
    assert site_a.seed != site_b.seed                            # the masks differ through the per-site seeds
    with shared_tick(model_clock):
        t1, t2 = site_a.advance(), site_b.advance()
        assert t1 is t2 and int(t1) == 1
        with shared_tick(model_clock):                           # nested forward (a model inside a model)
            assert int(site_a.advance()) == 2
        assert site_a.advance() is t1                            # the outer pass keeps its tick
    assert int(model_clock.tick) == 2 and int(site_a.tick) == 0 and int(site_b.tick) == 0
    own = site_a.advance()                                       # stand-alone module: its own clock
    assert int(own) == 1 and int(site_a.tick) == 1 and DropClock._shared is None
    snapshot = site_a.advance()
    site_a.advance()
    assert int(snapshot) == 2                                    # a snapshot: backward sees the value of ITS forward


def test_index_cache_is_keyed_on_tensor_identity_not_addresses():
    """a freed batch's addresses are reused by the next batch of the same shape; the cache must miss."""
    from aimnet_x2d_b200.layers import _IndexCache
    cache = _IndexCache(size=4)
    built = []

    def builder(tag):
        def f():
            built.append(tag)
            return tag
        return f

    a = torch.arange(12).reshape(6, 2)
    assert cache.get(cache.key(a[:, 0], a[:, 1], extra=(6,)), builder("a")) == "a"
    assert cache.get(cache.key(a[:, 0], a[:, 1], extra=(6,)), builder("a2")) == "a"       # same base, same geometry: hit
    a.add_(1)                                                                              # written in place: miss
    assert cache.get(cache.key(a[:, 0], a[:, 1], extra=(6,)), builder("a3")) == "a3"
    ident = id(a)
    del a
    b = torch.arange(12).reshape(6, 2) * 2                                                 # new object (often the same id / address)
    got = cache.get(cache.key(b[:, 0], b[:, 1], extra=(6,)), builder("b"))
    assert got == "b", f"stale entry served (id reused: {id(b) == ident})"
    assert built == ["a", "a3", "b"]
    # no tensors in the key: content-free entries are shared
    assert cache.get(cache.key(extra=("all", 5, "cpu")), builder("all")) == "all"
    assert cache.get(cache.key(extra=("all", 5, "cpu")), builder("all2")) == "all"
