"""CPU, gloo, world size 2: the host-side logic of the N>1 path.

* helper collectives keep the reference's semantics (utils/distributed.py:49-228);
* batch sharding is contiguous ceil(n/W) ranges (molecular.py:228-237, pipeline.py:282-310);
* the gradient exchange = ONE all-reduce(SUM) over a flat arena followed by a 1/W scale reproduces
  DistributedDataParallel's mean-of-per-rank-means, checked against the single-process oracle gradient of
  the concatenated batch when both ranks hold equally sized batches.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model_port as MP
from oracle.fixtures import FEATURE_SIZES, det_state


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rand_batch(rng, n_mol, atoms_per_mol):
    """Tiny chain molecules: hop-1 edges only, both directions (enough to drive gradients)."""
    N = n_mol * atoms_per_mol
    feats = {k: torch.from_numpy(rng.integers(0, v, size=N)) for k, v in FEATURE_SIZES.items()}
    e = []
    for m in range(n_mol):
        o = m * atoms_per_mol
        for a in range(atoms_per_mol - 1):
            e += [(o + a, o + a + 1), (o + a + 1, o + a)]
    e.sort()
    empty = torch.empty((0, 2), dtype=torch.long)
    return dict(atom_features_map=feats, multi_hop_edge_indices=torch.tensor(e, dtype=torch.long),
                batch_indices=torch.arange(n_mol).repeat_interleave(atoms_per_mol), total_charges=torch.zeros(n_mol),
                final_tetrahedral_chiral_tensor=torch.empty((0, 4), dtype=torch.long), final_cis_tensor=empty,
                final_trans_tensor=empty, targets=torch.from_numpy(rng.normal(size=(n_mol, 2)).astype(np.float32)))


def _cat_batches(a, b):
    na = int(a["batch_indices"].shape[0])
    ma = int(a["total_charges"].shape[0])
    out = dict(a)
    out["atom_features_map"] = {k: torch.cat([a["atom_features_map"][k], b["atom_features_map"][k]]) for k in FEATURE_SIZES}
    out["multi_hop_edge_indices"] = torch.cat([a["multi_hop_edge_indices"], b["multi_hop_edge_indices"] + na])
    out["batch_indices"] = torch.cat([a["batch_indices"], b["batch_indices"] + ma])
    out["total_charges"] = torch.cat([a["total_charges"], b["total_charges"]])
    out["targets"] = torch.cat([a["targets"], b["targets"]])
    return out


CFG = dict(hidden_dim=32, num_shells=2, num_message_passing_layers=1)


def _shapes():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import gnn_shapes
    return gnn_shapes(CFG, 2)


def _flat_grad(P, batch):
    for v in P.values():
        v.grad = None
    out, _, _, _ = MP.gnn_forward(P, CFG, batch)
    MP.weighted_l1(out, batch["targets"], torch.ones(2)).backward()
    return torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for v in P.values()])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from aimnet_x2d_b200 import distributed as D
    device, is_ddp, local_rank, ws = D.setup_distributed_environment(backend="gloo")
    try:
        assert is_ddp and ws == world and local_rank == rank and D.safe_get_rank() == rank and D.get_world_size() == world
        assert D.is_main_process() == (rank == 0)
        # gather_ndarray_to_rank0: ragged first dimension, rank 0 gets the concatenation, others an empty array
        arr = np.arange((rank + 2) * 3, dtype=np.float32).reshape(rank + 2, 3) + 100 * rank
        g = D.gather_ndarray_to_rank0(arr)
        if rank == 0:
            exp = np.concatenate([np.arange((r + 2) * 3, dtype=np.float32).reshape(r + 2, 3) + 100 * r for r in range(world)])
            assert np.array_equal(g, exp)
        else:
            assert g.size == 0
        # device-resident variant (evaluation path): same result, dtype preserved (no float cast), tensors in / out
        tg = D.gather_tensor_to_rank0(torch.from_numpy(arr).to(torch.float64))
        ti = D.gather_tensor_to_rank0(torch.arange(rank + 1, dtype=torch.int64) + 10 * rank)
        if rank == 0:
            assert tg.dtype == torch.float64 and np.array_equal(tg.numpy(), exp.astype(np.float64))
            assert ti.tolist() == [0, 10, 11]
        else:
            assert tg.shape == (0, 3) and ti.numel() == 0
        # per-rank sharded output files, merged by rank 0 in rank order (replaces pickle merging / padded gathers)
        import tempfile
        from aimnet_x2d_b200.inference import ShardedOutputWriter, merge_rank_outputs
        out_dir = D.broadcast_object(tempfile.mkdtemp() if rank == 0 else None)
        wtr = ShardedOutputWriter(out_dir, rank, world)
        for b in range(2):                                   # two batches per rank: 2 + rank molecules each
            n_mol = 2 + rank
            wtr.append("outputs", np.full((n_mol, 3), 100 * rank + b, dtype=np.float32))
            counts = np.arange(1, n_mol + 1)
            wtr.append_ragged("partial_charges", np.arange(counts.sum(), dtype=np.float32) + 1000 * rank, counts)
        wtr.close()
        D.barrier()
        if rank == 0:
            m = merge_rank_outputs(out_dir, world)
            assert m["outputs"].shape == (2 * 2 + 2 * 3, 3)
            assert m["outputs"][:, 0].tolist() == [0, 0, 1, 1, 100, 100, 100, 101, 101, 101]
            ptr = m["partial_charges_ptr"]
            assert ptr.tolist() == [0, 1, 3, 4, 6, 7, 9, 12, 13, 15] and len(m["partial_charges"]) == 18
            assert m["partial_charges"][ptr[4]] == 1000.0          # first atom of rank 1's first molecule
        D.barrier()
        s = D.gather_strings_to_rank0([f"r{rank}a", f"r{rank}b"])
        assert s == (["r0a", "r0b", "r1a", "r1b"] if rank == 0 else [])
        assert D.broadcast_object({"best": 1.5} if rank == 0 else None) == {"best": 1.5}
        t = D.all_reduce_tensor(torch.tensor([1.0 + rank, 2.0]), op="mean")
        assert torch.allclose(t, torch.tensor([1.5, 2.0]))
        with pytest.raises(ValueError):
            D.all_reduce_tensor(torch.zeros(1), op="prod")
        D.barrier()
        # the gradient exchange: one SUM all-reduce over the flat arena, then 1/W  (== DDP averaging)
        rng = np.random.Generator(np.random.PCG64(50))
        batches = [_rand_batch(rng, 3, 4) for _ in range(world)]
        P = {k: v.requires_grad_(True) for k, v in det_state(_shapes(), 3).items()}
        flat = _flat_grad(P, batches[rank]).clone()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= world
        if rank == 0:
            ref = _flat_grad(P, _cat_batches(batches[0], batches[1]))     # mean over the global batch
            q.put(float((flat - ref).abs().max() / ref.abs().max()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_helpers_and_gradient_exchange():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() < 1e-5


def test_shard_indices_cover_everything_once():
    from aimnet_x2d_b200.distributed import shard_indices
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert np.array_equal(np.concatenate(parts), np.arange(n))
            per = (n + w - 1) // w
            assert all(len(p) <= per for p in parts)


def test_helpers_degrade_without_process_group():
    """distributed.py:20,30-33,61-62,...: every helper is a no-op when torch.distributed is not initialised."""
    from aimnet_x2d_b200 import distributed as D
    a = np.arange(4.0)
    assert D.gather_ndarray_to_rank0(a) is a and D.gather_strings_to_rank0(["x"]) == ["x"]
    assert D.broadcast_object(3) == 3 and D.is_main_process() and D.safe_get_rank() == 0 and D.get_world_size() == 1
    t = torch.ones(2)
    assert D.all_reduce_tensor(t) is t
    D.barrier()
