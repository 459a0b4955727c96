"""Worker of tests/test_gpu_multi.py (one process per GPU under torch.distributed.run): the data-parallel training step of
the CUDA path -- FlatAdam (rank-0 parameter broadcast, NCCL all-reduce of the flat gradient arena, 1/W folded into the fused
clip + Adam) under GraphedTrainStep -- against the single-process CPU oracle on the concatenation of the per-rank batches
(DDP semantics: mean of per-rank mean losses, src/main/runner.py:703-707)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main(out_path):
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from aimnet_x2d_b200.collate import pad_batch
    from helpers import gnn_shapes
    from oracle import model_port as MP
    from oracle.fixtures import FEATURE_SIZES, det_state
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(hidden_dim=128, num_shells=3, num_message_passing_layers=2)
    T, B = 4, 48
    P = det_state(gnn_shapes(cfg, T), 7)
    model = ax.GNN(FEATURE_SIZES, 128, T, num_shells=3, num_message_passing_layers=2, task_type="multitask",
                   shell_conv_dropout=0.0, ffn_dropout=0.0)
    model.load_state_dict(P)
    if rank != 0:                      # replicas start DIFFERENT: FlatAdam must take rank 0's parameters (what DDP does)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.01 * (rank + 1))
    model.to(dev).train()
    opt = ax.FlatAdam(model.parameters(), lr=2.5e-4, max_grad_norm=1.0)
    w = torch.linspace(0.5, 1.5, T)
    crit = ax.WeightedL1Loss(w).to(dev)
    batches = [S.make_batch(900 + r, B, 3, "qm9", T) for r in range(world)]
    n_pad = max(b.graph_index.num_atoms for b in batches) + 64
    e_cap = max(b.graph_index.num_edges for b in batches) + 64
    padded = [pad_batch(b, (n_pad + 127) // 128 * 128, (e_cap + 1023) // 1024 * 1024, 8) for b in batches]
    t_cap = max(p.graph_index.n_tiles for p in padded) + 2
    me_cap = max(p.graph_index.max_tile_edges for p in padded) + 64
    padded = [pad_batch(b, (n_pad + 127) // 128 * 128, (e_cap + 1023) // 1024 * 1024, 8, t_cap, max_tile_edges=me_cap) for b in batches]
    step = ax.GraphedTrainStep(model, crit, opt, dev)
    loss = step(padded[rank].pin_memory())
    torch.cuda.synchronize()
    grad = (opt.flat_grad / world).cpu()              # the arena holds the SUM over ranks; 1/W is applied inside the update
    after = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, torch.tensor([loss], device=dev))
    # every rank must hold identical parameters after the step
    flat = opt.flat_param.clone()
    ref0 = flat.clone()
    dist.broadcast(ref0, src=0)
    same = bool(torch.equal(flat, ref0))
    res = {"rank": rank, "same_params": same}
    if rank == 0:
        Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        total = 0.0
        for b in batches:                               # mean over ranks of per-rank mean losses
            ob = dict(atom_features_map=b.atom_features_map, multi_hop_edge_indices=b.multi_hop_edge_indices,
                      batch_indices=b.batch_indices, total_charges=b.total_charges,
                      final_tetrahedral_chiral_tensor=b.final_tetrahedral_chiral_tensor, final_cis_tensor=b.final_cis_tensor,
                      final_trans_tensor=b.final_trans_tensor)
            out, _, _, _ = MP.gnn_forward(Pr, cfg, ob)
            l = MP.weighted_l1(out, b.targets, w) / world
            l.backward()
            total += float(l)
        keys = [k for k in Pr if Pr[k].grad is not None]
        worst = 0.0
        off = dict(zip([id(p) for p in opt.params], opt.offsets))
        names = {id(p): k for k, p in model.named_parameters()}
        for p in opt.params:
            k = names[id(p)]
            ref = Pr[k].grad.numpy() if Pr[k].grad is not None else np.zeros(tuple(p.shape), np.float32)
            got = grad[off[id(p)]: off[id(p)] + p.numel()].view(p.shape).numpy()
            if "attention_weights." in k and k.endswith(".bias"):
                continue
            scale = float(np.abs(ref).max())
            if scale > 0:
                worst = max(worst, float(np.abs(got - ref).max()) / scale)
        state = dict(m=[torch.zeros_like(Pr[k]) for k in keys], v=[torch.zeros_like(Pr[k]) for k in keys])
        with torch.no_grad():
            norm = MP.clip_and_adam([Pr[k] for k in keys], [Pr[k].grad for k in keys], state, step=1)
        upd = 0.0
        for k in keys:
            if "attention_weights." in k and k.endswith(".bias"):
                continue
            g = np.abs(Pr[k].grad.numpy())
            firm = g >= 1e-4 * max(float(g.max()), 1e-30)
            d = np.abs(after[k].astype(np.float64) - Pr[k].detach().numpy())
            upd = max(upd, float(np.where(firm, d, 0.0).max()) / max(float(np.abs(Pr[k].detach().numpy()).max()), 1e-30))
        res.update(worst_grad_rel=worst, worst_param_rel=upd, loss_mean=float(sum(float(x) for x in losses) / world),
                   oracle_loss=total, grad_norm=float(opt.grad_norm()), oracle_norm=norm)
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump(gathered, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
