"""Development tool: error of the bf16 CUDA path against the fp32 golden vectors, next to the error of the reference's op
sequence under torch.autocast('cpu', bfloat16) (oracle port) on the same fixtures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from conftest import load_golden  # noqa: E402
from helpers import cfg_from_golden, gnn_shapes  # noqa: E402
from oracle import model_port as MP  # noqa: E402
from oracle.fixtures import batch_to_torch, det_state  # noqa: E402
import test_gpu_bf16 as TB  # noqa: E402


def rel(a, ref):
    a = np.asarray(a, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


for name in ([a for a in sys.argv[1:] if not a.startswith("--")] or (["gnn_small", "gnn_default", "gnn_h4_l3"] if "--c4" not in sys.argv else [])):
    g = load_golden(name)
    cfg = cfg_from_golden(g); T = int(g["T"])
    P = {k: v.requires_grad_(True) for k, v in det_state(gnn_shapes(cfg, T), int(g["seed"])).items()}
    b = batch_to_torch(g)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out, attn, q, _ = MP.gnn_forward(P, cfg, b)
        loss = MP.weighted_l1(out.float(), b["targets"], torch.from_numpy(g["loss_weights"]))
    (out.float() * torch.from_numpy(TB._upstream(g))).sum().backward()
    model, _, _ = TB._model(g)
    model.compute_dtype = torch.bfloat16
    o2, a2, q2, l2 = TB._run(model, g)
    (o2 * torch.from_numpy(TB._upstream(g)).cuda()).sum().backward()
    print(f"== {name}: out ours {rel(o2.detach().cpu().numpy(), g['out']):.4f} autocast {rel(out.detach().float().numpy(), g['out']):.4f} | "
          f"loss ours {abs(float(l2) - float(g['loss'])) / abs(float(g['loss'])):.4f} autocast {abs(float(loss) - float(g['loss'])) / abs(float(g['loss'])):.4f}"
          + (f" | attn ours {rel(a2.detach().float().cpu().numpy(), g['attn']):.4f} autocast {rel(attn.detach().float().numpy(), g['attn']):.4f}" if "attn" in g else ""))
    rows = []
    for k, p in model.named_parameters():
        if "attention_weights." in k and k.endswith(".bias"):
            continue
        if "g_" + k in g:
            ours = rel(p.grad.cpu().numpy() if p.grad is not None else np.zeros(p.shape), g["g_" + k])
            ac = rel(P[k].grad.numpy() if P[k].grad is not None else np.zeros(p.shape), g["g_" + k])
        else:
            rn = float(g["gn_" + k])
            if rn == 0:
                continue
            ours = abs(float(p.grad.double().norm()) - rn) / rn if p.grad is not None else 1.0
            ac = abs(float(P[k].grad.double().norm()) - rn) / rn if P[k].grad is not None else 1.0
        rows.append((ours, ac, k))
    rows.sort(reverse=True)
    for ours, ac, k in rows[:12]:
        print(f"   {k:60s} ours {ours:.4f}  autocast {ac:.4f}")


def c4_shape(n_mol=512, layers=6, hops=4):
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200 import synthetic as S
    from oracle.fixtures import FEATURE_SIZES
    cfg = dict(hidden_dim=512, num_shells=hops, num_message_passing_layers=layers)
    T = 12
    batch = S.make_batch(4321, n_mol, hops, "qm9", T)
    P = det_state(gnn_shapes(cfg, T), 21)
    ob = dict(atom_features_map=batch.atom_features_map, multi_hop_edge_indices=batch.multi_hop_edge_indices,
              batch_indices=batch.batch_indices, total_charges=batch.total_charges,
              final_tetrahedral_chiral_tensor=batch.final_tetrahedral_chiral_tensor, final_cis_tensor=batch.final_cis_tensor,
              final_trans_tensor=batch.final_trans_tensor)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    ro, _, _, _ = MP.gnn_forward(Pr, cfg, ob)
    up = torch.sign(ro.detach() - batch.targets) / ro.shape[0]
    (ro * up).sum().backward()
    Pa = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ao, _, _, _ = MP.gnn_forward(Pa, cfg, ob)
    (ao.float() * up).sum().backward()
    res = {}
    for mode in ("bf16", "fp32"):
        model = ax.GNN(FEATURE_SIZES, 512, T, num_shells=hops, num_message_passing_layers=layers, task_type="multitask",
                       ffn_dropout=0.0, shell_conv_dropout=0.0)
        model.load_state_dict(P, strict=True)
        model.cuda().train()
        model.compute_dtype = torch.bfloat16 if mode == "bf16" else torch.float32
        bd = batch.to("cuda")
        out, _, _ = model(bd.atom_features_map, bd.multi_hop_edge_indices, bd.batch_indices, bd.total_charges,
                          bd.final_tetrahedral_chiral_tensor, bd.final_cis_tensor, bd.final_trans_tensor, graph_index=bd.graph_index)
        (out * up.cuda()).sum().backward()
        res[mode] = ({k: p.grad.double().cpu() for k, p in model.named_parameters() if p.grad is not None}, out.detach().cpu())
    print(f"== c4 shape L={layers} H={hops} B={n_mol}: out ours {rel(res['bf16'][1].numpy(), ro.detach().numpy()):.4f} "
          f"autocast {rel(ao.detach().float().numpy(), ro.detach().numpy()):.4f} fp32-cuda {rel(res['fp32'][1].numpy(), ro.detach().numpy()):.2e}")
    for k in Pr:
        if Pr[k].grad is None or k not in res["bf16"][0] or ("attention_weights" in k and k.endswith("bias")):
            continue
        ref = Pr[k].grad.double()
        s = float(ref.abs().max())
        if s == 0:
            continue
        print(f"   {k:58s} ours {float((res['bf16'][0][k] - ref).abs().max()) / s:.4f} autocast "
              f"{float((Pa[k].grad.double() - ref).abs().max()) / s:.4f} fp32-cuda {float((res['fp32'][0][k] - ref).abs().max()) / s:.1e} "
              f"| norm ours {float(res['bf16'][0][k].norm()) / float(ref.norm()):.4f} autocast {float(Pa[k].grad.double().norm()) / float(ref.norm()):.4f}")


if "--c4" in sys.argv:
    c4_shape()
